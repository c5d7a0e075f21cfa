"""Control-plane star reducer for small Python objects.

Rank 0 hosts a server thread; every replica (rank 0 included) is a client
over TCP. A collective is: every client sends one framed, pickled object
tagged with a per-client sequence number (the *key*); once the server holds
the objects of all replicas for a key it folds them in rank order with the
reduce function that rank 0 registered for that key and sends the result
back to everyone. Futures may be waited out of order.

Capabilities match the reference's ``adaptdl/adaptdl/reducer.py:30-160``
(allreduce / allreduce_async / broadcast, local mode with port 0, connect
retries); the design differs: length-prefixed frames instead of streaming
``pickle.load``, a fully non-blocking selector-driven server (per-connection
input buffers and output queues: contributions are accepted in any arrival
order and several large asynchronous all-reduces may be in flight without
the send/receive deadlock a blocking server has), an event instead of
sleep-polling for local-mode port discovery, and a real ``close``.

All replicas must invoke collectives in the same order.
"""

import logging
import pickle
import selectors
import socket
import struct
import threading
import time

LOG = logging.getLogger(__name__)

_HDR = struct.Struct("!IQ")          # key (u32), payload length (u64)
_HELLO = 0xFFFFFFFF                  # key of the rank-announcement frame


def default_reduce_fn(a, b):
    a += b
    return a


def _send_frame(sock, key, obj):
    payload = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(_HDR.pack(key, len(payload)) + payload)


def _recv_exact(sock, n):
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        r = sock.recv_into(view[got:], n - got)
        if r == 0:
            raise ConnectionError("control-plane peer closed the connection")
        got += r
    return bytes(buf)


def _recv_frame(sock):
    key, length = _HDR.unpack(_recv_exact(sock, _HDR.size))
    return key, pickle.loads(_recv_exact(sock, length))


class Future(object):
    """Handle to the result of an asynchronous all-reduce."""

    _UNSET = object()

    def __init__(self, reducer, key):
        self._reducer = reducer
        self._key = key
        self._value = Future._UNSET

    def result(self):
        if self._value is Future._UNSET:
            self._value = self._reducer._wait_for(self._key)
        return self._value


class _Server(threading.Thread):
    """Rank-0 server: gathers one object per replica per key, reduces,
    replies."""

    def __init__(self, port, replicas, reduce_fns, reduce_fns_lock):
        super().__init__(daemon=True, name="adaptdl-b200-reducer")
        self._replicas = replicas
        self._reduce_fns = reduce_fns
        self._lock = reduce_fns_lock
        self._listener = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._listener.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._listener.bind(("0.0.0.0", port))
        self._listener.listen(max(replicas, 8))
        self.port = self._listener.getsockname()[1]
        self._stop_event = threading.Event()
        self.error = None

    def stop(self):
        self._stop_event.set()
        try:
            self._listener.close()
        except OSError:
            pass

    def _accept_all(self):
        clients = [None] * self._replicas
        while any(c is None for c in clients):
            conn, _ = self._listener.accept()
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            key, rank = _recv_frame(conn)
            if key != _HELLO or not (0 <= rank < self._replicas) \
                    or clients[rank] is not None:
                conn.close()
                raise RuntimeError("bad control-plane handshake")
            clients[rank] = conn
        return clients

    def run(self):
        conns = []
        try:
            conns = [_Conn(rank, sock)
                     for rank, sock in enumerate(self._accept_all())]
            sel = selectors.DefaultSelector()
            for conn in conns:
                conn.sock.setblocking(False)
                sel.register(conn.sock, selectors.EVENT_READ, conn)
            self._loop(sel, conns)
        except Exception as exc:  # noqa: BLE001 - surfaced to clients
            if not self._stop_event.is_set():
                self.error = exc
                LOG.exception("control-plane reducer server failed")
        finally:
            try:
                self._listener.close()
            except OSError:
                pass
            if self.error is not None:
                # fail fast: everybody blocked on a result gets a
                # ConnectionError instead of waiting for a replica that died
                for conn in conns:
                    try:
                        conn.sock.shutdown(socket.SHUT_RDWR)
                    except OSError:
                        pass

    def _loop(self, sel, conns):
        """Event loop. Nothing in here blocks on one client: frames are
        parsed out of per-connection input buffers as bytes arrive and
        replies leave through per-connection output queues, so a replica
        that is still busy sending (several large asynchronous all-reduces
        in flight) cannot stall the replies the others -- or itself -- are
        waiting for (the reference's blocking server deadlocks there)."""
        pending = {}                         # key -> {rank: obj}
        dead = set()                         # ranks that hung up
        alive = self._replicas
        held_for_rank0 = []                  # see _finish

        def doomed(slot):
            """A reduction some departed replica never contributed to can
            never complete."""
            return any(rank not in slot for rank in dead)

        def want_write(conn):
            events = selectors.EVENT_READ if not conn.closed else 0
            if conn.out:
                events |= selectors.EVENT_WRITE
            if events:
                sel.modify(conn.sock, events, conn)

        drain_deadline = None
        while alive:
            if self._stop_event.is_set():
                # orderly shutdown: flush what is already queued, briefly
                if drain_deadline is None:
                    drain_deadline = time.time() + 5.0
                if not any(c.out for c in conns) and not held_for_rank0 \
                        or time.time() > drain_deadline:
                    break
            for skey, events in sel.select(timeout=0.2):
                conn = skey.data
                if events & selectors.EVENT_WRITE:
                    conn.flush()
                    want_write(conn)
                if not events & selectors.EVENT_READ:
                    continue
                try:
                    frames = conn.read_frames()
                except (ConnectionError, OSError):
                    conn.closed = True
                    if conn.out:
                        want_write(conn)
                    else:
                        sel.unregister(conn.sock)
                    alive -= 1
                    dead.add(conn.rank)
                    if any(doomed(slot) for slot in pending.values()):
                        raise ConnectionError(
                            "replica {} left in the middle of a "
                            "collective".format(conn.rank))
                    continue
                for key, obj in frames:
                    slot = pending.setdefault(key, {})
                    slot[conn.rank] = obj
                    if len(slot) == self._replicas:
                        del pending[key]
                        frame = self._finish(key, slot)
                        for other in conns[1:]:
                            if not other.closed:
                                other.out.append(memoryview(frame))
                                want_write(other)
                        held_for_rank0.append(frame)
                    elif doomed(slot):
                        raise ConnectionError(
                            "replica(s) {} are gone: the collective cannot "
                            "complete".format(sorted(dead)))
            # Rank 0 is answered last: it hosts this thread, and once it has
            # its result it may run ahead and even exit; by then every other
            # replica's reply must have left.
            if held_for_rank0 and not any(c.out for c in conns[1:]):
                if not conns[0].closed:
                    conns[0].out.extend(memoryview(f)
                                        for f in held_for_rank0)
                    want_write(conns[0])
                held_for_rank0 = []

    def _finish(self, key, slot):
        with self._lock:
            reduce_fn = self._reduce_fns.pop(key)
        result = slot[0]
        for rank in range(1, self._replicas):
            result = reduce_fn(result, slot[rank])
        payload = pickle.dumps(result, protocol=pickle.HIGHEST_PROTOCOL)
        return _HDR.pack(key, len(payload)) + payload


class _Conn(object):
    """Server side of one replica's connection (non-blocking)."""

    CHUNK = 1 << 20

    def __init__(self, rank, sock):
        self.rank = rank
        self.sock = sock
        self.inbuf = bytearray()
        self.out = []               # memoryviews still to be sent, in order
        self.closed = False

    def read_frames(self):
        """Drain the socket; returns the complete ``(key, obj)`` frames."""
        while True:
            try:
                chunk = self.sock.recv(self.CHUNK)
            except (BlockingIOError, InterruptedError):
                break
            if not chunk:
                raise ConnectionError("peer closed the connection")
            self.inbuf += chunk
            if len(chunk) < self.CHUNK:
                break
        frames = []
        view = self.inbuf
        offset = 0
        while len(view) - offset >= _HDR.size:
            key, length = _HDR.unpack_from(view, offset)
            end = offset + _HDR.size + length
            if len(view) < end:
                break
            frames.append((key, pickle.loads(
                bytes(view[offset + _HDR.size:end]))))
            offset = end
        if offset:
            del self.inbuf[:offset]
        return frames

    def flush(self):
        while self.out:
            head = self.out[0]
            try:
                sent = self.sock.send(head)
            except (BlockingIOError, InterruptedError):
                return
            except OSError:
                self.out = []       # the peer is gone
                return
            if sent < len(head):
                self.out[0] = head[sent:]
                return
            self.out.pop(0)


class Reducer(object):
    """Asynchronous (all)reduce of Python objects over a TCP star.

    Arguments:
        rank, replicas: this replica's rank and the job size.
        root_host, root_port: where rank 0 listens. ``root_port == 0`` is
            *local mode*: rank 0 binds an ephemeral port (only meaningful
            when every replica lives in this process or the port is
            communicated out of band).
    """

    CONNECT_RETRIES = 25

    def __init__(self, rank, replicas, root_host, root_port):
        self._rank = rank
        self._replicas = replicas
        self._next_key = 0
        self._results = {}
        self._server = None
        self._root_port = root_port
        if rank == 0:
            self._reduce_fns = {}
            self._fn_lock = threading.Lock()
            self._server = _Server(root_port, replicas, self._reduce_fns,
                                   self._fn_lock)
            self._root_port = self._server.port
            self._server.start()
            root_host = "127.0.0.1" if root_host in ("0.0.0.0", "") \
                else root_host
        elif root_host in ("0.0.0.0", ""):
            root_host = "127.0.0.1"
        self._sock = self._connect(root_host, self._root_port)
        _send_frame(self._sock, _HELLO, rank)

    @property
    def root_port(self):
        return self._root_port

    def _connect(self, host, port):
        delay = 0.05
        last = None
        for attempt in range(self.CONNECT_RETRIES + 1):
            sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            try:
                if port == 0:
                    raise ConnectionRefusedError("root port not known yet")
                sock.connect((host, port))
                sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                return sock
            except (ConnectionRefusedError, socket.gaierror, OSError) as e:
                last = e
                sock.close()
                LOG.debug("rank %d: root %s:%s not ready (%s), retrying",
                          self._rank, host, port, e)
                time.sleep(delay)
                delay = min(delay * 2, 5.0)
        raise ConnectionError("rank {} could not reach the root reducer at "
                              "{}:{}: {}".format(self._rank, host, port, last))

    # -- collectives -------------------------------------------------------

    def allreduce_async(self, obj, reduce_fn=default_reduce_fn):
        key = self._next_key
        self._next_key = (self._next_key + 1) % _HELLO
        if self._rank == 0:
            with self._fn_lock:
                self._reduce_fns[key] = reduce_fn
        _send_frame(self._sock, key, obj)
        return Future(self, key)

    def allreduce(self, obj, reduce_fn=default_reduce_fn):
        return self.allreduce_async(obj, reduce_fn).result()

    def broadcast(self, obj):
        """Rank 0's value wins (all-reduce with the left projection)."""
        return self.allreduce(obj, lambda x, y: x)

    def _wait_for(self, key):
        while key not in self._results:
            try:
                got_key, value = _recv_frame(self._sock)
            except Exception as exc:
                LOG.error("rank %d lost the control plane: %s",
                          self._rank, exc)
                raise
            self._results[got_key] = value
        return self._results.pop(key)

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass
        if self._server is not None:
            self._server.stop()
