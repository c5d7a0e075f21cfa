"""Symmetric (peer-mapped) device memory for one box of NVLink/NVSwitch GPUs.

A :class:`SymmetricRegion` is one allocation of identical size on every rank
whose physical pages are mapped into *every* rank's address space, so a kernel
on rank r can ``ld.global``/``st.global`` rank p's copy directly over NVLink.

Two providers:

* ``native`` (default): our own runtime (``csrc/adl_symm.cpp``) -- CUDA VMM
  allocations exported as POSIX fds, exchanged between the rank processes
  over unix datagram sockets with ``SCM_RIGHTS``, imported and mapped by each
  peer. Rebuilt from scratch at every elastic restart, at whatever world size
  the new generation has.
* ``torch``: ``torch.distributed._symmetric_memory`` as a bring-up fallback.
"""

import array
import logging
import os
import socket
import struct
import uuid

import torch
import torch.distributed as dist

from adaptdl_b200 import _native

LOG = logging.getLogger(__name__)


class _DeviceMemory(object):
    """Expose a raw device pointer through ``__cuda_array_interface__`` so
    torch can wrap it without copying."""

    def __init__(self, ptr, nbytes, owner=None):
        self.__cuda_array_interface__ = {
            "shape": (int(nbytes),), "typestr": "|u1",
            "data": (int(ptr), False), "version": 2, "strides": None,
        }
        self._owner = owner


def _all_gather_object(obj, group):
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def send_fd(sock, address, payload, fd):
    """One datagram carrying ``payload`` and the file descriptor ``fd``
    (SCM_RIGHTS). (``socket.send_fds`` drops its ``address`` argument on
    CPython <= 3.12, hence the explicit ``sendmsg``.)"""
    sock.sendmsg([payload],
                 [(socket.SOL_SOCKET, socket.SCM_RIGHTS,
                   array.array("i", [fd]))], 0, address)


def recv_fd(sock, bufsize=64):
    msg, fds, _, _ = socket.recv_fds(sock, bufsize, 4)
    if len(fds) != 1:
        for extra in fds:
            os.close(extra)
        raise RuntimeError("expected exactly one file descriptor")
    return msg, fds[0]


class FdExchange(object):
    """All-to-all exchange of one file descriptor per rank per round over
    unix datagram sockets in the abstract namespace."""

    def __init__(self, rank, world, gather_addresses):
        self.rank, self.world = rank, world
        self.sock = socket.socket(socket.AF_UNIX, socket.SOCK_DGRAM)
        self.addr = "\0adl-b200-{}-{}".format(uuid.uuid4().hex, rank)
        self.sock.bind(self.addr)
        self.sock.settimeout(120.0)
        # gathering the addresses doubles as "everyone has bound"
        self.addrs = gather_addresses(self.addr)
        self._serial = 0
        self._early = {}

    def exchange(self, fd):
        """Send ``fd`` to every peer; returns ``{rank: fd}`` of the peers'
        descriptors for the same round (caller closes them)."""
        serial = self._serial
        self._serial += 1
        header = struct.pack("!II", serial, self.rank)
        for peer, addr in enumerate(self.addrs):
            if peer != self.rank:
                send_fd(self.sock, addr, header, fd)
        got = {}
        while len(got) < self.world - 1:
            key = next((k for k in self._early if k[0] == serial), None)
            if key is not None:
                got[key[1]] = self._early.pop(key)
                continue
            msg, peer_fd = recv_fd(self.sock)
            got_serial, src = struct.unpack("!II", msg)
            if got_serial != serial:          # a faster peer is a round ahead
                self._early[(got_serial, src)] = peer_fd
            else:
                got[src] = peer_fd
        return got

    def close(self):
        try:
            self.sock.close()
        except OSError:
            pass


class SymmetricRegion(object):
    """``tensor``: this rank's bytes (uint8, on ``device``); ``ptrs[p]``:
    address of rank p's bytes in *this* process."""

    def __init__(self, tensor, ptrs, nbytes, provider, keepalive=None,
                 mc_ptr=0):
        self.tensor = tensor
        self.ptrs = list(ptrs)
        self.nbytes = nbytes
        self.provider = provider
        self.mc_ptr = int(mc_ptr or 0)   # NVLS multicast address (0 = none)
        self._keepalive = keepalive

    def carve(self, offset, nbytes, dtype=torch.uint8):
        """A typed view of ``[offset, offset+nbytes)`` plus its per-rank
        addresses."""
        view = self.tensor[offset:offset + nbytes].view(dtype)
        return view, [p + offset for p in self.ptrs]


class _NativeProvider(object):
    name = "native"

    def __init__(self, group, device):
        self.group = group
        self.device = device
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.lib = _native.load()
        if self.lib.adl_symm_init() != 0:
            raise RuntimeError("symmetric memory runtime: "
                               + self.lib.adl_symm_last_error().decode())
        if not self.lib.adl_topo_vmm_fd_supported(device.index):
            raise RuntimeError("device lacks POSIX-fd shareable VMM handles")
        self.fds = FdExchange(self.rank, self.world,
                              lambda addr: _all_gather_object(addr, group))
        self._mapped = []

    def _check(self, code, what):
        if code != 0:
            raise RuntimeError("{}: {}".format(
                what, self.lib.adl_symm_last_error().decode()))

    def allocate(self, nbytes):
        import ctypes
        lib, dev = self.lib, self.device.index
        size = ctypes.c_size_t()
        self._check(lib.adl_symm_round_size(dev, nbytes, ctypes.byref(size)),
                    "round_size")
        size = size.value
        want_mc = os.environ.get("ADAPTDL_B200_NVLS", "auto") != "0" \
            and bool(lib.adl_topo_multicast_supported(dev))
        if want_mc:              # multicast binding has its own granularity
            mc_size = ctypes.c_size_t()
            if lib.adl_mc_round_size(self.world, size,
                                     ctypes.byref(mc_size)) == 0:
                size = max(size, mc_size.value)
            else:
                want_mc = False
        handle, fd = ctypes.c_ulonglong(), ctypes.c_int()
        self._check(lib.adl_symm_create(dev, size, ctypes.byref(handle),
                                        ctypes.byref(fd)), "create")
        handles = {self.rank: handle.value}
        try:
            for src, peer_fd in self.fds.exchange(fd.value).items():
                imported = ctypes.c_ulonglong()
                try:
                    self._check(lib.adl_symm_import(
                        peer_fd, ctypes.byref(imported)), "import")
                finally:
                    os.close(peer_fd)
                handles[src] = imported.value
        finally:
            os.close(fd.value)
        ptrs = []
        for peer in range(self.world):
            ptr = ctypes.c_ulonglong()
            self._check(lib.adl_symm_map(handles[peer], size, dev,
                                         ctypes.byref(ptr)), "map")
            ptrs.append(ptr.value)
            self._mapped.append((ptr.value, size, handles[peer]))
        holder = _DeviceMemory(ptrs[self.rank], size, owner=self)
        tensor = torch.as_tensor(holder, device=self.device)
        tensor.zero_()
        torch.cuda.synchronize(self.device)
        # nobody may touch a peer's bytes before that peer zeroed them
        flags = _all_gather_object(bool(want_mc), self.group)
        mc_ptr = 0
        if all(flags):
            mc_ptr = self._bind_multicast(handle.value, size)
        return SymmetricRegion(tensor, ptrs, size, self.name, keepalive=self,
                               mc_ptr=mc_ptr)

    def _bind_multicast(self, mem_handle, size):
        """Create (rank 0) / import the NVLS multicast object, add every
        device, bind this rank's physical memory and map the multicast
        address. Any failure on any rank disables it everywhere."""
        import ctypes
        lib, dev = self.lib, self.device.index
        ok, mc_handle, fd = True, ctypes.c_ulonglong(), ctypes.c_int(-1)
        if self.rank == 0:
            ok = lib.adl_mc_create(self.world, size, ctypes.byref(mc_handle),
                                   ctypes.byref(fd)) == 0
        send_fd_ = fd.value if (self.rank == 0 and ok) \
            else os.open(os.devnull, os.O_RDONLY)
        try:
            got = self.fds.exchange(send_fd_)
        finally:
            os.close(send_fd_)
        for src, peer_fd in got.items():
            if src == 0 and self.rank != 0:
                ok = lib.adl_symm_import(peer_fd,
                                         ctypes.byref(mc_handle)) == 0
            os.close(peer_fd)
        ok = all(_all_gather_object(bool(ok), self.group))
        if ok:
            ok = lib.adl_mc_add_device(mc_handle.value, dev) == 0
        ok = all(_all_gather_object(bool(ok), self.group))
        if ok:
            ok = lib.adl_mc_bind(mc_handle.value, mem_handle, size) == 0
        ok = all(_all_gather_object(bool(ok), self.group))
        ptr = ctypes.c_ulonglong()
        if ok:
            ok = lib.adl_symm_map(mc_handle.value, size, dev,
                                  ctypes.byref(ptr)) == 0
        ok = all(_all_gather_object(bool(ok), self.group))
        if not ok:
            LOG.info("NVLS multicast unavailable (%s); using P2P loads",
                     lib.adl_symm_last_error().decode())
            return 0
        self._mapped.append((ptr.value, size, mc_handle.value))
        return ptr.value

    def close(self):
        for ptr, size, handle in self._mapped:
            self.lib.adl_symm_unmap(ptr, size)
            self.lib.adl_symm_release(handle)
        self._mapped = []
        self.fds.close()


class _TorchProvider(object):
    name = "torch"

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.device = device

    def allocate(self, nbytes):
        nbytes = (nbytes + 511) // 512 * 512
        try:
            self.symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001 - not needed on newer torch
            pass
        t = self.symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        hdl = self.symm_mem.rendezvous(t, group=self.group.group_name)
        t.zero_()
        torch.cuda.synchronize(self.device)
        hdl.barrier()
        mc_ptr = 0
        if os.environ.get("ADAPTDL_B200_NVLS", "auto") != "0":
            try:
                mc_ptr = int(hdl.multicast_ptr or 0)
            except Exception:  # noqa: BLE001
                mc_ptr = 0
        return SymmetricRegion(t, [int(p) for p in hdl.buffer_ptrs], nbytes,
                               self.name, keepalive=(hdl, t), mc_ptr=mc_ptr)

    def close(self):
        pass


class _LocalProvider(object):
    """World size 1: plain device memory."""
    name = "local"

    def __init__(self, device):
        self.device = device

    def allocate(self, nbytes):
        nbytes = (nbytes + 511) // 512 * 512
        t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        return SymmetricRegion(t, [t.data_ptr()], nbytes, self.name)

    def close(self):
        pass


def _native_capable(device):
    """Local probe (no collectives): can this rank use the native runtime?"""
    try:
        lib = _native.load()
        if lib.adl_symm_init() != 0:
            return False, lib.adl_symm_last_error().decode()
        if not lib.adl_topo_vmm_fd_supported(device.index):
            return False, "no POSIX-fd VMM handles"
        return True, ""
    except Exception as exc:  # noqa: BLE001
        return False, str(exc)


def make_provider(group, device, world_size):
    """Provider selection: ``ADAPTDL_B200_SYMM`` = native | torch | auto.
    Every rank takes the same branch (the capability probe is agreed on)."""
    if world_size <= 1:
        return _LocalProvider(device)
    want = os.environ.get("ADAPTDL_B200_SYMM", "auto").lower()
    if want in ("auto", "native"):
        ok, why = _native_capable(device)
        everyone = _all_gather_object((ok, why), group)
        if all(flag for flag, _ in everyone):
            return _NativeProvider(group, device)
        reasons = [w for flag, w in everyone if not flag]
        if want == "native":
            raise RuntimeError("native symmetric memory unavailable: "
                               + "; ".join(reasons))
        LOG.warning("native symmetric memory unavailable (%s); using "
                    "torch.distributed._symmetric_memory", reasons[0])
    return _TorchProvider(group, device)
