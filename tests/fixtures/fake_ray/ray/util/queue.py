import queue as _queue

Empty = _queue.Empty
Full = _queue.Full


class Queue(object):
    """``ray.util.queue.Queue``: an actor-backed queue in Ray, a plain
    thread-safe one here (same put / get / empty / qsize surface)."""

    def __init__(self, maxsize=0, actor_options=None):
        self._q = _queue.Queue(maxsize)

    def put(self, item, block=True, timeout=None):
        self._q.put(item, block, timeout)

    def get(self, block=True, timeout=None):
        return self._q.get(block, timeout)

    def empty(self):
        return self._q.empty()

    def qsize(self):
        return self._q.qsize()

    def shutdown(self, force=False):
        pass
