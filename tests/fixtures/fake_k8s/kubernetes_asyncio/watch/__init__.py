import asyncio

from .. import client


class Watch(object):
    """``async with Watch() as w: async for event in w.stream(fn, ...)``:
    replays the fixture's event log for the listed kind, then idles."""

    async def __aenter__(self):
        return self

    async def __aexit__(self, *exc):
        return False

    async def stream(self, fn, *args, timeout_seconds=None, **kwargs):
        kind = "job" if "custom_object" in fn.__name__ else "pod"
        seen = 0
        idle = 0
        while idle < 3:
            events = [e for e in client.STATE["events"] if e[0] == kind]
            if seen < len(events):
                idle = 0
                for _, what, obj in events[seen:]:
                    seen += 1
                    yield {"type": what,
                           "object": client._Model(obj) if kind == "pod"
                           else dict(obj)}
            else:
                idle += 1
                await asyncio.sleep(0.01)
