"""``kubernetes_asyncio.config`` of the stand-in: nothing to authenticate
against. ``FAKE_K8S_STATE`` (a JSON file with ``nodes`` / ``jobs`` / ``pods``
lists) seeds the in-memory API server of a process started from a test."""
import json
import os

from . import client


def _seed():
    path = os.environ.get("FAKE_K8S_STATE")
    if not path or client.STATE["nodes"] or client.STATE["jobs"]:
        return
    with open(path) as f:
        state = json.load(f)
    for node in state.get("nodes", []):
        client.STATE["nodes"][node["metadata"]["name"]] = node
    for job in state.get("jobs", []):
        key = (job["metadata"]["namespace"], job["metadata"]["name"])
        client.STATE["jobs"][key] = job
        client.STATE["events"].append(("job", "ADDED", job))
    for pod in state.get("pods", []):
        key = (pod["metadata"]["namespace"], pod["metadata"]["name"])
        client.STATE["pods"][key] = pod
        client.STATE["events"].append(("pod", "ADDED", pod))


LOADED = []


def load_incluster_config():
    LOADED.append("incluster")
    marker = os.environ.get("FAKE_K8S_CONFIG_MARKER")
    if marker:
        open(marker, "w").write("incluster")
    _seed()


async def load_kube_config(*args, **kwargs):
    _seed()
