"""Worker-side pieces for Ray-on-AWS (reference: ``aws/worker.py``).

:func:`poll_spot_termination` is plain Python (testable against a fake
metadata endpoint); :func:`worker_environment` builds the ``ADAPTDL_*``
environment of a worker; the ``ray.remote`` wrappers are created on demand by
:func:`remote_functions`.
"""

import importlib.util
import logging
import os
import shutil
import sys
import time
from pathlib import Path

from adaptdl_b200 import __version__
from adaptdl_b200.ray.aws.utils import (Status, checkpoint_obj_to_dir,
                                        serialize_checkpoint)

LOG = logging.getLogger(__name__)
SPOT_ENDPOINT = "169.254.169.254"
MASTER_PORT_BASE = 47000


def poll_spot_termination(endpoint=SPOT_ENDPOINT, timeout=None, period=5.0,
                          request_timeout=0.5):
    """Block until the EC2 spot *instance-action* notice (2-minute warning)
    appears; returns the action (``"terminate"`` / ``"stop"``), or ``None``
    after ``timeout`` seconds."""
    import requests
    url = "http://{}/latest/meta-data/spot/instance-action".format(endpoint)
    start = time.time()
    while True:
        try:
            resp = requests.get(url, timeout=request_timeout)
            if 200 <= resp.status_code < 300:
                action = resp.json().get("action")
                if action in ("terminate", "stop"):
                    return action
            elif resp.status_code != 404:
                raise RuntimeError("spot interruption endpoint not "
                                   "responding ({})".format(resp.status_code))
        except requests.RequestException as exc:
            LOG.debug("spot endpoint: %s", exc)
        if timeout is not None and time.time() - start > timeout:
            return None
        time.sleep(period)


def worker_environment(job_key, job_uid, rank, replicas, num_restarts,
                       supervisor_url, offset=0, base_dir="/tmp"):
    suffix = "{}-{}".format(job_uid, rank)
    return {
        "ADAPTDL_MASTER_PORT": str(MASTER_PORT_BASE + num_restarts + offset),
        "ADAPTDL_REPLICA_RANK": str(rank),
        "ADAPTDL_NUM_REPLICAS": str(replicas),
        "ADAPTDL_SUPERVISOR_URL": supervisor_url,
        "ADAPTDL_JOB_ID": job_key,
        "ADAPTDL_NUM_RESTARTS": str(num_restarts),
        "ADAPTDL_SCHED_VERSION": __version__,
        "ADAPTDL_CHECKPOINT_PATH": os.path.join(
            base_dir, "checkpoint-" + suffix),
        "ADAPTDL_SHARE_PATH": os.path.join(base_dir, "share-" + suffix),
    }


def _forget_previous_generation():
    """Ray reuses idle worker processes: a task of the next generation can
    land in the process that ran the previous one to its checkpoint (the
    script left through ``SystemExit``, the process lived on). The
    framework's process-global state -- control-plane connection, torch
    process group, state registry, epoch / loader / metrics singletons, exit
    flag -- belongs to that finished generation and must go before the
    script starts again."""
    if "adaptdl_b200.torch" not in sys.modules and \
            "adaptdl_b200.collective" not in sys.modules:
        return
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        LOG.warning("could not destroy the previous process group",
                    exc_info=True)
    from adaptdl_b200 import collective
    if collective.is_initialized():
        try:
            collective.teardown()
        except Exception:  # noqa: BLE001
            pass
    from adaptdl_b200.utils.testing import reset_global_state
    reset_global_state()


def run_script(path, argv, environment, checkpoint=None):
    """Execute the user's training script as ``__main__`` inside this
    process. Returns ``(Status, checkpoint_obj|None)``; a ``SystemExit``
    (rescale request: the trainer checkpointed and exited with 143) yields
    the serialised checkpoint directory."""
    _forget_previous_generation()
    os.environ.update(environment)
    ckpt_dir = environment["ADAPTDL_CHECKPOINT_PATH"]
    shutil.rmtree(ckpt_dir, ignore_errors=True)
    os.makedirs(ckpt_dir)
    os.makedirs(environment["ADAPTDL_SHARE_PATH"], exist_ok=True)
    if checkpoint:
        checkpoint_obj_to_dir(ckpt_dir, checkpoint)
    sys.argv = [Path(path).name] + list(argv or [])
    spec = importlib.util.spec_from_file_location("__main__", path)
    module = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(module)
    except SystemExit as exc:
        if exc.code in (0, None):
            return Status.SUCCEEDED, None
        return Status.RUNNING, serialize_checkpoint(ckpt_dir)
    except Exception:  # noqa: BLE001
        LOG.exception("worker failed")
        return Status.FAILED, None
    return Status.SUCCEEDED, None


def remote_functions():
    """``(listen_for_spot_termination, run_adaptdl)`` as Ray remote
    functions."""
    from adaptdl_b200.ray import require_ray
    ray = require_ray()

    @ray.remote(num_cpus=0.1, max_retries=0)
    def listen_for_spot_termination(timeout=None):
        mock = os.environ.get("MOCK", "False").lower() == "true"
        ip = ray.util.get_node_ip_address()
        endpoint = "{}:8234".format(ip) if mock else SPOT_ENDPOINT
        return ip if poll_spot_termination(endpoint, timeout) else None

    @ray.remote(max_retries=0)
    def run_adaptdl(job_key, job_uid, rank, replicas, num_restarts,
                    checkpoint=None, offset=0, path="", argv=None):
        controller = ray.get_actor("AdaptDLController")
        url = ray.get(controller.get_url.remote())
        environment = worker_environment(job_key, job_uid, rank, replicas,
                                         int(num_restarts), url, offset)
        controller.register_worker.remote(rank,
                                          ray.util.get_node_ip_address())
        status, ckpt = run_script(path, argv, environment, checkpoint)
        if status is Status.RUNNING and rank == 0:
            ray.get(controller.register_checkpoint.remote(ckpt))
        elif status is not Status.RUNNING and (rank == 0 or
                                               status is Status.FAILED):
            controller.register_status.remote(status.value)
        return status.value
    return listen_for_spot_termination, run_adaptdl
