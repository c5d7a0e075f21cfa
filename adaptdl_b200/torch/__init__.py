"""PyTorch trainer API (the reference's ``adaptdl.torch``).

    import adaptdl_b200.torch as adl

    adl.init_process_group("nccl")
    model = adl.AdaptiveDataParallel(model, optimizer, lr_scheduler)
    loader = adl.AdaptiveDataLoader(dataset, batch_size=128, shuffle=True)
    loader.autoscale_batch_size(4096, local_bsz_bounds=(32, 1024))
    for epoch in adl.remaining_epochs_until(30):
        for x, y in loader:
            optimizer.zero_grad(); loss_fn(model(x), y).backward()
            optimizer.step()
"""

import logging
import os

# First of all: the SIGTERM / SIGINT handlers. Importing torch takes seconds,
# and a replica that is preempted during that window would die of the signal
# instead of leaving through the checkpoint-and-exit-143 path.
from adaptdl_b200 import _signal  # noqa: F401  (isort: skip)

import torch.distributed

from adaptdl_b200.utils import rescale_trace as _rescale_trace

_rescale_trace.mark("interpreter_up")        # python + torch are loaded

from adaptdl_b200 import __version__, collective, env
from adaptdl_b200.utils import parse_version, pick_unused_port
from .epoch import current_epoch, finished_epochs, remaining_epochs_until
from .data import (current_dataloader, AdaptiveDataLoader, ElasticSampler,
                   DevicePrefetcher)
from .parallel import AdaptiveDataParallel, mixed_precision_params
from .accumulator import Accumulator
from adaptdl_b200.parallel.graph import GraphedTrainStep

LOG = logging.getLogger(__name__)

__all__ = [
    "init_process_group",
    "current_epoch",
    "finished_epochs",
    "remaining_epochs_until",
    "current_dataloader",
    "AdaptiveDataLoader",
    "ElasticSampler",
    "AdaptiveDataParallel",
    "Accumulator",
    "DevicePrefetcher",
    "GraphedTrainStep",
    "mixed_precision_params",
]


def version_check(version):
    """True for a real semantic version (``0.0.0`` means a dev build)."""
    return parse_version(version) is not None and version != "0.0.0"


def _adopt_torchrun_env():
    """Launched by ``torchrun``/``torch.distributed.run`` without ADAPTDL_*
    variables: map RANK / WORLD_SIZE / MASTER_ADDR onto them."""
    if "ADAPTDL_NUM_REPLICAS" in os.environ or "WORLD_SIZE" not in os.environ:
        return
    os.environ["ADAPTDL_NUM_REPLICAS"] = os.environ["WORLD_SIZE"]
    os.environ["ADAPTDL_REPLICA_RANK"] = os.environ.get("RANK", "0")
    os.environ.setdefault("ADAPTDL_NUM_NODES", "1")
    if "MASTER_ADDR" in os.environ:
        os.environ.setdefault("ADAPTDL_MASTER_ADDR",
                              os.environ["MASTER_ADDR"])
    if "MASTER_PORT" in os.environ:
        # the control plane takes the port next to torch's store
        os.environ.setdefault("ADAPTDL_MASTER_PORT",
                              str(int(os.environ["MASTER_PORT"]) + 1))


def _discover_master(url):
    """Long-poll the supervisor until every replica of this restart
    generation has an address; returns rank 0's."""
    import requests
    key, group = env.job_id(), env.num_restarts()
    while True:
        response = requests.get(url="{}/discover/{}/{}".format(url, key,
                                                               group))
        if response.status_code != 408:     # 408 = not all pods up yet
            break
    response.raise_for_status()
    master_addr = response.json()[0]
    sched_version = env.adaptdl_sched_version()
    if version_check(sched_version) and version_check(__version__):
        if parse_version(__version__)[0] != parse_version(sched_version)[0]:
            raise Exception("adaptdl_b200 version {} is incompatible with "
                            "scheduler version {}".format(__version__,
                                                          sched_version))
    return master_addr


def init_process_group(backend, init_method=None, world_size=None,
                       rank=None):
    """Initialise the control plane (:mod:`adaptdl_b200.collective`) and the
    default ``torch.distributed`` process group.

    Arguments:
        backend: ``"nccl"`` for multi-GPU training, else ``"gloo"``.
        init_method (str, optional): ``tcp://host:port`` of rank 0.
        world_size, rank (int, optional): default to ``ADAPTDL_NUM_REPLICAS``
            / ``ADAPTDL_REPLICA_RANK``.

    Rank 0's address comes from ``init_method``, else from the supervisor
    (``ADAPTDL_SUPERVISOR_URL``), else from ``ADAPTDL_MASTER_ADDR``.
    """
    _adopt_torchrun_env()
    if env.from_ray():
        # Ray Tune hands the rendezvous over explicitly; mirror it into the
        # environment the rest of the library reads
        if None in (init_method, world_size, rank):
            raise ValueError("under Ray, init_method, world_size and rank "
                             "must all be given")
        from adaptdl_b200.ray.utils import unique_nodes_pg
        os.environ.update(ADAPTDL_NUM_REPLICAS=str(world_size),
                          ADAPTDL_REPLICA_RANK=str(rank),
                          ADAPTDL_NUM_NODES=str(unique_nodes_pg()))
    if rank is None:
        rank = env.replica_rank()
    if world_size is None:
        world_size = env.num_replicas()

    # where rank 0 listens: explicit URL > supervisor discovery > environment
    master_port = env.master_port()
    supervisor = env.supervisor_url()
    if init_method is not None:
        scheme, _, location = init_method.partition("://")
        master_addr, _, port = location.rpartition(":")
        master_port = int(port)
    elif supervisor:
        master_addr = _discover_master(supervisor)
    else:
        master_addr = env.master_addr()

    collective.initialize(master_addr, master_port, rank, world_size)

    # rank 0 picks a free port for torch's store and tells everyone
    torch_port = collective.broadcast(pick_unused_port())
    torch_addr = "127.0.0.1" if master_addr in ("0.0.0.0", "") \
        else master_addr
    store_url = "tcp://{}:{}?rank={}&world_size={}".format(
        torch_addr, torch_port, rank, world_size)
    LOG.info("Initializing torch.distributed using %s", store_url)
    if str(backend).startswith("nccl") and torch.cuda.is_available():
        torch.cuda.set_device(env.local_rank() % torch.cuda.device_count())
    # Under torchrun the elastic agent hosts the rendezvous store and sets
    # TORCHELASTIC_USE_AGENT_STORE, which makes rank 0 NOT start a server for
    # a fresh tcp:// URL (every rank would wait forever): we bring our own
    # store (port agreed over the control plane), so opt out of the agent's.
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    torch.distributed.init_process_group(backend, store_url)
    LOG.info("torch.distributed initialized")
    from adaptdl_b200.utils import rescale_trace
    rescale_trace.mark("process_group_ready")
