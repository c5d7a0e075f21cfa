"""AdaptiveDataParallel: elastic data parallelism with adaptive batch size.

Public behaviour follows the reference's
``adaptdl/adaptdl/torch/parallel.py:39-239`` (constructor signature,
``forward`` accumulation control, ``gain``, ``to_tensorboard``, checkpoint
format) but the implementation is not a ``DistributedDataParallel`` subclass:
gradients live in flat, NVLink-mapped arenas owned by a
:class:`~adaptdl_b200.parallel.reducer_base.GradReducer`, and one fused
sm_100a kernel per bucket performs the all-reduce, the ``1/(N*accum)``
scaling and the gradient-noise-scale statistics while backward is still
running. On CPU (gloo) the same control flow runs on stock torch ops.
"""

import contextlib
import logging
import os
import warnings
from typing import Optional

import numpy as np
import torch
import torch.distributed

from adaptdl_b200 import checkpoint, env
from adaptdl_b200.parallel import make_reducer
from adaptdl_b200.torch import _metrics
from adaptdl_b200.torch.data import current_dataloader
from adaptdl_b200.torch.gradient_noise_scale import (
    AdamGradientNoiseScale, GradientNoiseScale)
from adaptdl_b200.torch.scaling_rules import (
    AdaScale, AdamScale, ScalingRuleBase)
from adaptdl_b200.utils.trace import traced

LOG = logging.getLogger(__name__)

_IGNORED_DDP_KWARGS = (
    "device_ids", "output_device", "dim", "find_unused_parameters",
    "check_reduction", "gradient_as_bucket_view", "static_graph",
    "delay_all_reduce_named_params", "param_to_hook_all_reduce",
    "mixed_precision", "device_mesh", "init_sync", "skip_all_reduce_unused_params")


class AdaptiveDataParallel(torch.nn.Module):
    """Wrap ``model`` for elastic, adaptive-batch-size data parallelism.

    Saves/restores the model, optimizer, LR scheduler and AMP scaler as part
    of every checkpoint, patches ``optimizer.step`` / ``zero_grad`` with the
    chosen LR scaling rule, and keeps replicas' gradients averaged.

    Arguments:
        model (torch.nn.Module): model to distribute (already on its device).
        optimizer (torch.optim.Optimizer): updates ``model``'s parameters.
        lr_scheduler: optional LR scheduler to checkpoint.
        mp_scaler: optional ``torch.amp.GradScaler`` (AMP loss scaling).
        scaling_rule (ScalingRuleBase): defaults to ``AdamScale`` for
            Adam/AdamW, else ``AdaScale``.
        name (str): unique name, needed only with several instances.
        **kwargs: ``DistributedDataParallel`` keyword arguments are accepted
            for compatibility: ``broadcast_buffers``, ``bucket_cap_mb`` and
            ``process_group`` are honoured, the rest ignored.
            ``reducer`` = ``"auto" | "cuda" | "torch"`` picks the gradient
            reducer.
    """

    def __init__(self, model, optimizer, lr_scheduler=None, mp_scaler=None,
                 scaling_rule: Optional[ScalingRuleBase] = None,
                 name="adaptdl-dataparallel", **kwargs):
        super().__init__()
        self.module = model
        self._key = id(self)
        self.require_backward_grad_sync = True
        self.broadcast_buffers = kwargs.pop("broadcast_buffers", True)
        bucket_cap_mb = kwargs.pop("bucket_cap_mb", None)
        process_group = kwargs.pop("process_group", None)
        backend = kwargs.pop("reducer", "auto")
        fused_step = kwargs.pop("fused_step", None)
        for key in list(kwargs):
            if key in _IGNORED_DDP_KWARGS:
                kwargs.pop(key)
        if kwargs:
            raise TypeError("unexpected arguments: {}".format(sorted(kwargs)))

        if torch.distributed.is_available() and \
                torch.distributed.is_initialized():
            self._world_size = torch.distributed.get_world_size(process_group)
            self._rank = torch.distributed.get_rank(process_group)
        else:
            self._world_size, self._rank = 1, 0

        if not scaling_rule and isinstance(
                optimizer, (torch.optim.Adam, torch.optim.AdamW)):
            self.scaling_rule = AdamScale()
        else:
            self.scaling_rule = scaling_rule or AdaScale()

        self._reducer = make_reducer(
            optimizer.param_groups, self._world_size, self._rank,
            lambda: self.require_backward_grad_sync,
            bucket_cap_mb=bucket_cap_mb, process_group=process_group,
            backend=backend, name=name)
        # Reference quirk kept (App. D1): preconditioned statistics only
        # when an AdamScale instance was passed explicitly.
        gns_cls = AdamGradientNoiseScale if isinstance(scaling_rule, AdamScale) \
            else GradientNoiseScale
        self.gns = gns_cls(self, optimizer, mp_scaler=mp_scaler,
                           num_replicas=self._world_size,
                           reducer=self._reducer)
        self.gns.add_listener(self._on_stats)
        self.gns.add_backward_listener(self._on_backward_end)
        self.scaling_rule.initialize(self, optimizer, patch_optimizer=True)
        self._engine = self._make_engine(optimizer, mp_scaler, fused_step)
        self._warn_if_unmastered_16bit(optimizer)

        self._state = _AdaptiveDataParallelState(
            model, optimizer, lr_scheduler, mp_scaler, name, self._engine)
        self._state.adp = self
        checkpoint.load_state(self._state)
        exact_masters = False
        if self._engine is not None:
            self._engine.adopt_optimizer_state()
            self._engine.push_gns_state(optimizer.state["gns"])
            if self._state.wide_state is not None:
                # exact fp32 masters / optimizer state of 16-bit parameters
                self._engine.load_wide_state(self._state.wide_state)
                self._state.wide_state = None
                exact_masters = True
        self._sync_module_states()
        if self._engine is not None and not exact_masters:
            # fresh start, or a checkpoint written without the engine: the
            # masters are the (broadcast / loaded) 16-bit weights
            self._engine.resync_master()
        from adaptdl_b200.utils import rescale_trace
        rescale_trace.mark("wrapper_ready")

    # ------------------------------------------------------------------

    @property
    def reducer(self):
        return self._reducer

    @property
    def engine(self):
        """The device-resident step engine, or ``None`` (host path)."""
        return self._engine

    def _make_engine(self, optimizer, mp_scaler, fused_step):
        import os
        if fused_step is None:
            fused_step = os.environ.get("ADAPTDL_B200_FUSED", "1") != "0"
        if not fused_step \
                or type(self._reducer).__name__ != "CudaGradReducer":
            return None
        from adaptdl_b200.parallel.engine import DeviceEngine
        try:
            engine = DeviceEngine(
                self._reducer, optimizer, self.scaling_rule,
                optimizer.state["gns"],
                precondition_stats=isinstance(self.gns,
                                              AdamGradientNoiseScale),
                mp_scaler=mp_scaler)
        except ValueError as exc:
            LOG.info("device engine unavailable (%s); using the host path",
                     exc)
            return None
        self.gns.attach_engine(engine)
        if _metrics.device_timer() is None:
            # step / sync durations of the goodput profile come from the
            # device's %globaltimer stamps (reference: host clocks,
            # torch/_metrics.py:43-59)
            from adaptdl_b200.parallel.timer import DeviceStepTimer
            self._step_timer = DeviceStepTimer(self._reducer)
            _metrics.set_device_timer(self._step_timer)
        return engine

    def _warn_if_unmastered_16bit(self, optimizer):
        if self._engine is not None:
            return
        narrow = [p for g in optimizer.param_groups for p in g["params"]
                  if p.dtype in (torch.bfloat16, torch.float16)]
        if narrow:
            LOG.warning(
                "%d parameters are stored in 16 bits but the device engine "
                "(fp32 master weights in the fused optimizer) is not "
                "active: the optimizer updates them in 16-bit precision",
                len(narrow))

    def _sync_engine_ctrl(self):
        engine = self._engine
        if engine is None or not engine.enabled:
            return
        legw_unit = 0.0
        rule = self.scaling_rule
        if hasattr(rule, "_base_warmup_epochs"):
            dataloader = current_dataloader()
            if dataloader is not None:
                legw_unit = (rule._base_warmup_epochs * rule._data_size
                             / dataloader.batch_size)
        engine.sync_ctrl(self.gns.accum_scale, self.gns._smoothing,
                         legw_unit)

    def _module_tensors(self, buffers_only=False):
        tensors = [] if buffers_only else \
            [p.data for p in self.module.parameters()]
        tensors += [b for b in self.module.buffers()
                    if torch.is_tensor(b) and b.numel() > 0]
        return tensors

    def _split_buffers(self):
        """``(eager, deferred)`` buffers for ``broadcast_buffers``.

        DDP (and the reference on top of it) re-broadcasts rank 0's buffers
        before every training forward. For normalisation layers that track
        running statistics the training forward only *updates* those buffers
        (``r <- (1-m) r + m batch``), it never feeds them into the output, and
        rank 0's sequence of values does not depend on the other ranks'. So
        broadcasting rank 0's running statistics right before they are first
        *used* -- the next evaluation-mode / no-grad forward -- leaves every
        rank with exactly the values the per-step broadcast would have given
        it, without two peer barriers and three launches on every step's
        critical path. Buffers of any other module are broadcast every
        forward, as before. ``ADAPTDL_B200_EAGER_BUFFER_BCAST=1`` restores
        the per-forward broadcast for everything."""
        import os
        from torch.nn.modules.batchnorm import _NormBase
        eager, deferred = [], []
        lazy_ok = os.environ.get("ADAPTDL_B200_EAGER_BUFFER_BCAST",
                                 "0") != "1"
        seen = set()
        for mod in self.module.modules():
            is_norm = isinstance(mod, _NormBase) and mod.track_running_stats
            for buf in mod.buffers(recurse=False):
                if not torch.is_tensor(buf) or buf.numel() == 0 \
                        or id(buf) in seen:
                    continue
                seen.add(id(buf))
                (deferred if (is_norm and lazy_ok) else eager).append(buf)
        return eager, deferred

    def _sync_module_states(self):
        """Rank 0's parameters and buffers win (construction and after every
        elastic restart)."""
        if self._world_size > 1:
            self._reducer.broadcast_parameters(self._module_tensors())

    def _pre_forward(self):
        """Host-side step set-up (also run before a CUDA-graph replay)."""
        dataloader = current_dataloader()
        if dataloader is not None and dataloader.training:
            # no gradient synchronisation on accumulation micro-steps
            self.require_backward_grad_sync = dataloader.is_optim_step()
            accum_scale = (dataloader.current_local_bsz
                           * env.num_replicas() / dataloader.batch_size)
            self.gns.set_accum_scale(accum_scale)
        self._sync_engine_ctrl()

    @traced("forward")
    def forward(self, *args, **kwargs):
        self._pre_forward()
        if self.broadcast_buffers and self._world_size > 1:
            training = self.module.training and torch.is_grad_enabled()
            if training and self.require_backward_grad_sync:
                eager, deferred = self._split_buffers()
                if eager:
                    self._reducer.broadcast_parameters(eager)
                self._deferred_buffers_stale = bool(deferred)
            elif not training and self._deferred_buffers_stale:
                self.sync_buffers()
        return self.module(*args, **kwargs)

    _deferred_buffers_stale = False

    def train(self, mode=True):
        # leaving training mode (``net.eval()``, called on every replica):
        # the deferred running statistics are about to be used
        if not mode and self._deferred_buffers_stale:
            self.sync_buffers()
        return super().train(mode)

    def sync_buffers(self):
        """Give every replica rank 0's buffers now (what the next
        evaluation-mode forward does on its own)."""
        if self._world_size > 1:
            buffers = self._module_tensors(buffers_only=True)
            if buffers:
                self._reducer.broadcast_parameters(buffers)
        self._deferred_buffers_stale = False

    @contextlib.contextmanager
    def no_sync(self):
        """DDP-compatible context: backward passes inside only accumulate."""
        old = self.require_backward_grad_sync
        self.require_backward_grad_sync = False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    def _on_backward_end(self, sync):
        dataloader = current_dataloader()
        if dataloader is None:
            raise RuntimeError("backpropagation outside AdaptiveDataLoader")
        dataloader.train()

    def _on_stats(self, stats):
        # invoked when the statistics of a synchronised backward have been
        # folded into the running averages
        if stats.sync_time is not None:
            try:
                _metrics.profile_sync_time(stats.sync_time)
            except AttributeError:
                pass       # not inside a profiled step
        dataloader = current_dataloader()
        if dataloader is None:
            return
        scale = dataloader.current_batch_size / dataloader.batch_size
        self._state.gain = self.gns.gain(scale)
        self._state.lr_factor = float(
            np.average(self.scaling_rule.scale_lr(scale)))
        _metrics.update_progress(self.gns.get_progress())
        if dataloader.max_batch_size and \
                dataloader.max_batch_size > dataloader.batch_size:
            _metrics.update_grad_params(self._key, self.gns.sqr_avg(),
                                        self.gns.var_avg())

    def zero_grad(self, *args, **kwargs):
        warnings.warn("zero_grad has no effect with AdaptiveDataParallel")

    @property
    def gain(self):
        """Current estimate of the AdaScale gain (r_t)."""
        self.gns._flush()
        return self._state.gain

    def to_tensorboard(self, writer, global_step, tag_prefix=""):
        """Write gradient statistics to a TensorBoard ``SummaryWriter``."""
        prefix = tag_prefix.rstrip("/") + "/" if tag_prefix else ""
        gns = self.gns
        scalars = [("Gradient_Norm_Sqr", gns.sqr_avg()),
                   ("Gradient_Variance", gns.var_avg()),
                   ("Gain", self._state.gain),
                   ("Learning_Rate_Factor", self._state.lr_factor),
                   ("Accum_Scale", gns.accum_scale),
                   ("Progress", gns.get_progress())]
        if gns.accum_count > 0:
            scalars.append(("Accum_Count", gns.accum_count))
        for tag, value in scalars:
            writer.add_scalar(prefix + tag, value, global_step)


def mixed_precision_params(model, dtype=torch.bfloat16, min_dim=2):
    """Store the matrix-like parameters of ``model`` (``dim >= min_dim``:
    convolution / linear / embedding weights) in ``dtype`` -- call it BEFORE
    building the optimizer. With :class:`AdaptiveDataParallel`'s device
    engine the fused optimizer keeps fp32 master weights and fp32 momentum /
    Adam moments for them and writes the rounded 16-bit weights itself, so a
    bf16-autocast step runs without a single cast kernel: the forward reads
    the 16-bit weights directly, weight gradients arrive (and are
    all-reduced) in 16 bits. Normalisation parameters and biases stay fp32.
    Returns ``model``."""
    for p in model.parameters():
        if p.dim() >= min_dim and p.is_floating_point():
            p.data = p.data.to(dtype)
    return model


def _save_fast(obj, fileobj):
    """``torch.save`` without the per-record CRC-32. The zip writer's
    checksum runs at 0.4-1 GB/s on one core and is most of the time a
    checkpoint of a large model takes (1.36 GB of model + AdamW state:
    1.2-3.2 s with, 0.55 s without; ``profiles/r2_elastic/README.md``), on the
    path where a preempted job races its grace period. Nothing reads the
    checksum back (``torch.load`` does not verify it), and the file stays an
    ordinary ``torch.save`` file for the reference and for older
    checkpoints' readers."""
    get = getattr(torch.serialization, "get_crc32_options", None)
    put = getattr(torch.serialization, "set_crc32_options", None)
    if get is None or put is None or \
            os.environ.get("ADAPTDL_B200_CHECKPOINT_CRC", "0") == "1":
        torch.save(obj, fileobj)
        return
    previous = get()
    put(False)
    try:
        torch.save(obj, fileobj)
    finally:
        put(previous)


def _load_fast(fileobj):
    """``torch.load`` of a data-parallel state, memory-mapped when the file
    object is a real file: tensors are not read into fresh host memory first
    (2.1 s + 2.3 s of page faults and copies for 1.36 GB against 0.1 s +
    0.4 s), every replica of a node shares the same page-cache pages instead
    of holding its own copy, and ``load_state_dict`` moves them straight to
    the device. ``weights_only=False``: the numpy arrays inside
    ``optimizer.state["gns"]`` need full unpickling."""
    path = getattr(fileobj, "name", None)
    if isinstance(path, str) and os.path.isfile(path) and \
            os.environ.get("ADAPTDL_B200_CHECKPOINT_MMAP", "1") != "0":
        try:
            return torch.load(path, map_location="cpu", weights_only=False,
                              mmap=True)
        except (RuntimeError, ValueError, TypeError, OSError):
            fileobj.seek(0)          # legacy (non-zip) file, odd filesystem
    return torch.load(fileobj, map_location="cpu", weights_only=False)


class _AdaptiveDataParallelState(checkpoint.State):
    """``torch.save(([model_sd, optim_sd, sched_sd|None, scaler_sd|None],
    gain, lr_factor))`` -- the reference's layout (App. B); the GNS running
    averages ride inside ``optim_sd["state"]["gns"]``."""

    def __init__(self, model, optimizer, lr_scheduler, mp_scaler,
                 name="adaptdl-dataparallel", engine=None):
        super().__init__(name)
        self.engine = engine
        self.model = model
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.mp_scaler = mp_scaler
        self.gain = 1.0
        self.lr_factor = 1.0
        self.wide_state = None       # loaded, not yet applied to the engine
        self.adp = None              # set by AdaptiveDataParallel

    def sync(self):
        # called on every replica before a checkpoint is written
        if self.adp is not None and self.adp._deferred_buffers_stale:
            self.adp.sync_buffers()
        # device-resident estimator / Adam step counters -> host dicts
        if self.engine is not None and self.engine.enabled:
            self.engine.pull_gns_state(self.optimizer.state["gns"])

    def save(self, fileobj):
        state_dicts = [
            self.model.state_dict(),
            self.optimizer.state_dict(),
            self.lr_scheduler.state_dict()
            if self.lr_scheduler is not None else None,
            self.mp_scaler.state_dict()
            if self.mp_scaler is not None else None,
        ]
        if self.engine is not None and self.engine.enabled:
            wide = self.engine.wide_state()
            if wide:
                # fifth entry (ignored by the reference's loader): fp32
                # masters + optimizer state of bf16/fp16 parameters, which
                # Optimizer.load_state_dict would round to the param dtype
                state_dicts.append(wide)
        _save_fast((state_dicts, self.gain, self.lr_factor), fileobj)

    def load(self, fileobj):
        state_dicts, self.gain, self.lr_factor = _load_fast(fileobj)
        self.model.load_state_dict(state_dicts[0])
        self.optimizer.load_state_dict(state_dicts[1])
        if state_dicts[2] is not None and self.lr_scheduler is not None:
            self.lr_scheduler.load_state_dict(state_dicts[2])
        if state_dicts[3] is not None and self.mp_scaler is not None:
            self.mp_scaler.load_state_dict(state_dicts[3])
        if len(state_dicts) > 4 and state_dicts[4]:
            self.wide_state = state_dicts[4]
