"""Device-resident training-step engine.

With the engine enabled, nothing in an optimizer step needs the host:

* the gradient-noise-scale **estimator** (EMA state, AdaScale gain,
  per-group LR factors, scale-invariant progress) runs inside
  ``adl_finalize_stats`` on the device, right after the last bucket's fused
  all-reduce;
* the **optimizer** is one fused kernel per gradient arena
  (``adl_fused_optim``: SGD-momentum / Adam / AdamW) that multiplies each
  group's learning rate by the factor the estimator just wrote to device
  memory;
* the host sees the statistics through a pinned **mailbox ring**, consumed
  with a fixed lag (deterministic and identical on every replica), so there
  is no ``.item()`` / event synchronisation on the step path -- and the whole
  step can be captured in a CUDA graph (:mod:`adaptdl_b200.parallel.graph`).

The reference computes the same quantities on the host from numpy state
(``torch/scaling_rules.py:64-125``, ``torch/gradient_noise_scale.py:212-273``)
and pays two device synchronisations per step for it.
"""

import ctypes
import os

import numpy as np
import torch

from adaptdl_b200 import _native
from adaptdl_b200._native import (
    OptimArgs, HYPER_STRIDE, MBOX_HDR, MB_ERR, MB_PROGRESS, GNS_TAIL,
    GNS_SQR_UNBIAS,
    GNS_VAR_UNBIAS, GNS_PROGRESS, GNS_BIASED, CTL_ACCUM_SCALE, CTL_SMOOTHING,
    CTL_RULE, CTL_RULE_ARG, CTL_ENABLED, RULE_ADASCALE, RULE_ADAMSCALE,
    RULE_LINEAR, RULE_SQRT, RULE_LEGW, check)
from adaptdl_b200.parallel import layout
from adaptdl_b200.utils.trace import traced

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def stats_lag():
    """How many optimizer steps the host mirror trails the device by."""
    return max(0, int(os.environ.get("ADAPTDL_B200_STATS_LAG", "1")))


def _rule_code(rule):
    from adaptdl_b200.torch import scaling_rules as sr
    table = {sr.AdaScale: RULE_ADASCALE, sr.AdamScale: RULE_ADAMSCALE,
             sr.LinearScale: RULE_LINEAR, sr.SqrtScale: RULE_SQRT,
             sr.LEGWScale: RULE_LEGW}
    return table.get(type(rule))


def supported_optimizer(optimizer):
    """'sgd' / 'adam' / None (exact torch classes with options the fused
    kernel reproduces bit-for-bit in exact arithmetic)."""
    kind = type(optimizer)
    if kind is torch.optim.SGD:
        for g in optimizer.param_groups:
            if g.get("dampening", 0) != 0 or g.get("maximize", False):
                return None
        return "sgd"
    if kind in (torch.optim.Adam, torch.optim.AdamW):
        for g in optimizer.param_groups:
            if g.get("amsgrad", False) or g.get("maximize", False):
                return None
            if torch.is_tensor(g.get("lr")):
                return None
        return "adam"
    return None


class DeviceEngine(object):

    def __init__(self, reducer, optimizer, rule, gns_state,
                 precondition_stats=False, mp_scaler=None):
        self.reducer = reducer
        self.optimizer = optimizer
        self.rule = rule
        self.kind = supported_optimizer(optimizer)
        self.rule_code = _rule_code(rule)
        if self.kind is None or self.rule_code is None:
            raise ValueError("optimizer / scaling rule not supported by the "
                             "device engine")
        if precondition_stats and self.kind != "adam":
            raise ValueError("Adam-preconditioned statistics need Adam / "
                             "AdamW")
        # statistics of g / (sqrt(v_hat) + eps): the bucket kernels read the
        # fused optimizer's second-moment arena in place
        self.precondition_stats = bool(precondition_stats)
        self.mp_scaler = None
        self._lib = _native.load()
        dev = reducer.device
        G = reducer.num_groups
        self.num_groups = G
        self.enabled = True
        self.state = torch.zeros(4 * G + GNS_TAIL, dtype=torch.float64,
                                 device=dev)
        self.ctrl = torch.zeros(8, dtype=torch.float64, device=dev)
        self._ctrl_host = torch.zeros(8, dtype=torch.float64).pin_memory()
        self._ctrl_last = None
        self.lr_factor = torch.ones(G + 1, dtype=torch.float32, device=dev)
        self.hyper = torch.zeros(G, HYPER_STRIDE, dtype=torch.float32,
                                 device=dev)
        self._hyper_host = torch.zeros(G, HYPER_STRIDE,
                                       dtype=torch.float32).pin_memory()
        self._hyper_last = None
        self.opt_steps = torch.zeros(1, dtype=torch.int32, device=dev)
        self._zero_i32 = torch.zeros(1, dtype=torch.int32, device=dev)
        # per group: (rsqrt(1 - beta2^t) or 0 = no preconditioning, eps)
        self.pinv_coef = torch.zeros(G, 2, dtype=torch.float32, device=dev)
        self._opt_steps_host = 0
        self._consumed = -1          # last optimizer step mirrored on host
        self._tables = []
        for arena in reducer.arenas:
            self._tables.append(self._build_tables(arena))
        self.push_gns_state(gns_state)
        reducer.engine = self
        if mp_scaler is not None:
            self.attach_scaler(mp_scaler)

    # ------------------------------------------------------------------
    # AMP loss scaling without host synchronisation
    # ------------------------------------------------------------------

    def attach_scaler(self, scaler):
        """Keep the step host-sync free under ``torch.amp.GradScaler``: the
        scaler hands ``optimizer.grad_scale`` (a device tensor) to optimizers
        that declare ``_step_supports_amp_scaling`` instead of unscaling and
        checking for infs with a blocking ``.item()``. The estimator divides
        the scale out of the statistics on the device, flags non-finite
        gradients (``lr_factor[G]``) and the fused optimizer unscales /
        skips accordingly."""
        if not scaler.is_enabled():
            return
        dev = self.reducer.device
        if scaler._scale is None:
            scaler._lazy_init_scale_growth_tracker(dev)
        self.mp_scaler = scaler
        self.reducer.set_amp_scale(scaler._scale)
        self.optimizer._step_supports_amp_scaling = True

    def second_moments(self, arena_index):
        """``(flat exp_avg_sq arena, is_fp32_next_to_16bit_grads)``."""
        table = self._tables[arena_index]
        return table["state1"], table["master"] is not None

    def reset_adam_state(self):
        """What the reference does when the batch-size scale changes
        (``gradient_noise_scale.py:313-330`` with step=0: the moments are
        zeroed and the step count restarts), on the device."""
        for table in self._tables:
            for key in ("state0", "state1"):
                if table[key] is not None:
                    table[key].zero_()
        self.opt_steps.zero_()
        self.pinv_coef[:, 0].zero_()
        self._opt_steps_host = 0

    # ------------------------------------------------------------------
    # layout tables + optimizer state arenas
    # ------------------------------------------------------------------

    def _build_tables(self, arena):
        dev = self.reducer.device
        itemsize = torch.empty((), dtype=arena.dtype).element_size()
        vec = layout.VEC_BYTES // itemsize
        segs = sorted((seg for b in arena.buckets for seg in b.segments),
                      key=lambda s: s.start)
        n_vec = max(arena.total, 1) // vec
        ends = []
        for j, seg in enumerate(segs):
            nxt = segs[j + 1].start // vec if j + 1 < len(segs) else n_vec
            ends.append(nxt)
        params = [arena.params[seg.param_index] for seg in segs]
        for p in params:
            if not (p.is_contiguous() or _dense(p)):
                raise ValueError("parameter with a non-dense layout")
        t = {
            "segs": segs, "params": params, "n_vec": n_vec,
            "seg_end": torch.tensor(ends, dtype=torch.int32, device=dev),
            "seg_group": torch.tensor([s.group for s in segs],
                                      dtype=torch.int32, device=dev),
            "seg_start": torch.tensor([s.start for s in segs],
                                      dtype=torch.int32, device=dev),
            "seg_numel": torch.tensor([s.length for s in segs],
                                      dtype=torch.int32, device=dev),
            "param_ptr": torch.tensor([p.data_ptr() for p in params],
                                      dtype=torch.int64, device=dev),
            "state0": None, "state1": None, "master": None,
        }
        # 16-bit parameters are trained in mixed precision: fp32 master
        # weights and fp32 optimizer state in flat arenas, the fused kernel
        # writes the rounded 16-bit weights the forward pass reads
        wide = arena.dtype in (torch.bfloat16, torch.float16)
        state_dtype = torch.float32 if wide else arena.dtype
        needs0 = self.kind == "adam" or any(
            g.get("momentum", 0) != 0 for g in self.optimizer.param_groups)
        if needs0:
            t["state0"] = torch.zeros(max(arena.total, 1), dtype=state_dtype,
                                      device=dev)
        if self.kind == "adam":
            t["state1"] = torch.zeros(max(arena.total, 1), dtype=state_dtype,
                                      device=dev)
        if wide:
            t["master"] = torch.zeros(max(arena.total, 1),
                                      dtype=torch.float32, device=dev)
            for view, p in zip(self._state_views(t, "master"), params):
                view.copy_(p.detach())
        return t

    def _state_views(self, table, key):
        flat = table[key]
        out = []
        for seg, p in zip(table["segs"], table["params"]):
            piece = flat[seg.start:seg.start + seg.length]
            out.append(piece.as_strided(p.shape, p.stride())
                       if not p.is_contiguous() else piece.view(p.shape))
        return out

    def wide_state(self):
        """fp32 master weights + optimizer state of the 16-bit arenas, keyed
        by (arena index, segment index): what a checkpoint must carry on top
        of ``optimizer.state_dict()`` (whose loader rounds state tensors to
        the parameter dtype)."""
        out = {}
        for ai, table in enumerate(self._tables):
            if table["master"] is None:
                continue
            for key in ("master", "state0", "state1"):
                if table[key] is not None:
                    for si, view in enumerate(self._state_views(table, key)):
                        out[(ai, si, key)] = view.detach().cpu().clone()
        return out

    def load_wide_state(self, saved):
        for ai, table in enumerate(self._tables):
            if table["master"] is None:
                continue
            for key in ("master", "state0", "state1"):
                if table[key] is None:
                    continue
                views = self._state_views(table, key)
                for si, view in enumerate(views):
                    src = saved.get((ai, si, key))
                    if src is not None and src.shape == view.shape:
                        view.copy_(src)
            # the 16-bit weights are the rounded masters
            for view, p in zip(self._state_views(table, "master"),
                               table["params"]):
                p.data.copy_(view)

    def resync_master(self):
        """Re-derive the masters from the (just loaded / broadcast) 16-bit
        parameters -- used when no saved master is available."""
        for table in self._tables:
            if table["master"] is not None:
                for view, p in zip(self._state_views(table, "master"),
                                   table["params"]):
                    view.copy_(p.detach())

    def adopt_optimizer_state(self):
        """Move whatever state the torch optimizer holds (fresh, or just
        loaded from a checkpoint) into the flat arenas and expose views of
        the arenas as ``optimizer.state[p][...]``."""
        opt = self.optimizer
        names = ("momentum_buffer", None) if self.kind == "sgd" \
            else ("exp_avg", "exp_avg_sq")
        loaded_step = None
        for table in self._tables:
            views0 = self._state_views(table, "state0") \
                if table["state0"] is not None else None
            views1 = self._state_views(table, "state1") \
                if table["state1"] is not None else None
            for i, p in enumerate(table["params"]):
                st = opt.state.get(p, None)
                if st is None:
                    st = {}
                    opt.state[p] = st
                for views, name in ((views0, names[0]), (views1, names[1])):
                    if views is None or name is None:
                        continue
                    old = st.get(name)
                    if old is not None and old.data_ptr() != \
                            views[i].data_ptr():
                        views[i].copy_(old)
                    st[name] = views[i]
                if table["master"] is not None:
                    st["master_param"] = self._master_views(table)[i]
                if self.kind == "adam":
                    step = st.get("step")
                    if step is not None:
                        loaded_step = int(float(step))
                    else:
                        st["step"] = torch.tensor(0.0)
        if self.kind == "adam" and loaded_step is not None:
            self.opt_steps.fill_(loaded_step)
            self._opt_steps_host = loaded_step
        if self.kind == "adam":
            self._init_pinv_coef(loaded_step or 0)

    def _init_pinv_coef(self, step):
        rows = []
        for g in self.optimizer.param_groups:
            beta2, eps = float(g["betas"][1]), float(g["eps"])
            coef = (1.0 - beta2 ** step) ** -0.5 if step >= 5 else 0.0
            rows.append((coef, eps))
        self.pinv_coef.copy_(torch.tensor(rows, dtype=torch.float32))

    def _master_views(self, table):
        cached = table.get("_master_views")
        if cached is None:
            cached = self._state_views(table, "master")
            table["_master_views"] = cached
        return cached

    def refresh_param_pointers(self):
        for table in self._tables:
            ptrs = [p.data_ptr() for p in table["params"]]
            table["param_ptr"].copy_(torch.tensor(ptrs, dtype=torch.int64))

    # ------------------------------------------------------------------
    # estimator state <-> host
    # ------------------------------------------------------------------

    def push_gns_state(self, gns):
        """Host dict (``optimizer.state['gns']``) -> device."""
        G = self.num_groups
        host = np.zeros(4 * G + GNS_TAIL)
        sqr_unbias = float(gns.get("sqr_avg_unbias", 0.0))
        var_unbias = float(gns.get("var_avg_unbias", 0.0))
        host[0:G] = np.broadcast_to(gns.get("sqr_avg_biased", 0.0), (G,))
        host[G:2 * G] = np.broadcast_to(gns.get("var_avg_biased", 0.0), (G,))
        host[2 * G:3 * G] = np.broadcast_to(gns["sqr_avg"], (G,))
        host[3 * G:4 * G] = np.broadcast_to(gns["var_avg"], (G,))
        host[4 * G + GNS_SQR_UNBIAS] = sqr_unbias
        host[4 * G + GNS_VAR_UNBIAS] = var_unbias
        host[4 * G + GNS_PROGRESS] = float(gns.get("progress", 0.0))
        host[4 * G + GNS_BIASED] = 1.0 if gns.get("biased", False) else 0.0
        self.state.copy_(torch.from_numpy(host))

    def pull_gns_state(self, gns):
        """Device -> host dict (synchronises; used for checkpoints)."""
        G = self.num_groups
        host = self.state.cpu().numpy()
        gns["sqr_avg_biased"] = host[0:G].copy()
        gns["var_avg_biased"] = host[G:2 * G].copy()
        gns["sqr_avg"] = host[2 * G:3 * G].copy()
        gns["var_avg"] = host[3 * G:4 * G].copy()
        gns["sqr_avg_unbias"] = float(host[4 * G + GNS_SQR_UNBIAS])
        gns["var_avg_unbias"] = float(host[4 * G + GNS_VAR_UNBIAS])
        gns["progress"] = float(host[4 * G + GNS_PROGRESS])
        gns["biased"] = bool(host[4 * G + GNS_BIASED])
        if self.kind == "adam":
            steps = int(self.opt_steps.item())
            for table in self._tables:
                for p in table["params"]:
                    self.optimizer.state[p]["step"] = torch.tensor(
                        float(steps))

    def set_progress(self, progress):
        G = self.num_groups
        self.state[4 * G + GNS_PROGRESS] = float(progress)

    # ------------------------------------------------------------------
    # host -> device control (only copies when something changed)
    # ------------------------------------------------------------------

    def sync_ctrl(self, accum_scale, smoothing, legw_unit=0.0):
        vals = (float(accum_scale), float(smoothing), float(self.rule_code),
                float(legw_unit), 1.0 if self.enabled else 0.0)
        if vals == self._ctrl_last:
            return
        self._ctrl_last = vals
        # a fresh pinned staging buffer per change (changes are rare): the
        # previous async copy may not have executed yet
        h = self._ctrl_host = torch.zeros(8, dtype=torch.float64).pin_memory()
        h[CTL_ACCUM_SCALE], h[CTL_SMOOTHING] = vals[0], vals[1]
        h[CTL_RULE], h[CTL_RULE_ARG], h[CTL_ENABLED] = vals[2], vals[3], \
            vals[4]
        self.ctrl.copy_(h, non_blocking=True)

    def sync_hyper(self):
        rows = []
        for g in self.optimizer.param_groups:
            if self.kind == "sgd":
                rows.append((float(g["lr"]), float(g.get("momentum", 0)),
                             float(g.get("weight_decay", 0)),
                             1.0 if g.get("nesterov", False) else 0.0,
                             0.0, 0.0, 0.0, 0.0))
            else:
                b1, b2 = g["betas"]
                adamw = type(self.optimizer) is torch.optim.AdamW or \
                    g.get("decoupled_weight_decay", False)
                rows.append((float(g["lr"]), float(b1),
                             float(g.get("weight_decay", 0)),
                             1.0 if adamw else 0.0, float(b2),
                             float(g["eps"]), 0.0, 0.0))
        rows = tuple(rows)
        if rows == self._hyper_last:
            return
        self._hyper_last = rows
        self._hyper_host = torch.tensor(rows, dtype=torch.float32) \
            .pin_memory()
        self.hyper.copy_(self._hyper_host, non_blocking=True)

    # ------------------------------------------------------------------
    # the fused optimizer step
    # ------------------------------------------------------------------

    @traced("optimizer_step")
    def optimizer_step(self):
        self.sync_hyper()
        red = self.reducer
        stream = torch.cuda.current_stream(red.device).cuda_stream
        grad_scale = None
        if self.mp_scaler is not None:
            # handed over by GradScaler.step (None = already unscaled)
            grad_scale = getattr(self.optimizer, "grad_scale", None)
        G = self.num_groups
        if self.kind == "adam":
            # the step count advances only when the update is applied
            check(self._lib.adl_optim_advance(
                self.opt_steps.data_ptr(),
                self.lr_factor.data_ptr() + 4 * G, stream),
                "adl_optim_advance")
            red.launches += 1
        for arena, table in zip(red.arenas, self._tables):
            args = OptimArgs()
            args.grad = arena.grad.data_ptr()
            args.state0 = table["state0"].data_ptr() \
                if table["state0"] is not None else None
            args.state1 = table["state1"].data_ptr() \
                if table["state1"] is not None else None
            args.param_ptr = table["param_ptr"].data_ptr()
            args.seg_start = table["seg_start"].data_ptr()
            args.seg_numel = table["seg_numel"].data_ptr()
            args.segs.seg_end = table["seg_end"].data_ptr()
            args.segs.seg_group = table["seg_group"].data_ptr()
            args.segs.n_seg = table["seg_end"].numel()
            args.n_vec = table["n_vec"]
            args.hyper = self.hyper.data_ptr()
            args.lr_factor = self.lr_factor.data_ptr()
            # adam step = opt_steps (device), advanced just above
            args.step_ctr = self.opt_steps.data_ptr()
            args.step_offset = self._zero_i32.data_ptr()
            args.n_groups = self.num_groups
            args.master = table["master"].data_ptr() \
                if table["master"] is not None else None
            if grad_scale is not None:
                args.grad_scale = grad_scale.data_ptr()
            if self.kind == "adam" and self.precondition_stats:
                args.pinv_coef = self.pinv_coef.data_ptr()
            grid = max(1, min(2 * red._sm_count,
                              (table["n_vec"] + 2 * 512 - 1) // (2 * 512)))
            check(self._lib.adl_fused_optim(
                ctypes.byref(args), 1 if self.kind == "adam" else 0,
                _DTYPE_CODE[arena.dtype], grid, stream), "adl_fused_optim")
            red.launches += 1
        self._opt_steps_host += 1
        self.optimizer._opt_called = True    # keep LR schedulers quiet

    # ------------------------------------------------------------------
    # host mirror of the device statistics
    # ------------------------------------------------------------------

    def mirror(self, gns_dict, force_latest=False):
        """Fold the newest mailbox slot the lag policy allows into the host
        dict. Returns the slot header (numpy) or ``None`` if nothing new."""
        red = self.reducer
        latest = red._steps - 1
        target = latest if force_latest else latest - stats_lag()
        if target < 0 or target <= self._consumed:
            return None
        arr = red.read_slot(target)
        self._consumed = target
        if int(arr[MB_ERR]) != 0:
            raise RuntimeError("fused all-reduce timed out waiting for a "
                               "peer (error word {})".format(int(arr[MB_ERR])))
        G = self.num_groups
        gns_dict["sqr_avg"] = np.array(arr[MBOX_HDR:MBOX_HDR + G])
        gns_dict["var_avg"] = np.array(arr[MBOX_HDR + G:MBOX_HDR + 2 * G])
        gns_dict["progress"] = float(arr[MB_PROGRESS])
        return np.array(arr[:MBOX_HDR])


def _dense(p):
    try:
        return torch.empty_like(p).stride() == p.stride()
    except RuntimeError:
        return False
