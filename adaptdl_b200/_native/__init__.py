"""Native (CUDA/C++) side of adaptdl_b200: build + load + ctypes bindings.

Sources live in ``csrc/``; the shared library is built IN-TREE
(``adaptdl_b200/_native/libadl_b200.so``) with plain ``nvcc`` for sm_100a so
that it travels with the repo snapshot to GPU machines. A stamp file holds a
hash of the sources + flags; :func:`build` rebuilds only when it changes,
:func:`load` never rebuilds implicitly unless the library is missing/stale
*and* nvcc is present.

The library exposes a C ABI (kernels take raw device pointers and a
``cudaStream_t``); torch tensors are passed as ``tensor.data_ptr()`` and the
launching stream as ``torch.cuda.current_stream().cuda_stream``.
"""

import ctypes
import hashlib
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
# a source checkout keeps the kernels in <repo>/csrc; an installed package
# (sdist / wheel) carries a copy next to this file (setup.py: BuildPyWithCsrc)
CSRC = os.path.join(_ROOT, "csrc")
if not os.path.isdir(CSRC):
    CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libadl_b200.so")
STAMP_PATH = os.path.join(_HERE, "libadl_b200.stamp")

SOURCES = ["adl_kernels.cu", "adl_optim.cu", "adl_gemm.cu", "adl_bn.cu",
           "adl_ln.cu", "adl_transformer.cu", "adl_symm.cpp"]
HEADERS = ["adl_common.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math",
    "-Xcompiler", "-fPIC",
    "-shared",
]

MAX_RANKS = 16
MAX_CTAS = 128

_lock = threading.Lock()
_lib = None


def _nvcc():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") \
        or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _existing_sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _source_hash():
    h = hashlib.sha256()
    for name in _existing_sources() + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile csrc/ into the in-tree shared library (sm_100a)."""
    with _lock:
        if not force and not is_stale():
            return LIB_PATH
        nvcc = _nvcc()
        if nvcc is None:
            raise RuntimeError("nvcc not found; cannot build " + LIB_PATH)
        srcs = [os.path.join(CSRC, s) for s in _existing_sources()]
        tmp = LIB_PATH + ".tmp.{}".format(os.getpid())
        cmd = [nvcc] + NVCC_FLAGS + ["-I", CSRC, "-o", tmp] + srcs + ["-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        proc = subprocess.run(cmd, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + proc.stdout)
        os.replace(tmp, LIB_PATH)
        with open(STAMP_PATH, "w") as f:
            f.write(_source_hash())
        if verbose:
            print(proc.stdout)
        return LIB_PATH


class SegTable(ctypes.Structure):
    _fields_ = [("seg_end", ctypes.c_void_p), ("seg_group", ctypes.c_void_p),
                ("n_seg", ctypes.c_int)]


class ReduceArgs(ctypes.Structure):
    _fields_ = [
        ("buf", ctypes.c_void_p * MAX_RANKS),
        ("pad", ctypes.c_void_p * MAX_RANKS),
        ("rank", ctypes.c_int), ("world", ctypes.c_int),
        ("step_ctr", ctypes.c_void_p),
        ("site", ctypes.c_uint32),
        ("n_vec", ctypes.c_int),
        ("scale", ctypes.c_float),
        ("want_local", ctypes.c_int),
        ("segs", SegTable),
        ("n_groups", ctypes.c_int),
        ("pinv", ctypes.c_void_p),
        ("L", ctypes.c_void_p), ("T", ctypes.c_void_p),
        ("err", ctypes.c_void_p),
        ("timeout_ns", ctypes.c_ulonglong),
        ("mc_buf", ctypes.c_void_p),
        ("pinv_mode", ctypes.c_int), ("pinv_wide", ctypes.c_int),
        ("pinv_coef", ctypes.c_void_p),
        ("stage", ctypes.c_void_p * MAX_RANKS),
        ("fuse_fin", ctypes.c_int),
        ("ticket", ctypes.c_void_p),
    ]


class LocalArgs(ctypes.Structure):
    _fields_ = [
        ("g", ctypes.c_void_p), ("a", ctypes.c_void_p),
        ("pv", ctypes.c_void_p), ("pinv", ctypes.c_void_p),
        ("n_vec", ctypes.c_int),
        ("segs", SegTable),
        ("n_groups", ctypes.c_int),
        ("s0", ctypes.c_void_p), ("s1", ctypes.c_void_p),
        ("s2", ctypes.c_void_p),
        ("flag", ctypes.c_int),
        ("flag_ptr", ctypes.c_void_p),
        ("pinv_mode", ctypes.c_int), ("pinv_wide", ctypes.c_int),
        ("pinv_coef", ctypes.c_void_p),
        ("fuse_fin", ctypes.c_int),
        ("ticket", ctypes.c_void_p),
    ]


class FinalizeArgs(ctypes.Structure):
    _fields_ = [
        ("xchg", ctypes.c_void_p * MAX_RANKS),
        ("pad", ctypes.c_void_p * MAX_RANKS),
        ("rank", ctypes.c_int), ("world", ctypes.c_int),
        ("step_ctr", ctypes.c_void_p),
        ("site", ctypes.c_uint32),
        ("n_rows", ctypes.c_int),
        ("n_groups", ctypes.c_int),
        ("rows", ctypes.c_void_p * 4),
        ("sum_mask", ctypes.c_int),
        ("micro_steps", ctypes.c_int),
        ("pair_mode", ctypes.c_int),
        ("pair_flag", ctypes.c_int),
        ("pair_state", ctypes.c_void_p),
        ("mailbox", ctypes.c_void_p),
        ("ring", ctypes.c_int), ("slot_doubles", ctypes.c_int),
        ("result", ctypes.c_void_p),
        ("t_start", ctypes.c_void_p),
        ("gns_state", ctypes.c_void_p),
        ("gns_ctrl", ctypes.c_void_p),
        ("lr_factor", ctypes.c_void_p),
        ("err", ctypes.c_void_p),
        ("timeout_ns", ctypes.c_ulonglong),
        ("last_stamp", ctypes.c_void_p),
        ("clock", ctypes.c_void_p),
        ("amp_scale", ctypes.c_void_p),
    ]


class BcastArgs(ctypes.Structure):
    _fields_ = [
        ("staging", ctypes.c_void_p * MAX_RANKS),
        ("pad", ctypes.c_void_p * MAX_RANKS),
        ("rank", ctypes.c_int), ("world", ctypes.c_int),
        ("src", ctypes.c_int),
        ("step_ctr", ctypes.c_void_p),
        ("site", ctypes.c_uint32),
        ("dst", ctypes.c_void_p),
        ("n_vec", ctypes.c_longlong),
        ("err", ctypes.c_void_p),
        ("timeout_ns", ctypes.c_ulonglong),
    ]


class OptimArgs(ctypes.Structure):
    _fields_ = [
        ("grad", ctypes.c_void_p),
        ("state0", ctypes.c_void_p), ("state1", ctypes.c_void_p),
        ("param_ptr", ctypes.c_void_p),
        ("seg_start", ctypes.c_void_p), ("seg_numel", ctypes.c_void_p),
        ("segs", SegTable),
        ("n_vec", ctypes.c_int),
        ("hyper", ctypes.c_void_p),
        ("lr_factor", ctypes.c_void_p),
        ("step_ctr", ctypes.c_void_p),
        ("step_offset", ctypes.c_void_p),
        ("n_groups", ctypes.c_int),
        ("master", ctypes.c_void_p),
        ("grad_scale", ctypes.c_void_p),
        ("pinv_coef", ctypes.c_void_p),
    ]


MBOX_HDR = 16           # mailbox header doubles (adl_kernels.cu ADL_MBOX_HDR)
XCHG_TAIL = 4           # timing doubles after the statistic rows of an exchange record
CLOCK_DOUBLES = 4
# mailbox header fields
(MB_SEQ, MB_FINITE, MB_GAIN, MB_PROGRESS, MB_SYNC_NS, MB_ERR, MB_SCALE,
 MB_NROWS, MB_STEP_NS, MB_ACCUM_NS, MB_ACCUM_COUNT, MB_AMP_SCALE) = range(12)
# reduce flavours (adl_allreduce_gns)
(FLAVOUR_TWOSHOT, FLAVOUR_ONESHOT, FLAVOUR_NVLS) = range(3)
# preconditioner modes of the statistics kernels
(PINV_NONE, PINV_FLAT, PINV_ADAM) = range(3)
SITES_PER_STEP = 1024
HYPER_STRIDE = 8
(RULE_ADASCALE, RULE_ADAMSCALE, RULE_LINEAR, RULE_SQRT, RULE_LEGW) = range(5)
(CTL_ACCUM_SCALE, CTL_SMOOTHING, CTL_RULE, CTL_RULE_ARG, CTL_ENABLED) = \
    range(5)
GNS_TAIL = 8
(GNS_SQR_UNBIAS, GNS_VAR_UNBIAS, GNS_PROGRESS, GNS_BIASED) = range(4)


class BnArgs(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("res", ctypes.c_void_p),
        ("y", ctypes.c_void_p), ("dy", ctypes.c_void_p),
        ("dx", ctypes.c_void_p), ("dres", ctypes.c_void_p),
        ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
        ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p),
        ("num_batches_tracked", ctypes.c_void_p),
        ("dgamma", ctypes.c_void_p), ("dbeta", ctypes.c_void_p),
        ("partial", ctypes.c_void_p), ("counters", ctypes.c_void_p),
        ("mask", ctypes.c_void_p),
        ("cb", ctypes.c_int), ("coef", ctypes.c_void_p),
        ("M", ctypes.c_int), ("C", ctypes.c_int),
        ("n_partial", ctypes.c_int), ("relu", ctypes.c_int),
        ("eps", ctypes.c_float), ("momentum", ctypes.c_float),
    ]


class LnArgs(ctypes.Structure):
    _fields_ = [
        ("x", ctypes.c_void_p), ("h", ctypes.c_void_p),
        ("mask", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("z", ctypes.c_void_p), ("dh", ctypes.c_void_p),
        ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
        ("mean", ctypes.c_void_p), ("rstd", ctypes.c_void_p),
        ("partial", ctypes.c_void_p),
        ("dgamma", ctypes.c_void_p), ("dbeta", ctypes.c_void_p),
        ("M", ctypes.c_int), ("D", ctypes.c_int),
        ("n_partial", ctypes.c_int),
        ("scale", ctypes.c_float), ("eps", ctypes.c_float),
    ]


def _declare(lib):
    c = ctypes
    lib.adl_set_device.argtypes = [c.c_int]
    lib.adl_error_string.restype = c.c_char_p
    lib.adl_error_string.argtypes = [c.c_int]
    lib.adl_sm_count.argtypes = [c.c_int]
    lib.adl_allreduce_gns.argtypes = [c.POINTER(ReduceArgs),
                                      c.POINTER(FinalizeArgs), c.c_int,
                                      c.c_int, c.c_int, c.c_int, c.c_void_p]
    lib.adl_local.argtypes = [c.POINTER(LocalArgs), c.POINTER(FinalizeArgs),
                              c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
    lib.adl_finalize_stats.argtypes = [c.POINTER(FinalizeArgs), c.c_void_p]
    lib.adl_stamp.argtypes = [c.c_void_p, c.c_void_p]
    lib.adl_step_mark.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    lib.adl_optim_advance.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
    lib.adl_bcast_pull.argtypes = [c.POINTER(BcastArgs), c.c_int,
                                   c.c_void_p]
    lib.adl_symm_last_error.restype = c.c_char_p
    lib.adl_symm_round_size.argtypes = [c.c_int, c.c_size_t,
                                        c.POINTER(c.c_size_t)]
    lib.adl_symm_create.argtypes = [c.c_int, c.c_size_t,
                                    c.POINTER(c.c_ulonglong),
                                    c.POINTER(c.c_int)]
    lib.adl_symm_import.argtypes = [c.c_int, c.POINTER(c.c_ulonglong)]
    lib.adl_symm_map.argtypes = [c.c_ulonglong, c.c_size_t, c.c_int,
                                 c.POINTER(c.c_ulonglong)]
    lib.adl_symm_unmap.argtypes = [c.c_ulonglong, c.c_size_t]
    lib.adl_symm_release.argtypes = [c.c_ulonglong]
    lib.adl_topo_can_access_peer.argtypes = [
        c.c_int, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_int),
        c.POINTER(c.c_int)]
    lib.adl_topo_multicast_supported.argtypes = [c.c_int]
    lib.adl_topo_vmm_fd_supported.argtypes = [c.c_int]
    lib.adl_mc_round_size.argtypes = [c.c_int, c.c_size_t,
                                      c.POINTER(c.c_size_t)]
    lib.adl_mc_create.argtypes = [c.c_int, c.c_size_t,
                                  c.POINTER(c.c_ulonglong),
                                  c.POINTER(c.c_int)]
    lib.adl_mc_add_device.argtypes = [c.c_ulonglong, c.c_int]
    lib.adl_mc_bind.argtypes = [c.c_ulonglong, c.c_ulonglong, c.c_size_t]
    lib.adl_fused_optim.argtypes = [c.POINTER(OptimArgs), c.c_int, c.c_int,
                                    c.c_int, c.c_void_p]
    lib.adl_gemm_bias_act.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
        c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int,
        c.c_void_p, c.c_void_p, c.c_void_p]
    lib.adl_bn_act.argtypes = [c.POINTER(BnArgs), c.c_int, c.c_int, c.c_int,
                               c.c_int, c.c_void_p]

    lib.adl_dropout_add_ln.argtypes = [c.POINTER(LnArgs), c.c_int, c.c_int,
                                       c.c_int, c.c_void_p]
    lib.adl_heads_permute.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p,
                                      c.c_void_p, c.c_int, c.c_int, c.c_int,
                                      c.c_int, c.c_int, c.c_int, c.c_void_p]
    lib.adl_gelu_dropout_bwd.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p,
                                         c.c_void_p, c.c_longlong, c.c_float,
                                         c.c_int, c.c_void_p]
    lib.adl_colsum.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p,
                               c.c_void_p, c.c_int, c.c_int, c.c_int,
                               c.c_int, c.c_void_p]
    lib.adl_slice_cast.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                   c.c_int, c.c_int, c.c_void_p]
    for name, struct in (("adl_sizeof_bn_args", BnArgs),
                         ("adl_sizeof_ln_args", LnArgs),
                         ("adl_sizeof_optim_args", OptimArgs),
                         ("adl_sizeof_reduce_args", ReduceArgs),
                         ("adl_sizeof_local_args", LocalArgs),
                         ("adl_sizeof_finalize_args", FinalizeArgs),
                         ("adl_sizeof_bcast_args", BcastArgs)):
        got = getattr(lib, name)()
        if got != c.sizeof(struct):
            raise RuntimeError(
                "ABI mismatch for {}: C {} vs ctypes {}".format(
                    struct.__name__, got, c.sizeof(struct)))
    if lib.adl_mbox_hdr() != MBOX_HDR or lib.adl_xchg_tail() != XCHG_TAIL:
        raise RuntimeError("ABI mismatch: mailbox / exchange layout")
    return lib


def load(build_if_needed=True):
    """Load (building first if stale and nvcc is available) and return the
    ctypes handle. Raises if the library cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    if is_stale() and build_if_needed and _nvcc() is not None:
        build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("native library missing: " + LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    _lib = _declare(lib)
    return _lib


def check(code, what="cuda call"):
    if code != 0:
        lib = load()
        msg = lib.adl_error_string(code)
        raise RuntimeError("{} failed: {} ({})".format(
            what, msg.decode() if msg else "?", code))
