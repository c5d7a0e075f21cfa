"""Host-only native code (no CUDA): build + load + ctypes bindings.

``csrc/host/*.cpp`` -> ``adaptdl_b200/_native/libadl_host.so`` with plain
``g++``: the parts of the framework that run on machines without a GPU or a
CUDA toolkit (the scheduler pods). Today that is the core of the Pollux genetic
search (``csrc/host/adl_pollux.cpp``). Same stamp-file protocol as the CUDA
library (:mod:`adaptdl_b200._native`): :func:`build` recompiles only when the
sources change, :func:`load` builds on first use when a compiler is around and
returns ``None`` (callers fall back to numpy) when it is not.
"""

import ctypes
import hashlib
import logging
import os
import shutil
import subprocess
import threading

LOG = logging.getLogger(__name__)

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
CSRC = os.path.join(_ROOT, "csrc", "host")
if not os.path.isdir(CSRC):
    CSRC = os.path.join(_HERE, "csrc", "host")
LIB_PATH = os.path.join(_HERE, "libadl_host.so")
STAMP_PATH = os.path.join(_HERE, "libadl_host.stamp")

SOURCES = ["adl_pollux.cpp"]
CXX_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-pthread"]

_lock = threading.Lock()
_lib = None
_failed = False


def _cxx():
    for cand in (os.environ.get("CXX"), "g++", "c++", "clang++"):
        if cand and shutil.which(cand):
            return shutil.which(cand)
    return None


def _source_hash():
    h = hashlib.sha256()
    for name in SOURCES:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(CXX_FLAGS).encode())
    return h.hexdigest()


def is_stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    if not os.path.isdir(CSRC):
        return False              # library shipped without its sources
    with open(STAMP_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force=False):
    """Compile csrc/host into the in-tree shared library."""
    with _lock:
        if not force and not is_stale():
            return LIB_PATH
        cxx = _cxx()
        if cxx is None:
            raise RuntimeError("no C++ compiler; cannot build " + LIB_PATH)
        srcs = [os.path.join(CSRC, s) for s in SOURCES]
        tmp = LIB_PATH + ".tmp.{}".format(os.getpid())
        proc = subprocess.run([cxx] + CXX_FLAGS + ["-o", tmp] + srcs,
                              stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("{} failed:\n{}".format(cxx, proc.stdout))
        os.replace(tmp, LIB_PATH)
        with open(STAMP_PATH, "w") as f:
            f.write(_source_hash())
        return LIB_PATH


def _bind(lib):
    c = ctypes
    p64 = c.POINTER(c.c_int64)
    p32 = c.POINTER(c.c_int32)
    pu8 = c.POINTER(c.c_uint8)
    pf64 = c.POINTER(c.c_double)
    lib.adl_pollux_create.restype = c.c_void_p
    lib.adl_pollux_create.argtypes = [
        c.c_int, c.c_int, c.c_int, p64, p64, p32, pu8, p32, p32, p32, p32,
        pf64, c.c_double, c.c_int, c.c_int, c.c_uint64, c.c_int]
    lib.adl_pollux_destroy.restype = None
    lib.adl_pollux_destroy.argtypes = [c.c_void_p]
    lib.adl_pollux_seed.restype = c.c_int
    lib.adl_pollux_seed.argtypes = [c.c_void_p, p32, c.c_int]
    lib.adl_pollux_run.restype = c.c_int
    lib.adl_pollux_run.argtypes = [c.c_void_p]
    lib.adl_pollux_missing.restype = c.c_int
    lib.adl_pollux_missing.argtypes = [c.c_void_p, p32, p32, p32, c.c_int]
    lib.adl_pollux_fill.restype = c.c_int
    lib.adl_pollux_fill.argtypes = [c.c_void_p, c.c_int, p32, p32, p32, pf64]
    lib.adl_pollux_timing.restype = None
    lib.adl_pollux_timing.argtypes = [c.c_void_p, pf64]
    lib.adl_pollux_population.restype = c.c_int
    lib.adl_pollux_population.argtypes = [c.c_void_p]
    lib.adl_pollux_generation.restype = c.c_int
    lib.adl_pollux_generation.argtypes = [c.c_void_p]
    lib.adl_pollux_result.restype = c.c_int
    lib.adl_pollux_result.argtypes = [c.c_void_p, p32, pf64]
    lib.adl_pollux_repair.restype = c.c_int
    lib.adl_pollux_repair.argtypes = [c.c_void_p, p32, c.c_uint64]
    lib.adl_pollux_mutate.restype = c.c_int
    lib.adl_pollux_mutate.argtypes = [c.c_void_p, p32, c.c_uint64]
    return lib


def load():
    """The bound library, or ``None`` when it cannot be had on this machine
    (no compiler and no pre-built copy). ``ADAPTDL_B200_NATIVE_HOST=0``
    forces ``None``."""
    global _lib, _failed
    if os.environ.get("ADAPTDL_B200_NATIVE_HOST", "1") == "0":
        return None
    if _lib is not None or _failed:
        return _lib
    try:
        # a differently built copy (tools/sanitize_host.sh: ASan / UBSan /
        # TSan instrumented) instead of the in-tree one
        override = os.environ.get("ADAPTDL_B200_HOST_LIB")
        if override:
            lib = _bind(ctypes.CDLL(override))
        else:
            if is_stale():
                build()
            lib = _bind(ctypes.CDLL(LIB_PATH))
    except (OSError, RuntimeError, AttributeError) as exc:
        LOG.warning("host native library unavailable (%s); using the numpy "
                    "implementations", str(exc).splitlines()[0])
        _failed = True
        return None
    _lib = lib
    return _lib
