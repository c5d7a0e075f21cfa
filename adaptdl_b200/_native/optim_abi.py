"""ctypes declarations for the fused optimizer / estimator kernels
(csrc/adl_optim.cu). Kept separate so the reducer ABI stays readable."""

import ctypes


def declare(lib):
    if not hasattr(lib, "adl_sgd_step"):
        return
    c = ctypes
    lib.adl_sgd_step.argtypes = [c.c_void_p] * 3 + [c.c_longlong] + \
        [c.c_void_p] * 3 + [c.c_int] + [c.c_float] * 5 + [c.c_int] * 3 + \
        [c.c_void_p]
