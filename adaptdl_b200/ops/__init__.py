"""Hand-written sm_100a compute ops (tensor-core path)."""

from adaptdl_b200.ops.linear_act import (LinearGELU, check_errors,
                                         gemm_bias_act, linear_act)

__all__ = ["LinearGELU", "linear_act", "gemm_bias_act", "check_errors"]
