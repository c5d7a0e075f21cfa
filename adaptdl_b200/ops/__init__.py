"""Hand-written sm_100a compute ops."""

from adaptdl_b200.ops.linear_act import (LinearGELU, check_errors,
                                         gemm_bias_act, linear_act)
from adaptdl_b200.ops.bn_act import BatchNormAct2d, bn_act
from adaptdl_b200.ops.layer_norm import dropout_add_layer_norm

from adaptdl_b200.ops._count import total as launch_count  # noqa: E402

__all__ = ["launch_count", "LinearGELU", "linear_act", "gemm_bias_act", "check_errors",
           "BatchNormAct2d", "bn_act", "dropout_add_layer_norm"]
