"""B200-native knowledge-distillation losses (LoCa logit KL, temperature KL, InfoNCE, feature MSE).

Hand-written sm_100a CUDA behind a C ABI (``include/kd_b200.h`` / ``libkd_b200.so``), exposed as
``torch.autograd.Function``s that drop into the reference's loss call sites.  See DESIGN.md.
"""
from .functional import feature_mse, gemm, info_nce, loca_kl, loca_kl_ce, shift_labels, temperature_kl  # noqa: F401
from .losses import KDLoss, KDLossMethods, patch_reference  # noqa: F401

__all__ = ["loca_kl", "loca_kl_ce", "shift_labels", "temperature_kl", "info_nce", "feature_mse", "gemm", "KDLoss", "KDLossMethods", "patch_reference"]
