"""ctypes binding of libkd_b200.so (the C ABI declared in include/kd_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is absent the
operators raise.  ``load()`` builds the library in-tree when it is missing or stale and nvcc exists.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

_lock = threading.Lock()
_lib: C.CDLL | None = None

# mirrors of the header enums
ABI_VERSION = 5
DTYPE_BF16, DTYPE_F16, DTYPE_F32 = 0, 1, 2
LOCA_REFERENCE, LOCA_PER_ROW, KL_PLAIN, KL_LOG_TARGET = 0, 1, 2, 3
FLAG_NO_GRAD, FLAG_FORCE_STREAMING = 1, 2
STATUS_BAD_LABEL = 1
TOKEN_RECORD_BYTES = 16
VSHARD_STAT_WORDS = 12

_p, _i64, _sz, _f, _d, _i, _u = C.c_void_p, C.c_int64, C.c_size_t, C.c_float, C.c_double, C.c_int, C.c_uint

# every symbol include/kd_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "kd_abi_version": (_i, []),
    "kd_last_error": (C.c_char_p, []),
    "kd_device_info": (_i, [C.POINTER(_i)] * 4),
    "kd_logit_kd_workspace_bytes": (_i, [_i64, _i64, C.POINTER(_sz)]),
    "kd_logit_kd_status_offset": (_i, [_i64, _i64, C.POINTER(_sz)]),
    "kd_logit_kd_rowaux_offset": (_i, [_i64, _i64, C.POINTER(_sz)]),
    "kd_logit_kd_fwd_bwd": (_i, [_p, _i64, _p, _i64, _p, _p, _i64, _p, _p, _sz, _i64, _i64, _i, _f, _f, _f, _i, _d, _p, _u, _p]),
    "kd_loca_teacher_stats": (_i, [_p, _i64, _p, _p, _p, _p, _i64, _i64, _i, _f, _f, _p]),
    "kd_loca_teacher_stats_peers": (_i, [_p, _i64, _p, _p, _i, _i64, _p, _p, _i64, _i64, _i, _f, _f, _p]),
    "kd_loca_table_workspace_bytes": (_i, [_i64, C.POINTER(_sz)]),
    "kd_loca_build_table": (_i, [_p, _i64, _i64, _p, _p, _p, _sz, _p]),
    "kd_loca_emit": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i, _f, _f, _d, _p, _p, _u, _p]),
    "kd_loca_emit_ce": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _f, _p, _i64, _p, _p, _p, _p, _i64, _i64, _i, _f, _f, _d, _p,
                             _p, _u, _p]),
    "kd_loca_regrad": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _f, _p, _p, _i64, _i64, _i64, _i, _f, _f, _d, _p, _p, _p]),
    "kd_logit_kd_regrad": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _sz, _i64, _i64, _i, _f, _f, _i, _d, _p, _p, _p]),
    "kd_logit_kd_ce_fwd_bwd": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _f, _p, _i64, _p, _p, _p, _sz, _i64, _i64, _i, _f, _f, _f, _d, _p, _u, _p]),
    "kd_logit_kd_ce_regrad": (_i, [_p, _i64, _p, _i64, _p, _p, _f, _p, _i64, _p, _sz, _i64, _i64, _i, _f, _f, _d, _p, _p, _p]),
    "kd_scale_rows": (_i, [_p, _i64, _i64, _i64, _i, _p, _p]),
    "kd_vshard_stats": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i, _f, _p]),
    "kd_vshard_combine": (_i, [_p, _i, _i64, _p, _i64, _f, _f, _i, _p, _p, _p, _p, _p]),
    "kd_vshard_accum": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _i64, _i64, _i, _f, _f, _i, _p]),
    "kd_vshard_emit": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _i64, _i64, _i, _f, _f, _i, _d, _p, _p]),
    "kd_lm_head_stat_tiles": (_i, [_i64]),
    "kd_lm_head_fwd": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _i64, _i, _f, _p]),
    "kd_lm_head_stats_merge": (_i, [_p, _i64, _p, _i64, _f, _p]),
    "kd_gemm": (_i, [_p, _i64, _i, _p, _i64, _i, _p, _i64, _i, _i64, _i64, _i64, _i, _p]),
    "kd_gemm_ex": (_i, [_p, _i64, _i, _p, _i64, _i, _p, _i64, _i, _i64, _i64, _i64, _i, _i, _p]),
    "kd_proj_mse_workspace_bytes": (_i, [_i64, _i64, C.POINTER(_sz)]),
    "kd_proj_mse_fwd_bwd": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p, _sz, _i64, _i64, _i64, _i, _p, _p]),
    "kd_info_nce_workspace_bytes": (_i, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "kd_info_nce_fwd_bwd": (_i, [_p, _p, _p, _p, _p, _sz, _i64, _i64, _i64, _i, _f, _p, _u, _p]),
    "kd_info_nce_regrad": (_i, [_p, _sz, _p, _i64, _i64, _i64, _i, _p, _p, _p]),
    "kd_info_nce_proj_workspace_bytes": (_i, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "kd_info_nce_proj_fwd_bwd": (_i, [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _sz, _i64, _i64, _i64, _i, _i, _f, _p, _u, _p]),
    "kd_info_nce_proj_regrad": (_i, [_p, _sz, _p, _p, _i64, _i64, _i64, _i64, _i, _i, _p, _p, _p]),
    "kd_feature_pool_workspace_bytes": (_i, [_i64, _i64, _i64, C.POINTER(_sz)]),
    "kd_feature_pool_fwd": (_i, [_p, _p, _p, _sz, _i64, _i64, _i64, _i, _p]),
    "kd_feature_pool_bwd": (_i, [_p, _p, _i64, _i64, _i64, _i, _p, _p]),
    "kd_feature_mse_workspace_bytes": (_i, [_i64, C.POINTER(_sz)]),
    "kd_feature_mse_fwd_bwd": (_i, [_p, _p, _p, _p, _p, _sz, _i64, _i, _p, _u, _p]),
    "kd_feature_mse_regrad": (_i, [_p, _p, _p, _i64, _i, _p, _p, _p]),
}


class KDError(RuntimeError):
    pass


def library_path() -> str:
    """In-tree library; ``KD_B200_LIB`` points at an alternative build (tuning experiments)."""
    return os.environ.get("KD_B200_LIB") or _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the shared library and declare every prototype."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if build_if_missing and not os.environ.get("KD_B200_LIB") and _build.is_stale():
            try:
                _build.build()
            except Exception as exc:  # noqa: BLE001
                if not os.path.isfile(path):
                    raise KDError(f"libkd_b200.so is missing and could not be built: {exc}") from exc
        if not os.path.isfile(path):
            raise KDError(f"{path} not found — run `python -m {__package__}.build` (no CPU fallback exists)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.kd_abi_version() != ABI_VERSION:
            raise KDError(f"ABI version mismatch: library {lib.kd_abi_version()}, binding {ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().kd_last_error().decode("utf-8", "replace")
        raise KDError(f"libkd_b200 error {rc}: {msg}")


_size_cache: dict = {}


def size_query(fn_name: str, *args) -> int:
    """Workspace-size queries are pure functions of their integer arguments: answered from a cache after the first call."""
    key = (fn_name, args)
    hit = _size_cache.get(key)
    if hit is not None:
        return hit
    out = _sz(0)
    check(getattr(load(), fn_name)(*args, C.byref(out)))
    if len(_size_cache) < 4096:
        _size_cache[key] = out.value
    return out.value


def device_info() -> dict:
    vals = [_i(0) for _ in range(4)]
    check(load().kd_device_info(*[C.byref(v) for v in vals]))
    keys = ("sm_count", "max_smem_per_block", "cluster4_max_active", "cluster8_max_active")
    return dict(zip(keys, (v.value for v in vals)))
