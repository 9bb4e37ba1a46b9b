"""CPU-side checks: the C-ABI library loads and exports every symbol include/kd_b200.h declares (no compute calls),
argument validation of the entry points, and the host logic of the Python mirror."""
import ctypes
import os
import re
import types

import pytest
import torch

from conftest import ROOT
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib, functional, losses, sharded

HEADER = os.path.join(ROOT, "include", "kd_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in kd_b200.h but not exported by libkd_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.kd_abi_version() == _lib.ABI_VERSION == 5


def test_size_queries_and_argument_validation_need_no_gpu():
    lib = _lib.load()
    n = _lib.size_query("kd_logit_kd_workspace_bytes", 16384, 151936)
    assert 2_000_000 < n < 4_000_000          # O(N) records + O(V) table, nothing O(N*V)
    assert _lib.size_query("kd_logit_kd_rowaux_offset", 16384, 151936) == 0
    assert _lib.size_query("kd_info_nce_workspace_bytes", 64, 729, 1152) > 3 * 64 * 1152 * 4
    out = ctypes.c_size_t(0)
    assert lib.kd_logit_kd_workspace_bytes(-1, 10, ctypes.byref(out)) != 0
    assert b"bad arguments" in lib.kd_last_error()
    # null pointers / bad shapes are rejected before any CUDA call
    rc = lib.kd_logit_kd_fwd_bwd(None, 8, None, 8, None, None, 8, None, None, 0, 4, 8, 0, 2.0, 0.8, 1e-8, 0, 1.0, None, 0, None)
    assert rc == 1 and b"null pointer" in lib.kd_last_error()
    rc = lib.kd_feature_mse_fwd_bwd(None, None, None, None, None, 0, 0, 0, None, 0, None)
    assert rc == 1
    # the tensor-core student head: tile count of the statistics block, argument checks before any CUDA call
    assert lib.kd_lm_head_stat_tiles(151936) == 1187 and lib.kd_lm_head_stat_tiles(256) == 2 and lib.kd_lm_head_stat_tiles(0) == 0
    assert lib.kd_lm_head_fwd(None, 896, None, 896, None, 151936, None, 16, 151936, 896, _lib.DTYPE_BF16, 2.0, None) == 1
    assert b"null pointer" in lib.kd_last_error()
    assert lib.kd_lm_head_fwd(None, 896, None, 896, None, 151936, None, 16, 151936, 896, _lib.DTYPE_F32, 2.0, None) == 2
    assert b"bf16 / fp16" in lib.kd_last_error()
    assert lib.kd_lm_head_fwd(None, 896, None, 896, None, 151936, None, 0, 151936, 896, _lib.DTYPE_BF16, 2.0, None) == 0
    assert lib.kd_lm_head_stats_merge(None, 0, None, 4, 2.0, None) == 1


def test_cpu_tensors_are_rejected_loudly():
    t = torch.zeros(2, 16)
    lab = torch.zeros(2, dtype=torch.long)
    for call in (lambda: kd.loca_kl(t, t, lab, 2.0), lambda: kd.temperature_kl(t, t, 2.0),
                 lambda: kd.info_nce(t, t), lambda: kd.feature_mse(t, t),
                 lambda: sharded.lm_head_fwd(t.to(torch.bfloat16), t.to(torch.bfloat16))):
        with pytest.raises(RuntimeError, match="CUDA-only"):
            call()


def test_reduction_normalisers():
    assert functional._inv_norm("mean", 6, 10, 2, None) == pytest.approx(1 / 60)
    assert functional._inv_norm("mean", 6, 10, 2, 24) == pytest.approx(1 / 240)      # token-sharded: global N
    assert functional._inv_norm("batchmean", 6, 10, 2, None) == pytest.approx(0.5)   # EV:130-136
    assert functional._inv_norm("sum", 6, 10, 2, None) == 1.0
    with pytest.raises(ValueError):
        functional._inv_norm("median", 1, 1, 1, None)


def test_row_view_is_zero_copy_for_strided_teacher():
    t = torch.zeros(3, 5, 40)[..., :32]           # wider teacher sliced like P1:155
    r = functional._rows(t)
    assert r.shape == (15, 32) and r.stride(0) == 40 and r.data_ptr() == t.data_ptr()
    c = torch.zeros(4, 32)
    assert functional._rows(c).data_ptr() == c.data_ptr()


def test_shard_rows_partitions_evenly():
    for n, w in ((16384, 8), (10, 3), (2, 4), (0, 2)):
        spans = [sharded.shard_rows(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_patch_reference_rebinds_the_reference_method_names():
    class Fake:                     # stands in for the reference LightningModule (method names as in P1/FB/EV)
        T = 2.0
        soft_target_loss_weight, ce_loss_weight = 0.1, 0.5

        def compute_loca_loss(self, teacher_logits, student_logits, student_loss, labels, alpha=0.8):
            raise AssertionError("reference body must not run after patching")

        def contrastive_loss(self, student_features, teacher_features, temperature=0.07):
            raise AssertionError

    patched = losses.patch_reference(Fake)
    assert patched == ["compute_loca_loss", "contrastive_loss"]
    assert Fake.compute_loca_loss is losses.KDLossMethods.compute_loca_loss
    obj = Fake()
    with pytest.raises(RuntimeError, match="CUDA-only"):      # now routed to the CUDA op (which refuses CPU tensors)
        obj.compute_loca_loss(torch.zeros(1, 2, 8), torch.zeros(1, 2, 8), torch.zeros(()), torch.zeros(1, 2, dtype=torch.long))
    inst = types.SimpleNamespace(T=1.0)
    assert losses.patch_reference(inst, only_existing=False) == list(losses._PATCHABLE)


def test_reference_class_can_be_patched_when_available():
    from oracle import ref_import as R
    if not R.available():
        pytest.skip("/root/reference not present")
    import copy
    K = R.reference_class("P1")
    sub = type("Patched", (K,), {})
    names = losses.patch_reference(sub)
    assert {"compute_loca_loss", "compute_vision_loss", "contrastive_loss"} <= set(names)
    import inspect
    for n in names:                 # same positional parameters as the reference methods
        assert list(inspect.signature(getattr(sub, n)).parameters) == list(inspect.signature(getattr(K, n)).parameters)
