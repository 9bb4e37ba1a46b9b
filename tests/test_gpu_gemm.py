"""GPU parity of the general tcgen05 product (kd_gemm): K-major and MN-major operands, 16-bit and fp32 outputs, ragged shapes —
against the fp32 product of the same (16-bit) operands — and of the projected feature MSE (BASELINE configs[3]) against
oracle/kd_oracle.feature_mse applied to the library-rounded projection, incl. the gradients w.r.t. input and projection."""
import pytest
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib
from oracle import kd_oracle as O
from _util import rel_err

pytestmark = pytest.mark.gpu

# (M, N, K): whole tiles, ragged everywhere, K not a multiple of the 64-wide block, the lm_head backward shapes (one of 8 shards, fewer rows)
_SHAPES = [(256, 256, 64), (128, 264, 128), (304, 1000, 896), (520, 896, 2056), (1000, 896, 18992), (18992, 896, 1024), (8, 8, 8)]
_MAJORS = [(False, False), (False, True), (True, True)]


def _mats(M, N, K, a_mn, b_mn, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to("cuda", dtype)
    b = (torch.randn(N, K, generator=g) * (2.0 / K ** 0.5)).to("cuda", dtype)
    ref = (a.double() @ b.double().t()).float()           # exact product of the 16-bit operands
    return (a.t().contiguous() if a_mn else a), (b.t().contiguous() if b_mn else b), ref


@pytest.mark.parametrize("a_mn,b_mn", _MAJORS, ids=["KK", "K-MN", "MN-MN"])
@pytest.mark.parametrize("M,N,K", _SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_gemm_matches_fp32_product(M, N, K, a_mn, b_mn, dtype):
    a, b, ref = _mats(M, N, K, a_mn, b_mn, dtype, seed=M + N + K)
    out = kd.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn)
    assert out.shape == (M, N) and out.dtype == dtype
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11      # one rounding of the fp32 sum (+ summation order)
    err = (out.float() - ref).abs()
    assert bool((err <= ulp * ref.abs() + 2e-4 * ref.abs().max()).all()), float(err.max())
    out32 = kd.gemm(a, b, a_mn_major=a_mn, b_mn_major=b_mn, out_dtype=torch.float32)
    assert out32.dtype == torch.float32 and rel_err(out32, ref) <= 5e-5      # fp32 accumulation in tensor memory over K <= 18 992


def test_gemm_strided_operands_and_unsupported_layouts():
    a, b, ref = _mats(200, 520, 128, False, True, torch.bfloat16, seed=5)
    big_a = torch.zeros(200, 256, device="cuda", dtype=torch.bfloat16)
    big_a[:, :128] = a
    big_b = torch.zeros(128, 1024, device="cuda", dtype=torch.bfloat16)
    big_b[:, :520] = b
    out = kd.gemm(big_a[:, :128], big_b[:, :520], b_mn_major=True)
    assert torch.equal(out, kd.gemm(a, b, b_mn_major=True))
    with pytest.raises(_lib.KDError):
        kd.gemm(a.t().contiguous(), b.t().contiguous(), a_mn_major=True, b_mn_major=False)       # not built
    with pytest.raises(_lib.KDError):
        kd.gemm(a.float(), b.float(), b_mn_major=True)


@pytest.mark.parametrize("M,K,N", [(300, 896, 3584), (65, 64, 72), (4096, 896, 3584)])
def test_projected_mse_matches_oracle(M, K, N):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(K, N, generator=g) * 0.03).to(torch.bfloat16)
    t = torch.randn(M, N, generator=g).to(torch.bfloat16)
    # oracle: the reference's mse_loss (EV:160-165) on the projection as a bf16 GEMM hands it on, gradients by autograd in fp64
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    proj = xd @ wd
    proj_r = proj + (proj.detach().float().to(torch.bfloat16).double() - proj.detach())       # value rounded, gradient straight through
    ref = O.feature_mse(proj_r, t)
    loss_ref = ((proj_r - t.double()) ** 2).mean()
    loss_ref.backward()
    assert abs(ref.loss.item() - loss_ref.item()) <= 1e-12
    xs, ws = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True)
    loss = kd.feature_mse(xs, t.cuda(), proj=ws)
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    assert rel_err(xs.grad.float().cpu(), xd.grad) <= 4e-3         # two bf16 roundings: the projected gradient, then the product
    assert rel_err(ws.grad.float().cpu(), wd.grad) <= 4e-3
    # and the same through a 3-D input, the way hidden states arrive ([B, S, K])
    xs3 = x.reshape(5, -1, K).cuda().requires_grad_(True) if M % 5 == 0 else None
    if xs3 is not None:
        l3 = kd.feature_mse(xs3, t.reshape(5, -1, N).cuda(), proj=w.cuda())
        assert abs(l3.item() - loss.item()) <= 1e-7 * abs(loss.item())


def test_gemm_on_a_capped_grid_gives_the_same_product():
    """`kd_gemm_ex(max_ctas=...)`: the persistent kernel on fewer CTA pairs (what the vocabulary-parallel head's dW product uses
    while the dH all-reduce runs) walks the same tiles — bit-identical output."""
    g = torch.Generator().manual_seed(5)
    a = torch.randn(1500, 520, generator=g).to(torch.bfloat16).cuda()
    b = torch.randn(704, 520, generator=g).to(torch.bfloat16).cuda()
    full = kd.gemm(a, b)
    for cap in (2, 6, 120, 10_000):
        assert torch.equal(kd.gemm(a, b, max_ctas=cap), full)
    with pytest.raises(_lib.KDError):
        kd.gemm(a, b, max_ctas=-1)

