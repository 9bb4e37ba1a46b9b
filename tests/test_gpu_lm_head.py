"""GPU parity of the hand-written tcgen05 student lm_head (kd_lm_head_fwd) and of the statistics its epilogue feeds into the
vocabulary-sharded KD (BASELINE configs[4]): against the fp32 product, against the streamed statistics kernel, and end to end
against the library-GEMM form of the same autograd op."""
import pytest
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib, sharded

pytestmark = pytest.mark.gpu

# (n_rows, V_local, H): full tiles, ragged rows / columns, H not a multiple of the 64-wide K block, more tiles than SMs
_SHAPES = [(128, 256, 64), (77, 264, 64), (300, 1000, 896), (1000, 5000, 896), (513, 2056, 200), (4096, 18992, 896), (1, 8, 8)]


def _operands(n, v, h, dtype, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    hidden = torch.randn(n, h, generator=g).to("cuda", dtype)
    weight = (torch.randn(v, h, generator=g) * (3.0 / h ** 0.5)).to("cuda", dtype)
    return hidden, weight


@pytest.mark.parametrize("n,v,h", _SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
def test_lm_head_gemm_matches_fp32_product(n, v, h, dtype):
    hidden, weight = _operands(n, v, h, dtype, seed=n + v)
    logits, partials = sharded.lm_head_fwd(hidden, weight)
    assert partials is None and logits.shape == (n, v) and logits.dtype == dtype
    ref = hidden.float() @ weight.float().t()
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11      # one rounding of the exact fp32 sum (+ summation order)
    err = (logits.float() - ref).abs()
    assert bool((err <= ulp * ref.abs() + 1e-4).all()), float(err.max())
    # and it is the same rounding the library GEMM produces, up to ties of the summation order
    lib = torch.matmul(hidden, weight.t())
    assert float((lib != logits).float().mean()) < 2e-2


def test_lm_head_gemm_strided_operands():
    hidden, weight = _operands(200, 520, 128, torch.bfloat16, seed=5)
    big_h = torch.zeros(200, 256, device="cuda", dtype=torch.bfloat16)
    big_h[:, :128] = hidden
    big_w = torch.zeros(520, 192, device="cuda", dtype=torch.bfloat16)
    big_w[:, :128] = weight
    a, _ = sharded.lm_head_fwd(big_h[:, :128], big_w[:, :128])
    b, _ = sharded.lm_head_fwd(hidden, weight)
    assert torch.equal(a, b)


def test_lm_head_gemm_rejects_what_tma_cannot_address():
    hidden, weight = _operands(16, 12, 64, torch.bfloat16)          # V_local % 8 != 0
    with pytest.raises(_lib.KDError):
        sharded.lm_head_fwd(hidden, weight)
    with pytest.raises(_lib.KDError):
        sharded.lm_head_fwd(hidden.float(), weight.float())


@pytest.mark.parametrize("n,v,h", [(77, 264, 64), (1000, 5000, 896), (4096, 18992, 896)])
@pytest.mark.parametrize("T", [2.0, 0.8])
def test_epilogue_statistics_match_the_streamed_kernel(n, v, h, T):
    hidden, weight = _operands(n, v, h, torch.bfloat16, seed=3)
    logits, partials = sharded.lm_head_fwd(hidden, weight, T=T)
    plain, _ = sharded.lm_head_fwd(hidden, weight)
    assert torch.equal(logits, plain)                                # the statistics do not perturb the product
    teacher = torch.randn(n, v, device="cuda").to(torch.bfloat16)
    labels = torch.randint(0, v, (n,), device="cuda")
    ph = sharded.CudaVocabPhases()
    streamed = ph.stats(teacher, logits, labels, 0, T)
    fused = ph.stats(teacher, logits, labels, 0, T, student_partials=partials)
    assert torch.equal(streamed[:, :7], fused[:, :7])               # teacher half: the same kernel
    assert torch.equal(streamed[:, 7], fused[:, 7])                 # student max
    k2 = 1.4426950408889634 / T
    lse_a = streamed[:, 7].double() * k2 + streamed[:, 8].double().log2()
    lse_b = fused[:, 7].double() * k2 + fused[:, 8].double().log2()
    assert float((lse_a - lse_b).abs().max()) <= 2e-6 * max(1.0, float(lse_a.abs().max()))


@pytest.mark.parametrize("mode", ["reference", "per_row", "kl"])
def test_lm_head_kd_matches_the_library_gemm_form(mode):
    n, v, h, T = 640, 4096, 896, 2.0
    hidden, weight = _operands(n, v, h, torch.bfloat16, seed=11)
    teacher = (torch.randn(n, v, device="cuda") * 2).to(torch.bfloat16)
    labels = torch.randint(0, v, (n,), device="cuda")
    out = {}
    for gemm in ("tcgen05", "library"):
        hd, wd = hidden.clone().requires_grad_(True), weight.clone().requires_grad_(True)
        loss = sharded.lm_head_kd_vocab_sharded(hd, wd, teacher, labels, 0, v, T, mode=mode, gemm=gemm)
        loss.backward()
        out[gemm] = (loss.item(), hd.grad.float(), wd.grad.float())
    la, ha, wa = out["tcgen05"]
    lb, hb, wb = out["library"]
    # the two GEMMs round a handful of logits differently (summation order): compare at the bf16 tolerance of the path
    assert abs(la - lb) <= 2e-4 * abs(lb)
    assert float((ha - hb).norm() / hb.norm()) <= 2e-3
    assert float((wa - wb).norm() / wb.norm()) <= 2e-3


def test_lm_head_kd_against_single_gpu_op():
    """configs[4] on one shard == kd.loca_kl on the logits the head produces."""
    n, v, h, T = 512, 2048, 256, 2.0
    hidden, weight = _operands(n, v, h, torch.bfloat16, seed=17)
    teacher = (torch.randn(n, v, device="cuda") * 2).to(torch.bfloat16)
    labels = torch.randint(0, v, (n,), device="cuda")
    logits, _ = sharded.lm_head_fwd(hidden, weight)
    ref = kd.loca_kl(teacher, logits.clone().requires_grad_(True), labels, T=T, alpha=0.8)
    got = sharded.lm_head_kd_vocab_sharded(hidden, weight, teacher, labels, 0, v, T, alpha=0.8)
    assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())


def test_lm_head_kd_backward_scales_with_grad_output():
    n, v, h, T = 256, 1024, 128, 2.0
    hidden, weight = _operands(n, v, h, torch.bfloat16, seed=23)
    teacher = (torch.randn(n, v, device="cuda") * 2).to(torch.bfloat16)
    labels = torch.randint(0, v, (n,), device="cuda")
    grads = []
    for scale in (1.0, 4.0):          # a power of two: the bf16 gradients must scale exactly
        hd, wd = hidden.clone().requires_grad_(True), weight.clone().requires_grad_(True)
        (sharded.lm_head_kd_vocab_sharded(hd, wd, teacher, labels, 0, v, T) * scale).backward()
        grads.append((hd.grad.float(), wd.grad.float()))
    assert float((grads[1][0] - 4.0 * grads[0][0]).norm() / grads[1][0].norm()) <= 2e-3
    assert float((grads[1][1] - 4.0 * grads[0][1]).norm() / grads[1][1].norm()) <= 2e-3
