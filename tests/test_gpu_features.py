"""GPU parity of InfoNCE and feature MSE against goldens / oracle (tolerances as in BASELINE.json)."""
import pytest
import torch

from _util import goldens_of_kind, grad_close, load_golden, rel_err
from oracle import kd_oracle as O
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", goldens_of_kind("info_nce"))
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_info_nce_goldens(name, dtype):
    g = load_golden(name)
    xs = torch.from_numpy(g["student"]).to("cuda", dtype).requires_grad_(True)
    xt = torch.from_numpy(g["teacher"]).to("cuda", dtype)
    loss = kd.info_nce(xs, xt, float(g["temperature"]))
    loss.backward()
    # losses near zero (well-separated diagonal) are compared absolutely at the scale of log-softmax rounding
    assert abs(loss.item() - float(g["loss"])) <= max(1e-5 * abs(float(g["loss"])), 2e-7)
    assert grad_close(xs.grad.float().cpu(), g["grad"]) if dtype == torch.bfloat16 else rel_err(xs.grad.float().cpu(), g["grad"]) <= 1e-4


@pytest.mark.parametrize("M,P,D", [(64, 729, 1152), (16, 10, 1152), (8, 1, 896), (5, 3, 100), (3, 1000, 4104), (2, 5000, 16), (1, 1, 8),
                                   (300, 7, 64)])
def test_info_nce_against_oracle(M, P, D):
    xs, xt = synthetic.vision_feature_batch(M, P, D, seed=M)
    if M == 16:     # unrelated teacher: every softmax entry matters
        xt = synthetic.vision_feature_batch(M, P, D, seed=999)[0]
    ref = O.info_nce(xs, xt)
    xd = xs.cuda().requires_grad_(True)
    loss = kd.info_nce(xd, xt.cuda())
    loss.backward()
    assert abs(loss.item() - ref.loss.item()) <= max(1e-5 * abs(ref.loss.item()), 2e-7)
    assert grad_close(xd.grad.float().cpu(), ref.grad)


def test_info_nce_2d_features_as_the_reference_passes_them():
    """contrastive_loss receives already pooled+normalised [B, 1152] features (P1:243-248 -> P1:345)."""
    xs, xt = synthetic.vision_feature_batch(12, 4, 1152, seed=3)
    fs = torch.nn.functional.normalize(xs.float().mean(1), dim=-1)
    ft = torch.nn.functional.normalize(synthetic.vision_feature_batch(12, 4, 1152, seed=4)[0].float().mean(1), dim=-1)
    ref = O.info_nce(fs, ft, pool=False)
    fd = fs.cuda().requires_grad_(True)
    loss = kd.info_nce(fd, ft.cuda())
    loss.backward()
    assert abs(loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item())
    assert rel_err(fd.grad.cpu(), ref.grad) <= 1e-4


@pytest.mark.parametrize("M,P,D,related", [(2048, 1, 1152, True), (1032, 1, 256, False), (4096, 1, 1152, False), (1024, 3, 512, True)])
def test_info_nce_token_level_on_the_tensor_cores(M, P, D, related):
    """contrastive_loss accepts any [M, D] (P1:393-416); at token-level M the similarity GEMM (P1:410) runs on the tcgen05
    kernel with the row log-softmax in its epilogue.  Oracle: the fp64 restatement on the GPU (same bf16-rounded inputs).
    Gates: loss 1e-5 relative, bf16 gradient 2e-3 (tests/_util.grad_close), and agreement with the exact fp32 row kernel."""
    xs, xt = synthetic.vision_feature_batch(M, P, D, seed=M + P)
    if not related:
        xt = synthetic.vision_feature_batch(M, P, D, seed=4242)[0]
    if P == 1:
        xs, xt = xs[:, 0], xt[:, 0]
    ref = O.info_nce(xs.cuda(), xt.cuda())
    xd = xs.cuda().requires_grad_(True)
    loss = kd.info_nce(xd, xt.cuda())
    loss.backward()
    assert abs(loss.item() - ref.loss.item()) <= max(1e-5 * abs(ref.loss.item()), 2e-7)
    assert grad_close(xd.grad.float().cpu(), ref.grad.cpu())
    xc = xs.cuda().requires_grad_(True)
    loss_c = kd.info_nce(xc, xt.cuda(), similarity="cuda_core")
    loss_c.backward()
    assert abs(loss.item() - loss_c.item()) <= max(1e-5 * abs(loss_c.item()), 2e-7)
    assert rel_err(xd.grad.float(), xc.grad.float()) <= 2e-3
    # a different incoming gradient re-emits from the pooled gradient kept in the workspace
    xe = xs.cuda().requires_grad_(True)
    (kd.info_nce(xe, xt.cuda()) * 0.5).backward()
    assert grad_close(xe.grad.float().cpu(), ref.grad.cpu() * 0.5)


def test_mse_golden_and_oracle():
    g = load_golden("mse_small")
    a = torch.from_numpy(g["input"]).to("cuda", torch.bfloat16).requires_grad_(True)
    b = torch.from_numpy(g["target"]).to("cuda", torch.bfloat16)
    loss = kd.feature_mse(a, b)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * float(g["loss"])
    assert grad_close(a.grad.float().cpu(), g["grad"])
    hs, ht, w = synthetic.hidden_state_batch(4096, 896, 3584, seed=2)
    proj = (hs.float() @ w.float()).to(torch.bfloat16)       # projection is the caller's GEMM (not in the reference)
    ref = O.feature_mse(proj, ht)
    pd = proj.cuda().requires_grad_(True)
    loss = kd.feature_mse(pd, ht.cuda())
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.loss.item()) <= 1e-5 * ref.loss.item()
    assert grad_close(pd.grad.float().cpu(), ref.grad * 3.0)


def test_info_nce_grad_output_scaling():
    xs, xt = synthetic.vision_feature_batch(16, 10, 1152, seed=5)
    xt = synthetic.vision_feature_batch(16, 10, 1152, seed=77)[0]
    ref = O.info_nce(xs, xt)
    xd = xs.cuda().requires_grad_(True)
    (kd.info_nce(xd, xt.cuda()) * 0.5).backward()         # weight 0.5 of phase 1 (P1:68, P1:351)
    assert grad_close(xd.grad.float().cpu(), ref.grad * 0.5)


@pytest.mark.parametrize("M,P,D,dtype", [(64, 729, 1152, torch.bfloat16), (7, 33, 100, torch.bfloat16), (1, 5, 8, torch.float16),
                                         (130, 3, 72, torch.float32), (2, 3000, 264, torch.bfloat16)])
def test_pool_tokens_is_the_fp32_mean_over_tokens(M, P, D, dtype):
    """`kd_feature_pool_fwd` / `_bwd`: the `.mean(dim=1)` of P1:243-246 in fp32 and its store-only broadcast backward."""
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.functional import pool_tokens
    g = torch.Generator().manual_seed(M + P)
    x = torch.randn(M, P, D, generator=g).to(dtype)
    w = torch.randn(M, D, generator=g)
    xd = x.cuda().requires_grad_(True)
    pooled = pool_tokens(xd)
    assert pooled.dtype == torch.float32
    assert rel_err(pooled.cpu(), x.double().mean(1)) <= 2e-6
    (pooled * w.cuda()).sum().backward()
    want = (w.double() / P).unsqueeze(1).expand(M, P, D)
    assert rel_err(xd.grad.cpu(), want) <= (1e-6 if dtype == torch.float32 else 3e-3)


def test_pool_into_an_unaligned_output_through_the_c_abi():
    """A C caller may hand `kd_feature_pool_fwd` an output that is only 4-byte aligned: the finishing kernel then stores scalars."""
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib
    lib = _lib.load()
    M, P, D = 5, 11, 64
    x = torch.randn(M, P, D, device="cuda").to(torch.bfloat16)
    buf = torch.zeros(M * D + 1, device="cuda")
    out = buf[1:]                                                   # 4 bytes past a 16-byte boundary
    need = _lib.size_query("kd_feature_pool_workspace_bytes", M, P, D)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    assert out.data_ptr() % 16 == 4
    assert lib.kd_feature_pool_fwd(x.data_ptr(), out.data_ptr(), ws.data_ptr(), need, M, P, D, _lib.DTYPE_BF16,
                                   torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert rel_err(out.view(M, D).cpu(), x.double().mean(1).cpu()) <= 2e-6


def _proj_reference(xs, xt, W, temperature=0.07):
    """contrastive_loss(x_s @ W, x_t) after the pooling of P1:243-248, in fp64 with autograd gradients."""
    x64, W64 = xs.double().requires_grad_(True), W.double().requires_grad_(True)
    y = x64 @ W64
    us = torch.nn.functional.normalize(y.mean(1) if y.dim() == 3 else y, dim=-1)
    t64 = xt.double()
    ut = torch.nn.functional.normalize(t64.mean(1) if t64.dim() == 3 else t64, dim=-1)
    lg = us @ ut.T / temperature
    loss = (torch.logsumexp(lg, -1) - lg.diagonal()).mean()
    gx, gW = torch.autograd.grad(loss, (x64, W64))
    return loss.item(), gx, gW


@pytest.mark.parametrize("M,P,Ds,Dt,xdtype,wdtype,teacher_2d", [
    (64, 27, 1152, 896, torch.bfloat16, torch.bfloat16, False),      # BASELINE configs[2] widths, fewer tokens
    (12, 9, 256, 640, torch.bfloat16, torch.float32, False),
    (7, 5, 100, 36, torch.float32, torch.float32, False),            # nothing 16-byte vectorisable
    (33, 4, 128, 72, torch.float16, torch.float16, True),            # teacher already pooled
    (130, 1, 64, 200, torch.bfloat16, torch.bfloat16, True),         # 2-D student rows, more than one row tile
])
def test_pooled_info_nce_through_a_projection_in_one_call(M, P, Ds, Dt, xdtype, wdtype, teacher_2d):
    """BASELINE configs[2]: the fused projected form (`kd_info_nce_proj_fwd_bwd`).  The oracle is the reference's
    contrastive_loss applied to x_s @ W (the projection is not in the reference); gradients w.r.t. the features AND the
    projection against fp64 autograd of the same formula."""
    g = torch.Generator().manual_seed(M * 1000 + Ds)
    xs = torch.randn(M, P, Ds, generator=g).to(xdtype)
    W = (torch.randn(Ds, Dt, generator=g) * 0.1).to(wdtype)
    # a related teacher: its features are a noisy linear image of the student's, so the diagonal means something
    xt = ((xs.float() @ W.float()) + 0.5 * torch.randn(M, P, Dt, generator=g)).to(xdtype)
    if teacher_2d:
        xt = xt.float().mean(1).to(xdtype)
    if P == 1:
        xs = xs[:, 0].contiguous()
        if not teacher_2d:
            xt = xt[:, 0].contiguous()
    want, gx, gW = _proj_reference(xs, xt, W)
    y64 = xs.double() @ W.double()
    if y64.dim() == 3 and xt.dim() == 2:            # the oracle pools both or neither
        y64 = y64.mean(1)
    assert abs(O.info_nce(y64, xt).loss.item() - want) <= 1e-9 * max(abs(want), 1e-3)
    xd, Wd = xs.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
    # fp16 gradients of a mean loss sit in fp16's subnormal range: scale the loss as the reference's fp16 AMP does (TR2:145)
    scale = 4096.0 if torch.float16 in (xdtype, wdtype) else 1.0
    loss = kd.info_nce(xd, xt.cuda(), proj=Wd, upstream=scale if scale != 1.0 else None)
    (loss * scale).backward()
    gx, gW = gx * scale, gW * scale
    assert abs(loss.item() - want) <= max(2e-5 * abs(want), 2e-7)
    if wdtype == torch.float32:
        assert rel_err(Wd.grad.cpu(), gW) <= 1e-4
    else:
        assert grad_close(Wd.grad.float().cpu(), gW, dtype=wdtype)
    if xdtype == torch.float32:
        assert rel_err(xd.grad.cpu(), gx) <= 1e-4
    else:
        assert grad_close(xd.grad.float().cpu(), gx, dtype=xdtype)


def test_projected_info_nce_weights_and_repeated_backward():
    """phase-1 weight 0.5 (P1:68, P1:351) through `upstream` and through a plain multiplication; a second backward through the
    same node re-emits both gradients into fresh buffers."""
    g = torch.Generator().manual_seed(77)
    xs = torch.randn(10, 6, 64, generator=g).to(torch.bfloat16)
    W = (torch.randn(64, 48, generator=g) * 0.1).to(torch.bfloat16)
    xt = torch.randn(10, 6, 48, generator=g).to(torch.bfloat16)
    _, gx, gW = _proj_reference(xs, xt, W)
    for use_upstream in (True, False):
        xd, Wd = xs.cuda().requires_grad_(True), W.cuda().requires_grad_(True)
        loss = kd.info_nce(xd, xt.cuda(), proj=Wd, upstream=0.5 if use_upstream else None)
        (loss * 0.5).backward(retain_graph=True)
        assert grad_close(xd.grad.float().cpu(), gx * 0.5) and grad_close(Wd.grad.float().cpu(), gW * 0.5)
        xd.grad = None
        Wd.grad = None
        (loss * 2.0).backward()                 # repeated backward, another weight
        assert grad_close(xd.grad.float().cpu(), gx * 2.0) and grad_close(Wd.grad.float().cpu(), gW * 2.0)
    # only the projection is trained (frozen features): no feature gradient is written
    Wd = W.cuda().requires_grad_(True)
    kd.info_nce(xs.cuda(), xt.cuda(), proj=Wd).backward()
    assert grad_close(Wd.grad.float().cpu(), gW)


def test_projected_info_nce_error_paths_through_the_c_abi():
    """Raw ctypes calls of `kd_info_nce_proj_*`: a short workspace, more rows than the CUDA-core form takes, a null pointer and a bad
    temperature are reported through the return code + kd_last_error(), nothing is launched."""
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib
    lib = _lib.load()
    M, P, Ds, Dt = 8, 3, 64, 32
    xs = torch.randn(M, P, Ds, device="cuda").to(torch.bfloat16)
    xt = torch.randn(M, P, Dt, device="cuda").to(torch.bfloat16)
    w = torch.randn(Ds, Dt, device="cuda").to(torch.bfloat16)
    gx, gw = torch.empty_like(xs), torch.empty_like(w)
    loss = torch.zeros((), device="cuda")
    need = _lib.size_query("kd_info_nce_proj_workspace_bytes", M, Ds, Dt)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def call(m=M, ws_bytes=need, temp=0.07, student=xs.data_ptr()):
        return lib.kd_info_nce_proj_fwd_bwd(student, P, xt.data_ptr(), P, w.data_ptr(), gx.data_ptr(), gw.data_ptr(), loss.data_ptr(),
                                            ws.data_ptr(), ws_bytes, m, Ds, Dt, _lib.DTYPE_BF16, _lib.DTYPE_BF16, temp, None, 0, st)
    assert call() == 0
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item()
    assert call(ws_bytes=need - 1) == 3 and b"workspace" in lib.kd_last_error()          # KD_ERR_WORKSPACE_TOO_SMALL
    assert call(m=5000) == 2 and b"kd_gemm" in lib.kd_last_error()                       # KD_ERR_UNSUPPORTED: use the tensor-core route
    assert call(temp=0.0) == 1                                                            # KD_ERR_INVALID_ARGUMENT
    assert call(student=None) == 1
    ratio = torch.ones((), device="cuda")
    assert lib.kd_info_nce_proj_regrad(ws.data_ptr(), need - 1, gx.data_ptr(), gw.data_ptr(), M, P, Ds, Dt, _lib.DTYPE_BF16,
                                       _lib.DTYPE_BF16, ratio.data_ptr(), None, st) == 3
    assert lib.kd_info_nce_proj_regrad(ws.data_ptr(), need, None, None, M, P, Ds, Dt, _lib.DTYPE_BF16, _lib.DTYPE_BF16,
                                       ratio.data_ptr(), None, st) == 1
    with pytest.raises(ValueError):
        kd.info_nce(xs, xt, proj=w.t().contiguous())                                      # [D_t, D_s]: wrong orientation


def test_mse_odd_sizes_and_unaligned():
    for n in (1, 7, 1025, 8191):
        a = torch.randn(n + 1, device="cuda", dtype=torch.bfloat16)[1:]     # 2-byte offset: not 16-B aligned
        b = torch.randn(n, device="cuda", dtype=torch.bfloat16)
        ref = O.feature_mse(a.cpu(), b.cpu())
        ad = a.clone().requires_grad_(True) if n % 2 else a.detach().requires_grad_(True)
        loss = kd.feature_mse(ad, b)
        loss.backward()
        assert abs(loss.item() - ref.loss.item()) <= 1e-5 * ref.loss.item()
        assert grad_close(ad.grad.float().cpu(), ref.grad)


def test_learned_projections_match_the_reference_applied_to_projected_features():
    """`proj` is not in the reference (both towers are SigLIP-1152): the oracle is the reference loss applied to x_s @ W."""
    g = torch.Generator().manual_seed(9)
    xs, xt = synthetic.vision_feature_batch(6, 19, 96, seed=4)
    W = (torch.randn(96, 160, generator=g) * 0.1)
    xt_wide = torch.randn(6, 19, 160, generator=g).to(torch.bfloat16)
    ref = O.info_nce((xs.double() @ W.double()), xt_wide)
    xd = xs.cuda().requires_grad_(True)
    Wd = W.cuda().requires_grad_(True)
    loss = kd.info_nce(xd, xt_wide.cuda(), proj=Wd)
    loss.backward()
    assert abs(loss.item() - ref.loss.item()) <= 2e-5 * abs(ref.loss.item())
    # gradients w.r.t. the features AND the projection (tcgen05 products of hi + lo parts) against fp64 autograd of the same formula
    x64, W64 = xs.double().requires_grad_(True), W.double().requires_grad_(True)
    us = torch.nn.functional.normalize((x64 @ W64).mean(1), dim=-1)
    ut = torch.nn.functional.normalize(xt_wide.double().mean(1), dim=-1)
    lg = us @ ut.T / 0.07
    l64 = (torch.logsumexp(lg, -1) - lg.diagonal()).mean()
    gx, gW = torch.autograd.grad(l64, (x64, W64))
    assert abs(l64.item() - ref.loss.item()) <= 1e-12
    assert rel_err(Wd.grad.cpu(), gW) <= 1e-4
    assert grad_close(xd.grad.float().cpu(), gx)
    # token-level 2-D bf16 features through a bf16 projection (the 1152 -> 896 case of BASELINE configs[2], at a size the oracle fits)
    xs2, xt2 = synthetic.vision_feature_batch(1536, 1, 256, seed=8)
    xs2, xt2 = xs2[:, 0].contiguous(), synthetic.vision_feature_batch(1536, 1, 128, seed=9)[1][:, 0].contiguous()
    Wb = (torch.randn(256, 128, generator=g) * 0.08).to(torch.bfloat16)
    y_ref = (xs2.double() @ Wb.double()).to(torch.bfloat16)          # what a bf16 GEMM hands to the loss
    ref2 = O.info_nce(y_ref.cuda(), xt2.cuda())
    x2d, Wbd = xs2.cuda().requires_grad_(True), Wb.cuda().requires_grad_(True)
    l2 = kd.info_nce(x2d, xt2.cuda(), proj=Wbd)
    l2.backward()
    assert abs(l2.item() - ref2.loss.item()) <= 1e-4 * abs(ref2.loss.item())      # bf16 rounding boundaries of the projected tensor
    gx2 = ref2.grad.double() @ Wb.double().cuda().T
    assert rel_err(x2d.grad.double(), gx2) <= 6e-3                                # two bf16 roundings (d projected, d features)
    assert torch.isfinite(Wbd.grad.float()).all()
    a = torch.randn(300, 48, generator=g).to(torch.bfloat16)
    b = torch.randn(300, 80, generator=g).to(torch.bfloat16)
    P = (torch.randn(48, 80, generator=g) * 0.2).to(torch.bfloat16)
    proj_ref = (a.float() @ P.float()).to(torch.bfloat16)            # what the bf16 GEMM hands to the loss
    ref = O.feature_mse(proj_ref, b)
    ad, Pd = a.cuda().requires_grad_(True), P.cuda().requires_grad_(True)
    loss = kd.feature_mse(ad, b.cuda(), proj=Pd)
    loss.backward()
    assert abs(loss.item() - ref.loss.item()) <= 1e-3 * ref.loss.item()      # cuBLAS bf16 accumulation order vs the fp32 product
    d_proj = ref.grad.float() @ P.float().t()
    assert ((ad.grad.float().cpu() - d_proj).norm() / d_proj.norm()).item() <= 1e-2
