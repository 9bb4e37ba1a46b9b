"""GPU parity of the fused LoCa KL + student cross-entropy: ``kd.loca_kl_ce`` through the C ABI against
the oracle's ``loca_kl`` + ``student_cross_entropy`` (fp64 on the GPU) — the sum the reference forms at P1:194 from the HF
student's loss (P1:238-240).  Gates: losses 1e-5 relative, bf16 gradient 2e-3 relative (tests/_util.grad_close)."""
import pytest
import torch

from _util import grad_close
from oracle import kd_oracle as O
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib, synthetic

pytestmark = pytest.mark.gpu


def _batch(n, V, seed, peaked, ignore_frac):
    t, s, lab = synthetic.logit_batch(n, V, seed=seed, peaked=peaked)
    g = torch.Generator().manual_seed(seed + 17)
    cel = torch.randint(0, V, (n,), generator=g)
    cel[torch.rand(n, generator=g) < ignore_frac] = -100
    if ignore_frac > 0:
        cel[0] = -100          # at least one ignored and one valid row
        cel[1] = 5
    return t.cuda(), s.cuda(), lab.cuda(), cel.cuda()


def _oracle(t, s, lab, cel, T, w):
    kl = O.loca_kl(t, s, lab, T)
    ce = O.student_cross_entropy(s, cel)
    return kl.loss + w * ce.loss, ce.loss, kl.grad + w * ce.grad.reshape(kl.grad.shape)


@pytest.mark.parametrize("n,V,T,peaked,ignore,w", [
    (64, synthetic.VOCAB, 2.0, False, 0.5, 1.0),     # BASELINE configs[1] temperature; half the rows are prompt / padding
    (48, synthetic.VOCAB, 0.8, True, 0.3, 1.0),      # the reference's T (P1:71): every row takes the clamped path
    (40, synthetic.VOCAB, 1.0, False, 0.0, 1.25),    # LG:75 T = 1: the CE shares the KD's exponentials; weight 1/gamma of phase 3
    (33, 40000, 2.0, False, 0.6, 0.5),               # another vocabulary, odd row count
])
def test_fused_ce_matches_kl_plus_cross_entropy(n, V, T, peaked, ignore, w):
    t, s, lab, cel = _batch(n, V, seed=n, peaked=peaked, ignore_frac=ignore)
    ref_loss, ref_ce, ref_grad = _oracle(t, s, lab, cel, T, w)
    sd = s.clone().requires_grad_(True)
    loss, ce = kd.loca_kl_ce(t, sd, lab, cel, T, ce_weight=w)
    loss.backward()
    assert abs(ce.item() - ref_ce.item()) <= 1e-5 * abs(ref_ce.item())
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert grad_close(sd.grad.float().cpu(), ref_grad.cpu())
    # the KD part alone is what loca_kl returns
    kl_only = kd.loca_kl(t, s, lab, T)
    assert abs((loss.item() - w * ce.item()) - kl_only.item()) <= 1e-5 * abs(kl_only.item()) + 1e-6 * abs(w * ce.item())


def test_fused_ce_backward_with_another_grad_output_and_repeated_backward():
    t, s, lab, cel = _batch(24, synthetic.VOCAB, seed=5, peaked=False, ignore_frac=0.4)
    _, _, ref_grad = _oracle(t, s, lab, cel, 2.0, 1.0)
    sd = s.clone().requires_grad_(True)
    loss, _ = kd.loca_kl_ce(t, sd, lab, cel, 2.0)
    (0.25 * loss).backward(retain_graph=True)                 # re-emission kernel, CE terms included
    assert grad_close(sd.grad.float().cpu(), 0.25 * ref_grad.cpu())
    sd.grad = None
    loss.backward()                                           # second backward through the same node: fresh buffer
    assert grad_close(sd.grad.float().cpu(), ref_grad.cpu())
    # phase 3 (P1:257-260): gamma * (loca + CE) + (1 - gamma) * CE = gamma * (loca + CE / gamma), gamma folded in as `upstream`
    gamma = 0.8
    _, _, ref3 = _oracle(t, s, lab, cel, 2.0, 1.0 / gamma)
    s3 = s.clone().requires_grad_(True)
    l3, _ = kd.loca_kl_ce(t, s3, lab, cel, 2.0, ce_weight=1.0 / gamma, upstream=gamma)
    (gamma * l3).backward()
    assert grad_close(s3.grad.float().cpu(), gamma * ref3.cpu())


def test_fused_ce_label_errors_and_shapes():
    t, s, lab, cel = _batch(16, synthetic.VOCAB, seed=9, peaked=False, ignore_frac=0.3)
    bad = cel.clone()
    bad[3] = synthetic.VOCAB + 7                              # F.cross_entropy raises on it; here: status bit + NaN, no host sync
    loss, ce = kd.loca_kl_ce(t, s.clone().requires_grad_(True), lab, bad, 2.0)
    assert torch.isnan(loss).item() and torch.isnan(ce).item()
    with pytest.raises(_lib.KDError):                         # no fused kernel for fp32 logits: the caller adds its own CE
        kd.loca_kl_ce(t.float(), s.float(), lab, cel, 2.0)
    shifted = kd.shift_labels(torch.tensor([[1, 2, 3], [4, -100, 6]]))
    assert shifted.tolist() == [[2, 3, -100], [-100, 6, -100]]


def test_patched_reference_method_fuses_the_ce_when_student_loss_is_none():
    """compute_loca_loss(teacher, student, None, labels) on a patched reference-shaped object == loca + the HF student's loss
    (P1:194 with P1:238-240's shifted, ignore_index=-100 cross-entropy), labels [B, S] as the reference passes them."""
    class Ref:
        T = 2.0

        def compute_loca_loss(self, teacher_logits, student_logits, student_loss, labels, alpha=0.8):
            raise AssertionError("not patched")
    obj = Ref()
    kd.patch_reference(obj)
    B, S, V = 2, 12, 40000
    t, s, lab = synthetic.logit_batch(B * S, V, seed=21)
    t, s, lab = t.cuda().view(B, S, V), s.cuda().view(B, S, V), lab.cuda().view(B, S)
    sd = s.clone().requires_grad_(True)
    total = obj.compute_loca_loss(t, sd, None, lab)
    total.backward()
    kl = O.loca_kl(t, s, lab, 2.0)
    hf = torch.nn.functional.cross_entropy(s[:, :-1].double().reshape(-1, V), lab[:, 1:].reshape(-1), ignore_index=-100)   # HF's shift
    ce = O.student_cross_entropy(s, kd.shift_labels(lab))
    assert abs(ce.loss.item() - hf.item()) <= 1e-12 * abs(hf.item())
    assert abs(total.item() - (kl.loss + ce.loss).item()) <= 1e-5 * abs((kl.loss + ce.loss).item())
    assert abs(obj.last_student_loss.item() - hf.item()) <= 1e-5 * abs(hf.item())
    assert grad_close(sd.grad.float().cpu(), (kl.grad.reshape(B, S, V) + ce.grad).cpu())


def test_token_sharded_emit_with_fused_ce_matches_the_single_call():
    """kd_loca_teacher_stats -> kd_loca_build_table -> kd_loca_emit_ce (the token-sharded building blocks with the fused CE):
    one shard == kd_logit_kd_ce_fwd_bwd bit for bit; two simulated shards (records concatenated in global token order, the CE's
    divisor = valid targets of both) sum to the full batch's losses and reproduce its gradient rows."""
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded
    n, V, T, w = 37, synthetic.VOCAB, 2.0, 1.25
    t, s, lab, cel = _batch(n, V, seed=41, peaked=False, ignore_frac=0.4)
    ref_loss, ref_ce, ref_grad = _oracle(t, s, lab, cel, T, w)
    sd = s.clone().requires_grad_(True)
    one_loss, one_ce = kd.loca_kl_ce(t, sd, lab, cel, T, ce_weight=w)
    one_loss.backward()
    loss, ce, grad = sharded.loca_kl_token_sharded(t, s, lab, T, ce_labels=cel, ce_weight=w)
    assert loss.item() == one_loss.item() and ce.item() == one_ce.item()
    assert torch.equal(grad, sd.grad)
    # two shards, ragged on purpose
    ph = sharded.CudaPhases()
    cut = 14
    parts = [(t[:cut], s[:cut], lab[:cut], cel[:cut]), (t[cut:], s[cut:], lab[cut:], cel[cut:])]
    stats = [ph.teacher_stats(tt, ll, T, 0.8, V) for tt, _, ll, _ in parts]
    table, consts = ph.build_table(torch.cat([st[0] for st in stats]), V)
    inv_count = (1.0 / (cel != -100).sum().float()).reshape(1)
    tot, tot_ce, grads = 0.0, 0.0, []
    for (tt, ss, _, cc), (_, lse2, status) in zip(parts, stats):
        l, c, g = ph.emit_ce(tt, ss, lse2, table, consts, cc, inv_count, w, T, 1e-8, 1.0 / (n * V), None, status=status)
        tot, tot_ce = tot + l.item(), tot_ce + c.item()
        grads.append(g)
    assert abs(tot - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    assert abs(tot_ce - ref_ce.item()) <= 1e-5 * abs(ref_ce.item())
    assert grad_close(torch.cat(grads).float().cpu(), ref_grad.cpu())
    # through autograd with another grad_output (the local gradient is rescaled on the device)
    s2 = s.clone().requires_grad_(True)
    l2, _ = sharded.loca_kl_ce_token_sharded_autograd(t, s2, lab, cel, T, ce_weight=w)
    (0.5 * l2).backward(retain_graph=True)                   # kd_loca_regrad: re-emitted from the saved row statistics, ONE rounding
    assert grad_close(s2.grad.float().cpu(), 0.5 * ref_grad.cpu())
    s2.grad = None
    l2.backward()                                            # second backward through the same node: fresh buffer
    assert grad_close(s2.grad.float().cpu(), ref_grad.cpu())
    # without the CE: the gamma of phase 3 (P1:260) arrives as grad_output; with upstream given the forward's gradient is final
    kl = O.loca_kl(t, s, lab, T)
    s3 = s.clone().requires_grad_(True)
    (0.8 * sharded.loca_kl_token_sharded_autograd(t, s3, lab, T)).backward()
    assert grad_close(s3.grad.float().cpu(), 0.8 * kl.grad.cpu())
    s4 = s.clone().requires_grad_(True)
    (0.8 * sharded.loca_kl_token_sharded_autograd(t, s4, lab, T, upstream=0.8)).backward()
    assert grad_close(s4.grad.float().cpu(), 0.8 * kl.grad.cpu())
    # the re-emission sweep and the forward launch evaluate the same cells in another order: only roundings on a boundary differ
    assert ((s3.grad.float() - s4.grad.float()).norm() / s4.grad.float().norm()).item() <= 1e-3


def test_token_sharded_fused_ce_bad_target_and_unsupported_shape():
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded
    t, s, lab, cel = _batch(12, synthetic.VOCAB, seed=43, peaked=False, ignore_frac=0.3)
    bad = cel.clone()
    bad[5] = synthetic.VOCAB + 1
    loss, ce, _ = sharded.loca_kl_token_sharded(t, s, lab, 2.0, ce_labels=bad)
    assert torch.isnan(loss).item() and torch.isnan(ce).item()
    with pytest.raises(_lib.KDError):                          # fp32 logits: no fused kernel, the caller adds its own CE
        sharded.loca_kl_token_sharded(t.float(), s.float(), lab, 2.0, ce_labels=cel)
    with pytest.raises(ValueError):
        sharded.loca_kl_token_sharded(t, s, lab, 2.0, ce_labels=cel[:-1])


def test_phase_split_ce_entry_points_error_paths_through_the_c_abi():
    """kd_loca_emit_ce / kd_loca_regrad called raw: null pointers and a short leading dimension are KD_ERR_INVALID_ARGUMENT (1) with a
    message; fp32 logits are KD_ERR_UNSUPPORTED (2) for the fused cross-entropy — nothing is launched in either case."""
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.functional import _stream_ptr
    lib = _lib.load()
    n, V, T = 8, 40000, 2.0
    t, s, lab, cel = _batch(n, V, seed=3, peaked=False, ignore_frac=0.3)
    ph = sharded.CudaPhases()
    rec, lse2, status = ph.teacher_stats(t, lab, T, 0.8, V)
    table, consts = ph.build_table(rec, V)
    dev = s.device
    loss, ce = torch.empty((), device=dev), torch.empty((), device=dev)
    rowaux, ceaux = torch.empty((n + 1, 4), device=dev), torch.empty((n, 2), device=dev)
    grad = torch.empty_like(s)
    inv = torch.full((1,), 0.2, device=dev)
    st = _stream_ptr(dev)

    def emit(ce_labels=cel.data_ptr(), status_ptr=status.data_ptr(), ld_g=V, dtype=0, tt=t, ss=s):
        return lib.kd_loca_emit_ce(tt.data_ptr(), tt.stride(0), ss.data_ptr(), ss.stride(0), lse2.data_ptr(), table.data_ptr(),
                                   consts.data_ptr(), ce_labels, inv.data_ptr(), 1.0, grad.data_ptr(), ld_g, loss.data_ptr(), ce.data_ptr(),
                                   rowaux.data_ptr(), ceaux.data_ptr(), n, V, dtype, T, 1e-8, 1.0 / (n * V), None, status_ptr, 0, st)
    assert emit() == 0
    assert emit(ce_labels=None) == 1 and b"kd_loca_emit_ce" in lib.kd_last_error()
    assert emit(status_ptr=None) == 1
    assert emit(ld_g=V - 8) == 1 and b"leading dimension" in lib.kd_last_error()
    assert emit(dtype=2, tt=t.float(), ss=s.float()) == 2 and b"cross-entropy" in lib.kd_last_error()
    torch.cuda.synchronize()
    up = torch.full((1,), 0.5, device=dev)

    def regrad(table_ptr=table.data_ptr(), ce_labels=cel.data_ptr(), ceaux_ptr=ceaux.data_ptr()):
        return lib.kd_loca_regrad(t.data_ptr(), t.stride(0), s.data_ptr(), s.stride(0), table_ptr, rowaux.data_ptr(), ce_labels,
                                  inv.data_ptr(), 1.0, ceaux_ptr, grad.data_ptr(), V, n, V, 0, T, 1e-8, 1.0 / (n * V), up.data_ptr(), None, st)
    g_fwd = grad.clone()
    assert regrad() == 0
    torch.cuda.synchronize()
    assert ((grad.float() - 0.5 * g_fwd.float()).norm() / (0.5 * g_fwd.float()).norm()).item() <= 1e-3     # half the forward's gradient
    assert regrad(table_ptr=None) == 1 and b"kd_loca_regrad" in lib.kd_last_error()
    assert regrad(ceaux_ptr=None) == 1
