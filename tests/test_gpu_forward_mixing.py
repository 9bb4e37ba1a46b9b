"""GPU parity of the loss MIXING through the patched reference interface.

The goldens (``oracle/make_forward_goldens.py``) come from the imported reference class: ``compute_vision_loss`` (P1:316-354),
``compute_loca_loss`` (P1:141-194, LG:208-261), ``compute_language_loss`` (P1:356-390), ``compute_loss`` (FB:191-230) and the
phase switch of ``forward`` (P1:250-260), with the reference's own hyper-parameters.  Here a class with the reference's
method names (all raising) is patched with ``patch_reference`` and driven exactly like the reference's ``forward`` drives it;
``KDLoss`` / ``combine_phase`` are checked against the same goldens."""
import types

import pytest
import torch
import torch.nn.functional as F

from _util import goldens_of_kind, grad_close, load_golden, rel_err
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5


class ReferenceShaped:
    """The reference module's loss interface with nothing behind it: every method must be replaced by patch_reference."""

    def __init__(self, **hyper):
        for k, v in hyper.items():
            setattr(self, k, v)

    def compute_loca_loss(self, teacher_logits, student_logits, student_loss, labels, alpha=0.8):
        raise AssertionError("not patched")

    def compute_vision_loss(self, student_features, teacher_features, teacher_logits, student_outputs):
        raise AssertionError("not patched")

    def compute_loss(self, teacher_logits, student_outputs, contrastive_loss_value):
        raise AssertionError("not patched")

    def compute_language_loss(self, teacher_logits, student_outputs):
        raise AssertionError("not patched")

    def contrastive_loss(self, student_features, teacher_features, temperature=0.07):
        raise AssertionError("not patched")

    def forward_tail(self, teacher_logits, student_outputs, labels, student_tokens, teacher_tokens):
        """P1:243-260 verbatim in structure: pool + normalise the hooked features, then the phase switch."""
        teacher_features = F.normalize(teacher_tokens.mean(dim=1), p=2, dim=-1)
        student_features = F.normalize(student_tokens.mean(dim=1), p=2, dim=-1)
        if self.phase == 1:
            return self.compute_vision_loss(student_features, teacher_features, teacher_logits, student_outputs)
        if self.phase == 2:
            return self.compute_loca_loss(teacher_logits, student_outputs.logits, student_outputs.loss, labels)
        loca_loss = self.compute_loca_loss(teacher_logits, student_outputs.logits, student_outputs.loss, labels)
        return self.gamma * loca_loss + (1 - self.gamma) * student_outputs.loss


def _load(name, logit_dtype=torch.bfloat16):
    g = load_golden(name)
    hyper = {k: float(g[k]) for k in ("T", "soft_target_loss_weight", "ce_loss_weight", "gamma") if k in g}
    t = torch.from_numpy(g["teacher_logits"]).to("cuda", logit_dtype)
    s = torch.from_numpy(g["student_logits"]).to("cuda", logit_dtype).requires_grad_(True)
    ce = torch.tensor(float(g["student_ce"]), dtype=torch.float64, device="cuda", requires_grad=True)
    xs = torch.from_numpy(g["student_tokens"]).to("cuda", torch.float32).requires_grad_(True)
    xt = torch.from_numpy(g["teacher_tokens"]).to("cuda", torch.float32)
    lab = torch.from_numpy(g["labels"]).cuda()
    return g, hyper, t, s, ce, xs, xt, lab


def _check(g, loss, s, ce, xs=None, token_tol=2e-4):
    # the CE scalar dominates some totals (2.6 against a KD term of 0.03): compare the part our kernels produce
    want_kd = float(g["loss"]) - float(g["grad_ce"]) * float(g["student_ce"])
    got_kd = loss.double().item() - float(g["grad_ce"]) * float(g["student_ce"])
    assert abs(got_kd - want_kd) <= LOSS_RTOL * abs(want_kd) + 1e-12, (got_kd, want_kd)
    assert abs(loss.double().item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    loss.backward()
    assert grad_close(s.grad.float().cpu(), g["grad_logits"])
    want_ce = float(g["grad_ce"])
    got_ce = 0.0 if ce.grad is None else ce.grad.item()
    assert abs(got_ce - want_ce) <= 1e-12
    if xs is not None:
        want = torch.from_numpy(g["grad_tokens"])
        if want.abs().max() == 0:
            assert xs.grad is None or xs.grad.abs().max().item() == 0.0
        else:
            assert rel_err(xs.grad.float().cpu(), want) <= token_tol


@pytest.mark.parametrize("phase", [1, 2, 3])
def test_patched_forward_tail_matches_reference(phase):
    g, hyper, t, s, ce, xs, xt, lab = _load(f"forward_p1_phase{phase}")
    mod = ReferenceShaped(phase=phase, **hyper)
    assert set(kd.patch_reference(mod)) >= {"compute_loca_loss", "compute_vision_loss", "contrastive_loss"}
    outs = types.SimpleNamespace(logits=s, loss=ce)
    _check(g, mod.forward_tail(t, outs, lab, xs, xt), s, ce, xs)


@pytest.mark.parametrize("phase", [1, 2, 3])
def test_kdloss_module_and_combine_phase(phase):
    """KDLoss.forward == combine_phase == P1:250-260; raw [B', tokens, D] features take the fused pooling path."""
    g, hyper, t, s, ce, xs, xt, lab = _load(f"forward_p1_phase{phase}")
    mod = kd.KDLoss(phase=phase, **hyper)
    outs = types.SimpleNamespace(logits=s, loss=ce)
    xs16 = xs.detach().to(torch.bfloat16).requires_grad_(True)            # goldens hold bf16-rounded tokens: exact
    loss = mod(t, outs, lab, student_features=xs16, teacher_features=xt.to(torch.bfloat16))
    _check(g, loss, s, ce)
    if phase == 1:
        assert grad_close(xs16.grad.float().cpu(), g["grad_tokens"])


def test_patched_logit_based_module():
    """LG:165 — compute_loca_loss at T = 1 with the wider teacher vocabulary (LG:222 slices it)."""
    g, hyper, t, s, ce, xs, xt, lab = _load("forward_lg_loca")
    mod = ReferenceShaped(**hyper)
    kd.patch_reference(mod)
    _check(g, mod.compute_loca_loss(t, s, ce, lab), s, ce)


def test_patched_language_loss():
    """P1:356-390 — sum / size(0) normaliser + ce_loss_weight * CE."""
    g, hyper, t, s, ce, xs, xt, lab = _load("forward_p1_language")
    mod = ReferenceShaped(**hyper)
    kd.patch_reference(mod)
    _check(g, mod.compute_language_loss(t, types.SimpleNamespace(logits=s, loss=ce)), s, ce)


def test_patched_feature_based_module():
    """FB:153-165 — pooled contrastive value, then compute_loss with the log_target quirk (FB:206-219) and its weights (FB:227)."""
    g, hyper, t, s, ce, xs, xt, lab = _load("forward_fb")
    mod = ReferenceShaped(**hyper)
    kd.patch_reference(mod)
    fs, ft = F.normalize(xs.mean(dim=1), p=2, dim=-1), F.normalize(xt.mean(dim=1), p=2, dim=-1)
    loss = mod.compute_loss(t, types.SimpleNamespace(logits=s, loss=ce), mod.contrastive_loss(fs, ft))
    _check(g, loss, s, ce, xs)


def test_retain_graph_backward_twice():
    """A second backward through the same graph must return the gradient again (it silently returned None)."""
    g, hyper, t, s, ce, xs, xt, lab = _load("forward_p1_phase2")
    cases = [(lambda: kd.loca_kl(t, s, lab, 0.8), s), (lambda: kd.loca_kl(t, s, lab, 0.8, mode="per_row"), s),
             (lambda: kd.temperature_kl(t, s, 0.8), s), (lambda: kd.feature_mse(s, t[..., : s.shape[-1]]), s),
             (lambda: kd.info_nce(xs, xt), xs)]
    for fn, leaf in cases:
        leaf.grad = None
        loss = fn()
        loss.backward(retain_graph=True)
        first = leaf.grad.clone()
        leaf.grad = None
        (loss * 2.0).backward()
        assert leaf.grad is not None and rel_err(leaf.grad.float(), 2.0 * first.float()) <= 2.0 ** -7
