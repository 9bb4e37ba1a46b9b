"""Pin the oracle against the reference's OWN functions imported from /root/reference.

Runs only in the dev container (the reference tree does not travel to the GPU box); the frozen
goldens in tests/golden carry the same pin everywhere else.
"""
import warnings

import pytest
import torch
from hypothesis import given, settings, strategies as st

from _util import rel_err
from oracle import kd_oracle as O
from oracle import ref_import as R
from oracle import reference_port as P
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference not present")
warnings.filterwarnings("ignore", message=".*reduction: 'mean' divides.*")


def _ref_loca(t, s, lab, T, alpha, which="P1"):
    sv = s.clone().requires_grad_(True)
    loss = R.loca_loss(t, sv, lab, T, alpha, which=which)
    (g,) = torch.autograd.grad(loss, sv)
    return loss.detach(), g


@settings(max_examples=25, deadline=None)
@given(B=st.integers(1, 3), S=st.integers(1, 7), V=st.integers(5, 300), T=st.sampled_from([0.8, 1.0, 2.0]),
       alpha=st.sampled_from([0.5, 0.8]), seed=st.integers(0, 10_000), boost=st.sampled_from([1.0, 5.0]),
       which=st.sampled_from(["P1", "LG"]))
def test_loca_reference_mode(B, S, V, T, alpha, seed, boost, which):
    t, s, lab = synthetic.logit_batch(B * S, V, seed=seed)
    t = t.double().reshape(B, S, V)
    s = (s.double() * boost).reshape(B, S, V)
    lab = lab.reshape(B, S)
    loss, grad = _ref_loca(t, s, lab, T, alpha, which)
    res = O.loca_kl(t, s, lab, T, alpha, mode=O.LOCA_REFERENCE, chunk_rows=3)
    assert abs(res.loss.item() - loss.item()) <= 1e-11 * abs(loss.item())
    assert rel_err(res.grad, grad) <= 1e-10
    # eager port: same ops -> same bits
    sv = s.clone().requires_grad_(True)
    pl = P.loca_loss_eager(t, sv, lab, T, alpha)
    assert pl.item() == loss.item()


@settings(max_examples=10, deadline=None)
@given(N=st.integers(1, 6), V=st.integers(5, 200), T=st.sampled_from([0.8, 2.0]), seed=st.integers(0, 10_000))
def test_loca_per_row_mode_is_reference_applied_per_token(N, V, T, seed):
    t, s, lab = synthetic.logit_batch(N, V, seed=seed)
    t, s = t.double(), s.double()
    sv = s.clone().requires_grad_(True)
    tot = 0.0
    for r in range(N):
        tot = tot + R.loca_loss(t[r][None, None], sv[r][None, None], lab[r][None, None], T, 0.8)
    tot = tot / N
    (g,) = torch.autograd.grad(tot, sv)
    res = O.loca_kl(t, s, lab, T, 0.8, mode=O.LOCA_PER_ROW)
    assert abs(res.loss.item() - tot.item()) <= 1e-11 * abs(tot.item())
    assert rel_err(res.grad, g) <= 1e-10


def test_reference_rejects_negative_labels_and_so_does_the_oracle():
    t, s, lab = synthetic.logit_batch(4, 50, seed=3)
    lab[1] = -100
    with pytest.raises((IndexError, RuntimeError)):
        R.loca_loss(t.float()[None], s.float()[None], lab[None], 2.0)
    with pytest.raises(IndexError):
        O.loca_kl(t, s, lab, 2.0)


@settings(max_examples=10, deadline=None)
@given(B=st.integers(1, 3), S=st.integers(1, 5), V=st.integers(4, 150), T=st.sampled_from([0.8, 2.0]),
       seed=st.integers(0, 10_000), quirk=st.booleans())
def test_temperature_kl(B, S, V, T, seed, quirk):
    t, s, _ = synthetic.logit_batch(B * S, V, seed=seed)
    t = t.double().reshape(B, S, V)
    s = s.double().reshape(B, S, V)
    sv = s.clone().requires_grad_(True)
    z = torch.zeros((), dtype=torch.float64)
    if quirk:
        loss = R.feature_based_loss(t, sv, z, z, T, soft_w=1.0, ce_w=0.0)
    else:
        f = torch.zeros(2, 3, dtype=torch.float64)
        loss = R.vision_loss(f, f, t, sv, T, soft_w=1.0, ce_w=0.0)
    (g,) = torch.autograd.grad(loss, sv)
    res = O.temperature_kl(t, s, T, log_target_quirk=quirk)
    assert abs(res.loss.item() - loss.item()) <= 1e-11 * abs(loss.item())
    assert rel_err(res.grad, g) <= 1e-10


def test_batchmean_variant_is_a_normaliser_change():
    """EV:130-136 kl_divergence_loss(batchmean) == sum / size(0)."""
    t, s, _ = synthetic.logit_batch(6, 40, seed=5)
    t, s = t.double().reshape(2, 3, 40), s.double().reshape(2, 3, 40)
    K = R.reference_class("EV")
    sp = torch.softmax(s / 2.0, -1)
    tp = torch.softmax(t / 2.0, -1)
    ref = K.kl_divergence_loss(R.fake_self(), sp, tp, 2.0)
    res = O.temperature_kl(t, s, 2.0, inv_norm=1.0 / 2)
    assert abs(res.loss.item() - ref.item()) <= 1e-11 * abs(ref.item())


def test_contrastive_and_mse():
    xs, xt = synthetic.vision_feature_batch(5, 4, 32, seed=9)
    xs64 = xs.double().requires_grad_(True)
    fs = torch.nn.functional.normalize(xs64.mean(1), dim=-1)
    ft = torch.nn.functional.normalize(xt.double().mean(1), dim=-1)
    loss = R.contrastive_loss(fs, ft)
    (g,) = torch.autograd.grad(loss, xs64)
    res = O.info_nce(xs, xt)
    assert abs(res.loss.item() - loss.item()) <= 1e-10 * max(abs(loss.item()), 1e-9)
    assert rel_err(res.grad, g) <= 1e-9
    a, b = torch.randn(7, 9), torch.randn(7, 9)
    assert abs(O.feature_mse(a, b).loss.item() - R.mse_loss(a, b).item()) <= 1e-6
