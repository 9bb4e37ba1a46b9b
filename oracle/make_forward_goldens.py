"""TEST INFRASTRUCTURE ONLY — goldens for the reference's loss MIXING, from the IMPORTED reference methods.

Run in the dev container (needs /root/reference):  ``python -m oracle.make_forward_goldens``
Writes ``tests/golden/forward_*.npz``.  Each case calls the reference class's own methods with its own hyper-parameters
(P1:67-71, LG:73-75, FB:72-74) in float64 on bf16-rounded inputs; the only lines restated here are the ones of ``forward``
that need no model: the pooling + normalisation of the hooked vision features (P1:243-248, FB:153-158) and the phase switch
(P1:250-260).  Gradients w.r.t. the student logits, the raw student vision tokens and the student CE scalar come from autograd.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", message=".*reduction: 'mean' divides.*")

HYPER = {   # the attributes each reference __init__ sets
    "P1": dict(T=0.8, soft_target_loss_weight=0.1, ce_loss_weight=0.5, gamma=0.8),      # P1:67-71
    "LG": dict(T=1.0, soft_target_loss_weight=0.1, ce_loss_weight=0.5),                  # LG:73-75
    "FB": dict(T=0.8, soft_target_loss_weight=0.1, ce_loss_weight=0.8),                  # FB:72-74
}


def instance(which: str, **extra):
    """A stand-in for the LightningModule carrying the reference's hyper-parameters, with the reference's methods bound."""
    K = R.reference_class(which)
    me = types.SimpleNamespace(**HYPER[which], **extra)
    for name in ("compute_loca_loss", "compute_vision_loss", "compute_loss", "compute_language_loss", "contrastive_loss"):
        if hasattr(K, name):
            setattr(me, name, types.MethodType(getattr(K, name), me))
    return me


def inputs(seed, B=2, S=6, V=640, Vt=700, Bp=3, P=5, D=64):
    t, s, lab = synthetic.logit_batch(B * S, V, seed=seed, teacher_vocab=Vt)
    xs, xt = synthetic.vision_feature_batch(Bp, P, D, seed=seed + 1)
    return dict(teacher_logits=t.double().reshape(B, S, Vt), student_logits=s.double().reshape(B, S, V),
                labels=lab.reshape(B, S), student_tokens=xs.double(), teacher_tokens=xt.double(),
                student_ce=torch.tensor(1.0 + 0.03125 * seed, dtype=torch.float64))


def pooled(x):          # P1:243-248
    return F.normalize(x.mean(dim=1), p=2, dim=-1)


def case(name, which, seed, fn):
    d = inputs(seed)
    sl = d["student_logits"].clone().requires_grad_(True)
    xs = d["student_tokens"].clone().requires_grad_(True)
    ce = d["student_ce"].clone().requires_grad_(True)
    outs = types.SimpleNamespace(logits=sl, loss=ce)
    loss = fn(instance(which), d, outs, pooled(xs), pooled(d["teacher_tokens"]))
    grads = torch.autograd.grad(loss, [sl, xs, ce], allow_unused=True)
    z = lambda g, like: (torch.zeros_like(like) if g is None else g).numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="forward", which=which,
                        teacher_logits=d["teacher_logits"].numpy().astype(np.float32),
                        student_logits=d["student_logits"].numpy().astype(np.float32), labels=d["labels"].numpy(),
                        student_tokens=d["student_tokens"].numpy().astype(np.float32),
                        teacher_tokens=d["teacher_tokens"].numpy().astype(np.float32), student_ce=d["student_ce"].numpy(),
                        loss=loss.detach().numpy(), grad_logits=z(grads[0], sl), grad_tokens=z(grads[1], xs), grad_ce=z(grads[2], ce),
                        **{k: v for k, v in HYPER[which].items()})
    print(f"{name}: loss={loss.item():.12e}")


def phase_switch(phase):            # P1:250-260
    def run(me, d, outs, fs, ft):
        if phase == 1:
            return me.compute_vision_loss(fs, ft, d["teacher_logits"], outs)
        if phase == 2:
            return me.compute_loca_loss(d["teacher_logits"], outs.logits, outs.loss, d["labels"])
        loca = me.compute_loca_loss(d["teacher_logits"], outs.logits, outs.loss, d["labels"])
        return me.gamma * loca + (1 - me.gamma) * outs.loss
    return run


def main():
    os.makedirs(OUT, exist_ok=True)
    case("forward_p1_phase1", "P1", 51, phase_switch(1))
    case("forward_p1_phase2", "P1", 52, phase_switch(2))
    case("forward_p1_phase3", "P1", 53, phase_switch(3))
    case("forward_p1_language", "P1", 54, lambda me, d, outs, fs, ft: me.compute_language_loss(d["teacher_logits"], outs))     # P1:356-390
    case("forward_lg_loca", "LG", 55, lambda me, d, outs, fs, ft: me.compute_loca_loss(d["teacher_logits"], outs.logits, outs.loss, d["labels"]))  # LG:165
    case("forward_fb", "FB", 56, lambda me, d, outs, fs, ft: me.compute_loss(d["teacher_logits"], outs, me.contrastive_loss(fs, ft)))   # FB:162-165


if __name__ == "__main__":
    main()
