"""TEST INFRASTRUCTURE ONLY — freeze golden vectors from the IMPORTED reference functions.

Run in the dev container (needs /root/reference):  ``python -m oracle.make_goldens``
Writes ``tests/golden/*.npz``.  Every output below comes from calling the reference's own
methods (``oracle/ref_import.py``) in float64 on bf16-rounded inputs — NOT from our
restatement — with ``torch.autograd`` providing the gradient w.r.t. the student tensor.
The full-vocabulary case regenerates its inputs from ``synthetic.logit_batch`` (seeded) and
stores a SHA-256 of them plus a strided sample of the gradient to stay small.
"""
from __future__ import annotations

import hashlib
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", message=".*reduction: 'mean' divides.*")


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float64)


def sha(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.contiguous().view(torch.uint8).numpy().tobytes() if t.dtype != torch.bfloat16
                 else t.contiguous().view(torch.int16).numpy().tobytes())
    return h.hexdigest()


def run_loca(teacher, student, labels, T, alpha, which="P1"):
    s = student.clone().requires_grad_(True)
    loss = R.loca_loss(teacher, s, labels, T, alpha, which=which)
    (g,) = torch.autograd.grad(loss, s)
    return loss.detach(), g


def run_loca_per_row(teacher, student, labels, T, alpha):
    """Reference applied to each [1,1,V] token, averaged over tokens (SURVEY §8c oracle 2)."""
    B, S, V = student.shape
    s = student.clone().requires_grad_(True)
    tot = 0.0
    for b in range(B):
        for i in range(S):
            tot = tot + R.loca_loss(teacher[b:b + 1, i:i + 1], s[b:b + 1, i:i + 1], labels[b:b + 1, i:i + 1], T, alpha)
    loss = tot / (B * S)
    (g,) = torch.autograd.grad(loss, s)
    return loss.detach(), g


def logit_case(name, B, S, V, T, alpha, seed, *, Vt=None, sigma=2.0, peaked=False, student_boost=0.0,
               dup_labels=False, label_is_runner_up=False, which="P1", per_row=False):
    t, s, lab = synthetic.logit_batch(B * S, V, seed=seed, sigma=sigma, peaked=peaked, teacher_vocab=Vt)
    if student_boost:   # make the student sharply peaked so that most sp < 1e-8 (clamp active)
        s = (s.float() * student_boost).to(torch.bfloat16)
    if dup_labels:
        lab[1::2] = lab[0]
    if label_is_runner_up:
        k = t[:, :V].float().topk(2, dim=-1).indices[:, 1]
        lab[::3] = k[::3]
    t64, s64 = t.to(torch.float64).reshape(B, S, -1), s.to(torch.float64).reshape(B, S, V)
    lab = lab.reshape(B, S)
    fn = run_loca_per_row if per_row else (lambda *a: run_loca(*a, which=which))
    loss, grad = fn(t64, s64, lab, T, alpha)
    loss32, _ = fn(t64.float(), s64.float(), lab, T, alpha)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="loca_per_row" if per_row else "loca_reference",
                        teacher=t64.numpy().astype(np.float32), student=s64.numpy().astype(np.float32),
                        labels=lab.numpy(), T=T, alpha=alpha, loss=loss.numpy(), loss_fp32=loss32.numpy(),
                        grad=grad.numpy())
    print(f"{name}: loss={loss.item():.12e} (fp32 run {loss32.item():.9e}) |grad|={grad.norm().item():.6e}")


def kl_case(name, B, S, V, T, seed, quirk, Vt=None):
    t, s, _ = synthetic.logit_batch(B * S, V, seed=seed, teacher_vocab=Vt)
    t64, s64 = t.to(torch.float64).reshape(B, S, -1), s.to(torch.float64).reshape(B, S, V)
    sv = s64.clone().requires_grad_(True)
    feats = torch.zeros(2, 4, dtype=torch.float64)
    if quirk:   # FB:191-230 with CE = 0 and contrastive = 0: loss = 0.1 * KL
        loss = R.feature_based_loss(t64, sv, torch.zeros((), dtype=torch.float64),
                                    torch.zeros((), dtype=torch.float64), T, soft_w=1.0, ce_w=0.0)
    else:       # P1:316-354 with soft weight 1 and contrastive weight 0
        loss = R.vision_loss(feats, feats, t64, sv, T, soft_w=1.0, ce_w=0.0)
    (g,) = torch.autograd.grad(loss, sv)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="kl_log_target" if quirk else "kl_plain",
                        teacher=t64.numpy().astype(np.float32), student=s64.numpy().astype(np.float32),
                        T=T, loss=loss.detach().numpy(), grad=g.numpy())
    print(f"{name}: loss={loss.item():.12e} |grad|={g.norm().item():.6e}")


def contrastive_case(name, M, P, D, seed, temperature=0.07, independent=False):
    xs, xt = synthetic.vision_feature_batch(M, P, D, seed=seed)
    if independent:   # unrelated teacher: loss ~ log(M), all softmax entries matter
        xt = synthetic.vision_feature_batch(M, P, D, seed=seed + 1000)[0]
    xs64 = xs.to(torch.float64).requires_grad_(True)
    xt64 = xt.to(torch.float64)
    # caller-side pooling exactly as P1:243-248, then the reference loss
    fs = torch.nn.functional.normalize(xs64.mean(dim=1), p=2, dim=-1)
    ft = torch.nn.functional.normalize(xt64.mean(dim=1), p=2, dim=-1)
    loss = R.contrastive_loss(fs, ft, temperature)
    (g,) = torch.autograd.grad(loss, xs64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="info_nce",
                        student=xs64.detach().numpy().astype(np.float32), teacher=xt64.numpy().astype(np.float32),
                        temperature=temperature, loss=loss.detach().numpy(), grad=g.numpy())
    print(f"{name}: loss={loss.item():.12e} |grad|={g.norm().item():.6e}")


def mse_case(name, n, da, seed):
    g_ = torch.Generator().manual_seed(seed)
    a = bf16_round(torch.randn(n, da, generator=g_)).requires_grad_(True)
    b = bf16_round(torch.randn(n, da, generator=g_))
    loss = R.mse_loss(a, b)            # EV:160-165 casts to fp32 internally
    (g,) = torch.autograd.grad(loss, a)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="mse", input=a.detach().numpy().astype(np.float32),
                        target=b.numpy().astype(np.float32), loss=loss.detach().numpy(), grad=g.numpy())
    print(f"{name}: loss={loss.item():.9e}")


def full_vocab_case(name, N, T, alpha, seed, peaked):
    V = synthetic.VOCAB
    t, s, lab = synthetic.logit_batch(N, V, seed=seed, peaked=peaked)
    digest = sha(t, s, lab)
    loss, grad = run_loca(t.to(torch.float64)[None], s.to(torch.float64)[None], lab[None], T, alpha)
    grad = grad[0]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), kind="loca_reference_fullvocab", n_rows=N, seed=seed,
                        peaked=peaked, T=T, alpha=alpha, input_sha256=digest, labels=lab.numpy(),
                        loss=loss.numpy(), grad_norm=grad.norm().numpy(), grad_stride=37,
                        grad_sample=grad[:, ::37].numpy(), grad_at_labels=grad[:, lab].numpy())
    print(f"{name}: loss={loss.item():.12e} |grad|={grad.norm().item():.6e} sha={digest[:16]}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    logit_case("loca_ref_T2_small", 2, 9, 257, 2.0, 0.8, 11)
    logit_case("loca_ref_T08_dups", 3, 8, 1000, 0.8, 0.8, 12, dup_labels=True, label_is_runner_up=True)
    logit_case("loca_ref_T1_lg_teacher_wider", 2, 6, 640, 1.0, 0.8, 13, Vt=700, which="LG")
    logit_case("loca_ref_clamp_active", 2, 8, 2048, 0.8, 0.8, 14, peaked=True, student_boost=6.0)
    logit_case("loca_ref_alpha05_single_token", 1, 1, 333, 2.0, 0.5, 15)
    logit_case("loca_row_T2", 2, 5, 515, 2.0, 0.8, 16, per_row=True, label_is_runner_up=True)
    logit_case("loca_row_clamp_active", 1, 6, 1536, 0.8, 0.8, 17, per_row=True, peaked=True, student_boost=6.0)
    kl_case("kl_plain_T08", 2, 7, 900, 0.8, 21, quirk=False)
    kl_case("kl_plain_T2_teacher_wider", 1, 5, 512, 2.0, 22, quirk=False, Vt=520)
    kl_case("kl_logtarget_T08", 2, 7, 900, 0.8, 23, quirk=True)
    contrastive_case("info_nce_pooled", 6, 7, 48, 31)
    contrastive_case("info_nce_pooled_siglip_dim", 5, 3, 1152, 32)
    contrastive_case("info_nce_pooled_independent", 9, 4, 96, 33, independent=True)
    mse_case("mse_small", 37, 53, 41)
    full_vocab_case("loca_ref_fullvocab_T2", 4, 2.0, 0.8, 1234, peaked=False)
    full_vocab_case("loca_ref_fullvocab_T08_peaked", 3, 0.8, 0.8, 77, peaked=True)


if __name__ == "__main__":
    main()
