"""Deterministic synthetic workloads for the KD-loss path.

Shared by ``bench.py``, ``__graft_entry__.smoke()`` and the tests.  Teacher rows are made
tie-free in their top-3 (the reference's ``topk`` leaves ties implementation-defined,
P1:170-171), labels lie in ``[0, V)`` as the reference requires (P1:166 raises otherwise).
"""
from __future__ import annotations

import torch

VOCAB = 151_936          # Qwen2 vocabulary of the LLaVA-OneVision teacher/student (BASELINE.json)


def _generator(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def make_tie_free_(teacher: torch.Tensor) -> torch.Tensor:
    """Lift each row's best / runner-up so that top1 > top2 > top3 strictly (in-place, any float dtype)."""
    vals, idx = teacher.float().topk(2, dim=-1)
    bump = torch.stack([vals[:, 0] + 1.0, vals[:, 1] + 0.5], dim=-1).to(teacher.dtype)
    teacher.scatter_(1, idx, bump)
    return teacher


def logit_batch(n_rows: int, vocab: int = VOCAB, *, seed: int = 1234, device="cpu", dtype=torch.bfloat16,
                sigma: float = 2.0, peaked: bool = False, teacher_vocab: int | None = None,
                chunk_rows: int = 2048):
    """Teacher / student logit rows and labels: ``sigma * randn`` rounded to ``dtype``.

    ``peaked`` adds +8 at the label column of ~70 % of the teacher rows (top-1 == label, student
    clamp region exercised).  ``teacher_vocab > vocab`` gives the teacher extra (ignored) columns,
    mirroring the slice at P1:155.  Generated in row chunks so a 16k x 151 936 batch never needs
    an fp32 copy of the whole tensor.
    """
    g = _generator(seed, device)
    vt = teacher_vocab or vocab
    teacher = torch.empty((n_rows, vt), dtype=dtype, device=device)
    student = torch.empty((n_rows, vocab), dtype=dtype, device=device)
    labels = torch.randint(0, vocab, (n_rows,), generator=g, device=device, dtype=torch.int64)
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        t = torch.randn((r1 - r0, vt), generator=g, device=device, dtype=torch.float32) * sigma
        s = torch.randn((r1 - r0, vocab), generator=g, device=device, dtype=torch.float32) * sigma
        if peaked:
            hit = torch.rand((r1 - r0,), generator=g, device=device) < 0.7
            t[:, :vocab].scatter_add_(1, labels[r0:r1, None], (8.0 * hit.float())[:, None])
        t = t.to(dtype)
        t[:, :vocab] = make_tie_free_(t[:, :vocab].contiguous())
        teacher[r0:r1] = t
        student[r0:r1] = s.to(dtype)
    return teacher, student, labels


def vision_feature_batch(n_images: int = 64, tokens: int = 729, dim: int = 1152, *, seed: int = 1234,
                         device="cpu", dtype=torch.bfloat16):
    """SigLIP post-layernorm features (P1:110-121): teacher = student + 0.5 * noise so the diagonal is meaningful."""
    g = _generator(seed, device)
    xs = torch.randn((n_images, tokens, dim), generator=g, device=device, dtype=torch.float32)
    xt = xs + 0.5 * torch.randn((n_images, tokens, dim), generator=g, device=device, dtype=torch.float32)
    return xs.to(dtype), xt.to(dtype)


def hidden_state_batch(n_rows: int = 65_536, d_student: int = 896, d_teacher: int = 3584, *, seed: int = 1234,
                       device="cpu", dtype=torch.bfloat16):
    """Student / teacher hidden states and a student->teacher projection (BASELINE.json config 4)."""
    g = _generator(seed, device)
    hs = torch.randn((n_rows, d_student), generator=g, device=device, dtype=torch.float32).to(dtype)
    ht = torch.randn((n_rows, d_teacher), generator=g, device=device, dtype=torch.float32).to(dtype)
    w = (torch.randn((d_student, d_teacher), generator=g, device=device, dtype=torch.float32) * 0.03).to(dtype)
    return hs, ht, w
