"""TEST INFRASTRUCTURE ONLY — eager-PyTorch port of the reference loss calls.

This is the CPU *baseline* (``bench.py`` ``cpu_baseline`` / ``--impl reference``,
kind "port"): the same sequence of ATen ops the reference executes, so that its
timing on the GPU box's host cores stands in for the reference, which is a
Python tree that cannot travel to the box.  It is pinned bit-for-bit against
the imported reference in ``tests/test_oracle_vs_reference.py`` (same ops on the
same backend give identical bits).  The product never imports it.

Citations: P1 = distillation/knowledge_distillation7b_double_trouble/phase1/
OnlineKnowledgeDistillationLLavaOneVision.py; FB = ..._feature_based/...;
EV = evaluation/onevisionv3/... (all under /root/reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def loca_loss_eager(teacher_logits, student_logits, labels, T, alpha=0.8, student_loss=None):
    """P1:141-194 — inputs [B,S,Vt], [B,S,V], [B,S]; returns loca + student_loss."""
    V = student_logits.size(-1)
    tp = F.softmax(teacher_logits[..., :V] / T, dim=-1)                       # P1:155,158
    sp = F.softmax(student_logits / T, dim=-1)                                # P1:159
    log_sp = sp.clamp(min=1e-8).log()                                         # P1:161-162,189
    idx = labels.unsqueeze(-1)
    p_gt = tp.gather(2, idx).squeeze(-1)                                      # P1:166
    _ = sp.gather(2, idx).squeeze(-1)                                         # P1:167 (dead value, kept for timing parity)
    runner_up = tp.topk(2, dim=-1).indices[..., 1]                            # P1:170-171
    p_k = tp.gather(2, runner_up.unsqueeze(-1)).squeeze(-1)                   # P1:174
    scale = alpha * (1 / (1 - p_gt + p_k))                                    # P1:177-180
    q = tp.clone()                                                            # P1:183
    q[:, :, labels] = 1 - scale * (tp.sum(dim=-1) - p_gt)                     # P1:184 (global column write)
    q[:, :, runner_up] = scale * p_k                                          # P1:185
    kd = F.kl_div(log_sp, q, reduction="mean") * (T ** 2)                     # P1:188-192
    return kd if student_loss is None else kd + student_loss                  # P1:194


def temperature_kl_eager(teacher_logits, student_logits, T, log_target_quirk=False):
    """P1:331-343 (plain) or FB:206-219 (``log_target=True`` passed with a probability target)."""
    V = student_logits.size(-1)
    target = F.softmax(teacher_logits[..., :V] / T, dim=-1)
    log_sp = F.log_softmax(student_logits / T, dim=-1)
    if log_target_quirk:
        return F.kl_div(log_sp, target, reduction="mean", log_target=True) * (T ** 2)
    return F.kl_div(log_sp, target, reduction="mean") * (T ** 2)


def pooled_features_eager(feats):
    """P1:243-248 — token mean then L2 normalise."""
    return F.normalize(feats.mean(dim=1), p=2, dim=-1)


def contrastive_eager(student_feats, teacher_feats, temperature=0.07):
    """P1:393-416 — one-directional NT-Xent against the diagonal."""
    a = F.normalize(student_feats, p=2, dim=-1)
    b = F.normalize(teacher_feats, p=2, dim=-1)
    sim = torch.mm(a, b.T) / temperature
    return F.cross_entropy(sim, torch.arange(sim.size(0), device=sim.device))


def mse_eager(inp, target):
    """EV:160-165."""
    return torch.mean((inp.to(torch.float32) - target.to(torch.float32)) ** 2)
