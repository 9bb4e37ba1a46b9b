"""TEST INFRASTRUCTURE ONLY — closed-form CPU restatement of the reference KD losses.

Never imported by the product package.  Every function cites the reference
lines it restates (paths relative to ``/root/reference``; abbreviations as in
SURVEY.md: P1 = distillation/knowledge_distillation7b_double_trouble/phase1/
OnlineKnowledgeDistillationLLavaOneVision.py, LG = ..._logit_based/..., FB =
..._feature_based/..., EV = evaluation/onevisionv3/...).

The restatement works on flattened rows ``[N, V]`` (N = B*S tokens), returns
the loss AND the analytic gradient w.r.t. the student tensor, and is chunked
over rows so that shapes the eager reference cannot hold in memory still have
an oracle.  It is pinned against the imported reference functions
(``oracle/ref_import.py``) by ``tests/test_oracle_vs_reference.py`` and
against the frozen goldens in ``tests/golden`` by ``tests/test_oracle_goldens.py``.

Conventions shared with the CUDA kernels:
  * 2nd-best class ``k`` = runner-up under "largest value, lowest index first"
    (torch.topk leaves ties implementation-defined; goldens are tie-free).
  * duplicate override columns: the LAST token in flattened order wins, and the
    ``k`` pass runs after the ``labels`` pass (CPU ``index_put_`` behaviour of
    P1:184-185).
  * ``Σ_v tp[r, v]`` in P1:184 is taken as the computed fp sum (≈ 1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

LOCA_REFERENCE = "reference"   # as written: one global override table shared by every row
LOCA_PER_ROW = "per_row"       # paper-intended: each row overrides only its own two cells


@dataclass
class KDResult:
    loss: torch.Tensor            # 0-d, dtype of the computation
    grad: torch.Tensor | None     # same shape as the student tensor
    aux: dict


def _as_rows(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-1])


# --------------------------------------------------------------------------------------
# A1 — LoCa-calibrated logit KL  (P1:141-194, LG:208-261)
# --------------------------------------------------------------------------------------

def top2_lowest_index(p: torch.Tensor):
    """(best idx, runner-up idx) per row with ties resolved towards the lowest index."""
    V = p.shape[-1]
    ar = torch.arange(V, device=p.device)
    m1 = p.max(dim=-1, keepdim=True).values
    i1 = torch.where(p == m1, ar, V).min(dim=-1).values
    p2 = p.clone()
    p2.scatter_(1, i1[:, None], -math.inf)
    m2 = p2.max(dim=-1, keepdim=True).values
    i2 = torch.where(p2 == m2, ar, V).min(dim=-1).values
    return i1, i2


def loca_row_stats(teacher: torch.Tensor, labels: torch.Tensor, T: float, alpha: float, V: int,
                   dtype=torch.float64, chunk_rows: int | None = None):
    """Per-token calibration scalars of P1:155-180.

    Returns dict of [N] tensors: p_gt (P1:166), k (P1:170-171), p_k (P1:174),
    v1 = 1 - s*(sum tp - p_gt) (P1:184 rhs), v2 = s*p_k (P1:185 rhs), tsum.
    Runs on the device of ``teacher`` (the tests run it in fp64 on the GPU for the full benchmark shape), row chunk by
    row chunk when ``chunk_rows`` is given.
    """
    rows_all = _as_rows(teacher)
    if chunk_rows is not None and rows_all.shape[0] > chunk_rows:
        lab_all = labels.reshape(-1)
        parts = [loca_row_stats(rows_all[r0:r0 + chunk_rows], lab_all[r0:r0 + chunk_rows], T, alpha, V, dtype)
                 for r0 in range(0, rows_all.shape[0], chunk_rows)]
        return {k: torch.cat([p_[k] for p_ in parts]) for k in parts[0]}
    t = rows_all[:, :V].to(dtype)                               # P1:155 (vocab slice)
    lab = labels.reshape(-1).long().to(t.device)
    if lab.numel() and (lab.min() < 0 or lab.max() >= V):
        raise IndexError("labels must lie in [0, V) — the reference raises on -100 (SURVEY §8a)")
    tp = torch.softmax(t / T, dim=-1)                             # P1:158
    p_gt = tp.gather(1, lab[:, None])[:, 0]                       # P1:166
    _, k = top2_lowest_index(t)                                   # P1:170-171 (order of logits == order of probs)
    p_k = tp.gather(1, k[:, None])[:, 0]                          # P1:174
    scale = alpha / (1.0 - p_gt + p_k)                            # P1:177-180
    tsum = tp.sum(dim=-1)
    v1 = 1.0 - scale * (tsum - p_gt)                              # P1:184
    v2 = scale * p_k                                              # P1:185
    return {"p_gt": p_gt, "k": k, "p_k": p_k, "v1": v1, "v2": v2, "tsum": tsum, "labels": lab}


def build_override_table(labels: torch.Tensor, k: torch.Tensor, v1: torch.Tensor, v2: torch.Tensor, V: int):
    """Global column override of P1:184-185: returns ``ov`` [V], NaN where no token wrote.

    For every row the SAME columns are overwritten: column labels[r] <- v1[r]
    (row order, last wins), then column k[r] <- v2[r] (row order, last wins).
    """
    dev = v1.device
    ov = torch.full((V,), math.nan, dtype=v1.dtype, device=dev)
    n = labels.numel()
    order = torch.arange(n, device=dev)
    own_l = torch.full((V,), -1, dtype=torch.long, device=dev).scatter_reduce(0, labels, order, "amax", include_self=True)
    own_k = torch.full((V,), -1, dtype=torch.long, device=dev).scatter_reduce(0, k, order, "amax", include_self=True)
    sel = own_l >= 0
    ov[sel] = v1[own_l[sel]]
    sel = own_k >= 0
    ov[sel] = v2[own_k[sel]]
    return ov


def loca_kl(teacher, student, labels, T: float, alpha: float = 0.8, eps: float = 1e-8,
            mode: str = LOCA_REFERENCE, inv_norm: float | None = None, dtype=torch.float64,
            want_grad: bool = True, chunk_rows: int = 64, stats_override=None, grad_stride: int | None = None) -> KDResult:
    """LoCa KL (without the ``+ student_loss`` term of P1:194) and d loss / d student.

    loss = T^2 * inv_norm * sum_{r,v} [ xlogy(q,q) - q*log(max(sp,eps)) ]      (P1:188-192)
    grad = T   * inv_norm * ( sp * Q_r - q * [sp >= eps] ),  Q_r = sum_v q*[sp >= eps]
    with inv_norm = 1/(N*V) for the reference's reduction='mean'.

    ``stats_override`` (dict with labels,k,v1,v2 of ALL tokens in global order) lets a caller
    emulate token-sharded execution: the table is built from the global list while
    ``teacher``/``student`` hold only the local rows.
    ``grad_stride``: keep only every ``grad_stride``-th column of the gradient (``grad`` is then [N, ceil(V/stride)]) plus,
    in ``aux``, the gradient at each row's label / runner-up cell and the per-row gradient norm — the full fp64 gradient
    of the benchmark shape (16 384 x 151 936) would be 20 GB.  Device-agnostic: runs where ``student`` lives.
    """
    s_rows = _as_rows(student)
    t_rows = _as_rows(teacher)
    N, V = s_rows.shape
    dev = s_rows.device
    if inv_norm is None:
        inv_norm = 1.0 / (N * V)
    st = loca_row_stats(t_rows, labels, T, alpha, V, dtype=dtype, chunk_rows=chunk_rows)
    ov = None
    if mode == LOCA_REFERENCE:
        src = st if stats_override is None else stats_override
        ov = build_override_table(src["labels"].reshape(-1).long().to(dev), src["k"].reshape(-1).long().to(dev),
                                  src["v1"].to(dev, dtype), src["v2"].to(dev, dtype), V)
    elif mode != LOCA_PER_ROW:
        raise ValueError(mode)

    total = torch.zeros((), dtype=dtype, device=dev)
    n_keep = V if grad_stride is None else (V + grad_stride - 1) // grad_stride
    grad = torch.empty((N, n_keep), dtype=dtype, device=dev) if want_grad else None
    g_lab = torch.empty(N, dtype=dtype, device=dev)
    g_k = torch.empty(N, dtype=dtype, device=dev)
    g_norm = torch.empty(N, dtype=dtype, device=dev)
    Q_all = torch.empty(N, dtype=dtype, device=dev)
    for r0 in range(0, N, chunk_rows):
        r1 = min(N, r0 + chunk_rows)
        t = t_rows[r0:r1, :V].to(dtype)
        s = s_rows[r0:r1].to(dtype)
        tp = torch.softmax(t / T, dim=-1)
        sp = torch.softmax(s / T, dim=-1)                         # P1:159
        q = tp.clone()                                            # P1:183
        if mode == LOCA_REFERENCE:
            q = torch.where(torch.isnan(ov)[None, :], q, ov[None, :].expand_as(q))
        else:
            rows = torch.arange(r1 - r0, device=dev)
            q[rows, st["labels"][r0:r1]] = st["v1"][r0:r1]
            q[rows, st["k"][r0:r1]] = st["v2"][r0:r1]
        keep = sp >= eps                                          # clamp(min=eps) passes grad where sp >= eps (P1:161-162)
        logsp = torch.log(torch.clamp(sp, min=eps))
        total = total + (torch.xlogy(q, q) - q * logsp).sum()     # F.kl_div pointwise (P1:188-192)
        Q = (q * keep).sum(dim=-1)
        Q_all[r0:r1] = Q
        if want_grad:
            gchunk = (T * inv_norm) * (sp * Q[:, None] - q * keep)
            rows = torch.arange(r1 - r0, device=dev)
            g_lab[r0:r1] = gchunk[rows, st["labels"][r0:r1]]
            g_k[r0:r1] = gchunk[rows, st["k"][r0:r1]]
            g_norm[r0:r1] = gchunk.norm(dim=1)
            grad[r0:r1] = gchunk if grad_stride is None else gchunk[:, ::grad_stride]
    loss = total * (T * T * inv_norm)
    aux = {"stats": st, "override": ov, "Q": Q_all, "grad_at_label": g_lab, "grad_at_k": g_k, "grad_row_norm": g_norm}
    if want_grad and grad_stride is None:
        grad = grad.reshape(student.shape)
    return KDResult(loss, grad if want_grad else None, aux)


# --------------------------------------------------------------------------------------
# A2 / A3 — plain temperature KL and the log_target quirk
# --------------------------------------------------------------------------------------

def temperature_kl(teacher, student, T: float, inv_norm: float | None = None, log_target_quirk: bool = False,
                   dtype=torch.float64, want_grad: bool = True, chunk_rows: int = 64) -> KDResult:
    """A2: F.kl_div(log_softmax(s/T), softmax(t/T))*T^2 (P1:331-343, EV:261-273);
    batchmean / sum-over-size(0) variants (EV:130-136, P1:378-381) differ only in ``inv_norm``.
    A3 (``log_target_quirk``): same call with log_target=True although the target is a
    probability (FB:206-219, LG:115-122): pointwise exp(tp)*(tp - log sp).
    """
    s_rows = _as_rows(student)
    t_rows = _as_rows(teacher)
    N, V = s_rows.shape
    dev = s_rows.device
    if inv_norm is None:
        inv_norm = 1.0 / (N * V)
    total = torch.zeros((), dtype=dtype, device=dev)
    grad = torch.empty((N, V), dtype=dtype, device=dev) if want_grad else None
    for r0 in range(0, N, chunk_rows):
        r1 = min(N, r0 + chunk_rows)
        t = t_rows[r0:r1, :V].to(dtype)
        s = s_rows[r0:r1].to(dtype)
        tp = torch.softmax(t / T, dim=-1)
        lsp = torch.log_softmax(s / T, dim=-1)
        if log_target_quirk:
            w = torch.exp(tp)
            total = total + (w * (tp - lsp)).sum()
        else:
            w = tp
            total = total + (torch.xlogy(tp, tp) - tp * lsp).sum()
        if want_grad:
            grad[r0:r1] = (T * inv_norm) * (torch.exp(lsp) * w.sum(dim=-1, keepdim=True) - w)
    return KDResult(total * (T * T * inv_norm), grad.reshape(student.shape) if want_grad else None, {})


# --------------------------------------------------------------------------------------
# (f)1 — the student's cross-entropy the reference adds at P1:194 / P1:260
# --------------------------------------------------------------------------------------

def student_cross_entropy(student, ce_labels, ignore_index: int = -100, dtype=torch.float64, want_grad: bool = True) -> KDResult:
    """``student_outputs.loss`` of the HF causal LM called at P1:238-240 with ``labels=labels``: ``F.cross_entropy`` of the
    logits against the NEXT token (the model shifts: logits[..., :-1, :] vs labels[..., 1:]), ``ignore_index=-100``
    (the data module's padding value, DM:145-146), mean over the valid targets.  ``ce_labels`` are already aligned with the
    logit rows (row (b, s) carries labels[b, s + 1]; the last position of a sequence carries ``ignore_index``), which is the
    same sum over the same (row, target) pairs."""
    s = student.detach().to(dtype).reshape(-1, student.shape[-1]).clone().requires_grad_(want_grad)
    loss = torch.nn.functional.cross_entropy(s, ce_labels.reshape(-1).to(torch.int64), ignore_index=ignore_index)
    grad = None
    if want_grad:
        (grad,) = torch.autograd.grad(loss, s)
        grad = grad.reshape(student.shape)
    return KDResult(loss.detach(), grad, {})


# --------------------------------------------------------------------------------------
# A4 — InfoNCE on (optionally mean-pooled) features (P1:243-248 + P1:393-416)
# --------------------------------------------------------------------------------------

def info_nce(student_feats, teacher_feats, temperature: float = 0.07, pool: bool | None = None,
             dtype=torch.float64, want_grad: bool = True) -> KDResult:
    """``pool`` (default: inputs are 3-D) applies ``.mean(dim=1)`` + ``F.normalize`` as the
    caller does at P1:243-248; the loss itself normalises again (idempotent, P1:406-407),
    forms student x teacher^T / temperature (P1:410) and takes cross-entropy against the
    diagonal (P1:413-414).  Gradient w.r.t. the *un-pooled* student features via autograd.
    """
    xs = student_feats.detach().to(dtype).clone().requires_grad_(want_grad)
    xt = teacher_feats.detach().to(dtype)
    if pool is None:
        pool = xs.dim() == 3
    fs, ft = (xs.mean(dim=1), xt.mean(dim=1)) if pool else (xs, xt)
    us = torch.nn.functional.normalize(fs, p=2, dim=-1)
    ut = torch.nn.functional.normalize(ft, p=2, dim=-1)
    logits = us @ ut.T / temperature
    M = logits.shape[0]
    loss = (torch.logsumexp(logits, dim=-1) - logits.diagonal()).sum() / M
    grad = None
    if want_grad:
        (grad,) = torch.autograd.grad(loss, xs)
    return KDResult(loss.detach(), grad, {"logits": logits.detach()})


# --------------------------------------------------------------------------------------
# A5 — feature MSE (EV:160-165)
# --------------------------------------------------------------------------------------

def feature_mse(inp, target, dtype=torch.float64, want_grad: bool = True) -> KDResult:
    """mean((input.float() - target.float())**2) over all elements; grad = 2*(a-b)/numel."""
    a = inp.to(dtype)
    b = target.to(dtype)
    d = a - b
    loss = (d * d).mean()
    grad = (2.0 / d.numel()) * d if want_grad else None
    return KDResult(loss, grad, {})
