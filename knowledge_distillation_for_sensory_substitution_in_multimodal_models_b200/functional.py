"""torch.autograd.Function front-ends of the B200 KD-loss kernels.

These are the drop-in replacements for the reference's loss *calls* (SURVEY.md §8(b)):

* :func:`loca_kl`        — body of ``compute_loca_loss`` (P1:141-194, LG:208-261) minus ``+ student_loss``
* :func:`temperature_kl` — KL term of ``compute_vision_loss`` (P1:331-343) / of ``compute_loss`` (FB:206-219)
* :func:`info_nce`       — pooling (P1:243-248) + ``contrastive_loss`` (P1:393-416)
* :func:`feature_mse`    — ``mse_loss`` (EV:160-165)

Forward and backward are ONE fused device pass: the gradient w.r.t. the student tensor is produced in
the forward launch (scaled by ``upstream`` if given).  ``backward`` launches a re-emission kernel that
exits immediately when the incoming ``grad_output`` equals that scale (checked on the device, no host
sync) and otherwise recomputes the gradient from the saved row statistics in one streaming sweep, so the
result is rounded to the output dtype exactly once either way.  PyTorch is used for
memory, streams and autograd plumbing only; all arithmetic runs in ``libkd_b200.so``.
"""
from __future__ import annotations

import torch

from . import _lib

_DTYPES = {torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16, torch.float32: _lib.DTYPE_F32}
_VARIANTS = {"reference": _lib.LOCA_REFERENCE, "per_row": _lib.LOCA_PER_ROW}


def _require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = tensors[0].device
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("KD-loss kernels are CUDA-only (sm_100a); got a CPU tensor — there is no CPU fallback")
        if t.device != dev:
            raise RuntimeError(f"tensors on different devices: {t.device} vs {dev}")
    return dev


class _device_guard:
    """``torch.cuda.device(dev)`` that costs nothing when ``dev`` already is the current device (the usual case: one process
    per GPU) — the stock context manager is ~10 us of the ~140 us a small op spends on the host."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev: torch.device):
        self.idx = dev.index

    def __enter__(self):
        self.prev = -1
        if self.idx is not None:
            cur = torch.cuda.current_device()
            if cur != self.idx:
                torch.cuda.set_device(self.idx)
                self.prev = cur
        return self

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; use bfloat16, float16 or float32") from None


def _rows(x: torch.Tensor) -> torch.Tensor:
    """View as [N, V] without copying when the last dim is contiguous and rows are evenly strided."""
    v = x.shape[-1]
    if x.dim() == 2 and x.stride(1) == 1:
        return x
    if x.stride(-1) == 1:
        try:
            return x.view(-1, v) if x.is_contiguous() else x.reshape(-1, v)
        except RuntimeError:
            pass
    return x.reshape(-1, v).contiguous()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)      # what torch's own generated code uses: ~0.3 us


def _stream_ptr(dev: torch.device) -> int:
    """The current CUDA stream of ``dev`` as an integer handle (``torch.cuda.current_stream(dev).cuda_stream`` builds a Stream
    object per call: ~4 us, twice per step of a small op)."""
    if _raw_stream is not None and dev.index is not None:
        return _raw_stream(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def _backward_scalars(grad_output, upstream, dev):
    """(grad_output, grad_output / scale folded in at forward time) as device scalars — no host sync."""
    go = grad_output
    if go.dtype != torch.float32 or go.device != dev or go.dim() != 0 or go.requires_grad:      # (the usual case needs none of it)
        go = go.detach().to(device=dev, dtype=torch.float32).reshape(()).contiguous()
    ratio = go if upstream is None else (go / upstream).contiguous()
    return go, ratio


def _upstream_ptr(upstream, dev):
    if upstream is None:
        return None, None
    t = torch.as_tensor(upstream, dtype=torch.float32, device=dev).reshape(())
    return t, t.data_ptr()


def _inv_norm(reduction: str, n_rows: int, vocab: int, batch: int, n_total) -> float:
    n = int(n_total) if n_total is not None else n_rows
    if reduction == "mean":          # F.kl_div(reduction='mean'): divide by every element (P1:188-192)
        return 1.0 / (float(n) * float(vocab))
    if reduction == "batchmean":     # EV:130-136 / P1:378-381: divide by size(0)
        return 1.0 / float(batch)
    if reduction == "sum":
        return 1.0
    raise ValueError(f"unknown reduction {reduction!r}")


class _LogitKD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, labels, T, alpha, eps, variant, reduction, upstream, n_total, flags):
        dev = _require_cuda(teacher, student, labels)
        if teacher.dtype != student.dtype:
            raise TypeError(f"teacher ({teacher.dtype}) and student ({student.dtype}) must share a dtype")
        if teacher.shape[:-1] != student.shape[:-1]:
            raise ValueError(f"teacher {tuple(teacher.shape)} and student {tuple(student.shape)} disagree on token dims")
        V = student.shape[-1]
        if teacher.shape[-1] < V:
            raise ValueError("teacher vocabulary is smaller than the student's (the reference slices the teacher, P1:155)")
        t2, s2 = _rows(teacher.detach()), _rows(student.detach())
        n_rows = s2.shape[0]
        batch = student.shape[0] if student.dim() > 1 else 1
        lab = None
        if variant in (_lib.LOCA_REFERENCE, _lib.LOCA_PER_ROW):
            if labels is None:
                raise ValueError("LoCa needs labels")
            if labels.numel() != n_rows:
                raise ValueError(f"labels has {labels.numel()} entries for {n_rows} tokens")
            lab = labels.detach().reshape(-1).to(torch.int64).contiguous()
        need_grad = student.requires_grad and not (flags & _lib.FLAG_NO_GRAD)
        if not need_grad:
            flags |= _lib.FLAG_NO_GRAD
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_logit_kd_workspace_bytes", n_rows, V)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty((n_rows, V), dtype=student.dtype, device=dev) if need_grad else None
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_logit_kd_fwd_bwd(
                t2.data_ptr(), t2.stride(0), s2.data_ptr(), s2.stride(0),
                lab.data_ptr() if lab is not None else None,
                grad.data_ptr() if grad is not None else None, V, loss.data_ptr(),
                ws.data_ptr(), ws_bytes, n_rows, V, _dtype_code(student), float(T), float(alpha), float(eps),
                int(variant), _inv_norm(reduction, n_rows, V, batch, n_total), up_ptr, int(flags), _stream_ptr(dev)))
        ctx.grad = grad
        ctx.need_grad = need_grad
        ctx.student_shape = student.shape
        ctx.dev = dev
        ctx.upstream = up_t
        ctx.workspace = ws      # row statistics (kd_logit_kd_rowaux_offset), token records and the override table
        if need_grad:           # inputs of the backward re-emission
            ctx.rows = (t2, s2)
            ctx.call = (n_rows, V, _dtype_code(student), float(T), float(eps), int(variant),
                        _inv_norm(reduction, n_rows, V, batch, n_total))
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.need_grad:
            return (None,) * 11
        # First backward: the gradient the forward launch produced, re-emitted in place only if grad_output differs from the
        # scale folded in there.  A later backward through the same graph (retain_graph=True, two losses sharing the node)
        # must not touch that buffer — autograd may have handed it out as ``student.grad`` — so the gradient is recomputed
        # from the saved row statistics into a fresh one (skip pointer NULL = never skip).
        first = ctx.grad is not None
        go, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        t2, s2 = ctx.rows
        n_rows, V, dt, T, eps, variant, inv_norm = ctx.call
        ws = ctx.workspace
        with _device_guard(ctx.dev):
            grad = ctx.grad if first else torch.empty((n_rows, V), dtype=s2.dtype, device=ctx.dev)
            ctx.grad = None
            _lib.check(_lib.load().kd_logit_kd_regrad(t2.data_ptr(), t2.stride(0), s2.data_ptr(), s2.stride(0),
                                                      grad.data_ptr(), grad.stride(0), ws.data_ptr(), ws.numel(),
                                                      n_rows, V, dt, T, eps, variant, inv_norm, go.data_ptr(),
                                                      ratio.data_ptr() if first else None, _stream_ptr(ctx.dev)))
        return (None, grad.view(ctx.student_shape), None, None, None, None, None, None, None, None, None)


CE_IGNORE = -100


def shift_labels(labels: torch.Tensor, ignore_index: int = CE_IGNORE) -> torch.Tensor:
    """Next-token targets aligned with the logit rows: ``out[b, s] = labels[b, s + 1]`` and ``ignore_index`` at the last
    position — the shift HF's causal-LM loss applies to the ``labels`` the reference passes to the student (P1:238-240)."""
    pad = torch.full_like(labels[..., :1], ignore_index)
    return torch.cat([labels[..., 1:], pad], dim=-1)


class _LocaKLWithCE(torch.autograd.Function):
    """As-written LoCa KL + the student's shifted cross-entropy from one pass over the logits (``kd_logit_kd_ce_fwd_bwd``)."""

    @staticmethod
    def forward(ctx, teacher, student, labels, ce_labels, T, alpha, eps, ce_weight, upstream, n_total, ce_valid_count):
        dev = _require_cuda(teacher, student, labels, ce_labels)
        if teacher.dtype != student.dtype:
            raise TypeError(f"teacher ({teacher.dtype}) and student ({student.dtype}) must share a dtype")
        if teacher.shape[:-1] != student.shape[:-1]:
            raise ValueError(f"teacher {tuple(teacher.shape)} and student {tuple(student.shape)} disagree on token dims")
        V = student.shape[-1]
        if teacher.shape[-1] < V:
            raise ValueError("teacher vocabulary is smaller than the student's (the reference slices the teacher, P1:155)")
        t2, s2 = _rows(teacher.detach()), _rows(student.detach())
        n_rows = s2.shape[0]
        if labels.numel() != n_rows or ce_labels.numel() != n_rows:
            raise ValueError(f"labels / ce_labels have {labels.numel()} / {ce_labels.numel()} entries for {n_rows} tokens")
        lab = labels.detach().reshape(-1).to(torch.int64).contiguous()
        cel = ce_labels.detach().reshape(-1).to(torch.int64).contiguous()
        # 1 / (number of valid next-token targets) as a device scalar: nothing is read back (0 valid -> inf -> NaN like torch)
        count = (cel != CE_IGNORE).sum() if ce_valid_count is None else torch.as_tensor(ce_valid_count, device=dev)
        inv_count = (1.0 / count.to(torch.float32)).reshape(()).contiguous()
        need_grad = student.requires_grad
        flags = 0 if need_grad else _lib.FLAG_NO_GRAD
        inv_norm = _inv_norm("mean", n_rows, V, 1, n_total)
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_logit_kd_workspace_bytes", n_rows, V)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            ce = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty((n_rows, V), dtype=student.dtype, device=dev) if need_grad else None
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_logit_kd_ce_fwd_bwd(
                t2.data_ptr(), t2.stride(0), s2.data_ptr(), s2.stride(0), lab.data_ptr(), cel.data_ptr(), inv_count.data_ptr(),
                float(ce_weight), grad.data_ptr() if grad is not None else None, V, loss.data_ptr(), ce.data_ptr(),
                ws.data_ptr(), ws_bytes, n_rows, V, _dtype_code(student), float(T), float(alpha), float(eps), inv_norm, up_ptr,
                int(flags), _stream_ptr(dev)))
        ctx.grad, ctx.need_grad, ctx.student_shape, ctx.dev, ctx.upstream, ctx.workspace = grad, need_grad, student.shape, dev, up_t, ws
        if need_grad:
            ctx.rows = (t2, s2, cel, inv_count)
            ctx.call = (n_rows, V, _dtype_code(student), float(T), float(eps), float(ce_weight), inv_norm)
        ctx.mark_non_differentiable(ce)
        return loss, ce

    @staticmethod
    def backward(ctx, grad_output, _grad_ce):
        if not ctx.need_grad:
            return (None,) * 11
        first = ctx.grad is not None           # see _LogitKD.backward: a repeated backward re-emits into a fresh buffer
        go, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        t2, s2, cel, inv_count = ctx.rows
        n_rows, V, dt, T, eps, ce_weight, inv_norm = ctx.call
        ws = ctx.workspace
        with _device_guard(ctx.dev):
            grad = ctx.grad if first else torch.empty((n_rows, V), dtype=s2.dtype, device=ctx.dev)
            ctx.grad = None
            _lib.check(_lib.load().kd_logit_kd_ce_regrad(t2.data_ptr(), t2.stride(0), s2.data_ptr(), s2.stride(0), cel.data_ptr(),
                                                         inv_count.data_ptr(), ce_weight, grad.data_ptr(), grad.stride(0),
                                                         ws.data_ptr(), ws.numel(), n_rows, V, dt, T, eps, inv_norm, go.data_ptr(),
                                                         ratio.data_ptr() if first else None, _stream_ptr(ctx.dev)))
        return (None, grad.view(ctx.student_shape)) + (None,) * 9


def loca_kl_ce(teacher_logits, student_logits, labels, ce_labels, T: float, alpha: float = 0.8, eps: float = 1e-8,
               ce_weight: float = 1.0, upstream=None, n_total=None, ce_valid_count=None):
    """``(loca_kl + ce_weight * CE, CE)`` from ONE pass over the logits — SURVEY §8(f) rank 1.

    ``CE`` is the loss the HF student returns at P1:238-240 and the reference adds at P1:194 / P1:260: mean over the valid
    next-token targets of ``logsumexp(s) - s[target]`` (temperature 1).  ``ce_labels`` are those targets aligned with the
    logit rows (:func:`shift_labels` of the ``labels`` given to the model; ``-100`` = ignored); ``labels`` are the UNSHIFTED
    ids LoCa calibrates against (P1:166), all of which must lie in ``[0, V)``.  Call the student with ``labels=None`` so that
    it does not compute its own loss.  The second result is the CE alone (detached, for logging).  As-written LoCa mode,
    16-bit logits, ``V % 8 == 0``; ``ce_valid_count`` overrides the divisor when the rows are a shard of a larger batch."""
    return _LocaKLWithCE.apply(teacher_logits, student_logits, labels, ce_labels, T, alpha, eps, ce_weight, upstream, n_total,
                               ce_valid_count)


def loca_kl(teacher_logits, student_logits, labels, T: float, alpha: float = 0.8, eps: float = 1e-8,
            mode: str = "reference", reduction: str = "mean", upstream=None, n_total=None, force_streaming=False):
    """LoCa-calibrated logit KL (P1:141-194) without the trailing ``+ student_loss``.

    ``mode="reference"`` reproduces the as-written global column override (P1:184-185);
    ``mode="per_row"`` calibrates each token's own two cells.  ``upstream`` is an optional scalar
    (python float or 0-d tensor) folded into the gradient in the forward launch, e.g. ``gamma`` of
    phase 3 (P1:260) or an AMP loss scale: ``backward`` divides the incoming ``grad_output`` by it, so when
    the caller really multiplies the loss by that scalar the rescale pass is skipped on the device.
    """
    flags = _lib.FLAG_FORCE_STREAMING if force_streaming else 0
    return _LogitKD.apply(teacher_logits, student_logits, labels, T, alpha, eps, _VARIANTS[mode], reduction,
                          upstream, n_total, flags)


def temperature_kl(teacher_logits, student_logits, T: float, reduction: str = "mean", log_target_quirk: bool = False,
                   upstream=None, n_total=None, force_streaming=False):
    """``F.kl_div(log_softmax(s/T), softmax(t/T), reduction) * T**2`` (P1:331-343); ``log_target_quirk``
    reproduces FB:206-219 where the probability target is flagged ``log_target=True``."""
    variant = _lib.KL_LOG_TARGET if log_target_quirk else _lib.KL_PLAIN
    flags = _lib.FLAG_FORCE_STREAMING if force_streaming else 0
    return _LogitKD.apply(teacher_logits, student_logits, None, T, 0.0, 1e-8, variant, reduction, upstream, n_total, flags)


def _parts16(x: torch.Tensor, dtype):
    """A 16-bit tensor as itself; an fp32 one as (hi, lo) 16-bit parts with hi + lo = x to ~2^-17, so that tensor-core
    products carry an fp32 operand (the lo x lo term is dropped)."""
    if x.dtype != torch.float32:
        return [x.to(dtype)]
    hi = x.to(dtype)
    return [hi, (x - hi.float()).to(dtype)]


def _gemm_parts(a_parts, b_parts, **kw):
    out = None
    for i, a in enumerate(a_parts):
        for j, b in enumerate(b_parts):
            if i and j:
                continue
            y = gemm(a, b, out_dtype=torch.float32, **kw)
            out = y if out is None else out + y
    return out


class _Project(torch.autograd.Function):
    """``y = x @ W`` (x: [M, K], W: [K, N]) on ``kd_gemm`` — the learned projection of BASELINE configs[2] (1152 -> 896 / 3584).
    fp32 operands (pooled features, an fp32 master weight) are split into 16-bit hi + lo parts so that the tensor-core product
    keeps its digits; the result is fp32 unless both operands are 16-bit (then it has their type, like the library GEMM of an
    autocast forward).  No operand is transposed in HBM: W and the gradients are read through MN-major descriptors."""

    @staticmethod
    def forward(ctx, x, w):
        _require_cuda(x, w)
        if x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[0]:
            raise ValueError("project: x [M, K] and W [K, N] expected")
        d16 = torch.float16 if torch.float16 in (x.dtype, w.dtype) and torch.bfloat16 not in (x.dtype, w.dtype) else torch.bfloat16
        xd, wd = x.detach().contiguous(), w.detach().contiguous()
        ctx.save_for_backward(xd, wd)
        ctx.needs, ctx.d16 = (x.requires_grad, w.requires_grad), d16
        y = _gemm_parts(_parts16(xd, d16), _parts16(wd, d16), b_mn_major=True)
        return y if torch.float32 in (x.dtype, w.dtype) else y.to(d16)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        dx = dw = None
        g = _parts16(gy.contiguous(), ctx.d16)
        if ctx.needs[0]:                                   # dx[M, K] = gy[M, N] . W[K, N]^T  (both K-major as stored)
            dx = _gemm_parts(g, _parts16(w, ctx.d16)).to(x.dtype)
        if ctx.needs[1]:                                   # dW[K, N] = x[M, K]^T . gy[M, N]  (both MN-major as stored)
            dw = _gemm_parts(_parts16(x, ctx.d16), g, a_mn_major=True, b_mn_major=True).to(w.dtype)
        return dx, dw


def project(x: torch.Tensor, proj: torch.Tensor) -> torch.Tensor:
    """``x @ proj`` on the tcgen05 GEMM, differentiable w.r.t. both (see :class:`_Project`); leading dims are flattened."""
    lead = x.shape[:-1]
    y = _Project.apply(x.reshape(-1, x.shape[-1]), proj)
    return y.reshape(*lead, proj.shape[1])


class _PoolTokens(torch.autograd.Function):
    """``x.float().mean(dim=1)`` of ``[M, tokens, D]`` features (P1:243-246) on the library's pooling kernels; the backward is the
    store-only broadcast ``dx[m, p, :] = d_pooled[m, :] / tokens`` in ``x``'s dtype."""

    @staticmethod
    def forward(ctx, x):
        dev = _require_cuda(x)
        if x.dim() != 3:
            raise ValueError("pool_tokens: [M, tokens, D] features expected")
        xd = x.detach().contiguous()
        M, P, D = xd.shape
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_feature_pool_workspace_bytes", M, P, D)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            out = torch.empty((M, D), dtype=torch.float32, device=dev)
            _lib.check(_lib.load().kd_feature_pool_fwd(xd.data_ptr(), out.data_ptr(), ws.data_ptr(), ws_bytes, M, P, D, _dtype_code(xd),
                                                       _stream_ptr(dev)))
        ctx.shape3, ctx.dtype, ctx.dev = (M, P, D), xd.dtype, dev
        return out

    @staticmethod
    def backward(ctx, g):
        M, P, D = ctx.shape3
        gm = (g.to(torch.float32) * (1.0 / P)).contiguous()
        with _device_guard(ctx.dev):
            dx = torch.empty((M, P, D), dtype=ctx.dtype, device=ctx.dev)
            _lib.check(_lib.load().kd_feature_pool_bwd(gm.data_ptr(), dx.data_ptr(), M, P, D, _DTYPES[ctx.dtype], None, _stream_ptr(ctx.dev)))
        return dx


def pool_tokens(x: torch.Tensor) -> torch.Tensor:
    """fp32 mean over the token dimension of ``[M, tokens, D]`` features, differentiable (see :class:`_PoolTokens`)."""
    return _PoolTokens.apply(x)


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, temperature, upstream, similarity="auto"):
        dev = _require_cuda(student, teacher)
        if similarity not in ("auto", "cuda_core"):
            raise ValueError("similarity: 'auto' (tensor cores for token-level M) or 'cuda_core' (the exact fp32 row kernel)")
        if student.shape != teacher.shape or student.dtype != teacher.dtype:
            raise ValueError("student and teacher features must have the same shape and dtype")
        if student.dim() == 2:
            M, P, D = student.shape[0], 1, student.shape[1]
        elif student.dim() == 3:
            M, P, D = student.shape
        else:
            raise ValueError("features must be [M, D] or [M, tokens, D]")
        # (inside Function.forward the inputs carry no graph: no detach needed, and .contiguous() only when it does something)
        xs = student if student.is_contiguous() else student.contiguous()
        xt = teacher if teacher.is_contiguous() else teacher.contiguous()
        need_grad = student.requires_grad
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_info_nce_workspace_bytes", M, P, D)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(xs) if need_grad else None
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_info_nce_fwd_bwd(xs.data_ptr(), xt.data_ptr(), grad.data_ptr() if need_grad else None,
                                               loss.data_ptr(), ws.data_ptr(), ws_bytes, M, P, D, _dtype_code(xs),
                                               float(temperature), up_ptr,
                                               (0 if need_grad else _lib.FLAG_NO_GRAD) |
                                               (_lib.FLAG_FORCE_STREAMING if similarity == "cuda_core" else 0),
                                               _stream_ptr(dev)))
        ctx.grad, ctx.dev, ctx.upstream, ctx.need_grad = grad, dev, up_t, need_grad
        ctx.workspace, ctx.shape3, ctx.gdtype, ctx.gshape = ws, (M, P, D), xs.dtype, xs.shape
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.need_grad:
            return None, None, None, None, None
        first = ctx.grad is not None           # see _LogitKD.backward: a repeated backward re-emits into a fresh buffer
        _, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        M, P, D = ctx.shape3
        ws = ctx.workspace
        with _device_guard(ctx.dev):
            grad = ctx.grad if first else torch.empty(ctx.gshape, dtype=ctx.gdtype, device=ctx.dev)
            ctx.grad = None
            _lib.check(_lib.load().kd_info_nce_regrad(ws.data_ptr(), ws.numel(), grad.data_ptr(), M, P, D,
                                                      _dtype_code(grad), ratio.data_ptr(), ratio.data_ptr() if first else None,
                                                      _stream_ptr(ctx.dev)))
        return grad, None, None, None, None


class _InfoNCEProjected(torch.autograd.Function):
    """Pooled contrastive loss through a learned projection in ONE library call (``kd_info_nce_proj_fwd_bwd``): balanced pooling
    of both tensors, ``y = mean(x_s) @ proj`` and the two backward products in fp32 on the CUDA cores (a few dozen rows are no
    tensor-core tile), the row kernel, the store-only broadcast.  Gradients w.r.t. the student features and ``proj`` are written in
    the forward launch; ``backward`` re-emits them only when ``grad_output`` differs from what forward assumed."""

    @staticmethod
    def forward(ctx, student, teacher, proj, temperature, upstream):
        dev = _require_cuda(student, teacher, proj)
        if student.dim() not in (2, 3) or teacher.dim() not in (2, 3) or student.shape[0] != teacher.shape[0]:
            raise ValueError("features must be [M, D] or [M, tokens, D] with the same M")
        if teacher.dtype != student.dtype:
            teacher = teacher.to(student.dtype)
        xs, xt, w = student.detach().contiguous(), teacher.detach().contiguous(), proj.detach().contiguous()
        M, Ds, Dt = xs.shape[0], xs.shape[-1], xt.shape[-1]
        Ps = xs.shape[1] if xs.dim() == 3 else 1
        Pt = xt.shape[1] if xt.dim() == 3 else 1
        if w.dim() != 2 or tuple(w.shape) != (Ds, Dt):
            raise ValueError(f"proj must be [{Ds}, {Dt}] (student width x teacher width), got {tuple(w.shape)}")
        need_x, need_w = student.requires_grad, proj.requires_grad
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_info_nce_proj_workspace_bytes", M, Ds, Dt)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            gx = torch.empty_like(xs) if need_x else None
            gw = torch.empty_like(w) if need_w else None
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_info_nce_proj_fwd_bwd(
                xs.data_ptr(), Ps, xt.data_ptr(), Pt, w.data_ptr(), gx.data_ptr() if need_x else None,
                gw.data_ptr() if need_w else None, loss.data_ptr(), ws.data_ptr(), ws_bytes, M, Ds, Dt, _dtype_code(xs),
                _dtype_code(w), float(temperature), up_ptr, 0 if (need_x or need_w) else _lib.FLAG_NO_GRAD, _stream_ptr(dev)))
        ctx.grads, ctx.dev, ctx.upstream, ctx.needs = (gx, gw), dev, up_t, (need_x, need_w)
        ctx.workspace, ctx.dims = ws, (M, Ps, Ds, Dt)
        ctx.meta = (xs.shape, xs.dtype, w.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        need_x, need_w = ctx.needs
        if not (need_x or need_w):
            return None, None, None, None, None
        first = ctx.grads is not None           # a repeated backward re-emits into fresh buffers
        _, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        M, Ps, Ds, Dt = ctx.dims
        xshape, xdtype, wdtype = ctx.meta
        ws = ctx.workspace
        with _device_guard(ctx.dev):
            if first:
                gx, gw = ctx.grads
            else:
                gx = torch.empty(xshape, dtype=xdtype, device=ctx.dev) if need_x else None
                gw = torch.empty((Ds, Dt), dtype=wdtype, device=ctx.dev) if need_w else None
            ctx.grads = None
            _lib.check(_lib.load().kd_info_nce_proj_regrad(
                ws.data_ptr(), ws.numel(), gx.data_ptr() if need_x else None, gw.data_ptr() if need_w else None,
                M, Ps, Ds, Dt, _DTYPES[xdtype], _DTYPES[wdtype], ratio.data_ptr(), ratio.data_ptr() if first else None,
                _stream_ptr(ctx.dev)))
        return gx, None, gw, None, None


_PROJ_FUSED_MAX_ROWS = 4096        # kd_info_nce_proj_fwd_bwd's limit; more rows = token-level features -> tensor-core similarity


def info_nce(student_feats, teacher_feats, temperature: float = 0.07, upstream=None, proj=None, similarity: str = "auto"):
    """Mean-pool over tokens (3-D input), L2-normalise, student x teacher^T / temperature, cross-entropy
    against the diagonal — P1:243-248 followed by ``contrastive_loss`` (P1:393-416).

    ``proj`` ([D_s, D_t], optional, not in the reference where both towers are SigLIP-1152): a learned projection of the
    student features, ``contrastive_loss(x_s @ proj, x_t)``.  Mean-pooling is linear, so the projection is applied to the pooled
    student vectors.  Up to 4096 rows (the pooled 64-image batch of BASELINE configs[2]) the whole chain is one library call
    (:class:`_InfoNCEProjected`, fp32 products on the CUDA cores, differentiable w.r.t. ``proj``); token-level features are
    projected on the tcgen05 GEMM (``project``) and take the tensor-core similarity path.

    ``similarity``: with ``"auto"`` a token-level call (``M >= 1024`` rows, ``M`` and ``D`` multiples of 8 — e.g. the
    46 656 x 1152 un-pooled SigLIP tokens) computes ``student @ teacher.T`` (P1:410) on the tcgen05 kernel with the row
    log-softmax in its epilogue; the similarity matrix is never stored.  ``"cuda_core"`` forces the exact fp32 row kernel."""
    if proj is not None:
        M = student_feats.shape[0]
        token_level = (similarity == "auto" and student_feats.dim() == 2 and M >= 1024 and M % 8 == 0
                       and proj.shape[1] % 8 == 0)
        if M <= _PROJ_FUSED_MAX_ROWS and not token_level:
            return _InfoNCEProjected.apply(student_feats, teacher_feats, proj, temperature, upstream)
        # token-level features: pool (if 3-D), project the [M, D_s] rows on the tcgen05 kernel, then the 2-D form of the op
        pooled = pool_tokens(student_feats) if student_feats.dim() == 3 else student_feats
        student_feats = project(pooled, proj)
        teacher_feats = pool_tokens(teacher_feats.detach()) if teacher_feats.dim() == 3 else teacher_feats
        if teacher_feats.dtype != student_feats.dtype:
            teacher_feats = teacher_feats.to(student_feats.dtype)
    return _InfoNCE.apply(student_feats, teacher_feats, temperature, upstream, similarity)


class _FeatureMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target, upstream):
        dev = _require_cuda(inp, target)
        if inp.shape != target.shape or inp.dtype != target.dtype:
            raise ValueError("input and target must have the same shape and dtype")
        a, b = inp.detach().contiguous(), target.detach().contiguous()
        need_grad = inp.requires_grad
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_feature_mse_workspace_bytes", a.numel())
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(a) if need_grad else None
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_feature_mse_fwd_bwd(a.data_ptr(), b.data_ptr(), grad.data_ptr() if need_grad else None,
                                                  loss.data_ptr(), ws.data_ptr(), ws_bytes, a.numel(), _dtype_code(a),
                                                  up_ptr, 0 if need_grad else _lib.FLAG_NO_GRAD, _stream_ptr(dev)))
        ctx.grad, ctx.dev, ctx.upstream, ctx.need_grad = grad, dev, up_t, need_grad
        ctx.inputs = (a, b) if need_grad else None
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.need_grad:
            return None, None, None
        first = ctx.grad is not None           # see _LogitKD.backward: a repeated backward re-emits into a fresh buffer
        go, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        a, b = ctx.inputs
        with _device_guard(ctx.dev):
            grad = ctx.grad if first else torch.empty_like(a)
            ctx.grad = None
            _lib.check(_lib.load().kd_feature_mse_regrad(a.data_ptr(), b.data_ptr(), grad.data_ptr(), a.numel(),
                                                         _dtype_code(a), go.data_ptr(), ratio.data_ptr() if first else None,
                                                         _stream_ptr(ctx.dev)))
        return grad, None, None


def gemm(a: torch.Tensor, b: torch.Tensor, a_mn_major: bool = False, b_mn_major: bool = False, out_dtype=None, out=None,
         max_ctas: int = 0):
    """``C[M, N] = A[M, K] . B[N, K]^T`` on the tcgen05 kernel (``kd_gemm``).  ``a`` is ``[M, K]`` (or ``[K, M]`` with
    ``a_mn_major``: the operand is stored transposed and read through an MN-major descriptor, nothing is transposed in HBM),
    ``b`` is ``[N, K]`` (or ``[K, N]`` with ``b_mn_major``).  bf16 / fp16 CUDA operands; ``out_dtype``: theirs or float32.
    ``max_ctas`` > 0 caps the persistent grid (``kd_gemm_ex``) so that a collective running beside the product finds free SMs."""
    dev = _require_cuda(a, b)
    if a.dim() != 2 or b.dim() != 2 or a.dtype != b.dtype:
        raise ValueError("gemm: two 2-D operands of one dtype expected")
    a = a if a.stride(1) == 1 else a.contiguous()
    b = b if b.stride(1) == 1 else b.contiguous()
    M, K = (a.shape[1], a.shape[0]) if a_mn_major else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn_major else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm: contraction lengths differ ({K} vs {Kb})")
    out_dtype = out_dtype or a.dtype
    if out_dtype not in (a.dtype, torch.float32):
        raise TypeError("gemm: the output is the operand type or float32")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=dev)
    elif out.shape != (M, N) or out.dtype != out_dtype or out.stride(1) != 1:
        raise ValueError("gemm: `out` must be a row-major [M, N] tensor of the output dtype")
    with _device_guard(dev):
        _lib.check(_lib.load().kd_gemm_ex(a.data_ptr(), a.stride(0), int(a_mn_major), b.data_ptr(), b.stride(0), int(b_mn_major),
                                          out.data_ptr(), out.stride(0), int(out_dtype == torch.float32), M, N, K, _dtype_code(a),
                                          int(max_ctas), _stream_ptr(dev)))
    return out


class _ProjectedMSE(torch.autograd.Function):
    """``mse_loss(input @ proj, target)`` (EV:160-165 on a projected tensor) with the MSE as the projection GEMM's epilogue."""

    @staticmethod
    def forward(ctx, inp, proj, target, upstream):
        dev = _require_cuda(inp, proj, target)
        x = inp.detach().reshape(-1, inp.shape[-1])
        tgt = target.detach().reshape(-1, target.shape[-1])
        x = x if x.stride(1) == 1 else x.contiguous()
        tgt = tgt if tgt.stride(1) == 1 else tgt.contiguous()
        w = proj.detach() if proj.stride(1) == 1 else proj.detach().contiguous()
        M, K = x.shape
        N = w.shape[1]
        if w.shape[0] != K or tgt.shape != (M, N) or not (x.dtype == w.dtype == tgt.dtype):
            raise ValueError("feature_mse(proj=...): input [.., K], proj [K, N], target [.., N] of one 16-bit dtype expected")
        lib = _lib.load()
        with _device_guard(dev):
            ws_bytes = _lib.size_query("kd_proj_mse_workspace_bytes", M, N)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            gp = torch.empty((M, N), dtype=x.dtype, device=dev)          # d loss / d (input @ proj); the product itself is never stored
            up_t, up_ptr = _upstream_ptr(upstream, dev)
            _lib.check(lib.kd_proj_mse_fwd_bwd(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), tgt.data_ptr(), tgt.stride(0),
                                               gp.data_ptr(), gp.stride(0), loss.data_ptr(), ws.data_ptr(), ws_bytes, M, N, K,
                                               _dtype_code(x), up_ptr, _stream_ptr(dev)))
        ctx.saved = (x, w, gp)
        ctx.dev, ctx.upstream, ctx.in_shape = dev, up_t, inp.shape
        ctx.needs = (inp.requires_grad, proj.requires_grad)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.saved is None:
            raise RuntimeError("feature_mse(proj=...): backward through the same graph twice (the projected gradient is rescaled in place)")
        x, w, g = ctx.saved
        ctx.saved = None
        _, ratio = _backward_scalars(grad_output, ctx.upstream, ctx.dev)
        with _device_guard(ctx.dev):         # d loss / d projected *= grad_output / upstream, skipped on the device when that is 1
            _lib.check(_lib.load().kd_scale_rows(g.data_ptr(), g.stride(0), g.shape[0], g.shape[1], _dtype_code(g), ratio.data_ptr(),
                                                 _stream_ptr(ctx.dev)))
        d_in = gemm(g, w).view(ctx.in_shape) if ctx.needs[0] else None                       # [M, N] . [K, N]^T
        d_proj = gemm(x, g, a_mn_major=True, b_mn_major=True) if ctx.needs[1] else None     # [M, K]^T . [M, N]
        return d_in, d_proj, None, None


def feature_mse(inp, target, upstream=None, proj=None):
    """``mean((input.float() - target.float())**2)`` (EV:160-165); gradient in ``input``'s dtype.

    ``proj`` ([D_in, D_target], optional, not in the reference): ``mse_loss(input @ proj, target)`` for a student whose hidden
    size differs from the teacher's (896 vs 3584; BASELINE configs[3]).  The projection runs on the tcgen05 kernel with the MSE
    as its epilogue — the projected tensor never reaches HBM — and is differentiable w.r.t. ``input`` and ``proj`` (two more
    products on the same kernel).  16-bit operands, dimensions multiples of 8."""
    if proj is not None:
        return _ProjectedMSE.apply(inp, proj, target, upstream)
    return _FeatureMSE.apply(inp, target, upstream)
