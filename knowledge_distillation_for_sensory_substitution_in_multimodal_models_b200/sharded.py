"""Multi-GPU execution of the logit-KD path: one process per GPU, tokens partitioned across ranks.

Per-row variants (``per_row`` LoCa, plain KL, log-target KL) need no data-path collective at all:
every rank runs :func:`functional.loca_kl` / :func:`functional.temperature_kl` on its own rows with
``n_total`` set to the global token count, and the scalar losses are summed.

The as-written LoCa (``mode="reference"``, P1:184-185) couples all tokens through the shared override
table, so it has ONE exchange step: every rank computes 16-byte token records for its rows (phase A),
the records are all-gathered in global token order (rank-major), every rank rebuilds the identical
V-entry table (phase B) and then emits loss partial + gradient for its rows (phase C).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from .functional import _dtype_code, _require_cuda, _rows, _stream_ptr, _upstream_ptr
from .functional import gemm as tc_gemm


class CudaPhases:
    """The three phase-split C-ABI calls (kd_loca_teacher_stats / kd_loca_build_table / kd_loca_emit)."""

    def teacher_stats(self, teacher, labels, T, alpha, V):
        dev = teacher.device
        n = teacher.shape[0]
        rec = torch.empty((n, 4), dtype=torch.int32, device=dev)
        lse2 = torch.empty((n,), dtype=torch.float32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_loca_teacher_stats(teacher.data_ptr(), teacher.stride(0), labels.data_ptr(),
                                                         rec.data_ptr(), lse2.data_ptr(), status.data_ptr(), n, V,
                                                         _dtype_code(teacher), float(T), float(alpha), _stream_ptr(dev)))
        return rec, lse2, status

    def teacher_stats_peers(self, teacher, labels, T, alpha, V, peer_ptrs_dev, n_peers, row_offset):
        """Phase A that stores its records into every rank's buffer (peer-mapped memory) instead of a local one."""
        dev = teacher.device
        n = teacher.shape[0]
        lse2 = torch.empty((n,), dtype=torch.float32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_loca_teacher_stats_peers(teacher.data_ptr(), teacher.stride(0), labels.data_ptr(),
                                                               int(peer_ptrs_dev), int(n_peers), int(row_offset), lse2.data_ptr(),
                                                               status.data_ptr(), n, V, _dtype_code(teacher), float(T), float(alpha),
                                                               _stream_ptr(dev)))
        return lse2, status

    def build_table(self, records_all, V):
        dev = records_all.device
        table = torch.empty((V,), dtype=torch.float32, device=dev)
        consts = torch.empty((2,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ws_bytes = _lib.size_query("kd_loca_table_workspace_bytes", V)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(_lib.load().kd_loca_build_table(records_all.data_ptr(), records_all.shape[0], V, table.data_ptr(),
                                                       consts.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(dev)))
        return table, consts

    def emit(self, teacher, student, lse2_t, table, consts, T, eps, inv_norm, upstream, want_grad=True,
             force_streaming=False, status=None, saved=None):
        """``saved`` (a dict) receives the buffers a later re-emission needs (:meth:`regrad`)."""
        dev = student.device
        n, V = student.shape
        loss = torch.empty((), dtype=torch.float32, device=dev)
        rowaux = torch.empty((n + 1, 4), dtype=torch.float32, device=dev)     # +1: scratch entry (row queue)
        grad = torch.empty((n, V), dtype=student.dtype, device=dev) if want_grad else None
        up_t, up_ptr = _upstream_ptr(upstream, dev)
        flags = (0 if want_grad else _lib.FLAG_NO_GRAD) | (_lib.FLAG_FORCE_STREAMING if force_streaming else 0)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_loca_emit(teacher.data_ptr(), teacher.stride(0), student.data_ptr(), student.stride(0),
                                                lse2_t.data_ptr(), table.data_ptr(), consts.data_ptr(),
                                                grad.data_ptr() if want_grad else None, V, loss.data_ptr(),
                                                rowaux.data_ptr(), n, V, _dtype_code(student), float(T), float(eps),
                                                float(inv_norm), up_ptr, status.data_ptr() if status is not None else None,
                                                flags, _stream_ptr(dev)))
        if saved is not None:
            saved.update(teacher=teacher, student=student, table=table, rowaux=rowaux, T=float(T), eps=float(eps), inv_norm=float(inv_norm))
        return loss, grad

    def regrad(self, saved, grad, upstream, skip_if_one=None):
        """Re-emit the local gradient from a forward's ``saved`` buffers with ``upstream`` (device scalar) as its scale: one rounding
        to the output type (``kd_loca_regrad``); a no-op launch when ``*skip_if_one == 1``."""
        t, s = saved["teacher"], saved["student"]
        dev = s.device
        n, V = s.shape
        cel, ceaux, inv_count = saved.get("ce_labels"), saved.get("ceaux"), saved.get("inv_count")
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_loca_regrad(t.data_ptr(), t.stride(0), s.data_ptr(), s.stride(0), saved["table"].data_ptr(),
                                                  saved["rowaux"].data_ptr(), cel.data_ptr() if cel is not None else None,
                                                  inv_count.data_ptr() if cel is not None else None, float(saved.get("ce_weight", 0.0)),
                                                  ceaux.data_ptr() if cel is not None else None, grad.data_ptr(), grad.stride(0), n, V,
                                                  _dtype_code(s), saved["T"], saved["eps"], saved["inv_norm"], upstream.data_ptr(),
                                                  skip_if_one.data_ptr() if skip_if_one is not None else None, _stream_ptr(dev)))
        return grad

    def emit_ce(self, teacher, student, lse2_t, table, consts, ce_labels, inv_count, ce_weight, T, eps, inv_norm, upstream,
                want_grad=True, status=None, saved=None):
        """Phase C with the student's cross-entropy of the local rows fused in (``kd_loca_emit_ce``): ``(loss, ce, grad)``."""
        dev = student.device
        n, V = student.shape
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ce = torch.empty((), dtype=torch.float32, device=dev)
        rowaux = torch.empty((n + 1, 4), dtype=torch.float32, device=dev)
        ceaux = torch.empty((max(n, 1), 2), dtype=torch.float32, device=dev)
        grad = torch.empty((n, V), dtype=student.dtype, device=dev) if want_grad else None
        up_t, up_ptr = _upstream_ptr(upstream, dev)
        if status is None:
            status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_loca_emit_ce(teacher.data_ptr(), teacher.stride(0), student.data_ptr(), student.stride(0),
                                                   lse2_t.data_ptr(), table.data_ptr(), consts.data_ptr(), ce_labels.data_ptr(),
                                                   inv_count.data_ptr(), float(ce_weight), grad.data_ptr() if want_grad else None, V,
                                                   loss.data_ptr(), ce.data_ptr(), rowaux.data_ptr(), ceaux.data_ptr(), n, V,
                                                   _dtype_code(student), float(T), float(eps), float(inv_norm), up_ptr,
                                                   status.data_ptr(), 0 if want_grad else _lib.FLAG_NO_GRAD, _stream_ptr(dev)))
        if saved is not None:
            saved.update(teacher=teacher, student=student, table=table, rowaux=rowaux, T=float(T), eps=float(eps), inv_norm=float(inv_norm),
                         ce_labels=ce_labels, inv_count=inv_count, ce_weight=float(ce_weight), ceaux=ceaux)
        return loss, ce, grad


def shard_rows(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced token partition: ranks [0, n_total % world) get one extra row."""
    base, extra = divmod(n_total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def exchange_counts(n_local: int, device, group=None) -> list[int]:
    """Row count of every rank, in rank order.  One small all-gather plus a host read-back: call it once when the shard
    sizes are fixed, or pass ``counts=`` explicitly every step when they vary (each rank then knows them from the data
    pipeline).  Nothing is cached here: ranks that resize independently must not diverge in the collectives they issue."""
    world = dist.get_world_size(group)
    n = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    out = torch.empty((world,), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, n, group=group)
    return [int(c) for c in out.tolist()]


def gather_records(records: torch.Tensor, group=None, counts=None) -> torch.Tensor:
    """All-gather ragged [n_i, 4] int32 record blocks in rank order (global flattened token order).

    ``counts`` = rows of every rank; when omitted they are exchanged on every call (:func:`exchange_counts`, one host
    read-back) — never cached, so shards may change size from step to step on any subset of ranks."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return records
    world = dist.get_world_size(group)
    if world == 1:
        return records
    if counts is None:
        counts = exchange_counts(records.shape[0], records.device, group)
    counts = [int(c) for c in counts]
    if len(counts) != world or counts[dist.get_rank(group)] != records.shape[0]:
        raise ValueError(f"counts {counts} do not describe this rank's {records.shape[0]} rows / a world of {world}")
    if len(set(counts)) == 1:            # even shards: one contiguous all-gather, no padding or concatenation
        out = torch.empty((world * counts[0], 4), dtype=records.dtype, device=records.device)
        dist.all_gather_into_tensor(out, records.contiguous(), group=group)
        return out
    n_max = max(counts)
    padded = torch.zeros((n_max, 4), dtype=records.dtype, device=records.device)
    padded[: records.shape[0]] = records
    blocks = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded, group=group)
    return torch.cat([b[:c] for b, c in zip(blocks, counts)], dim=0).contiguous()


class PeerRecordExchange:
    """Record all-gather fused into phase A over NVLink peer memory (torch symmetric memory).

    Every rank owns two [n_total, 4] int32 record buffers in symmetric memory, used alternately step by step; the phase-A
    kernel of rank g stores its records at rows [offset_g, offset_g + n_g) of EVERY rank's current buffer (16-byte peer stores),
    and a device-side barrier on the launch stream separates phase A from the table build.  Two buffers are enough: to write the
    buffer of step i+2 a rank must have passed the barrier of step i+1, which every rank reaches only after its table build of
    step i has read that buffer.
    """

    def __init__(self, counts, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.counts = [int(c) for c in counts]
        if len(self.counts) != self.world:
            raise ValueError(f"counts {self.counts} do not describe a world of {self.world}")
        self.offset, self.n_total = sum(self.counts[: self.rank]), sum(self.counts)
        self.bufs, self.handles = [], []
        for _ in range(2):
            buf = symm_mem.empty((self.n_total, 4), dtype=torch.int32, device=device)
            self.handles.append(symm_mem.rendezvous(buf, self.group.group_name))
            self.bufs.append(buf)
        self.step = 0

    def run(self, phases, teacher, labels, T, alpha, V):
        """Returns (records of ALL tokens in global order, lse2_t of the local rows)."""
        i = self.step & 1
        self.step += 1
        hdl = self.handles[i]
        if teacher.shape[0] != self.counts[self.rank]:
            raise ValueError(f"this exchange was built for {self.counts[self.rank]} local rows, got {teacher.shape[0]}")
        lse2_t, status = phases.teacher_stats_peers(teacher, labels, T, alpha, V, hdl.buffer_ptrs_dev, self.world, self.offset)
        hdl.barrier(channel=0)
        return self.bufs[i], lse2_t, status


_peer_exchanges: dict = {}       # (group, device, counts of every rank) -> PeerRecordExchange; keyed by ALL ranks' counts


def _sum_status(status, group, multi):
    """A bad label on any rank must poison every rank's loss (the reference raises, P1:166): OR = max of the status words."""
    if status is not None and multi:
        dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group)
    return status


def loca_kl_token_sharded(teacher, student, labels, T, alpha=0.8, eps=1e-8, group=None, upstream=None,
                          phases=None, reduce_loss=True, want_grad=True, n_total=None, exchange="nccl", counts=None,
                          ce_labels=None, ce_weight=1.0, ce_valid_count=None, saved=None):
    """As-written LoCa over tokens sharded across the ranks of ``group``.

    ``teacher``/``student``/``labels`` hold this rank's rows.  Returns ``(loss, grad_local)`` where
    ``loss`` is the GLOBAL loss (identical on every rank) when ``reduce_loss`` and the local partial otherwise.
    ``counts``: rows of every rank in rank order; pass it when shard sizes vary from step to step (sequence lengths do) —
    without it the counts are exchanged on every call (one host read-back), nothing is cached across calls.
    ``phases`` injects the three phase implementations (tests run the host logic on CPU with an
    oracle-backed stand-in; the product default is :class:`CudaPhases`).  ``exchange="peer"`` replaces the NCCL
    all-gather of the records by peer stores issued from the phase-A kernel itself (:class:`PeerRecordExchange`).
    A label outside ``[0, V)`` on ANY rank NaN-poisons the loss on every rank (the reference raises at P1:166).

    ``ce_labels`` (this rank's next-token targets, ``-100`` = ignored; :func:`functional.shift_labels`) adds the student's
    cross-entropy from the same pass over the local rows (``kd_loca_emit_ce``) and changes the return value
    to ``(loss, ce, grad_local)`` with ``loss = KD + ce_weight * CE``.  The CE is the mean over the valid targets of ALL ranks —
    the concatenated batch, like the KD term — unless ``ce_valid_count`` (a number or device scalar) gives another divisor: the
    count is summed across the ranks on the device, nothing is read back.
    """
    if phases is None:
        phases = CudaPhases()
        _require_cuda(teacher, student, labels)
    t2, s2 = _rows(teacher), _rows(student)
    V = s2.shape[1]
    lab = labels.reshape(-1).to(torch.int64).contiguous()
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if multi and counts is None:
        counts = exchange_counts(t2.shape[0], t2.device, group)
    if exchange == "peer" and multi:
        # phase A + all-gather in one kernel: records go straight into every rank's buffer over NVLink
        key = (id(group), t2.device.index, tuple(int(c) for c in counts))
        ex = _peer_exchanges.get(key)
        if ex is None:
            ex = _peer_exchanges[key] = PeerRecordExchange(counts, t2.device, group)
        rec_all, lse2_t, status = ex.run(phases, t2, lab, T, alpha, V)
    else:
        rec, lse2_t, status = phases.teacher_stats(t2, lab, T, alpha, V)
        rec_all = gather_records(rec, group, counts) if multi else rec
    status = _sum_status(status, group, multi)
    if n_total is None:
        n_total = rec_all.shape[0]
    table, consts = phases.build_table(rec_all, V)
    inv_norm = 1.0 / (float(n_total) * float(V))
    if ce_labels is not None:
        cel = ce_labels.reshape(-1).to(torch.int64).contiguous()
        if cel.numel() != t2.shape[0]:
            raise ValueError(f"ce_labels has {cel.numel()} entries for {t2.shape[0]} local tokens")
        if ce_valid_count is None:
            count = (cel != -100).sum().to(torch.float32).reshape(1)
            if multi:
                dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
        else:
            count = torch.as_tensor(ce_valid_count, device=cel.device).to(torch.float32).reshape(1)
        inv_count = (1.0 / count).contiguous()              # 0 valid targets -> inf -> NaN, like F.cross_entropy
        kw = {} if saved is None else {"saved": saved}       # buffers for a re-emission in backward (CudaPhases.regrad)
        loss, ce, grad = phases.emit_ce(t2, s2, lse2_t, table, consts, cel, inv_count, ce_weight, T, eps, inv_norm, upstream,
                                        want_grad=want_grad, status=status, **kw)
        if reduce_loss and multi:
            both = torch.stack([loss, ce])
            dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
            loss, ce = both[0].clone(), both[1].clone()      # separate tensors: the autograd front marks the CE non-differentiable
        return loss, ce, grad
    kw = {} if saved is None else {"saved": saved}
    loss, grad = phases.emit(t2, s2, lse2_t, table, consts, T, eps, inv_norm, upstream, want_grad=want_grad, status=status, **kw)
    if reduce_loss and multi:
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss, grad


class _TokenShardedLoCa(torch.autograd.Function):
    """Differentiable front of :func:`loca_kl_token_sharded` (what ``patch_reference(..., group=...)`` binds): the gradient of
    the local rows is produced in the forward launch with ``upstream`` folded in; ``backward`` launches the re-emission kernel
    (``kd_loca_regrad``) with ``grad_output`` as the scale — it exits at once when ``grad_output / upstream`` is 1, the usual
    ``loss.backward()``, and otherwise recomputes the gradient from the saved row statistics so that it is rounded ONCE (P1:260
    multiplies the loss by gamma: rescaling the rounded bf16 gradient would miss the 2e-3 gate).  A second backward through the same
    node re-emits into a fresh buffer."""

    @staticmethod
    def forward(ctx, teacher, student, labels, T, alpha, eps, group, upstream, phases, exchange, counts,
                ce_labels=None, ce_weight=1.0, ce_valid_count=None):
        need_grad = student.requires_grad
        ph = phases if phases is not None else CudaPhases()
        saved = {} if (need_grad and hasattr(ph, "regrad")) else None
        res = loca_kl_token_sharded(teacher.detach(), student.detach(), labels, T, alpha, eps, group=group, upstream=upstream,
                                    phases=phases, want_grad=need_grad, exchange=exchange, counts=counts,
                                    ce_labels=ce_labels, ce_weight=ce_weight, ce_valid_count=ce_valid_count, saved=saved)
        loss, grad = res[0], res[-1]
        ctx.grad, ctx.shape, ctx.need_grad = grad, student.shape, need_grad
        ctx.saved, ctx.phases = saved, ph
        ctx.upstream = None if upstream is None else torch.as_tensor(upstream, dtype=torch.float32, device=student.device).reshape(())
        if ce_labels is None:
            return loss
        ctx.mark_non_differentiable(res[1])
        return loss, res[1]

    @staticmethod
    def backward(ctx, grad_output, _grad_ce=None):
        if not ctx.need_grad:
            return (None,) * 14
        if ctx.saved:                             # CUDA phases: re-emission from the saved row statistics, one rounding
            s2 = ctx.saved["student"]
            first = ctx.grad is not None
            go = grad_output.detach().to(device=s2.device, dtype=torch.float32).reshape(1).contiguous()
            ratio = (go if ctx.upstream is None else go / ctx.upstream).contiguous()
            g = ctx.grad if first else torch.empty(s2.shape, dtype=s2.dtype, device=s2.device)
            ctx.grad = None
            ctx.phases.regrad(ctx.saved, g, go, skip_if_one=ratio if first else None)
            return (None, g.reshape(ctx.shape)) + (None,) * 12
        if ctx.grad is None:
            raise RuntimeError("loca_kl_token_sharded_autograd: backward through the same graph twice (the gradient is rescaled in place)")
        g, ctx.grad = ctx.grad, None
        go = grad_output.detach().to(device=g.device, dtype=torch.float32).reshape(())
        ratio = (go if ctx.upstream is None else go / ctx.upstream).contiguous()
        if g.is_cuda:
            with torch.cuda.device(g.device):
                _lib.check(_lib.load().kd_scale_rows(g.data_ptr(), g.stride(0), g.shape[0], g.shape[1], _dtype_code(g), ratio.data_ptr(),
                                                     _stream_ptr(g.device)))
        else:                                     # host-side test doubles only (the product path is CUDA)
            g = (g * ratio.to(g.dtype)).to(g.dtype)
        return (None, g.reshape(ctx.shape)) + (None,) * 12


def loca_kl_ce_token_sharded_autograd(teacher, student, labels, ce_labels, T, alpha=0.8, eps=1e-8, ce_weight=1.0, group=None,
                                      upstream=None, phases=None, exchange="nccl", counts=None, ce_valid_count=None):
    """``(KD + ce_weight * CE, CE)`` of the concatenated batch over token shards, both from one pass over this rank's logits; the
    gradient w.r.t. the local student rows flows through the first result (the CE alone is detached, for logging)."""
    return _TokenShardedLoCa.apply(teacher, student, labels, float(T), float(alpha), float(eps), group, upstream, phases, exchange,
                                   counts, ce_labels, float(ce_weight), ce_valid_count)


def loca_kl_token_sharded_autograd(teacher, student, labels, T, alpha=0.8, eps=1e-8, group=None, upstream=None, phases=None,
                                   exchange="nccl", counts=None):
    """As-written LoCa over token shards as a differentiable loss: the GLOBAL loss value on every rank, the gradient w.r.t. this
    rank's student rows through autograd.  Same arguments as :func:`loca_kl_token_sharded`."""
    return _TokenShardedLoCa.apply(teacher, student, labels, float(T), float(alpha), float(eps), group, upstream, phases, exchange, counts,
                                   None, 1.0, None)


def per_row_token_sharded(fn, teacher, student, *args, group=None, reduce_loss=True, n_total=None, **kw):
    """Row-independent variants: run ``fn`` (``loca_kl(mode='per_row')`` / ``temperature_kl``) on the local rows
    normalised by the global token count; the only data-path collective is the scalar loss sum.  ``n_total`` = global
    token count; when omitted it is all-reduced on every call (one host read-back, nothing cached)."""
    n_local = _rows(student).shape[0]
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if n_total is None:
        n_total = n_local
        if multi:
            t = torch.tensor([n_local], dtype=torch.int64, device=student.device)
            dist.all_reduce(t, group=group)
            n_total = int(t.item())
    loss = fn(teacher, student, *args, n_total=int(n_total), **kw)
    if reduce_loss and multi:
        red = loss.detach().clone()
        dist.all_reduce(red, group=group)
        loss = loss + (red - loss.detach())       # value = global sum, gradient = local path
    return loss


# =====================================================================================================================
# Vocabulary-sharded execution (BASELINE.json configs[4]): every rank holds ALL rows but only columns
# [col0, col0 + V_local) of both logit tensors — what a vocabulary-parallel lm_head produces.  Two exchanges per step:
# an all-gather of 48-byte per-row statistics and an all-reduce of one float per row (Q_r); the loss is a scalar sum.
# =====================================================================================================================

_VS_VARIANTS = {"reference": _lib.LOCA_REFERENCE, "per_row": _lib.LOCA_PER_ROW, "kl": _lib.KL_PLAIN,
                "kl_log_target": _lib.KL_LOG_TARGET}
_LN2 = 0.6931471805599453


class CudaVocabPhases:
    """The four phase calls of the vocabulary-sharded path (kd_vshard_stats / _combine / _accum / _emit)."""

    def stats(self, teacher, student, labels, col0, T, student_partials=None):
        """``student_partials`` ([tiles, n, 2], from ``lm_head_fwd``): the student half comes from the GEMM epilogue and the
        student logits are not read here."""
        dev = student.device
        n, v_local = student.shape
        out = torch.empty((n, _lib.VSHARD_STAT_WORDS), dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.kd_vshard_stats(teacher.data_ptr(), teacher.stride(0),
                                           student.data_ptr() if student_partials is None else None, student.stride(0),
                                           labels.data_ptr() if labels is not None else None, int(col0), out.data_ptr(),
                                           n, v_local, _dtype_code(student), float(T), _stream_ptr(dev)))
            if student_partials is not None:
                if tuple(student_partials.shape) != (lib.kd_lm_head_stat_tiles(v_local), n, 2) or not student_partials.is_contiguous():
                    raise ValueError("student_partials must be the contiguous [tiles, n_rows, 2] block lm_head_fwd returned")
                _lib.check(lib.kd_lm_head_stats_merge(student_partials.data_ptr(), student_partials.shape[0], out.data_ptr(), n,
                                                      float(T), _stream_ptr(dev)))
        return out

    def combine(self, stats_all, labels, V_total, T, alpha, variant):
        dev = stats_all.device
        g, n = stats_all.shape[0], stats_all.shape[1]
        rec = torch.empty((n, 4), dtype=torch.int32, device=dev)
        lse2 = torch.empty((n, 2), dtype=torch.float32, device=dev)
        rowconst = torch.empty((n,), dtype=torch.float32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_vshard_combine(stats_all.data_ptr(), g, n, labels.data_ptr() if labels is not None else None,
                                                     int(V_total), float(T), float(alpha), variant, rec.data_ptr(), lse2.data_ptr(),
                                                     rowconst.data_ptr(), status.data_ptr(), _stream_ptr(dev)))
        return rec, lse2, rowconst, status

    def build_table(self, records, V_total):
        return CudaPhases().build_table(records, V_total)

    def accum(self, teacher, student, lse2, rec, table, col0, T, eps, variant):
        dev = student.device
        n, v_local = student.shape
        partial = torch.empty((n, 2), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_vshard_accum(teacher.data_ptr(), teacher.stride(0), student.data_ptr(), student.stride(0),
                                                   lse2.data_ptr(), rec.data_ptr(), table.data_ptr() if table is not None else None,
                                                   int(col0), partial.data_ptr(), n, v_local, _dtype_code(student), float(T),
                                                   float(eps), variant, _stream_ptr(dev)))
        return partial

    def emit(self, teacher, student, lse2, q_rows, rec, table, col0, T, eps, variant, inv_norm, upstream):
        dev = student.device
        n, v_local = student.shape
        grad = torch.empty((n, v_local), dtype=student.dtype, device=dev)
        up_t, up_ptr = _upstream_ptr(upstream, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().kd_vshard_emit(teacher.data_ptr(), teacher.stride(0), student.data_ptr(), student.stride(0),
                                                  lse2.data_ptr(), q_rows.data_ptr(), rec.data_ptr(),
                                                  table.data_ptr() if table is not None else None, int(col0), grad.data_ptr(),
                                                  grad.stride(0), n, v_local, _dtype_code(student), float(T), float(eps), variant,
                                                  float(inv_norm), up_ptr, _stream_ptr(dev)))
        return grad


class DistCollectives:
    """The two data-path exchanges + the scalar loss sum over ``group`` (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def gather_stats(self, stats: torch.Tensor) -> torch.Tensor:
        if not self.active:
            return stats.unsqueeze(0)
        world = dist.get_world_size(self.group)
        out = torch.empty((world * stats.shape[0],) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
        dist.all_gather_into_tensor(out, stats.contiguous(), group=self.group)      # concatenated along dim 0, shard-major
        return out.view((world,) + tuple(stats.shape))

    def sum_(self, x: torch.Tensor) -> torch.Tensor:
        if self.active:
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    def sum_async(self, x: torch.Tensor):
        """In-place sum started on the collective's own stream; ``.wait()`` on the returned handle (None when there is nothing
        to exchange) makes the current stream wait for it — kernels issued in between overlap the exchange."""
        if self.active:
            return dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return None


def logit_kd_vocab_sharded(teacher_shard, student_shard, labels, col0, V_total, T, mode="reference", alpha=0.8, eps=1e-8,
                           n_rows_norm=None, reduction="mean", upstream=None, group=None, phases=None, collectives=None,
                           want_grad=True, student_partials=None):
    """Logit KD over logits sharded along the vocabulary.

    ``teacher_shard`` / ``student_shard``: this rank's columns ``[col0, col0 + V_local)`` of all rows; ``labels``: all rows'
    labels (LoCa modes), global class indices.  ``mode``: ``"reference"`` (as written, P1:184-185), ``"per_row"``, ``"kl"``
    (P1:331-343) or ``"kl_log_target"`` (FB:206-219).  Returns ``(loss, grad_shard)``: the GLOBAL loss (identical on every
    rank) and the gradient w.r.t. this rank's student columns.  ``student_partials``: the per-tile student statistics
    ``lm_head_fwd(..., T=T)`` produced together with ``student_shard`` (the stats phase then streams only the teacher).
    """
    variant = _VS_VARIANTS[mode]
    if phases is None:
        phases = CudaVocabPhases()
        _require_cuda(teacher_shard, student_shard, labels)
    if collectives is None:
        collectives = DistCollectives(group)
    t2, s2 = _rows(teacher_shard), _rows(student_shard)
    n = s2.shape[0]
    loca = variant in (_lib.LOCA_REFERENCE, _lib.LOCA_PER_ROW)
    lab = labels.reshape(-1).to(torch.int64).contiguous() if (loca and labels is not None) else None
    if loca and lab is None:
        raise ValueError("the LoCa modes need labels")
    if reduction == "mean":
        inv_norm = 1.0 / (float(n if n_rows_norm is None else n_rows_norm) * float(V_total))
    elif reduction == "batchmean":
        inv_norm = 1.0 / float(n if n_rows_norm is None else n_rows_norm)
    else:
        raise ValueError(f"unknown reduction {reduction!r}")

    if student_partials is not None:                                      # student (max, sum-exp) left by the lm_head GEMM
        stats = phases.stats(t2, s2, lab, col0, T, student_partials=student_partials)
    else:
        stats = phases.stats(t2, s2, lab, col0, T)
    stats_all = collectives.gather_stats(stats)                             # exchange 1: [G, n, 12]
    rec, lse2, rowconst, status = phases.combine(stats_all, lab, V_total, T, alpha, variant)
    table, table_const = None, 0.0
    if variant == _lib.LOCA_REFERENCE:
        table, consts = phases.build_table(rec, V_total)
        table_const = consts[0].double() if consts is not None else 0.0
    partial = phases.accum(t2, s2, lse2, rec, table, col0, T, eps, variant)
    q_rows = collectives.sum_(partial[:, 1].contiguous())                    # exchange 2: Q_r
    grad = phases.emit(t2, s2, lse2, q_rows, rec, table, col0, T, eps, variant, inv_norm, upstream) if want_grad else None
    unit = 1.0 if variant == _lib.KL_LOG_TARGET else _LN2
    rows = collectives.sum_(partial[:, 0].double().sum().reshape(1) * unit)[0]
    loss = (rows + rowconst.double().sum() * _LN2 + float(n) * table_const) * (float(T) * float(T) * inv_norm)
    if loca and status is not None:          # a label outside [0, V_total): the reference raises (P1:166) — NaN, no host sync
        loss = torch.where(status.reshape(()) != 0, torch.full_like(loss, float("nan")), loss)
    return loss.float(), grad


# ---------------------------------------------------------------------------------------------------------------------
# Vocabulary-parallel student lm_head feeding the vocabulary-sharded KD loss (BASELINE.json configs[4]):
# logits_shard = hidden @ weight_shard^T is produced per rank by the hand-written tcgen05 kernel (kd_lm_head_fwd,
# csrc/kd_gemm.cu), whose epilogue also leaves the student's per-tile (max, sum-exp) so that the stats phase streams the
# teacher only; the loss and d loss / d logits_shard come from logit_kd_vocab_sharded, and the two weight-side products of the
# backward (the same tcgen05 kernel, MN-major operands) close the loop — dH needs the sum over the vocabulary shards: one all-reduce
# of the fp32 partial [n_rows, hidden], started under the dW product.  Keeping the logits out of HBM altogether does not pay at
# this shape: the recomputed product costs what the logit traffic it removes does.
# ---------------------------------------------------------------------------------------------------------------------
def lm_head_fwd(hidden, weight, T=None, out=None):
    """``hidden`` [n_rows, H] x ``weight`` [V_local, H] -> (``logits`` [n_rows, V_local], ``partials`` or None) through
    ``kd_lm_head_fwd``.  With ``T`` the epilogue also returns the per-128-column-tile student (max, sum-exp at temperature T)
    block [tiles, n_rows, 2] that ``logit_kd_vocab_sharded(student_partials=...)`` consumes.  bf16 / fp16 CUDA tensors only."""
    dev = _require_cuda(hidden, weight)
    if hidden.dim() != 2 or weight.dim() != 2 or hidden.shape[1] != weight.shape[1] or hidden.dtype != weight.dtype:
        raise ValueError("lm_head_fwd: hidden [n_rows, H] and weight [V_local, H] of one dtype expected")
    h = hidden if hidden.stride(1) == 1 else hidden.contiguous()
    w = weight if weight.stride(1) == 1 else weight.contiguous()
    n, v_local = h.shape[0], w.shape[0]
    lib = _lib.load()
    logits = out if out is not None else torch.empty((n, v_local), dtype=h.dtype, device=dev)
    if logits.shape != (n, v_local) or logits.dtype != h.dtype or logits.stride(1) != 1:
        raise ValueError("lm_head_fwd: `out` must be a row-major [n_rows, V_local] tensor of the operand dtype")
    partials = None
    if T is not None:
        partials = torch.empty((lib.kd_lm_head_stat_tiles(v_local), n, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.kd_lm_head_fwd(h.data_ptr(), h.stride(0), w.data_ptr(), w.stride(0), logits.data_ptr(), logits.stride(0),
                                      partials.data_ptr() if partials is not None else None, n, v_local, h.shape[1],
                                      _dtype_code(h), float(T) if T is not None else 1.0, _stream_ptr(dev)))
    return logits, partials


_MAX_RESERVED_SMS = 16      # SMs the d weight product may leave to the d hidden all-reduce (0: never cap the product)


def _dw_cta_cap(dev, m: int, n: int) -> int:
    """CTAs for the [m, n] d weight product when a collective runs beside it.  The persistent kernel works through its 256 x 256
    tiles in waves of one tile per CTA pair; 300 tiles (18 992 x 896) take five waves on 74 pairs and on 60 — so 28 of the 148 SMs
    can go to the all-reduce at no cost to the product.  Measured on 8 GPUs, two boxes (tools/vshard_overlap_check.py): blocking
    all-reduce 9.58 / 9.32 ms per step, asynchronous with every SM held by the product 9.62 / 9.68 (the collective cannot start:
    no gain), with 16 SMs free 9.28 / 9.15, with 28 free 9.31, with 32 free and a sixth wave 10.14 — a gain of ~2 %, close to the
    run-to-run spread (0.15 ms), kept because it never costs a wave."""
    if _MAX_RESERVED_SMS <= 0:
        return 0
    pairs = torch.cuda.get_device_properties(dev).multi_processor_count // 2
    tiles = ((m + 255) // 256) * ((n + 255) // 256)
    waves = (tiles + pairs - 1) // pairs
    keep = max((tiles + waves - 1) // waves, pairs - _MAX_RESERVED_SMS // 2, 1)      # fewest pairs with the same number of waves
    return 0 if keep >= pairs else 2 * keep


def _tensor_core_head_applies(hidden, weight) -> bool:
    return (hidden.is_cuda and hidden.dtype in (torch.bfloat16, torch.float16) and weight.dtype == hidden.dtype
            and hidden.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0)


class _LmHeadVocabShardedKD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight_shard, teacher_shard, labels, col0, V_total, T, mode, alpha, eps, n_rows_norm, reduction,
                group, phases, collectives, gemm):
        partials = None
        ctx.gemm = "library"
        if gemm == "tcgen05" and phases is None and _tensor_core_head_applies(hidden, weight_shard):
            ctx.gemm = "tcgen05"
            # (measured: running the teacher's half of the statistics beside this GEMM on a second stream gains nothing —
            # 5.85 against 5.80 ms for head + KD forward; the power-bound GEMM slows by what the overlap saves)
            logits, partials = lm_head_fwd(hidden.detach(), weight_shard.detach(), T=T)
        elif gemm == "tcgen05" and phases is None and hidden.is_cuda and hidden.dtype != torch.float32:
            raise ValueError("lm_head_kd_vocab_sharded: hidden size and V_local must be multiples of 8 for the tensor-core head "
                             "(pass gemm='library' otherwise)")
        else:                                                                       # fp32 operands / host-side test doubles
            logits = torch.matmul(hidden.detach(), weight_shard.detach().t())      # [n_rows, V_local], the student's dtype
        loss, grad_logits = logit_kd_vocab_sharded(teacher_shard, logits, labels, col0, V_total, T, mode=mode, alpha=alpha, eps=eps,
                                                   n_rows_norm=n_rows_norm, reduction=reduction, group=group, phases=phases,
                                                   collectives=collectives, student_partials=partials)
        ctx.save_for_backward(hidden, weight_shard)
        ctx.grad_logits = grad_logits                                               # ours: rescaled in place by the backward
        ctx.collectives = collectives if collectives is not None else DistCollectives(group)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        hidden, weight_shard = ctx.saved_tensors
        grad_logits, ctx.grad_logits = ctx.grad_logits, None
        if grad_logits is None:
            raise RuntimeError("lm_head_kd_vocab_sharded: backward through the same graph twice")
        if grad_logits.is_cuda and grad_logits.dtype == hidden.dtype and grad_logits.stride(1) == 1:
            # d loss / d logits *= grad_output on the device, skipped there when it is 1 (the usual loss.backward())
            dev = grad_logits.device
            go = grad_output.detach().to(device=dev, dtype=torch.float32).reshape(()).contiguous()
            with torch.cuda.device(dev):
                _lib.check(_lib.load().kd_scale_rows(grad_logits.data_ptr(), grad_logits.stride(0), grad_logits.shape[0],
                                                     grad_logits.shape[1], _dtype_code(grad_logits), go.data_ptr(), _stream_ptr(dev)))
            g = grad_logits
        else:
            g = grad_logits.to(hidden.dtype) * grad_output.to(hidden.dtype)
        if ctx.gemm == "tcgen05":
            # both backward products on the tcgen05 kernel, operands read in place through MN-major descriptors (kd_gemm):
            # the partial dH leaves the kernel in fp32 and is summed over the vocabulary shards in fp32, rounded ONCE
            # The all-reduce of the partial dH (117 MB at 65 536 x 896) travels while the dW product runs.
            d_hidden = tc_gemm(g, weight_shard.detach(), b_mn_major=True, out_dtype=torch.float32)       # [n, V_l] . [V_l, H]
            start = getattr(ctx.collectives, "sum_async", None)
            work = start(d_hidden) if start is not None else None
            if start is None:
                d_hidden = ctx.collectives.sum_(d_hidden)
            # ... on a few SMs fewer than the machine has: the persistent kernel would otherwise keep every SM to the end
            d_weight = tc_gemm(g, hidden.detach(), a_mn_major=True, b_mn_major=True,                     # [n, V_l]^T . [n, H]
                               max_ctas=_dw_cta_cap(g.device, g.shape[1], hidden.shape[1]) if work is not None else 0)
            if work is not None:
                work.wait()
            d_hidden = d_hidden.to(hidden.dtype)
        else:
            try:                                    # library GEMM with an fp32 result
                d_hidden = torch.mm(g, weight_shard.detach(), out_dtype=torch.float32)
            except (TypeError, RuntimeError):
                d_hidden = torch.matmul(g.float(), weight_shard.detach().float())
            d_hidden = ctx.collectives.sum_(d_hidden.contiguous()).to(hidden.dtype)                   # fp32 partials, one rounding
            d_weight = torch.matmul(g.t(), hidden.detach())
        return (d_hidden, d_weight) + (None,) * 14


def lm_head_kd_vocab_sharded(hidden, weight_shard, teacher_shard, labels, col0, V_total, T, mode="reference", alpha=0.8,
                             eps=1e-8, n_rows_norm=None, reduction="mean", group=None, phases=None, collectives=None,
                             gemm="tcgen05"):
    """KD loss of a vocabulary-parallel student ``lm_head``: ``hidden`` [n_rows, K] (replicated), ``weight_shard`` [V_local, K] (this
    rank's rows of the 151 936 x 896 head), ``teacher_shard`` [n_rows, V_local].  Differentiable w.r.t. ``hidden`` and
    ``weight_shard``; the returned loss is the global one.  ``gemm``: ``"tcgen05"`` (the hand-written kernel, statistics fused
    into its epilogue) or ``"library"`` (torch.matmul; also what fp32 operands use)."""
    if gemm not in ("tcgen05", "library"):
        raise ValueError(f"unknown gemm {gemm!r}")
    return _LmHeadVocabShardedKD.apply(hidden, weight_shard, teacher_shard, labels, int(col0), int(V_total), float(T), mode,
                                       float(alpha), float(eps), n_rows_norm, reduction, group, phases, collectives, gemm)
