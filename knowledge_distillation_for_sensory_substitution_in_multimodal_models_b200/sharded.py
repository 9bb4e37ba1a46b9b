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
             force_streaming=False):
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
                                                float(inv_norm), up_ptr, flags, _stream_ptr(dev)))
        return loss, grad


def shard_rows(n_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced token partition: ranks [0, n_total % world) get one extra row."""
    base, extra = divmod(n_total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


_shard_counts: dict = {}     # (group id, local rows) -> rows of every rank; shard sizes are exchanged once, not per step


def gather_records(records: torch.Tensor, group=None, counts=None) -> torch.Tensor:
    """All-gather ragged [n_i, 4] int32 record blocks in rank order (global flattened token order).

    The per-rank row counts need a host read-back, so they are exchanged only the first time a given local
    shard size is seen for ``group`` (pass ``counts`` explicitly when ranks resize independently)."""
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return records
    world = dist.get_world_size(group)
    if world == 1:
        return records
    key = (id(group), int(records.shape[0]))
    if counts is None:
        counts = _shard_counts.get(key)
    if counts is None:
        n_local = torch.tensor([records.shape[0]], dtype=torch.int64, device=records.device)
        gathered = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(gathered, n_local, group=group)
        counts = [int(c.item()) for c in gathered]
        _shard_counts[key] = counts
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


def loca_kl_token_sharded(teacher, student, labels, T, alpha=0.8, eps=1e-8, group=None, upstream=None,
                          phases=None, reduce_loss=True, want_grad=True, n_total=None):
    """As-written LoCa over tokens sharded across the ranks of ``group``.

    ``teacher``/``student``/``labels`` hold this rank's rows.  Returns ``(loss, grad_local)`` where
    ``loss`` is the GLOBAL loss (identical on every rank) when ``reduce_loss`` and the local partial otherwise.
    ``phases`` injects the three phase implementations (tests run the host logic on CPU with an
    oracle-backed stand-in; the product default is :class:`CudaPhases`).
    """
    if phases is None:
        phases = CudaPhases()
        _require_cuda(teacher, student, labels)
    t2, s2 = _rows(teacher), _rows(student)
    V = s2.shape[1]
    lab = labels.reshape(-1).to(torch.int64).contiguous()
    rec, lse2_t, _status = phases.teacher_stats(t2, lab, T, alpha, V)
    rec_all = gather_records(rec, group)
    if n_total is None:
        n_total = rec_all.shape[0]
    table, consts = phases.build_table(rec_all, V)
    inv_norm = 1.0 / (float(n_total) * float(V))
    loss, grad = phases.emit(t2, s2, lse2_t, table, consts, T, eps, inv_norm, upstream, want_grad=want_grad)
    if reduce_loss and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        loss = loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    return loss, grad


def per_row_token_sharded(fn, teacher, student, *args, group=None, reduce_loss=True, **kw):
    """Row-independent variants: run ``fn`` (``loca_kl(mode='per_row')`` / ``temperature_kl``) on the local rows
    normalised by the global token count; the only collective is the scalar loss sum."""
    n_local = _rows(student).shape[0]
    n_total = n_local
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([n_local], dtype=torch.int64, device=student.device)
        dist.all_reduce(t, group=group)
        n_total = int(t.item())
    loss = fn(teacher, student, *args, n_total=n_total, **kw)
    if reduce_loss and n_total != n_local:
        red = loss.detach().clone()
        dist.all_reduce(red, group=group)
        loss = loss + (red - loss.detach())       # value = global sum, gradient = local path
    return loss
