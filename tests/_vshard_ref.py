"""Plain-torch fp64 stand-ins for the four vocabulary-sharded phases (test infrastructure; same interface as
sharded.CudaVocabPhases) and an in-memory collective for simulating several shards inside one process."""
from __future__ import annotations

import math
import threading

import torch

from oracle import kd_oracle as O

LOCA_REFERENCE, LOCA_PER_ROW, KL_PLAIN, KL_LOG_TARGET = 0, 1, 2, 3
LOG2E = 1.4426950408889634


class TorchVocabPhases:
    """stats: [n, 12] float64 (m_t, S_t, a1, i1, a2, i2, x_label, m_s, S_s, -, -, -); records: [n, 4] float64 (label, v1, k, v2)."""

    def stats(self, teacher, student, labels, col0, T):
        t, s = teacher.double(), student.double()
        n, vl = s.shape
        k2 = LOG2E / T
        out = torch.zeros((n, 12), dtype=torch.float64)
        for x, (im, isum) in ((t, (0, 1)), (s, (7, 8))):
            m = x.max(dim=1).values
            out[:, im] = m
            out[:, isum] = torch.exp2((x - m[:, None]) * k2).sum(dim=1)
        i1, i2 = O.top2_lowest_index(t)
        out[:, 2] = t.gather(1, i1[:, None])[:, 0]
        out[:, 3] = (i1 + col0).double()
        out[:, 4] = t.gather(1, i2[:, None])[:, 0] if vl > 1 else -math.inf
        out[:, 5] = (i2 + col0).double() if vl > 1 else 2.0 ** 40
        out[:, 6] = -math.inf
        if labels is not None:
            loc = labels.long() - col0
            own = (loc >= 0) & (loc < vl)
            out[own, 6] = t[own].gather(1, loc[own][:, None])[:, 0]
        return out

    def combine(self, stats_all, labels, V_total, T, alpha, variant):
        g, n = stats_all.shape[0], stats_all.shape[1]
        k2 = LOG2E / T

        def lse2(im, isum):
            m = stats_all[:, :, im].max(dim=0).values
            ssum = (stats_all[:, :, isum] * torch.exp2((stats_all[:, :, im] - m[None]) * k2)).sum(dim=0)
            return m * k2 + torch.log2(ssum)

        l2t, l2s = lse2(0, 1), lse2(7, 8)
        # global runner-up: merge the 2G candidates per row with "largest value, lowest index first"
        vals = torch.cat([stats_all[:, :, 2], stats_all[:, :, 4]], dim=0).T          # [n, 2G]
        idxs = torch.cat([stats_all[:, :, 3], stats_all[:, :, 5]], dim=0).T
        order = torch.argsort(idxs, dim=1)
        vals, idxs = vals.gather(1, order), idxs.gather(1, order)
        o2 = torch.argsort(vals, dim=1, descending=True, stable=True)
        a2 = vals.gather(1, o2)[:, 1]
        k = idxs.gather(1, o2)[:, 1].long()
        x_label = stats_all[:, :, 6].max(dim=0).values
        rec = torch.full((n, 4), math.nan, dtype=torch.float64)
        rec[:, 0] = -1
        rec[:, 2] = k.double()
        rowconst = torch.zeros(n, dtype=torch.float64)
        if variant in (LOCA_REFERENCE, LOCA_PER_ROW):
            p_gt, p_k = torch.exp2(x_label * k2 - l2t), torch.exp2(a2 * k2 - l2t)
            sc = alpha / (1.0 - p_gt + p_k)
            v1, v2 = 1.0 - sc * (1.0 - p_gt), sc * p_k
            rec[:, 0], rec[:, 1], rec[:, 3] = labels.double(), v1, v2
            if variant == LOCA_PER_ROW:
                rowconst = torch.where(k != labels, torch.xlogy(v1, v1), torch.zeros_like(v1)) / math.log(2.0) + torch.xlogy(v2, v2) / math.log(2.0)
        return rec, torch.stack([l2t, l2s], dim=1), rowconst, None

    def build_table(self, rec, V_total):
        table = O.build_override_table(rec[:, 0].long(), rec[:, 2].long(), rec[:, 1], rec[:, 3], V_total)
        sel = ~torch.isnan(table)
        return table, torch.stack([torch.xlogy(table[sel], table[sel]).sum(), table[sel].sum()])

    def _cells(self, teacher, student, lse2, rec, table, col0, T, eps, variant):
        t, s = teacher.double(), student.double()
        n, vl = s.shape
        k2 = LOG2E / T
        a = s * k2 - lse2[:, 1:2]
        b = t * k2 - lse2[:, 0:1]
        tp = torch.exp2(b)
        if variant == KL_PLAIN:
            return tp, tp * (b - a), torch.ones_like(tp, dtype=torch.bool), a
        if variant == KL_LOG_TARGET:
            w = torch.exp(tp)
            return w, w * (tp - a * math.log(2.0)), torch.ones_like(tp, dtype=torch.bool), a
        ov = torch.full_like(tp, math.nan)
        if variant == LOCA_REFERENCE:
            ov[:] = table[col0:col0 + vl][None]
        else:
            rows = torch.arange(n)
            for col, val in ((rec[:, 0], rec[:, 1]), (rec[:, 2], rec[:, 3])):       # label pass, then the runner-up pass
                loc = col.long() - col0
                own = (loc >= 0) & (loc < vl) & (rec[:, 0] >= 0)
                ov[rows[own], loc[own]] = val[own]
        over = ~torch.isnan(ov)
        q = torch.where(over, ov, tp)
        lq = torch.where(over, torch.zeros_like(b), b)
        thr2 = math.log2(eps)
        keep = a >= thr2
        return q, q * (lq - torch.maximum(a, torch.full_like(a, thr2))), keep, a

    def accum(self, teacher, student, lse2, rec, table, col0, T, eps, variant):
        q, term, keep, _ = self._cells(teacher, student, lse2, rec, table, col0, T, eps, variant)
        return torch.stack([term.sum(dim=1), (q * keep).sum(dim=1)], dim=1)

    def emit(self, teacher, student, lse2, q_rows, rec, table, col0, T, eps, variant, inv_norm, upstream):
        q, _, keep, a = self._cells(teacher, student, lse2, rec, table, col0, T, eps, variant)
        g = T * inv_norm * (1.0 if upstream is None else float(upstream))
        return g * (torch.exp2(a) * q_rows[:, None] - q * keep)


class ThreadCollectives:
    """all-gather / all-reduce between threads of one process (each thread plays one shard)."""

    def __init__(self, world):
        self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)

    def bind(self, rank):
        outer = self

        class Bound:
            def gather_stats(self, stats):
                outer.slots[rank] = stats
                if stats.is_cuda:
                    torch.cuda.synchronize()
                outer.barrier.wait()
                out = torch.stack(list(outer.slots), dim=0)
                outer.barrier.wait()
                return out

            def sum_(self, x):
                outer.slots[rank] = x.clone()
                if x.is_cuda:
                    torch.cuda.synchronize()
                outer.barrier.wait()
                out = torch.stack(list(outer.slots), dim=0).sum(dim=0)
                outer.barrier.wait()
                x.copy_(out)
                return x

        return Bound()
