"""World-size-2 (gloo, CPU) test of the token-sharded LoCa host logic: ragged record all-gather in global token
order, identical table on every rank, global normaliser, loss all-reduce.  The three device phases are replaced by
oracle-backed stand-ins (tests may use oracle/); the CUDA phases are covered by the GPU tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kd_oracle as O
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded, synthetic

T, ALPHA, V = 2.0, 0.8, 257


class OraclePhases:
    """CPU stand-in for sharded.CudaPhases with the same record layout ([n,4] int32: label, v1 bits, k, v2 bits)."""

    def teacher_stats(self, teacher, labels, T, alpha, V):
        bad = (labels < 0) | (labels >= V)
        status = torch.tensor([1 if bool(bad.any()) else 0], dtype=torch.int32)      # KD_STATUS_BAD_LABEL, like phase A
        st = O.loca_row_stats(teacher, labels.clamp(0, V - 1), T, alpha, V, dtype=torch.float64)
        rec = torch.empty((labels.numel(), 4), dtype=torch.int32)
        rec[:, 0] = st["labels"].int()
        rec[:, 1] = st["v1"].float().view(torch.int32)
        rec[:, 2] = st["k"].int()
        rec[:, 3] = st["v2"].float().view(torch.int32)
        return rec, None, status

    def build_table(self, records_all, V):
        lab, k = records_all[:, 0].long(), records_all[:, 2].long()
        v1 = records_all[:, 1].contiguous().view(torch.float32).double()
        v2 = records_all[:, 3].contiguous().view(torch.float32).double()
        self.stats_all = {"labels": lab, "k": k, "v1": v1, "v2": v2}
        return O.build_override_table(lab, k, v1, v2, V), None

    def emit(self, teacher, student, lse2_t, table, consts, T, eps, inv_norm, upstream, want_grad=True, force_streaming=False,
             status=None):
        lab_local = torch.zeros(student.shape[0], dtype=torch.long)     # labels only matter through the table here
        res = O.loca_kl(teacher, student, lab_local, T, ALPHA, eps, mode=O.LOCA_REFERENCE, inv_norm=inv_norm,
                        stats_override=self.stats_all)
        loss = res.loss.float()
        if status is not None and int(status.item()) != 0:               # what loss_finalize_kernel does with the status word
            loss = torch.full_like(loss, float("nan"))
        return loss, res.grad


    def emit_ce(self, teacher, student, lse2_t, table, consts, ce_labels, inv_count, ce_weight, T, eps, inv_norm, upstream,
                want_grad=True, status=None):
        """kd_loca_emit_ce: the KD partial of emit() + sum over the local valid rows of (lse - s[c]) * inv_count."""
        loss, grad = self.emit(teacher, student, lse2_t, table, consts, T, eps, inv_norm, upstream, want_grad, status=status)
        with torch.enable_grad():                                    # also called from inside an autograd.Function.forward
            sd = student.detach().double().requires_grad_(True)
            ce = torch.nn.functional.cross_entropy(sd, ce_labels, ignore_index=-100, reduction="sum") * inv_count.double().reshape(())
            (gce,) = torch.autograd.grad(ce, sd)
        ce = ce.detach().float()
        return loss + ce_weight * ce, ce, grad + ce_weight * gce


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t, s, lab = synthetic.logit_batch(n_total, V, seed=3)
        # ragged shards on purpose: rank 0 gets fewer rows than an even split
        cut = n_total // 3
        lo, hi = (0, cut) if rank == 0 else (cut, n_total)
        loss, grad = sharded.loca_kl_token_sharded(t[lo:hi], s[lo:hi], lab[lo:hi], T, ALPHA, phases=OraclePhases())
        full = O.loca_kl(t, s, lab, T, ALPHA)
        ok_loss = abs(loss.item() - full.loss.item()) <= 1e-6 * abs(full.loss.item())
        ok_grad = ((grad - full.grad[lo:hi]).norm() / full.grad[lo:hi].norm()).item() <= 1e-6
        # per-row variants: only the normaliser and the scalar sum cross ranks
        part = sharded.per_row_token_sharded(
            lambda a, b, n_total=None: O.temperature_kl(a, b, T, inv_norm=1.0 / (n_total * V)).loss.float(), t[lo:hi], s[lo:hi])
        ok_kl = abs(part.item() - O.temperature_kl(t, s, T).loss.item()) <= 1e-6 * abs(part.item())
        out[rank] = (ok_loss, ok_grad, ok_kl)
    finally:
        dist.destroy_process_group()


def _worker_resizing(rank, world, port, out):
    """Shard sizes change from step to step, on one rank only at first: nothing may be cached across calls (a stale count
    made one rank skip the count collective the other issued — gloo 'collective mismatch', a hang on NCCL)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = []
        # (rows of rank 0, rows of rank 1): rank 1 keeps 2 rows while rank 0 goes 2 -> 3, then both change, then explicit counts
        for step, (n0, n1, explicit) in enumerate([(2, 2, False), (3, 2, False), (3, 5, False), (4, 1, True), (2, 2, True)]):
            n_total = n0 + n1
            t, s, lab = synthetic.logit_batch(n_total, V, seed=50 + step)
            lo, hi = (0, n0) if rank == 0 else (n0, n_total)
            counts = [n0, n1] if explicit else None
            loss, grad = sharded.loca_kl_token_sharded(t[lo:hi], s[lo:hi], lab[lo:hi], T, ALPHA, phases=OraclePhases(), counts=counts)
            full = O.loca_kl(t, s, lab, T, ALPHA)
            ok.append(abs(loss.item() - full.loss.item()) <= 1e-6 * abs(full.loss.item())
                      and ((grad - full.grad[lo:hi]).norm() / full.grad[lo:hi].norm()).item() <= 1e-6)
            part = sharded.per_row_token_sharded(
                lambda a, b, n_total=None: O.temperature_kl(a, b, T, inv_norm=1.0 / (n_total * V)).loss.float(), t[lo:hi], s[lo:hi],
                n_total=n_total if explicit else None)
            ok.append(abs(part.item() - O.temperature_kl(t, s, T).loss.item()) <= 1e-6 * abs(part.item()))
        # a label outside [0, V) on ONE rank poisons the loss on BOTH (the reference raises, P1:166)
        t, s, lab = synthetic.logit_batch(6, V, seed=70)
        lab[4] = -100
        lo, hi = (0, 3) if rank == 0 else (3, 6)
        loss, _ = sharded.loca_kl_token_sharded(t[lo:hi], s[lo:hi], lab[lo:hi], T, ALPHA, phases=OraclePhases())
        ok.append(bool(torch.isnan(loss).item()))
        out[rank] = all(ok)
    finally:
        dist.destroy_process_group()


def _worker_autograd(rank, world, port, out):
    """patch_reference(..., group=...) -> compute_loca_loss on token shards: global loss value, local gradient through autograd,
    grad_output != 1 handled in backward (P1:260 multiplies the loss by gamma)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import patch_reference
        n_total, cut = 11, 4
        t, s, lab = synthetic.logit_batch(n_total, V, seed=9)
        lo, hi = (0, cut) if rank == 0 else (cut, n_total)
        mod = types.SimpleNamespace(T=T, compute_loca_loss=None)
        patch_reference(mod, group=dist.group.WORLD)
        mod.kd_phases = OraclePhases()
        sl = s[lo:hi].double().requires_grad_(True)
        ce = torch.tensor(0.25, dtype=torch.float64, requires_grad=True)
        total = mod.compute_loca_loss(t[lo:hi].double(), sl, ce, lab[lo:hi], ALPHA)
        (0.8 * total).backward()
        full = O.loca_kl(t, s, lab, T, ALPHA)
        ok_loss = abs(total.item() - (full.loss.item() + 0.25)) <= 1e-6 * abs(total.item())
        ok_grad = ((sl.grad - 0.8 * full.grad[lo:hi]).norm() / (0.8 * full.grad[lo:hi]).norm()).item() <= 1e-6
        out[rank] = (ok_loss, ok_grad, abs(ce.grad.item() - 0.8) <= 1e-12)
    finally:
        dist.destroy_process_group()


def _worker_fused_ce(rank, world, port, out):
    """Token shards + the fused cross-entropy: KD and CE are those of the CONCATENATED batch (the CE's divisor is the valid-target
    count of all ranks, summed on the device), the gradient is the local rows'; then the same through the patched method."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import patch_reference
        n_total, cut, w = 13, 5, 1.25
        t, s, lab = synthetic.logit_batch(n_total, V, seed=12)
        g = torch.Generator().manual_seed(4)
        cel = torch.randint(0, V, (n_total,), generator=g)
        cel[[0, 3, 7, 12]] = -100                                    # uneven numbers of ignored targets on the two ranks
        lo, hi = (0, cut) if rank == 0 else (cut, n_total)
        full = O.loca_kl(t, s, lab, T, ALPHA)
        ce_full = O.student_cross_entropy(s, cel)
        ref_loss = full.loss + w * ce_full.loss
        ref_grad = (full.grad + w * ce_full.grad)[lo:hi]
        loss, ce, grad = sharded.loca_kl_token_sharded(t[lo:hi], s[lo:hi], lab[lo:hi], T, ALPHA, phases=OraclePhases(),
                                                       ce_labels=cel[lo:hi], ce_weight=w)
        ok = [abs(loss.item() - ref_loss.item()) <= 1e-6 * abs(ref_loss.item()),
              abs(ce.item() - ce_full.loss.item()) <= 1e-6 * abs(ce_full.loss.item()),
              ((grad - ref_grad).norm() / ref_grad.norm()).item() <= 1e-6]
        # the patched reference method: student_loss=None asks for the fused form; labels [B, S], CE targets = their next-token shift
        B, S = 2, 4
        t, s, lab = synthetic.logit_batch(world * B * S, V, seed=13)
        lab3 = lab.view(world * B, S)
        cel3 = torch.cat([lab3[:, 1:], torch.full((world * B, 1), -100)], dim=1)
        full = O.loca_kl(t, s, lab, T, ALPHA)
        ce_full = O.student_cross_entropy(s, cel3.reshape(-1))
        lo, hi = rank * B * S, (rank + 1) * B * S
        mod = types.SimpleNamespace(T=T, compute_loca_loss=None)
        patch_reference(mod, group=dist.group.WORLD)
        mod.kd_phases = OraclePhases()
        sl = s[lo:hi].double().view(B, S, V).requires_grad_(True)
        total = mod.compute_loca_loss(t[lo:hi].double().view(B, S, V), sl, None, lab3[rank * B:(rank + 1) * B], ALPHA)
        (0.5 * total).backward()
        ref = full.loss + ce_full.loss
        ref_grad = 0.5 * (full.grad + ce_full.grad)[lo:hi]
        ok += [abs(total.item() - ref.item()) <= 1e-6 * abs(ref.item()),
               abs(mod.last_student_loss.item() - ce_full.loss.item()) <= 1e-6 * abs(ce_full.loss.item()),
               ((sl.grad.reshape(-1, V) - ref_grad).norm() / ref_grad.norm()).item() <= 1e-6]
        out[rank] = tuple(ok)
    finally:
        dist.destroy_process_group()


def _worker_module_fused(rank, world, port, out):
    """The Lightning-shaped module with fuse_ce=True and a group: every rank steps on ITS batch (one with padded labels), the loss
    is KD + CE of the two batches concatenated — the CE's divisor counts the valid targets of both ranks — times the phase-3 mix."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import train_online_kd
        from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.module import (
            OnlineKnowledgeDistillationModule, TinyCausalLM)
        from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import functional as F_kd
        batches = list(train_online_kd.synthetic_batches(2, 2, 16, V, 5, 64, torch.device("cpu"), pad_last=False))
        gamma, ok = 0.8, []
        for phase in (2, 3):
            student, teacher = TinyCausalLM(V, 32, seed=1, dtype=torch.float32), TinyCausalLM(V, 48, seed=2, dtype=torch.float32)
            m = OnlineKnowledgeDistillationModule(student, teacher, phase=phase, T=T, gamma=gamma, fuse_ce=True, group=dist.group.WORLD)
            m.kd_phases = OraclePhases()
            loss = m.training_step(batches[rank], 0)
            loss.backward()
            # the concatenated batch, single process, from the same models
            with torch.no_grad():
                sl = torch.cat([student(b["depth_input_ids"], b["depth_pixel_values"]).logits.reshape(-1, V) for b in batches])
                tl = torch.cat([teacher(b["rgb_input_ids"], b["rgb_pixel_values"]).logits.reshape(-1, V) for b in batches])
            lab = torch.cat([b["labels"].reshape(-1) for b in batches])
            cel = torch.cat([F_kd.shift_labels(b["labels"]).reshape(-1) for b in batches])
            kl, ce = O.loca_kl(tl, sl, lab, T, ALPHA), O.student_cross_entropy(sl, cel)
            want = kl.loss + ce.loss if phase == 2 else gamma * (kl.loss + ce.loss) + (1 - gamma) * ce.loss
            ok.append(abs(loss.item() - want.item()) <= 1e-5 * abs(want.item()))
            ok.append(abs(m.last_student_loss.item() - ce.loss.item()) <= 1e-5 * abs(ce.loss.item()))
            ok.append(all(p.grad is not None and torch.isfinite(p.grad).all().item() for p in student.lm_head.parameters()))
        out[rank] = tuple(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_module_with_fused_cross_entropy_on_token_shards_world_size_2():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_module_fused, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True,) * 6, 1: (True,) * 6}


@pytest.mark.timeout(180)
def test_token_sharded_fused_cross_entropy_world_size_2():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_fused_ce, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True,) * 6, 1: (True,) * 6}


@pytest.mark.timeout(180)
def test_patched_compute_loca_loss_on_token_shards_is_differentiable():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_autograd, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True, True), 1: (True, True, True)}


@pytest.mark.timeout(180)
def test_token_sharded_shards_that_resize_between_steps():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker_resizing, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


@pytest.mark.timeout(180)
def test_token_sharded_loca_world_size_2():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 23, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True, True), 1: (True, True, True)}


def test_single_process_passthrough():
    rec = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    assert sharded.gather_records(rec) is rec
