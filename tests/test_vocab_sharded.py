"""Vocabulary-sharded logit KD (BASELINE configs[4]).

* CPU: world-size-2 gloo run of the host logic (statistics all-gather, Q_r all-reduce, loss sum) with plain-torch fp64
  phases, and a three-shard in-process run of every variant — both against the full-vocabulary oracle.
* GPU: the four CUDA phases (through the C ABI) on two simulated shards of one device against the oracle."""
import os
import socket
import threading

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import grad_close
from _vshard_ref import ThreadCollectives, TorchVocabPhases
from oracle import kd_oracle as O
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded, synthetic

T, ALPHA = 2.0, 0.8


def _oracle(mode, t, s, lab, T):
    if mode in ("reference", "per_row"):
        return O.loca_kl(t, s, lab, T, ALPHA, mode=mode)
    return O.temperature_kl(t, s, T, log_target_quirk=(mode == "kl_log_target"))


def _run_shards(mode, t, s, lab, cuts, T, phases_factory, device="cpu", dtype=None):
    """Every shard in its own thread; returns [(loss, grad_shard)] in shard order."""
    world = len(cuts) - 1
    coll = ThreadCollectives(world)
    out = [None] * world
    err = []

    def work(rank):
        try:
            lo, hi = cuts[rank], cuts[rank + 1]
            ts, ss = t[:, lo:hi].contiguous(), s[:, lo:hi].contiguous()
            if dtype is not None:
                ts, ss = ts.to(device, dtype), ss.to(device, dtype)
            out[rank] = sharded.logit_kd_vocab_sharded(ts, ss, lab.to(device) if lab is not None else None, lo, t.shape[1], T, mode=mode,
                                                       alpha=ALPHA, phases=phases_factory(), collectives=coll.bind(rank))
        except Exception as exc:  # noqa: BLE001
            err.append(exc)
            coll.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [th.start() for th in threads]
    [th.join() for th in threads]
    if err:
        raise err[0]
    return out


@pytest.mark.parametrize("mode", ["reference", "per_row", "kl", "kl_log_target"])
def test_three_shards_in_process_cpu(mode):
    n, V = 19, 301
    t, s, lab = synthetic.logit_batch(n, V, seed=11)
    t, s = t.float(), s.float()
    res = _run_shards(mode, t, s, lab, [0, 97, 97 + 128, V], T, TorchVocabPhases)
    ref = _oracle(mode, t, s, lab, T)
    grad = torch.cat([g for _, g in res], dim=1)
    for loss, _ in res:
        assert abs(loss.item() - ref.loss.item()) <= 2e-6 * abs(ref.loss.item())
    assert ((grad - ref.grad).norm() / ref.grad.norm()).item() <= 1e-6


def _free_port():
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        return sck.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, V = 13, 257
        t, s, lab = synthetic.logit_batch(n, V, seed=4)
        t, s = t.float(), s.float()
        cut = 129
        lo, hi = (0, cut) if rank == 0 else (cut, V)
        oks = []
        for mode in ("reference", "per_row", "kl"):
            loss, grad = sharded.logit_kd_vocab_sharded(t[:, lo:hi].contiguous(), s[:, lo:hi].contiguous(), lab, lo, V, T, mode=mode,
                                                        alpha=ALPHA, phases=TorchVocabPhases())
            ref = _oracle(mode, t, s, lab, T)
            oks.append(abs(loss.item() - ref.loss.item()) <= 2e-6 * abs(ref.loss.item()))
            oks.append(((grad - ref.grad[:, lo:hi]).norm() / ref.grad[:, lo:hi].norm()).item() <= 1e-6)
        out[rank] = tuple(oks)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_vocab_sharded_world_size_2_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True,) * 6, 1: (True,) * 6}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["reference", "per_row", "kl", "kl_log_target"])
@pytest.mark.parametrize("n,V,cuts,Tk", [(24, 8192, [0, 4096, 8192], 2.0), (9, 5000, [0, 1237, 5000], 0.8),
                                         # >= 64 rows per SM: the warp-per-row kernels (vector and ragged scalar walks)
                                         (9600, 1040, [0, 520, 1040], 2.0), (9500, 777, [0, 301, 777], 0.8)])
def test_cuda_phases_two_simulated_shards(mode, n, V, cuts, Tk):
    t, s, lab = synthetic.logit_batch(n, V, seed=21, peaked=(Tk < 1.0))
    res = _run_shards(mode, t, s, lab, cuts, Tk, sharded.CudaVocabPhases, device="cuda", dtype=torch.bfloat16)
    ref = _oracle(mode, t, s, lab, Tk)
    grad = torch.cat([g.float().cpu() for _, g in res], dim=1)
    for loss, _ in res:
        assert abs(loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item())
    assert grad_close(grad, ref.grad)


def test_vocab_parallel_lm_head_gradients_cpu():
    """hidden @ weight_shard^T -> vocabulary-sharded KD: loss, d/d hidden (summed over shards) and d/d weight_shard against
    autograd through the full-vocabulary oracle gradient."""
    n, K, V = 11, 24, 264
    g = torch.Generator().manual_seed(2)
    H = torch.randn(n, K, generator=g, dtype=torch.float64)
    W = torch.randn(V, K, generator=g, dtype=torch.float64) * 0.5
    t, _, lab = synthetic.logit_batch(n, V, seed=5)
    t = t.double()
    logits = H @ W.t()
    ref = O.loca_kl(t, logits, lab, T, ALPHA, mode="reference")
    dH_ref, dW_ref = ref.grad @ W, ref.grad.t() @ H
    cuts = [0, 96, 200, V]
    world = len(cuts) - 1
    coll = ThreadCollectives(world)
    out, err = [None] * world, []

    def work(rank):
        try:
            lo, hi = cuts[rank], cuts[rank + 1]
            h = H.clone().requires_grad_(True)
            w = W[lo:hi].clone().requires_grad_(True)
            loss = sharded.lm_head_kd_vocab_sharded(h, w, t[:, lo:hi].contiguous(), lab, lo, V, T, alpha=ALPHA,
                                                    phases=TorchVocabPhases(), collectives=coll.bind(rank))
            loss.backward()
            out[rank] = (loss.item(), h.grad, w.grad)
        except Exception as exc:  # noqa: BLE001
            err.append(exc)
            coll.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [th.start() for th in threads]
    [th.join() for th in threads]
    assert not err, err
    for rank, (loss, dH, dW) in enumerate(out):
        lo, hi = cuts[rank], cuts[rank + 1]
        assert abs(loss - ref.loss.item()) <= 2e-6 * abs(ref.loss.item())
        assert ((dH - dH_ref).norm() / dH_ref.norm()).item() <= 1e-6
        assert ((dW - dW_ref[lo:hi]).norm() / dW_ref[lo:hi].norm()).item() <= 1e-6
