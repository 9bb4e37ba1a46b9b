#!/usr/bin/env python
"""Benchmark of the KD-loss hot path (BASELINE.json metric: KD loss fwd+bwd tokens/s @ V=151,936).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode reference|per_row]

One "step" = one LoCa logit-KD forward+backward (loss + bf16 gradient w.r.t. the student logits) over a
batch of B*S = 16 384 synthetic token rows per GPU (BASELINE.json configs[1]: B=8, S=2048, V=151 936, T=2.0).
N > 1 is launched by torchrun (one rank per GPU); tokens are sharded across ranks (weak scaling: 16 384 rows
per rank) and, in the as-written "reference" mode, the 16-byte token records are all-gathered over NCCL
between the two phases.  Rank 0 prints ONE JSON line.

``--impl reference`` times the reference's own CPU path (the eager-PyTorch port in oracle/reference_port.py —
the reference is a Python tree that cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

V = 151_936
ROWS_PER_GPU = 16_384          # B=8 x S=2048
T_KD, ALPHA = 2.0, 0.8
CPU_SAMPLE_ROWS = 512          # BASELINE.json configs[0]: B=1, S=512 on the host CPU
METRIC = "KD loss fwd+bwd tokens/s @ V=151,936"


def load_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU during the timed region.

    B200_PROFILING.md's recipe: an `nvidia-smi --query-gpu=... -lms 100` process started BEFORE the timed region and stopped after
    it; only the samples whose timestamps fall inside the region are reported.  (An in-process NVML poller thread stalled kernel
    launches by ~10 ms per sample in multi-rank runs: 2 GPUs, 10 steps: 5.25 ms/step with it, 3.98 ms/step without.)  Only the
    reporting rank samples."""

    FIELDS = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        import shutil
        import subprocess
        import tempfile
        self.proc, self.path, self.t0, self.t1 = None, None, None, None
        if os.environ.get("KD_BENCH_NO_CLOCKS") or not enabled or not shutil.which("nvidia-smi"):
            return
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = visible.split(",")[index] if visible else str(index)
            fd, self.path = tempfile.mkstemp(prefix="kd_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", phys], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def __enter__(self):
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()

    def summary(self) -> dict:
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)                      # let one more sample land, then stop the poller
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, int(float(f[1])), int(float(f[2])), [n for n, v in zip(self.NAMES, f[3:7]) if v == "Active"]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except OSError:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        note = None
        if not inside and rows:               # region shorter than the polling period: the samples bracketing it
            inside = sorted(rows, key=lambda r: min(abs(r[0] - self.t0), abs(r[0] - (self.t1 or self.t0))))[:2]
            note = "timed region shorter than the 100 ms polling period: nearest samples"
        if inside:
            mhz = sorted(r[1] for r in inside)
            out = {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": inside[0][2], "reasons": sorted({n for r in inside for n in r[3]}),
                   "samples": len(inside)}
            if note:
                out["note"] = note
        return out


# --------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline (oracle port; the only place bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------------
def cpu_reference_step_time(repeats: int, warmup: int) -> tuple[float, int]:
    from oracle import reference_port as P
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
    import warnings
    warnings.filterwarnings("ignore", message=".*reduction: 'mean' divides.*")
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t, s, lab = synthetic.logit_batch(CPU_SAMPLE_ROWS, V, seed=1234, dtype=torch.float32)
    t, s, lab = t.reshape(1, CPU_SAMPLE_ROWS, V), s.reshape(1, CPU_SAMPLE_ROWS, V), lab.reshape(1, CPU_SAMPLE_ROWS)
    best = float("inf")
    for i in range(warmup + repeats):
        sv = s.clone().requires_grad_(True)
        t0 = time.perf_counter()
        loss = P.loca_loss_eager(t, sv, lab, T_KD, ALPHA)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            best = min(best, dt)
    return best, threads


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    times = []
    from oracle import reference_port as P
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
    import warnings
    warnings.filterwarnings("ignore", message=".*reduction: 'mean' divides.*")
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    t, s, lab = synthetic.logit_batch(CPU_SAMPLE_ROWS, V, seed=1234, dtype=torch.float32)
    t, s, lab = t.reshape(1, CPU_SAMPLE_ROWS, V), s.reshape(1, CPU_SAMPLE_ROWS, V), lab.reshape(1, CPU_SAMPLE_ROWS)
    for i in range(args.warmup + args.steps):
        sv = s.clone().requires_grad_(True)
        a = time.perf_counter()
        P.loca_loss_eager(t, sv, lab, T_KD, ALPHA).backward()
        if i >= args.warmup:
            times.append(time.perf_counter() - a)
    ms = 1e3 * sum(times) / len(times)
    value = CPU_SAMPLE_ROWS / (ms / 1e3)
    sample = (f"each step = B=1,S={CPU_SAMPLE_ROWS} fp32 rows of the workload (BASELINE configs[0]); the eager path needs "
              f"~10 [N,V] fp32 temporaries so 16 384 rows do not fit; {threads} threads on {cpu_model_name()}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LoCa logit-KD fwd+bwd, B=8,S=2048,V=151936,T=2.0 (CPU sample B=1,S=512)", "mode": "reference"},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }))


# --------------------------------------------------------------------------------------------------------
# parity of the timed configuration itself (outside the timed region; oracle = checker only)
# --------------------------------------------------------------------------------------------------------
def parity_block(kd, sharded, dist, teacher, student, labels, mode, world, rank, dev, loss_value, step, exchange, phases):
    """Loss and a strided gradient sample of THE BENCHMARKED BATCH against oracle/kd_oracle.py run chunked in fp64 on the GPU.
    N > 1: every rank runs the oracle on its rows with the GLOBAL table (its fp64 token records are all-gathered), the partial
    losses are summed, and — on 256 rows per rank — the sharded loss is compared with the one-GPU op on the concatenated batch."""
    from oracle import kd_oracle as O
    N = teacher.shape[0]
    n_total = N * world
    stride = 97
    step()                                              # leaves the gradient of the benchmarked batch in student.grad
    grad = student.grad
    override = None
    if world > 1 and mode == "reference":
        st = O.loca_row_stats(teacher, labels, T_KD, ALPHA, V, chunk_rows=128)
        override = {}
        for key in ("labels", "k", "v1", "v2"):
            part = st[key].contiguous()
            full = torch.empty((n_total,), dtype=part.dtype, device=dev)
            dist.all_gather_into_tensor(full, part)
            override[key] = full
    ref = O.loca_kl(teacher, student.detach(), labels, T_KD, ALPHA, mode=mode, inv_norm=1.0 / (float(n_total) * V), chunk_rows=128,
                    grad_stride=stride, stats_override=override)
    ref_loss = ref.loss.reshape(1).clone()
    num = (grad[:, ::stride].double() - ref.grad).pow(2).sum().reshape(1)
    den = ref.grad.pow(2).sum().reshape(1)
    if world > 1:
        for x in (ref_loss, num, den):
            dist.all_reduce(x)
    out = {"oracle": "oracle/kd_oracle.py, fp64 on the GPU, chunked; every 97th gradient column", "rows_checked": n_total,
           "loss": loss_value, "oracle_loss": ref_loss.item(), "loss_rel_err": abs(loss_value - ref_loss.item()) / abs(ref_loss.item()),
           "grad_rel_err": (num.sqrt() / den.sqrt()).item(), "gates": {"loss": 1e-5, "grad": 2e-3}}
    del ref
    if world > 1:
        m = 256
        tt, ss, ll = teacher[:m].contiguous(), student.detach()[:m].contiguous(), labels[:m].contiguous()
        if mode == "reference":
            sl, _ = sharded.loca_kl_token_sharded(tt, ss, ll, T_KD, ALPHA, phases=phases, exchange="nccl", counts=[m] * world)
        else:
            sl = sharded.per_row_token_sharded(kd.loca_kl, tt, ss, ll, T_KD, alpha=ALPHA, mode="per_row", n_total=m * world)
        gt = torch.empty((world * m, tt.shape[1]), dtype=tt.dtype, device=dev)
        gs = torch.empty((world * m, ss.shape[1]), dtype=ss.dtype, device=dev)
        gl = torch.empty((world * m,), dtype=ll.dtype, device=dev)
        dist.all_gather_into_tensor(gt, tt)
        dist.all_gather_into_tensor(gs, ss)
        dist.all_gather_into_tensor(gl, ll)
        one = kd.loca_kl(gt, gs, gl, T_KD, alpha=ALPHA, mode=mode)
        out["sharded_vs_one_gpu"] = {"rows": world * m, "sharded_loss": sl.item(), "one_gpu_loss_of_concatenated_batch": one.item(),
                                     "rel_diff": abs(sl.item() - one.item()) / abs(one.item())}
        del gt, gs, gl
    out["ok"] = bool(out["loss_rel_err"] <= 1e-5 and out["grad_rel_err"] <= 2e-3 and
                     out.get("sharded_vs_one_gpu", {}).get("rel_diff", 0.0) <= 1e-5)
    return out


# --------------------------------------------------------------------------------------------------------
# BASELINE configs[4] (N >= 2): fused student lm_head (896 x 151 936) + LoCa-KL, vocabulary sharded, 64k tokens
# --------------------------------------------------------------------------------------------------------
def vshard_line(sharded, synthetic, dist, world, rank, dev):
    """Every rank holds all 65 536 hidden rows and V / world columns of the head and of the teacher logits; per step: tcgen05 head
    GEMM (+ fused student statistics) -> stats -> all-gather -> combine -> table -> accum -> all-reduce(Q_r) -> emit -> dH (fp32 partial,
    all-reduced) and dW on the tcgen05 kernel.  Parity at 64 rows against the fp64 oracle first; then the timed 64k-token step."""
    from oracle import kd_oracle as O
    K, n_tok = 896, 65_536
    lo = rank * (V // world)
    hi = (rank + 1) * (V // world) if rank < world - 1 else V
    out = {"tokens": n_tok, "head": f"{K} x {V}", "columns_per_gpu": hi - lo, "mode": "reference (as written)", "T": T_KD}
    # ---- parity, 64 rows, identical operands on every rank (CPU generator) ----
    g = torch.Generator().manual_seed(11)
    n = 64
    Hs = torch.randn(n, K, generator=g).to(torch.bfloat16)
    Wf = (torch.randn(V, K, generator=g) * 0.02).to(torch.bfloat16)
    t, _, lab = synthetic.logit_batch(n, V, seed=12)
    Hd, Wd, td = Hs.to(dev), Wf.to(dev), t.to(dev)
    logits = torch.matmul(Hd.float(), Wd.float().t()).to(torch.bfloat16)           # what a bf16 GEMM hands to the loss
    ref = O.loca_kl(td, logits, lab.to(dev), T_KD, ALPHA, chunk_rows=16)
    dH_ref = ref.grad @ Wd.double()
    dW_ref = ref.grad.t() @ Hd.double()
    h = Hd.clone().requires_grad_(True)
    w = Wd[lo:hi].clone().requires_grad_(True)
    loss = sharded.lm_head_kd_vocab_sharded(h, w, td[:, lo:hi].contiguous(), lab.to(dev), lo, V, T_KD)
    loss.backward()
    errs = torch.tensor([abs(loss.item() - ref.loss.item()) / abs(ref.loss.item()),
                         ((h.grad.double() - dH_ref).norm() / dH_ref.norm()).item(),
                         ((w.grad.double() - dW_ref[lo:hi]).norm() / dW_ref[lo:hi].norm()).item()], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    out["parity_64_rows"] = {"loss_rel_err": errs[0].item(), "dH_rel_err": errs[1].item(), "dW_rel_err": errs[2].item(),
                             "gates": {"loss": 1e-5, "dH": 2e-3, "dW": 2e-3},
                             "ok": bool(errs[0].item() <= 1e-5 and errs[1].item() <= 2e-3 and errs[2].item() <= 2e-3)}
    del Hd, Wd, td, logits, ref, dH_ref, dW_ref, h, w, loss
    torch.cuda.empty_cache()
    # ---- timed step ----
    gd = torch.Generator(device=dev).manual_seed(5)
    H = torch.randn(n_tok, K, device=dev, generator=gd).to(torch.bfloat16).requires_grad_(True)           # same seed: replicated
    Ws = (torch.randn(hi - lo, K, device=dev) * 0.02).to(torch.bfloat16).requires_grad_(True)
    ts = torch.empty(n_tok, hi - lo, device=dev, dtype=torch.bfloat16)
    for r0 in range(0, n_tok, 8192):
        ts[r0:r0 + 8192] = (torch.randn(8192, hi - lo, device=dev) * 2.0).to(torch.bfloat16)
    labs = torch.randint(0, V, (n_tok,), device=dev, generator=gd)

    def step():
        H.grad = None
        Ws.grad = None
        sharded.lm_head_kd_vocab_sharded(H, Ws, ts, labs, lo, V, T_KD).backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 5
    a.record()
    for _ in range(k):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / k], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tf_peak = 1401.6
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf_peak = float(json.load(f)["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        pass
    flops = 3 * 2.0 * n_tok * K * V                      # forward product + dH + dW, whole job
    out.update({"ms_per_step": ms.item(), "tokens_per_s": n_tok / (ms.item() / 1e3), "tflops_whole_job": flops / (ms.item() / 1e3) / 1e12,
                "bound": "tensor", "peak_tflops_per_gpu": tf_peak, "frac": flops / (ms.item() / 1e3) / 1e12 / (world * tf_peak),
                "scaling": "strong (65 536 tokens, head sharded over the GPUs)",
                "note": "frac = achieved whole-job TFLOP/s / (n_gpus x measured sustained bf16 peak); the streaming KD phases "
                        "(3 teacher reads, 2 student reads, 1 gradient write per shard) are HBM-bound and sit inside the same step"})
    del H, Ws, ts, labs
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="reference", choices=["reference", "per_row"])
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="token rows per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the per_row / KL / InfoNCE / MSE side measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the benchmarked batch")
    ap.add_argument("--force-streaming", action="store_true", help="time the general streaming emit kernel instead of the cluster kernel")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1, as-written mode: record all-gather by peer stores from the phase-A kernel (NVLink symmetric memory) or by NCCL")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
    from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib, sharded, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(args.gpus, 1) and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    peak_gbs, peak_src = load_peaks()
    N = args.rows
    mode = args.mode

    teacher, student, labels = synthetic.logit_batch(N, V, seed=1234 + rank, device=dev)
    student.requires_grad_(True)
    n_total = N * world
    counts = [N] * world                  # fixed, equal shards: passed explicitly so no step pays a count exchange
    phases = sharded.CudaPhases()
    exchange = args.exchange if world > 1 else "none"
    if world > 1 and mode == "reference" and exchange == "peer":
        try:        # symmetric memory needs P2P access between the ranks' GPUs; fall back to the NCCL all-gather otherwise
            sharded.loca_kl_token_sharded(teacher[:64], student.detach()[:64], labels[:64], T_KD, ALPHA, phases=phases, exchange="peer")
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(f"[bench] peer exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            exchange = "nccl"
        flag = torch.tensor([1 if exchange == "peer" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        exchange = "peer" if int(flag.item()) == 1 else "nccl"

    def step():
        """The public call a user makes: forward + backward through the autograd Function(s)."""
        student.grad = None
        if world == 1:
            loss = kd.loca_kl(teacher, student, labels, T_KD, alpha=ALPHA, mode=mode)
            loss.backward()
        elif mode == "reference":
            loss, grad = sharded.loca_kl_token_sharded(teacher, student.detach(), labels, T_KD, ALPHA, phases=phases, exchange=exchange,
                                                       counts=counts)
            student.grad = grad
        else:
            loss = sharded.per_row_token_sharded(kd.loca_kl, teacher, student, labels, T_KD, alpha=ALPHA, mode="per_row", n_total=n_total)
            loss.backward()
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank, enabled=(rank == 0))      # the poller starts before the warm-up, well before the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clocks:
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        sync_all()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = ms_total.item() / args.steps
    value = n_total / (ms_step / 1e3)
    loss_value = float(loss.item())
    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(kd, sharded, dist, teacher, student, labels, mode, world, rank, dev, loss_value, step, exchange, phases)
        except Exception as exc:  # noqa: BLE001 - report, do not hide the headline
            parity = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    # ---- dominant kernel (phase C / emit) timed alone with CUDA events on its launch stream ------------
    roof = None
    side = {}
    if rank == 0:
        esize = 2
        if mode == "reference":
            rec, lse2_t, _ = phases.teacher_stats(teacher, labels, T_KD, ALPHA, V)
            table, consts = phases.build_table(rec, V)
            inv_norm = 1.0 / (float(N) * V)

            def emit():
                return phases.emit(teacher, student.detach(), lse2_t, table, consts, T_KD, 1e-8, inv_norm, None,
                                   force_streaming=args.force_streaming)

            def stats():
                return phases.teacher_stats(teacher, labels, T_KD, ALPHA, V)
            for _ in range(3):
                emit(); stats()
            torch.cuda.synchronize()
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            k = max(args.steps // 2, 5)
            a.record()
            for _ in range(k):
                emit()
            b.record()
            for _ in range(k):
                stats()
            c.record()
            torch.cuda.synchronize()
            emit_ms, stats_ms = a.elapsed_time(b) / k, b.elapsed_time(c) / k
            emit_bytes = 3 * esize * V * N            # read teacher + read student + write grad
            achieved = emit_bytes / (emit_ms / 1e3) / 1e9
            traffic, traffic_src = None, None
            if not args.force_streaming and N == ROWS_PER_GPU:
                try:      # dram__bytes_read+write per row from the committed ncu --set full capture of THIS kernel and shape
                    with open(os.path.join(ROOT, "profiles", "emit_traffic.json")) as f:
                        tj = json.load(f)
                    traffic = tj["dram_bytes_per_row"] * N
                    traffic_src = (f"NOT measured in this run: static ncu --set full capture ({tj.get('rows', '?')} rows, "
                                   f"{tj.get('source', 'profiles/')}, commit {tj.get('commit', '?')}) x {N} rows")
                except Exception:  # noqa: BLE001
                    pass
            roof = {"bound": "hbm", "kernel": "LoCa emit (phase C: q/Q_r/loss/grad)", "achieved": achieved, "peak": peak_gbs,
                    "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes": emit_bytes,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_token": 3 * esize * V, "kernel_ms": emit_ms}
            side["teacher_stats"] = {"kernel_ms": stats_ms, "achieved_gbs": esize * V * N / (stats_ms / 1e3) / 1e9,
                                     "frac": esize * V * N / (stats_ms / 1e3) / 1e9 / peak_gbs,
                                     "algorithmic_bytes_per_token": esize * V}
            step_bytes = 4 * esize * V * N
        else:
            step_bytes = 3 * esize * V * N
            achieved = step_bytes / (ms_step / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "LoCa per-row fused emit", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                    "frac": achieved / peak_gbs, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_token": 3 * esize * V, "kernel_ms": ms_step}
        side["whole_step"] = {"algorithmic_bytes_per_token": step_bytes // N,
                              "achieved_gbs": step_bytes / (ms_step / 1e3) / 1e9,
                              "frac": step_bytes / (ms_step / 1e3) / 1e9 / peak_gbs}

    # ---- secondary paths (one GPU, rank 0): per_row LoCa, plain KL, InfoNCE, MSE -----------------------
    if rank == 0 and not args.no_secondary:
        def timeit(fn, k=10, w=3):
            import gc
            for _ in range(w):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            gc.collect()
            gc.disable()                      # (the sub-millisecond ops are host-bound: keep collector pauses out of their loop)
            try:
                a.record()
                for _ in range(k):
                    fn()
                b.record()
                torch.cuda.synchronize()
            finally:
                gc.enable()
            return a.elapsed_time(b) / k

        def timeit_graph(fn, k=50):
            """The same step replayed from a CUDA graph: the device time of its launch chain without the host between launches."""
            side_stream = torch.cuda.Stream()
            side_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side_stream):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / k

        def fb(f):
            def run():
                student.grad = None
                f().backward()
            return run
        other = "per_row" if mode == "reference" else "reference"
        for name, f, byt in (
            (f"loca_{other}", fb(lambda: kd.loca_kl(teacher, student, labels, T_KD, alpha=ALPHA, mode=other)), (6 if other == "per_row" else 8) * V * N),
            ("kl_plain_T0.8", fb(lambda: kd.temperature_kl(teacher, student, 0.8)), 6 * V * N),
        ):
            ms = timeit(f, k=5, w=2)
            side[name] = {"ms": ms, "tokens_per_s": N / (ms / 1e3), "achieved_gbs": byt / (ms / 1e3) / 1e9,
                          "frac": byt / (ms / 1e3) / 1e9 / peak_gbs}
        # backward re-emission: what `(w * loss).backward()` (grad_output != the scale folded in at forward time; a changing AMP loss
        # scale, TR2:145) costs on top of the step — one streaming 6V sweep from the saved row statistics
        try:
            def weighted():
                student.grad = None
                (0.5 * kd.loca_kl(teacher, student, labels, T_KD, alpha=ALPHA, mode=mode)).backward()
            ms_w = timeit(weighted, k=5, w=2)
            ms_p = timeit(fb(lambda: kd.loca_kl(teacher, student, labels, T_KD, alpha=ALPHA, mode=mode)), k=5, w=2)
            re_ms = max(ms_w - ms_p, 1e-6)
            side["backward_re_emission"] = {"step_ms": ms_p, "step_with_re_emission_ms": ms_w, "re_emission_ms": re_ms,
                                            "achieved_gbs": 6 * V * N / (re_ms / 1e3) / 1e9, "frac": 6 * V * N / (re_ms / 1e3) / 1e9 / peak_gbs,
                                            "algorithmic_bytes_per_token": 6 * V,
                                            "note": "difference of two step times (includes the two scalar multiplies of autograd)"}
        except Exception as exc:  # noqa: BLE001
            side["backward_re_emission"] = {"error": f"{type(exc).__name__}: {exc}"}
        # the reference's own temperatures (P1:71 T = 0.8, LG:75 T = 1) on peaked logits (label = teacher's best class on ~70 % of
        # the rows, student 4x sharper): every row takes the sp >= eps clamp path (pass B' + second reduction), the slow case
        tpk, spk, lpk = synthetic.logit_batch(N, V, seed=4321, device=dev, peaked=True)
        spk = (spk.float() * 4.0).to(torch.bfloat16).requires_grad_(True)

        def fbp(f):
            def run():
                spk.grad = None
                f().backward()
            return run
        for name, f, byt in (
            ("loca_reference_T0.8_peaked", fbp(lambda: kd.loca_kl(tpk, spk, lpk, 0.8, alpha=ALPHA)), 8 * V * N),
            ("loca_reference_T1_peaked", fbp(lambda: kd.loca_kl(tpk, spk, lpk, 1.0, alpha=ALPHA)), 8 * V * N),
            ("loca_per_row_T0.8_peaked", fbp(lambda: kd.loca_kl(tpk, spk, lpk, 0.8, alpha=ALPHA, mode="per_row")), 6 * V * N),
            ("kl_log_target_T0.8", fb(lambda: kd.temperature_kl(teacher, student, 0.8, log_target_quirk=True)), 6 * V * N),
        ):
            ms = timeit(f, k=5, w=2)
            side[name] = {"ms": ms, "tokens_per_s": N / (ms / 1e3), "achieved_gbs": byt / (ms / 1e3) / 1e9,
                          "frac": byt / (ms / 1e3) / 1e9 / peak_gbs, "algorithmic_bytes_per_token": byt // N}
        # the student's shifted cross-entropy (student_outputs.loss, P1:238-240 -> P1:194) out of the SAME pass
        # as the logit KD.  Half of the rows carry a next-token target (the rest are prompt / padding, -100).  Beside it: what the
        # unfused step costs — the KD op plus the library F.cross_entropy on the same bf16 logits (chunked over rows so that its
        # fp32 temporaries fit; a HF model's own loss upcasts the whole [N, V] tensor at once).
        try:
            gce = torch.Generator(device="cpu").manual_seed(77)
            cel = torch.randint(0, V, (N,), generator=gce)
            cel[torch.rand(N, generator=gce) < 0.5] = -100
            cel = cel.to(dev)

            def lib_ce_step():
                student.grad = None
                kl = kd.loca_kl(teacher, student, labels, T_KD, alpha=ALPHA)
                n_valid = (cel != -100).sum()
                ce = 0.0
                for r0 in range(0, N, 2048):
                    ce = ce + torch.nn.functional.cross_entropy(student[r0:r0 + 2048].float(), cel[r0:r0 + 2048], ignore_index=-100,
                                                                reduction="sum")
                (kl + ce / n_valid).backward()
            ce_lines = {}
            for name, f in (("T2.0", fb(lambda: kd.loca_kl_ce(teacher, student, labels, cel, T_KD, alpha=ALPHA)[0])),
                            ("T1.0", fb(lambda: kd.loca_kl_ce(teacher, student, labels, cel, 1.0, alpha=ALPHA)[0]))):
                ms = timeit(f, k=5, w=2)
                ce_lines[name] = {"ms": ms, "tokens_per_s": N / (ms / 1e3), "achieved_gbs": 8 * V * N / (ms / 1e3) / 1e9,
                                  "frac": 8 * V * N / (ms / 1e3) / 1e9 / peak_gbs}
            ms_lib = timeit(lib_ce_step, k=3, w=1)
            student.grad = None
            torch.cuda.empty_cache()
            side["loca_reference_plus_fused_ce"] = {**ce_lines, "valid_target_fraction": 0.5, "algorithmic_bytes_per_token": 8 * V,
                                                    "unfused_kd_plus_library_cross_entropy_ms": ms_lib,
                                                    "note": "loss = LoCa KL + CE, gradient of both, one pass over the student logits; "
                                                            "T = 1 shares the KD exponentials, T = 2 pays one more ex2 per cell and sweep on valid rows"}
        except Exception as exc:  # noqa: BLE001
            side["loca_reference_plus_fused_ce"] = {"error": f"{type(exc).__name__}: {exc}"}
        del tpk, spk, lpk
        torch.cuda.empty_cache()
        xs, xt = synthetic.vision_feature_batch(64, 729, 1152, device=dev)
        xs.requires_grad_(True)

        def nce():
            xs.grad = None
            kd.info_nce(xs, xt).backward()
        ms = timeit(nce, k=50)
        nce_bytes = 3 * xs.numel() * 2           # read both feature tensors, write the gradient
        side["info_nce_64x729x1152"] = {"us": ms * 1e3, "achieved_gbs": nce_bytes / (ms / 1e3) / 1e9,
                                        "frac": nce_bytes / (ms / 1e3) / 1e9 / peak_gbs,
                                        "note": "eager autograd step: host-bound (five launches + the autograd engine per step)"}
        try:
            msg = timeit_graph(nce)
            side["info_nce_64x729x1152"].update({"graph_replay_us": msg * 1e3, "graph_replay_gbs": nce_bytes / (msg / 1e3) / 1e9,
                                                 "graph_replay_frac": nce_bytes / (msg / 1e3) / 1e9 / peak_gbs})
        except Exception as exc:  # noqa: BLE001
            side["info_nce_64x729x1152"]["graph_replay_error"] = f"{type(exc).__name__}: {exc}"
        xs.grad = None
        # BASELINE configs[2] with its learned projection (1152 -> 896 on the pooled student vectors; the repo's tcgen05 GEMM on
        # hi + lo bf16 parts of the fp32 pooled features), teacher features 896 wide
        try:
            wp = (torch.randn(1152, 896, device=dev) * 0.03).to(torch.bfloat16).requires_grad_(True)
            xt896 = torch.randn(64, 729, 896, device=dev).to(torch.bfloat16)

            def nce_proj():
                xs.grad = None
                wp.grad = None
                kd.info_nce(xs, xt896, proj=wp).backward()
            ms = timeit(nce_proj, k=50)
            proj_bytes = (2 * xs.numel() + xt896.numel()) * 2      # read student + teacher features, write the student gradient
            side["info_nce_64x729x1152_proj896"] = {"us": ms * 1e3, "achieved_gbs": proj_bytes / (ms / 1e3) / 1e9,
                                                    "frac": proj_bytes / (ms / 1e3) / 1e9 / peak_gbs,
                                                    "note": "one library call (kd_info_nce_proj_fwd_bwd): balanced pooling of both tensors, "
                                                            "y = mean(x_s) @ W, dW and d mean as fp32 CUDA-core products, row kernel, "
                                                            "store-only broadcast; gradients w.r.t. the features and W"}
            try:
                msg = timeit_graph(nce_proj)
                side["info_nce_64x729x1152_proj896"].update({"graph_replay_us": msg * 1e3,
                                                             "graph_replay_frac": proj_bytes / (msg / 1e3) / 1e9 / peak_gbs})
            except Exception as exc:  # noqa: BLE001
                side["info_nce_64x729x1152_proj896"]["graph_replay_error"] = f"{type(exc).__name__}: {exc}"
            xs.grad = None
            del wp, xt896
        except Exception as exc:  # noqa: BLE001
            side["info_nce_64x729x1152_proj896"] = {"error": f"{type(exc).__name__}: {exc}"}
        del xs, xt
        # token-level contrastive loss: all 64 x 729 = 46 656 un-pooled SigLIP tokens as rows (contrastive_loss takes any [M, D],
        # P1:393-416): the similarity GEMM (P1:410) on the tensor cores with the row log-softmax in its epilogue, 3 x 5.0 TFLOP
        try:
            tf_pk = 1401.6
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    tf_pk = float(json.load(f)["bf16_tflops_sustained"])
            except Exception:  # noqa: BLE001
                pass
            Mt, Dt = 46_656, 1152
            ts_, tt_ = synthetic.vision_feature_batch(Mt, 1, Dt, seed=7, device=dev)
            ts_, tt_ = ts_[:, 0].contiguous().requires_grad_(True), tt_[:, 0].contiguous()

            def nce_tok():
                ts_.grad = None
                kd.info_nce(ts_, tt_).backward()
            ms = timeit(nce_tok, k=5, w=2)
            fl = 3 * 2.0 * Mt * Mt * Dt / 1e12
            side["info_nce_token_level_46656x1152"] = {"ms": ms, "tflops": fl / (ms / 1e3), "bound": "tensor", "peak_tflops": tf_pk,
                                                       "frac": fl / (ms / 1e3) / tf_pk, "rows_per_s": Mt / (ms / 1e3),
                                                       "note": "forward + backward; three M x M x D products on the tcgen05 kernel (statistics, "
                                                               "softmax - I from a recomputation, G . u_t); the similarity matrix is never stored"}
            del ts_, tt_
        except Exception as exc:  # noqa: BLE001
            side["info_nce_token_level_46656x1152"] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        ha = torch.randn(65_536, 3584, device=dev, dtype=torch.bfloat16).requires_grad_(True)
        hb = torch.randn(65_536, 3584, device=dev, dtype=torch.bfloat16)

        def mse():
            ha.grad = None
            kd.feature_mse(ha, hb).backward()
        ms = timeit(mse)
        side["feature_mse_65536x3584"] = {"ms": ms, "achieved_gbs": 6 * ha.numel() / (ms / 1e3) / 1e9,
                                          "frac": 6 * ha.numel() / (ms / 1e3) / 1e9 / peak_gbs}
        try:      # the same step without the host between its launches (a 0.22 ms kernel under ~20 us of autograd)
            msg = timeit_graph(mse)
            side["feature_mse_65536x3584"].update({"graph_replay_ms": msg, "graph_replay_frac": 6 * ha.numel() / (msg / 1e3) / 1e9 / peak_gbs})
        except Exception as exc:  # noqa: BLE001
            side["feature_mse_65536x3584"]["graph_replay_error"] = f"{type(exc).__name__}: {exc}"
        del ha, hb
        torch.cuda.empty_cache()
        tf_peak, tf_src = 1401.6, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                tf_peak, tf_src = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
        except Exception:  # noqa: BLE001
            pass
        # BASELINE configs[3]: feature MSE on PROJECTED hidden states (student 896-d -> teacher 3584-d), B=16, S=4096: the projection
        # on the tcgen05 kernel with the MSE as its epilogue, fwd + bwd (d input, d proj on the same kernel); bound: tensor pipe
        try:
            hs_, ht_, wp_ = synthetic.hidden_state_batch(65_536, 896, 3584, device=dev)
            hs_.requires_grad_(True)
            wp_.requires_grad_(True)

            def pmse():
                hs_.grad = None
                wp_.grad = None
                kd.feature_mse(hs_, ht_, proj=wp_).backward()

            def pmse_fwd():
                with torch.no_grad():
                    kd.feature_mse(hs_, ht_, proj=wp_)

            def pmse_lib():
                hs_.grad = None
                wp_.grad = None
                kd.feature_mse(torch.matmul(hs_, wp_), ht_).backward()
            ms, ms_f, ms_l = timeit(pmse), timeit(pmse_fwd), timeit(pmse_lib)
            fl = 2.0 * 65_536 * 896 * 3584 / 1e12
            side["projected_feature_mse_65536x896to3584"] = {
                "ms": ms, "forward_ms": ms_f, "tflops": 3 * fl / (ms / 1e3), "forward_tflops": fl / (ms_f / 1e3), "bound": "tensor",
                "peak_tflops": tf_peak, "frac": 3 * fl / (ms / 1e3) / tf_peak, "forward_frac": fl / (ms_f / 1e3) / tf_peak,
                "peak_source": tf_src, "library_gemm_plus_mse_kernel_ms": ms_l,
                "note": "tokens/s = 65536 / (ms / 1e3); the [65536, 3584] projected tensor never reaches HBM (MSE in the GEMM epilogue)"}
            del hs_, ht_, wp_
        except Exception as exc:  # noqa: BLE001
            side["projected_feature_mse"] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        # the student lm_head (896 x 151 936) on the tensor cores with its fused row statistics, next to the library GEMM on the
        # same operands; bound: tensor pipe, against the measured sustained bf16 peak
        try:
            hd = torch.randn(N, 896, device=dev, dtype=torch.bfloat16)
            wt = (torch.randn(V, 896, device=dev) * 0.05).to(torch.bfloat16)
            out = torch.empty(N, V, device=dev, dtype=torch.bfloat16)
            out2 = torch.empty(N, V, device=dev, dtype=torch.bfloat16)
            ms_own = timeit(lambda: sharded.lm_head_fwd(hd, wt, T=T_KD, out=out2), k=5, w=2)
            ms_lib = timeit(lambda: torch.matmul(hd, wt.t(), out=out), k=5, w=2)
            del out2
            tf = 2.0 * N * V * 896 / (ms_own / 1e3) / 1e12
            side[f"lm_head_gemm_{N}x{V}x896"] = {"ms": ms_own, "tflops": tf, "bound": "tensor", "peak_tflops": tf_peak,
                                                "frac": tf / tf_peak, "peak_source": tf_src, "library_gemm_ms": ms_lib,
                                                "note": "hand-written tcgen05 cta_group::2 kernel incl. fused row statistics; both it and "
                                                        "the library GEMM (torch.matmul) write a preallocated output"}
            del hd, wt, out
        except Exception as exc:  # noqa: BLE001 - a side measurement must not take the headline down
            side["lm_head_gemm"] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

        # one vocabulary shard of BASELINE configs[4] on THIS GPU (65 536 tokens x 18 992 of the 151 936 columns = what each of 8 GPUs
        # computes; no exchange partner, so the collectives are no-ops): head GEMM + vocabulary-sharded LoCa KD, forward + backward
        if world == 1:
            try:
                n_s, v_s = 65536, V // 8
                hd = torch.randn(n_s, 896, device=dev, dtype=torch.bfloat16)
                ws_ = (torch.randn(v_s, 896, device=dev) * 0.02).to(torch.bfloat16)
                ts_ = (torch.randn(n_s, v_s, device=dev) * 2.0).to(torch.bfloat16)
                lb_ = torch.randint(0, v_s, (n_s,), device=dev)

                def shard_step():
                    h_, w_ = hd.detach().requires_grad_(True), ws_.detach().requires_grad_(True)
                    sharded.lm_head_kd_vocab_sharded(h_, w_, ts_, lb_, 0, V, T_KD).backward()
                ms = timeit(shard_step, k=5, w=2)
                fl = 3 * 2.0 * n_s * v_s * 896
                side["lm_head_kd_one_shard_65536x18992"] = {
                    "ms": ms, "tokens_per_s": n_s / (ms / 1e3), "tflops": fl / (ms / 1e3) / 1e12, "bound": "tensor",
                    "peak_tflops": tf_peak, "frac": fl / (ms / 1e3) / 1e12 / tf_peak,
                    "note": "one of the 8 shards of configs[4] on one GPU: tcgen05 head with fused statistics, streamed KD phases "
                            "(stats / accum / emit), dH and dW on the same GEMM; the 8-GPU step adds ~0.7 ms of exchanges"}
                del hd, ws_, ts_, lb_
            except Exception as exc:  # noqa: BLE001
                side["lm_head_kd_one_shard"] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    # ---- BASELINE configs[4] at N >= 2 (all ranks take part) ----------------------------------------------
    vshard = None
    if world > 1 and not args.no_secondary:
        try:
            vshard = vshard_line(sharded, synthetic, dist, world, rank, dev)
        except Exception as exc:  # noqa: BLE001
            vshard = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()
        if rank == 0:
            side["lm_head_kd_vshard"] = vshard

    # ---- the token-sharded step with the student's cross-entropy fused in (kd_loca_emit_ce; all ranks take part) -------------
    if world > 1 and mode == "reference" and not args.no_secondary:
        try:
            gce = torch.Generator(device="cpu").manual_seed(77 + rank)
            cel = torch.randint(0, V, (N,), generator=gce)
            cel[torch.rand(N, generator=gce) < 0.5] = -100
            cel = cel.to(dev)

            def ce_step():
                return sharded.loca_kl_token_sharded(teacher, student.detach(), labels, T_KD, ALPHA, phases=phases, exchange=exchange,
                                                     counts=counts, ce_labels=cel)
            for _ in range(3):
                l_ce, c_ce, _g = ce_step()
            sync_all()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                l_ce, c_ce, _g = ce_step()
            c1.record()
            sync_all()
            ms_ce = torch.tensor([c0.elapsed_time(c1) / 10], device=dev)
            dist.all_reduce(ms_ce, op=dist.ReduceOp.MAX)
            # parity: the KD part against the plain step's loss, the CE against the library cross-entropy of the concatenated batch
            part = torch.zeros(2, dtype=torch.float64, device=dev)
            for r0 in range(0, N, 2048):
                part[0] += torch.nn.functional.cross_entropy(student.detach()[r0:r0 + 2048].float(), cel[r0:r0 + 2048], ignore_index=-100,
                                                             reduction="sum").double()
            part[1] = (cel != -100).sum()
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
            ce_ref = (part[0] / part[1]).item()
            del _g
            torch.cuda.empty_cache()
            if rank == 0:
                kd_part = l_ce.item() - c_ce.item()
                side["loca_reference_plus_fused_ce_token_sharded"] = {
                    "ms": ms_ce.item(), "tokens_per_s": n_total / (ms_ce.item() / 1e3), "valid_target_fraction": 0.5,
                    "ce": c_ce.item(), "library_cross_entropy_of_the_concatenated_batch": ce_ref,
                    "ce_rel_err": abs(c_ce.item() - ce_ref) / abs(ce_ref),
                    "kd_part_rel_err_vs_plain_step": abs(kd_part - loss_value) / abs(loss_value),
                    "ok": abs(c_ce.item() - ce_ref) <= 1e-5 * abs(ce_ref) and abs(kd_part - loss_value) <= 1e-4 * abs(loss_value),
                    "note": "KD + CE of the concatenated batch from one pass over each rank's logits; the CE's divisor is summed over "
                            "the ranks on the device; kd_part = loss - ce in fp32 (the CE is ~15x the KD term, hence the 1e-4 gate)"}
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                side["loca_reference_plus_fused_ce_token_sharded"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- end-to-end through the public API with HOST buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        ht = torch.empty((N, V), dtype=torch.bfloat16, pin_memory=True)
        hs = torch.empty((N, V), dtype=torch.bfloat16, pin_memory=True)
        hl = torch.empty((N,), dtype=torch.int64, pin_memory=True)
        ht.copy_(teacher, non_blocking=True)
        hs.copy_(student.detach(), non_blocking=True)
        hl.copy_(labels, non_blocking=True)
        torch.cuda.synchronize()
        dt_, ds_, dl_ = torch.empty_like(teacher), torch.empty_like(teacher), torch.empty_like(labels)
        ds_.requires_grad_(True)

        def e2e_step():
            dt_.copy_(ht, non_blocking=True)
            ds_.data.copy_(hs, non_blocking=True)
            dl_.copy_(hl, non_blocking=True)
            ds_.grad = None
            if world == 1 or mode != "reference":
                l_ = kd.loca_kl(dt_, ds_, dl_, T_KD, alpha=ALPHA, mode=mode, n_total=n_total)
                l_.backward()
            else:
                l_, g_ = sharded.loca_kl_token_sharded(dt_, ds_.detach(), dl_, T_KD, ALPHA, phases=phases, exchange=exchange, counts=counts)
            return float(l_.item())          # device -> host read of the step's result

        ke = max(2, min(args.steps, 5))
        e2e_step()
        sync_all()
        e0.record()
        for _ in range(ke):
            e2e_step()
        e1.record()
        sync_all()
        ms_e = torch.tensor([e0.elapsed_time(e1) / ke], device=dev)
        if world > 1:
            dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total / (ms_e.item() / 1e3), "unit": "tokens/s", "ms_per_step": ms_e.item(),
               "h2d_bytes_per_step": 2 * N * V * 2 + N * 8, "d2h_bytes_per_step": 4, "steps": ke,
               "h2d_gbs": (2 * N * V * 2 + N * 8) / (ms_e.item() / 1e3) / 1e9,
               "note": "pinned host logits + labels copied H2D every step, loss read back; per-GPU bytes; the step is the "
                       "host->device copy of 9.96 GB (h2d_gbs = the PCIe rate it ran at) plus ~4 ms of kernels"}
        del ht, hs, hl, dt_, ds_, dl_

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            best, threads = cpu_reference_step_time(repeats=3, warmup=1)
            cpu = {"value": CPU_SAMPLE_ROWS / best, "unit": "tokens/s", "cores": threads, "kind": "port",
                   "sample": f"oracle/reference_port.py eager LoCa fwd+bwd, fp32, B=1,S={CPU_SAMPLE_ROWS},V={V} "
                             f"(BASELINE configs[0]), best of 3 after 1 warm-up = {best:.3f} s on {cpu_model_name()}"}
        # our kernels per step (the helper-stream pair queue_init + streaming emit only runs with clusters of 8):
        #   1 GPU, as-written: status_clear, teacher_stats, owner_init, owner_scatter, table_fill, table_consts, fused emit,
        #                      loss_finalize, regrad (early exit)                                                    = 9
        #   1 GPU, per-row   : status_clear, teacher_stats, fused emit, loss_finalize, regrad                       = 5
        #   N GPUs, as-written (phase-split ABI): teacher_stats[_peers], owner_init, owner_scatter, table_fill, table_consts,
        #                      fused emit, loss_finalize                                                             = 7
        if mode == "reference":
            launches_per_step = 9 if world == 1 else 7
        else:
            launches_per_step = 5
        out = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"LoCa logit-KD fwd+bwd, B=8,S=2048 ({N} rows/GPU), V={V}, T={T_KD}, alpha={ALPHA}",
                       "mode": mode, "parallelism": f"token-sharded x{world}" if world > 1 else "single GPU",
                       **({"exchange": "phase A stores records into every rank's buffer over NVLink (symmetric memory) + device barrier"
                           if exchange == "peer" else "NCCL all-gather of 16 B/token records"} if (world > 1 and mode == "reference") else {}),
                       "l2": "inputs (2 x 4.98 GB/GPU) exceed the 126 MB L2; no flush needed"},
            "loss": loss_value, "parity": parity, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks.summary(), "secondary": side,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
