"""Per-phase times of the vocabulary-sharded KD on one shard (one GPU): stats / combine / table / accum / emit, with the HBM
floor of each (bytes streamed / MEASURED_PEAKS hbm_gbs)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib, sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
v_local = int(sys.argv[2]) if len(sys.argv) > 2 else 18992
mode = sys.argv[3] if len(sys.argv) > 3 else "reference"
V, T, eps, alpha = 151936, 2.0, 1e-8, 0.8
dev = "cuda"
peak = 6548.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
g = torch.Generator(device=dev).manual_seed(0)
teacher = (torch.randn(n, v_local, device=dev, generator=g) * 2).to(torch.bfloat16)
student = (torch.randn(n, v_local, device=dev, generator=g) * 2).to(torch.bfloat16)
labels = torch.randint(0, v_local, (n,), device=dev, generator=g)
variant = sharded._VS_VARIANTS[mode]
ph = sharded.CudaVocabPhases()


def timed(fn, reps=5):
    for _ in range(2):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


cell = n * v_local * 2 / 1e9      # GB per logit matrix
t_stats, stats = timed(lambda: ph.stats(teacher, student, labels, 0, T))
t_comb, (rec, lse2, rowconst, _st) = timed(lambda: ph.combine(stats.unsqueeze(0), labels, V, T, alpha, variant))
table = None
t_tab = 0.0
if variant == _lib.LOCA_REFERENCE:
    t_tab, (table, consts) = timed(lambda: ph.build_table(rec, V))
t_acc, partial = timed(lambda: ph.accum(teacher, student, lse2, rec, table, 0, T, eps, variant))
q_rows = partial[:, 1].contiguous()
t_emit, grad = timed(lambda: ph.emit(teacher, student, lse2, q_rows, rec, table, 0, T, eps, variant, 1.0 / n, None))
print(f"n_rows {n}, V_local {v_local}, mode {mode}; HBM peak {peak:.0f} GB/s")
for name, t, gb in (("stats", t_stats, 2 * cell), ("combine", t_comb, 0), ("table", t_tab, 0), ("accum", t_acc, 2 * cell), ("emit", t_emit, 3 * cell)):
    floor = gb / peak * 1e3
    print(f"  {name:8s} {t:7.3f} ms" + (f"   floor {floor:.3f} ms -> {floor / t * 100:.0f} %" if gb else ""))
