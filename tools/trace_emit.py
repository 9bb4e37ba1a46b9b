"""Decode the KD_TRACE dump: per-phase durations of the cluster emit kernel (debug build only)."""
import sys
import numpy as np
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 16
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, nw, 64, 8).astype(np.int64)      # [cta][warp][row_iter][stamp]
ncta = a.shape[0]
names = ["tma_wait", "sweepA", "redA_wait", "sweepB", "redB_wait", "sweepC"]
valid = a[:, :, 2:40, :]                      # skip warm-up rows
ok = (valid[..., 6] > 0) & (valid[..., 0] > 0)
d = np.diff(valid, axis=-1)
print("ctas", ncta, "samples", ok.sum())
for i, n in enumerate(names):
    x = d[..., i][ok]
    print(f"{n:10s} mean {x.mean():8.0f} ns   p50 {np.median(x):8.0f}   p90 {np.percentile(x, 90):8.0f}")
tot = (valid[..., 6] - valid[..., 0])[ok]
print(f"row total  mean {tot.mean():8.0f} ns")
nxt = (valid[:, :, 1:, 0] - valid[:, :, :-1, 6])
print(f"refill+loop gap mean {nxt[ok[:, :, 1:] & ok[:, :, :-1]].mean():8.0f} ns")
# per-warp view for cta 0
for w in (0, nw // 4, nw // 2 - 1, nw // 2, nw - 2):
    x = d[0, w][ok[0, w]]
    print("cta0 warp", w, " ".join(f"{n}={x[:, i].mean():.0f}" for i, n in enumerate(names)))
