"""Decode the KD_TRACE dump of the two-pipeline fused emit kernel: per-phase durations per warp and row (debug build only)."""
import sys
import numpy as np
nw = 16
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, nw, 64, 8).astype(np.int64)      # [cta][warp][row_iter][stamp]
names = ["tma_wait", "pass0", "passAB", "red_wait", "passC", "refill"]
valid = a[:, :, 3:60, :]
ok = (valid[..., 6] > 0) & (valid[..., 0] > 0)
d = np.diff(valid, axis=-1)
print("ctas", a.shape[0], "samples", ok.sum())
for i, n in enumerate(names):
    x = d[..., i][ok]
    print(f"{n:10s} mean {x.mean():8.0f} ns   p50 {np.median(x):8.0f}   p90 {np.percentile(x, 90):8.0f}")
tot = (valid[..., 6] - valid[..., 0])[ok]
print(f"row total  mean {tot.mean():8.0f} ns")
for w in (0, 3, 7, 8, 15):
    x = d[0, w][ok[0, w]]
    print("cta0 warp", w, " ".join(f"{n}={x[:, i].mean():.0f}" for i, n in enumerate(names)))
# timeline of cta 0, warps 0 and 8 (the two pipelines), rows 10..13, relative to the first stamp
t0 = a[0, 0, 10, 0]
for w in (0, 8):
    for r in range(10, 14):
        print("cta0 warp", w, "row_iter", r, " ".join(str(int(x - t0)) for x in a[0, w, r, :8]))
# fold warps: push (stamp 3) -> all partials arrived (stamp 7) -> result broadcast (stamp 4)
f = a[:, [0, 8], 3:60, :]
okf = (f[..., 7] > 0) & (f[..., 3] > 0)
print("fold warp: push->arrived mean", (f[..., 7] - f[..., 3])[okf].mean(), " arrived->broadcast mean", (f[..., 4] - f[..., 7])[okf].mean())
# skew between the CTAs of cluster 0 at the push of the slowest warp (warp 0)
c = a[0:8, 0, 3:60, 3]
print("cluster 0: spread of warp-0 push time across its 8 CTAs, mean", (c.max(axis=0) - c.min(axis=0)).mean())
lastpush = a[0:8, 0:8, 3:60, 3].max(axis=(0, 1))
print("cluster 0 pipe 0: last push -> fold-warp arrived (cta0)", (a[0, 0, 3:60, 7] - lastpush).mean())
