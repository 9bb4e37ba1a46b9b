"""Backward products of one lm_head shard and the configs[3] projection on the tcgen05 kernel next to the library GEMM (dev tool)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
def timed(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n, V, H = 65536, 18992, 896
g = (torch.randn(n, V, device="cuda") * 1e-3).to(torch.bfloat16)
W = (torch.randn(V, H, device="cuda") * 0.02).to(torch.bfloat16)
hid = torch.randn(n, H, device="cuda").to(torch.bfloat16)
fl = 2.0 * n * V * H / 1e9
o32 = torch.empty(n, H, device="cuda", dtype=torch.float32); o16 = torch.empty(V, H, device="cuda", dtype=torch.bfloat16)
t = timed(lambda: kd.gemm(g, W, b_mn_major=True, out_dtype=torch.float32, out=o32)); tl = timed(lambda: torch.matmul(g, W))
print(f"dH  [65536x18992]x[18992x896] fp32 out: tcgen05 {t:.3f} ms ({fl/t:.0f} TF/s) library(bf16 out) {tl:.3f} ms ({fl/tl:.0f} TF/s)", flush=True)
t = timed(lambda: kd.gemm(g, hid, a_mn_major=True, b_mn_major=True, out=o16)); tl = timed(lambda: torch.matmul(g.t(), hid))
print(f"dW  [18992x65536]x[65536x896]: tcgen05 {t:.3f} ms ({fl/t:.0f} TF/s) library {tl:.3f} ms ({fl/tl:.0f} TF/s)", flush=True)
del g, W, o32, o16
M, K, N = 65536, 896, 3584
x = torch.randn(M, K, device="cuda").to(torch.bfloat16).requires_grad_(True)
w = (torch.randn(K, N, device="cuda") * 0.03).to(torch.bfloat16).requires_grad_(True)
tg = torch.randn(M, N, device="cuda").to(torch.bfloat16)
fl = 2.0 * M * K * N / 1e9
def ours():
    x.grad = None; w.grad = None
    kd.feature_mse(x, tg, proj=w).backward()
def lib():
    x.grad = None; w.grad = None
    kd.feature_mse(torch.matmul(x, w), tg).backward()
def fwd_only():
    with torch.no_grad(): kd.feature_mse(x, tg, proj=w)
t, tl, tf = timed(ours), timed(lib), timed(fwd_only)
print(f"configs[3] projected MSE 65536x896->3584 fwd+bwd: fused epilogue {t:.3f} ms ({3*fl/t:.0f} TF/s)  library GEMM + mse kernel {tl:.3f} ms; forward (loss + grad_projected) {tf:.3f} ms ({fl/tf:.0f} TF/s)", flush=True)
