"""One shard of BASELINE configs[4] on one GPU: student lm_head (896 x V_local) + vocabulary-sharded LoCa KD, forward and
forward+backward, with the hand-written tcgen05 head (statistics fused into its epilogue) and with the library GEMM."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
v_local = int(sys.argv[2]) if len(sys.argv) > 2 else 18992
V, H, T = 151936, 896, 2.0
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
hidden = torch.randn(n, H, device=dev, generator=g).to(torch.bfloat16)
weight = (torch.randn(v_local, H, device=dev, generator=g) * 0.1).to(torch.bfloat16)
teacher = (torch.randn(n, v_local, device=dev, generator=g) * 2).to(torch.bfloat16)
labels = torch.randint(0, v_local, (n,), device=dev, generator=g)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"n_rows {n}, V_local {v_local}, H {H}", flush=True)
t_own = timed(lambda: sharded.lm_head_fwd(hidden, weight))
t_own_s = timed(lambda: sharded.lm_head_fwd(hidden, weight, T=T))
t_lib = timed(lambda: torch.matmul(hidden, weight.t()))
fl = 2.0 * n * v_local * H / 1e9
print(f"head GEMM: tcgen05 {t_own:.3f} ms ({fl / t_own:.0f} TFLOP/s), with statistics {t_own_s:.3f} ms, library {t_lib:.3f} ms ({fl / t_lib:.0f} TFLOP/s)", flush=True)
for gemm in ("tcgen05", "library"):
    def fwd():
        with torch.no_grad():
            return sharded.lm_head_kd_vocab_sharded(hidden, weight, teacher, labels, 0, V, T, gemm=gemm)

    def step():
        h, w = hidden.detach().requires_grad_(True), weight.detach().requires_grad_(True)
        sharded.lm_head_kd_vocab_sharded(h, w, teacher, labels, 0, V, T, gemm=gemm).backward()

    tf, ts = timed(fwd), timed(step)
    print(f"{gemm:8s}: head + KD forward {tf:.3f} ms, forward + backward {ts:.3f} ms = {n / ts * 1e3:.0f} tokens/s", flush=True)
