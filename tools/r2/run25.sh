#!/bin/bash
cd $GRAFT_REPO_ROOT
for vl in 18992 75968; do
timeout 600 python tools/lmhead_time.py 65536 $vl 2>&1 | tail -5
timeout 300 python tools/vshard_phase_times.py 65536 $vl 2>&1 | tail -7
done
python - <<'PY'
import torch, sys, os
sys.path.insert(0, os.getcwd())
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
def timed(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n, H = 65536, 896
for V in (18992, 75968):
    g = (torch.randn(n, V, device="cuda") * 1e-3).to(torch.bfloat16)
    W = (torch.randn(V, H, device="cuda") * 0.02).to(torch.bfloat16)
    hid = torch.randn(n, H, device="cuda").to(torch.bfloat16)
    fl = 2.0 * n * V * H / 1e9
    o32 = torch.empty(n, H, device="cuda", dtype=torch.float32); o16 = torch.empty(V, H, device="cuda", dtype=torch.bfloat16)
    t = timed(lambda: kd.gemm(g, W, b_mn_major=True, out_dtype=torch.float32, out=o32)); tl = timed(lambda: torch.matmul(g, W))
    print(f"V={V} dH: tcgen05 {t:.3f} ms ({fl/t:.0f} TF/s) library {tl:.3f} ms ({fl/tl:.0f} TF/s)", flush=True)
    t = timed(lambda: kd.gemm(g, hid, a_mn_major=True, b_mn_major=True, out=o16)); tl = timed(lambda: torch.matmul(g.t(), hid))
    print(f"V={V} dW: tcgen05 {t:.3f} ms ({fl/t:.0f} TF/s) library {tl:.3f} ms ({fl/tl:.0f} TF/s)", flush=True)
    del g, W, hid, o32, o16
PY
