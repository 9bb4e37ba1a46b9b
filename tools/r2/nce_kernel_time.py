"""Device time of the InfoNCE forward+backward launch(es) at 64 x 729 x 1152 through the C ABI (no torch allocations in the loop)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib
lib = _lib.load()
M, P, D = 64, 729, 1152
xs = torch.randn(M, P, D, device="cuda", dtype=torch.bfloat16)
xt = (xs.float() + 0.5 * torch.randn_like(xs, dtype=torch.float32)).to(torch.bfloat16)
grad = torch.empty_like(xs); loss = torch.empty((), device="cuda")
wsb = _lib.size_query("kd_info_nce_workspace_bytes", M, P, D)
ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
def call():
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.kd_info_nce_fwd_bwd(xs.data_ptr(), xt.data_ptr(), grad.data_ptr(), loss.data_ptr(), ws.data_ptr(), wsb, M, P, D, 0, 0.07, None, 0, st))
for _ in range(5): call()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): call()
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); g.replay(); b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 10 * 1e3
print(f"info_nce fwd+bwd device time {us:.1f} us = {3 * xs.numel() * 2 / us / 1e3:.0f} GB/s (floor 49 us at 6548.8 GB/s) loss {loss.item():.6e}")
