import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
xs = torch.randn(64, 729, 1152, device="cuda", dtype=torch.bfloat16)
xt = (xs.float() + 0.5 * torch.randn_like(xs, dtype=torch.float32)).to(torch.bfloat16)
xs.requires_grad_(True)
for _ in range(4):
    xs.grad = None
    loss = kd.info_nce(xs, xt, 0.07)
    loss.backward()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    xs.grad = None
    loss = kd.info_nce(xs, xt, 0.07)
    loss.backward()
b.record(); torch.cuda.synchronize()
print("info_nce fwd+bwd us", a.elapsed_time(b) / 20 * 1e3, "loss", loss.item())
