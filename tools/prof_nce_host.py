import os, sys, cProfile, pstats
sys.path.insert(0, os.getcwd())
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
xs = torch.randn(64, 729, 1152, device="cuda", dtype=torch.bfloat16)
xt = (xs.float() + 0.5 * torch.randn_like(xs, dtype=torch.float32)).to(torch.bfloat16)
xs.requires_grad_(True)
def plain():
    xs.grad = None
    kd.info_nce(xs, xt, 0.07).backward()
for _ in range(10): plain()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200): plain()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
