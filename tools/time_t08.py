"""LoCa as-written step at the reference's phase-2 temperature (T = 0.8): every row has clamped cells (masked path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
N = 16384
for T, peaked in ((2.0, False), (0.8, False), (0.8, True)):
    t, s, lab = synthetic.logit_batch(N, synthetic.VOCAB, seed=1, device="cuda", peaked=peaked)
    s.requires_grad_(True)
    def step():
        s.grad = None
        kd.loca_kl(t, s, lab, T).backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): step()
    b.record(); torch.cuda.synchronize()
    print(f"T={T} peaked={peaked}: {a.elapsed_time(b)/10:.3f} ms/step")
    del t, s, lab
