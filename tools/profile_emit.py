"""Tiny driver for ncu: a few LoCa steps at a reduced row count (same per-row behaviour as the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "reference"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t, s, lab = synthetic.logit_batch(rows, synthetic.VOCAB, seed=1, device="cuda")
s.requires_grad_(True)
for _ in range(steps):
    s.grad = None
    loss = kd.loca_kl(t, s, lab, 2.0, mode=mode)
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
