"""CUDA-graph capture of one LoCa fwd+bwd step (the C ABI claims capture safety: no host sync, no allocation)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
t, s, lab = synthetic.logit_batch(N, synthetic.VOCAB, seed=3, device="cuda")
s.requires_grad_(True)
def step():
    s.grad = None
    loss = kd.loca_kl(t, s, lab, 2.0)
    loss.backward()
    return loss
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        ref_loss = step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
ref_grad = s.grad.clone()
g = torch.cuda.CUDAGraph()
s.grad = None
with torch.cuda.graph(g):
    loss = kd.loca_kl(t, s, lab, 2.0)
    loss.backward()
g.replay(); torch.cuda.synchronize()
same = (loss.item() == ref_loss.item()) and ((s.grad.float() - ref_grad.float()).norm() / ref_grad.float().norm()).item() < 1e-3
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): g.replay()
b.record(); torch.cuda.synchronize()
ms_graph = a.elapsed_time(b) / 20
a.record()
for _ in range(20): step()
b.record(); torch.cuda.synchronize()
print(f"graph replay {ms_graph:.3f} ms/step, eager {a.elapsed_time(b) / 20:.3f} ms/step, identical loss + grad {same}")
sys.exit(0 if same else 1)
