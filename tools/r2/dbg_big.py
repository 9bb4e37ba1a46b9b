import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
from oracle import kd_oracle as O
N, V, stride, T = 16384, 151936, 37, 0.8
mode = sys.argv[1] if len(sys.argv) > 1 else "reference"
t, s, lab = synthetic.logit_batch(N, V, seed=1234, device="cuda", peaked=True)
s = (s.float() * 4.0).to(torch.bfloat16)
sd = s.clone().requires_grad_(True)
loss = kd.loca_kl(t, sd, lab, T, alpha=0.8, mode=mode); loss.backward()
ref = O.loca_kl(t, s, lab, T, 0.8, mode=mode, chunk_rows=128, grad_stride=stride)
g = sd.grad
print("loss", loss.item(), ref.loss.item(), "finite ours", torch.isfinite(g).all().item(), "finite ref", torch.isfinite(ref.grad).all().item())
gs = g[:, ::stride].double()
e = ((gs - ref.grad).norm() / ref.grad.norm()).item()
rr = ref.grad.float().to(torch.bfloat16).double()
print("rel err", e, "vs rounded", ((gs - rr).norm() / rr.norm()).item())
rowe = (gs - ref.grad).norm(dim=1)
top = rowe.topk(5)
print("worst rows abs", top.values.tolist(), top.indices.tolist(), "row norms", ref.grad.norm(dim=1)[top.indices].tolist())
r = int(top.indices[0])
d = (gs[r] - ref.grad[r]).abs()
c = int(d.argmax())
print("row", r, "col", c * stride, "ours", gs[r, c].item(), "ref", ref.grad[r, c].item(), "label", int(lab[r]), "k", int(ref.aux["stats"]["k"][r]),
      "Q", ref.aux["Q"][r].item())
