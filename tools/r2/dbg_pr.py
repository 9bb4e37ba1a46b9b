import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
from oracle import kd_oracle as O
V, T, boost = 151936, 0.8, 4.0
N = 24
t, s, lab = synthetic.logit_batch(N, V, seed=300 + V % 11)
s = (s.float() * boost).to(torch.bfloat16)
t = t.float()
top = t.topk(2, dim=-1)
lab[0] = top.indices[0, 1]; lab[1] = top.indices[1, 0]; lab[2], lab[3] = 0, V - 1
t[4, 8], t[4, 9] = 30.0, 29.0
t[5, V - 1], t[5, 0] = 30.0, 29.0
t[6, V // 9 * 4 - 1], t[6, V // 9 * 4] = 31.0, 30.5
lab[6] = V // 9 * 4
t[7, 5], lab[7] = 8.0 + 16.0 * T, 5
t = t.to(torch.bfloat16)
for r in range(N):
    ref = O.loca_kl(t[r:r+1], s[r:r+1], lab[r:r+1], T, 0.8, mode="per_row")
    sd = s[r:r+1].cuda().requires_grad_(True)
    l = kd.loca_kl(t[r:r+1].cuda(), sd, lab[r:r+1].cuda(), T, alpha=0.8, mode="per_row")
    l.backward()
    ls = kd.loca_kl(t[r:r+1].cuda(), s[r:r+1].cuda(), lab[r:r+1].cuda(), T, alpha=0.8, mode="per_row", force_streaming=True)
    ge = ((sd.grad.float().cpu().double() - ref.grad).norm() / ref.grad.norm()).item()
    st = ref.aux["stats"]
    print(r, f"ours {l.item():.6e} stream {ls.item():.6e} ref {ref.loss.item():.6e} rel {abs(l.item()-ref.loss.item())/abs(ref.loss.item()):.2e} grad {ge:.2e}",
          "lab", int(lab[r]), "k", int(st['k'][0]), f"p_gt {st['p_gt'][0]:.4f} p_k {st['p_k'][0]:.4f} v1 {st['v1'][0]:.4f} v2 {st['v2'][0]:.4f}")
