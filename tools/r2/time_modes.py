"""Time the logit-KD modes at the bench shape through the public ops (dev tool): ms per fwd+bwd step."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
N = int(os.environ.get("ROWS", 16384)); V = 151936
modes = sys.argv[1:] or ["per_row", "per_row_T0.8p", "reference", "reference_T0.8p", "kl", "kl_T0.8", "kl_lt_T0.8"]
t, s, lab = synthetic.logit_batch(N, V, seed=1234, device="cuda")
tp, sp_, labp = synthetic.logit_batch(N, V, seed=1234, device="cuda", peaked=True)
sp_ = (sp_.float() * 4).to(torch.bfloat16)
s.requires_grad_(True); sp_.requires_grad_(True)
g = torch.Generator().manual_seed(3)
cel = torch.randint(0, V, (N,), generator=g)
cel[torch.rand(N, generator=g) < 0.5] = -100          # half the rows are prompt / padding
cel = cel.cuda()
def run(name):
    if name == "reference_ce": return kd.loca_kl_ce(t, s, lab, cel, 2.0)[0]
    if name == "reference_T0.8p_ce": return kd.loca_kl_ce(tp, sp_, labp, cel, 0.8)[0]
    if name == "reference_T1_ce": return kd.loca_kl_ce(t, s, lab, cel, 1.0)[0]
    if name == "per_row": return kd.loca_kl(t, s, lab, 2.0, mode="per_row")
    if name == "reference": return kd.loca_kl(t, s, lab, 2.0)
    if name == "per_row_T0.8p": return kd.loca_kl(tp, sp_, labp, 0.8, mode="per_row")
    if name == "reference_T0.8p": return kd.loca_kl(tp, sp_, labp, 0.8)
    if name == "reference_T1": return kd.loca_kl(t, s, lab, 1.0)
    if name == "kl": return kd.temperature_kl(t, s, 2.0)
    if name == "kl_T0.8": return kd.temperature_kl(t, s, 0.8)
    if name == "kl_lt_T0.8": return kd.temperature_kl(t, s, 0.8, log_target_quirk=True)
for m in modes:
    for _ in range(3):
        s.grad = None; sp_.grad = None; run(m).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 10
    a.record()
    for _ in range(k):
        s.grad = None; sp_.grad = None; l = run(m); l.backward()
    b.record(); torch.cuda.synchronize()
    print(f"{m:18s} {a.elapsed_time(b)/k:7.3f} ms/step  loss {l.item():.6e}", flush=True)
