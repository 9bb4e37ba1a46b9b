"""Small end-to-end cases for compute-sanitizer (memcheck / synccheck / initcheck): every logit-KD variant once."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
for (N, V, T) in ((21, 32000, 2.0), (20, 32000, 0.8), (3, 151936, 2.0)):
    t, s, lab = synthetic.logit_batch(N, V, seed=N, device="cuda", peaked=True)
    s.requires_grad_(True)
    for fn in (lambda: kd.loca_kl(t, s, lab, T), lambda: kd.loca_kl(t, s, lab, T, mode="per_row"), lambda: kd.temperature_kl(t, s, T),
               lambda: kd.temperature_kl(t, s, T, log_target_quirk=True)):
        s.grad = None
        loss = fn()
        loss.backward()
        torch.cuda.synchronize()
        print(N, V, T, float(loss), flush=True)
print("done")
