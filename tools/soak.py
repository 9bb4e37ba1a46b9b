"""Soak run: many back-to-back steps of every logit-KD variant (looking for rare hangs in the cluster protocols)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for N in (16384, 4099, 37):
    t, s, lab = synthetic.logit_batch(N, synthetic.VOCAB, seed=N, device="cuda", peaked=True)
    s.requires_grad_(True)
    cel = torch.randint(0, synthetic.VOCAB, (N,), device="cuda")
    cel[::2] = -100
    only = os.environ.get("SOAK_ONLY")
    for name, fn in (("loca T=2", lambda: kd.loca_kl(t, s, lab, 2.0)), ("loca T=0.8", lambda: kd.loca_kl(t, s, lab, 0.8)),
                     ("per_row T=0.8", lambda: kd.loca_kl(t, s, lab, 0.8, mode="per_row")), ("kl T=0.8", lambda: kd.temperature_kl(t, s, 0.8)),
                     ("loca+ce T=2", lambda: kd.loca_kl_ce(t, s, lab, cel, 2.0)[0]), ("loca+ce T=0.8", lambda: kd.loca_kl_ce(t, s, lab, cel, 0.8)[0]),
                     ("loca+ce T=1", lambda: kd.loca_kl_ce(t, s, lab, cel, 1.0)[0]),
                     ("kl_log_target T=0.8", lambda: kd.temperature_kl(t, s, 0.8, log_target_quirk=True))):
        if only and only not in name:
            continue
        k = steps if N == 16384 else steps * 4
        t0 = time.perf_counter()
        first = None
        for i in range(k):
            s.grad = None
            loss = fn()
            loss.backward()
            if i == 0:
                first = loss.item()
        last = loss.item()
        torch.cuda.synchronize()
        print(f"N={N} {name}: {k} steps in {time.perf_counter() - t0:.1f} s, loss first {first:.9e} last {last:.9e} same={first == last}", flush=True)
    del t, s, lab
