"""Step time (forward + backward through the autograd op) of every fused-kernel variant at 16 384 x 151 936 bf16 for the cluster
configuration selected by the environment (KD_CLUSTER_SIZE=8|9, KD_SPARE_SMS); run once per setting, see tools/cluster_ab.sh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

N, V = 16384, synthetic.VOCAB
tag = " ".join(f"{k}={os.environ[k]}" for k in ("KD_CLUSTER_SIZE", "KD_SPARE_SMS") if k in os.environ) or "default"


def timed(fn, k=20, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


for peaked in (False, True):
    t, s, lab = synthetic.logit_batch(N, V, seed=1, device="cuda", peaked=peaked)
    s.requires_grad_(True)
    cel = torch.randint(0, V, (N,), device="cuda")
    cel[::2] = -100
    cases = ([("as-written T=2", lambda: kd.loca_kl(t, s, lab, 2.0)), ("per-row T=2", lambda: kd.loca_kl(t, s, lab, 2.0, mode="per_row")),
              ("plain KL T=0.8", lambda: kd.temperature_kl(t, s, 0.8)), ("log_target T=0.8", lambda: kd.temperature_kl(t, s, 0.8, log_target_quirk=True)),
              ("as-written + CE T=2", lambda: kd.loca_kl_ce(t, s, lab, cel, 2.0)[0])] if not peaked else
             [("as-written T=0.8 peaked", lambda: kd.loca_kl(t, s, lab, 0.8)), ("per-row T=0.8 peaked", lambda: kd.loca_kl(t, s, lab, 0.8, mode="per_row")),
              ("as-written T=1 peaked", lambda: kd.loca_kl(t, s, lab, 1.0))])
    for name, fn in cases:
        if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
            continue
        def step():
            s.grad = None
            fn().backward()
        ms = min(timed(step) for _ in range(2))
        print(f"[{tag}] {name}: {ms:.3f} ms", flush=True)
    del t, s, lab, cel
    torch.cuda.empty_cache()
