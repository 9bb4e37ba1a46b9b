"""Backward re-emission (grad_output != what forward assumed, e.g. `(0.5 * loss).backward()` or a dynamic AMP loss scale):
time of the regrad kernel alone and of the whole forward + backward step, 16 384 x 151 936 bf16."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

N, V = 16384, synthetic.VOCAB
t, s, lab = synthetic.logit_batch(N, V, seed=1, device="cuda")
s.requires_grad_(True)


def timed(fn, k=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


for mode in ("reference", "per_row"):
    def plain():
        s.grad = None
        kd.loca_kl(t, s, lab, 2.0, mode=mode).backward()

    def weighted():
        s.grad = None
        (0.5 * kd.loca_kl(t, s, lab, 2.0, mode=mode)).backward()

    a, b = timed(plain), timed(weighted)
    print(f"{mode:10s} step {a:.3f} ms   step with re-emission {b:.3f} ms   re-emission {b - a:.3f} ms = "
          f"{6 * V * N / ((b - a) / 1e3) / 1e9:.0f} GB/s")
