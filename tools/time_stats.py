"""Time phase A (teacher statistics) alone: 16 384 x 151 936 bf16 rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic, sharded
N = 16384
t, s, lab = synthetic.logit_batch(N, synthetic.VOCAB, seed=1, device="cuda")
del s
ph = sharded.CudaPhases()
def timeit(fn, k=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
ms = timeit(lambda: ph.teacher_stats(t, lab, 2.0, 0.8, synthetic.VOCAB))
rec, lse, _ = ph.teacher_stats(t, lab, 2.0, 0.8, synthetic.VOCAB)
print("teacher_stats ms %.4f  GB/s %.0f  checksum %.6f %d" % (ms, N * synthetic.VOCAB * 2 / ms / 1e6, lse.double().sum().item(), int(rec.view(torch.int32).long().sum().item()) if hasattr(rec, "view") else 0))
