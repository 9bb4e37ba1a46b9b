import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic, sharded
N = 16384
t, s, lab = synthetic.logit_batch(N, synthetic.VOCAB, seed=1, device="cuda")
s.requires_grad_(True)
ph = sharded.CudaPhases()
def timeit(fn, k=10, w=3):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / k
def full():
    s.grad = None
    kd.loca_kl(t, s, lab, 2.0).backward()
def fwd_only():
    with torch.no_grad():
        pass
    s.grad = None
    loss = kd.loca_kl(t, s, lab, 2.0)
rec, lse, _ = ph.teacher_stats(t, lab, 2.0, 0.8, synthetic.VOCAB)
table, consts = ph.build_table(rec, synthetic.VOCAB)
print("full step      ", timeit(full))
print("forward only   ", timeit(fwd_only))
print("teacher_stats  ", timeit(lambda: ph.teacher_stats(t, lab, 2.0, 0.8, synthetic.VOCAB)))
print("build_table    ", timeit(lambda: ph.build_table(rec, synthetic.VOCAB)))
print("emit           ", timeit(lambda: ph.emit(t, s.detach(), lse, table, consts, 2.0, 1e-8, 1.0 / (N * synthetic.VOCAB), None)))
def chain():
    r, l, _ = ph.teacher_stats(t, lab, 2.0, 0.8, synthetic.VOCAB)
    tb, c = ph.build_table(r, synthetic.VOCAB)
    ph.emit(t, s.detach(), l, tb, c, 2.0, 1e-8, 1.0 / (N * synthetic.VOCAB), None)
print("stats+table+emit", timeit(chain))
