"""Times the three streaming kernels whose loads-in-flight batch is a compile-time knob (KD_REGRAD_BATCH in logit_kd.cu,
KD_MSE_BATCH in feature_mse.cu, KD_POOL_BATCH in info_nce.cu) for the library selected with KD_B200_LIB
(tools/build_variant.sh <name> --src <file> -DKD_..._BATCH=n).  Usage: python tools/knob_sweep.py [regrad] [mse] [pool]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

what = sys.argv[1:] or ["regrad", "mse", "pool"]
tag = os.environ.get("KD_B200_LIB", "default")


def timed(fn, k=10, w=3):
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


def graph_timed(fn, k=50):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timed(g.replay, k=k)


if "regrad" in what:
    N, V = 16384, synthetic.VOCAB
    t, s, lab = synthetic.logit_batch(N, V, seed=1, device="cuda")
    s.requires_grad_(True)

    half = torch.tensor(0.5, device="cuda")
    for name, make in (("as-written", lambda: kd.loca_kl(t, s, lab, 2.0)), ("per-row", lambda: kd.loca_kl(t, s, lab, 2.0, mode="per_row")),
                       ("plain KL", lambda: kd.temperature_kl(t, s, 0.8))):
        loss = make()

        def re_emit():   # a backward with grad_output != 1 through the retained node: the re-emission sweep (+ two scalar kernels)
            torch.autograd.grad(loss, s, grad_outputs=half, retain_graph=True)
        best = min(timed(re_emit) for _ in range(3))
        print(f"[{tag}] regrad {name} {best:.3f} ms = {6 * V * N / best / 1e6:.0f} GB/s", flush=True)
        del loss
    del t, s, lab
    torch.cuda.empty_cache()

if "mse" in what:
    hs, ht, _w = synthetic.hidden_state_batch(65536, 896, 3584, device="cuda")
    a = torch.randn_like(ht).requires_grad_(True)

    def mse():
        a.grad = None
        kd.feature_mse(a, ht).backward()
    best = min(timed(mse, k=20) for _ in range(3))
    print(f"[{tag}] mse {best * 1e3:.1f} us = {6 * a.numel() / best / 1e6:.0f} GB/s", flush=True)
    del hs, ht, a
    torch.cuda.empty_cache()

if "pool" in what:
    xs, xt = synthetic.vision_feature_batch(64, 729, 1152, device="cuda")
    xs.requires_grad_(True)

    def nce():
        xs.grad = None
        kd.info_nce(xs, xt).backward()
    best = min(graph_timed(nce) for _ in range(3))
    print(f"[{tag}] pooled info_nce (graph replay) {best * 1e3:.1f} us", flush=True)
