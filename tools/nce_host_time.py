"""Pooled InfoNCE (BASELINE configs[2]): eager fwd+bwd per call (CUDA events), host time per call (wall clock with the GPU
left behind), and the same step replayed from a CUDA graph (pure device time of the launch chain)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd

dev = "cuda"
xs = torch.randn(64, 729, 1152, device=dev, dtype=torch.bfloat16)
xt = (xs.float() + 0.5 * torch.randn_like(xs, dtype=torch.float32)).to(torch.bfloat16)
xs.requires_grad_(True)
wp = (torch.randn(1152, 896, device=dev) * 0.03).to(torch.bfloat16).requires_grad_(True)
xt896 = torch.randn(64, 729, 896, device=dev).to(torch.bfloat16)


def plain():
    xs.grad = None
    kd.info_nce(xs, xt, 0.07).backward()


def proj():
    xs.grad = None
    wp.grad = None
    kd.info_nce(xs, xt896, 0.07, proj=wp).backward()


def measure(name, f, k=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(k):
        f()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    eager = a.elapsed_time(b) / k * 1e3
    host = (t1 - t0) / k * 1e6
    # graph replay
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        f()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(k):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:8s} eager {eager:7.1f} us/step   host issue {host:7.1f} us/step   graph replay {a.elapsed_time(b) / k * 1e3:7.1f} us/step")


which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "plain"):
    measure("plain", plain)
if which in ("both", "proj"):
    measure("proj896", proj)
