"""Token-level InfoNCE (contrastive_loss on the un-pooled SigLIP tokens, M = 64 x 729 = 46 656 rows of 1152) on the tensor
cores: ms per forward+backward, TFLOP/s over the three M x M x D products, and a check against the fp64 oracle at a size it fits."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
from oracle import kd_oracle as O

for M in (int(a) for a in (sys.argv[1:] or ["8192", "46656"])):
    D = 1152
    xs, xt = synthetic.vision_feature_batch(M, 1, D, seed=7, device="cuda")
    xs, xt = xs[:, 0].contiguous(), xt[:, 0].contiguous()
    xd = xs.clone().requires_grad_(True)
    for _ in range(2):
        xd.grad = None
        kd.info_nce(xd, xt).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 5
    a.record()
    for _ in range(k):
        xd.grad = None
        loss = kd.info_nce(xd, xt)
        loss.backward()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / k
    line = f"M={M} D={D}: {ms:.3f} ms fwd+bwd = {3 * 2 * M * M * D / ms / 1e9:.0f} TFLOP/s over 3 products; loss {loss.item():.6f}"
    if M <= 16384:
        ref = O.info_nce(xs, xt)
        g = xd.grad.double()
        line += f"; oracle loss rel {abs(loss.item() - ref.loss.item()) / abs(ref.loss.item()):.2e}, grad rel {((g - ref.grad).norm() / ref.grad.norm()).item():.2e}"
    print(line, flush=True)
