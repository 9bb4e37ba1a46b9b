"""One launch of the tcgen05 student lm_head (with its fused row statistics) at 16 384 x 151 936 x 896 — the target of the ncu
capture summarised in profiles/r2_lm_head_gemm_full.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
v = int(sys.argv[2]) if len(sys.argv) > 2 else 151936
hidden = torch.randn(n, 896, device="cuda").to(torch.bfloat16)
weight = (torch.randn(v, 896, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(3):
    logits, partials = sharded.lm_head_fwd(hidden, weight, T=2.0)
torch.cuda.synchronize()
print("ok", logits.shape, partials.shape)
