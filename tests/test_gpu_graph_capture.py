"""The C ABI promises capture safety (no host synchronisation, no allocation inside the library): one LoCa forward+backward
is captured into a CUDA graph and replayed; loss and gradient must match the eager step."""
import pytest
import torch

import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic

pytestmark = pytest.mark.gpu


# 1100 rows of per-row LoCa: clusters of 8 + the streaming helper on a second (internal) stream — a fork / join inside the capture
@pytest.mark.parametrize("mode,rows", [("reference", 96), ("per_row", 96), ("per_row", 1100)])
def test_step_replays_from_a_cuda_graph(mode, rows):
    t, s, lab = synthetic.logit_batch(rows, synthetic.VOCAB, seed=13, device="cuda")
    s.requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):          # warm-up outside the capture: library loads, attributes, allocator pools
        for _ in range(2):
            s.grad = None
            ref = kd.loca_kl(t, s, lab, 2.0, mode=mode)
            ref.backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref_loss, ref_grad = ref.item(), s.grad.clone()
    s.grad = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = kd.loca_kl(t, s, lab, 2.0, mode=mode)
        loss.backward()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert loss.item() == ref_loss
    assert ((s.grad.float() - ref_grad.float()).norm() / ref_grad.float().norm()).item() <= 1e-3
