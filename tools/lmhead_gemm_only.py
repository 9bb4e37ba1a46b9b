"""kd_lm_head_fwd alone (no allocation inside the timed loop) next to the library GEMM on the SAME random operands, 20 back-to-back
launches each (operand entropy changes tensor-core power and with it the clock: compare like with like)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import _lib
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200.functional import _stream_ptr

lib = _lib.load()
for n, v, h in [(65536, 18992, 896), (16384, 151936, 896), (16384, 8192, 3584)]:
    hidden = torch.randn(n, h, device="cuda").to(torch.bfloat16)
    weight = (torch.randn(v, h, device="cuda") * 0.1).to(torch.bfloat16)
    logits = torch.empty(n, v, device="cuda", dtype=torch.bfloat16)
    parts = torch.empty(lib.kd_lm_head_stat_tiles(v), n, 2, device="cuda")
    st = _stream_ptr(hidden.device)
    for label, pp in (("product only", None), ("+ student statistics", parts.data_ptr())):
        def call():
            _lib.check(lib.kd_lm_head_fwd(hidden.data_ptr(), h, weight.data_ptr(), h, logits.data_ptr(), v, pp, n, v, h, _lib.DTYPE_BF16, 2.0, st))
        for _ in range(3):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{n} x {v} x {h} {label}: {ms:.3f} ms, {2.0 * n * v * h / ms / 1e9:.0f} TFLOP/s", flush=True)
    for _ in range(3):
        torch.matmul(hidden, weight.t(), out=logits)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        torch.matmul(hidden, weight.t(), out=logits)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{n} x {v} x {h} library GEMM (torch.matmul): {ms:.3f} ms, {2.0 * n * v * h / ms / 1e9:.0f} TFLOP/s", flush=True)
    del hidden, weight, logits, parts
