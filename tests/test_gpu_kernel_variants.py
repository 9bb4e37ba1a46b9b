"""The alternative emit kernels are selected by environment variables that the library reads once, so each variant
runs in its own interpreter: the three-sweep cluster kernel (KD_FUSED=0: the exact per-cell path that serves alpha > 1 and the
log_target variant), clusters of 8 with the streaming helper on the spare SMs (KD_CLUSTER_SIZE=8) and without it
(KD_SPARE_SMS=0).  Each must meet the same gates as the default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from _util import grad_close
from oracle import kd_oracle as O
import knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 as kd
from knowledge_distillation_for_sensory_substitution_in_multimodal_models_b200 import synthetic
ok = True
for (N, V, T, peaked, boost, mode) in [(40, 151936, 2.0, False, 1.0, "reference"), (24, 151936, 0.8, True, 4.0, "reference"),
                                       (33, 32000, 1.0, False, 1.0, "reference"), (17, 151936, 2.0, True, 1.0, "per_row")]:
    t, s, lab = synthetic.logit_batch(N, V, seed=300 + N, peaked=peaked)
    s = (s.float() * boost).to(torch.bfloat16)
    ref = O.loca_kl(t, s, lab, T, 0.8, mode=mode, chunk_rows=8)
    sd = s.cuda().requires_grad_(True)
    loss = kd.loca_kl(t.cuda(), sd, lab.cuda(), T, alpha=0.8, mode=mode)
    loss.backward()
    good = abs(loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item()) and grad_close(sd.grad.float().cpu(), ref.grad)
    print(N, V, T, mode, "loss rel", abs(loss.item() - ref.loss.item()) / abs(ref.loss.item()), "OK" if good else "FAIL")
    ok &= good
t, s, _ = synthetic.logit_batch(21, 151936, seed=9)
ref = O.temperature_kl(t, s, 0.8, chunk_rows=8)
sd = s.cuda().requires_grad_(True)
loss = kd.temperature_kl(t.cuda(), sd, 0.8)
loss.backward()
ok &= abs(loss.item() - ref.loss.item()) <= 1e-5 * abs(ref.loss.item()) and grad_close(sd.grad.float().cpu(), ref.grad)
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("env", [{"KD_FUSED": "0"}, {"KD_CLUSTER_SIZE": "8"}, {"KD_CLUSTER_SIZE": "8", "KD_SPARE_SMS": "0"}], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_alternative_kernels_meet_the_gates(env):
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env={**os.environ, **env}, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
