"""Helpers shared by the tests (oracle access lives here: tests are allowed to use ``oracle/``)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def goldens_of_kind(*kinds: str) -> list[str]:
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        z = np.load(p, allow_pickle=False)
        if str(z["kind"]) in kinds:
            out.append(os.path.splitext(os.path.basename(p))[0])
    return out


def rel_err(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


GRAD_ERRORS: list = []      # (test id, relative error) of every grad_close call, summarised at the end of the run (conftest.py)


def grad_close(ours, ref, dtype=torch.bfloat16, tol=2e-3) -> bool:
    """Gradient parity for a 16-bit output: BASELINE.json's gate, 2e-3 relative on the L2 norm — no looser fallback.

    One exact rounding to bf16 costs at most 2^-9 = 1.95e-3 per element (about 1.1e-3 on the norm of smoothly distributed
    values), so the gate leaves the kernels ~1e-4 of their own.  Every measured error is recorded and the worst cases are
    printed in the terminal summary, so the margin is visible in GPUTEST logs.
    """
    import os
    ours = torch.as_tensor(ours, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    e = rel_err(ours, ref)
    GRAD_ERRORS.append((os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], e))
    return e <= tol
