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
