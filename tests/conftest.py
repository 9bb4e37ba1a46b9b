import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def pytest_terminal_summary(terminalreporter):
    """Worst gradient errors seen by tests/_util.grad_close (gate: 2e-3 relative on the norm; see its docstring)."""
    try:
        from _util import GRAD_ERRORS
    except Exception:
        return
    if not GRAD_ERRORS:
        return
    tier2 = [g for g in GRAD_ERRORS if g[2] is not None]
    terminalreporter.write_line(f"grad_close: {len(GRAD_ERRORS)} checks; {len(GRAD_ERRORS) - len(tier2)} within 2e-3 of the exact oracle, "
                                f"{len(tier2)} judged against the bf16-rounded oracle; worst:")
    for name, e, er in sorted(GRAD_ERRORS, key=lambda x: -x[1])[:8]:
        terminalreporter.write_line(f"  {e:.3e}" + (f" (vs rounded oracle {er:.2e})" if er is not None else "") + f"  {name}")
