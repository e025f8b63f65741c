import json
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import advspec_loader  # noqa: E402

advspec_loader.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


_DIAG: dict = {}


@pytest.fixture(scope="session")
def diag():
    """Numbers worth keeping from a GPU run (max errors, timings); dumped to gpurun_out/."""
    yield _DIAG


def pytest_sessionfinish(session, exitstatus):
    if _DIAG:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        with open(out / "test_diag.json", "w") as f:
            json.dump(_DIAG, f, indent=1, sort_keys=True)


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return 0
