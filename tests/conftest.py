"""pytest configuration: registers the `gpu` marker and puts the repo root / oracle on sys.path.

`-m "not gpu"` runs everywhere (oracle vs golden vectors, host logic, C-ABI symbol checks, gloo sharding);
`-m gpu` are the parity tests proper: they call the CUDA kernels through the C ABI on a real B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_reference(monkeypatch):
    """Schedulers built inside the test reproduce the reference as executed on CPU tensors (true divisions, 16-bit
    scalar rounding) — the rules the CPU-made fixtures (tests/golden/{sd,sd16,fm}_*) and the oracle's default
    functions follow.  Without it the product default applies: the reference as executed on CUDA tensors (cuda_*)."""
    from consolver_b200 import _sched_common

    monkeypatch.setattr(_sched_common, "DEFAULT_REFERENCE_DEVICE", "cpu")
