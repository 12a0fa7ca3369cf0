import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_to_max(a, b, floor=1e-30):
    """max|a-b| / max(max|b|, floor) (the parity metric of SURVEY.md §7 'hard parts').  `floor` guards
    tensors that are analytically zero (e.g. d<Z>/d(angle) when RZ acts on |0...0>)."""
    import torch
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = max(b.abs().max().item(), floor)
    return (a - b).abs().max().item() / denom
