import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def sub(d, prefix):
    """entries of a flat npz dict under `prefix.` as torch tensors"""
    n = len(prefix) + 1
    return {k[n:]: torch.from_numpy(v) for k, v in d.items() if k.startswith(prefix + '.')}


def rel_err(a, b):
    """max-norm relative error: max|a-b| / max|b| (the tolerance definition used throughout:
    1e-5 fp32 / 1e-3 fp16 storage, BASELINE.json north_star)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope='session')
def golden():
    return load_golden
