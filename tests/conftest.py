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
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name, device='cpu'):
    """-> (dict of tensors, state-dict) from tests/golden/<name>.npz (written by tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    arrays, sd = {}, {}
    for k in z.files:
        t = torch.from_numpy(z[k]).to(device)
        if k.startswith('sd/'):
            sd[k[3:]] = t
        else:
            arrays[k] = t
    return arrays, sd


@pytest.fixture(scope='session')
def golden():
    return load_golden
