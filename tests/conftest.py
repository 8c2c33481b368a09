import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Golden:
    """Lazy view over one tests/golden/*.npz with 'group/case/key' names."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name))

    def cases(self, group):
        return [str(c) for c in self.z[f'{group}/_cases']]

    def case(self, group, case):
        pre = f'{group}/{case}/'
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def sub(self, prefix):
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


@pytest.fixture(scope='session')
def ops_golden():
    return golden('ops.npz')


@pytest.fixture(scope='session')
def modconv_golden():
    return golden('modconv.npz')


@pytest.fixture(scope='session')
def tiny_golden():
    return golden('tiny.npz')


def rel_err(a, b):
    """max |a-b| / max |b|  -- the 'max relative error' of BASELINE.json's north_star."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
