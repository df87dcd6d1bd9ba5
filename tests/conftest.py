import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def assets():
    """Synthetic model tensors / prior / regressor shared by all tests (seed 0)."""
    from bodyfitting_b200 import synthetic as syn
    cache = {}

    def get(kind):
        if kind not in cache:
            if kind in ('smpl', 'smplx'):
                cache[kind] = syn.make_model(kind, 0)
            elif kind == 'gmm':
                cache[kind] = syn.make_gmm(0)
            elif kind == 'jx':
                cache[kind] = syn.make_J_regressor_extra(seed=0)
        return cache[kind]
    return get
