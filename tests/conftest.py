import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 and the built libb2jax.so (run with `-m gpu` on the GPU box)')


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(1234)
