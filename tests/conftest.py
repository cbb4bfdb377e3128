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


def _gpu_unavailable_reason():
    """None when the CUDA path can run here; otherwise why not (no library / no B200 / no driver)."""
    try:
        from vkjax_b200 import runtime
        runtime.Context.get()
        return None
    except Exception as exc:                                   # noqa: BLE001 - any failure means "cannot run GPU tests here"
        return f'{type(exc).__name__}: {exc}'


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a B200 skips the gpu-marked tests instead of failing at the first one.
    `-m gpu` on the GPU box is unaffected: there the context is created and nothing is skipped -- and if the library
    is missing THERE the tests must fail loudly, so an explicit `-m gpu` selection never skips."""
    gpu_items = [it for it in items if it.get_closest_marker('gpu') is not None]
    if not gpu_items or 'gpu' in (config.getoption('-m') or '').replace('not gpu', ''):
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason='needs a B200 + libb2jax.so: ' + reason)
    for it in gpu_items:
        it.add_marker(skip)
