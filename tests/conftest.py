import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def lib_built():
    """The CUDA library must exist (built by __graft_entry__.build()); build it if missing."""
    from manifoldem_esper_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _capi.LIB_PATH


@pytest.fixture(scope='session')
def ctx(lib_built):
    """GPU context; GPU tests FAIL (not skip) if the extension cannot run on the device."""
    from manifoldem_esper_b200 import _capi
    c = _capi.Context(0)
    yield c
    c.close()


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLD, name))
