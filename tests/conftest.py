import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def clouds():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "clouds.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def ctx():
    import pcl_feature_extraction_b200 as pfx
    c = pfx.Context(0)
    yield c
    c.close()
