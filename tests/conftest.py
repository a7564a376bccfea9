import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One product context on cuda:0.  No fallback: if the library or the GPU is missing the test errors."""
    import slam3d_gx_b200 as s3d
    c = s3d.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def small_cam():
    from slam3d_gx_b200 import synth
    return synth.Camera().scaled(0.25)   # 160x120 = 19 200 points


@pytest.fixture(scope="session")
def small_pair(small_cam):
    from slam3d_gx_b200 import synth
    return synth.make_pair(0, cam=small_cam)


@pytest.fixture(scope="session")
def full_pair():
    from slam3d_gx_b200 import synth
    return synth.make_pair(0)


def pose_close(Ta, Tb, rot_tol=1e-4, trans_tol=1e-4):
    from slam3d_gx_b200 import synth
    r, t = synth.pose_error(np.asarray(Ta), np.asarray(Tb))
    return r <= rot_tol and t <= trans_tol, (r, t)
