import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session; fails loudly without libicpb200.so or a GPU."""
    import icpb200
    c = icpb200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def pair10k(orc):
    """BASELINE config 1: Kinect v1 frame pair, 5 deg / 5 cm, 10k points each, clouds at world (5,5,5)."""
    import numpy as np
    from icpb200 import synth
    d0, d1, col, _ = synth.frame_pair()
    p0, _, _ = orc.backproject(d0, col)
    p1, _, _ = orc.backproject(d1, col)
    cam = np.array([5, 5, 5], np.float32)   # icp.cpp:53
    p0 = orc.translate(p0, cam)
    p1 = orc.translate(p1, cam)
    data = synth.subsample_exact(p1, 10000, 1)
    target = synth.subsample_exact(p0, 10000, 2)
    return data, target
