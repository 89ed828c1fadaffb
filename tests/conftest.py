import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib  # noqa: E402  (also puts the product's python/ dir on sys.path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def synth():
    return oracle_lib.Synth()


@pytest.fixture(scope="session")
def engine():
    import slam_b200
    e = slam_b200.Engine(0)  # raises loudly without the CUDA library or a GPU: there is no CPU fallback
    yield e
    e.close()


@pytest.fixture(scope="session")
def scene(synth):
    return synth.scene(1, n_boxes=400)


@pytest.fixture(scope="session")
def small_pair(synth, scene, oracle):
    """Two reduced-resolution scans 1 m apart, voxel 0.5 (a fast stand-in for config C1)."""
    s = oracle_lib.small_sensor(32, 600)
    a = synth.scan(s, scene, (0.0, 0.0, 0.0), 7)
    b = synth.scan(s, scene, (1.0, 0.1, 0.01), 8)
    da, _ = oracle.voxel_downsample(a, 0.5)
    db, _ = oracle.voxel_downsample(b, 0.5)
    return dict(raw_a=a, raw_b=b, a=da, b=db)


def rot_angle(R):
    return float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))
