import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box)')


def has_cuda():
    try:
        from urban_pointcloud_processing_b200 import _lib
        return _lib.load().upcp_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def from_mm(a):
    return a.astype(np.float64) * 0.001


def unpack(bits, n):
    return np.unpackbits(bits)[:n].astype(bool)


def config1_cloud(g):
    """BASELINE config 1: the 4 M-point synthetic cloud over the REAL demo tile, regenerated from
    the seed stored in tests/golden/config1_golden.npz (which holds only the reference's labels)."""
    import ast
    import pandas as pd
    from urban_pointcloud_processing_b200 import synthetic
    from urban_pointcloud_processing_b200.utils import ahn_utils
    demo = os.path.join(GOLDEN, 'demo')
    ahn = ahn_utils.load_ahn_tile(os.path.join(demo, 'ahn_2386_9702.npz'))
    df = pd.read_csv(os.path.join(demo, 'bgt_buildings_demo.csv'))
    rings = [np.array(ast.literal_eval(p), dtype=np.float64) for p in df.polygon.values]
    points, kind = synthetic.make_demo_cloud(int(g['n']), ahn, rings, seed=int(g['seed']))
    assert int(np.round(points * 1000.0).astype(np.int64).sum()) == int(g['checksum_points'])
    return points, kind


CONFIG1_LAYERS = [{'bottom': 12., 'grid_size': 0.1, 'threshold': 0.5},
                  {'bottom': 0., 'top': 12., 'grid_size': 0.05, 'threshold': 0.5}]
