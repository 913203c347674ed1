import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fandisk():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "fandisk_denoise.npz")))


@pytest.fixture(scope="session")
def until_min():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "until_min.npz")))


@pytest.fixture(scope="session")
def cube():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "cube386_labels.npz")))


@pytest.fixture(scope="session")
def cpsd():
    """Yadav-2018 baseline path on fandisk, recorded from the reference (tests/golden/make_golden_cpsd.py)"""
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "cpsd_fandisk.npz")))


@pytest.fixture(scope="session")
def fandisk_k32():
    """Processor.getMyFeatureDecomposition(32) + class steps on fandisk, recorded from the reference (make_golden_k32.py)"""
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "fandisk_k32.npz")))


@pytest.fixture(scope="session")
def ours():
    """PostProcessing.ipynb#c9 rows "Ours" (snapshot classes + displacement clamp) and "CTD-QEM" on fandisk, recorded from
    the reference (tests/golden/make_golden_ours.py)"""
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "ours_fandisk.npz")))


def angle_between(a, b):
    """fp64 atan2(|a x b|, a.b): fp32 acos has a ~5e-4 rad floor near 0."""
    import numpy as np
    a = a.astype(np.float64); b = b.astype(np.float64)
    return np.arctan2(np.linalg.norm(np.cross(a, b), axis=1), (a * b).sum(1))
