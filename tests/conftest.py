import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def corc():
    import oracle
    return oracle.corc()


def seeded_load(npart, nx=128, ny=64, seed=1234):
    """config-1 densities (fortran/particles.F90:68-103) from a seeded numpy stream"""
    import oracle
    m = oracle.mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
    rng = np.random.default_rng(seed)
    u = rng.random(npart * 80 + 1000)
    x, v, _ = oracle.corc().plasma_from_uniforms(m, npart, 0.05, 0.5, u)
    return m, x, v


def periodic_diff(a, b, period):
    return np.abs(np.mod(a - b + period / 2, period) - period / 2)


def golden_files():
    """tests/golden/bupdate_*.npz (frozen agreement of the two oracles, make_golden.py) and, when somebody has run
    tools/pin_against_reference.sh on a machine with gfortran + FFTW, pinned_*.npz: outputs of the reference's own Fortran"""
    import glob
    return sorted(glob.glob(os.path.join(GOLDEN, "bupdate_*.npz")) + glob.glob(os.path.join(GOLDEN, "pinned_*.npz")))
