"""shared by tests/test_oracle.py, tests/test_gpu_referee.py and tools/small_eps_table.py: distances of a double-precision
result to the extended-precision referee vectors (tests/golden/referee_*.npz, made by tests/golden/make_referee.py)"""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIMX, DIMY = 4 * np.pi, 2 * np.pi


def referee_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "referee_*.npz")))


def dist_to_referee(g, x, v, energy):
    """(max |dx| / box, max |dv| / max|v|, max |dE| / max|E|) of a double result against referee hi + lo"""
    def per(a, hi, lo, period):
        d = (a - hi) - lo
        return np.abs(np.mod(d + period / 2, period) - period / 2)
    dx = max(per(x[0], g["x_hi"][0], g["x_lo"][0], DIMX).max() / DIMX, per(x[1], g["x_hi"][1], g["x_lo"][1], DIMY).max() / DIMY)
    dv = np.abs((v - g["v_hi"]) - g["v_lo"]).max() / np.abs(g["v_hi"]).max()
    de = np.abs((energy - g["energy_hi"]) - g["energy_lo"]).max() / np.abs(g["energy_hi"]).max()
    return float(dx), float(dv), float(de)


def v_tolerance(g):
    """velocity tolerance of a double implementation against the referee: ten times the distance of the reference's own
    arithmetic order in double (the C oracle) from the extended-precision answer, never tighter than 5e-13"""
    return max(5e-13, 10.0 * float(g["c_oracle_dist"][1]))
