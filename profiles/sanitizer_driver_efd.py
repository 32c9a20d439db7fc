"""compute-sanitizer target for the external-field kernel: both tau policies (lane-per-sample, warp-per-particle), tails that
leave idle particle groups in the last warp, the in-place device entry point, and the CIC stage entry points."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uapic_b200 as ub  # noqa: E402

rng = np.random.default_rng(0)
cases = ((16, 1003), (32, 77), (2, 65), (8, 130), (12, 37), (50, 9), (250, 5))
if os.environ.get("EFD_SANITIZER_LANES_ONLY"):          # the power-of-two lane policies only (short run)
    cases = ((16, 203), (32, 77), (4, 33), (8, 130))
for ntau, n in cases:
    x = np.asfortranarray(rng.random((2, n)) * [[4 * np.pi], [2 * np.pi]])
    v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
    xo, vo = ub.efd_run(x, v, ntau=ntau, nstep=2)
    assert np.isfinite(xo).all() and np.isfinite(vo).all()
mesh = ub.Mesh(0, 4 * np.pi, 20, 0, 2 * np.pi, 12)
p = ub.Particles(999, 1.0)
p.x[0], p.x[1] = rng.uniform(-9, 9, 999), rng.uniform(-9, 9, 999)
f = ub.MeshFields(mesh)
ub.compute_rho_cic(f, p)
ub.interpol_eb_cic(p, f)
print("sanitizer driver (efd + cic stages) done")
