"""debug driver: one-pass modes vs the oracle, prints errors instead of asserting"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle
import uapic_b200 as ub
from conftest import seeded_load, periodic_diff
DT = np.pi / 16
DIMX, DIMY = 4 * np.pi, 2 * np.pi
corc = oracle.corc()
for (ntau, nx, ny, npart, nstep, eps) in [(32, 128, 128, 6000, 1, 0.1), (32, 128, 128, 6000, 4, 0.1), (16, 128, 64, 6001, 4, 0.1), (8, 64, 32, 5003, 4, 0.1), (16, 128, 64, 6000, 4, 1e-3)]:
    om, x0, v0 = seeded_load(npart, nx, ny, seed=3)
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    w = DIMX * DIMY / npart
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    eno, _, _, _ = corc.run_bupdate(om, ntau, eps, DT, nstep, xo, vo, w)
    for name, mode in (("full128", ub.STORE_FULL), ("onepass72", ub.STORE_ONEPASS), ("onepass48", ub.STORE_ONEPASS_LEAN)):
        try:
            xg, vg, eng, _ = ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w, storage_mode=mode)
        except Exception as e:
            print(ntau, nx, ny, npart, nstep, eps, name, "FAILED", e); continue
        dx = periodic_diff(xg[0], xo[0], DIMX).max() / DIMX
        dy = periodic_diff(xg[1], xo[1], DIMY).max() / DIMY
        dv = np.abs(vg - vo).max() / np.abs(vo).max()
        de = np.abs(eng - eno).max() / np.abs(eno).max()
        print(f"ntau={ntau} {nx}x{ny} np={npart} steps={nstep} eps={eps} {name}: dx={dx:.2e} dy={dy:.2e} dv={dv:.2e} de={de:.2e}", flush=True)
