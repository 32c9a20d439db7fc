#!/usr/bin/env python
"""Measured parity at small eps: for every extended-precision referee case (tests/golden/referee_*.npz)
run the C oracle, the one-pass kernels and the two-barrier kernels on the same double inputs and record the max relative
distance of x, v and the energy history to the referee, and of the GPU to the oracle.  Needs a GPU.
    python tools/small_eps_table.py > profiles/r2_small_eps_parity.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import uapic_b200 as ub  # noqa: E402
from referee_util import DIMX, DIMY, dist_to_referee, referee_cases  # noqa: E402


def rel(a, b, scale):
    return float(np.abs(a - b).max() / scale)


rows = []
for path in referee_cases():
    g = np.load(path)
    nx, ny, ntau, nstep = int(g["nx"]), int(g["ny"]), int(g["ntau"]), int(g["nstep"])
    eps, dt, w = float(g["eps"]), float(g["dt"]), float(g["w"])
    om = oracle.mesh(0, DIMX, nx, 0, DIMY, ny)
    xo, vo = g["x0"].copy(order="F"), g["v0"].copy(order="F")
    eno, _, _, _ = oracle.corc().run_bupdate(om, ntau, eps, dt, nstep, xo, vo, w)
    row = {"case": os.path.basename(path)[:-4], "eps": eps, "ntau": ntau, "mesh": [nx, ny], "particles": int(xo.shape[1]), "steps": nstep,
           "c_oracle_vs_referee": dict(zip("xvE", dist_to_referee(g, xo, vo, eno))),
           "numpy_twin_vs_referee": dict(zip("xvE", map(float, g["np_twin_dist"])))}
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    for name, mode in (("onepass_lean", ub.STORE_ONEPASS_LEAN), ("onepass_full", ub.STORE_ONEPASS), ("two_barrier", ub.STORE_FULL)):
        x, v, en, _ = ub.run_bupdate(mesh, ntau, eps, dt, nstep, g["x0"], g["v0"], w, storage_mode=mode)
        row[f"gpu_{name}_vs_referee"] = dict(zip("xvE", dist_to_referee(g, x, v, en)))
        dxo = max(np.abs(np.mod(x[0] - xo[0] + DIMX / 2, DIMX) - DIMX / 2).max() / DIMX, np.abs(np.mod(x[1] - xo[1] + DIMY / 2, DIMY) - DIMY / 2).max() / DIMY)
        row[f"gpu_{name}_vs_c_oracle"] = {"x": float(dxo), "v": rel(v, vo, np.abs(vo).max()), "E": rel(en, eno, np.abs(eno).max())}
    rows.append(row)
    print(f"{row['case']}: v: oracle->ref {row['c_oracle_vs_referee']['v']:.1e}  onepass->ref {row['gpu_onepass_lean_vs_referee']['v']:.1e}  "
          f"two-barrier->ref {row['gpu_two_barrier_vs_referee']['v']:.1e}  onepass->oracle {row['gpu_onepass_lean_vs_c_oracle']['v']:.1e}", file=sys.stderr)
json.dump({"what": "max relative distances after 8 UA steps; referee = numpy twin in x87 long double (tests/golden/make_referee.py)", "rows": rows},
          sys.stdout, indent=1)
