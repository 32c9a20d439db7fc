#!/usr/bin/env python
"""Throughput of the sibling 3D scheme (fortran/uapic3d.f90 as shipped: 64 x 64 x 4 cells, 409 600 particles, ep = 2^-10,
delta = 3e-3, Nmrc = Nmrcm = 128 -> N0mrc = 4, 32 768 sub-steps) on one B200, beside the CPU oracle (C restatement of the
Fortran, one core -- the Fortran is single-threaded) on a bounded number of outer iterations.  Prints one JSON line.
    python tools/bench_uapic3d.py [--outer K]      (K outer iterations of the MRC loop = 256 K sub-steps; 0 = the whole program)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import uapic_b200 as ub  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--outer", type=int, default=8)
ap.add_argument("--cpu-outer", type=int, default=1)
a = ap.parse_args()
nx, ny, nz = 64, 64, 4
npart = nx * ny * 100
mesh = ub.Mesh3D((0, 0, 0), (18, 18, 1), (nx, ny, nz))
om = oracle.mesh3((0, 0, 0), (18, 18, 1), (nx, ny, nz))
with ub.Session3D(mesh, npart) as s:
    s.generate_particles(seed=20190101)
    s.init_fields()
    s.run(128, 128, np.pi, 1)                       # warm-up: one outer iteration
    x, v, _ = s.download_particles()
    t0 = time.perf_counter()
    n = s.run(128, 128, np.pi, a.outer)
    s.download_particles()
    dt_gpu = time.perf_counter() - t0
    launches = s.launch_count
n_cpu, dt_cpu = 1, float("nan")
if a.cpu_outer > 0:          # 0 = skip the CPU leg (NOT "the whole program": that is half an hour on one core)
    xo, vo = oracle.corc3().generate(om, 20190101, npart)
    t0 = time.perf_counter()
    n_cpu, _, _, _ = oracle.corc3().run(om, xo, vo, 18 * 18 / npart, 0.5 ** 10, 3e-3, 128, 128, np.pi, max_outer=a.cpu_outer)
    dt_cpu = time.perf_counter() - t0
print(json.dumps({"workload": "fortran/uapic3d.f90 as shipped: 64x64x4, 409600 particles, Nmrc = Nmrcm = 128 (N0mrc = 4)",
                  "cuda_graph": os.environ.get("UAPIC3D_NO_GRAPH") != "1", "gpu": {"substeps": n, "seconds": dt_gpu, "us_per_substep": 1e6 * dt_gpu / n, "particle_substeps_per_s": npart * n / dt_gpu,
                          "whole_program_seconds_extrapolated": dt_gpu / n * 32768},
                  "cpu_oracle_1_core": {"substeps": n_cpu, "seconds": dt_cpu, "us_per_substep": 1e6 * dt_cpu / n_cpu,
                                        "particle_substeps_per_s": npart * n_cpu / dt_cpu, "whole_program_seconds_extrapolated": dt_cpu / n_cpu * 32768},
                  "speedup": (dt_cpu / n_cpu) / (dt_gpu / n)}))
