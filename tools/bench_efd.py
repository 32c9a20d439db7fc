#!/usr/bin/env python
"""Throughput of the external-field program (fortran/efd.f90: ntau = 16, eps = 1e-3, dt = pi/16, 8 IMEX2 steps) on one B200:
the kernel alone (CUDA events, device-resident particles, `uapic_efd_run_device`) and end to end through host buffers
(`uapic_efd_run`), beside the C restatement on all host cores and on one core.  Prints one JSON line.
    python tools/bench_efd.py [--particles N] [--cpu-particles M]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import uapic_b200 as ub  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=204800)        # efd.f90:19
ap.add_argument("--cpu-particles", type=int, default=204800)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
ntau, nstep = 16, 8
rng = np.random.default_rng(1)
n = a.particles
x = np.asfortranarray(rng.random((2, n)) * [[4 * np.pi], [2 * np.pi]])
v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
vd = torch.from_numpy(np.ascontiguousarray(v.T)).cuda()
xo, vo = torch.empty_like(xd), torch.empty_like(vd)
for _ in range(3):
    ub.efd_run_device(xd, vd, xo, vo)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ub.efd_run_device(xd, vd, xo, vo)
e1.record()
torch.cuda.synchronize()
ms_kernel = e0.elapsed_time(e1) / a.reps
ub.efd_run(x, v)                                  # first call of this entry point: context work, not the steady state
e2e = []
for _ in range(3):
    t0 = time.perf_counter()
    ub.efd_run(x, v)
    e2e.append(time.perf_counter() - t0)
s_e2e = min(e2e)
m = min(a.cpu_particles, n)
c = oracle.corc()
cores = c.max_threads()
c.set_threads(cores)
t0 = time.perf_counter()
c.efd_run(x[:, :m], v[:, :m])
s_cpu_all = time.perf_counter() - t0
c.set_threads(1)
m1 = max(1, m // 16)
t0 = time.perf_counter()
c.efd_run(x[:, :m1], v[:, :m1])
s_cpu_one = time.perf_counter() - t0
unit = "particle-tau IMEX2 steps/s"
print(json.dumps({
    "workload": f"fortran/efd.f90 over all particles: {n} particles, ntau = {ntau}, eps = 1e-3, {nstep} steps after the third-order preparation",
    "unit": unit,
    "gpu_kernel": {"ms": ms_kernel, "value": n * ntau * nstep / (ms_kernel * 1e-3)},
    "gpu_e2e_host_buffers": {"ms": 1e3 * s_e2e, "all_ms": [1e3 * t for t in e2e], "value": n * ntau * nstep / s_e2e, "h2d_bytes": 32 * n, "d2h_bytes": 32 * n},
    "cpu_oracle_all_cores": {"cores": cores, "particles": m, "seconds": s_cpu_all, "value": m * ntau * nstep / s_cpu_all},
    "cpu_oracle_one_core": {"particles": m1, "seconds": s_cpu_one, "value": m1 * ntau * nstep / s_cpu_one},
    "speedup_kernel_vs_all_cores": (s_cpu_all / m) / (ms_kernel * 1e-3 / n),
}))
