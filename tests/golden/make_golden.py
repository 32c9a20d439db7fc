"""Generate the committed golden vectors under tests/golden/.

The reference (Julia / Fortran + FFTW) cannot execute in this image, so these vectors are NOT
outputs of the reference itself: they are outputs of the numpy twin (oracle/uapic_oracle_np.py,
a restatement of src/*.jl + test/bupdate.jl) on a seeded load, cross-checked here against the
C oracle (a restatement of fortran/*.F90).  They freeze today's agreed answer so that a later
change to either oracle, or to the CUDA kernels, shows up as a diff.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import corc, nporc  # noqa: E402


def load(seed, npart, m):
    rng = np.random.default_rng(seed)
    u = rng.random(npart * 80)
    x, v, _ = corc().plasma_from_uniforms(m, npart, 0.05, 0.5, u)
    return x, v


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    cases = {
        # name: (nx, ny, ntau, eps, nstep, npart, seed)
        "bupdate_n16_eps1e-1": (128, 64, 16, 0.1, 4, 3000, 20190101),
        "bupdate_n32_eps1e-1": (128, 128, 32, 0.1, 3, 2000, 20190102),
        "bupdate_n16_eps1e-3": (128, 64, 16, 1e-3, 4, 3000, 20190103),
        "bupdate_n8_mesh32x16": (32, 16, 8, 0.1, 3, 1500, 20190104),
    }
    for name, (nx, ny, ntau, eps, nstep, npart, seed) in cases.items():
        m = oracle.mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
        mm = nporc.Mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
        x0, v0 = load(seed, npart, m)
        w = (4 * np.pi * 2 * np.pi) / npart
        dt = np.pi / 2 / 8
        x, v, energy, sumv, emesh = nporc.run_bupdate(mm, ntau, eps, dt, nstep, x0, v0, w)
        xc, vc = x0.copy(order="F"), v0.copy(order="F")
        en_c, sv_c, _, em_c = corc().run_bupdate(m, ntau, eps, dt, nstep, xc, vc, w)
        dimx, dimy = 4 * np.pi, 2 * np.pi
        dx = np.abs(np.mod(xc[0] - x[0] + dimx / 2, dimx) - dimx / 2).max()
        dy = np.abs(np.mod(xc[1] - x[1] + dimy / 2, dimy) - dimy / 2).max()
        print(f"{name}: C-vs-numpy  dx={dx:.2e} dy={dy:.2e} dv={np.abs(vc - v).max():.2e} "
              f"dE={np.abs(en_c - energy).max() / np.abs(energy).max():.2e}")
        np.savez_compressed(os.path.join(out, name + ".npz"), nx=nx, ny=ny, ntau=ntau, eps=eps, nstep=nstep, dt=dt, w=w,
                            x0=x0, v0=v0, x=x, v=v, energy=energy, sumv=sumv, emesh=emesh)


if __name__ == "__main__":
    main()
