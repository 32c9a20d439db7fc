"""Extended-precision referee vectors under tests/golden/referee_*.npz.

The reference's own programs cannot run here (no Julia, no Fortran compiler, no FFTW), so these are NOT reference
outputs.  They are the numpy twin (oracle/uapic_oracle_np.py, a restatement of src/*.jl + test/bupdate.jl) evaluated in
x87 LONG DOUBLE (64-bit mantissa, eps 1.1e-19) from double-precision inputs: the same formulas with 2000x smaller
rounding.  Two double implementations that disagree (C oracle vs CUDA kernels, or the reference's own Fortran vs Julia)
can be ranked against it, and the parity tolerances of the GPU tests at small eps are set from the distances measured
here instead of from an amplification argument ("What's weak" item 2).

Every referee value is stored as a (hi, lo) pair of doubles, hi + lo = the long-double value, so that a double result
can be compared with it without the referee's own rounding to double getting in the way.

    python tests/golden/make_referee.py          # ~1 minute
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import corc, nporc  # noqa: E402

CASES = {
    # name: (nx, ny, ntau, eps, nstep, npart, seed)
    "referee_n16_eps1e-1": (64, 32, 16, 1e-1, 8, 400, 7101),
    "referee_n16_eps1e-2": (64, 32, 16, 1e-2, 8, 400, 7102),
    "referee_n16_eps1e-3": (64, 32, 16, 1e-3, 8, 400, 7103),
    "referee_n16_eps1e-4": (64, 32, 16, 1e-4, 8, 400, 7104),
    "referee_n16_eps1e-5": (64, 32, 16, 1e-5, 8, 400, 7105),
    "referee_n32_eps1e-1": (64, 64, 32, 1e-1, 8, 400, 7111),
    "referee_n32_eps1e-3": (64, 64, 32, 1e-3, 8, 400, 7113),
    "referee_n32_eps1e-5": (64, 64, 32, 1e-5, 8, 400, 7115),
    "referee_n8_eps1e-4": (32, 32, 8, 1e-4, 8, 400, 7124),
}


def split(a):
    hi = np.asarray(a, dtype=np.float64)
    lo = np.asarray(np.asarray(a, dtype=np.longdouble) - hi.astype(np.longdouble), dtype=np.float64)
    return hi, lo


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    dimx, dimy = 4 * np.pi, 2 * np.pi
    for name, (nx, ny, ntau, eps, nstep, npart, seed) in CASES.items():
        m = oracle.mesh(0, dimx, nx, 0, dimy, ny)
        rng = np.random.default_rng(seed)
        x0, v0, _ = corc().plasma_from_uniforms(m, npart, 0.05, 0.5, rng.random(npart * 80))
        w = dimx * dimy / npart
        dt = np.pi / 16
        with nporc.precision("longdouble"):
            mm = nporc.Mesh(0, dimx, nx, 0, dimy, ny)
            x, v, energy, sumv, emesh = nporc.run_bupdate(mm, ntau, eps, dt, nstep, x0, v0, w)
        assert x.dtype == np.longdouble and energy.dtype == np.longdouble
        # how far are the two double oracles from it?
        xc, vc = x0.copy(order="F"), v0.copy(order="F")
        en_c, _, _, _ = corc().run_bupdate(m, ntau, eps, dt, nstep, xc, vc, w)
        mm64 = nporc.Mesh(0, dimx, nx, 0, dimy, ny)
        xn, vn, en_n, _, _ = nporc.run_bupdate(mm64, ntau, eps, dt, nstep, x0, v0, w)
        L = np.longdouble

        def dist(xa, va, ea):
            ddx = np.abs(np.mod(xa[0].astype(L) - x[0] + dimx / 2, dimx) - dimx / 2).max() / dimx
            ddy = np.abs(np.mod(xa[1].astype(L) - x[1] + dimy / 2, dimy) - dimy / 2).max() / dimy
            return float(max(ddx, ddy)), float(np.abs(va.astype(L) - v).max() / np.abs(v).max()), float(np.abs(ea.astype(L) - energy).max() / np.abs(energy).max())
        print(f"{name}: C oracle vs referee x {dist(xc, vc, en_c)[0]:.1e} v {dist(xc, vc, en_c)[1]:.1e} E {dist(xc, vc, en_c)[2]:.1e} | "
              f"numpy twin vs referee x {dist(xn, vn, en_n)[0]:.1e} v {dist(xn, vn, en_n)[1]:.1e} E {dist(xn, vn, en_n)[2]:.1e}", flush=True)
        xh, xl = split(x); vh, vl = split(v); eh, el = split(energy)
        np.savez_compressed(os.path.join(out, name + ".npz"), nx=nx, ny=ny, ntau=ntau, eps=eps, nstep=nstep, dt=dt, w=w,
                            x0=x0, v0=v0, x_hi=xh, x_lo=xl, v_hi=vh, v_lo=vl, energy_hi=eh, energy_lo=el,
                            c_oracle_dist=np.array(dist(xc, vc, en_c)), np_twin_dist=np.array(dist(xn, vn, en_n)))


if __name__ == "__main__":
    main()
