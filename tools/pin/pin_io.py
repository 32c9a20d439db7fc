"""binary exchange files of tools/pin_against_reference.sh (pin_driver.F90 / pin_julia.jl) and the comparison"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
DIMX, DIMY = 4 * np.pi, 2 * np.pi


def cases():
    """(name, nx, ny, ntau, eps, nstep, x0, v0): the referee cases plus BASELINE config 1 itself (204 800 particles, as shipped)"""
    from referee_util import referee_cases
    out = []
    for p in referee_cases():
        g = np.load(p)
        out.append((os.path.basename(p)[:-4].replace("referee", "pinned"), int(g["nx"]), int(g["ny"]), int(g["ntau"]), float(g["eps"]),
                    int(g["nstep"]), g["x0"], g["v0"]))
    import oracle
    m = oracle.mesh(0, DIMX, 128, 0, DIMY, 64)
    x0, v0 = oracle.corc().generate(m, "plasma", 20190101, 204800)
    out.append(("pinned_config1_as_shipped", 128, 64, 16, 0.1, 8, x0, v0))
    return out


def write_input(path, nx, ny, ntau, eps, nstep, x0, v0):
    n = x0.shape[1]
    with open(path, "wb") as f:
        np.array([n], dtype=np.int64).tofile(f)
        np.array([nx, ny, ntau, nstep], dtype=np.int32).tofile(f)
        np.array([eps, np.pi / 16, DIMX, DIMY, DIMX * DIMY / n], dtype=np.float64).tofile(f)
        np.asfortranarray(x0).T.tofile(f)          # (2,np) column-major = particle-major pairs
        np.asfortranarray(v0).T.tofile(f)


def read_output(path):
    with open(path, "rb") as f:
        n = int(np.fromfile(f, np.int64, 1)[0])
        nx, ny, ntau, nstep = map(int, np.fromfile(f, np.int32, 4))
        x = np.fromfile(f, np.float64, 2 * n).reshape(n, 2).T
        v = np.fromfile(f, np.float64, 2 * n).reshape(n, 2).T
        en = np.fromfile(f, np.float64, 1 + 2 * nstep)
        e = np.fromfile(f, np.float64, 2 * (nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1, 2).transpose(2, 1, 0)
    return x, v, en, e


def rel(xa, va, ea, xb, vb, eb):
    dx = max(np.abs(np.mod(xa[0] - xb[0] + DIMX / 2, DIMX) - DIMX / 2).max() / DIMX,
             np.abs(np.mod(xa[1] - xb[1] + DIMY / 2, DIMY) - DIMY / 2).max() / DIMY)
    return float(dx), float(np.abs(va - vb).max() / np.abs(vb).max()), float(np.abs(ea - eb).max() / np.abs(eb).max())


def main():
    cmd, work = sys.argv[1], sys.argv[2]
    if cmd == "inputs":
        for name, nx, ny, ntau, eps, nstep, x0, v0 in cases():
            write_input(os.path.join(work, name + ".in"), nx, ny, ntau, eps, nstep, x0, v0)
            print(name)
        return
    # compare: reference outputs (Fortran, and Julia if present) vs the C oracle, the numpy twin and -- if a GPU is there -- the library
    import oracle
    try:
        import uapic_b200 as ub
        have_gpu = ub.device_count() > 0
    except Exception:
        have_gpu = False
    worst = 0.0
    for name, nx, ny, ntau, eps, nstep, x0, v0 in cases():
        for impl in ("fortran", "julia"):
            outp = os.path.join(work, f"{name}.{impl}.out")
            if not os.path.exists(outp):
                continue
            xr, vr, enr, er = read_output(outp)
            om = oracle.mesh(0, DIMX, nx, 0, DIMY, ny)
            xo, vo = x0.copy(order="F"), v0.copy(order="F")
            wrap = oracle.WRAP_FORTRAN if impl == "fortran" else oracle.WRAP_JULIA
            eno, _, _, _ = oracle.corc().run_bupdate(om, ntau, eps, np.pi / 16, nstep, xo, vo, DIMX * DIMY / x0.shape[1], wrap=wrap)
            d = rel(xo, vo, eno, xr, vr, enr)
            line = f"{name:32s} {impl:8s} oracle-vs-reference x {d[0]:.1e} v {d[1]:.1e} E {d[2]:.1e}"
            tolv = 1e-10 * max(1.0, 1e-3 / eps)
            ok = d[0] < 1e-10 and d[2] < 1e-10 and d[1] < tolv
            if have_gpu:
                xg, vg, eng, _ = ub.run_bupdate(ub.Mesh(0, DIMX, nx, 0, DIMY, ny), ntau, eps, np.pi / 16, nstep, x0, v0, DIMX * DIMY / x0.shape[1],
                                                wrap=ub.WRAP_FORTRAN if impl == "fortran" else ub.WRAP_JULIA)
                dg = rel(xg, vg, eng, xr, vr, enr)
                line += f" | GPU-vs-reference x {dg[0]:.1e} v {dg[1]:.1e} E {dg[2]:.1e}"
                ok = ok and dg[0] < 1e-10 and dg[2] < 1e-10 and dg[1] < tolv
            print(line, "OK" if ok else "MISMATCH")
            worst = max(worst, 0.0 if ok else 1.0)
            if impl == "fortran":      # these ARE reference outputs: commit them as golden vectors
                np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), nx=nx, ny=ny, ntau=ntau, eps=eps, nstep=nstep,
                                    dt=np.pi / 16, w=DIMX * DIMY / x0.shape[1], x0=x0, v0=v0, x=xr, v=vr, energy=enr, emesh=er,
                                    source="fortran/bupdate.F90 modules via tools/pin/pin_driver.F90")
    sys.exit(int(worst))


if __name__ == "__main__":
    main()
