"""CPU checks of the oracle's restatement of the reference's 3D path (fortran/uapic3d.f90 and its modules) against the
reference-owned programs that exercise it: fortran/test_poisson_3d.f90 (analytic Poisson solution) and fortran/test_pic_3d.f90
(total charge, CIC interpolation of a smooth field).  These pin orc3_poisson / orc3_compute_rho_cic / orc3_interpolate_eb_cic."""
import numpy as np

import oracle


def test_poisson_3d_reference_program():
    """test_poisson_3d.f90:14-68: rho = -3 sin x sin y sin z on [0,2pi]x[0,4pi]x[0,6pi], 32 x 64 x 128 -> E = grad(sin sin sin).
    The program prints the summed absolute errors; here they must vanish to round-off at every node, ghost planes included."""
    nx, ny, nz = 32, 64, 128
    m = oracle.mesh3((0, 0, 0), (2 * np.pi, 4 * np.pi, 6 * np.pi), (nx, ny, nz))
    x = np.arange(nx + 1) * (2 * np.pi / nx)
    y = np.arange(ny + 1) * (4 * np.pi / ny)
    z = np.arange(nz + 1) * (6 * np.pi / nz)
    sx, sy, sz = np.sin(x)[:, None, None], np.sin(y)[None, :, None], np.sin(z)[None, None, :]
    cx, cy, cz = np.cos(x)[:, None, None], np.cos(y)[None, :, None], np.cos(z)[None, None, :]
    rho = np.asfortranarray(-3 * sx * sy * sz)
    e = oracle.corc3().poisson(m, rho)
    assert np.abs(e[0] - cx * sy * sz).max() < 1e-14
    assert np.abs(e[1] - sx * cy * sz).max() < 1e-14
    assert np.abs(e[2] - sx * sy * cz).max() < 1e-14


def test_pic_3d_reference_program():
    """test_pic_3d.f90:45-84: total deposited charge vs nbpart*w, and CIC interpolation of (sin x, sin y, sin z)"""
    nx, ny, nz, npart = 64, 64, 4, 64 * 64 * 10
    m = oracle.mesh3((0, 0, 0), (18, 18, 1), (nx, ny, nz))
    o = oracle.corc3()
    x, v = o.generate(m, 20190101, npart)
    assert x.min() >= 0 and x[0].max() < 18 and x[1].max() < 18 and x[2].max() < 1
    w = 18 * 18 * 1 / npart
    rho = o.compute_rho_cic(m, x, w)
    dx, dy, dz = 18 / nx, 18 / ny, 1 / nz
    # the ghost planes are periodic images (the reference OVERWRITES what was deposited there, compute_rho_cic.f90:69-71)
    assert np.array_equal(rho[nx], rho[0]) and np.array_equal(rho[:, ny], rho[:, 0]) and np.array_equal(rho[:, :, nz], rho[:, :, 0])
    total = rho[:nx, :ny, :nz].sum() * dx * dy * dz
    assert 0.70 * npart * w < total <= npart * w * (1 + 1e-12)       # charge that fell on the ghost planes is lost in the reference
    X = np.arange(nx + 1) * dx
    Y = np.arange(ny + 1) * dy
    Z = np.arange(nz + 1) * dz
    e = np.zeros((3, nx + 1, ny + 1, nz + 1), order="F")
    e[0] = np.sin(X)[:, None, None]
    e[1] = np.sin(Y)[None, :, None]
    e[2] = np.sin(Z)[None, None, :]
    ep = o.interpolate_eb_cic(m, e, x)
    assert np.abs(ep[0] - np.sin(x[0])).mean() < dx * dx and np.abs(ep[1] - np.sin(x[1])).mean() < dy * dy
    assert np.abs(ep[2] - np.sin(x[2])).mean() < dz * dz


def test_both_branches_run_and_quirk_matters():
    m = oracle.mesh3((0, 0, 0), (18, 18, 1), (16, 16, 4))
    o = oracle.corc3()
    x0, v0 = o.generate(m, 7, 500)
    w = 18 * 18 / 500
    for nmrc, tfinal, expect in ((8, 0.05, 65), (4, np.pi, 2 * 4 * 4)):          # N0mrc = 1 -> plain steps ; N0mrc = 128 -> MRC
        x, v = x0.copy(order="F"), v0.copy(order="F")
        n, ep, e, rho = o.run(m, x, v, w, 0.5 ** 10, 3e-3, nmrc, 4, tfinal)
        assert n == expect and np.all(np.isfinite(x)) and np.all(np.isfinite(v)) and np.all(np.isfinite(e))
    xa, va = x0.copy(order="F"), v0.copy(order="F")
    xb, vb = x0.copy(order="F"), v0.copy(order="F")
    o.run(m, xa, va, w, 0.5 ** 10, 0.03, 4, 4, np.pi, index_quirk=1)
    o.run(m, xb, vb, w, 0.5 ** 10, 0.03, 4, 4, np.pi, index_quirk=0)
    assert np.abs(va - vb).max() > 1e-8          # p%x(m,1) instead of p%x(1,m) changes the answer (uapic3d.f90:179,182)
