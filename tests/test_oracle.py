"""CPU tests that pin the oracle: the reference's own known-answer tests, the numpy twin
(Julia conventions) against the C restatement (Fortran conventions), and the committed golden vectors."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle import nporc

from conftest import golden_files, GOLDEN, periodic_diff, seeded_load


# ---- test/test_poisson.jl:1-49 -----------------------------------------------------------------
def test_poisson_analytic_reference_test(corc):
    m = oracle.mesh(0, 2 * np.pi / 0.5, 64, 0, 2 * np.pi / 1.0, 128)
    x = np.linspace(m.xmin, m.xmax, m.nx + 1)
    y = np.linspace(m.ymin, m.ymax, m.ny + 1)
    rho = np.asfortranarray(-8 * np.sin(2 * x)[:, None] * np.cos(2 * y)[None, :])
    e = np.zeros((2, m.nx + 1, m.ny + 1), order="F")
    sol = np.zeros_like(e)
    sol[0] = 2 * np.cos(2 * x)[:, None] * np.cos(2 * y)[None, :]
    sol[1] = -2 * np.sin(2 * x)[:, None] * np.sin(2 * y)[None, :]
    corc.poisson(m, rho, e)
    assert np.abs(e - sol).max() < 1e-14          # test_poisson.jl:28
    rho = np.asfortranarray(-4 * (np.sin(2 * x)[:, None] + np.cos(2 * y)[None, :]))
    corc.poisson(m, rho, e)
    sol[0] = (2 * np.cos(2 * x))[:, None] + 0 * y[None, :]
    sol[1] = 0 * x[:, None] - (2 * np.sin(2 * y))[None, :]
    assert np.abs(e - sol).max() < 1e-14          # test_poisson.jl:47


# ---- test/test_particles.jl:16-76 --------------------------------------------------------------
@pytest.mark.parametrize("wrap", [oracle.WRAP_FORTRAN, oracle.WRAP_JULIA])
def test_particles_meshfields_interaction_reference_test(corc, wrap):
    m = oracle.mesh(0.0, 20.0, 20, 0.0, 20.0, 20)
    nx = ny = 20
    dx = dy = 1.0
    pts = [((i - 0.5) * dx, (j - 0.5) * dx) for i in range(5, nx - 5 + 1) for j in range(5, ny - 5 + 1)]
    assert len(pts) == 121
    x = np.asfortranarray(np.array(pts).T.copy())
    rho = np.zeros((nx + 1, ny + 1), order="F")
    corc.compute_rho_m6(m, x, 1.0 / 121, rho, wrap)
    assert abs(rho[:nx, :ny].sum() * dx * dy) < 1e-4          # test_particles.jl:45
    e = np.zeros((2, nx + 1, ny + 1), order="F")
    for i in range(nx + 1):
        for j in range(ny + 1):
            e[0, i, j] = i * dx
            e[1, i, j] = j * dy
    ep = np.zeros((2, 121), order="F")
    corc.interpol_eb_m6(m, e, x, ep, wrap)
    assert np.abs(ep[0] - x[0]).mean() < 1e-6                 # test_particles.jl:73
    assert np.abs(ep[1] - x[1]).mean() < 1e-6                 # test_particles.jl:74


def test_f_m6_partition_of_unity(corc):
    for d in np.linspace(0, 0.999, 37):
        s = sum(corc.f_m6(abs(a - d)) for a in range(-3, 4))
        assert abs(s - 1.0) < 1e-15
    assert corc.f_m6(3.0) == 0.0 and corc.f_m6(3.5) == 0.0     # cm3 taps are identically zero
    assert np.allclose([corc.f_m6(q) for q in (0.2, 1.3, 2.6)], nporc.f_m6(np.array([0.2, 1.3, 2.6])), rtol=0, atol=1e-15)


def test_fft_matches_numpy(corc):
    rng = np.random.default_rng(3)
    for n in (4, 8, 16, 32, 12):
        a = np.asfortranarray(rng.standard_normal((n, 5)) + 1j * rng.standard_normal((n, 5)))
        assert np.abs(corc.fft_tau(a, -1) - np.fft.fft(a, axis=0)).max() < 1e-13
        assert np.abs(corc.fft_tau(a, +1, normalise=True) - np.fft.ifft(a, axis=0)).max() < 1e-14


def test_poisson_c_vs_numpy_on_noise(corc):
    """white-noise rho exercises FFTW's c2r handling of the non-Hermitian Nyquist / DC bins"""
    rng = np.random.default_rng(0)
    for nx, ny in ((128, 64), (64, 128), (32, 32), (20, 12)):
        m = oracle.mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
        mm = nporc.Mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
        rho = np.asfortranarray(rng.standard_normal((nx + 1, ny + 1)))
        e1 = np.zeros((2, nx + 1, ny + 1), order="F")
        e2 = np.zeros((2, nx + 1, ny + 1))
        n1 = corc.poisson(m, rho, e1)
        n2 = nporc.Poisson(mm)(np.array(rho), e2)
        assert np.abs(e1 - e2).max() < 1e-13
        assert abs(n1 - n2) / n2 < 1e-13


def test_stagewise_c_vs_numpy(corc):
    m, x, v = seeded_load(500)
    mm = nporc.Mesh(m.xmin, m.xmax, m.nx, m.ymin, m.ymax, m.ny)
    rng = np.random.default_rng(5)
    ep = np.asfortranarray(rng.standard_normal((2, 500)))
    ntau, eps, dt = 16, 0.1, np.pi / 16
    b, t, pl, ql, xt, yt = corc.preparation(ntau, eps, dt, x, v, ep)
    b2, t2, pl2, ql2, xt2, yt2 = nporc.preparation(ntau, eps, dt, x, v, ep)
    for a, c in ((b, b2), (t, t2), (pl, pl2), (ql, ql2), (xt, xt2), (yt, yt2)):
        assert np.abs(a - c).max() < 1e-13
    assert np.abs(yt.imag).max() > 1e-9        # imaginary parts are load-bearing (Nyquist filter)
    emesh = np.asfortranarray(rng.standard_normal((2, m.nx + 1, m.ny + 1)))
    emesh[:, m.nx, :] = emesh[:, 0, :]
    emesh[:, :, m.ny] = emesh[:, :, 0]
    et = np.zeros((ntau, 2, 500), order="F")
    et2 = np.zeros((ntau, 2, 500))
    corc.interpol_eb_m6_tau(m, emesh, xt, et)
    nporc.interpol_eb_m6_tau(mm, np.array(emesh), xt2, et2)
    assert np.abs(et - et2).max() < 1e-13
    fx, fy = corc.compute_f(eps, b, xt, yt, et, normalise=False)
    fx2, fy2 = nporc.compute_f(eps, b2, xt2, yt2, et2)
    assert np.abs(fx - fx2).max() < 1e-12 and np.abs(fy - fy2).max() < 1e-12
    rho = np.zeros((m.nx + 1, m.ny + 1), order="F")
    rho2 = np.zeros((m.nx + 1, m.ny + 1))
    xo = np.zeros((2, 500), order="F")
    xo2 = np.zeros((2, 500))
    w = 8 * np.pi ** 2 / 500
    corc.compute_rho_m6_tau(m, eps, xt, t, w, rho, xo, oracle.WRAP_JULIA)
    nporc.compute_rho_m6_tau(mm, rho2, xo2, w, xt2, t2, eps)
    assert np.abs(rho - rho2).max() < 1e-12 and np.abs(xo - xo2).max() < 1e-13
    assert abs(rho[:m.nx, :m.ny].sum()) * m.dx * m.dy < 1e-12       # neutralised


@pytest.mark.parametrize("eps,tolv", [(0.1, 1e-11), (1e-3, 1e-9)])
def test_full_run_c_vs_numpy(corc, eps, tolv):
    m, x0, v0 = seeded_load(1500, seed=77)
    mm = nporc.Mesh(m.xmin, m.xmax, m.nx, m.ymin, m.ymax, m.ny)
    w = 8 * np.pi ** 2 / 1500
    dt = np.pi / 16
    xa, va = x0.copy(order="F"), v0.copy(order="F")
    en, sv, _, _ = corc.run_bupdate(m, 16, eps, dt, 4, xa, va, w)
    xb, vb, en2, sv2, _ = nporc.run_bupdate(mm, 16, eps, dt, 4, x0, v0, w)
    assert periodic_diff(xa[0], xb[0], 4 * np.pi).max() < 1e-12
    assert periodic_diff(xa[1], xb[1], 2 * np.pi).max() < 1e-12
    assert np.abs(va - vb).max() < tolv
    assert np.abs(en - en2).max() / np.abs(en2).max() < 1e-13


def test_threads_do_not_change_results_beyond_roundoff(corc):
    m, x0, v0 = seeded_load(1200, seed=5)
    w = 8 * np.pi ** 2 / 1200
    outs = []
    for nt in (1, 4):
        corc.set_threads(nt)
        xa, va = x0.copy(order="F"), v0.copy(order="F")
        en, _, _, _ = corc.run_bupdate(m, 16, 0.1, np.pi / 16, 2, xa, va, w)
        outs.append((xa, va, en))
    corc.set_threads(1)
    assert np.abs(outs[0][0] - outs[1][0]).max() < 1e-12
    assert np.abs(outs[0][1] - outs[1][1]).max() < 1e-12
    assert np.abs(outs[0][2] - outs[1][2]).max() / outs[0][2].max() < 1e-13


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_c_oracle_against_golden(corc, path):
    g = np.load(path)
    nx, ny, ntau, nstep = int(g["nx"]), int(g["ny"]), int(g["ntau"]), int(g["nstep"])
    eps, dt, w = float(g["eps"]), float(g["dt"]), float(g["w"])
    m = oracle.mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)
    x, v = np.asfortranarray(g["x0"]).copy(order="F"), np.asfortranarray(g["v0"]).copy(order="F")
    en, sv, _, emesh = corc.run_bupdate(m, ntau, eps, dt, nstep, x, v, w)
    tolv = 1e-10 * max(1.0, 0.1 / eps)
    assert periodic_diff(x[0], g["x"][0], 4 * np.pi).max() < 1e-11
    assert periodic_diff(x[1], g["x"][1], 2 * np.pi).max() < 1e-11
    assert np.abs(v - g["v"]).max() < tolv
    assert np.abs(en - g["energy"]).max() / np.abs(g["energy"]).max() < 1e-12
    assert np.abs(emesh - g["emesh"]).max() < 1e-11


def test_cic_variant_c_vs_numpy_and_properties():
    """the build-defined CIC scheme (bilinear weights of performance/test_cic.F90:73-76 inside the UA loop; the reference has no
    such path): the two oracles state it independently and must agree; bilinear interpolation reproduces a bilinear field
    exactly away from the periodic seam; the deposit is neutral after the epilogue"""
    import oracle
    from oracle import uapic_oracle_np as onp
    c = oracle.corc()
    dimx, dimy = 4 * np.pi, 2 * np.pi
    om = oracle.mesh(0, dimx, 64, 0, dimy, 32)
    m = onp.Mesh(0, dimx, 64, 0, dimy, 32)
    rng = np.random.default_rng(5)
    npart = 2000
    x0, v0, _ = c.plasma_from_uniforms(om, npart, 0.05, 0.5, rng.random(npart * 80))
    w = dimx * dimy / npart
    dt = np.pi / 16
    try:
        c.set_scheme("cic"); onp.SCHEME = "cic"
        assert c.scheme() == 1
        xo, vo = x0.copy(order="F"), v0.copy(order="F")
        en, _, _, _ = c.run_bupdate(om, 16, 0.1, dt, 4, xo, vo, w)
        xn, vn, enn, _, _ = onp.run_bupdate(m, 16, 0.1, dt, 4, x0, v0, w)
        assert np.abs(np.mod(xo[0] - xn[0] + dimx / 2, dimx) - dimx / 2).max() < 1e-12 * dimx
        assert np.abs(vo - vn).max() < 1e-11 * np.abs(vo).max()
        assert np.abs(en - enn).max() < 1e-13 * np.abs(en).max()
        # bilinear field e1 = x*y, e2 = x - 2y on the nodes: reproduced exactly in the interior
        e = np.zeros((2, 65, 33), order="F")
        xs, ys = np.arange(65) * m.dx, np.arange(33) * m.dy
        e[0], e[1] = np.outer(xs, ys), xs[:, None] - 2 * ys[None, :]
        xp = np.asfortranarray(np.stack([rng.random(500) * (dimx - 2 * m.dx) + 0.5 * m.dx, rng.random(500) * (dimy - 2 * m.dy) + 0.5 * m.dy]))
        ep = np.zeros_like(xp)
        onp.interpol_eb_m6(m, e, xp.copy(), ep)
        assert np.abs(ep[0] - xp[0] * xp[1]).max() < 1e-12 and np.abs(ep[1] - (xp[0] - 2 * xp[1])).max() < 1e-12
        rho = np.zeros((65, 33))
        onp.compute_rho_m6(m, rho, x0.copy(), w)
        assert abs(rho[:64, :32].sum() * m.dx * m.dy) < 1e-10
    finally:
        c.set_scheme("m6"); onp.SCHEME = "m6"
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    en6, _, _, _ = c.run_bupdate(om, 16, 0.1, dt, 4, xo, vo, w)
    assert np.abs(en6 - en).max() > 1e-6 * np.abs(en6).max()      # M6 and CIC are different schemes


def test_gnuplot_dump_layout(tmp_path):
    """src/gnuplot.jl:4-27: (nx+1)(ny+1) records `x  y  e1  e2  rho`, x outer, y inner, a blank line after every x column"""
    import uapic_b200 as ub
    m = ub.Mesh(0, 4 * np.pi, 4, 0, 2 * np.pi, 2)
    f = ub.MeshFields(m)
    f.e[0], f.e[1] = 1.5, -2.0
    f.rho[:] = np.arange(15).reshape(5, 3)
    p = tmp_path / "fields.dat"
    ub.gnuplot(str(p), f)
    lines = p.read_text().split("\n")
    assert len(lines) == 5 * 3 + 5 + 1 and lines[3] == "" and lines[-1] == ""
    rec = [float(t) for t in lines[4].split()]            # i = 1, j = 0
    assert rec == [m.dx, 0.0, 1.5, -2.0, 3.0]
    assert lines[0].count("  ") == 4


# ---- both double oracles against the extended-precision referee (tests/golden/make_referee.py) -----------------------------
import glob as _glob  # noqa: E402

from referee_util import dist_to_referee, referee_cases  # noqa: E402


@pytest.mark.parametrize("path", referee_cases(), ids=lambda p: p.split("referee_")[-1][:-4])
def test_c_oracle_vs_extended_precision_referee(corc, path):
    """the C oracle (Fortran operation order, double) must sit where make_referee.py found it: x and the energy history within
    1e-13 of the long-double evaluation, v within 4.2e-14 * (0.1/eps) * 10 -- the eps-amplified rounding of b(x)"""
    g = np.load(path)
    om = oracle.mesh(0, 4 * np.pi, int(g["nx"]), 0, 2 * np.pi, int(g["ny"]))
    x, v = g["x0"].copy(order="F"), g["v0"].copy(order="F")
    en, _, _, _ = corc.run_bupdate(om, int(g["ntau"]), float(g["eps"]), float(g["dt"]), int(g["nstep"]), x, v, float(g["w"]))
    dx, dv, de = dist_to_referee(g, x, v, en)
    assert dx < 1e-13 and de < 1e-13
    assert dv < 5e-13 * (0.1 / float(g["eps"]))
    assert dv < 2.0 * float(g["c_oracle_dist"][1]) + 1e-15      # and has not moved since the vectors were made
