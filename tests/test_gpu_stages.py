"""GPU parity tests, stage by stage: every exported stage function of libuapic_b200.so (through the Python mirror
of the Julia API, i.e. through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances: bit-exact where the kernel follows the reference's operation order without FMA contraction
(M6 gathers, fixed-point deposits); 1e-12 relative elsewhere (the tau FFT runs across lanes in a different,
equally valid, summation order than the oracle's).
"""
import numpy as np
import pytest

import oracle
import uapic_b200 as ub
from oracle import nporc

from conftest import seeded_load

pytestmark = pytest.mark.gpu

DT = np.pi / 16


def _mesh_pair(nx=128, ny=64):
    return ub.Mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny), oracle.mesh(0, 4 * np.pi, nx, 0, 2 * np.pi, ny)


def _random_emesh(m, rng):
    e = np.asfortranarray(rng.standard_normal((2, m.nx + 1, m.ny + 1)))
    e[:, m.nx, :] = e[:, 0, :]
    e[:, :, m.ny] = e[:, :, 0]
    return e


def _particles(npart, nx=128, ny=64, seed=11):
    _, x, v = seeded_load(npart, nx, ny, seed)
    return x, v


# ---- reference-owned known-answer tests, on the GPU ------------------------------------------------
def test_poisson_reference_test():
    """test/test_poisson.jl:1-49"""
    mesh = ub.Mesh(0, 2 * np.pi / 0.5, 64, 0, 2 * np.pi / 1.0, 128)
    fields, sol = ub.MeshFields(mesh), ub.MeshFields(mesh)
    x = np.linspace(mesh.xmin, mesh.xmax, mesh.nx + 1)
    y = np.linspace(mesh.ymin, mesh.ymax, mesh.ny + 1)
    fields.rho[:] = -8 * np.sin(2 * x)[:, None] * np.cos(2 * y)[None, :]
    sol.e[0] = 2 * np.cos(2 * x)[:, None] * np.cos(2 * y)[None, :]
    sol.e[1] = -2 * np.sin(2 * x)[:, None] * np.sin(2 * y)[None, :]
    poisson = ub.Poisson(mesh)
    poisson(fields)
    assert ub.errors(fields, sol) < 1e-14
    fields.rho[:] = -4 * (np.sin(2 * x)[:, None] + np.cos(2 * y)[None, :])
    poisson(fields)
    sol.e[0] = (2 * np.cos(2 * x))[:, None] + 0 * y[None, :]
    sol.e[1] = 0 * x[:, None] - (2 * np.sin(2 * y))[None, :]
    assert ub.errors(fields, sol) < 1e-14


def test_particles_meshfields_interaction_reference_test():
    """test/test_particles.jl:16-76 (20 x 20 mesh: not a power of two)"""
    mesh = ub.Mesh(0.0, 20.0, 20, 0.0, 20.0, 20)
    fields = ub.MeshFields(mesh)
    p = ub.Particles(121, 1.0 / 121)
    k = 0
    for i in range(5, 16):
        for j in range(5, 16):
            p.x[0, k], p.x[1, k] = (i - 0.5) * mesh.dx, (j - 0.5) * mesh.dx
            k += 1
    ub.compute_rho_m6(fields, p)
    assert abs(ub.integrate(fields.rho, mesh)) < 1e-4
    for i in range(21):
        for j in range(21):
            fields.e[0, i, j], fields.e[1, i, j] = i * mesh.dx, j * mesh.dy
    ub.interpol_eb_m6(p, fields)
    assert np.abs(p.e[0] - p.x[0]).mean() < 1e-6
    assert np.abs(p.e[1] - p.x[1]).mean() < 1e-6


@pytest.mark.parametrize("nx,ny", [(128, 64), (64, 128), (256, 256), (32, 16), (20, 12)])
def test_poisson_vs_oracle_on_noise(corc, nx, ny):
    mesh, om = _mesh_pair(nx, ny)
    rng = np.random.default_rng(nx * 1000 + ny)
    f = ub.MeshFields(mesh)
    f.rho[:] = rng.standard_normal(f.rho.shape)
    e_ref = np.zeros_like(f.e, order="F")
    nrj_ref = corc.poisson(om, f.rho, e_ref)
    nrj = ub.Poisson(mesh)(f)
    scale = np.abs(e_ref).max()
    assert np.abs(f.e - e_ref).max() / scale < 1e-13
    assert abs(nrj - nrj_ref) / nrj_ref < 1e-13


# ---- mesh <-> particle kernels -----------------------------------------------------------------------
@pytest.mark.parametrize("wrap", [ub.WRAP_FORTRAN, ub.WRAP_JULIA])
def test_plain_gather_bit_exact(corc, wrap):
    mesh, om = _mesh_pair()
    rng = np.random.default_rng(1)
    x, _ = _particles(5000)
    x = np.asfortranarray(x + rng.integers(-2, 3, size=x.shape) * np.array([[4 * np.pi], [2 * np.pi]]))  # out-of-box too
    f = ub.MeshFields(mesh)
    f.e[:] = _random_emesh(mesh, rng)
    p = ub.Particles(5000, 1.0)
    p.x[:] = x
    xo = x.copy(order="F")
    ep_ref = np.zeros((2, 5000), order="F")
    corc.interpol_eb_m6(om, f.e, xo, ep_ref, wrap)
    ub.interpol_eb_m6(p, f, wrap=wrap)
    assert np.array_equal(p.e, ep_ref)
    assert np.array_equal(p.x, xo)          # Julia wraps x in place; Fortran leaves it


def test_tau_gather_bit_exact(corc):
    mesh, om = _mesh_pair()
    rng = np.random.default_rng(2)
    x, v = _particles(700)
    ep = np.asfortranarray(rng.standard_normal((2, 700)))
    for ntau in (8, 16, 32):
        _, _, _, _, xt, _ = corc.preparation(ntau, 0.1, DT, x, v, ep)
        f = ub.MeshFields(mesh)
        f.e[:] = _random_emesh(mesh, rng)
        et_ref = np.zeros((ntau, 2, 700), order="F")
        et = np.zeros((ntau, 2, 700), order="F")
        for wrap in (ub.WRAP_FORTRAN, ub.WRAP_JULIA):
            corc.interpol_eb_m6_tau(om, f.e, xt, et_ref, wrap)
            ub.interpol_eb_m6(et, f, xt, 700, ntau, wrap=wrap)
            assert np.array_equal(et, et_ref)


@pytest.mark.parametrize("wrap", [ub.WRAP_FORTRAN, ub.WRAP_JULIA])
def test_plain_deposit(corc, wrap):
    mesh, om = _mesh_pair()
    x, _ = _particles(20000)
    w = 8 * np.pi ** 2 / 20000
    f = ub.MeshFields(mesh)
    p = ub.Particles(20000, w)
    p.x[:] = x
    rho_ref = np.zeros_like(f.rho, order="F")
    xo = x.copy(order="F")
    tot_ref = corc.compute_rho_m6(om, xo, w, rho_ref, wrap)
    tot = ub.compute_rho_m6(f, p, wrap=wrap)
    assert np.abs(f.rho - rho_ref).max() < 1e-13 * np.abs(rho_ref).max() + 1e-14
    assert abs(tot - tot_ref) < 1e-12
    assert np.array_equal(p.x, xo)
    # fixed point: bit-exact against the oracle's fixed-point twin, and reproducible
    scale = ub._lib.fixed_point_scale(8 * np.pi ** 2)
    rho_fx = np.zeros_like(f.rho, order="F")
    tot_fx = corc.compute_rho_m6_fixed(om, x, w, scale, rho_fx, wrap)
    p.x[:] = x
    tot1 = ub.compute_rho_m6(f, p, wrap=wrap, deposit_mode=ub.DEPOSIT_FIXED_POINT)
    assert np.array_equal(f.rho, rho_fx)
    assert tot1 == tot_fx
    assert np.abs(rho_fx - rho_ref).max() < 1e-10 * np.abs(rho_ref).max()


def test_deposit_edge_particle_on_upper_boundary(corc):
    """edge case (i) of SURVEY 8a: a tiny negative coordinate makes modulo() return exactly nx"""
    mesh, om = _mesh_pair(32, 16)
    x = np.asfortranarray(np.array([[-1e-18, 3.0, 4 * np.pi - 1e-15], [1.0, -1e-19, 2.0]]))
    w = 1.0
    f = ub.MeshFields(mesh)
    p = ub.Particles(3, w)
    p.x[:] = x
    rho_ref = np.zeros_like(f.rho, order="F")
    corc.compute_rho_m6(om, x.copy(order="F"), w, rho_ref, oracle.WRAP_FORTRAN)
    ub.compute_rho_m6(f, p, wrap=ub.WRAP_FORTRAN)
    assert np.abs(f.rho - rho_ref).max() < 1e-13 * np.abs(rho_ref).max()


# ---- UA stages ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("ntau", [2, 4, 8, 16, 32])
def test_fft_tau(corc, ntau):
    rng = np.random.default_rng(ntau)
    a = np.asfortranarray(rng.standard_normal((ntau, 2, 333)) + 1j * rng.standard_normal((ntau, 2, 333)))
    out = np.zeros_like(a, order="F")
    ub.fft_tau(out, ntau, a)
    assert np.abs(out - np.fft.fft(a, axis=0)).max() < 1e-13
    b = out.copy(order="F")
    ub.ifft_tau(b)
    assert np.abs(b - a).max() < 1e-14


@pytest.mark.parametrize("ntau,eps", [(16, 0.1), (32, 0.1), (8, 0.1), (16, 1e-3),
                                      (6, 0.1), (12, 0.1), (20, 1e-2), (48, 0.1), (64, 0.1), (250, 0.1)])    # general kernels (uapic_generic.cu)
def test_stage_chain_vs_oracle(corc, ntau, eps):
    """preparation -> gather -> compute_f -> ua_step1 x2 -> deposit -> gather -> compute_f -> ua_step2 x2 -> deposit -> compute_v,
    each GPU stage fed with the ORACLE's inputs so errors do not accumulate across stages"""
    mesh, om = _mesh_pair()
    rng = np.random.default_rng(7)
    npart = 1201          # deliberately not a multiple of the particles-per-warp
    x, v = _particles(npart, seed=3)
    w = 8 * np.pi ** 2 / npart
    emesh = _random_emesh(mesh, rng) * 0.3
    ep = np.asfortranarray(rng.standard_normal((2, npart)) * 0.3)
    tol = 2e-12
    shp = (ntau, 2, npart)

    def close(a, b, t=tol):
        return np.abs(a - b).max() <= t * max(1.0, np.abs(b).max())

    # oracle chain (Fortran conventions)
    b, t, pl, ql, xt, yt = corc.preparation(ntau, eps, DT, x, v, ep)
    # GPU preparation
    p = ub.Particles(npart, w)
    p.x[:], p.v[:], p.e[:] = x, v, ep
    ua = ub.UA(ntau, eps, npart, wrap=ub.WRAP_FORTRAN)
    gxt, gyt = np.zeros(shp, np.complex128, order="F"), np.zeros(shp, np.complex128, order="F")
    ub.preparation(ua, DT, p, gxt, gyt)
    assert close(p.b, b) and close(p.t, t) and close(ua.pl, pl) and close(ua.ql, ql)
    assert close(gxt, xt) and close(gyt, yt)

    et = np.zeros(shp, order="F")
    corc.interpol_eb_m6_tau(om, emesh, xt, et)
    fx, fy = corc.compute_f(eps, b, xt, yt, et, normalise=True)
    p.b[:], p.t[:] = b, t
    ua.pl[:], ua.ql[:] = pl, ql
    gfx, gfy = np.zeros(shp, np.complex128, order="F"), np.zeros(shp, np.complex128, order="F")
    ub.compute_f(gfx, gfy, ua, p, xt, yt, et, normalise=True)
    assert close(gfx, fx) and close(gfy, fy)
    ub.compute_f(gfx, gfy, ua, p, xt, yt, et, normalise=False)       # Julia normalisation
    assert close(gfx, fx * ntau) and close(gfy, fy * ntau)

    # predictor: Fortran form and Julia form
    xt_p, yt_p = xt.copy(order="F"), yt.copy(order="F")
    xf = corc.ua_step1(eps, t, pl, xt_p, fx)
    yf = corc.ua_step1(eps, t, pl, yt_p, fy)
    g1, g1f = xt.copy(order="F"), np.zeros(shp, np.complex128, order="F")
    ub.ua_step1(g1, g1f, ua, p, fx)
    assert close(g1f, xf) and close(g1, xt_p)
    g2, g2f = yt.copy(order="F"), np.zeros(shp, np.complex128, order="F")
    ub.ua_step1(g2, g2f, ua, p, fy)
    assert close(g2f, yf) and close(g2, yt_p)
    jl = np.zeros(shp, np.complex128, order="F")
    ub.ua_step(jl, xf, ua, p, np.asfortranarray(fx * ntau))          # unnormalised Julia inputs
    ub.ifft_tau(jl)
    assert close(jl, xt_p)

    # predictor deposit
    rho_ref = np.zeros((mesh.nx + 1, mesh.ny + 1), order="F")
    x_ref = np.zeros((2, npart), order="F")
    for wrap in (ub.WRAP_FORTRAN, ub.WRAP_JULIA):
        corc.compute_rho_m6_tau(om, eps, xt_p, t, w, rho_ref, x_ref, wrap)
        f = ub.MeshFields(mesh)
        ua.wrap = wrap
        ub.update_particles_x(p, f, ua, xt_p)
        assert close(p.x, x_ref)
        assert np.abs(f.rho - rho_ref).max() < 1e-11 * np.abs(rho_ref).max()
    ua.wrap = ub.WRAP_FORTRAN

    # corrector
    corc.interpol_eb_m6_tau(om, emesh, xt_p, et)
    gx, gy = corc.compute_f(eps, b, xt_p, yt_p, et, normalise=True)
    xt_c, yt_c = xt_p.copy(order="F"), yt_p.copy(order="F")
    corc.ua_step2(eps, t, pl, ql, xt_c, xf, fx, gx)
    corc.ua_step2(eps, t, pl, ql, yt_c, yf, fy, gy)
    g3 = np.zeros(shp, np.complex128, order="F")
    ub.ua_step2(g3, xf, ua, p, fx, gx)
    assert close(g3, xt_c)
    g4 = np.zeros(shp, np.complex128, order="F")
    ub.ua_step2(g4, yf, ua, p, fy, gy)
    assert close(g4, yt_c)
    jl2 = np.zeros(shp, np.complex128, order="F")
    ub.ua_step(jl2, yf, ua, p, np.asfortranarray(fy * ntau), np.asfortranarray(gy * ntau))
    v_ref = np.zeros((2, npart), order="F")
    corc.compute_v(eps, t, yt_c, v_ref)
    tolv = min(tol, 1e-12) * max(1.0, 0.1 / eps)      # 30x the measured GPU-vs-oracle distance (profiles/r2b_small_eps_parity.json)
    ub.compute_v(jl2, p, ua, yt_is_fourier=True)                      # Julia: Fourier coefficients (unnormalised)
    assert np.abs(p.v - v_ref).max() < tolv * np.abs(v_ref).max()
    ub.compute_v(yt_c, p, ua, yt_is_fourier=False)                    # Fortran: time domain
    assert np.abs(p.v - v_ref).max() < tolv * np.abs(v_ref).max()


def test_bad_arguments_fail_loudly():
    mesh = ub.Mesh(0, 1, 8, 0, 1, 8)
    with pytest.raises(ValueError):
        ub.UA(13, 0.1, 10)                  # odd
    with pytest.raises(ValueError):
        ub.UA(258, 0.1, 10)                 # beyond the general kernels
    p = ub.Particles(4, 1.0)
    ua = ub.UA(16, 0.1, 4)
    with pytest.raises(ValueError):
        ub.preparation(ua, 0.1, p, np.zeros((8, 2, 4), np.complex128, order="F"), np.zeros((16, 2, 4), np.complex128, order="F"))
    f = ub.MeshFields(ub.Mesh(0, 1, 1030, 0, 1, 8))
    with pytest.raises(ub.UapicError):
        ub.Poisson(f.mesh)(f)


@pytest.mark.parametrize("nx,ny,wrap", [(128, 64, ub.WRAP_FORTRAN), (20, 12, ub.WRAP_JULIA)])
def test_cic_stage_entry_points_vs_oracle(corc, nx, ny, wrap):
    """uapic_compute_rho_cic / uapic_interpol_eb_cic (build-defined bilinear shape, include/uapic_b200.h) against the oracle in
    its CIC mode; a linear field is reproduced exactly by bilinear interpolation (the check test/test_particles.jl:52-74 makes for M6)"""
    mesh, om = _mesh_pair(nx, ny)
    rng = np.random.default_rng(nx + ny)
    n = 20000
    p = ub.Particles(n, (mesh.xmax - mesh.xmin) * (mesh.ymax - mesh.ymin) / n)
    p.x[0], p.x[1] = rng.uniform(-1, 1, n) * 3 * (mesh.xmax - mesh.xmin), rng.uniform(-1, 1, n) * 3 * (mesh.ymax - mesh.ymin)
    f = ub.MeshFields(mesh)
    f.e[:] = rng.standard_normal(f.e.shape)
    f.e[:, nx, :], f.e[:, :, ny] = f.e[:, 0, :], f.e[:, :, 0]             # periodic ghosts, as every solve leaves them
    f.e[:, nx, ny] = f.e[:, 0, 0]
    corc.set_scheme("cic")
    try:
        xo, rho_o, ep_o = p.x.copy(order="F"), np.zeros_like(f.rho, order="F"), np.zeros_like(p.e, order="F")
        tot_o = corc.compute_rho_m6(om, xo, p.w, rho_o, wrap=wrap)
        xo2 = p.x.copy(order="F")
        corc.interpol_eb_m6(om, f.e, xo2, ep_o, wrap=wrap)
    finally:
        corc.set_scheme("m6")
    x0 = p.x.copy(order="F")
    tot = ub.compute_rho_cic(f, p, wrap=wrap)
    assert np.abs(f.rho - rho_o).max() < 1e-12 * max(1.0, np.abs(rho_o).max()) and abs(tot - tot_o) < 1e-9 * abs(tot_o)
    assert np.abs(p.x - xo).max() < 1e-12 * (mesh.xmax - mesh.xmin)        # the wrap convention's effect on x, as the oracle
    p.x[:] = x0
    ub.interpol_eb_cic(p, f, wrap=wrap)
    assert np.abs(p.e - ep_o).max() < 1e-13 * np.abs(ep_o).max()
    # bilinear interpolation reproduces a field that is linear inside every cell
    q = ub.Particles(500, 1.0)
    q.x[0], q.x[1] = rng.uniform(0.05, 0.9, 500) * (mesh.xmax - mesh.xmin), rng.uniform(0.05, 0.9, 500) * (mesh.ymax - mesh.ymin)
    for i in range(nx + 1):
        f.e[0, i, :], f.e[1, i, :] = i * mesh.dx, np.arange(ny + 1) * mesh.dy
    ub.interpol_eb_cic(q, f, wrap=ub.WRAP_FORTRAN)
    assert np.abs(q.e - q.x).max() < 1e-12
