"""GPU parity tests of the fused session path (the performance path) against the CPU oracle and the committed
golden vectors: the whole loop of fortran/bupdate.F90:89-128.

Tolerance (north star): 1e-10 relative on positions, velocities and the electric-energy history at eps = 0.1.
For smaller eps the velocity tolerance is scaled by 0.1/eps: the scheme multiplies every rounding difference in
b(x) by t/eps (SURVEY.md section 7 "eps-amplified rounding"); the reference's own two implementations (Julia vs
Fortran conventions, tests/test_oracle.py) differ from each other by the same amount.
"""
import glob
import os

import numpy as np
import pytest

import oracle
import uapic_b200 as ub

from conftest import golden_files, GOLDEN, periodic_diff, seeded_load

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _two_barrier_kernels_by_default():
    """this module targets the two-barrier kernels (uapic_fused.cu); the one-pass kernels, which a Session picks by default
    for ntau = 8, 16, 32, have their own module (test_gpu_onepass.py)"""
    import uapic_b200.session as sess
    old = sess.DEFAULT_STORAGE
    sess.DEFAULT_STORAGE = ub.STORE_FULL
    yield
    sess.DEFAULT_STORAGE = old

DT = np.pi / 16
DIMX, DIMY = 4 * np.pi, 2 * np.pi


def _compare(xg, vg, eng, xo, vo, eno, eps, tol=1e-10):
    # measured (profiles/r2b_small_eps_parity.json): GPU-vs-oracle distance of v is 3.4e-14 at eps = 0.1 and grows like 1/eps
    # (3.1e-10 at 1e-5) -- both sit within 2e-10 of the extended-precision referee there.  30x that, instead of round 1's
    # 1e-10 * 0.1/eps (which was 1e-6 at eps = 1e-5).
    tolv = min(tol, 1e-12) * max(1.0, 0.1 / eps)
    assert periodic_diff(xg[0], xo[0], DIMX).max() < tol * DIMX
    assert periodic_diff(xg[1], xo[1], DIMY).max() < tol * DIMY
    assert np.abs(vg - vo).max() < tolv * np.abs(vo).max()
    assert eng.shape == eno.shape
    assert np.abs(eng - eno).max() / np.abs(eno).max() < tol


@pytest.mark.parametrize("ntau,nx,ny,npart,nstep,eps", [
    (16, 128, 64, 20000, 8, 0.1),       # config 1 (as shipped) at reduced particle count, all 8 steps
    (32, 128, 128, 6000, 4, 0.1),       # config 3 shape
    (8, 64, 32, 5000, 4, 0.1),
    (16, 128, 64, 6001, 4, 1e-2),       # odd particle count: ragged last warp
    (16, 128, 64, 6000, 4, 1e-3),
])
def test_session_vs_oracle(corc, ntau, nx, ny, npart, nstep, eps):
    om, x0, v0 = seeded_load(npart, nx, ny, seed=100 + ntau)
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    w = DIMX * DIMY / npart
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    eno, _, ep_o, emesh_o = corc.run_bupdate(om, ntau, eps, DT, nstep, xo, vo, w)
    xg, vg, eng, emesh_g = ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w)
    _compare(xg, vg, eng, xo, vo, eno, eps)
    assert np.abs(emesh_g - emesh_o).max() < 1e-10 * np.abs(emesh_o).max()


def test_session_julia_wrap_vs_oracle(corc):
    npart, ntau, eps, nstep = 5000, 16, 0.1, 3
    om, x0, v0 = seeded_load(npart, seed=9)
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    w = DIMX * DIMY / npart
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    eno, _, _, _ = corc.run_bupdate(om, ntau, eps, DT, nstep, xo, vo, w, wrap=oracle.WRAP_JULIA)
    xg, vg, eng, _ = ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w, wrap=ub.WRAP_JULIA)
    _compare(xg, vg, eng, xo, vo, eno, eps)
    assert xg[0].min() >= 0 and xg[0].max() < DIMX and xg[1].min() >= 0 and xg[1].max() < DIMY   # stored wrapped


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_session_against_golden(path):
    g = np.load(path)
    nx, ny, ntau, nstep = int(g["nx"]), int(g["ny"]), int(g["ntau"]), int(g["nstep"])
    eps, dt, w = float(g["eps"]), float(g["dt"]), float(g["w"])
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    xg, vg, eng, emesh = ub.run_bupdate(mesh, ntau, eps, dt, nstep, np.asfortranarray(g["x0"]), np.asfortranarray(g["v0"]), w)
    _compare(xg, vg, eng, g["x"], g["v"], g["energy"], eps)
    assert np.abs(emesh - g["emesh"]).max() < 1e-10 * np.abs(g["emesh"]).max()


def test_fixed_point_mode_is_bit_reproducible_and_close(corc):
    npart, ntau, eps, nstep = 8000, 16, 0.1, 3
    om, x0, v0 = seeded_load(npart, seed=21)
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    w = DIMX * DIMY / npart
    runs = [ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w, deposit_mode=ub.DEPOSIT_FIXED_POINT) for _ in range(3)]
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert np.array_equal(a, b)                      # bit-identical run to run
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    eno, _, _, _ = corc.run_bupdate(om, ntau, eps, DT, nstep, xo, vo, w)
    _compare(runs[0][0], runs[0][1], runs[0][2], xo, vo, eno, eps)


def test_sharded_fixed_point_equals_unsharded():
    """bit-exactness across "GPU counts" emulated on one device: the int64 raw meshes of two half-size shards are summed
    through the allreduce hook exactly as NCCL would, and every later stage must then reproduce the unsharded bits."""
    import threading
    import torch

    npart, ntau, eps, nstep = 6000, 16, 0.1, 2
    _, x0, v0 = seeded_load(npart, seed=33)
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    w = DIMX * DIMY / npart
    ref = ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w, deposit_mode=ub.DEPOSIT_FIXED_POINT)

    world = 2
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world
    errors = []

    def worker(rank):
        try:
            lo, hi = ub.dist.shard_range(npart, rank, world)
            s = ub.Session(mesh, ntau, eps, DT, hi - lo, weight=w, nbpart_global=npart, deposit_mode=ub.DEPOSIT_FIXED_POINT)

            def reduce(ptr, count, dtype, stream):
                assert dtype == 1
                torch.cuda.synchronize()
                view = torch.as_tensor(ub.dist._CudaView(ptr, count, "<i8"), device="cuda")
                slots[rank] = view
                barrier.wait()
                total = slots[0] + slots[1]
                torch.cuda.synchronize()
                barrier.wait()
                view.copy_(total)
                torch.cuda.synchronize()
                barrier.wait()
                return 0

            s.set_allreduce(reduce)
            s.upload_particles(np.asfortranarray(x0[:, lo:hi]), np.asfortranarray(v0[:, lo:hi]))
            s.init_fields()
            s.step(nstep)
            s.synchronize()
            x, v = s.download_particles()
            e, _ = s.download_fields()
            results[rank] = (x, v, s.energy_history(), e)
            s.close()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)
            barrier.abort()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    x = np.concatenate([results[0][0], results[1][0]], axis=1)
    v = np.concatenate([results[0][1], results[1][1]], axis=1)
    assert np.array_equal(x, ref[0]) and np.array_equal(v, ref[1])
    assert np.array_equal(results[0][2], ref[2]) and np.array_equal(results[1][2], ref[2])
    assert np.array_equal(results[0][3], ref[3])


def test_device_loaders_and_diagnostics():
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    npart = 200000
    with ub.Session(mesh, 16, 0.1, DT, npart) as s:
        s.generate_particles("plasma", seed=5)
        x, v = s.download_particles()
        assert x[0].min() >= 0 and x[0].max() < DIMX and x[1].min() >= 0 and x[1].max() < DIMY
        assert np.abs(v).max() <= 5.0
        # densities of particles.F90:77,94: <sin y> = 1/2 under 1+sin y ; <vx^2> = 1 + 4 for the two bumps at +-2
        assert abs(np.sin(x[1]).mean() - 0.5) < 0.01
        assert abs((v[0] ** 2).mean() - 5.0) < 0.1 and abs((v[1] ** 2).mean() - 1.0) < 0.05
        s.init_fields()
        s.step(1)
        s.synchronize()
        x1, v1 = s.download_particles()
        assert np.allclose(s.sum_v(), v1.sum(axis=1), rtol=1e-10, atol=1e-8)
        assert s.energy_history().shape == (3,)
        assert s.launch_count > 0 and s.device_bytes > npart * 16 * 8 * 16
        s.generate_particles("landau", seed=6)
        x, v = s.download_particles()
        r2 = (v ** 2).sum(axis=0)
        assert abs(r2.mean() - 2.0) < 0.02                 # |v|^2 = -2 ln u  => mean 2
        assert x[0].min() >= 0 and x[0].max() <= DIMX


@pytest.mark.parametrize("ntau,nx,ny,eps", [(16, 128, 64, 0.1), (32, 128, 128, 0.1), (16, 128, 64, 1e-3)])
def test_hybrid_storage_vs_oracle_and_store_full(corc, ntau, nx, ny, eps):
    """UAPIC_STORE_HYBRID keeps 16 B per particle-tau (E at the tau samples) across the barrier and recomputes the
    predictor in phase B: same answer as the oracle, and as the store-full layout to round-off."""
    npart, nstep = 6000, 4
    om, x0, v0 = seeded_load(npart, nx, ny, seed=71)
    mesh = ub.Mesh(0, DIMX, nx, 0, DIMY, ny)
    w = DIMX * DIMY / npart
    xo, vo = x0.copy(order="F"), v0.copy(order="F")
    eno, _, _, _ = corc.run_bupdate(om, ntau, eps, DT, nstep, xo, vo, w)
    res = {}
    for name, mode in (("full", ub.STORE_FULL), ("hybrid", ub.STORE_HYBRID)):
        with ub.Session(mesh, ntau, eps, DT, npart, weight=w, storage_mode=mode) as s:
            s.upload_particles(x0, v0)
            s.init_fields()
            s.step(nstep)
            s.synchronize()
            x, v = s.download_particles()
            res[name] = (x, v, s.energy_history(), s.device_bytes)
    _compare(res["hybrid"][0], res["hybrid"][1], res["hybrid"][2], xo, vo, eno, eps)
    _compare(res["hybrid"][0], res["hybrid"][1], res["hybrid"][2], res["full"][0], res["full"][1], res["full"][2], eps, tol=1e-11)
    assert res["hybrid"][3] < res["full"][3] * 0.3          # 16 B instead of 128 B per particle-tau


def test_session_state_errors():
    mesh = ub.Mesh(0, DIMX, 32, 0, DIMY, 16)
    with ub.Session(mesh, 16, 0.1, DT, 100) as s:
        with pytest.raises(ub.UapicError):
            s.init_fields()                 # no particles yet
        with pytest.raises(ub.UapicError):
            s.step(1)
    with pytest.raises(ub.UapicError):
        ub.Session(mesh, 7, 0.1, DT, 100)                  # odd ntau: the reference's tables need an even one too (ua_type.F90:51-56)


def test_stage_api_runs_the_reference_script_sequence():
    """test/bupdate.jl:63-114 call for call through the stage API (Julia conventions: unnormalised fft!, yt left in Fourier
    space after the corrector), against the numpy twin of the Julia sources and against the fused session."""
    from oracle import nporc
    npart, ntau, eps, nstep = 3000, 16, 0.1, 3
    _, x0, v0 = seeded_load(npart, seed=61)
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    w = DIMX * DIMY / npart
    fields = ub.MeshFields(mesh)
    p = ub.Particles(npart, w)
    p.x[:], p.v[:] = x0, v0
    poisson = ub.Poisson(mesh)
    ua = ub.UA(ntau, eps, npart)
    shp = (ntau, 2, npart)
    z = lambda: np.zeros(shp, np.complex128, order="F")  # noqa: E731
    et = np.zeros(shp, order="F")
    xt, xft, yt, yft, fx, fy, gx, gy = z(), z(), z(), z(), z(), z(), z(), z()
    nrj = []
    ub.compute_rho_m6(fields, p)                               # :63
    nrj.append(poisson(fields))                                # :65
    ub.interpol_eb_m6(p, fields)                               # :67
    for _ in range(nstep):
        ub.preparation(ua, DT, p, xt, yt)                      # :71
        ub.update_particles_e(p, et, fields, ua, xt)           # :73
        ub.compute_f(fx, fy, ua, p, xt, yt, et)                # :77
        ub.fft_tau(xft, ua, xt)                                # :79
        ub.ua_step(xt, xft, ua, p, fx)                         # :80
        ub.fft_tau(yft, ua, yt)                                # :82
        ub.ua_step(yt, yft, ua, p, fy)                         # :83
        ub.ifft_tau(xt)                                        # :85
        ub.ifft_tau(yt)                                        # :86
        ub.update_particles_x(p, fields, ua, xt)               # :88
        nrj.append(poisson(fields))                            # :90
        ub.update_particles_e(p, et, fields, ua, xt)           # :92
        ub.compute_f(gx, gy, ua, p, xt, yt, et)                # :96
        ub.ua_step(xt, xft, ua, p, fx, gx)                     # :98
        ub.ua_step(yt, yft, ua, p, fy, gy)                     # :100
        ub.ifft_tau(xt)                                        # :102
        ub.update_particles_x(p, fields, ua, xt)               # :104
        nrj.append(poisson(fields))                            # :106
        ub.compute_v(yt, p, ua)                                # :110
    nrj = np.array(nrj)
    mm = nporc.Mesh(0, DIMX, 128, 0, DIMY, 64)
    xo, vo, eno, _, _ = nporc.run_bupdate(mm, ntau, eps, DT, nstep, x0, v0, w)
    _compare(p.x, p.v, nrj, xo, vo, eno, eps)
    xs, vs, ens, _ = ub.run_bupdate(mesh, ntau, eps, DT, nstep, x0, v0, w, wrap=ub.WRAP_JULIA)
    _compare(xs, vs, ens, p.x, p.v, nrj, eps)
