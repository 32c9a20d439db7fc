"""The external-field program (fortran/efd.f90) on the GPU: `uapic_efd_run` against the two CPU restatements, and against the
numbers the reference itself prints (efd.f90:481) when run on init_particles_2d's own load."""
import numpy as np
import pytest

import oracle
import uapic_b200 as ub

pytestmark = pytest.mark.gpu

DIMX, DIMY = 4 * np.pi, 2 * np.pi
# The map is ill-conditioned in v like the UA loop: moving the input by ONE ulp moves v by 1.8e-13 / 1.8e-12 / 1.2e-11 / 1.4e-10
# at eps = 1e-1 / 1e-2 / 1e-3 / 1e-4 (and x by 3e-14), and the two CPU restatements differ from each other by the same amounts.
# So x is held to 1e-12 of the box and v to 1e-12 * 0.1/eps of max|v| -- inside the north star's 1e-10 down to eps = 1e-3.


def _load(n, seed):
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.random((2, n)) * [[DIMX], [DIMY]])
    v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
    return x, v


def _close(xg, vg, xo, vo, eps=1e-3):
    tol, tolv = 1e-12, 1e-12 * max(1.0, 0.1 / eps)
    dx = np.abs(xg - xo)
    dx[0] = np.minimum(dx[0], DIMX - dx[0])          # a value wrapped on one side of the box edge and not on the other
    dx[1] = np.minimum(dx[1], DIMY - dx[1])
    assert dx[0].max() < tol * DIMX and dx[1].max() < tol * DIMY
    assert np.abs(vg - vo).max() < tolv * max(1.0, np.abs(vo).max())


@pytest.mark.parametrize("ntau", [2, 4, 8, 16, 32])
@pytest.mark.parametrize("eps", [1e-1, 1e-3])
def test_lane_path_vs_oracle(corc, ntau, eps):
    x, v = _load(1003, ntau)                          # 1003: the last warp has idle particle groups
    xg, vg = ub.efd_run(x, v, ntau=ntau, eps=eps)
    xo, vo = corc.efd_run(x, v, ntau=ntau, eps=eps)
    _close(xg, vg, xo, vo, eps)


@pytest.mark.parametrize("ntau", [6, 12, 20, 50, 64, 100, 250])
def test_any_even_ntau_vs_oracle(corc, ntau):
    x, v = _load(131, ntau)
    for eps in (1e-1, 1e-3):
        xg, vg = ub.efd_run(x, v, ntau=ntau, eps=eps)
        xo, vo = corc.efd_run(x, v, ntau=ntau, eps=eps)
        _close(xg, vg, xo, vo, eps)


def test_vs_numpy_restatement_other_parameters():
    x, v = _load(4096, 5)
    kw = dict(ntau=16, eps=1e-2, dt=np.pi / 32, tfinal=np.pi / 4)
    xg, vg = ub.efd_run(x, v, **kw)
    xn, vn = oracle.efd_np.efd_run(x, v, **kw)
    _close(xg, vg, xn, vn, 1e-2)
    xg2, vg2 = ub.efd_run(x, v, nstep=3, **kw)        # explicit step count instead of nint(tfinal/dt) = 8
    assert np.abs(vg2 - vg).max() > 1e-6


@pytest.mark.parametrize("eps", [1e-1, 1e-3, 1e-4])
def test_vs_extended_precision_referee(corc, eps):
    """against the formulas evaluated in long double (oracle/efd_np.py, real=np.longdouble): the kernel must be as close to the
    exact evaluation as the C restatement is (within 10x), and inside the conditioning bound"""
    x, v = _load(2000, 21)
    xg, vg = ub.efd_run(x, v, eps=eps)
    xr, vr = oracle.efd_np.efd_run(x, v, eps=eps, real=np.longdouble)
    xo, vo = corc.efd_run(x, v, eps=eps)
    d_gpu, d_orc = float(np.abs(vg - vr).max()), float(np.abs(vo - vr).max())
    assert d_gpu < 1e-12 * max(1.0, 0.1 / eps) * float(np.abs(vr).max())
    assert d_gpu < 10 * max(d_orc, 1e-14)
    assert float(np.abs(xg - xr).max()) < 1e-12


def test_reference_program_reproduces_the_printed_constants():
    """the whole program on the reference's own load: the two numbers efd.f90:481 prints must vanish"""
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    p, src = ub.plasma(mesh, 204800, use_gfortran=True, return_source=True)
    if "libgfortran" not in src:
        pytest.skip("libgfortran not loadable here")
    p, f, printed = ub.efd(16, particles=p)
    assert abs(printed[0]) < 1e-8 and abs(printed[1]) < 1e-8          # 11 digits of the reference's constants
    assert 0 <= p.x[0].min() and p.x[0].max() <= DIMX and 0 <= p.x[1].min() and p.x[1].max() <= DIMY
    assert abs(ub.integrate(f.rho, mesh)) < 1e-9      # neutralised deposit of the final positions (efd.f90:484)


def test_reference_program_as_shipped_one_particle(corc):
    """`efd(16, 1)` (test/test_efd.jl:483; `do m=1,1`, efd.f90:131): only the first particle moves"""
    mesh = ub.Mesh(0, DIMX, 128, 0, DIMY, 64)
    p0 = ub.plasma(mesh, 2048, seed=11)
    x0, v0 = p0.x.copy(order="F"), p0.v.copy(order="F")
    p, f, _ = ub.efd(16, 1, particles=p0)
    assert np.array_equal(p.x[:, 1:], x0[:, 1:]) and np.array_equal(p.v[:, 1:], v0[:, 1:])
    xo, vo = corc.efd_run(x0[:, :1], v0[:, :1])
    _close(p.x[:, :1], p.v[:, :1], xo, vo)


def test_particles_are_independent_and_runs_repeatable():
    x, v = _load(1_000_000, 9)
    xa, va = ub.efd_run(x, v)
    xb, vb = ub.efd_run(x, v)
    assert np.array_equal(xa, xb) and np.array_equal(va, vb)
    sl = slice(333_333, 333_333 + 4097)
    xs, vs = ub.efd_run(x[:, sl], v[:, sl])
    assert np.array_equal(xs, xa[:, sl]) and np.array_equal(vs, va[:, sl])
    assert np.isfinite(va).all() and 0 <= xa[0].min() and xa[0].max() <= DIMX


def test_device_entry_point_in_place():
    import torch
    x, v = _load(5000, 13)
    xg, vg = ub.efd_run(x, v)
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
    vd = torch.from_numpy(np.ascontiguousarray(v.T)).cuda()
    ub.efd_run_device(xd, vd, x_out=xd, v_out=vd)
    torch.cuda.synchronize()
    assert np.array_equal(xd.cpu().numpy().T, xg) and np.array_equal(vd.cpu().numpy().T, vg)


def test_edge_cases_and_errors():
    x, v = _load(0, 1)
    xg, vg = ub.efd_run(x, v)
    assert xg.shape == (2, 0) and vg.shape == (2, 0)
    x, v = _load(1, 2)
    xg, vg = ub.efd_run(x, v, tfinal=0.0)             # no steps: the prepared datum read back at tau = 0 is the input
    assert np.abs(xg - x).max() < 1e-12 and np.abs(vg - v).max() < 1e-12
    for bad in (dict(ntau=7), dict(ntau=258), dict(eps=0.0), dict(dt=-1.0), dict(box=(0.0, 0.0, 0.0, 1.0))):
        with pytest.raises(ub.UapicError):
            ub.efd_run(x, v, **bad)
    with pytest.raises(ValueError):
        ub.efd_run(np.zeros((3, 4)), np.zeros((3, 4)))


def test_one_sample_per_lane_policy(corc):
    """the library picks the tau policy once per process (UAPIC_EFD_SPL2, default: two samples per lane); the other lane policy is
    held to the same oracle in a child process"""
    import json
    import os
    import subprocess
    import sys
    x, v = _load(333, 17)
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import uapic_b200 as ub; "
            "d = np.load(sys.argv[1]); out = {}\n"
            "for ntau in (4, 8, 16, 32):\n"
            "    xg, vg = ub.efd_run(d['x'], d['v'], ntau=ntau)\n"
            "    out[str(ntau)] = [xg.tolist(), vg.tolist()]\n"
            "print('RESULT' + json.dumps(out))") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "load.npz")
        np.savez(path, x=x, v=v)
        env = dict(os.environ, UAPIC_EFD_SPL2="0")
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([line for line in r.stdout.splitlines() if line.startswith("RESULT")][-1][6:])
    for ntau in (4, 8, 16, 32):
        xo, vo = corc.efd_run(x, v, ntau=ntau)
        xg, vg = (np.array(a) for a in out[str(ntau)])
        _close(xg, vg, xo, vo)
        x2, v2 = ub.efd_run(x, v, ntau=ntau)                      # this process: the default policy
        _close(x2, v2, xo, vo)
