"""The external-field program fortran/efd.f90 -- oracle pinned by the reference's OWN numbers.

efd.f90:481 (and test/test_efd.jl:468) print  sum(v1) + 857.95049281063064,  sum(v2) + 593.40700170710875 : the sums of
the final velocities of a complete 204 800-particle run, recorded by the reference's authors.  The load is
init_particles_2d (particles.F90:54-103: gfortran `random_number` under a fixed 33-word seed); the bundled libgfortran
gives that stream here, so both restatements of the program can be held to those constants."""
import numpy as np
import pytest

import oracle
import uapic_b200 as ub

REF_SUM_V = (-857.95049281063064, -593.40700170710875)       # efd.f90:481


@pytest.fixture(scope="module")
def reference_load():
    mesh = ub.Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)      # efd.f90:64-65,94-98
    p, src = ub.plasma(mesh, 204800, use_gfortran=True, return_source=True)
    if "libgfortran" not in src:
        pytest.skip("libgfortran not loadable here")
    return p


@pytest.fixture(scope="module")
def numpy_run(reference_load):
    return oracle.efd_np.efd_run(reference_load.x, reference_load.v)


def _c_run(x, v, **kw):
    c = oracle.corc()
    c.set_threads(c.max_threads())
    try:
        return c.efd_run(x, v, **kw)
    finally:
        c.set_threads(1)


def test_numpy_restatement_reproduces_the_reference_constants(numpy_run):
    x, v = numpy_run
    # 2e5 values of size <= 5 summed: 1e-9 absolute is 12 digits of the printed constants
    assert abs(v[0].sum() - REF_SUM_V[0]) < 1e-9 and abs(v[1].sum() - REF_SUM_V[1]) < 1e-9
    assert x[0].min() >= 0 and x[0].max() <= 4 * np.pi and x[1].min() >= 0 and x[1].max() <= 2 * np.pi


def test_c_restatement_reproduces_the_reference_constants(reference_load, numpy_run):
    x, v = _c_run(reference_load.x, reference_load.v)
    assert abs(v[0].sum() - REF_SUM_V[0]) < 1e-9 and abs(v[1].sum() - REF_SUM_V[1]) < 1e-9
    xn, vn = numpy_run
    assert np.abs(v - vn).max() < 1e-10 and np.abs(x - xn).max() < 1e-10


def test_load_is_not_incidental():
    """the constants are sensitive to the load: a different stream misses them by O(100)"""
    mesh = ub.Mesh(0, 4 * np.pi, 128, 0, 2 * np.pi, 64)
    p = ub.plasma(mesh, 204800, seed=7)
    _, v = _c_run(p.x, p.v)
    assert abs(v[0].sum() - REF_SUM_V[0]) > 1.0


@pytest.mark.parametrize("ntau", [8, 16, 32, 12, 50])
def test_two_restatements_agree(ntau):
    rng = np.random.default_rng(ntau)
    n = 300
    x = np.asfortranarray(rng.random((2, n)) * [[4 * np.pi], [2 * np.pi]])
    v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
    for eps in (1e-1, 1e-3):
        xc, vc = oracle.corc().efd_run(x, v, ntau=ntau, eps=eps)
        xn, vn = oracle.efd_np.efd_run(x, v, ntau=ntau, eps=eps)
        assert np.abs(vc - vn).max() < 1e-10 and np.abs(xc - xn).max() < 1e-10


@pytest.mark.parametrize("eps", [1e-1, 1e-2, 1e-3, 1e-4])
def test_extended_precision_referee_ranks_the_restatements(eps):
    """the same formulas in x87 long double (numpy's FFT included) from the same double inputs: both double restatements sit
    within the conditioning of the map -- one input ulp moves v by ~1e-12 * 0.1/eps -- of the exact evaluation"""
    rng = np.random.default_rng(3)
    n = 1000
    x = np.asfortranarray(rng.random((2, n)) * [[4 * np.pi], [2 * np.pi]])
    v = np.asfortranarray(rng.normal(size=(2, n)) * 2)
    xr, vr = oracle.efd_np.efd_run(x, v, eps=eps, real=np.longdouble)
    assert xr.dtype == np.longdouble
    bound = 1e-12 * max(1.0, 0.1 / eps) * float(np.abs(vr).max())
    for xo, vo in (oracle.corc().efd_run(x, v, eps=eps), oracle.efd_np.efd_run(x, v, eps=eps)):
        assert float(np.abs(xo - xr).max()) < 1e-12 and float(np.abs(vo - vr).max()) < bound
