"""CPU check of the algebra behind the one-pass kernels (uapic.jl_b200/csrc/uapic_onepass.cu): the reorganised step of
tests/onepass_algebra_np.py -- both deposits before the field solves, compute_v as a time-domain weighted sum -- must
reproduce the reference sequence of test/bupdate.jl:63-114 (numpy twin of the Julia sources) to round-off."""
import numpy as np
import pytest

import oracle
from oracle import uapic_oracle_np as onp

import onepass_algebra_np as op

DT = np.pi / 16
DIMX, DIMY = 4 * np.pi, 2 * np.pi


@pytest.mark.parametrize("ntau,eps,nx,ny,tolv", [(16, 0.1, 128, 64, 1e-11), (32, 0.1, 32, 32, 1e-11), (8, 1e-2, 32, 16, 1e-10),
                                                 (16, 1e-3, 64, 32, 1e-9)])
def test_onepass_algebra_equals_reference_sequence(ntau, eps, nx, ny, tolv):
    m = onp.Mesh(0, DIMX, nx, 0, DIMY, ny)
    om = oracle.mesh(0, DIMX, nx, 0, DIMY, ny)
    npart, nstep = 1500, 3
    rng = np.random.default_rng(7)
    x0, v0, _ = oracle.corc().plasma_from_uniforms(om, npart, 0.05, 0.5, rng.random(npart * 80))
    w = DIMX * DIMY / npart
    xr, vr, er, _, _ = onp.run_bupdate(m, ntau, eps, DT, nstep, x0, v0, w)
    xo, vo, eo = op.run_bupdate_onepass(m, ntau, eps, DT, nstep, x0, v0, w)
    assert np.abs(np.mod(xr[0] - xo[0] + DIMX / 2, DIMX) - DIMX / 2).max() < 1e-12 * DIMX
    assert np.abs(np.mod(xr[1] - xo[1] + DIMY / 2, DIMY) - DIMY / 2).max() < 1e-12 * DIMY
    assert np.abs(vr - vo).max() < tolv * np.abs(vr).max()
    assert np.abs(er - eo).max() < 1e-13 * np.abs(er).max()


def test_corrector_position_does_not_depend_on_the_predictor_field():
    """identity 1: zeroing the predictor field changes v but leaves the corrector deposit position untouched"""
    ntau, eps, nx, ny, npart = 16, 0.1, 32, 16, 400
    m = onp.Mesh(0, DIMX, nx, 0, DIMY, ny)
    om = oracle.mesh(0, DIMX, nx, 0, DIMY, ny)
    rng = np.random.default_rng(3)
    x0, v0, _ = oracle.corc().plasma_from_uniforms(om, npart, 0.05, 0.5, rng.random(npart * 80))
    w = DIMX * DIMY / npart
    rho = np.zeros((nx + 1, ny + 1)); e = np.zeros((2, nx + 1, ny + 1)); ep = np.zeros_like(x0)
    x = x0.copy()
    onp.compute_rho_m6(m, rho, x, w)
    onp.Poisson(m)(rho, e)
    onp.interpol_eb_m6(m, e, x, ep)
    A = op.phase_a(m, ntau, eps, DT, x, v0, ep, e)
    v_a = op.phase_b(m, ntau, eps, A, e)
    v_b = op.phase_b(m, ntau, eps, A, np.zeros_like(e))
    assert np.abs(v_a - v_b).max() > 1e-6          # the field matters for v ...
    assert "pos_c" in A and np.isfinite(A["pos_c"]).all()   # ... but pos_c was fixed before any predictor field existed
