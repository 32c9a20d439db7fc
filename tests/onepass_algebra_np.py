"""numpy statement of the "one-pass" reorganisation of the UA step used by the sm_100a kernels of
``uapic.jl_b200/csrc/uapic_onepass.cu`` -- TEST INFRASTRUCTURE (it documents and checks the algebra; the product is CUDA).

The reference (fortran/bupdate.F90:97-123, test/bupdate.jl:71-110) runs, per step,

    preparation -> gather(E_n) -> compute_f -> ua_step1 -> deposit -> Poisson -> gather(E_p) -> compute_f
                -> ua_step2 -> deposit -> Poisson -> compute_v

Two observations (exact identities of the reference's formulas, no approximation):

 1. ``gx`` of the second compute_f is ``R(-tau) yt_pred / b`` (ua_steps.F90:174-175): it does not depend on the new
    field E_p.  Hence the corrected x-coefficients ``elt*xf + pl*fx + ql*(gx-fx)/t`` (ua_steps.F90:260), the corrector
    deposit position (compute_rho_m6.F90:74-87) and therefore rho_{n+1} are known BEFORE the predictor Poisson solve.
    Both deposits of a step can be made by the same kernel; one field barrier per step remains.
 2. only the tau* evaluation ``sum_k Yc_k exp(+i l_k t/eps)`` of the corrected y-coefficients is ever used
    (ua_steps.F90:293-303), and it is linear in ``gy``:
        sum_k (Yp_k + qt_k (Gy_k - Fy_k)) conj(elt_k) = [sum_k (Yp_k - qt_k Fy_k) conj(elt_k)] + sum_n gy(tau_n) W_n
    with ``qt = ql/t`` and ``W_n = (1/N) sum_k qt_k conj(elt_k) exp(-i k tau_n)``, a per-particle weight function that
    depends on t only.  The bracket is known in the first kernel; the second kernel needs no tau-FFT at all.

What crosses the barrier per particle-tau: Re xt_pred (2 doubles), yt_pred (2 complex), and optionally W_n (1 complex)
and interv (1 double) = 48 or 72 bytes instead of 128; per particle: t, b and the two bracket sums.
"""
from __future__ import annotations

import numpy as np

from oracle import uapic_oracle_np as onp


def phase_a(mesh, ntau, eps, dt, x, v, ep, e_mesh):
    """everything of one UA step that does not need the predictor field.
    returns dict with: pos_p (2,np) predictor deposit position, pos_c (2,np) corrector deposit position (= new particles.x),
    and the barrier-crossing data xtr (N,2,np) real, ytp (N,2,np) complex, W (N,np) complex, t, b, qa (2,np) complex."""
    N = ntau
    tau, ltau = onp.ua_tables(N)
    b, t, pl, ql, xt, yt = onp.preparation(N, eps, dt, x, v, ep)
    et = np.zeros((N, 2, x.shape[1]))
    onp.interpol_eb_m6_tau(mesh, e_mesh, xt, et)
    fx, fy = onp.compute_f(eps, b, xt, yt, et)
    fx /= N; fy /= N                                       # Fortran normalisation (ua_steps.F90:194-195)
    xf = np.fft.fft(xt, axis=0) / N
    yf = np.fft.fft(yt, axis=0) / N
    elt = np.exp(-1j * ltau[:, None] * t[None, :] / eps)   # (N,np)
    E = elt[:, None, :]
    P = pl[:, None, :]
    Xp = E * xf + P * fx                                   # ua_steps.F90:226 (normalised coefficients of predicted xt)
    Yp = E * yf + P * fy
    xtp = np.fft.ifft(Xp, axis=0) * N                      # time domain (:231)
    ytp = np.fft.ifft(Yp, axis=0) * N
    cE = np.conj(E)
    pos_p = np.real(np.sum(Xp * cE, axis=0))               # compute_rho_m6.F90:74-87
    # corrector for x: gx depends on predicted yt only
    ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
    gx = np.empty_like(fx)
    gx[:, 0, :] = (ct * ytp[:, 0, :] + st * ytp[:, 1, :]) / b
    gx[:, 1, :] = (-st * ytp[:, 0, :] + ct * ytp[:, 1, :]) / b
    gx = np.fft.fft(gx, axis=0) / N
    qt = (ql / t[None, :])[:, None, :]
    w = qt * cE                                            # qt_k conj(elt_k)
    pos_c = pos_p + np.real(np.sum(w * (gx - fx), axis=0))
    qa = np.sum((Yp - qt * fy) * cE, axis=0)               # bracket of observation 2 (complex; only Re is used)
    W = np.fft.fft(w[:, 0, :], axis=0) / N                 # W_n = (1/N) sum_k w_k exp(-i k tau_n)
    return dict(b=b, t=t, pos_p=pos_p, pos_c=pos_c, xtr=np.real(xtp), ytp=ytp, W=W, qa=qa)


def phase_b(mesh, ntau, eps, A, e_pred):
    """v_{n+1} from the barrier-crossing data and the predictor field (ua_steps.F90:160-185 for gy, :293-303)."""
    N = ntau
    tau, _ = onp.ua_tables(N)
    ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
    b, t = A["b"], A["t"]
    xt1, xt2 = A["xtr"][:, 0, :], A["xtr"][:, 1, :]
    yt1, yt2 = A["ytp"][:, 0, :], A["ytp"][:, 1, :]
    et = np.zeros((N, 2, b.shape[0]))
    xt = np.zeros((N, 2, b.shape[0]), dtype=np.complex128)
    xt[:, 0, :], xt[:, 1, :] = xt1, xt2
    onp.interpol_eb_m6_tau(mesh, e_pred, xt, et)
    interv = (1 + 0.5 * np.sin(xt1) * np.sin(xt2) - b) / eps
    tmp1 = et[:, 0, :] + (ct * yt2 - st * yt1) * interv
    tmp2 = et[:, 1, :] + (-ct * yt1 - st * yt2) * interv
    gy1 = (ct * tmp1 - st * tmp2) / b
    gy2 = (st * tmp1 + ct * tmp2) / b
    px = A["qa"][0] + np.sum(gy1 * A["W"], axis=0)
    py = A["qa"][1] + np.sum(gy2 * A["W"], axis=0)
    c, s = np.cos(t / eps), np.sin(t / eps)
    return np.stack([np.real(c * px + s * py), np.real(c * py - s * px)])


def run_bupdate_onepass(mesh, ntau, eps, dt, nstep, x, v, w):
    """same contract as oracle.uapic_oracle_np.run_bupdate, through the reorganised step"""
    x = np.array(x, dtype=np.float64, copy=True)
    v = np.array(v, dtype=np.float64, copy=True)
    nx, ny = mesh.nx, mesh.ny
    rho = np.zeros((nx + 1, ny + 1))
    e = np.zeros((2, nx + 1, ny + 1))
    e_pred = np.zeros((2, nx + 1, ny + 1))
    ep = np.zeros_like(x)
    poisson = onp.Poisson(mesh)
    energy = []
    onp.compute_rho_m6(mesh, rho, x, w)
    energy.append(poisson(rho, e))
    onp.interpol_eb_m6(mesh, e, x, ep)
    for _ in range(nstep):
        A = phase_a(mesh, ntau, eps, dt, x, v, ep, e)
        # the one barrier: both deposits, both Poisson solves
        xp = A["pos_p"].copy()
        onp.compute_rho_m6(mesh, rho, xp, w)
        energy.append(poisson(rho, e_pred))
        x = A["pos_c"].copy()
        onp.compute_rho_m6(mesh, rho, x, w)          # wraps x like src/compute_rho.jl:69-70
        energy.append(poisson(rho, e))
        v = phase_b(mesh, ntau, eps, A, e_pred)
    return x, v, np.array(energy)
