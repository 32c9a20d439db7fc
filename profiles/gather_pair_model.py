"""Model of the two-samples-per-lane M6 gather (round 2): same bank-group model as gather_line_model.py (data-pipe wavefronts
of one LDG.128 = sum over quarter-warps of the max number of distinct 16-byte addresses in the same 16-byte slot of a
128-byte line; it reproduced the measured 5.3 wavefronts per tap of the one-sample-per-lane gather), applied to lanes that
each own two ADJACENT tau samples and load the 7 x 7 union stencil of their two cells once.
Prints wavefronts per LDG.128 and per 32 samples for both mappings, and the share of lanes whose two cells are more than
one cell apart (they take the one-sample fallback for their second sample)."""
import numpy as np
rng = np.random.default_rng(0)
eps = 0.1
npart = 4000
DIMX, DIMY = 4 * np.pi, 2 * np.pi
tau = np.arange(32) * 2 * np.pi / 32
ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
slot = lambda i, j: (i % 2) + 2 * (j % 4)                 # 2 x 4 tiling (uapic_fast.cuh)
addr = lambda i, j: i + 1000 * j


def cost(A, S, group=8):
    """A, S: (32 lanes, npart): per-particle wavefronts of one load instruction"""
    tot = np.zeros(A.shape[1])
    for q in range(32 // group):
        Aq, Sq = A[q * group:(q + 1) * group], S[q * group:(q + 1) * group]
        w = np.zeros(A.shape[1], int)
        for s in range(8):
            m = (Sq == s)
            Am = np.where(m, Aq, -1)
            Am = np.sort(Am, axis=0)
            d = (np.diff(Am, axis=0) != 0).sum(axis=0) + 1 - (Am[0] == -1)
            w = np.maximum(w, np.where(m.any(axis=0), d, 0))
        tot += w
    return tot


for nx, ny, name in ((128, 128, "config 3 (128x128)"), (128, 64, "configs 1/2/4 (128x64)"), (256, 256, "config 5 (256x256)")):
    dx, dy = DIMX / nx, DIMY / ny
    x = rng.random(npart) * DIMX; y = rng.random(npart) * DIMY
    vr = np.sqrt(-2 * np.log(rng.random(npart))); th = rng.random(npart) * 2 * np.pi
    vx, vy = vr * np.cos(th), vr * np.sin(th)
    b = 1 + 0.5 * np.sin(x) * np.sin(y)
    xt1 = x + eps * (st * vx / b - ct * vy / b) + eps * vy / b
    xt2 = y + eps * (st * vy / b + ct * vx / b) - eps * vx / b
    I = np.floor(np.mod(xt1 / dx, nx)).astype(int)
    J = np.floor(np.mod(xt2 / dy, ny)).astype(int)
    # one sample per lane
    one = np.mean([cost(addr(I + a, J + bb), slot(I + a, J + bb)).mean() for a in range(6) for bb in range(6)])
    # two adjacent samples per lane: lanes 0-15 particle p, lanes 16-31 particle p+1 (neighbour in the sorted order ~ same bin:
    # modelled by an independent particle shifted into the same 8x8-cell bin)
    I0, I1, J0, J1 = I[0::2], I[1::2], J[0::2], J[1::2]
    wrapx = np.abs(I0 - I1) > nx // 2; wrapy = np.abs(J0 - J1) > ny // 2
    I0 = np.where(wrapx & (I0 < I1), I0 + nx, I0); I1 = np.where(wrapx & (I1 < I0), I1 + nx, I1)
    J0 = np.where(wrapy & (J0 < J1), J0 + ny, J0); J1 = np.where(wrapy & (J1 < J0), J1 + ny, J1)
    BI, BJ = np.minimum(I0, I1), np.minimum(J0, J1)
    far = (np.abs(I0 - I1) > 1) | (np.abs(J0 - J1) > 1)
    perm = rng.permutation(npart)
    BI2 = np.concatenate([BI, (BI[:, perm] - BI[:1, perm] + BI[:1]) ], axis=0)      # second particle moved next to the first
    BJ2 = np.concatenate([BJ, (BJ[:, perm] - BJ[:1, perm] + BJ[:1]) ], axis=0)
    two = np.mean([cost(addr(BI2 + a, BJ2 + bb), slot(BI2 + a, BJ2 + bb)).mean() for a in range(7) for bb in range(7)])
    per32_one, per32_two = 36 * one, 49 * two / 2
    print(f"{name}: one sample/lane {one:.2f} wavefronts per LDG.128 -> {per32_one:.0f} per 32 samples;  two samples/lane {two:.2f} -> "
          f"{per32_two:.0f} per 32 samples ({100 * (per32_two / per32_one - 1):+.0f} %); far pairs {100 * far.mean():.1f} % of lanes, "
          f"{100 * far.any(axis=0).mean():.1f} % of particles")
