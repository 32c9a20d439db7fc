"""model: data-pipe wavefronts of one LDG.128 tap of the M6 gather = sum over quarter-warps of the max number of DISTINCT
16-byte addresses that fall into the same 16-byte slot of a 128-byte line (bank-group conflict).  lanes = 32 tau samples of
one particle's first-order orbit.  Compare mesh layouts."""
import numpy as np
rng = np.random.default_rng(0)
nx = ny = 128
DIMX, DIMY = 4*np.pi, 2*np.pi
dx, dy = DIMX/nx, DIMY/ny
eps = 0.1
npart = 4000
x = rng.random(npart)*DIMX; y = rng.random(npart)*DIMY
vr = np.sqrt(-2*np.log(rng.random(npart))); th = rng.random(npart)*2*np.pi
vx, vy = vr*np.cos(th), vr*np.sin(th)
b = 1 + 0.5*np.sin(x)*np.sin(y)
tau = np.arange(32)*2*np.pi/32
ct, st = np.cos(tau)[:, None], np.sin(tau)[:, None]
xt1 = x + eps*(st*vx/b - ct*vy/b) + eps*vy/b
xt2 = y + eps*(st*vy/b + ct*vx/b) - eps*vx/b
I = np.floor(np.mod(xt1/dx, nx)).astype(int)      # (32, np) cell index; halo index = I (node i-2 at I)
J = np.floor(np.mod(xt2/dy, ny)).astype(int)

def wavefronts(slot_fn, addr_fn, group=8):
    """average wavefronts per LDG over taps (a,b) and particles"""
    tot = 0.0; cnt = 0
    for a in range(0, 6, 1):
        for bb in range(0, 6, 2):
            A = addr_fn(I + a, J + bb)          # unique node id
            S = slot_fn(I + a, J + bb)          # slot 0..7
            for q in range(32 // group):
                Aq, Sq = A[q*group:(q+1)*group], S[q*group:(q+1)*group]
                # per particle: max over slots of number of distinct addresses in that slot
                w = np.zeros(npart, int)
                for s in range(8):
                    m = (Sq == s)
                    # count distinct addresses among lanes with slot s
                    Am = np.where(m, Aq, -1)
                    Am.sort(axis=0)
                    d = (np.diff(Am, axis=0) != 0).sum(axis=0) + 1 - (Am[0] == -1)   # distinct values minus the -1 filler
                    d = np.where(m.any(axis=0), d, 0)
                    w = np.maximum(w, d)
                tot += w.sum(); cnt += npart
    return tot / cnt * (32 // group)

LD = lambda L: (lambda i, j: (i + (134 + ((L - 6) % 8)) * j))
print("linear ld=134 (L=6, current):", wavefronts(lambda i, j: (i + 6*j) % 8, LD(6)))
for L in (1, 2, 3, 5, 7):
    print(f"linear L={L}:", wavefronts(lambda i, j, L=L: (i + L*j) % 8, LD(L)))
print("tile 2x4 :", wavefronts(lambda i, j: (i % 2) + 2*(j % 4), LD(6)))
print("tile 4x2 :", wavefronts(lambda i, j: (i % 4) + 4*(j % 2), LD(6)))
print("tile 1x8 (y-major):", wavefronts(lambda i, j: (j % 8), LD(6)))
print("skew: (i%2)+2*((j + (i//2))%4):", wavefronts(lambda i, j: (i % 2) + 2*((j + (i//2)) % 4), LD(6)))
print("skew2: (i + 2*j + 4*(j//4... ) ", wavefronts(lambda i, j: (i + 2*j + ((j//4) % 2)*1 ) % 8, LD(6)))
print("xor: (i ^ (2*j)) % 8:", wavefronts(lambda i, j: ((i) ^ (2*j)) % 8, LD(6)))
print("(i%2) + 2*((j + 2*(i//2))%4):", wavefronts(lambda i, j: (i % 2) + 2*((j + 2*(i//2)) % 4), LD(6)))

def lines(line_fn):
    tot = 0.0; cnt = 0
    for a in range(6):
        for bb in range(6):
            Ln = line_fn(I + a, J + bb)
            Ls = np.sort(Ln, axis=0)
            d = (np.diff(Ls, axis=0) != 0).sum(axis=0) + 1
            tot += d.sum(); cnt += npart
    return tot / cnt
print("---- distinct 128-B lines per tap instruction (32 lanes = one particle's orbit)")
ld = 134
print("x-major 8x1 (current, ld=134):", lines(lambda i, j: (i + ld*j) // 8))
print("x-major 8x1 aligned rows (ld=136):", lines(lambda i, j: (i // 8) + 1000*j))
print("y-major 1x8:", lines(lambda i, j: (j // 8) + 1000*i))
print("tile 2x4:", lines(lambda i, j: (i // 2) + 1000*(j // 4)))
print("tile 4x2:", lines(lambda i, j: (i // 4) + 1000*(j // 2)))
print("tile 2x4 (sectors 32B = 2x1) ...")

def lines_grouped(line_fn, group):
    tot = 0.0; cnt = 0
    for a in range(6):
        for bb in range(6):
            Ln = line_fn(I + a, J + bb)
            for q in range(32 // group):
                Ls = np.sort(Ln[q*group:(q+1)*group], axis=0)
                tot += ((np.diff(Ls, axis=0) != 0).sum(axis=0) + 1).sum()
            cnt += npart
    return tot / cnt
print("---- sum over lane groups of distinct lines in the group")
for grp in (8, 16):
    print(f"group {grp}: x-major", lines_grouped(lambda i, j: (i + ld*j) // 8, grp), " y-major", lines_grouped(lambda i, j: (j // 8) + 1000*i, grp),
          " tile2x4", lines_grouped(lambda i, j: (i // 2) + 1000*(j // 4), grp), " tile4x2", lines_grouped(lambda i, j: (i // 4) + 1000*(j // 2), grp))
