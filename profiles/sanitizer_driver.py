import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uapic_b200 as ub
DT = np.pi / 16
DIMX, DIMY = 4 * np.pi, 2 * np.pi
for ntau, scheme in ((32, ub.SCHEME_M6), (16, ub.SCHEME_M6), (8, ub.SCHEME_CIC)):
    mesh = ub.Mesh(0, DIMX, 64, 0, DIMY, 32)
    with ub.Session(mesh, ntau, 0.1, DT, 3001, scheme=scheme) as s:
        s.generate_particles("landau", seed=1)
        s.init_fields(); s.step(2); s.synchronize()
        x, v = s.download_particles()
        print(ntau, scheme, s.energy_history()[-1], float(np.abs(v).max()))
# host-resident stepping (chunked, copy streams) and the two-barrier kernels
mesh = ub.Mesh(0, DIMX, 64, 0, DIMY, 32)
with ub.Session(mesh, 32, 0.1, DT, 70001) as s:
    s.generate_particles("plasma", seed=2)
    s.init_fields()
    x, v = s.download_particles(); e = s.download_particle_e()
    s.step_host(x, v, e, x, v); s.step_host(x, v, None, x, v)
    print("step_host", s.energy_history()[-1])
for mode in (ub.STORE_FULL, ub.STORE_HYBRID):
    with ub.Session(mesh, 16, 0.1, DT, 3001, storage_mode=mode) as s:
        s.generate_particles("plasma", seed=2); s.set_sort(1, 3)
        s.init_fields(); s.step(2); s.synchronize()
        print("legacy", mode, s.energy_history()[-1])
