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
# ---- round 2: fused phase B-in-A kernel, one-launch field solve vs split, general-ntau kernels, the 3D path ----
with ub.Session(mesh, 32, 0.1, DT, 3001) as s:
    s.set_fusion(True)
    s.generate_particles("landau", seed=3)
    s.init_fields(); s.step(5); s.synchronize()
    x, v = s.download_particles()
    print("fused", s.energy_history()[-1], s.sum_v())
for ntau in (12, 64):
    with ub.Session(mesh, ntau, 0.1, DT, 1001) as s:
        s.generate_particles("plasma", seed=4)
        s.init_fields(); s.step(2); s.synchronize()
        print("general ntau", ntau, s.energy_history()[-1])
m3 = ub.Mesh3D((0, 0, 0), (18, 18, 1), (16, 12, 4))
with ub.Session3D(m3, 3001) as s3:
    s3.generate_particles(seed=5)
    s3.init_fields()
    print("uapic3d substeps", s3.run(4, 4, np.pi, 1), s3.run(8, 4, 0.05, 2))
    x, v, ep = s3.download_particles()
f3 = ub.Fields3D(m3)
ub.mrc3d.compute_rho_cic(f3, x, 0.1); ub.mrc3d.solve_poisson(f3); ub.mrc3d.interpolate_eb_cic(x, f3)
print("uapic3d stages ok")
