# The reference's test files (test/runtests.jl:6-9) against the ccall-backed module.  test_efd.jl is out of scope
# (external-E variant, no assertions).  Needs a B200 and libuapic_b200.so; Julia is absent from the development image,
# so these are exercised through their Python twins in tests/test_gpu_stages.py and tests/test_gpu_session.py.
using Test
using UAPIC

include("test_poisson.jl")
include("test_particles.jl")
include("bupdate.jl")
