# Runs the ccall-backed module through the reference's own test cases (reference test/runtests.jl:6-9 lists them).
# Needs a B200 and libuapic_b200.so; Julia is absent from the development image, so the same cases are exercised by
# their Python twins in tests/test_gpu_stages.py, tests/test_gpu_session.py and tests/test_gpu_efd.py.
using Test, UAPIC

const CASES = ("test_efd", "test_poisson", "test_particles", "bupdate")

@testset "UAPIC on libuapic_b200" begin
    for case in CASES
        include(joinpath(@__DIR__, case * ".jl"))
    end
end
