# The reference's test/test_efd.jl carries the whole external-field algorithm inline and ends with `efd(16, 1)` (:483, no
# assertion).  Here the algorithm is the library's kernel (uapic_efd_run); the case keeps the reference's call and adds what
# the program prints: over the full load the two numbers of efd.f90:481 vanish.
@testset "external-field program" begin
    p, f, printed = efd(16, 1)
    @test p.nbpart == 204800
    p, f, printed = efd(16, 204800)
    @test abs(printed[1]) < 1e-8 && abs(printed[2]) < 1e-8
    @test integrate(f.ρ, f.mesh) ≈ 0.0 atol = 1e-9
end
