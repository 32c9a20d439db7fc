@testset "Poisson 2D on rectangular grid" begin
    kx, ky = 0.5, 1.0
    mesh = Mesh(0, 2π / kx, 64, 0, 2π / ky, 128)
    fields = MeshFields(mesh)
    solutions = MeshFields(mesh)
    x = LinRange(mesh.xmin, mesh.xmax, mesh.nx + 1)
    y = LinRange(mesh.ymin, mesh.ymax, mesh.ny + 1)
    fields.ρ .= -8 .* sin.(2 .* x) .* cos.(2 .* y')
    solutions.e[1, :, :] .= 2 .* cos.(2 .* x) .* cos.(2 .* y')
    solutions.e[2, :, :] .= -2 .* sin.(2 .* x) .* sin.(2 .* y')
    poisson! = Poisson(mesh)
    poisson!(fields)
    @test errors(fields, solutions) ≈ 0.0 atol = 1e-14
    fields.ρ .= -4 * (sin.(2 * x) .+ cos.(2 * y'))
    poisson!(fields)
    for j in 1:mesh.ny+1, i in 1:mesh.nx+1
        solutions.e[1, i, j] = 2 * cos(2 * x[i])
        solutions.e[2, i, j] = -2 * sin(2 * y[j])
    end
    @test errors(fields, solutions) ≈ 0.0 atol = 1e-14
end
