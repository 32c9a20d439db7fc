@testset " Test particles input file read " begin
    mesh = Mesh(0, 4π, 128, 0, 2π, 64)
    # regenerate with: python -c "import uapic_b200 as u; u.make_particles_dat('particles.dat')"
    particles = read_particles("particles.dat", mesh)
    @test particles.nbpart == 204800
end

@testset " Test Particles-MeshFields interaction " begin
    nx, ny = 20, 20
    mesh = Mesh(0.0, 20.0, nx, 0.0, 20.0, ny)
    dx, dy = mesh.dx, mesh.dy
    fields = MeshFields(mesh)
    nbpart = 121
    particles = Particles(nbpart, 1 / nbpart)
    k = 1
    for i = 5:nx-5, j = 5:ny-5
        particles.x[1, k] = (i - 0.5) * dx
        particles.x[2, k] = (j - 0.5) * dx
        k += 1
    end
    compute_rho_m6!(fields, particles)
    @test integrate(fields.ρ, mesh) ≈ 0.0 atol = 1e-4
    for i = 1:nx+1, j = 1:ny+1
        fields.e[1, i, j] = (i - 1) * dx
        fields.e[2, i, j] = (j - 1) * dy
    end
    interpol_eb_m6!(particles, fields)
    err_x = sum(abs.(particles.e[1, :] .- particles.x[1, :])) / nbpart
    err_y = sum(abs.(particles.e[2, :] .- particles.x[2, :])) / nbpart
    @test err_x ≈ 0.0 atol = 1e-6
    @test err_y ≈ 0.0 atol = 1e-6
end
