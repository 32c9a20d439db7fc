# test/bupdate.jl of the reference, twice: (1) stage by stage through the ccall wrappers, exactly the sequence of
# the reference script; (2) the same run through the device-resident Session.  Both must agree.
@testset "pic2d" begin
    ntau = 16
    kx, ky = 0.50, 1.0
    dimx, dimy = 2π / kx, 2π / ky
    nx, ny = 128, 64
    mesh = Mesh(0.0, dimx, nx, 0.0, dimy, ny)
    dt = π / 2 / (2^3)
    tfinal = π / 2
    nstep = trunc(Int64, tfinal / dt)
    fields = MeshFields(mesh)
    particles = plasma(mesh, 204800)
    x0, v0 = copy(particles.x), copy(particles.v)
    nbpart = particles.nbpart
    poisson! = Poisson(mesh)
    ε = 0.1
    ua = UA(ntau, ε, nbpart)
    et = zeros(Float64, (ntau, 2, nbpart))
    xt = zeros(ComplexF64, (ntau, 2, nbpart)); x̃t = zeros(ComplexF64, (ntau, 2, nbpart))
    yt = zeros(ComplexF64, (ntau, 2, nbpart)); ỹt = zeros(ComplexF64, (ntau, 2, nbpart))
    fx = zeros(ComplexF64, (ntau, 2, nbpart)); fy = zeros(ComplexF64, (ntau, 2, nbpart))
    gx = zeros(ComplexF64, (ntau, 2, nbpart)); gy = zeros(ComplexF64, (ntau, 2, nbpart))
    nrj = Float64[]
    compute_rho_m6!(fields, particles)
    push!(nrj, poisson!(fields))
    interpol_eb_m6!(particles, fields)
    for istep = 1:nstep
        preparation!(ua, dt, particles, xt, yt)
        update_particles_e!(particles, et, fields, ua, xt)
        compute_f!(fx, fy, ua, particles, xt, yt, et)
        fft_tau!(x̃t, xt)
        ua_step!(xt, x̃t, ua, particles, fx)
        fft_tau!(ỹt, yt)
        ua_step!(yt, ỹt, ua, particles, fy)
        ifft_tau!(xt)
        ifft_tau!(yt)
        update_particles_x!(particles, fields, ua, xt)
        push!(nrj, poisson!(fields))
        update_particles_e!(particles, et, fields, ua, xt)
        compute_f!(gx, gy, ua, particles, xt, yt, et)
        ua_step!(xt, x̃t, ua, particles, fx, gx)
        ua_step!(yt, ỹt, ua, particles, fy, gy)
        ifft_tau!(xt)
        update_particles_x!(particles, fields, ua, xt)
        push!(nrj, poisson!(fields))
        compute_v!(yt, particles, ua)
        @show sum(particles.v[1, :]), sum(particles.v[2, :])
    end

    s = Session(mesh, ntau, ε, dt, nbpart)
    p2 = Particles(nbpart, particles.w)
    p2.x .= x0; p2.v .= v0
    upload_particles!(s, p2)
    init_fields!(s)
    step!(s, nstep)
    download_particles(s, p2)
    @test maximum(abs.(p2.v .- particles.v)) < 1e-9
    @test maximum(abs.(energy_history(s) .- nrj)) / maximum(nrj) < 1e-10
end
