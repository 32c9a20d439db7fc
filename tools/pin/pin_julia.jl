# pin_julia.jl -- run the REFERENCE's own Julia package (from its checkout, nothing copied) on the same binary inputs as
# pin_driver.F90 and write the same binary output.  The loop body is test/bupdate.jl:63-114.
#     julia --project=$REF tools/pin/pin_julia.jl $REF in.bin out.bin        (needs FFTW.jl, Sobol.jl as the reference does)
ref, fin, fout = ARGS
using UAPIC, FFTW, LinearAlgebra          # UAPIC = the reference package itself (--project=$REF)

io = open(fin)
nbpart = read(io, Int64); nx, ny, ntau, nstep = read(io, Int32), read(io, Int32), read(io, Int32), read(io, Int32)
eps, dt, xmax, ymax, w = [read(io, Float64) for _ = 1:5]
x = Array{Float64}(undef, 2, nbpart); read!(io, x)
v = Array{Float64}(undef, 2, nbpart); read!(io, v)
close(io)

mesh = Mesh(0.0, xmax, Int(nx), 0.0, ymax, Int(ny))
fields = MeshFields(mesh)
particles = Particles(Int(nbpart), w)
particles.x .= x; particles.v .= v
poisson! = Poisson(mesh)
ua = UA(Int(ntau), eps, Int(nbpart))
energy = Float64[]
et = zeros(Float64, (ntau, 2, nbpart))
xt, x̃t, yt, ỹt, fx, fy, gx, gy = [zeros(ComplexF64, (ntau, 2, nbpart)) for _ = 1:8]
ftau = plan_fft(xt, 1)

compute_rho_m6!(fields, particles)
push!(energy, poisson!(fields))
interpol_eb_m6!(particles, fields)
for istep = 1:nstep
    preparation!(ua, dt, particles, xt, yt)
    update_particles_e!(particles, et, fields, ua, xt)
    compute_f!(fx, fy, ua, particles, xt, yt, et)
    mul!(x̃t, ftau, xt); ua_step!(xt, x̃t, ua, particles, fx)
    mul!(ỹt, ftau, yt); ua_step!(yt, ỹt, ua, particles, fy)
    ifft!(xt, 1); ifft!(yt, 1)
    update_particles_x!(particles, fields, ua, xt)
    push!(energy, poisson!(fields))
    update_particles_e!(particles, et, fields, ua, xt)
    compute_f!(gx, gy, ua, particles, xt, yt, et)
    ua_step!(xt, x̃t, ua, particles, fx, gx)
    ua_step!(yt, ỹt, ua, particles, fy, gy)
    ifft!(xt, 1)
    update_particles_x!(particles, fields, ua, xt)
    push!(energy, poisson!(fields))
    compute_v!(yt, particles, ua)
end
open(fout, "w") do o
    write(o, Int64(nbpart), Int32(nx), Int32(ny), Int32(ntau), Int32(nstep))
    write(o, particles.x); write(o, particles.v); write(o, energy); write(o, fields.e)
end
