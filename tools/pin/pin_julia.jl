# pin_julia.jl -- run the REFERENCE's own Julia package (from its checkout, nothing copied) on the same binary inputs as
# pin_driver.F90 and write the same binary output.  The loop body is test/bupdate.jl:63-114.
#     julia --project=$REF tools/pin/pin_julia.jl $REF in.bin out.bin        (needs FFTW.jl, Sobol.jl as the reference does)
ref, fin, fout = ARGS
using UAPIC, FFTW, LinearAlgebra          # UAPIC = the reference package itself (--project=$REF)

io = open(fin)
npt = read(io, Int64); nx, ny, nta, nst = read(io, Int32), read(io, Int32), read(io, Int32), read(io, Int32)
eps, dt, xmax, ymax, w = [read(io, Float64) for _ = 1:5]
x = Array{Float64}(undef, 2, npt); read!(io, x)
v = Array{Float64}(undef, 2, npt); read!(io, v)
close(io)

mesh = Mesh(0.0, xmax, Int(nx), 0.0, ymax, Int(ny))
fld = MeshFields(mesh)
pt = Particles(Int(npt), w)
pt.x .= x; pt.v .= v
solve! = Poisson(mesh)
uat = UA(Int(nta), eps, Int(npt))
energy = Float64[]
etau = zeros(Float64, (nta, 2, npt))
xtau, xhat, ytau, yhat, fxh, fyh, gxh, gyh = [zeros(ComplexF64, (nta, 2, npt)) for _ = 1:8]
planτ = plan_fft(xtau, 1)

compute_rho_m6!(fld, pt)
push!(energy, solve!(fld))
interpol_eb_m6!(pt, fld)
for it = 1:nst
    preparation!(uat, dt, pt, xtau, ytau)
    update_particles_e!(pt, etau, fld, uat, xtau)
    compute_f!(fxh, fyh, uat, pt, xtau, ytau, etau)
    mul!(xhat, planτ, xtau); ua_step!(xtau, xhat, uat, pt, fxh)
    mul!(yhat, planτ, ytau); ua_step!(ytau, yhat, uat, pt, fyh)
    ifft!(xtau, 1); ifft!(ytau, 1)
    update_particles_x!(pt, fld, uat, xtau)
    push!(energy, solve!(fld))
    update_particles_e!(pt, etau, fld, uat, xtau)
    compute_f!(gxh, gyh, uat, pt, xtau, ytau, etau)
    ua_step!(xtau, xhat, uat, pt, fxh, gxh)
    ua_step!(ytau, yhat, uat, pt, fyh, gyh)
    ifft!(xtau, 1)
    update_particles_x!(pt, fld, uat, xtau)
    push!(energy, solve!(fld))
    compute_v!(ytau, pt, uat)
end
open(fout, "w") do o
    write(o, Int64(npt), Int32(nx), Int32(ny), Int32(nta), Int32(nst))
    write(o, pt.x); write(o, pt.v); write(o, energy); write(o, fld.e)
end
