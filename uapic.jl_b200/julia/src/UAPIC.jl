# UAPIC.jl -- drop-in host module for the `bupdate` path of JuliaVlasov/UAPIC.jl on B200.
#
# Same exported names, argument order and mutation semantics as the reference package (src/*.jl), but every
# function on the hot path is a thin `ccall` into libuapic_b200.so (include/uapic_b200.h).  No FFTW: the tau
# FFTs and the Poisson FFTs run inside the library.  There is no CPU fallback; without an sm_100 device every
# call throws `UAPICError`.
#
# NOTE: Julia is not installed in the image this repo is developed in, so this file is exercised only through
# its Python twin (uapic.jl_b200/api.py, same C ABI, same argument order).  It is written against Julia >= 1.6.
#
#   ENV["UAPIC_B200_LIB"] may point at the shared library; default: ../../libuapic_b200.so next to this package.
module UAPIC

export Mesh, MeshFields, Particle, Particles, UA, Poisson
export read_particles, plasma, landau_sampling
export compute_rho_m6!, interpol_eb_m6!, compute_rho_cic!, interpol_eb_cic!
export preparation!, update_particles_e!, update_particles_x!, compute_f!, ua_step!, compute_v!
export fft_tau!, ifft_tau!
export integrate, gnuplot, errors
export Session, upload_particles!, init_fields!, step!, step_host!, generate_particles!, set_sort!, download_particles, download_fields, energy_history
export STORE_FULL, STORE_HYBRID, STORE_ONEPASS, STORE_ONEPASS_LEAN, SCHEME_M6, SCHEME_CIC
export nccl_unique_id, init_nccl!, sum_v, peer_handle, init_peers!, close_peers!
export Mesh3D, Session3D, run_uapic3d!
export efd, efd_run!
export UAPICError

const libuapic = get(ENV, "UAPIC_B200_LIB", joinpath(@__DIR__, "..", "..", "libuapic_b200.so"))

const WRAP_FORTRAN = Cint(0)
const WRAP_JULIA   = Cint(1)
const DEPOSIT_FP64_ATOMIC = Cint(0)
const DEPOSIT_FIXED_POINT = Cint(1)
const SCHEME_M6  = Cint(0)           # quintic spline: what the reference ships
const SCHEME_CIC = Cint(1)           # bilinear, build-defined (include/uapic_b200.h); STORE_ONEPASS_LEAN only
const STORE_FULL         = Cint(0)   # two field barriers per step, 128 B per particle-tau between them
const STORE_HYBRID       = Cint(1)   # ... 16 B, predictor recomputed
const STORE_ONEPASS      = Cint(2)   # one field barrier per step, 72 B per particle-tau across it
const STORE_ONEPASS_LEAN = Cint(3)   # ... 48 B

struct UAPICError <: Exception
    code :: Cint
    msg  :: String
end

Base.showerror(io::IO, e::UAPICError) = print(io, "UAPICError(", e.code, "): ", e.msg)

@inline function check(rc::Cint)
    if rc != 0
        msg = unsafe_string(ccall((:uapic_last_error, libuapic), Cstring, ()))
        throw(UAPICError(rc, msg))
    end
    nothing
end

# ---------------------------------------------------------------------------------------------------
# types (src/meshfields.jl, src/particles.jl, src/ua_type.jl)
# ---------------------------------------------------------------------------------------------------

struct Mesh
    xmin :: Float64
    xmax :: Float64
    nx   :: Int
    dx   :: Float64
    ymin :: Float64
    ymax :: Float64
    ny   :: Int
    dy   :: Float64
    function Mesh(xmin, xmax, nx, ymin, ymax, ny)
        new(xmin, xmax, nx, (xmax - xmin) / nx, ymin, ymax, ny, (ymax - ymin) / ny)
    end
end

# mirror of `uapic_mesh_t`
struct CMesh
    xmin :: Cdouble
    xmax :: Cdouble
    ymin :: Cdouble
    ymax :: Cdouble
    nx   :: Int32
    ny   :: Int32
end
CMesh(m::Mesh) = CMesh(m.xmin, m.xmax, m.ymin, m.ymax, Int32(m.nx), Int32(m.ny))

struct MeshFields
    mesh :: Mesh
    e :: Array{Float64,3}
    ρ :: Array{Float64,2}
    function MeshFields(mesh::Mesh)
        nx, ny = mesh.nx, mesh.ny
        new(mesh, zeros(Float64, (2, nx + 1, ny + 1)), zeros(Float64, (nx + 1, ny + 1)))
    end
end

struct Particle
    x :: Float64
    v :: ComplexF64
    e :: ComplexF64
    b :: Float64
end

mutable struct Particles
    nbpart :: Int64
    x :: Array{Float64,2}
    v :: Array{Float64,2}
    e :: Array{Float64,2}
    b :: Vector{Float64}
    t :: Vector{Float64}
    w :: Float64
    function Particles(nbpart::Int64, w::Float64)
        new(nbpart, zeros(2, nbpart), zeros(2, nbpart), zeros(2, nbpart), zeros(nbpart), zeros(nbpart), w)
    end
end

mutable struct UA
    ntau :: Int64
    ε    :: Float64
    tau  :: Vector{Float64}
    ltau :: Vector{Float64}
    pl   :: Array{ComplexF64,2}
    ql   :: Array{ComplexF64,2}
    function UA(ntau, ε, nbpart)
        ntau in (2, 4, 8, 16, 32) || throw(ArgumentError("ntau must be a power of two in [2, 32]"))
        dtau = 2π / ntau
        ltau = Float64.(vcat(0:ntau÷2-1, -ntau÷2:-1))
        tau  = [i * dtau for i = 0:ntau-1]
        new(ntau, ε, tau, ltau, zeros(ComplexF64, (ntau, nbpart)), zeros(ComplexF64, (ntau, nbpart)))
    end
end

# ---------------------------------------------------------------------------------------------------
# Poisson (src/poisson.jl:62-83)
# ---------------------------------------------------------------------------------------------------

struct Poisson
    mesh :: Mesh
end

function (p::Poisson)(fields::MeshFields)
    nrj = Ref{Cdouble}(0.0)
    check(ccall((:uapic_poisson, libuapic), Cint, (Ref{CMesh}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                CMesh(p.mesh), fields.ρ, fields.e, nrj))
    nrj[]
end

# ---------------------------------------------------------------------------------------------------
# particles <-> mesh (src/compute_rho.jl, src/interpolation.jl)
# ---------------------------------------------------------------------------------------------------

function compute_rho_m6!(fields::MeshFields, particles::Particles; deposit_mode = DEPOSIT_FP64_ATOMIC)
    tot = Ref{Cdouble}(0.0)
    check(ccall((:uapic_compute_rho_m6, libuapic), Cint,
                (Ref{CMesh}, Int64, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cint, Cint, Ref{Cdouble}),
                CMesh(fields.mesh), particles.nbpart, particles.x, particles.w, fields.ρ, WRAP_JULIA, deposit_mode, tot))
    println(" rho_total = $(tot[]) ")       # src/compute_rho.jl:311
    nothing
end

function compute_rho_m6!(fields::MeshFields, particles::Particles, xt::Array{ComplexF64,3}, ua::UA;
                         deposit_mode = DEPOSIT_FP64_ATOMIC)
    tot = Ref{Cdouble}(0.0)
    check(ccall((:uapic_compute_rho_m6_tau, libuapic), Cint,
                (Ref{CMesh}, Cint, Cdouble, Int64, Ptr{ComplexF64}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint, Ref{Cdouble}),
                CMesh(fields.mesh), ua.ntau, ua.ε, particles.nbpart, xt, particles.t, particles.w, fields.ρ, particles.x,
                WRAP_JULIA, deposit_mode, tot))
    nothing
end

function interpol_eb_m6!(particles::Particles, fields::MeshFields)
    check(ccall((:uapic_interpol_eb_m6, libuapic), Cint,
                (Ref{CMesh}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
                CMesh(fields.mesh), fields.e, particles.nbpart, particles.x, particles.e, WRAP_JULIA))
    nothing
end

# the same two stages with the bilinear shape of SCHEME_CIC (build-defined: the reference has no 2D CIC deposit)
function compute_rho_cic!(fields::MeshFields, particles::Particles; deposit_mode = DEPOSIT_FP64_ATOMIC)
    tot = Ref{Cdouble}(0.0)
    check(ccall((:uapic_compute_rho_cic, libuapic), Cint,
                (Ref{CMesh}, Int64, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Cint, Cint, Ref{Cdouble}),
                CMesh(fields.mesh), particles.nbpart, particles.x, particles.w, fields.ρ, WRAP_JULIA, deposit_mode, tot))
    tot[]
end

function interpol_eb_cic!(particles::Particles, fields::MeshFields)
    check(ccall((:uapic_interpol_eb_cic, libuapic), Cint,
                (Ref{CMesh}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Cint),
                CMesh(fields.mesh), fields.e, particles.nbpart, particles.x, particles.e, WRAP_JULIA))
    nothing
end

function interpol_eb_m6!(e::Array{Float64,3}, fields::MeshFields, x::Array{ComplexF64,3}, nbpart::Int64, ntau::Int64)
    check(ccall((:uapic_interpol_eb_m6_tau, libuapic), Cint,
                (Ref{CMesh}, Ptr{Cdouble}, Cint, Int64, Ptr{ComplexF64}, Ptr{Cdouble}, Cint),
                CMesh(fields.mesh), fields.e, ntau, nbpart, x, e, WRAP_JULIA))
    nothing
end

# ---------------------------------------------------------------------------------------------------
# UA stages (src/ua_steps.jl)
# ---------------------------------------------------------------------------------------------------

function preparation!(ua::UA, dt::Float64, particles::Particles, xt::Array{ComplexF64,3}, yt::Array{ComplexF64,3})
    check(ccall((:uapic_preparation, libuapic), Cint,
                (Cint, Cdouble, Cdouble, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                ua.ntau, ua.ε, dt, particles.nbpart, particles.x, particles.v, particles.e, particles.b, particles.t,
                ua.pl, ua.ql, xt, yt))
    nothing
end

update_particles_e!(particles::Particles, et::Array{Float64,3}, fields::MeshFields, ua::UA, xt::Array{ComplexF64,3}) =
    interpol_eb_m6!(et, fields, xt, particles.nbpart, ua.ntau)

update_particles_x!(particles::Particles, fields::MeshFields, ua::UA, xt::Array{ComplexF64,3}) =
    compute_rho_m6!(fields, particles, xt, ua)

function compute_f!(fx::Array{ComplexF64,3}, fy::Array{ComplexF64,3}, ua::UA, particles::Particles,
                    xt::Array{ComplexF64,3}, yt::Array{ComplexF64,3}, et::Array{Float64,3})
    check(ccall((:uapic_compute_f, libuapic), Cint,
                (Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Cdouble}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint),
                ua.ntau, ua.ε, particles.nbpart, particles.b, xt, yt, et, fx, fy, 0))   # 0: unnormalised fft!, src/ua_steps.jl:142-143
    nothing
end

# mul!(x̃t, ftau, xt) of test/bupdate.jl:79,82
function fft_tau!(x̃t::Array{ComplexF64,3}, xt::Array{ComplexF64,3})
    ntau = size(xt, 1)
    check(ccall((:uapic_fft_tau, libuapic), Cint, (Cint, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint),
                ntau, length(xt) ÷ ntau, xt, x̃t, -1, 0))
    nothing
end

# ifft!(xt, 1) of test/bupdate.jl:85-86,102
function ifft_tau!(xt::Array{ComplexF64,3})
    ntau = size(xt, 1)
    check(ccall((:uapic_fft_tau, libuapic), Cint, (Cint, Int64, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint),
                ntau, length(xt) ÷ ntau, xt, xt, 1, 1))
    nothing
end

function ua_step!(xt::Array{ComplexF64,3}, x̃t::Array{ComplexF64,3}, ua::UA, particles::Particles, fx::Array{ComplexF64,3})
    check(ccall((:uapic_ua_step_predict, libuapic), Cint,
                (Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                ua.ntau, ua.ε, particles.nbpart, particles.t, ua.pl, x̃t, fx, xt))
    nothing
end

function ua_step!(xt::Array{ComplexF64,3}, x̃t::Array{ComplexF64,3}, ua::UA, particles::Particles,
                  fx::Array{ComplexF64,3}, gx::Array{ComplexF64,3})
    check(ccall((:uapic_ua_step_correct, libuapic), Cint,
                (Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
                ua.ntau, ua.ε, particles.nbpart, particles.t, ua.pl, ua.ql, x̃t, fx, gx, xt))
    nothing
end

function compute_v!(yt::Array{ComplexF64,3}, particles::Particles, ua::UA)
    check(ccall((:uapic_compute_v, libuapic), Cint,
                (Cint, Cdouble, Int64, Ptr{Cdouble}, Ptr{ComplexF64}, Cint, Ptr{Cdouble}),
                ua.ntau, ua.ε, particles.nbpart, particles.t, yt, 1, particles.v))   # 1: yt holds Fourier coefficients
    nothing
end

# ---------------------------------------------------------------------------------------------------
# loaders and diagnostics (host side, outside the step)
# ---------------------------------------------------------------------------------------------------

function read_particles(filename, mesh::Mesh)                      # src/read_particles.jl:3-35
    nbpart = countlines(filename)
    println(" nbpart   : ", nbpart)
    println(" filename : ", filename)
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    particles = Particles(nbpart, (dimx * dimy) / nbpart)
    open(filename) do f
        for (k, line) in enumerate(eachline(f))
            s = split(line)
            ix, iy = parse(Int32, s[1]), parse(Int32, s[2])
            dpx, dpy = parse(Float64, s[3]), parse(Float64, s[4])
            particles.v[1, k] = parse(Float64, s[5])
            particles.v[2, k] = parse(Float64, s[6])
            particles.x[1, k] = (dpx + ix) * mesh.dx
            particles.x[2, k] = (dpy + iy) * mesh.dy
        end
    end
    particles
end

function plasma(mesh::Mesh, nbpart::Int64)                         # src/plasma.jl:3-52
    kx, alpha = 0.5, 0.05
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    particles = Particles(nbpart, (dimx * dimy) / nbpart)
    k = 1
    while k <= nbpart
        xi, yi, zi = rand() * dimx, rand() * dimy, (2.0 + alpha) * rand()
        if 1.0 + sin(yi) + alpha * cos(kx * xi) >= zi
            particles.x[1, k], particles.x[2, k] = xi, yi
            k += 1
        end
    end
    k = 1
    while k <= nbpart
        xi, yi, zi = (rand() - 0.5) * 10, (rand() - 0.5) * 10, rand()
        temm = (exp(-((xi - 2)^2 + yi^2) / 2) + exp(-((xi + 2)^2 + yi^2) / 2)) / 2
        if temm >= zi
            particles.v[1, k], particles.v[2, k] = xi, yi
            k += 1
        end
    end
    particles
end

# the load src/landau.jl describes (that file references undefined names); pseudo-random instead of Sobol
function landau_sampling(mesh::Mesh, nbpart::Int64; kx = 0.5, alpha = 0.05)
    dimx, dimy = mesh.xmax - mesh.xmin, mesh.ymax - mesh.ymin
    particles = Particles(nbpart, (dimx * dimy) / nbpart)
    function newton(r)
        target = r * 2π / kx
        x0 = target
        for _ = 1:50
            p = x0 + alpha * sin(kx * x0) / kx
            f = 1 + alpha * cos(kx * x0)
            x1 = x0 - (p - target) / f
            done = abs(x1 - x0) <= 1e-12
            x0 = x1
            done && break
        end
        x0
    end
    for i = 1:nbpart
        v = sqrt(-2 * log((i - 0.5) / nbpart))
        r1, r2, r3 = rand(), rand(), rand()
        θ = r1 * 2π
        particles.x[1, i] = mesh.xmin + newton(r2)
        particles.x[2, i] = mesh.ymin + r3 * dimy
        particles.v[1, i] = v * cos(θ)
        particles.v[2, i] = v * sin(θ)
    end
    particles
end

integrate(field::Array{Float64,2}, mesh::Mesh) = sum(view(field, 1:mesh.nx, 1:mesh.ny)) * mesh.dx * mesh.dy

errors(computed::MeshFields, reference::MeshFields) = maximum(abs.(computed.e .- reference.e))

function gnuplot(filename::String, fields::MeshFields)             # src/gnuplot.jl:4-27
    open(filename, "w") do f
        nx, ny, dx, dy = fields.mesh.nx, fields.mesh.ny, fields.mesh.dx, fields.mesh.dy
        for i in 1:nx+1
            for j in 1:ny+1
                write(f, string((i - 1) * dx), "  ", string((j - 1) * dy), "  ", string(fields.e[1, i, j]), "  ",
                      string(fields.e[2, i, j]), "  ", string(fields.ρ[i, j]), "\n")
            end
            write(f, "\n")
        end
    end
end

# ---------------------------------------------------------------------------------------------------
# device-resident session: the whole loop of test/bupdate.jl:63-114 without PCIe round trips
# ---------------------------------------------------------------------------------------------------

struct CConfig        # mirror of `uapic_config_t`
    mesh :: CMesh
    ntau :: Int32
    wrap :: Int32
    deposit_mode :: Int32
    scheme :: Int32
    storage_mode :: Int32
    device :: Int32
    eps :: Cdouble
    dt :: Cdouble
    nbpart :: Int64
    weight :: Cdouble
    total_mass :: Cdouble
    stream :: Ptr{Cvoid}
end

mutable struct Session
    handle :: Ptr{Cvoid}
    mesh   :: Mesh
    nbpart :: Int64
    # storage_mode: STORE_ONEPASS_LEAN (one field barrier per step, 48 B per particle-tau across it) for ntau = 8, 16, 32,
    # STORE_FULL (the literal two-barrier sequence, 128 B) otherwise; see include/uapic_b200.h
    function Session(mesh::Mesh, ntau, ε, dt, nbpart; weight = (mesh.xmax - mesh.xmin) * (mesh.ymax - mesh.ymin) / nbpart,
                     nbpart_global = nbpart, wrap = WRAP_JULIA, deposit_mode = DEPOSIT_FP64_ATOMIC, device = 0,
                     scheme = SCHEME_M6, storage_mode = ntau in (8, 16, 32) ? STORE_ONEPASS_LEAN : STORE_FULL)
        cfg = CConfig(CMesh(mesh), ntau, wrap, deposit_mode, scheme, storage_mode, device, ε, dt, nbpart, weight, weight * nbpart_global, C_NULL)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:uapic_session_create, libuapic), Cint, (Ref{CConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        s = new(h[], mesh, nbpart)
        finalizer(x -> (x.handle != C_NULL && ccall((:uapic_session_destroy, libuapic), Cint, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), s)
        s
    end
end

# ---- several GPUs, one Julia process per GPU ------------------------------------------------------------------------
# The library binds libnccl.so.2 itself (dlopen) and sums the raw ρ meshes with ncclAllReduce on the session's stream; the
# caller only has to hand rank 0's 128-byte id to the other ranks (MPI.jl: `MPI.Bcast!(id, 0, comm)`; or a shared file):
#     id = rank == 0 ? nccl_unique_id() : zeros(UInt8, 128);  MPI.Bcast!(id, 0, comm)
#     s  = Session(mesh, ntau, ε, dt, n_local; nbpart_global = n_global, device = local_rank)
#     init_nccl!(s, id, nranks, rank)
function nccl_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:uapic_nccl_unique_id, libuapic), Cint, (Ptr{UInt8},), id))
    id
end
init_nccl!(s::Session, id::Vector{UInt8}, nranks, rank) =
    (length(id) == 128 || error("the NCCL unique id is 128 bytes");
     check(ccall((:uapic_session_init_nccl, libuapic), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), s.handle, id, nranks, rank)))
# the same exchange without any collective (one node): the field-solve kernel adds the ranks' deposit meshes out of each other's
# memory over NVLink.  handles = the 64-byte handles of all ranks, concatenated in rank order (e.g. MPI.Allgather(peer_handle(s), comm))
function peer_handle(s::Session)
    h = zeros(UInt8, 64)
    check(ccall((:uapic_session_peer_handle, libuapic), Cint, (Ptr{Cvoid}, Ptr{UInt8}), s.handle, h))
    h
end
init_peers!(s::Session, handles::Vector{UInt8}, nranks, rank) =
    check(ccall((:uapic_session_init_peers, libuapic), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), s.handle, handles, nranks, rank))
close_peers!(s::Session) = check(ccall((:uapic_session_close_peers, libuapic), Cint, (Ptr{Cvoid},), s.handle))   # all ranks, after a barrier
# sum(v[1,:]), sum(v[2,:]) of this shard -- what test/bupdate.jl:112 prints every step
function sum_v(s::Session)
    out = zeros(2)
    check(ccall((:uapic_session_sum_v, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), s.handle, out))
    out
end

# reorder the device copies of the particle arrays by coarse mesh bin every `interval` steps (0 = never); uploads and
# downloads stay in the caller's particle order
set_sort!(s::Session, interval, bin_cells_log2 = 3) =
    check(ccall((:uapic_session_set_sort, libuapic), Cint, (Ptr{Cvoid}, Cint, Cint), s.handle, interval, bin_cells_log2))
upload_particles!(s::Session, p::Particles) =
    check(ccall((:uapic_session_upload_particles, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, p.x, p.v))
# one loop iteration of test/bupdate.jl:69-114 for particles kept on the host: copies and kernels pipelined in the library
step_host!(s::Session, p::Particles) =
    check(ccall((:uapic_session_step_host, libuapic), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, p.x, p.v, p.e, p.x, p.v))
# device-side loads: kind 0 = plasma (src/plasma.jl), 1 = Landau (src/landau.jl); shard = global indices first:stride:...
generate_particles!(s::Session, kind; seed = 20190101, first = 0, stride = 1, α = 0.05, kx = 0.5) =
    check(ccall((:uapic_session_generate_particles_strided, libuapic), Cint,
                (Ptr{Cvoid}, Cint, UInt64, Int64, Int64, Cdouble, Cdouble), s.handle, kind, seed, first, stride, α, kx))
init_fields!(s::Session) = check(ccall((:uapic_session_init_fields, libuapic), Cint, (Ptr{Cvoid},), s.handle))
step!(s::Session, nsteps = 1) = check(ccall((:uapic_session_step, libuapic), Cint, (Ptr{Cvoid}, Cint), s.handle, nsteps))

function download_particles(s::Session, p::Particles)
    check(ccall((:uapic_session_download_particles, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, p.x, p.v))
    p
end

function download_fields(s::Session, fields::MeshFields)
    check(ccall((:uapic_session_download_fields, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, fields.e, fields.ρ))
    fields
end

function energy_history(s::Session)
    n = Ref{Int64}(0)
    check(ccall((:uapic_session_energy_history, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{Int64}), s.handle, C_NULL, 0, n))
    out = zeros(n[])
    n[] > 0 && check(ccall((:uapic_session_energy_history, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{Int64}), s.handle, out, n[], n))
    out
end

# ---------------------------------------------------------------------------------------------------
# the 3D program fortran/uapic3d.f90 (rotation push + multi-revolution composition, CIC, 3D spectral Poisson)
# ---------------------------------------------------------------------------------------------------
struct CMesh3D                # uapic3d_mesh_t
    xmin :: NTuple{3,Cdouble}
    xmax :: NTuple{3,Cdouble}
    n    :: NTuple{3,Int32}
end
struct CConfig3D              # uapic3d_config_t
    mesh :: CMesh3D
    nbpart :: Int64
    nbpart_global :: Int64
    weight :: Cdouble
    ep :: Cdouble
    delta :: Cdouble
    deposit_mode :: Int32
    index_quirk :: Int32
    device :: Int32
    stream :: Ptr{Cvoid}
end
struct Mesh3D
    xmin :: NTuple{3,Float64}
    xmax :: NTuple{3,Float64}
    n    :: NTuple{3,Int}
end
CMesh3D(m::Mesh3D) = CMesh3D(m.xmin, m.xmax, Int32.(m.n))

mutable struct Session3D
    handle :: Ptr{Cvoid}
    mesh   :: Mesh3D
    nbpart :: Int64
    function Session3D(mesh::Mesh3D, nbpart; ep = 0.5^10, delta = 3e-3, nbpart_global = nbpart,
                       weight = prod(mesh.xmax .- mesh.xmin) / nbpart_global, deposit_mode = DEPOSIT_FP64_ATOMIC,
                       index_quirk = true, device = 0)
        cfg = CConfig3D(CMesh3D(mesh), nbpart, nbpart_global, weight, ep, delta, deposit_mode, index_quirk ? 1 : 0, device, C_NULL)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:uapic3d_create, libuapic), Cint, (Ref{CConfig3D}, Ref{Ptr{Cvoid}}), cfg, h))
        s = new(h[], mesh, nbpart)
        finalizer(x -> (x.handle != C_NULL && ccall((:uapic3d_destroy, libuapic), Cint, (Ptr{Cvoid},), x.handle); x.handle = C_NULL), s)
        s
    end
end

# x, v :: Array{Float64,2}(3, nbpart); runs uapic3d.f90:74-206 and returns the number of sub-steps; x, v are updated in place
function run_uapic3d!(s::Session3D, x::Array{Float64,2}, v::Array{Float64,2}; Nmrc = 2^7, Nmrcm = 2^7, tfinal = π)
    check(ccall((:uapic3d_upload_particles, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, x, v))
    check(ccall((:uapic3d_init_fields, libuapic), Cint, (Ptr{Cvoid},), s.handle))
    n = Ref{Int64}(0)
    check(ccall((:uapic3d_run, libuapic), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cint, Ref{Int64}), s.handle, Nmrc, Nmrcm, tfinal, 0, n))
    check(ccall((:uapic3d_download_particles, libuapic), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), s.handle, x, v, C_NULL))
    n[]
end

# ---------------------------------------------------------------------------------------------------------------------
# the external-field program (fortran/efd.f90, test/test_efd.jl) -- uapic_efd_run, one kernel for all particles
# ---------------------------------------------------------------------------------------------------------------------
struct EfdConfig            # uapic_efd_config_t
    ntau   :: Int32
    nstep  :: Int32
    eps    :: Float64
    dt     :: Float64
    tfinal :: Float64
    xmin   :: Float64
    xmax   :: Float64
    ymin   :: Float64
    ymax   :: Float64
end

# x, v :: Array{Float64,2}(2, nbpart), advanced in place to tfinal (efd.f90:133-478 over all particles)
function efd_run!(x::Array{Float64,2}, v::Array{Float64,2}, mesh::Mesh; ntau = 16, eps = 1e-3, dt = π/2/2^3, tfinal = π/2, nstep = 0)
    cfg = Ref(EfdConfig(ntau, nstep, eps, dt, tfinal, mesh.xmin, mesh.xmax, mesh.ymin, mesh.ymax))
    check(ccall((:uapic_efd_run, libuapic), Cint, (Ref{EfdConfig}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                cfg, size(x, 2), x, v, x, v))
    x, v
end

# `efd(ntau, nbpart)` of test/test_efd.jl:8: the program with its own constants on the first nbpart particles of
# particles.dat, then the M6 deposit; returns (particles, fields, the pair the program prints against efd.f90:481)
function efd(ntau, nbpart; filename = "particles.dat", eps = 1e-3)
    kx, ky = 0.5, 1.0
    mesh = Mesh(0.0, 2π/kx, 128, 0.0, 2π/ky, 64)
    f = MeshFields(mesh)
    p = read_particles(filename, mesh)
    x, v = p.x[:, 1:nbpart], p.v[:, 1:nbpart]
    efd_run!(x, v, mesh; ntau = ntau, eps = eps)
    p.x[:, 1:nbpart] .= x
    p.v[:, 1:nbpart] .= v
    compute_rho_m6!(f, p)
    p, f, (sum(p.v[1, :]) + 857.95049281063064, sum(p.v[2, :]) + 593.40700170710875)
end

end # module
