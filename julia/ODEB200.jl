# ODEB200.jl -- Julia host shim over libode_b200.so (include/ode_b200.h).
#
# Keeps ODE.jl's surface for the RK / ode23s hot path:
#     tout, yout = ode45(F, y0, tspan; reltol, abstol, minstep, maxstep, initstep, points)
#     sol = solve(prob, ode45(); reltol, abstol, saveat, ...)        (see solve_b200 below)
# with `F` a registered DeviceRHS token instead of a closure (a Julia closure cannot run on the
# device).  A DeviceRHS is also callable, `F(t, y)`, and evaluates the same expression in Julia,
# so the unmodified ODE.jl solvers accept it for reference runs.
#
# NOTE: written against the C ABI but NOT executed in the build container (Julia is not
# installed there, SURVEY.md F2).  The Python mirror ode.jl_b200/api.py is the tested twin; the
# two files marshal the same structs in the same order (tests/test_abi.py checks the Julia struct
# declarations below against the C header field by field).
#
# Like the reference (src/algorithm_types.jl), every solver name is a TYPE whose constructor-call
# with arguments is the solver function and whose instance is the algorithm marker:
#     tout, yout = ode45(F, y0, tspan; kw...)          # src/runge_kutta.jl:218-224
#     sol = solve(ODEProblem(F, u0, tspan), ode45())   # src/common.jl:1-123, when DiffEqBase is loaded
module ODEB200

using Libdl

# DiffEqBase is optional: with it the algorithm types join its hierarchy and `DiffEqBase.solve(prob, alg)`
# gets a method for them (the end of this file); without it only the odeXX / solve_b200 surface exists.
const HAVE_DIFFEQBASE = try
    @eval import DiffEqBase
    true
catch
    false
end
@static if HAVE_DIFFEQBASE
    abstract type ODEB200Algorithm <: DiffEqBase.AbstractODEAlgorithm end
else
    abstract type ODEB200Algorithm end
end

const LIB = Ref{String}(get(ENV, "ODE_B200_LIB", joinpath(@__DIR__, "..", "ode.jl_b200", "libode_b200.so")))

# ---- ids (include/ode_b200.h) -------------------------------------------------------------------
const METHODS = Dict(:ode1=>0, :ode2_midpoint=>1, :ode2_heun=>2, :ode4=>3, :ode21=>4, :ode23=>5,
                     :ode45_fe=>6, :ode45_dp=>7, :ode45=>7, :ode78=>8, :ode23s=>9,
                     :ode4s=>10, :ode4s_s=>10, :ode4s_kr=>11, :ode4ms=>12, :ode5ms=>13,
                     :ode_ms1=>14, :ode_ms2=>15, :ode_ms3=>16, :ode_ms4=>12, :ode_ms5=>13,
                     :ode45_ck=>17)  # ode45_ck: Cash-Karp, an extension (README.md:57; no tableau in the reference source)
const RHS = Dict(:linear=>0, :const=>1, :timelin=>2, :rotation=>3, :lorenz=>4, :vdp=>5,
                 :robertson=>6, :kepler=>7, :reaction_diffusion=>8, :reaction_diffusion_field=>9)
const POINTS = Dict(:all=>0, :specified=>1, :final=>2)

struct DeviceRHS{Name}
    params::Vector{Float64}
end
DeviceRHS(name::Symbol, params...) = DeviceRHS{name}(Float64[params...])
rhs_id(::DeviceRHS{N}) where {N} = RHS[N]

# host evaluation: the same expressions, same association, as csrc/rhs.cuh
(f::DeviceRHS{:linear})(t, y) = f.params[1]*y
(f::DeviceRHS{:const})(t, y) = f.params[1]
(f::DeviceRHS{:timelin})(t, y) = f.params[1]*t
(f::DeviceRHS{:rotation})(t, y) = [-y[2]; y[1]]
(f::DeviceRHS{:lorenz})(t, y) = (p = f.params; [p[1]*(y[2]-y[1]); y[1]*(p[2]-y[3])-y[2]; y[1]*y[2]-p[3]*y[3]])
(f::DeviceRHS{:vdp})(t, y) = [y[2]; f.params[1]*((1.0-y[1]*y[1])*y[2]) - y[1]]
function (f::DeviceRHS{:robertson})(t, y)
    p = f.params
    d1 = -p[1]*y[1] + p[3]*y[2]*y[3]
    d3 = p[2]*y[2]*y[2]
    [d1; -d1 - d3; d3]
end
function (f::DeviceRHS{:kepler})(t, y)
    r2 = y[1]*y[1] + y[2]*y[2]; r3 = r2*sqrt(r2)
    [y[3]; y[4]; -y[1]/r3; -y[2]/r3]
end

# DiffEq-style signatures, so that ODEProblem(DeviceRHS(:lorenz, 10.0, 28.0, 8/3), u0, tspan) works with the stock
# CPU solvers too: f(u, p, t) and f(du, u, p, t) (the problem's `p` is ignored: the token carries its parameters)
(f::DeviceRHS)(u, p, t) = f(t, u)
(f::DeviceRHS)(du, u, p, t) = (du .= f(t, u); nothing)

# ---- right-hand sides registered at run time (CUDA C++ body, compiled by NVRTC inside the library) ------
struct UserRHS
    id::Int32
    params::Vector{Float64}
end
rhs_id(f::UserRHS) = f.id
"""register_rhs("brusselator", 2, 2, "f[0] = ...; f[1] = ...;"; jac = nothing, params = [1.0, 3.0])"""
function register_rhs(name::String, dof::Integer, n_params::Integer, eval_body::String; jac = nothing, params = Float64[])
    id = ccall((:ode_b200_register_rhs, LIB[]), Cint, (Cstring, Int32, Int32, Cstring, Cstring),
               name, dof, n_params, eval_body, jac === nothing ? C_NULL : jac)
    id < 0 && error("libode_b200: " * last_error())
    UserRHS(id, Float64[params...])
end

# ---- ode_b200_opts (field order and types as in the header) -----------------------------------------
struct Opts
    method::Int32; rhs::Int32; dof::Int32; n_params::Int32
    params_stride::Int32; points::Int32; jacobian::Int32; device::Int32
    max_steps::Int64
    reltol::Float64; abstol::Float64; minstep::Float64; maxstep::Float64; initstep::Float64
    mathmode::Int32
    host_output::Int32
    n_gpus::Int32      # fan-out inside the library over devices device .. device + n_gpus - 1
    fp_mode::Int32     # 0 = strict (reference parity), 1 = FMA contraction (opt-in)
    reserved::NTuple{4,Int32}
end
"""Opts with the tail fields spelled once (mathmode = 0, host_output = 0)."""
make_opts(method, rhs, dof, n_params, params_stride, points, jacobian, device, max_steps, reltol, abstol, minstep,
          maxstep, initstep; n_gpus = 1, strict_fp = true) =
    Opts(method, rhs, dof, n_params, params_stride, points, jacobian, device, max_steps, reltol, abstol, minstep,
         maxstep, initstep, 0, 0, n_gpus, strict_fp ? 0 : 1, ntuple(_ -> Int32(0), 4))

last_error() = unsafe_string(ccall((:ode_b200_last_error, LIB[]), Cstring, ()))

function check(rc::Integer)
    rc == 0 && return
    msg = last_error()
    rc == -2 && error("Zero time span")                 # src/ODE.jl:88
    rc == -3 && error("initstep has wrong sign.")       # src/runge_kutta.jl:294
    rc == -5 && error("Can only use this solver with an adaptive RK Butcher table")  # :249
    error("libode_b200 error $rc: $msg")
end

function _solve(method::Symbol, F::Union{DeviceRHS,UserRHS}, y0, tspan;
                reltol = 1.0e-5, abstol = 1.0e-8,
                minstep = abs(tspan[end] - tspan[1])/1e18,     # src/runge_kutta.jl:235
                maxstep = abs(tspan[end] - tspan[1])/2.5,      # :236
                initstep = 0.0, points = :all, jacobian = nothing, device = 0)
    haskey(POINTS, points) || error("Unrecognized option points==$points")   # :289
    scalar = !(y0 isa AbstractVector)
    y = scalar ? Float64[y0] : convert(Vector{Float64}, y0)
    ts = convert(Vector{Float64}, collect(tspan))
    # jacobian = true selects the registered analytic Jacobian of F (a closure cannot run on the device)
    o = Ref(make_opts(METHODS[method], rhs_id(F), length(y), length(F.params), 0, POINTS[points],
                      jacobian === true ? 1 : 0, device, 0, reltol, abstol, minstep, maxstep, initstep))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve y ts F begin
        check(ccall((:ode_b200_solve_single, LIB[]), Cint,
                    (Ref{Opts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
                    o, y, F.params, ts, length(ts), h))
    end
    n = ccall((:ode_b200_result_len, LIB[]), Int64, (Ptr{Cvoid},), h[])
    tout = Vector{Float64}(undef, n)
    yflat = Matrix{Float64}(undef, length(y), n)       # column = one output state (row-major [len][dof] in C)
    check(ccall((:ode_b200_result_copy, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h[], tout, yflat))
    status = ccall((:ode_b200_result_status, LIB[]), Int32, (Ptr{Cvoid},), h[])
    ccall((:ode_b200_result_free, LIB[]), Cvoid, (Ptr{Cvoid},), h[])
    status == 1 && method != :ode23s && println("Warning: dt < minstep.  Stopping.")   # :359
    yout = scalar ? vec(yflat) : [yflat[:, i] for i in 1:n]
    return tout, yout
end

for name in (:ode1, :ode2_midpoint, :ode2_heun, :ode4, :ode21, :ode23, :ode45_fe, :ode45_dp, :ode45, :ode45_ck, :ode78, :ode23s,
             :ode4s, :ode4s_s, :ode4s_kr, :ode4ms, :ode5ms)
    # the solver name is a type (its instance is the algorithm marker, src/algorithm_types.jl); calling the type with
    # (F, y0, tspan) is the solver function, exactly the reference's `AlgType(f, u0, Ts; ...)` at src/common.jl:93
    @eval struct $name <: ODEB200Algorithm end
    @eval $name(F::Union{DeviceRHS,UserRHS}, y0, tspan; kw...) = _solve($(QuoteNode(name)), F, y0, tspan; kw...)
    @eval $name(F, y0, tspan; kw...) =
        error("libode_b200 takes a registered DeviceRHS; a Julia closure cannot run on the device (no CPU fallback)")
    @eval alg_symbol(::$name) = $(QuoteNode(name))
    @eval export $name
end
"""ode_ms(F, x0, tspan, order): fixed-step Adams-Bashforth of order 1..5 (src/ODE.jl:175)."""
ode_ms(F::Union{DeviceRHS,UserRHS}, x0, tspan, order::Integer; kw...) =
    (1 <= order <= 5 || error("ode_ms is built for order 1..5"); _solve(Symbol("ode_ms", order), F, x0, tspan))
export DeviceRHS, UserRHS, register_rhs, ode_ms

# ---- ensembles: n independent odeXX calls in one launch --------------------------------------------
struct EnsembleIO
    n_traj::Int64; y0::Ptr{Float64}; params::Ptr{Float64}; tspan::Ptr{Float64}; n_tspan::Int32; pad0::Int32
    t_final::Ptr{Float64}; y_final::Ptr{Float64}; n_accept::Ptr{Int32}; n_reject::Ptr{Int32}; n_rhs::Ptr{Int32}
    status::Ptr{Int32}; pts_t::Ptr{Float64}; pts_y::Ptr{Float64}; pts_n::Ptr{Int32}; pts_cap::Int64
    trace_traj::Int64; trace::Ptr{Float64}; trace_cap::Int64; trace_n::Ptr{Int64}
end

"""solve_ensemble(:ode45, F, Y0, tspan; kw...) with Y0 a dof x n_traj matrix (one trajectory per column)."""
function solve_ensemble(method::Symbol, F::DeviceRHS, Y0::Matrix{Float64}, tspan;
                        reltol = 1.0e-5, abstol = 1.0e-8, minstep = abs(tspan[end]-tspan[1])/1e18,
                        maxstep = abs(tspan[end]-tspan[1])/2.5, initstep = 0.0, device = 0, n_gpus = 1, strict_fp = true)
    dof, n = size(Y0)
    ts = convert(Vector{Float64}, collect(tspan))
    tf = Vector{Float64}(undef, n); yf = Matrix{Float64}(undef, dof, n)
    na = Vector{Int32}(undef, n); nr = similar(na); nf = similar(na); st = similar(na)
    # n_gpus > 1: the library shards the trajectories over that many devices (one host thread each, no collective)
    o = Ref(make_opts(METHODS[method], rhs_id(F), dof, length(F.params), 0, POINTS[:final], 0, device, 0,
                      reltol, abstol, minstep, maxstep, initstep; n_gpus = n_gpus, strict_fp = strict_fp))
    GC.@preserve Y0 ts F tf yf na nr nf st begin
        io = Ref(EnsembleIO(n, pointer(Y0), pointer(F.params), pointer(ts), length(ts), 0,
                            pointer(tf), pointer(yf), pointer(na), pointer(nr), pointer(nf), pointer(st),
                            C_NULL, C_NULL, C_NULL, 0, -1, C_NULL, 0, C_NULL))
        check(ccall((:ode_b200_solve_ensemble, LIB[]), Cint, (Ref{Opts}, Ref{EnsembleIO}), o, io))
    end
    return (t_final = tf, y_final = yf, n_accept = na, n_reject = nr, n_rhs = nf, status = st)
end

# ---- common interface: what src/common.jl:1-123 does, minus DiffEqBase types --------------------------
"""solve_b200(f::DeviceRHS, u0, tspan, alg::Symbol; ...) mirrors ODE.solve's keyword mapping
(dtmin = span/1e-9 (sic, src/common.jl:13), dtmax = span/2.5, saveat -> Ts, save_everystep -> points)."""
function solve_b200(f::DeviceRHS, u0, tspan::Tuple, alg::Symbol; saveat = Float64[], reltol = 1e-5, abstol = 1e-8,
                    save_everystep = isempty(saveat), dtmin = abs(tspan[2]-tspan[1])/1e-9,
                    dtmax = abs(tspan[2]-tspan[1])/2.5, dt = 0.0)
    sv = saveat isa Number ? collect(tspan[1]+saveat:saveat:tspan[2]) : collect(Float64, saveat)
    !isempty(sv) && sv[end] == tspan[2] && pop!(sv)
    Ts = (!isempty(sv) && sv[1] == tspan[1]) ? unique([sv; tspan[2]]) : unique([tspan[1]; sv; tspan[2]])
    points = save_everystep ? :all : :specified
    ts, us = _solve(alg, f, u0 isa AbstractArray ? vec(u0) : u0, Ts; abstol = abstol, reltol = reltol,
                    maxstep = dtmax, minstep = dtmin, initstep = dt, points = points)
    return (t = ts, u = us, retcode = :Succss)   # src/common.jl:122
end

# ---- one large system (dof = N): registered stencils / fields, include/ode_b200.h -----------------------
struct LargeStats
    n_accept::Int64; n_reject::Int64; n_rhs::Int64; status::Int32; pad::Int32; t_final::Float64; kernel_ms::Float64
end
"""register_stencil("lap4", "return p[0]*((((-w[0]+16.0*w[1])-30.0*w[2])+16.0*w[3])-w[4])/12.0;"; radius = 2, params = [0.1])
`point(w, p, t, i, n)`: w[k] = u[i-radius+k] on a periodic ring."""
function register_stencil(name::String, body::String; radius::Integer = 1, params = Float64[])
    id = ccall((:ode_b200_register_stencil, LIB[]), Cint, (Cstring, Int32, Int32, Cstring), name, radius, length(params), body)
    id < 0 && error("libode_b200: " * last_error())
    UserRHS(id, Float64[params...])
end
"""register_field("mirror", "return p[0]*(y[n-1-i] - y[i]);"; params = [0.5]): a general rhs(i, n, y, p, t)."""
function register_field(name::String, body::String; params = Float64[])
    id = ccall((:ode_b200_register_field, LIB[]), Cint, (Cstring, Int32, Cstring), name, length(params), body)
    id < 0 && error("libode_b200: " * last_error())
    UserRHS(id, Float64[params...])
end
"""solve_large(:ode45, F, u0, (t0, tend); kw...) -> (t_final, u_final, stats): oderk_adapt on a dof = N state with the
built-in DeviceRHS(:reaction_diffusion, kappa, r), a registered stencil (fused kernel) or field (kernel per stage)."""
function solve_large(method::Symbol, F::Union{DeviceRHS,UserRHS}, u0::Vector{Float64}, tspan;
                     reltol = 1.0e-5, abstol = 1.0e-8, minstep = abs(tspan[end]-tspan[1])/1e18,
                     maxstep = abs(tspan[end]-tspan[1])/2.5, initstep = 0.0, device = 0, n_gpus = 1)
    n = length(u0)
    out = similar(u0)
    p = isempty(F.params) ? Float64[0.0] : F.params
    st = Ref(LargeStats(0, 0, 0, 0, 0, 0.0, 0.0))
    # n_gpus > 1: the grid is cut into that many slabs, one per device, which exchange one record per step attempt
    # over NVLink peer memory from inside the attempt kernel (ode_b200_solve_large_multi); same bits as one GPU
    o = Ref(make_opts(METHODS[method], rhs_id(F), min(n, typemax(Int32)), length(F.params), 0, POINTS[:final], 0, device, 0,
                      reltol, abstol, minstep, maxstep, initstep; n_gpus = n_gpus))
    GC.@preserve u0 out p begin
        check(ccall((:ode_b200_solve_large, LIB[]), Cint,
                    (Ref{Opts}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}, Ref{LargeStats},
                     Ptr{Float64}, Int64, Ptr{Int64}),
                    o, n, u0, p, Float64(tspan[1]), Float64(tspan[end]), out, st, C_NULL, 0, C_NULL))
    end
    st[].status == 1 && println("Warning: dt < minstep.  Stopping.")   # src/runge_kutta.jl:359
    return st[].t_final, out, st[]
end
"""solve_large_fixed(:ode4, F, u0, tspan) -> (tspan, [u(t) for t in tspan]): oderk_fixed / ode4ms on a dof = N state."""
function solve_large_fixed(method::Symbol, F::Union{DeviceRHS,UserRHS}, u0::Vector{Float64}, tspan; device = 0)
    n = length(u0)
    ts = convert(Vector{Float64}, collect(tspan))
    rows = Matrix{Float64}(undef, n, length(ts))      # column = one state (row-major [n_tspan][n] in C)
    p = isempty(F.params) ? Float64[0.0] : F.params
    st = Ref(LargeStats(0, 0, 0, 0, 0, 0.0, 0.0))
    o = Ref(make_opts(METHODS[method], rhs_id(F), min(n, typemax(Int32)), length(F.params), 0, POINTS[:all], 0, device, 0,
                      1.0e-5, 1.0e-8, 0.0, 0.0, 0.0))
    GC.@preserve u0 ts rows p begin
        check(ccall((:ode_b200_solve_large_fixed, LIB[]), Cint,
                    (Ref{Opts}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ref{LargeStats}),
                    o, n, u0, p, ts, length(ts), rows, C_NULL, st))
    end
    return ts, [rows[:, i] for i in 1:length(ts)]
end
export register_stencil, register_field, solve_large, solve_large_fixed

"""solve_large_multi(:ode45, F, u0, (t0, tend), devices; kw...): the same on an explicit list of CUDA devices."""
function solve_large_multi(method::Symbol, F::Union{DeviceRHS,UserRHS}, u0::Vector{Float64}, tspan, devices::Vector{<:Integer};
                           reltol = 1.0e-5, abstol = 1.0e-8, minstep = abs(tspan[end]-tspan[1])/1e18,
                           maxstep = abs(tspan[end]-tspan[1])/2.5, initstep = 0.0)
    n = length(u0)
    out = similar(u0)
    p = isempty(F.params) ? Float64[0.0] : F.params
    devs = convert(Vector{Int32}, devices)
    st = Ref(LargeStats(0, 0, 0, 0, 0, 0.0, 0.0))
    o = Ref(make_opts(METHODS[method], rhs_id(F), min(n, typemax(Int32)), length(F.params), 0, POINTS[:final], 0, devs[1], 0,
                      reltol, abstol, minstep, maxstep, initstep))
    GC.@preserve u0 out p devs begin
        check(ccall((:ode_b200_solve_large_multi, LIB[]), Cint,
                    (Ref{Opts}, Int32, Ptr{Int32}, Int64, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64},
                     Ref{LargeStats}, Ptr{Float64}, Int64, Ptr{Int64}),
                    o, length(devs), devs, n, u0, p, Float64(tspan[1]), Float64(tspan[end]), out, st, C_NULL, 0, C_NULL))
    end
    st[].status == 1 && println("Warning: dt < minstep.  Stopping.")
    return st[].t_final, out, st[]
end
export solve_large_multi

# ---- slab sessions (one process per GPU, e.g. under MPI.jl): device pointers, the caller's stream ----------------
# The per-rank flow with the exchange over peer memory:
#     s = slab_create(:ode45, F, n_local, n_global, rank, world, (t0, tend); device = local_rank)
#     h = slab_peer_alloc(s)                      # 64-byte handle; allgather it across the ranks (MPI.Allgather)
#     slab_peer_connect(s, all_handles)           # Vector{UInt8} of world*64 bytes, rank order; then MPI.Barrier
#     slab_set_state(s, d_u_local); ctl = slab_run(s); slab_get_state(s, d_u_local); slab_free(s)
# (d_u_local: a device pointer, e.g. pointer(::CuArray{Float64}) from CUDA.jl.)
const PEER_HANDLE_BYTES = 64
function slab_create(method::Symbol, F::Union{DeviceRHS,UserRHS}, n_local::Integer, n_global::Integer, rank::Integer,
                     world::Integer, tspan; reltol = 1.0e-5, abstol = 1.0e-8, minstep = abs(tspan[end]-tspan[1])/1e18,
                     maxstep = abs(tspan[end]-tspan[1])/2.5, initstep = 0.0, device = 0, stream = C_NULL)
    p = isempty(F.params) ? Float64[0.0] : F.params
    o = Ref(make_opts(METHODS[method], rhs_id(F), min(n_global, typemax(Int32)), length(F.params), 0, POINTS[:final], 0,
                      device, 0, reltol, abstol, minstep, maxstep, initstep))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve p begin
        check(ccall((:ode_b200_slab_create, LIB[]), Cint,
                    (Ref{Opts}, Int64, Int64, Int32, Int32, Ptr{Float64}, Float64, Float64, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                    o, n_local, n_global, rank, world, p, Float64(tspan[1]), Float64(tspan[end]), stream, h))
    end
    h[]
end
function slab_peer_alloc(s::Ptr{Cvoid})
    h = zeros(UInt8, PEER_HANDLE_BYTES)
    check(ccall((:ode_b200_slab_peer_alloc, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Ptr{Cvoid}}), s, h, C_NULL))
    h
end
slab_peer_connect(s::Ptr{Cvoid}, handles::Vector{UInt8}) =
    check(ccall((:ode_b200_slab_peer_connect, LIB[]), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Ptr{Ptr{Cvoid}}), s, handles, C_NULL))
slab_set_state(s::Ptr{Cvoid}, d_u::Ptr) = check(ccall((:ode_b200_slab_set_state, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s, d_u))
slab_get_state(s::Ptr{Cvoid}, d_u::Ptr) = check(ccall((:ode_b200_slab_get_state, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s, d_u))
slab_reset(s::Ptr{Cvoid}, t0, tend) = check(ccall((:ode_b200_slab_reset, LIB[]), Cint, (Ptr{Cvoid}, Float64, Float64), s, t0, tend))
"""slab_run(s) -> (done, status, n_accept, n_reject, t_final, attempts, kernel_ms): the whole integration in one call."""
function slab_run(s::Ptr{Cvoid})
    ctl = zeros(Float64, 16)
    ms = Ref{Float64}(0.0)
    check(ccall((:ode_b200_slab_run, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}), s, ctl, ms))
    (done = ctl[1] != 0, status = Int(ctl[2]), n_accept = Int(ctl[3]), n_reject = Int(ctl[4]), t_final = ctl[8],
     attempts = Int(ctl[10]), kernel_ms = ms[])
end
slab_free(s::Ptr{Cvoid}) = ccall((:ode_b200_slab_free, LIB[]), Cvoid, (Ptr{Cvoid},), s)
export slab_create, slab_peer_alloc, slab_peer_connect, slab_set_state, slab_get_state, slab_reset, slab_run, slab_free

# ---- DiffEqBase.solve(prob, alg): the reference's common-interface method (src/common.jl:1-123) --------------------
# Same signature, same keyword mapping (dtmin = span/1e-9 (sic, :13), dtmax = span/2.5, saveat -> Ts, save_everystep ->
# points, save_start), same reshaping and the same build_solution call with retcode = :Succss (:119-122).  The only
# difference: prob.f must wrap a DeviceRHS / UserRHS token (ODEProblem(DeviceRHS(:lorenz, ...), u0, tspan)); a plain
# Julia function raises, because there is no CPU fallback.
@static if HAVE_DIFFEQBASE
function DiffEqBase.solve(
    prob::DiffEqBase.AbstractODEProblem{uType,tupType,isinplace},
    alg::AlgType,
    timeseries=[], ts=[], ks=[];
    verbose=true,
    save_timeseries=nothing,
    saveat=eltype(tupType)[], reltol = 1e-5, abstol = 1e-8,
    save_everystep=isempty(saveat),
    dense = save_everystep && isempty(saveat),
    save_start = save_everystep || isempty(saveat) || saveat isa Number ?
                 true : prob.tspan[1] in saveat,
    callback=nothing,
    dtmin = abs(prob.tspan[2]-prob.tspan[1])/1e-9,
    dtmax = abs(prob.tspan[2]-prob.tspan[1])/2.5,
    timeseries_errors=true, dense_errors=false,
    dt = 0.0, norm = nothing,
    kwargs...) where {uType,tupType,isinplace,AlgType<:ODEB200Algorithm}

    tType = eltype(tupType)
    F = prob.f isa DiffEqBase.AbstractDiffEqFunction && hasproperty(prob.f, :f) ? prob.f.f : prob.f
    F isa Union{DeviceRHS,UserRHS} ||
        error("ODEB200: the problem's function must be a registered DeviceRHS token (no CPU fallback)")
    norm === nothing || error("ODEB200: only the default norm (LinearAlgebra.norm) is compiled into the kernels")
    if save_timeseries !== nothing
        verbose && @warn("save_timeseries is deprecated. Use save_everystep instead")
        save_everystep = save_timeseries
    end
    if hasproperty(prob.f, :mass_matrix) && prob.f.mass_matrix != DiffEqBase.I
        error("This solver is not able to use mass matrices.")            # src/common.jl:42-44
    end
    if callback !== nothing || :callback ∈ keys(prob.kwargs)
        error("ODE is not compatible with callbacks.")                    # :46-48
    end
    tspan = prob.tspan
    if saveat isa Number                                                   # :52-61
        if (tspan[1]:saveat:tspan[end])[end] == tspan[end]
            saveat_vec = convert(Vector{tType}, collect(tType, tspan[1]+saveat:saveat:tspan[end]))
        else
            saveat_vec = convert(Vector{tType}, collect(tType, tspan[1]+saveat:saveat:(tspan[end]-saveat)))
        end
    else
        saveat_vec = convert(Vector{tType}, collect(saveat))
    end
    if !isempty(saveat_vec) && saveat_vec[end] == tspan[2]                 # :63-65
        pop!(saveat_vec)
    end
    if !isempty(saveat_vec) && saveat_vec[1] == tspan[1]                   # :67-71
        Ts = unique([saveat_vec; tspan[2]])
    else
        Ts = unique([tspan[1]; saveat_vec; tspan[2]])
    end
    points = save_everystep ? :all : :specified                            # :72-76
    sizeu = size(prob.u0)
    u0 = uType <: AbstractArray ? vec(prob.u0) : prob.u0
    # AlgType(...) is ode45(...) etc., like :93-100; fixed-step / multistep front-ends take no keywords there either
    ts, timeseries_tmp = AlgType(F, u0, Ts; abstol = abstol, reltol = reltol, maxstep = dtmax, minstep = dtmin,
                                 initstep = dt, points = points)
    if save_start                                                          # :102-107
        start_idx = 1
    else
        start_idx = 2
        ts = ts[2:end]
    end
    if uType <: AbstractArray                                              # :110-117
        timeseries = Vector{uType}()
        for i = start_idx:length(timeseries_tmp)
            push!(timeseries, reshape(timeseries_tmp[i], sizeu))
        end
    else
        timeseries = timeseries_tmp
    end
    DiffEqBase.build_solution(prob, alg, ts, timeseries,
                              timeseries_errors = timeseries_errors,
                              dense_errors = dense_errors,
                              retcode = :Succss)                           # :119-122
end
end # HAVE_DIFFEQBASE

end # module
