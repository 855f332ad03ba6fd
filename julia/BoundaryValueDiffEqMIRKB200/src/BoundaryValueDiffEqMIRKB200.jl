# BoundaryValueDiffEqMIRKB200 — Julia host glue of the B200-native MIRK backend.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia`.  It is the thin, mechanical
# `ccall` layer over include/mirk_b200.h that a maintainer drops next to
# lib/BoundaryValueDiffEqMIRK; the same C ABI is exercised end to end by the Python mirror
# (boundaryvaluediffeq.jl_b200/api.py) and the tests.  Reference entry points it replaces:
#   SciMLBase.__init(prob::BVProblem, alg::AbstractMIRK; dt, ...)   lib/BoundaryValueDiffEqMIRK/src/mirk.jl:49-53
#   SciMLBase.solve!(cache)                                          lib/BoundaryValueDiffEqMIRK/src/mirk.jl:286-332
#   EnsembleProblem driver                                           SciMLBase (usage test/Core/ensemble_tests.jl:20-38)
module BoundaryValueDiffEqMIRKB200

using SciMLBase, Libdl
import BoundaryValueDiffEqCore: AbstractBoundaryValueDiffEqAlgorithm, AbstractBoundaryValueDiffEqCache,
                                DefectControl

const libmirkb200 = Ref{String}(get(ENV, "MIRK_B200_LIB", "libmirkb200.so"))

# ---- device functions ------------------------------------------------------------------------
"""
    BVPDeviceFunction(name)

Stands where `BVPFunction(f!, bc!)` stands: the registry name of a device functor (RHS + boundary
conditions, csrc/problems.cuh).  Arbitrary Julia closures cannot run inside a CUDA kernel.
"""
struct BVPDeviceFunction
    name::String
end

function problem_id(f::BVPDeviceFunction)
    id = Ref{Int32}(-1)
    check(ccall((:mirk_problem_lookup, libmirkb200[]), Cint, (Cstring, Ref{Int32}), f.name, id))
    return id[]
end

"register a functor compiled with nvcc from the plugin template (INTEGRATION.md)"
function register_plugin(name::AbstractString, so_path::AbstractString)
    id = Ref{Int32}(-1)
    check(ccall((:mirk_problem_register_plugin, libmirkb200[]), Cint, (Cstring, Cstring, Ref{Int32}), name, so_path, id))
    return BVPDeviceFunction(String(name))
end

# ---- algorithms (same fields as MIRK4()/MIRK6(), lib/BoundaryValueDiffEqMIRK/src/algorithms.jl:55-61)
abstract type AbstractMIRKB200 <: AbstractBoundaryValueDiffEqAlgorithm end
Base.@kwdef struct MIRK4B200 <: AbstractMIRKB200
    defect_threshold::Float64 = 0.1
    max_num_subintervals::Int = 3000
    device::Int = 0
end
Base.@kwdef struct MIRK6B200 <: AbstractMIRKB200
    defect_threshold::Float64 = 0.1
    max_num_subintervals::Int = 3000
    device::Int = 0
end
for (name, ord) in ((:MIRK2B200, 2), (:MIRK3B200, 3), (:MIRK5B200, 5), (:MIRK6IB200, 7))  # 7 = the C ABI's code of MIRK6I
    @eval begin
        Base.@kwdef struct $name <: AbstractMIRKB200
            defect_threshold::Float64 = 0.1
            max_num_subintervals::Int = 3000
            device::Int = 0
        end
        alg_order(::$name) = $ord
    end
end
alg_order(::MIRK4B200) = 4
alg_order(::MIRK6B200) = 6

# ---- C structs (layout of include/mirk_b200.h) -------------------------------------------------
struct MirkDesc
    problem_id::Int32
    order::Int32
    abstol::Float64
    adaptive::Int32
    defect_threshold::Float64
    max_num_subintervals::Int32
    maxiters::Int32
    reinterp_inplace::Int32
    chunk::Int32
    device::Int32
    n_params::Int32
    params::Ptr{Float64}
end

struct MirkResult
    retcode::Int32
    n_mesh::Int32
    outer_iters::Int32
    newton_iters::Int32
    resid_norm::Float64
    defect_norm::Float64
    n_hist::Int32
    hist_n_mesh::NTuple{64, Int32}
    hist_newton::NTuple{64, Int32}
    hist_defect::NTuple{64, Float64}
end

function check(code::Integer)
    code >= 0 && return code
    msg = unsafe_string(ccall((:mirk_last_error, libmirkb200[]), Cstring, ()))
    occursin("dt must be positive", msg) && throw(ArgumentError("dt must be positive"))  # CORE/src/utils.jl:354
    error("libmirkb200 error $code: $msg")
end

const RETCODES = (ReturnCode.Success, ReturnCode.Failure, ReturnCode.MaxIters, ReturnCode.Unstable, ReturnCode.Stalled)

# ---- cache -----------------------------------------------------------------------------------
mutable struct MIRKB200Cache{P, A} <: AbstractBoundaryValueDiffEqCache
    prob::P
    alg::A
    handle::Ptr{Cvoid}
    n::Int
    nlsolve_kwargs::Any
    optimize_kwargs::Any
    verbose::Any
end

function SciMLBase.__init(prob::BVProblem, alg::AbstractMIRKB200; dt = 0.0, abstol = 1e-6, adaptive = true,
        controller = DefectControl(), nlsolve_kwargs = (; abstol = abstol), optimize_kwargs = (;),
        verbose = nothing, kwargs...)
    f = prob.f.f
    f isa BVPDeviceFunction || throw(ArgumentError("the B200 backend needs prob.f to wrap a BVPDeviceFunction"))
    p = prob.p isa SciMLBase.NullParameters ? Float64[] : collect(Float64, prob.p)
    desc = MirkDesc(problem_id(f), alg_order(alg), get(nlsolve_kwargs, :abstol, abstol), adaptive,
        controller.defect_threshold, alg.max_num_subintervals, get(nlsolve_kwargs, :maxiters, 1000), 0, 0,
        alg.device, length(p), pointer(p))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve p check(ccall((:mirk_create, libmirkb200[]), Cint, (Ref{MirkDesc}, Ref{Ptr{Cvoid}}), desc, h))
    t0, t1 = prob.tspan
    u0 = prob.u0
    if u0 isa AbstractVector{<:Number}            # constant guess on a uniform mesh, CORE/src/utils.jl:766-769
        u = collect(Float64, u0)                   # prob.u0 is never mutated (mirk_basic_tests.jl:717-719)
        check(ccall((:mirk_set_uniform_guess, libmirkb200[]), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Float64}),
            h[], t0, t1, dt, u))
        n = length(u)
    else                                           # vector of vectors carries its own mesh length, utils.jl:339-348
        N, n = length(u0), length(first(u0))
        mesh = collect(range(t0; stop = t1, length = N))
        y = reduce(vcat, (collect(Float64, ui) for ui in u0))      # node-major like recursive_flatten
        check(ccall((:mirk_set_mesh_guess, libmirkb200[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
            h[], N, mesh, y))
    end
    cache = MIRKB200Cache(prob, alg, h[], n, nlsolve_kwargs, optimize_kwargs, verbose)
    finalizer(c -> ccall((:mirk_destroy, libmirkb200[]), Cint, (Ptr{Cvoid},), c.handle), cache)
    return cache
end

function SciMLBase.solve!(cache::MIRKB200Cache)
    res = Ref{MirkResult}()
    check(ccall((:mirk_solve, libmirkb200[]), Cint, (Ptr{Cvoid}, Ref{MirkResult}), cache.handle, res))
    N, n = Int(res[].n_mesh), cache.n
    mesh = Vector{Float64}(undef, N)
    y = Matrix{Float64}(undef, n, N)               # column i = node i: node-major in memory
    check(ccall((:mirk_get_solution, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), cache.handle, mesh, y))
    resid = Vector{Float64}(undef, n + (N - 1) * n)
    check(ccall((:mirk_get_residual, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}), cache.handle, resid))
    u = [y[:, i] for i in 1:N]
    sol = SciMLBase.build_solution(cache.prob, cache.alg, mesh, u; interp = MIRKB200Interpolation(cache),
        retcode = RETCODES[res[].retcode + 1], resid = resid, original = res[])
    return sol
end

# dense output: sol(t) and sol(t, Val{1}) evaluate on the device (lib/BoundaryValueDiffEqMIRK/src/interpolation.jl:17-96)
struct MIRKB200Interpolation{C} <: SciMLBase.AbstractDiffEqInterpolation
    cache::C
end
function (id::MIRKB200Interpolation)(tvals, idxs, deriv, p, continuity::Symbol = :left)
    ts = collect(Float64, tvals isa Number ? (tvals,) : tvals)
    out = Matrix{Float64}(undef, id.cache.n, length(ts))
    d = deriv === Val{1} ? 1 : 0
    check(ccall((:mirk_interp, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Float64}),
        id.cache.handle, ts, length(ts), d, out))
    sel = idxs === nothing ? Colon() : idxs
    return tvals isa Number ? out[sel, 1] : [out[sel, j] for j in axes(out, 2)]
end
SciMLBase.interp_summary(::MIRKB200Interpolation) = "MIRK continuous extension evaluated by libmirkb200"

# ---- ensembles -----------------------------------------------------------------------------------
"`solve(EnsembleProblem(prob; prob_func), MIRK4B200(), EnsembleB200(); trajectories, dt)`"
struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    node_cap::Int
end
EnsembleB200() = EnsembleB200(128)

struct MirkEnsembleDesc
    problem_id::Int32
    order::Int32
    abstol::Float64
    adaptive::Int32
    defect_threshold::Float64
    max_num_subintervals::Int32
    maxiters::Int32
    reinterp_inplace::Int32
    device::Int32
    node_cap::Int32
    t0::Float64
    t1::Float64
    dt::Float64
end

function SciMLBase.__solve(ens::SciMLBase.AbstractEnsembleProblem, alg::AbstractMIRKB200, ealg::EnsembleB200;
        trajectories, dt, abstol = 1e-6, adaptive = true, controller = DefectControl(), kwargs...)
    base = ens.prob
    f = base.f.f::BVPDeviceFunction
    probs = [ens.prob_func(base, i) for i in 1:trajectories]          # harvested on the host, 2-arg prob_func
    params = reduce(hcat, (collect(Float64, q.p) for q in probs))      # np × ntraj = [ntraj][np] row-major
    u0 = collect(Float64, base.u0)
    desc = MirkEnsembleDesc(problem_id(f), alg_order(alg), abstol, adaptive, controller.defect_threshold,
        alg.max_num_subintervals, 1000, 0, alg.device, ealg.node_cap, base.tspan[1], base.tspan[2], dt)
    ret = Vector{Int32}(undef, trajectories)
    nm = similar(ret); its = similar(ret)
    yfirst = Matrix{Float64}(undef, length(u0), trajectories)
    check(ccall((:mirk_ensemble_solve, libmirkb200[]), Cint,
        (Ref{MirkEnsembleDesc}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        desc, trajectories, params, u0, 0, ret, nm, its, yfirst))
    converged = all(==(0), ret)
    return (; retcodes = RETCODES[ret .+ 1], n_mesh = nm, newton_iters = its, u_first = yfirst, converged)
end

export BVPDeviceFunction, MIRK2B200, MIRK3B200, MIRK4B200, MIRK5B200, MIRK6B200, MIRK6IB200, EnsembleB200, register_plugin

end # module
