# BoundaryValueDiffEqMIRKB200 — Julia host glue of the B200-native MIRK backend.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no `julia`.  It is the thin, mechanical
# `ccall` layer over include/mirk_b200.h that a maintainer drops next to
# lib/BoundaryValueDiffEqMIRK; the same C ABI is exercised end to end by the Python mirror
# (boundaryvaluediffeq.jl_b200/api.py) and the tests.  The struct layouts below are checked against the
# header by tests/test_host_logic.py (it parses this file).  Reference entry points it replaces:
#   SciMLBase.__init(prob::BVProblem, alg::AbstractMIRK; dt, ...)   lib/BoundaryValueDiffEqMIRK/src/mirk.jl:49-53
#   SciMLBase.solve!(cache)                                          lib/BoundaryValueDiffEqMIRK/src/mirk.jl:286-332
#   EnsembleProblem driver                                           SciMLBase (usage test/Core/ensemble_tests.jl:20-38)
#
# How `solve(prob, MIRK4(); dt)` reaches this backend.  The right-hand side and boundary conditions of a device
# problem are a *device functor* named by `BVPDeviceFunction` (Julia closures cannot run inside a CUDA kernel), put
# where the closure goes: `BVProblem(BVPFunction(BVPDeviceFunction("pendulum"), nothing), u0, tspan, p)`.
#   (a) this module's own `MIRK2 … MIRK6, MIRK6I` carry the reference's fields (`nlsolve`, `optimize`, `jac_alg`,
#       `defect_threshold`, `max_num_subintervals`, algorithms.jl:55-61) plus `device`; with
#       `using BoundaryValueDiffEqMIRKB200: MIRK4, MIRK6` a reference script runs unchanged.
#   (b) with BoundaryValueDiffEqMIRK loaded, ext/BoundaryValueDiffEqMIRKB200MIRKExt.jl adds an `__init` method for
#       the REFERENCE's own `MIRK4()/MIRK6()` that is more specific in the problem's function type, so the
#       reference algorithms dispatch here whenever `prob.f.f isa BVPDeviceFunction`.
module BoundaryValueDiffEqMIRKB200

using SciMLBase, Libdl
import BoundaryValueDiffEqCore: AbstractBoundaryValueDiffEqAlgorithm, AbstractBoundaryValueDiffEqCache,
                                DefectControl, BVPJacobianAlgorithm, BVPVerbosity, DEFAULT_VERBOSE

const libmirkb200 = Ref{String}(get(ENV, "MIRK_B200_LIB", "libmirkb200.so"))

# ---- device functions ------------------------------------------------------------------------
"""
    BVPDeviceFunction(name)

Stands where the closure `f!` stands in `BVPFunction(f!, bc!)`: the registry name of a device functor (RHS +
boundary conditions + boundary times, csrc/problems.cuh).  Arbitrary Julia closures cannot run inside a CUDA kernel.
"""
struct BVPDeviceFunction
    name::String
end
# a device function is never called on the host
(f::BVPDeviceFunction)(args...) = error("BVPDeviceFunction($(f.name)) runs on the device only")

struct MirkProblemInfo
    n::Int32
    n_params::Int32
    problem_type::Int32
    n_bc::Int32
    n_bca::Int32
    max_bc_pts::Int32
end

function problem_id(f::BVPDeviceFunction)
    id = Ref{Int32}(-1)
    check(ccall((:mirk_problem_lookup, libmirkb200[]), Cint, (Cstring, Ref{Int32}), f.name, id))
    return id[]
end
function problem_info(f::BVPDeviceFunction)
    info = Ref{MirkProblemInfo}()
    check(ccall((:mirk_problem_info_get, libmirkb200[]), Cint, (Int32, Ref{MirkProblemInfo}), problem_id(f), info))
    return info[]
end

"register a functor compiled with nvcc from the plugin template (INTEGRATION.md)"
function register_plugin(name::AbstractString, so_path::AbstractString)
    id = Ref{Int32}(-1)
    check(ccall((:mirk_problem_register_plugin, libmirkb200[]), Cint, (Cstring, Cstring, Ref{Int32}), name, so_path, id))
    return BVPDeviceFunction(String(name))
end

device_function(prob) = prob.f.f isa BVPDeviceFunction ? prob.f.f : nothing

# ---- algorithms: the reference's fields (lib/BoundaryValueDiffEqMIRK/src/algorithms.jl:55-61) + `device` ------
abstract type AbstractMIRKB200 <: AbstractBoundaryValueDiffEqAlgorithm end
for (name, ord) in ((:MIRK2, 2), (:MIRK3, 3), (:MIRK4, 4), (:MIRK5, 5), (:MIRK6, 6), (:MIRK6I, 7))  # 7 = the C ABI's code of MIRK6I
    @eval begin
        Base.@kwdef struct $name{N, O, J <: BVPJacobianAlgorithm} <: AbstractMIRKB200
            nlsolve::N = nothing
            optimize::O = nothing
            jac_alg::J = BVPJacobianAlgorithm()
            defect_threshold::Float64 = 0.1
            max_num_subintervals::Int = 3000
            device::Int = 0
        end
        tableau_code(::$name) = $ord
    end
end
# the pre-rename spellings of round 1 stay as aliases
const MIRK2B200 = MIRK2; const MIRK3B200 = MIRK3; const MIRK4B200 = MIRK4
const MIRK5B200 = MIRK5; const MIRK6B200 = MIRK6; const MIRK6IB200 = MIRK6I

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
    nlsolve::Int32
    controller::Int32
    ge_method::Int32
    DE::Float64
    GE::Float64
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

# ---- cache: the fields the reference's tests read (.prob, .alg, .nlsolve_kwargs, .optimize_kwargs, .verbose) ----
mutable struct MIRKB200Cache{P, A, NK, OK} <: AbstractBoundaryValueDiffEqCache
    prob::P
    alg::A
    handle::Ptr{Cvoid}
    n::Int
    nlsolve_kwargs::NK
    optimize_kwargs::OK
    verbose::BVPVerbosity
end

# verbose = true / false / a BVPVerbosity, as the reference normalises it (CORE/src/utils.jl:882-895)
_verbosity(v::BVPVerbosity) = v
_verbosity(v::Bool) = v ? DEFAULT_VERBOSE : BVPVerbosity(SciMLBase.SciMLLogging.None())
_verbosity(v) = BVPVerbosity(v)

# Initial guess -> (mesh, node-major n x N matrix).  The cases of CORE/src/utils.jl:339-388,694-704,750-773:
#   vector of numbers          constant guess on range(t0, t1, cld(t1 - t0, dt) + 1)     (needs dt > 0)
#   function u0(p, t) / u0(t)  evaluated on the same uniform mesh                          (needs dt > 0)
#   vector of vectors / VectorOfArray   guess on range(t0, t1, length(u0))
#   DiffEqArray / ODESolution / any object with .t and .u   guess AND mesh
function _guess(u0, p, t0, t1, dt)
    if u0 isa AbstractVector{<:Number}
        return nothing, collect(Float64, u0)                       # handled by mirk_set_uniform_guess
    elseif u0 isa Function
        dt > 0 || throw(ArgumentError("dt must be positive"))
        mesh = collect(range(t0; stop = t1, length = Int(cld(t1 - t0, dt)) + 1))
        f = applicable(u0, p, t0) ? (t -> u0(p, t)) : u0            # u0(t) is deprecated upstream but accepted
        return mesh, reduce(hcat, (collect(Float64, f(t)) for t in mesh))
    elseif hasproperty(u0, :t) && hasproperty(u0, :u)               # DiffEqArray, ODESolution, a previous BVP solution
        return collect(Float64, u0.t), reduce(hcat, (collect(Float64, ui) for ui in u0.u))
    else                                                            # vector of vectors / VectorOfArray
        us = hasproperty(u0, :u) ? u0.u : u0
        mesh = collect(range(t0; stop = t1, length = length(us)))
        return mesh, reduce(hcat, (collect(Float64, ui) for ui in us))
    end
end

# alg.nlsolve -> the C ABI's solver code.  `nothing` is the reference default (polyalgorithm NewtonRaphson ->
# NewtonRaphson + BackTracking -> TrustRegion, CORE/src/default_internal_solve.jl:31-45); the three sub-solvers can be
# requested on their own; anything else cannot run on the device and is rejected rather than ignored.
nlsolve_code(::Nothing) = Int32(0)
function nlsolve_code(alg)
    name = string(nameof(typeof(alg)))
    occursin("TrustRegion", name) && return Int32(3)
    if occursin("GeneralizedFirstOrderAlgorithm", name) || occursin("NewtonRaphson", name)
        ls = hasproperty(alg, :linesearch) ? getproperty(alg, :linesearch) : nothing
        (ls === nothing || occursin("NoLineSearch", string(typeof(ls)))) && return Int32(1)
        occursin("BackTracking", string(typeof(ls))) && return Int32(2)
    end
    throw(ArgumentError("nlsolve = $(typeof(alg)) cannot run on the B200 backend (supported: nothing, NewtonRaphson(), NewtonRaphson(linesearch = BackTracking()), TrustRegion())"))
end

# controller -> (code, global-error method, DE, GE, defect_threshold) of the C ABI (CORE/src/calc_errors.jl:54-139)
ge_method_code(m) = occursin("REErrorControl", string(typeof(m))) ? Int32(1) : Int32(0)
controller_fields(c::DefectControl) = (Int32(0), Int32(0), 1.0, 1.0, Float64(c.defect_threshold))
function controller_fields(c)
    name = string(nameof(typeof(c)))
    name == "GlobalErrorControl" && return (Int32(1), ge_method_code(c.method), 1.0, 1.0, 0.1)
    name == "SequentialErrorControl" &&
        return (Int32(2), ge_method_code(c.global_error.method), 1.0, 1.0, Float64(c.defect.defect_threshold))
    name == "HybridErrorControl" &&
        return (Int32(3), ge_method_code(c.global_error.method), Float64(c.DE), Float64(c.GE), Float64(c.defect.defect_threshold))
    throw(ArgumentError("unknown error controller $(typeof(c))"))
end

function __init_b200(prob::BVProblem, alg, order::Integer, device::Integer; dt = 0.0, abstol = 1e-6, adaptive = true,
        controller = DefectControl(), nlsolve_kwargs = (; abstol = abstol), optimize_kwargs = (;),
        verbose = DEFAULT_VERBOSE, kwargs...)
    f = device_function(prob)
    f === nothing && throw(ArgumentError("the B200 backend needs prob.f to wrap a BVPDeviceFunction"))
    alg.optimize === nothing || throw(ArgumentError("`optimize` solvers are not supported by the B200 backend"))
    ctrl, gem, DE, GE, thr = controller_fields(controller)
    p = prob.p isa SciMLBase.NullParameters ? Float64[] : collect(Float64, prob.p)
    desc = MirkDesc(problem_id(f), order, get(nlsolve_kwargs, :abstol, abstol), adaptive,
        thr, alg.max_num_subintervals, get(nlsolve_kwargs, :maxiters, 1000), 0, 0,
        device, length(p), pointer(p), nlsolve_code(alg.nlsolve), ctrl, gem, DE, GE)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve p check(ccall((:mirk_create, libmirkb200[]), Cint, (Ref{MirkDesc}, Ref{Ptr{Cvoid}}), desc, h))
    t0, t1 = prob.tspan
    mesh, y = _guess(prob.u0, prob.p, t0, t1, dt)    # copies: prob.u0 is never mutated (mirk_basic_tests.jl:717-719)
    if mesh === nothing
        check(ccall((:mirk_set_uniform_guess, libmirkb200[]), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Float64}),
            h[], t0, t1, dt, y))
        n = length(y)
    else
        n = size(y, 1)
        check(ccall((:mirk_set_mesh_guess, libmirkb200[]), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
            h[], length(mesh), mesh, y))
    end
    # like the reference's cache, prob.u0 of the stored problem is a plain state vector (mirk.jl:71-77)
    prob_stored = prob.u0 isa AbstractVector{<:Number} ? prob : SciMLBase.remake(prob; u0 = y[:, 1])
    cache = MIRKB200Cache(prob_stored, alg, h[], n, nlsolve_kwargs, optimize_kwargs, _verbosity(verbose))
    finalizer(c -> ccall((:mirk_destroy, libmirkb200[]), Cint, (Ptr{Cvoid},), c.handle), cache)
    return cache
end

SciMLBase.__init(prob::BVProblem, alg::AbstractMIRKB200; kwargs...) =
    __init_b200(prob, alg, tableau_code(alg), alg.device; kwargs...)

function SciMLBase.solve!(cache::MIRKB200Cache)
    res = Ref{MirkResult}()
    check(ccall((:mirk_solve, libmirkb200[]), Cint, (Ptr{Cvoid}, Ref{MirkResult}), cache.handle, res))
    N, n = Int(res[].n_mesh), cache.n
    mesh = Vector{Float64}(undef, N)
    y = Matrix{Float64}(undef, n, N)               # column i = node i: node-major in memory
    check(ccall((:mirk_get_solution, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), cache.handle, mesh, y))
    L = Int(problem_info(device_function(cache.prob)).n_bc)
    resid = Vector{Float64}(undef, L + (N - 1) * n)
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
"`solve(EnsembleProblem(prob; prob_func), MIRK4(), EnsembleB200(); trajectories, dt)`"
struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    node_cap::Int       # 0: the library default (on-chip state up to its limit, HBM slab beyond)
end
EnsembleB200() = EnsembleB200(0)

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
    nlsolve::Int32
end

# generic functions the CUDA extension adds device-array methods to
function set_guess! end
function solution! end

# packed inputs of a sweep: (params np x ntraj, u0).  Host arrays here; the CUDA extension adds the CuArray methods
stage_params(p::AbstractMatrix) = Matrix{Float64}(p)
set_inputs!(h, params::Matrix{Float64}, u0::Vector{Float64}) =
    check(ccall((:mirk_ensemble_set_inputs, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32),
        h, params, u0, 0))

function ensemble_order(alg, order)
    order in (4, 6) || throw(ArgumentError("the batched ensemble kernel is instantiated for MIRK4 and MIRK6"))
    return order
end

function __solve_ensemble_b200(ens, alg, order, device, ealg::EnsembleB200; trajectories, dt, abstol = 1e-6,
        adaptive = true, controller = DefectControl(), packed_params = nothing, kwargs...)
    base = ens.prob
    f = device_function(base)
    f === nothing && throw(ArgumentError("the B200 backend needs prob.f to wrap a BVPDeviceFunction"))
    # prob_func is harvested on the host (2-argument form of this SciMLBase major, ensemble_tests.jl:20,37);
    # `packed_params` (np x trajectories, host Matrix or CuArray) is the fast path that skips 262 144 remake calls
    probs = packed_params === nothing ? [ens.prob_func(base, i) for i in 1:trajectories] : nothing
    params = packed_params === nothing ? reduce(hcat, (collect(Float64, q.p) for q in probs)) : packed_params
    u0 = collect(Float64, base.u0)
    desc = MirkEnsembleDesc(problem_id(f), ensemble_order(alg, order), abstol, adaptive, controller.defect_threshold,
        alg.max_num_subintervals, 1000, 0, device, ealg.node_cap, base.tspan[1], base.tspan[2], dt,
        nlsolve_code(alg.nlsolve) == 1 ? Int32(1) : Int32(0))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mirk_ensemble_create, libmirkb200[]), Cint, (Ref{MirkEnsembleDesc}, Int64, Ref{Ptr{Cvoid}}), desc, trajectories, h))
    try
        set_inputs!(h[], stage_params(params), u0)
        check(ccall((:mirk_ensemble_run, libmirkb200[]), Cint, (Ptr{Cvoid}, Ptr{Cfloat}), h[], C_NULL))
        ret = Vector{Int32}(undef, trajectories)
        resid = Vector{Float64}(undef, trajectories)
        check(ccall((:mirk_ensemble_get_results, libmirkb200[]), Cint,
            (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h[], ret, C_NULL, C_NULL, C_NULL, resid, C_NULL, C_NULL))
        # one SciMLBase solution per trajectory (sol.t, sol.u, sol.retcode), wrapped like the stock ensemble driver does
        n = length(u0)
        cap = Ref{Int32}(0)
        check(ccall((:mirk_ensemble_node_cap, libmirkb200[]), Cint, (Ptr{Cvoid}, Ref{Int32}), h[], cap))
        mesh = Vector{Float64}(undef, cap[]); y = Matrix{Float64}(undef, n, cap[])
        sols = map(1:trajectories) do i
            N = Ref{Int32}(0)
            check(ccall((:mirk_ensemble_get_trajectory, libmirkb200[]), Cint,
                (Ptr{Cvoid}, Int64, Ref{Int32}, Ptr{Float64}, Ptr{Float64}), h[], i - 1, N, mesh, y))
            prob_i = probs === nothing ? base : probs[i]
            SciMLBase.build_solution(prob_i, alg, mesh[1:N[]], [y[:, k] for k in 1:N[]]; retcode = RETCODES[ret[i] + 1],
                resid = [resid[i]])
        end
        return SciMLBase.EnsembleSolution(sols, 0.0, all(==(0), ret))
    finally
        ccall((:mirk_ensemble_destroy, libmirkb200[]), Cint, (Ptr{Cvoid},), h[])
    end
end

SciMLBase.__solve(ens::SciMLBase.AbstractEnsembleProblem, alg::AbstractMIRKB200, ealg::EnsembleB200; kwargs...) =
    __solve_ensemble_b200(ens, alg, tableau_code(alg), alg.device, ealg; kwargs...)

export BVPDeviceFunction, EnsembleB200, register_plugin
# MIRK2 … MIRK6I are deliberately NOT exported (they would clash with BoundaryValueDiffEqMIRK's): import them
# explicitly, `using BoundaryValueDiffEqMIRKB200: MIRK4, MIRK6`, or use the reference's own through the extension.

end # module
