# dump_reference.jl — run the REAL reference (SciML/BoundaryValueDiffEq.jl) on the problems this repository's oracle and
# CUDA path are tested on, and write golden vectors the test-suite consumes (tests/test_reference_goldens.py).
#
#   julia --project=<env with BoundaryValueDiffEq> julia/parity/dump_reference.jl [outdir = tests/golden]
#
# This image has no `julia`, so the fixtures are absent and parity is UNPINNED (DESIGN.md §2); any Julia box can produce
# them.  For every case it records what the north star asks to match: the Newton step count of every outer iteration
# (`sol_nlprob.stats.nsteps`), the mesh size history, the return code, the final mesh and the solution values.  The
# outer loop below is `SciMLBase.solve!(::MIRKCache)` (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:286-306) unrolled so the
# per-iteration quantities can be observed; it calls the reference's own `__perform_mirk_iteration`.
#
# The problems are the built-ins of boundaryvaluediffeq.jl_b200/csrc/problems.cuh (= oracle/mirk_problems.c), which
# themselves restate benchmark/simple_pendulum.jl and lib/BoundaryValueDiffEqMIRK/test/Core/mirk_basic_tests.jl.
using BoundaryValueDiffEq, BoundaryValueDiffEqMIRK, SciMLBase, Printf
import BoundaryValueDiffEqMIRK: __perform_mirk_iteration, __split_kwargs

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "..", "tests", "golden")
mkpath(outdir)

# ---- problems (name => (f!, bc!, bc style, n)) ----------------------------------------------------------------------
pendulum_f!(du, u, p, t) = (du[1] = u[2]; du[2] = -p[1] * sin(u[1]); nothing)
pendulum_bc!(res, sol, p, t) = (res[1] = sol(pi / 4)[1] + pi / 2; res[2] = sol(pi / 2)[1] - pi / 2; nothing)   # benchmark/simple_pendulum.jl:14-19

# u'' = -k u with conditions at times p[2], p[4] on components p[6], p[7] (0-based in the device functor)
linear2_f!(du, u, p, t) = (du[1] = u[2]; du[2] = -p[1] * u[1]; nothing)
linear2_bc!(res, sol, p, t) = (res[1] = sol(p[2])[Int(p[6]) + 1] - p[3]; res[2] = sol(p[4])[Int(p[7]) + 1] - p[5]; nothing)
linear2tp_bca!(res, ua, p) = (res[1] = ua[1] - p[2]; nothing)
linear2tp_bcb!(res, ub, p) = (res[1] = ub[1] - p[3]; nothing)

function swirling_f!(du, u, p, t)
    e = p[1]
    du[1] = u[2]; du[2] = (u[1] * u[4] - u[3] * u[2]) / e; du[3] = u[4]; du[4] = u[5]; du[5] = u[6]
    du[6] = (-u[3] * u[6] - u[1] * u[2]) / e
    return nothing
end
function swirling_bc!(res, sol, p, t)
    a, b = sol(0.0), sol(1.0)
    res[1] = a[1] + 1.0; res[2] = a[3]; res[3] = a[4]; res[4] = b[1] - 1.0; res[5] = b[3]; res[6] = b[4]
    return nothing
end
lotka_f!(du, u, p, t) = (du[1] = p[1] * u[1] - p[2] * u[1] * u[2]; du[2] = -p[3] * u[2] + p[4] * u[1] * u[2]; nothing)
lotka_bc!(res, sol, p, t) = (res[1] = sol(0.0)[1] - 1.0; res[2] = sol(0.0)[2] - 2.0; nothing)
function layer_f!(du, u, p, t)
    du[1] = u[2]
    du[2] = -t / p[1] * u[2] - pi^2 * cos(pi * t) - pi * t / p[1] * sin(pi * t)
    return nothing
end
layer_bc!(res, sol, p, t) = (res[1] = sol(-1.0)[1] + 2.0; res[2] = sol(1.0)[1]; nothing)

# u'' = -u with a boundary condition that reads the DERIVATIVE of the interpolant (sol(t, Val{1}),
# lib/BoundaryValueDiffEqMIRK/src/interpolation.jl:277-292); p = [alpha, c]: u1(0) = 0, u1(pi/2) - 1 + alpha (u1'(pi/4) - c) = 0.
# Pins what the repository assumes about it: the derivative is built from the Float64 stage buffers, so it does not enter
# the boundary Jacobian and Newton converges linearly in that row (5 steps at alpha = 0.1, not the 1 of a linear problem).
robin_f!(du, u, p, t) = (du[1] = u[2]; du[2] = -u[1]; nothing)
robin_bc!(res, sol, p, t) = (res[1] = sol(0.0)[1]; res[2] = sol(pi / 2)[1] - 1.0 + p[1] * (sol(pi / 4, Val{1})[1] - p[2]); nothing)

# chain of NP torsionally coupled pendula (BASELINE config C2 at a size the reference finishes quickly)
function chain_f!(du, u, p, t)
    NP = length(u) ÷ 2
    g, kap = p[1], p[2]
    for k in 1:NP
        du[k] = u[NP + k]
        acc = -g * sin(u[k]) - 2kap * u[k]
        k > 1 && (acc += kap * u[k - 1])
        k < NP && (acc += kap * u[k + 1])
        du[NP + k] = acc
    end
    return nothing
end
chain_bca!(res, ua, p) = (NP = length(ua) ÷ 2; for k in 1:NP; res[k] = ua[k] - p[2 + k]; end; nothing)
chain_bcb!(res, ub, p) = (NP = length(ub) ÷ 2; for k in 1:NP; res[k] = ub[k] - p[2 + NP + k]; end; nothing)

# ---- JSON without dependencies --------------------------------------------------------------------------------------
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isfinite(x) ? @sprintf("%.17g", x) : "null"
jarr(v::AbstractVector{<:Real}) = "[" * join(jnum.(v), ",") * "]"
jarr(v::AbstractVector) = "[" * join(jarr.(v), ",") * "]"

function dump_case(file, name, order, prob, alg; dt, abstol = 1.0e-6, adaptive = true, extra = "")
    cache = SciMLBase.__init(prob, alg; dt, abstol, adaptive)
    (abstol_, adaptive_, controller, _), _ = __split_kwargs(; cache.kwargs...)
    hist_n, hist_newton, hist_defect = Int[], Int[], Float64[]
    push!(hist_n, length(cache.mesh))
    sol_nl, info, err = __perform_mirk_iteration(cache, abstol_, adaptive_, controller)
    push!(hist_newton, sol_nl.stats.nsteps); push!(hist_defect, err)
    if adaptive_
        while SciMLBase.successful_retcode(info) && err > abstol_
            push!(hist_n, length(cache.mesh))
            sol_nl, info, err = __perform_mirk_iteration(cache, abstol_, adaptive_, controller)
            push!(hist_newton, sol_nl.stats.nsteps); push!(hist_defect, err)
        end
    end
    # the public path on a fresh cache, for the values themselves
    sol = solve(prob, alg; dt, abstol, adaptive)
    p = prob.p isa SciMLBase.NullParameters ? Float64[] : collect(Float64, prob.p)
    u0 = prob.u0 isa AbstractVector{<:Number} ? jarr(collect(Float64, prob.u0)) : "null"
    open(joinpath(outdir, file), "w") do io
        print(io, "{\"name\":\"$name\",\"order\":$order,\"p\":", jarr(p), ",\"u0\":", u0,
            ",\"tspan\":", jarr(collect(Float64, prob.tspan)), ",\"dt\":", jnum(dt), ",\"abstol\":", jnum(abstol),
            ",\"adaptive\":", adaptive ? "true" : "false", extra,
            ",\"retcode\":\"", string(sol.retcode), "\",\"hist_n_mesh\":", jarr(hist_n), ",\"hist_newton\":", jarr(hist_newton),
            ",\"hist_defect\":", jarr(hist_defect), ",\"nlsolve_retcode_last\":\"", string(sol_nl.retcode),
            "\",\"t\":", jarr(collect(Float64, sol.t)), ",\"u\":", jarr([collect(Float64, ui) for ui in sol.u]),
            ",\"julia\":\"", string(VERSION), "\",\"note\":\"written by julia/parity/dump_reference.jl from the unmodified reference\"}")
    end
    println(file, ": ", sol.retcode, " meshes ", hist_n, " newton ", hist_newton)
end

algs = ((4, MIRK4()), (6, MIRK6()))
for (order, alg) in algs
    # BASELINE config C1
    tspan = (0.0, pi / 2)
    dump_case("ref_pendulum_mirk$(order).json", "pendulum", order,
        BVProblem(BVPFunction(pendulum_f!, pendulum_bc!; bcresid_prototype = zeros(2)), [pi / 2, pi / 2], tspan, [9.81]), alg; dt = 0.05)
    # mirk_basic_tests.jl:16-47
    lp = [1.0, 0.0, 5.0, 5.0, 0.0, 0.0, 0.0]
    dump_case("ref_linear2_mirk$(order).json", "linear2", order,
        BVProblem(BVPFunction(linear2_f!, linear2_bc!; bcresid_prototype = zeros(2)), [5.0, -3.5], (0.0, 5.0), lp), alg; dt = 0.2)
    dump_case("ref_linear2_tp_mirk$(order).json", "linear2_tp", order,
        TwoPointBVProblem(linear2_f!, (linear2tp_bca!, linear2tp_bcb!), [5.0, -3.5], (0.0, 5.0), [1.0, 5.0, 0.0];
            bcresid_prototype = (zeros(1), zeros(1))), alg; dt = 0.2)
    dump_case("ref_swirling_mirk$(order).json", "swirling", order,
        BVProblem(BVPFunction(swirling_f!, swirling_bc!; bcresid_prototype = zeros(6)), zeros(6), (0.0, 1.0), [0.01]), alg;
        dt = 0.01, abstol = 1.0e-4)
    dump_case("ref_layer_mirk$(order).json", "layer", order,
        BVProblem(BVPFunction(layer_f!, layer_bc!; bcresid_prototype = zeros(2)), [0.0, 0.0], (-1.0, 1.0), [0.01]), alg; dt = 0.05)
    dump_case("ref_robin_sine_mirk$(order).json", "robin_sine", order,
        BVProblem(BVPFunction(robin_f!, robin_bc!; bcresid_prototype = zeros(2)), [0.0, 1.0], (0.0, pi / 2), [0.1, cos(pi / 4)]), alg;
        dt = order == 4 ? 0.05 : 0.1)
    # a start from which plain NewtonRaphson fails: pins the polyalgorithm fallbacks and the halve-and-zero path
    dump_case("ref_lotka_hard_mirk$(order).json", "lotka", order,
        BVProblem(BVPFunction(lotka_f!, lotka_bc!; bcresid_prototype = zeros(2)), [5.0, 5.0], (0.0, 10.0), [7.5, 4.0, 8.5, 5.0]), alg; dt = 0.1)
end
# BASELINE config C2's problem (n = 16, MIRK6, fixed mesh) at 400 intervals: Newton count with the reference's own
# (too narrow, quirk Q1) sparsity pattern and with a dense Jacobian
let NP = 8, T = 0.5, nint = 400
    # a, b ~ U(-1, 1) from numpy.random.default_rng(0): pasted so that both sides see identical numbers
    a = [0.2739233746429086, -0.4604265724722594, -0.9180529521276106, -0.9669447289429418, 0.6265404784005448, 0.8255111545554434, 0.21327155153435973, 0.4589931219679968]
    b = [0.08724998293084574, 0.8701448475755365, 0.6317071082430643, -0.9945229996597038, 0.7148085531751387, -0.9328288493890713, 0.45931089285988813, -0.648688758794882]
    p = vcat([9.81, 4.0], a, b)
    mesh = collect(range(0.0; stop = T, length = nint + 1))
    guess = [vcat(a .+ (b .- a) .* (t / T), (b .- a) ./ T) for t in mesh]
    prob = TwoPointBVProblem(chain_f!, (chain_bca!, chain_bcb!), guess, (0.0, T), p; bcresid_prototype = (zeros(NP), zeros(NP)))
    dump_case("ref_chain8_fixed_mirk6.json", "chain8", 6, prob, MIRK6(); dt = T / nint, adaptive = false,
        extra = ",\"nint\":$nint,\"jacobian\":\"default sparse pattern (quirk Q1)\"")
    dense = MIRK6(; jac_alg = BVPJacobianAlgorithm(AutoForwardDiff()))
    dump_case("ref_chain8_fixed_mirk6_dense.json", "chain8", 6, prob, dense; dt = T / nint, adaptive = false,
        extra = ",\"nint\":$nint,\"jacobian\":\"dense AutoForwardDiff\"")
end
