# Extension loaded when the REFERENCE package BoundaryValueDiffEqMIRK is in the session: its own `MIRK4()` /
# `MIRK6()` (… MIRK2, MIRK3, MIRK5, MIRK6I) dispatch to the B200 backend whenever the problem's function wraps a
# `BVPDeviceFunction`, so `solve(prob, MIRK4(); dt = 0.05)` and
# `solve(EnsembleProblem(prob; prob_func), MIRK4(), EnsembleB200(); trajectories, dt)` need no new algorithm type.
#
# The methods are more specific than the reference's `SciMLBase.__init(prob::BVProblem, alg::AbstractMIRK; ...)`
# (lib/BoundaryValueDiffEqMIRK/src/mirk.jl:49-53) in the problem's function type only.  `DeviceBVPFunction` spells the
# position of the user function among BVPFunction's type parameters (`BVPFunction{iip, specialize, twopoint, F, …}`)
# and `DeviceBVProblem` the position of the function among BVProblem's (`BVProblem{uType, tType, iip, nlls, P, F, …}`)
# as of SciMLBase 2.x / 3.x — check both against the installed SciMLBase when bumping its major version.
# NOT EXECUTED in this repository (no julia here).
module BoundaryValueDiffEqMIRKB200MIRKExt

using SciMLBase
import BoundaryValueDiffEqMIRK
import BoundaryValueDiffEqMIRK: AbstractMIRK, alg_order
import BoundaryValueDiffEqMIRKB200 as B200

const DeviceBVPFunction = SciMLBase.BVPFunction{<:Any, <:Any, <:Any, <:B200.BVPDeviceFunction}
const DeviceBVProblem = SciMLBase.BVProblem{<:Any, <:Any, <:Any, <:Any, <:Any, <:DeviceBVPFunction}

# the C ABI's tableau code: the convergence order, except MIRK6I = 7
tableau_code(alg::AbstractMIRK) = alg isa BoundaryValueDiffEqMIRK.MIRK6I ? 7 : alg_order(alg)

SciMLBase.__init(prob::DeviceBVProblem, alg::AbstractMIRK; device = 0, kwargs...) =
    B200.__init_b200(prob, alg, tableau_code(alg), device; kwargs...)

SciMLBase.__solve(ens::SciMLBase.AbstractEnsembleProblem, alg::AbstractMIRK, ealg::B200.EnsembleB200; device = 0, kwargs...) =
    B200.__solve_ensemble_b200(ens, alg, tableau_code(alg), device, ealg; kwargs...)

end
