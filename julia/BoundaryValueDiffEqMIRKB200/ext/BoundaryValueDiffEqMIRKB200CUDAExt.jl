# CUDA.jl extension (weak dependency, loaded only when CUDA.jl is in the session): device arrays go to the library
# as DEVICE POINTERS — mirk_ensemble_set_inputs_device / mirk_set_mesh_guess_device / mirk_get_solution_device copy
# device-to-device on the solver's stream, no host round trip — and algorithms can be bound to a CuDevice.
# NOT EXECUTED in this repository (no julia here); the three entry points are exercised from Python with torch
# device pointers (tests/test_gpu_ensemble.py, tests/test_gpu_parity.py).
module BoundaryValueDiffEqMIRKB200CUDAExt

using CUDA
import BoundaryValueDiffEqMIRKB200 as B200
import BoundaryValueDiffEqMIRKB200: libmirkb200, check

# `MIRK4(CUDA.device())`: the algorithm bound to a CUDA.jl device
for A in (:MIRK2, :MIRK3, :MIRK4, :MIRK5, :MIRK6, :MIRK6I)
    @eval B200.$A(dev::CuDevice; kw...) = B200.$A(; device = CUDA.deviceid(dev), kw...)
end

# a parameter sweep that already lives on the GPU (np x ntraj CuMatrix) stays there
B200.stage_params(p::CuMatrix{Float64}) = p
function B200.set_inputs!(h, params::CuMatrix{Float64}, u0::Vector{Float64})
    d_u0 = CuArray(u0)
    GC.@preserve params d_u0 check(ccall((:mirk_ensemble_set_inputs_device, libmirkb200[]), Cint,
        (Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Int32), h, pointer(params), pointer(d_u0), 0))
end

"`set_guess!(cache, mesh, y::CuMatrix)`: replace the guess of an initialised cache from a device array (n x N, node-major)"
function B200.set_guess!(cache::B200.MIRKB200Cache, mesh::Vector{Float64}, y::CuMatrix{Float64})
    size(y) == (cache.n, length(mesh)) || throw(DimensionMismatch("y must be n x length(mesh)"))
    GC.@preserve y check(ccall((:mirk_set_mesh_guess_device, libmirkb200[]), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, CuPtr{Float64}), cache.handle, length(mesh), mesh, pointer(y)))
    return cache
end

"`solution!(y::CuMatrix, cache)`: sol.u into a device array without leaving the GPU"
function B200.solution!(y::CuMatrix{Float64}, cache::B200.MIRKB200Cache)
    GC.@preserve y check(ccall((:mirk_get_solution_device, libmirkb200[]), Cint, (Ptr{Cvoid}, CuPtr{Float64}),
        cache.handle, pointer(y)))
    return y
end

end
