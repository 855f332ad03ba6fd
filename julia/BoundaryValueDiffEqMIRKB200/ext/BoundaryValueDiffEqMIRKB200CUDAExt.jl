# CUDA.jl extension: lets a caller that already holds device arrays hand them over without a host
# round trip, and selects the library's device from CUDA.jl's current device.  Loaded only when
# CUDA.jl is in the session (weak dependency).  NOT EXECUTED in this repository (no julia here).
module BoundaryValueDiffEqMIRKB200CUDAExt

using CUDA
import BoundaryValueDiffEqMIRKB200 as B200

"MIRK4B200 / MIRK6B200 bound to CUDA.jl's current device"
B200.MIRK4B200(dev::CuDevice; kw...) = B200.MIRK4B200(; device = CUDA.deviceid(dev), kw...)
B200.MIRK6B200(dev::CuDevice; kw...) = B200.MIRK6B200(; device = CUDA.deviceid(dev), kw...)

"parameter sweeps living in a CuArray are copied once to the host staging buffer the C ABI takes"
stage_params(p::CuArray{Float64}) = Array(p)

end
