// host_util.cpp — host-only helpers of the C ABI that need binary128 arithmetic (compiled by g++,
// nvcc's front end does not accept __float128 in .cu files).
#include "mirk_b200.h"

// collect(range(t0; stop = t1, length = nint + 1)) (lib/BoundaryValueDiffEqCore/src/utils.jl:694).
// Julia's range is a twice-precision StepRangeLen, i.e. effectively the correctly rounded
// t0 + i (t1 - t0) / nint; binary128 carries the extra bits here.
extern "C" void mirk_mesh_uniform_fill(double t0, double t1, int32_t nint, double* mesh) {
    const __float128 a = (__float128)t0, b = (__float128)t1;
    for (int32_t i = 0; i <= nint; i++) mesh[i] = (double)(a + ((b - a) * (__float128)i) / (__float128)nint);
    mesh[0] = t0;
    mesh[nint] = t1;
}
