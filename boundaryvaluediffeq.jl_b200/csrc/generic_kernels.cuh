// generic_kernels.cuh — problem-independent helper kernels (included by the host TU only).
#pragma once
#include <cuda_runtime.h>

namespace mirk {

// quirk Q3 (MIRK/src/mirk.jl:368-370): the reference adds y0[i_old] from the array it is rewriting.
// inc[j] = h sum w K was computed in parallel by k_interp(add_base=0); this resolves the chain
// sequentially, one thread per component.
__global__ void k_reinterp_inplace_chain(int n, int N_old, int N_new, const double* __restrict__ y_old,
                                         const double* __restrict__ inc, const int* __restrict__ iold,
                                         double* __restrict__ y_new) {
    const int k = threadIdx.x;
    if (k >= n) return;
    for (int j = 0; j < N_new; j++) {
        const int i = iold[j];
        const double base = (i < j) ? y_new[(size_t)i * n + k] : y_old[(size_t)i * n + k];
        y_new[(size_t)j * n + k] = inc[(size_t)j * n + k] + base;
    }
}

// ---- Newton update y <- y - delta ----------------------------------------------------------------
__global__ void k_axpy_neg(size_t len, double* __restrict__ y, const double* __restrict__ d) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) y[i] -= d[i];
}

__global__ void k_fill_nodes(int N, int n, const double* __restrict__ u0, double* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)N * n) y[i] = u0[i % n];
}

// ---- half_mesh! (MIRK/src/adaptivity.jl:287-304) -------------------------------------------------
__global__ void k_half_mesh(int N, const double* __restrict__ mesh, double* __restrict__ mesh_new) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N - 1) {
        mesh_new[2 * i] = mesh[i];
        mesh_new[2 * i + 1] = (mesh[i + 1] + mesh[i]) / 2.0;
    } else if (i == N - 1) {
        mesh_new[2 * i] = mesh[i];
    }
}

// ---- mesh_selector!(cache, DefectControl) + redistribute! (MIRK/src/adaptivity.jl:23-75,250-278) --
// One block.  s_hat is computed in parallel; the sums and the equidistribution sweep are done by
// ONE thread in the reference's order (they decide an integer mesh size through a round(), and the
// sweep restarts its integral at every emitted node, so the order is part of the result).
// est[i] = |errors_i|_inf.  out[0] = new node count, out[1] = 0 success / 1 failure.
// With use_smem the mesh and s_hat live in shared memory (2 N doubles), else s_hat overwrites est.
__global__ void __launch_bounds__(1024)
k_mesh_select(int order, int N, const double* __restrict__ mesh, double* __restrict__ est, double abstol,
              int max_sub, int cap, double* __restrict__ mesh_new, int* __restrict__ out, int use_smem) {
    extern __shared__ double sm_sel[];
    __shared__ int s_mode, s_ns;
    const int ni = N - 1, tid = threadIdx.x, T = blockDim.x;
    double* sh = use_smem ? sm_sel : est;
    const double* ms = mesh;
    if (use_smem) {
        double* m2 = sm_sel + ni;
        for (int i = tid; i < N; i += T) m2[i] = mesh[i];
        ms = m2;
    }
    const double ex = 1.0 / (double)(order + 1);
    for (int i = tid; i < ni; i += T) sh[i] = pow(est[i] / abstol, ex);
    __syncthreads();
    if (tid == 0) {
        double r1 = 0.0, r2 = 0.0;
        for (int i = 0; i < ni; i++) {
            if (sh[i] > r1) r1 = sh[i];
            r2 += sh[i];
        }
        const double r3 = r2 / ni;
        long long n_predict = (long long)nearbyint(1.3 * r2 + 1.0);  // round(Int, .): half to even
        const double n_ = 0.1 * ni;
        if (fabs((double)(n_predict - ni)) < n_) n_predict = (long long)nearbyint(ni + n_);
        int mode, ns;
        if (r1 <= 1.0 * r3) {  // rho = 1.0
            ns = 2 * ni;
            mode = 1;
        } else {
            const long long lb = N / 2, ub = 4LL * ni;
            ns = (int)(n_predict < lb ? lb : (n_predict > ub ? ub : n_predict));
            mode = 2;
        }
        if (ns > max_sub || ns + 1 > cap) mode = 0;
        s_mode = mode;
        s_ns = ns;
        out[0] = mode ? ns + 1 : N;
        out[1] = mode ? 0 : 1;
    }
    __syncthreads();
    const int mode = s_mode, ns = s_ns;
    if (mode == 0) return;
    if (mode == 1) {
        for (int i = tid; i < N; i += T) {
            mesh_new[2 * i] = ms[i];
            if (i < ni) mesh_new[2 * i + 1] = (ms[i + 1] + ms[i]) / 2.0;
        }
        return;
    }
    // resize!(mesh, ns+1) keeps the leading old entries; the tail defaults to t_end here
    for (int i = tid; i <= ns; i += T) mesh_new[i] = (i < N) ? ms[i] : ms[ni];
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < ni; i++) {
            const double h = ms[i + 1] - ms[i];
            sh[i] /= h;
            tot += sh[i] * h;
        }
        const double zeta = tot / (double)ns;
        int k = 0;
        long long i = 0;
        double t = ms[0], integral = 0.0;
        mesh_new[0] = ms[0];
        while (k < ni) {
            const double next_piece = sh[k] * (ms[k + 1] - t);
            const double int_next = integral + next_piece;
            if (int_next > zeta) {
                const double tn = (zeta - integral) / sh[k] + t;
                if (i + 1 <= ns) mesh_new[i + 1] = tn;
                t = tn;
                i++;
                integral = 0.0;
            } else {
                integral = int_next;
                t = ms[k + 1];
                k++;
            }
        }
        mesh_new[ns] = ms[ni];
    }
}

}  // namespace mirk
