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

// ---- small vector kernels of the line-search / trust-region fallbacks of the Newton polyalgorithm (failure paths
// only: simple mappings, deterministic reductions so that accept / reject decisions repeat from run to run) --------
__global__ void k_axpby(size_t len, double* __restrict__ out, double a, const double* __restrict__ x, double b,
                        const double* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = a * x[i] + b * y[i];
}
// out[0] = a . b, out[1] = min(a), out[2] = max(a): one block, fixed summation order
__global__ void __launch_bounds__(1024) k_dot_minmax(size_t len, const double* __restrict__ a, const double* __restrict__ b,
                                                     double* __restrict__ out) {
    __shared__ double sd[1024], smn[1024], smx[1024];
    double acc = 0.0, mn = INFINITY, mx = -INFINITY;
    for (size_t i = threadIdx.x; i < len; i += blockDim.x) {
        const double v = a[i];
        acc += v * b[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    sd[threadIdx.x] = acc; smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sd[threadIdx.x] += sd[threadIdx.x + o];
            smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
            smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sd[0]; out[1] = smn[0]; out[2] = smx[0]; }
}
// |delta|_2^2 and |y|_2^2 of a Newton step into out[0], out[1] (zeroed by the caller): feeds the stalled-step rule of
// the termination test (only compared against thresholds many orders away, so the atomic summation order is harmless)
__global__ void __launch_bounds__(256) k_step_norms(size_t len, const double* __restrict__ delta, const double* __restrict__ y,
                                                    double* __restrict__ out) {
    double a = 0.0, b = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        a += delta[i] * delta[i];
        b += y[i] * y[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
}
// residual-row index of boundary row q: Standard [bc; Phi], TwoPoint [bc_a; Phi; bc_b]
__device__ __forceinline__ size_t bc_row(int q, int La, int N, int n) {
    return q < La ? (size_t)q : (size_t)La + (size_t)(N - 1) * n + (q - La);
}
// out = J v on the block structure (Jacobian blocks of the last evaluation): one thread per residual row
__global__ void k_jvec(int n, int N, int L, int La, const double* __restrict__ Lb, const double* __restrict__ Rb,
                       const int* __restrict__ m_ptr, const int* __restrict__ bc_nodes, const double* __restrict__ Bc,
                       const double* __restrict__ v, double* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nphi = (size_t)(N - 1) * n;
    if (e < nphi) {
        const size_t i = e / n;
        const int r = (int)(e % n);
        const double *lr = Lb + (i * n + r) * n, *rr = Rb + (i * n + r) * n, *vi = v + i * n;
        double acc = 0.0;
        for (int j = 0; j < n; j++) acc += lr[j] * vi[j] + rr[j] * vi[n + j];
        out[La + e] = acc;
    } else if (e < nphi + L) {
        const int q = (int)(e - nphi), m = *m_ptr;
        double acc = 0.0;
        for (int k = 0; k < m; k++)
            for (int j = 0; j < n; j++) acc += Bc[((size_t)k * L + q) * n + j] * v[(size_t)bc_nodes[k] * n + j];
        out[bc_row(q, La, N, n)] = acc;
    }
}
// out = J^T w: one thread per unknown (node i, component j)
__global__ void k_jtvec(int n, int N, int L, int La, const double* __restrict__ Lb, const double* __restrict__ Rb,
                        const int* __restrict__ m_ptr, const int* __restrict__ bc_nodes, const double* __restrict__ Bc,
                        const double* __restrict__ w, double* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)N * n) return;
    const size_t i = e / n;
    const int j = (int)(e % n), m = *m_ptr;
    double acc = 0.0;
    if (i < (size_t)N - 1)
        for (int r = 0; r < n; r++) acc += Lb[(i * n + r) * n + j] * w[La + i * n + r];
    if (i > 0)
        for (int r = 0; r < n; r++) acc += Rb[((i - 1) * n + r) * n + j] * w[La + (i - 1) * n + r];
    for (int k = 0; k < m; k++)
        if ((size_t)bc_nodes[k] == i)
            for (int q = 0; q < L; q++) acc += Bc[((size_t)k * L + q) * n + j] * w[bc_row(q, La, N, n)];
    out[e] = acc;
}

__global__ void k_fill_nodes(int N, int n, const double* __restrict__ u0, double* __restrict__ y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)N * n) y[i] = u0[i % n];
}

// ---- global-error estimate (MIRK/src/adaptivity.jl:464-567) -------------------------------------------------------
// halve_sol: nodes copied, midpoints averaged (guess of the Richardson solve on the halved mesh)
__global__ void k_halve_sol(int n, int N, const double* __restrict__ y, double* __restrict__ y2) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)(2 * (N - 1) + 1) * n) return;
    const size_t j = e / n;
    const int k = (int)(e % n);
    y2[e] = (j & 1) ? (y[(j / 2 + 1) * n + k] + y[(j / 2) * n + k]) / 2.0 : y[(j / 2) * n + k];
}
// err = (y_high[stride * i] - y) / (1 + |y|) per node, and its max-norm per node into nm
__global__ void k_ge_node(int n, int N, int stride, const double* __restrict__ yh, const double* __restrict__ y,
                          double* __restrict__ err, double* __restrict__ nm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double m = 0.0;
    for (int k = 0; k < n; k++) {
        const double lo = y[(size_t)i * n + k];
        const double e = (yh[(size_t)i * stride * n + k] - lo) / (1.0 + fabs(lo));
        err[(size_t)i * n + k] = e;
        if (!(fabs(e) <= m)) m = fabs(e);
    }
    nm[i] = m;
}
// GE_subinterval!: interval i keeps node i's error vector if its norm is >= node i + 1's, else node i + 1's
__global__ void k_ge_pick(int n, int N, const double* __restrict__ err, const double* __restrict__ nm,
                          double* __restrict__ errors, double* __restrict__ est, unsigned long long* __restrict__ norm_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N - 1) return;
    const int pick = nm[i] >= nm[i + 1] ? i : i + 1;
    for (int k = 0; k < n; k++) errors[(size_t)i * n + k] = err[(size_t)pick * n + k];
    est[i] = nm[pick];
    atomicMax(norm_bits, (unsigned long long)__double_as_longlong(nm[pick]));
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
// The four controllers' selectors differ only in the exponent 1 / expo_den (order + 1; order for GlobalErrorControl),
// the halving threshold rho (1 for DefectControl, 2 otherwise) and, for HybridErrorControl, s_hat being the sum of the
// powers of two estimates (est2 != nullptr).
__global__ void __launch_bounds__(1024)
k_mesh_select(int expo_den, double rho, int N, const double* __restrict__ mesh, double* __restrict__ est,
              const double* __restrict__ est2, double abstol, int max_sub, int cap, double* __restrict__ mesh_new,
              int* __restrict__ out, int use_smem) {
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
    const double ex = 1.0 / (double)expo_den;
    for (int i = tid; i < ni; i += T) {
        double v = pow(est[i] / abstol, ex);
        if (est2) v = v + pow(est2[i] / abstol, ex);
        sh[i] = v;
    }
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
        if (r1 <= rho * r3) {
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
