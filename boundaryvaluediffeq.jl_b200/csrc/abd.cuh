// abd.cuh — parallel solve of the almost-block-diagonal Newton system
//
//     [ B_k ... (boundary rows, any pinned nodes) ] [d_1]   [bc ]
//     [ L_1 R_1                                   ] [d_2] = [Phi]
//     [     L_2 R_2  ...                          ] [ : ]
//
// The reference hands this matrix to LinearSolve (almost-banded QR / banded LU, one core;
// call site lib/BoundaryValueDiffEqCore/src/default_internal_solve.jl:107-110).  Here it is a
// stable block cyclic reduction in the style of Wright's structured elimination: a *relation*
// Lc d_a + Rc d_b = rc couples two nodes; a group of consecutive relations is collapsed to one by
// eliminating its interior nodes with row-pivoted Gauss-Jordan on the stacked 2n x n block
// (pivoting between the two stacked halves is what keeps dichotomic problems stable); the
// eliminated node c keeps   d_c = rt - TL d_a - TR d_right   for the back substitution.  Nodes the
// boundary condition touches are never eliminated; the few that survive are closed with the
// boundary rows in one small dense solve.
//
// This file holds the size-generic path (any n): working matrix in shared memory when
// 2n(3n+1) doubles fit, otherwise in a global scratch slab.  The register-resident warp path for
// n <= 16 is in abd_warp.cuh.
#pragma once
#include <cuda_runtime.h>

#include "warp_dense.cuh"

namespace mirk {

// warp argmax of (value, index); lower index wins ties
__device__ __forceinline__ void warp_argmax(double& v, int& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// Row-pivoted Gauss-Jordan on the first `ne` columns of W (rows x cols, pitch `ld`), restricted
// to rows with elig[r] != 0; on return pivrow[q] is the row that owns unit column q.
// All threads of the block call this.  Returns false (uniformly) on a zero / non-finite pivot.
__device__ bool block_gauss_jordan(double* W, int rows, int cols, int ld, int ne, int* elig,
                                   int* pivrow, double* mult, double* prow, int* s_p) {
    const int tid = threadIdx.x, T = blockDim.x;
    for (int q = 0; q < ne; q++) {
        if (tid < 32) {
            double best = -1.0;
            int bi = 0x7fffffff;
            for (int r = tid; r < rows; r += 32) {
                if (elig[r]) {
                    const double a = fabs(W[(size_t)r * ld + q]);
                    if (a > best || !(a == a)) { best = (a == a) ? a : INFINITY; bi = r; }
                }
            }
            warp_argmax(best, bi);
            if (tid == 0) *s_p = (best > 0.0 && best < INFINITY) ? bi : -1;
        }
        __syncthreads();
        const int pr = *s_p;
        if (pr < 0) return false;
        const double inv = 1.0 / W[(size_t)pr * ld + q];
        for (int r = tid; r < rows; r += T) mult[r] = W[(size_t)r * ld + q];
        for (int c = q + tid; c < cols; c += T) prow[c] = W[(size_t)pr * ld + c] * inv;
        __syncthreads();
        const int nc = cols - q;  // columns q..cols-1
        for (int e = tid; e < rows * nc; e += T) {
            const int r = e / nc, c = q + e % nc;
            double* w = &W[(size_t)r * ld + c];
            if (r == pr) *w = prow[c];
            else *w = (c == q) ? 0.0 : (*w - mult[r] * prow[c]);
        }
        if (tid == 0) { elig[pr] = 0; pivrow[q] = pr; }
        __syncthreads();
    }
    return true;
}

// One level of the reduction.  Block g collapses relations [gs[g], gs[g+1]) of the input level.
// Relation storage: L[k][n][n], R[k][n][n] row-major, r[k][n].  nodes[k], nodes[k+1] are the
// global node ids relation k couples.
__global__ void __launch_bounds__(256)
k_reduce_generic(int n, const double* __restrict__ inL, const double* __restrict__ inR,
                 const double* __restrict__ inr, double* __restrict__ outL, double* __restrict__ outR,
                 double* __restrict__ outr, const int* __restrict__ nodes, const int* __restrict__ gs,
                 double* __restrict__ TL, double* __restrict__ TR, double* __restrict__ rt,
                 double* __restrict__ scratch, int use_smem, int* __restrict__ status) {
    extern __shared__ double smem[];
    const int tid = threadIdx.x, T = blockDim.x;
    const int g = blockIdx.x, k0 = gs[g], k1 = gs[g + 1];
    const int rows = 2 * n, cols = 3 * n + 1, ld = cols;
    const size_t nn = (size_t)n * n;
    if (k1 - k0 == 1) {  // nothing to eliminate: pass the relation through
        for (int e = tid; e < (int)nn; e += T) {
            outL[g * nn + e] = inL[k0 * nn + e];
            outR[g * nn + e] = inR[k0 * nn + e];
        }
        for (int e = tid; e < n; e += T) outr[(size_t)g * n + e] = inr[(size_t)k0 * n + e];
        return;
    }
    double* W = use_smem ? smem : scratch + (size_t)g * rows * ld;
    double* mult = use_smem ? smem + (size_t)rows * ld : smem;
    double* prow = mult + rows;
    int* elig = (int*)(prow + cols);
    int* pivrow = elig + rows;
    int* rowlist = pivrow + n;   // rows that receive the next relation
    int* s_p = rowlist + n;

    // carried rows <- relation k0 :  [E | A | B | rhs] = [R | L | 0 | r]
    for (int e = tid; e < n * cols; e += T) {
        const int q = e / cols, c = e % cols;
        double v;
        if (c < n) v = inR[k0 * nn + (size_t)q * n + c];
        else if (c < 2 * n) v = inL[k0 * nn + (size_t)q * n + (c - n)];
        else if (c < 3 * n) v = 0.0;
        else v = inr[(size_t)k0 * n + q];
        W[(size_t)q * ld + c] = v;
    }
    for (int q = tid; q < n; q += T) rowlist[q] = n + q;
    __syncthreads();
    for (int j = k0 + 1; j < k1; j++) {
        // incoming relation j into the free rows: [E | A | B | rhs] = [L | 0 | R | r]
        for (int e = tid; e < n * cols; e += T) {
            const int q = e / cols, c = e % cols;
            double v;
            if (c < n) v = inL[j * nn + (size_t)q * n + c];
            else if (c < 2 * n) v = 0.0;
            else if (c < 3 * n) v = inR[j * nn + (size_t)q * n + (c - 2 * n)];
            else v = inr[(size_t)j * n + q];
            W[(size_t)rowlist[q] * ld + c] = v;
        }
        for (int r = tid; r < rows; r += T) elig[r] = 1;
        __syncthreads();
        if (!block_gauss_jordan(W, rows, cols, ld, n, elig, pivrow, mult, prow, s_p)) {
            if (tid == 0) atomicExch(status, 1);
            return;
        }
        // factors of the eliminated node: d_c = rt - TL d_a - TR d_right
        const int c_node = nodes[j];
        for (int e = tid; e < (int)nn; e += T) {
            const int q = e / n, c = e % n;
            const double* row = W + (size_t)pivrow[q] * ld;
            TL[(size_t)c_node * nn + e] = row[n + c];
            TR[(size_t)c_node * nn + e] = row[2 * n + c];
        }
        for (int q = tid; q < n; q += T) rt[(size_t)c_node * n + q] = W[(size_t)pivrow[q] * ld + 3 * n];
        __syncthreads();
        // survivors: E <- B, B <- 0 ; pivot rows become the free rows of the next merge
        for (int e = tid; e < rows * n; e += T) {
            const int r = e / n, c = e % n;
            if (elig[r]) {
                W[(size_t)r * ld + c] = W[(size_t)r * ld + 2 * n + c];
                W[(size_t)r * ld + 2 * n + c] = 0.0;
            }
        }
        for (int q = tid; q < n; q += T) rowlist[q] = pivrow[q];
        __syncthreads();
    }
    // emit the collapsed relation from the n surviving rows (rows not in rowlist)
    if (tid == 0) {
        for (int r = 0; r < rows; r++) elig[r] = 1;
        for (int q = 0; q < n; q++) elig[rowlist[q]] = 0;
        int cnt = 0;
        for (int r = 0; r < rows; r++) if (elig[r]) pivrow[cnt++] = r;
    }
    __syncthreads();
    for (int e = tid; e < (int)nn; e += T) {
        const int q = e / n, c = e % n;
        const double* row = W + (size_t)pivrow[q] * ld;
        outR[g * nn + e] = row[c];
        outL[g * nn + e] = row[n + c];
    }
    for (int q = tid; q < n; q += T) outr[(size_t)g * n + q] = W[(size_t)pivrow[q] * ld + 3 * n];
}

// Back substitution of one level: block g recovers the interior nodes of its group right to left.
__global__ void __launch_bounds__(256)
k_backsub_generic(int n, const int* __restrict__ nodes, const int* __restrict__ gs,
                  const double* __restrict__ TL, const double* __restrict__ TR,
                  const double* __restrict__ rt, double* __restrict__ delta) {
    const int g = blockIdx.x, k0 = gs[g], k1 = gs[g + 1];
    if (k1 - k0 == 1) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const size_t nn = (size_t)n * n;
    const double* da = delta + (size_t)nodes[k0] * n;
    for (int j = k1 - 1; j > k0; j--) {
        const int c = nodes[j];
        const double* dr = delta + (size_t)nodes[j + 1] * n;
        for (int q = warp; q < n; q += nw) {
            const double* tl = TL + (size_t)c * nn + (size_t)q * n;
            const double* tr = TR + (size_t)c * nn + (size_t)q * n;
            double acc = 0.0;
            for (int k = lane; k < n; k += 32) acc += tl[k] * da[k] + tr[k] * dr[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) delta[(size_t)c * n + q] = rt[(size_t)c * n + q] - acc;
        }
        __syncthreads();
    }
}

// Closing solve on the surviving nodes kept[0..Q): Q-1 relations + L boundary rows, dense
// Gauss-Jordan with row pivoting in a global scratch matrix M (D x (D+1), D = Q n).  One block.
__device__ void final_solve_body(int n, int Q, const int* kept, const double* relL, const double* relR,
                                 const double* relr, int L, int La, const int* m_ptr, const int* bc_nodes,
                                 const double* Bc, const double* resid, size_t tail_off, double* M, double* delta,
                                 int* status, double* smem) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int D = Q * n, cols = D + 1, ld = cols;
    const size_t nn = (size_t)n * n;
    double* mult = smem;
    double* prow = mult + D;
    int* elig = (int*)(prow + cols);
    int* pivrow = elig + D;
    int* s_p = pivrow + D;
    if (!M) M = (double*)(s_p + 4);  // the closing matrix fits in shared memory
    for (int e = tid; e < D * cols; e += T) M[e] = 0.0;
    __syncthreads();
    // boundary rows first (the reference's row order), accumulated per pinned node
    // (one thread per matrix entry: a thread per row would walk its n x m dependent load / add / store chain
    //  alone — ~15 us of pure latency at n = 16)
    const int m = *m_ptr;
    for (int e = tid; e < L * n; e += T) {
        const int q = e / n, c = e % n;
        for (int k = 0; k < m; k++) {
            int slot = -1;
            const int bn = bc_nodes[k];
            for (int s = 0; s < Q; s++) if (kept[s] == bn) slot = s;
            if (slot < 0) { atomicExch(status, 2); continue; }
            M[(size_t)q * ld + slot * n + c] += Bc[((size_t)k * L + q) * n + c];
        }
    }
    if (tid < L) M[(size_t)tid * ld + D] = tid < La ? resid[tid] : resid[tail_off + (tid - La)];
    for (int e = tid; e < (Q - 1) * (int)nn; e += T) {
        const int g = e / (int)nn, q = (e % (int)nn) / n, c = e % n;
        double* row = M + (size_t)(L + g * n + q) * ld;
        row[g * n + c] = relL[e];
        row[(g + 1) * n + c] = relR[e];
    }
    for (int e = tid; e < (Q - 1) * n; e += T) M[(size_t)(L + e) * ld + D] = relr[e];
    for (int r = tid; r < D; r += T) elig[r] = 1;
    __syncthreads();
    if (D <= 32) {
        // small closing systems (two-point problems: D = 2n <= 32; the pendulum: D = 6): one warp, one lane
        // per row in registers, REDUX pivot search and shuffle broadcast — no block barrier per pivot
        if (tid < 32) warp_dense_solve32(M, D, ld, kept, n, delta, status);
        return;
    }
    if (!block_gauss_jordan(M, D, cols, ld, D, elig, pivrow, mult, prow, s_p)) {
        if (tid == 0) atomicExch(status, 1);
        return;
    }
    for (int e = tid; e < D; e += T) {
        const int s = e / n, c = e % n;
        delta[(size_t)kept[s] * n + c] = M[(size_t)pivrow[e] * ld + D];
    }
}


// ---- mesh-partitioned mode: every rank reduces its mesh segment to ONE relation between its
// end nodes, the relations (+ the boundary blocks the end ranks own) are all-gathered, and every rank
// solves the small interface system redundantly.
// payload per rank: [relL n^2 | relR n^2 | relr n | Bc 2*L*n | bc rows L]
__host__ __device__ inline size_t part_payload_doubles(int n, int L) {
    return (size_t)2 * n * n + n + (size_t)2 * L * n + L;
}
__global__ void k_part_pack(int n, int L, int La, const double* __restrict__ relL, const double* __restrict__ relR,
                            const double* __restrict__ relr, const double* __restrict__ Bc,
                            const double* __restrict__ resid, size_t tail_off, double* __restrict__ out) {
    const int nn = n * n, tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    for (int e = tid; e < nn; e += T) { out[e] = relL[e]; out[nn + e] = relR[e]; }
    for (int e = tid; e < n; e += T) out[2 * nn + e] = relr[e];
    double* ob = out + 2 * nn + n;
    for (int e = tid; e < 2 * L * n; e += T) ob[e] = Bc[e];
    for (int q = tid; q < L; q += T) ob[2 * L * n + q] = q < La ? resid[q] : resid[tail_off + (q - La)];
}
// gathered payloads -> contiguous relation arrays of the interface system + its boundary blocks:
// block 0 (node 0) = rank 0's bc_a rows, block 1 (node G) = rank G-1's bc_b rows; if_resid = [bc_a ; bc_b]
// (standard != 0: a Standard problem, whose boundary rows couple both outer ends — every rank evaluated them on the
//  exchanged end states, so both blocks and all rows come from rank 0's payload)
__global__ void k_part_unpack(int n, int G, int L, int La, int standard, const double* __restrict__ recv, double* __restrict__ oL,
                              double* __restrict__ oR, double* __restrict__ orr, double* __restrict__ oBc,
                              double* __restrict__ oresid) {
    const int nn = n * n, tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    const size_t P = part_payload_doubles(n, L);
    for (int e = tid; e < G * nn; e += T) {
        const int g = e / nn, k = e % nn;
        oL[e] = recv[(size_t)g * P + k];
        oR[e] = recv[(size_t)g * P + nn + k];
    }
    for (int e = tid; e < G * n; e += T) orr[e] = recv[(size_t)(e / n) * P + 2 * nn + e % n];
    const double* b0 = recv + 2 * nn + n;
    const double* bG = recv + (size_t)(G - 1) * P + 2 * nn + n;
    for (int e = tid; e < L * n; e += T) {
        const int q = e / n;
        oBc[e] = q < La ? b0[e] : 0.0;                        // block 0: Bc[0][q][c]
        oBc[L * n + e] = standard ? b0[L * n + e] : (q < La ? 0.0 : bG[L * n + e]);  // block 1: Bc[1][q][c]
    }
    for (int q = tid; q < L; q += T) oresid[q] = q < La ? b0[2 * L * n + q] : bG[2 * L * n + q];
}
// ---- the same exchange over NVLink peer memory (no library collective) ---------------------------------
// Every rank owns one exchange buffer `xbuf` (cudaMalloc + CUDA IPC, mapped by all peers):
//     payload slots [2][G][P] doubles | norm slots [2][G][4] words | payload flags [2][G] | norm flags [2][G]
// A rank PUSHES its packed relation straight into slot [parity][rank] of every peer's buffer (remote stores
// through NVSwitch), fences system-wide and then stores the epoch into the peers' flag [parity][rank]; the
// consumer spins on its OWN (local) flags, so waiting costs no fabric traffic.  The pack of the relation and the
// all-gather are one kernel, the wait and the unpack into the interface system another; epochs live in device
// memory, so both replay inside CUDA graphs.  Double buffering by epoch parity is enough: a peer can only be one
// exchange ahead, because its next push waits on this rank's flag for the current one.
constexpr int kMaxPeers = 16;
struct XchgLayout {
    int G;
    size_t P;
    int n;  // states: sizes the two "ghost" slots [t, y(t)] of the outer ends (Standard problems: bc couples both ends)
    __host__ __device__ size_t off_nslot() const { return (size_t)2 * G * P; }
    __host__ __device__ size_t off_pflag() const { return off_nslot() + (size_t)2 * G * 4; }
    __host__ __device__ size_t off_nflag() const { return off_pflag() + (size_t)2 * G; }
    __host__ __device__ size_t off_gslot() const { return off_nflag() + (size_t)2 * G; }          // [parity][end][1 + n]
    __host__ __device__ size_t off_gflag() const { return off_gslot() + (size_t)2 * 2 * (n + 1); }  // [parity][end]
    __host__ __device__ size_t total() const { return off_gflag() + (size_t)2 * 2; }
};
struct XchgPeers {
    double* buf[kMaxPeers];
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= epoch; gives up after ~2 s of SM clocks (a peer that died must not hang the device)
__device__ __forceinline__ bool xchg_wait(const unsigned long long* flag, unsigned long long epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > 4000000000ll) return false;
        __nanosleep(64);
    }
    return true;
}

// pack this segment's collapsed relation (+ the boundary blocks and rows it owns) and push it to every rank
// (body: one whole block, any size)
__device__ __forceinline__ void part_push_body(int n, int L, int La, const double* __restrict__ relL, const double* __restrict__ relR,
                                               const double* __restrict__ relr, const double* __restrict__ Bc,
                                               const double* __restrict__ resid, size_t tail_off, const XchgPeers& peers,
                                               const XchgLayout& lay, int rank, const unsigned long long* __restrict__ epoch_ptr) {
    const int nn = n * n, tid = threadIdx.x, T = blockDim.x;
    const unsigned long long e = *epoch_ptr + 1ull;
    const size_t slot = ((size_t)(e & 1ull) * lay.G + rank) * lay.P;
    const int tot = (int)lay.P;
    for (int i = tid; i < tot; i += T) {
        double v;
        if (i < nn) v = relL[i];
        else if (i < 2 * nn) v = relR[i - nn];
        else if (i < 2 * nn + n) v = relr[i - 2 * nn];
        else if (i < 2 * nn + n + 2 * L * n) v = Bc[i - 2 * nn - n];
        else {
            const int q = i - 2 * nn - n - 2 * L * n;
            v = q < La ? resid[q] : resid[tail_off + (q - La)];
        }
        for (int r = 0; r < lay.G; r++) peers.buf[r][slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < lay.G)
        st_release_sys(reinterpret_cast<unsigned long long*>(peers.buf[tid] + lay.off_pflag()) + (e & 1ull) * lay.G + rank, e);
}
__global__ void __launch_bounds__(1024)
k_part_push(int n, int L, int La, const double* __restrict__ relL, const double* __restrict__ relR,
            const double* __restrict__ relr, const double* __restrict__ Bc, const double* __restrict__ resid,
            size_t tail_off, XchgPeers peers, XchgLayout lay, int rank, const unsigned long long* __restrict__ epoch_ptr) {
    part_push_body(n, L, La, relL, relR, relr, Bc, resid, tail_off, peers, lay, rank, epoch_ptr);
}
// wait for every rank's relation of this epoch, then unpack into the interface system (as k_part_unpack)
// (body: one whole block, any size >= G threads)
__device__ __forceinline__ void part_wait_unpack_body(int n, int L, int La, int standard, double* __restrict__ xbuf, const XchgLayout& lay,
                                                      double* __restrict__ oL, double* __restrict__ oR, double* __restrict__ orr,
                                                      double* __restrict__ oBc, double* __restrict__ oresid,
                                                      unsigned long long* __restrict__ epoch_ptr, int* __restrict__ status) {
    const int nn = n * n, tid = threadIdx.x, T = blockDim.x, G = lay.G;
    const unsigned long long e = *epoch_ptr + 1ull;
    if (tid < G) {
        const unsigned long long* fl = reinterpret_cast<const unsigned long long*>(xbuf + lay.off_pflag()) + (e & 1ull) * G + tid;
        if (!xchg_wait(fl, e)) atomicExch(status, 3);
    }
    __syncthreads();
    const double* recv = xbuf + (size_t)(e & 1ull) * G * lay.P;
    const size_t P = lay.P;
    for (int i = tid; i < G * nn; i += T) {
        const int g = i / nn, k = i % nn;
        oL[i] = __ldcg(recv + (size_t)g * P + k);  // peer-written data: read through L2
        oR[i] = __ldcg(recv + (size_t)g * P + nn + k);
    }
    for (int i = tid; i < G * n; i += T) orr[i] = __ldcg(recv + (size_t)(i / n) * P + 2 * nn + i % n);
    const double* b0 = recv + 2 * nn + n;
    const double* bG = recv + (size_t)(G - 1) * P + 2 * nn + n;
    for (int i = tid; i < L * n; i += T) {
        const int q = i / n;
        oBc[i] = q < La ? __ldcg(b0 + i) : 0.0;
        oBc[L * n + i] = standard ? __ldcg(b0 + L * n + i) : (q < La ? 0.0 : __ldcg(bG + L * n + i));
    }
    for (int q = tid; q < L; q += T) oresid[q] = q < La ? __ldcg(b0 + 2 * L * n + q) : __ldcg(bG + 2 * L * n + q);
    __syncthreads();
    if (tid == 0) *epoch_ptr = e;
}
__global__ void __launch_bounds__(1024)
k_part_wait_unpack(int n, int L, int La, int standard, double* __restrict__ xbuf, XchgLayout lay, double* __restrict__ oL,
                   double* __restrict__ oR, double* __restrict__ orr, double* __restrict__ oBc,
                   double* __restrict__ oresid, unsigned long long* __restrict__ epoch_ptr, int* __restrict__ status) {
    part_wait_unpack_body(n, L, La, standard, xbuf, lay, oL, oR, orr, oBc, oresid, epoch_ptr, status);
}
// all-reduce(max) of words[0..3) (|F|_inf bits, defect bits, status) over the ranks: push, wait, reduce — one warp
// (bc_* != nullptr: the rows of the boundary conditions this rank owns are folded into words[0] first — the work of
//  k_bc_norm_masked — so the norm exchange is one launch)
__global__ void __launch_bounds__(32)
k_words_allmax(unsigned long long* __restrict__ words, double* __restrict__ xbuf, XchgPeers peers, XchgLayout lay, int rank,
               unsigned long long* __restrict__ epoch_ptr, int* __restrict__ status, const double* __restrict__ bc_resid, int L,
               int La, size_t tail_off, int own_a, int own_b) {
    const int lane = threadIdx.x, G = lay.G;
    if (bc_resid) {
        unsigned long long m = 0ull;
        for (int q = lane; q < L; q += 32) {
            const bool is_a = q < La;
            if ((is_a && own_a) || (!is_a && own_b)) {
                const double v = is_a ? bc_resid[q] : bc_resid[tail_off + (q - La)];
                const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
                m = b > m ? b : m;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
            m = t > m ? t : m;
        }
        if (lane == 0 && m > words[0]) words[0] = m;
        __syncwarp();
    }
    const unsigned long long e = *epoch_ptr + 1ull, par = e & 1ull;
    unsigned long long w0 = 0ull, w1 = 0ull, w2 = 0ull;
    if (lane < G) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers.buf[lane] + lay.off_nslot()) + (par * G + rank) * 4;
        dst[0] = words[0]; dst[1] = words[1]; dst[2] = words[2];
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(peers.buf[lane] + lay.off_nflag()) + par * G + rank, e);
        const unsigned long long* fl = reinterpret_cast<const unsigned long long*>(xbuf + lay.off_nflag()) + par * G + lane;
        if (!xchg_wait(fl, e)) atomicExch(status, 3);
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(xbuf + lay.off_nslot()) + (par * G + lane) * 4;
        w0 = __ldcg(src); w1 = __ldcg(src + 1); w2 = __ldcg(src + 2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xffffffffu, w0, o), b = __shfl_xor_sync(0xffffffffu, w1, o),
                                 c = __shfl_xor_sync(0xffffffffu, w2, o);
        w0 = a > w0 ? a : w0; w1 = b > w1 ? b : w1; w2 = c > w2 ? c : w2;
    }
    if (lane == 0) { words[0] = w0; words[1] = w1; words[2] = w2; *epoch_ptr = e; }
}

// ---- Standard problems in the mesh-partitioned mode: bc!(res, sol, p, t) couples BOTH outer ends, so before every
// boundary evaluation rank 0 hands [t_first, y(t_first)] and the last rank [t_last, y(t_last)] to everybody; each rank
// then evaluates the boundary rows and their two blocks on that two-node "ghost" mesh (identical results everywhere).
// Peer-memory flavour: one warp; pushes into every peer's ghost slot, waits for both slots of this epoch, copies them out.
__global__ void __launch_bounds__(32)
k_part_ends_bcast(int n, int N, const double* __restrict__ mesh, const double* __restrict__ y, double* __restrict__ xbuf,
                  XchgPeers peers, XchgLayout lay, int rank, unsigned long long* __restrict__ epoch_ptr, int* __restrict__ status,
                  double* __restrict__ gmesh, double* __restrict__ gy) {
    const int lane = threadIdx.x, G = lay.G;
    const unsigned long long e = *epoch_ptr + 1ull, par = e & 1ull;
    for (int end = 0; end < 2; end++) {
        if (rank != (end ? G - 1 : 0)) continue;
        const int node = end ? N - 1 : 0;
        if (lane < G) {
            double* dst = peers.buf[lane] + lay.off_gslot() + (par * 2 + end) * (size_t)(n + 1);
            dst[0] = mesh[node];
            for (int c = 0; c < n; c++) dst[1 + c] = y[(size_t)node * n + c];
            st_release_sys(reinterpret_cast<unsigned long long*>(peers.buf[lane] + lay.off_gflag()) + par * 2 + end, e);
        }
    }
    __syncwarp();
    if (lane < 2) {
        const unsigned long long* fl = reinterpret_cast<const unsigned long long*>(xbuf + lay.off_gflag()) + par * 2 + lane;
        if (!xchg_wait(fl, e)) atomicExch(status, 3);
    }
    __syncwarp();
    for (int i = lane; i < 2 * (n + 1); i += 32) {
        const int end = i / (n + 1), k = i % (n + 1);
        const double v = __ldcg(xbuf + lay.off_gslot() + (par * 2 + end) * (size_t)(n + 1) + k);
        if (k == 0) gmesh[end] = v; else gy[(size_t)end * n + k - 1] = v;
    }
    if (lane == 0) *epoch_ptr = e;
}
// NCCL flavour: every rank contributes [t_first, y_first, t_last, y_last]; after the all-gather the ghost mesh is rank 0's
// first half and the last rank's second half
__global__ void k_part_ends_pack(int n, int N, const double* __restrict__ mesh, const double* __restrict__ y, double* __restrict__ out) {
    for (int i = threadIdx.x; i < 2 * (n + 1); i += blockDim.x) {
        const int end = i / (n + 1), k = i % (n + 1), node = end ? N - 1 : 0;
        out[i] = k == 0 ? mesh[node] : y[(size_t)node * n + k - 1];
    }
}
__global__ void k_part_ends_unpack(int n, int G, const double* __restrict__ recv, double* __restrict__ gmesh, double* __restrict__ gy) {
    for (int i = threadIdx.x; i < 2 * (n + 1); i += blockDim.x) {
        const int end = i / (n + 1), k = i % (n + 1);
        const double v = recv[(size_t)(end ? G - 1 : 0) * 2 * (n + 1) + i];
        if (k == 0) gmesh[end] = v; else gy[(size_t)end * n + k - 1] = v;
    }
}

// |bc rows|_inf of the rows this rank owns (a-rows on rank 0, b-rows on the last rank) into norm_bits
__global__ void k_bc_norm_masked(int L, int La, const double* __restrict__ resid, size_t tail_off, int own_a, int own_b,
                                 unsigned long long* __restrict__ norm_bits) {
    unsigned long long m = 0ull;
    for (int q = threadIdx.x; q < L; q += blockDim.x) {
        const bool is_a = q < La;
        if ((is_a && own_a) || (!is_a && own_b)) {
            const double v = is_a ? resid[q] : resid[tail_off + (q - La)];
            const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
            m = b > m ? b : m;
        }
    }
    if (m) atomicMax(norm_bits, m);
}

__global__ void __launch_bounds__(1024)
k_final_solve(int n, int Q, const int* __restrict__ kept, const double* __restrict__ relL,
              const double* __restrict__ relR, const double* __restrict__ relr, int L, int La,
              const int* __restrict__ m_ptr, const int* __restrict__ bc_nodes,
              const double* __restrict__ Bc, const double* __restrict__ resid, size_t tail_off,
              double* M, double* __restrict__ delta, int* __restrict__ status) {
    extern __shared__ double smem[];
    final_solve_body(n, Q, kept, relL, relR, relr, L, La, m_ptr, bc_nodes, Bc, resid, tail_off, M, delta, status, smem);
}

}  // namespace mirk
