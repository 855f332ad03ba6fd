// scratch experiment: k_reduce_mma32 (csrc/abd_mma32.cuh) against k_reduce_pair<32> and k_reduce_generic on
// synthetic relations; not part of the library.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../boundaryvaluediffeq.jl_b200/csrc -I../include
//        -DMIRK_ELIM_PANEL -DMIRK_ABD_MMA exp_mma32.cu -o bin/exp_mma32
#include <cstdio>
#include <vector>
#include <random>
#include <cmath>
#include "abd_pair.cuh"
#include "abd_mma32.cuh"
using namespace mirk;
static void run(int R, int chunk, bool with_generic) {
    constexpr int n = 32; const size_t nn = n * n;
    const int G = R / chunk;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); std::vector<int> hn(R + 1), hg(G + 1);
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    for (int i = 0; i <= R; i++) hn[i] = i;
    for (int i = 0; i <= G; i++) hg[i] = i * chunk;
    double *L, *Rr, *r, *o[3][3], *T[3][3]; int *nodes, *gs, *status;
    cudaMalloc(&L, 8 * R * nn); cudaMalloc(&Rr, 8 * R * nn); cudaMalloc(&r, 8 * R * n);
    for (int v = 0; v < 3; v++) {
        cudaMalloc(&o[v][0], 8 * G * nn); cudaMalloc(&o[v][1], 8 * G * nn); cudaMalloc(&o[v][2], 8 * G * n);
        cudaMalloc(&T[v][0], 8 * (R + 1) * nn); cudaMalloc(&T[v][1], 8 * (R + 1) * nn); cudaMalloc(&T[v][2], 8 * (R + 1) * n);
        cudaMemset(T[v][0], 0, 8 * (R + 1) * nn); cudaMemset(T[v][1], 0, 8 * (R + 1) * nn); cudaMemset(T[v][2], 0, 8 * (R + 1) * n);
    }
    cudaMalloc(&nodes, 4 * (R + 1)); cudaMalloc(&gs, 4 * (G + 1)); cudaMalloc(&status, 12); cudaMemset(status, 0, 12);
    cudaMemcpy(L, hL.data(), 8 * R * nn, cudaMemcpyHostToDevice); cudaMemcpy(Rr, hR.data(), 8 * R * nn, cudaMemcpyHostToDevice);
    cudaMemcpy(r, hr.data(), 8 * R * n, cudaMemcpyHostToDevice);
    cudaMemcpy(nodes, hn.data(), 4 * (R + 1), cudaMemcpyHostToDevice); cudaMemcpy(gs, hg.data(), 4 * (G + 1), cudaMemcpyHostToDevice);
    const int rows = 2 * n, cols = 3 * n + 1;
    const int smem = (int)(sizeof(double) * ((size_t)rows * cols + rows + cols) + sizeof(int) * (rows + 2 * n + 4));
    cudaFuncSetAttribute(k_reduce_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms[3] = {0, 1e9f, 1e9f};
    if (with_generic) {
        cudaEventRecord(e0);
        k_reduce_generic<<<G, 256, smem>>>(n, L, Rr, r, o[0][0], o[0][1], o[0][2], nodes, gs, T[0][0], T[0][1], T[0][2], nullptr, 1, status);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[0], e0, e1);
    }
    for (int rep = 0; rep < 3; rep++) {
        float m;
        cudaEventRecord(e0);
        k_reduce_pair<n><<<G, 64>>>(L, Rr, r, o[1][0], o[1][1], o[1][2], nodes, gs, T[1][0], T[1][1], T[1][2], status + 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&m, e0, e1); ms[1] = fminf(ms[1], m);
        cudaEventRecord(e0);
        k_reduce_mma32<<<G, 128>>>(L, Rr, r, o[2][0], o[2][1], o[2][2], nodes, gs, T[2][0], T[2][1], T[2][2], status + 2);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&m, e0, e1); ms[2] = fminf(ms[2], m);
    }
    int st[3]; cudaMemcpy(st, status, 12, cudaMemcpyDeviceToHost);
    printf("R=%d chunk=%d G=%d: generic %.1f us (st %d) | pair %.1f us (st %d) | mma32 %.1f us (st %d) | %s\n", R, chunk, G,
           ms[0] * 1e3, st[0], ms[1] * 1e3, st[1], ms[2] * 1e3, st[2], cudaGetErrorString(cudaGetLastError()));
    // the collapsed relations may differ by a row permutation; the factors are unique: compare TL, TR, rt with the pair kernel
    const char* nm[3] = {"TL", "TR", "rt"};
    for (int a = 0; a < 3; a++) {
        size_t len = (a < 2 ? (R + 1) * nn : (size_t)(R + 1) * n);
        std::vector<double> x(len), y(len);
        cudaMemcpy(x.data(), T[1][a], 8 * len, cudaMemcpyDeviceToHost); cudaMemcpy(y.data(), T[2][a], 8 * len, cudaMemcpyDeviceToHost);
        double md = 0, mx = 0; long nan = 0;
        for (size_t i = 0; i < len; i++) { if (!(y[i] == y[i])) nan++; else md = fmax(md, fabs(x[i] - y[i])); mx = fmax(mx, fabs(x[i])); }
        printf("  %s: max |pair - mma32| = %.3e (max |pair| %.3e), NaNs in mma32 %ld\n", nm[a], md, mx, nan);
    }
    // and the output relation's rhs checksum (row order is the same in both kernels)
    std::vector<double> ra(G * n), rb(G * n);
    cudaMemcpy(ra.data(), o[1][2], 8 * G * n, cudaMemcpyDeviceToHost); cudaMemcpy(rb.data(), o[2][2], 8 * G * n, cudaMemcpyDeviceToHost);
    double ca = 0, cb = 0; for (int i = 0; i < G * n; i++) { ca += ra[i]; cb += rb[i]; }
    printf("  relation rhs checksum pair %.12e mma32 %.12e\n", ca, cb);
    cudaFree(L); cudaFree(Rr); cudaFree(r); cudaFree(nodes); cudaFree(gs); cudaFree(status);
    for (int v = 0; v < 3; v++) for (int a = 0; a < 3; a++) { cudaFree(o[v][a]); cudaFree(T[v][a]); }
}
int main() {
    run(800, 8, true);
    run(16000, 8, false);
    return 0;
}
