// scratch experiment: k_reduce_mma32r (panel factorisation replicated per warp, two barriers per panel) against
// k_reduce_mma32 (one barrier per pivot) on synthetic n = 32 relations; not part of the library.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../boundaryvaluediffeq.jl_b200/csrc -I../include -I.
//        -DMIRK_ELIM_PANEL -DMIRK_ABD_MMA exp_mma32r.cu -o bin/exp_mma32r
#include <cstdio>
#include <vector>
#include <random>
#include <cmath>
#include "abd_pair.cuh"
#include "abd_mma32r.cuh"
using namespace mirk;
static void run(int R, int chunk) {
    constexpr int n = 32; const size_t nn = n * n;
    const int G = R / chunk;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); std::vector<int> hn(R + 1), hg(G + 1);
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    for (int i = 0; i <= R; i++) hn[i] = i;
    for (int i = 0; i <= G; i++) hg[i] = i * chunk;
    double *L, *Rr, *r, *o[2][3], *T[2][3]; int *nodes, *gs, *status;
    cudaMalloc(&L, 8 * R * nn); cudaMalloc(&Rr, 8 * R * nn); cudaMalloc(&r, 8 * R * n);
    for (int v = 0; v < 2; v++) {
        cudaMalloc(&o[v][0], 8 * G * nn); cudaMalloc(&o[v][1], 8 * G * nn); cudaMalloc(&o[v][2], 8 * G * n);
        cudaMalloc(&T[v][0], 8 * (R + 1) * nn); cudaMalloc(&T[v][1], 8 * (R + 1) * nn); cudaMalloc(&T[v][2], 8 * (R + 1) * n);
        cudaMemset(T[v][0], 0, 8 * (R + 1) * nn); cudaMemset(T[v][1], 0, 8 * (R + 1) * nn); cudaMemset(T[v][2], 0, 8 * (R + 1) * n);
    }
    cudaMalloc(&nodes, 4 * (R + 1)); cudaMalloc(&gs, 4 * (G + 1)); cudaMalloc(&status, 8); cudaMemset(status, 0, 8);
    cudaMemcpy(L, hL.data(), 8 * R * nn, cudaMemcpyHostToDevice); cudaMemcpy(Rr, hR.data(), 8 * R * nn, cudaMemcpyHostToDevice);
    cudaMemcpy(r, hr.data(), 8 * R * n, cudaMemcpyHostToDevice);
    cudaMemcpy(nodes, hn.data(), 4 * (R + 1), cudaMemcpyHostToDevice); cudaMemcpy(gs, hg.data(), 4 * (G + 1), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms[2] = {1e9f, 1e9f};
    for (int rep = 0; rep < 3; rep++) {
        float m;
        cudaEventRecord(e0);
        k_reduce_mma32<<<G, 128>>>(L, Rr, r, o[0][0], o[0][1], o[0][2], nodes, gs, T[0][0], T[0][1], T[0][2], status);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&m, e0, e1); ms[0] = fminf(ms[0], m);
        cudaEventRecord(e0);
        k_reduce_mma32r<<<G, 128>>>(L, Rr, r, o[1][0], o[1][1], o[1][2], nodes, gs, T[1][0], T[1][1], T[1][2], status + 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&m, e0, e1); ms[1] = fminf(ms[1], m);
    }
    int st[2]; cudaMemcpy(st, status, 8, cudaMemcpyDeviceToHost);
    printf("R=%d chunk=%d G=%d: mma32 %.1f us (st %d) | mma32r %.1f us (st %d) | %s\n", R, chunk, G, ms[0] * 1e3, st[0], ms[1] * 1e3, st[1],
           cudaGetErrorString(cudaGetLastError()));
    const char* nm[6] = {"TL", "TR", "rt", "outL", "outR", "outr"};
    for (int a = 0; a < 6; a++) {
        const size_t len = a < 2 ? (R + 1) * nn : a == 2 ? (size_t)(R + 1) * n : a < 5 ? G * nn : (size_t)G * n;
        std::vector<double> x(len), y(len);
        cudaMemcpy(x.data(), a < 3 ? T[0][a] : o[0][a - 3], 8 * len, cudaMemcpyDeviceToHost);
        cudaMemcpy(y.data(), a < 3 ? T[1][a] : o[1][a - 3], 8 * len, cudaMemcpyDeviceToHost);
        double md = 0; long nan = 0;
        for (size_t i = 0; i < len; i++) { if (!(y[i] == y[i])) nan++; else md = fmax(md, fabs(x[i] - y[i])); }
        printf("  %s: max |mma32 - mma32r| = %.3e, NaNs %ld\n", nm[a], md, nan);
    }
    cudaFree(L); cudaFree(Rr); cudaFree(r); cudaFree(nodes); cudaFree(gs); cudaFree(status);
    for (int v = 0; v < 2; v++) for (int a = 0; a < 3; a++) { cudaFree(o[v][a]); cudaFree(T[v][a]); }
}
#if defined(MIRK_R32_PROF)
static void prof() {
    constexpr int n = 32; const size_t nn = n * n; const int R = 2, G = 1;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); int hn[3] = {0, 1, 2}, hg[2] = {0, 2};
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    double *L, *Rr, *r, *o0, *o1, *o2, *T0, *T1, *T2; int *nodes, *gs, *status;
    cudaMalloc(&L, 8 * R * nn); cudaMalloc(&Rr, 8 * R * nn); cudaMalloc(&r, 8 * R * n); cudaMalloc(&o0, 8 * nn); cudaMalloc(&o1, 8 * nn); cudaMalloc(&o2, 8 * n);
    cudaMalloc(&T0, 8 * 3 * nn); cudaMalloc(&T1, 8 * 3 * nn); cudaMalloc(&T2, 8 * 3 * n); cudaMalloc(&nodes, 12); cudaMalloc(&gs, 8); cudaMalloc(&status, 4);
    cudaMemcpy(L, hL.data(), 8 * R * nn, cudaMemcpyHostToDevice); cudaMemcpy(Rr, hR.data(), 8 * R * nn, cudaMemcpyHostToDevice);
    cudaMemcpy(r, hr.data(), 8 * R * n, cudaMemcpyHostToDevice); cudaMemcpy(nodes, hn, 12, cudaMemcpyHostToDevice); cudaMemcpy(gs, hg, 8, cudaMemcpyHostToDevice);
    int zero = 0;
    for (int it = 0; it < 3; it++) { cudaMemcpyToSymbol(g_r32_prof_n, &zero, 4); k_reduce_mma32r<<<1, 128>>>(L, Rr, r, o0, o1, o2, nodes, gs, T0, T1, T2, status); cudaDeviceSynchronize(); }
    long long h[96]; int cnt; cudaMemcpyFromSymbol(h, g_r32_prof, sizeof(h)); cudaMemcpyFromSymbol(&cnt, g_r32_prof_n, 4);
    printf("mma32r stamps per panel (A start | after bar1 | after B | after C,D | after bar2), then end of eliminate:\n");
    for (int i = 0; i < cnt; i++) printf(" %lld%s", h[i] - h[0], i % 5 == 4 ? " |" : "");
    printf("\n");
}
#endif
int main() {
#if defined(MIRK_R32_PROF)
    prof(); return 0;
#endif
    run(800, 8);
    run(16000, 8);
    run(250000, 8);
    run(296 * 16, 16);
    return 0;
}
