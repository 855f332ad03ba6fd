// scratch experiment: latency / throughput of the warp-level ABD reduction (n = 16); not part of the library.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../boundaryvaluediffeq.jl_b200/csrc [-DMIRK_BCAST_SHFL] exp_abd.cu
#include <cstdio>
#include <vector>
#include <random>
#include "abd_warp.cuh"
using namespace mirk;
static void run(const char* tag, int R, int chunk, int wpb, int reps) {
    constexpr int n = 16; const size_t nn = n * n;
    const int G = (R + chunk - 1) / chunk;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); std::vector<int> hn(R + 1), hg(G + 1);
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    for (int i = 0; i <= R; i++) hn[i] = i;
    for (int i = 0; i <= G; i++) hg[i] = std::min(i * chunk, R);
    double *L, *Rr, *r, *oL, *oR, *orr, *TL, *TR, *rt; int *nodes, *gs, *status;
    cudaMalloc(&L, 8 * R * nn); cudaMalloc(&Rr, 8 * R * nn); cudaMalloc(&r, 8 * R * n);
    cudaMalloc(&oL, 8 * G * nn); cudaMalloc(&oR, 8 * G * nn); cudaMalloc(&orr, 8 * G * n);
    cudaMalloc(&TL, 8 * (R + 1) * nn); cudaMalloc(&TR, 8 * (R + 1) * nn); cudaMalloc(&rt, 8 * (R + 1) * n);
    cudaMalloc(&nodes, 4 * (R + 1)); cudaMalloc(&gs, 4 * (G + 1)); cudaMalloc(&status, 4); cudaMemset(status, 0, 4);
    cudaMemcpy(L, hL.data(), 8 * R * nn, cudaMemcpyHostToDevice); cudaMemcpy(Rr, hR.data(), 8 * R * nn, cudaMemcpyHostToDevice);
    cudaMemcpy(r, hr.data(), 8 * R * n, cudaMemcpyHostToDevice);
    cudaMemcpy(nodes, hn.data(), 4 * (R + 1), cudaMemcpyHostToDevice); cudaMemcpy(gs, hg.data(), 4 * (G + 1), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < reps; it++) {
        cudaEventRecord(e0);
        k_reduce_warp<n, 3><<<(G + wpb - 1) / wpb, 32 * wpb>>>(G, L, Rr, r, oL, oR, orr, nodes, gs, TL, TR, rt, status);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<double> ho(G * n); cudaMemcpy(ho.data(), orr, 8 * G * n, cudaMemcpyDeviceToHost);
    double cs = 0; for (double x : ho) cs += x;
    int st; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
    const int merges = chunk - 1;
    printf("%-28s R=%6d chunk=%2d G=%5d wpb=%d : %8.1f us  (%.2f us per merge-depth)  checksum %.12e status %d %s\n", tag, R, chunk, G, wpb,
           best * 1e3, best * 1e3 / merges, cs, st, cudaGetErrorString(cudaGetLastError()));
}
int main() {
#ifdef MIRK_BCAST_SHFL
    const char* v = "shfl";
#else
    const char* v = "smem";
#endif
    printf("variant %s\n", v);
    run("level0", 20000, 8, 4, 5);
    run("level0 chunk12", 20000, 12, 4, 5);
    run("level1 (4 warps/CTA)", 2504, 8, 4, 5);
    run("level1 (1 warp/CTA)", 2504, 8, 1, 5);
    run("one warp per SM", 148 * 8, 8, 1, 5);
    run("one warp total", 8, 8, 1, 5);
    run("one warp, 1 merge", 2, 2, 1, 5);
    return 0;
}
