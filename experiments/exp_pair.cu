// scratch experiment: k_reduce_pair<32> against k_reduce_generic on synthetic relations; not part of the library
#include <cstdio>
#include <vector>
#include <random>
#include <cmath>
#include "abd_pair.cuh"
using namespace mirk;
int main() {
    constexpr int n = 32; const size_t nn = n * n;
    const int R = 800, chunk = 8, G = R / chunk;
    std::vector<double> hL(R * nn), hR(R * nn), hr(R * n); std::vector<int> hn(R + 1), hg(G + 1);
    std::mt19937_64 g(7); std::uniform_real_distribution<double> U(-0.05, 0.05);
    for (int k = 0; k < R; k++) for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        hL[k * nn + i * n + j] = (i == j ? -1.0 : 0.0) + U(g); hR[k * nn + i * n + j] = (i == j ? 1.0 : 0.0) + U(g); }
    for (auto& x : hr) x = U(g);
    for (int i = 0; i <= R; i++) hn[i] = i;
    for (int i = 0; i <= G; i++) hg[i] = i * chunk;
    double *L, *Rr, *r, *o[2][3], *T[2][3], *scratch; int *nodes, *gs, *status;
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
    const int rows = 2 * n, cols = 3 * n + 1;
    const int smem = (int)(sizeof(double) * ((size_t)rows * cols + rows + cols) + sizeof(int) * (rows + 2 * n + 4));
    cudaFuncSetAttribute(k_reduce_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms0, ms1;
    cudaEventRecord(e0);
    k_reduce_generic<<<G, 256, smem>>>(n, L, Rr, r, o[0][0], o[0][1], o[0][2], nodes, gs, T[0][0], T[0][1], T[0][2], nullptr, 1, status);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms0, e0, e1);
    cudaEventRecord(e0);
    k_reduce_pair<n><<<G, 64>>>(L, Rr, r, o[1][0], o[1][1], o[1][2], nodes, gs, T[1][0], T[1][1], T[1][2], status + 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
    int st[2]; cudaMemcpy(st, status, 8, cudaMemcpyDeviceToHost);
    printf("generic %.1f us status %d | pair %.1f us status %d | %s\n", ms0 * 1e3, st[0], ms1 * 1e3, st[1], cudaGetErrorString(cudaGetLastError()));
    // the collapsed relations may differ by a row permutation/scaling; compare through the factors (unique): TL, TR, rt
    const char* nm[3] = {"TL", "TR", "rt"};
    for (int a = 0; a < 3; a++) {
        size_t len = (a < 2 ? (R + 1) * nn : (size_t)(R + 1) * n);
        std::vector<double> x(len), y(len);
        cudaMemcpy(x.data(), T[0][a], 8 * len, cudaMemcpyDeviceToHost); cudaMemcpy(y.data(), T[1][a], 8 * len, cudaMemcpyDeviceToHost);
        double md = 0, mx = 0; int nan = 0;
        for (size_t i = 0; i < len; i++) { if (!(y[i] == y[i])) nan++; md = fmax(md, fabs(x[i] - y[i])); mx = fmax(mx, fabs(x[i])); }
        printf("%s: max |generic - pair| = %.3e (max |generic| %.3e), NaNs in pair %d\n", nm[a], md, mx, nan);
    }
    return 0;
}
