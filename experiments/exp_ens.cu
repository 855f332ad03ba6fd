// scratch experiment: occupancy of the ensemble kernel (pendulum, MIRK4); not part of the library
#include <cstdio>
#include <vector>
#include <cmath>
#include <random>
#include "ensemble.cuh"
#include "problems.cuh"
using namespace mirk;
template <int MINB> float run(int ntraj, int reps) {
    using P = problems::Pendulum;
    using LY = EnsLayout<P, 4>;
    const int NC = 128, N0 = 33;
    long long stride = (ntraj + 31) / 32 * 32;
    double *work, *params, *u0, *mesh0, *rn, *dn; int *rc, *nm, *ni, *no;
    cudaMalloc(&work, sizeof(double) * LY::slots_per_node * NC * stride);
    cudaMalloc(&params, sizeof(double) * ntraj); cudaMalloc(&u0, 16); cudaMalloc(&mesh0, sizeof(double) * N0);
    cudaMalloc(&rn, sizeof(double) * ntraj); cudaMalloc(&dn, sizeof(double) * ntraj);
    cudaMalloc(&rc, 4 * ntraj); cudaMalloc(&nm, 4 * ntraj); cudaMalloc(&ni, 4 * ntraj); cudaMalloc(&no, 4 * ntraj);
    std::vector<double> hp(ntraj), hm(N0); std::mt19937_64 g(1); std::uniform_real_distribution<double> U(8, 12);
    for (auto& x : hp) x = U(g);
    const double T = M_PI / 2; for (int i = 0; i < N0; i++) hm[i] = T * i / (N0 - 1);
    double hu[2] = {M_PI / 2, M_PI / 2};
    cudaMemcpy(params, hp.data(), sizeof(double) * ntraj, cudaMemcpyHostToDevice);
    cudaMemcpy(mesh0, hm.data(), sizeof(double) * N0, cudaMemcpyHostToDevice); cudaMemcpy(u0, hu, 16, cudaMemcpyHostToDevice);
    EnsArgs a{ntraj, stride, NC, N0, mesh0, params, u0, 0, 1e-6, 0.1, 1, 3000, 1000, 0, 100, work, rc, nm, ni, no, rn, dn};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        k_ensemble_solve<P, 4, MINB><<<(ntraj + 63) / 64, 64>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    std::vector<int> hrc(ntraj), hni(ntraj); cudaMemcpy(hrc.data(), rc, 4 * ntraj, cudaMemcpyDeviceToHost); cudaMemcpy(hni.data(), ni, 4 * ntraj, cudaMemcpyDeviceToHost);
    long ok = 0, its = 0; for (int i = 0; i < ntraj; i++) { ok += hrc[i] == 0; its += hni[i]; }
    printf("MINB=%d ntraj=%d best %.2f ms  converged %ld  mean iters %.2f  err=%s\n", MINB, ntraj, best, ok, (double)its / ntraj, cudaGetErrorString(cudaGetLastError()));
    cudaFree(work); return best;
}
int main() { run<1>(262144, 3); run<6>(262144, 3); run<8>(262144, 3); run<10>(262144, 3); run<12>(262144, 3); return 0; }
