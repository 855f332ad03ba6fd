// ensemble_ops.cuh — launch table entries of the ensemble kernels (included by the ops_ensemble_*.cu units, which
// only differ in the problems they instantiate so that they compile in parallel).
#pragma once
#include "ensemble.cuh"
#include "ensemble_warp.cuh"
#include "ops.cuh"

namespace mirk {
using namespace problems;

template <class P, int ORDER> struct EnsImpl {
    // thread per trajectory, state in an HBM slab (any n <= 6; also the overflow path of the warp kernel)
    static void run(cudaStream_t st, const EnsArgs& a) {
        // n <= 2: cap registers at 96 (10 CTAs of 64 threads per SM): the kernel is memory-latency bound and
        // the extra resident warps buy 17 % (experiments/exp_ens.cu); larger n needs the registers
        constexpr int MINB = P::n <= 2 ? 10 : 1;
        // small shards: one warp per CTA so the CTAs spread evenly over the SMs
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int tpb = (a.ntraj + 63) / 64 < 8LL * sms ? 32 : 64;
        const long long blocks = (a.ntraj + tpb - 1) / tpb;
        k_ensemble_solve<P, ORDER, MINB><<<(unsigned)blocks, tpb, 0, st>>>(a);
    }
    // warp per trajectory, state in shared memory (n <= 2): persistent CTAs pulling trajectories from a counter
    static constexpr bool kWarp = P::n <= 2;
    // dynamic shared memory of ONE warp at capacity NC (a CTA holds 1, 2 or 4 warps, whatever fits)
    static size_t warp_smem_bytes(int NC) {
        if constexpr (kWarp) return sizeof(double) * EnsWarpLayout<P, ORDER>::warp_doubles(NC);
        else return 0;
    }
    static cudaError_t run_warp(cudaStream_t st, const EnsWarpArgs& w) {
        if constexpr (kWarp) {
            int wpb = kEnsWarpsPerBlock;
            while (wpb > 1 && warp_smem_bytes(w.NCs) * wpb > (size_t)220 * 1024) wpb >>= 1;
            const int smem = (int)(warp_smem_bytes(w.NCs) * wpb);
            cudaError_t e = cudaFuncSetAttribute(k_ensemble_warp<P, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 148, occ = 1;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ensemble_warp<P, ORDER>, wpb * 32, smem);
            if (occ < 1) occ = 1;
            long long blocks = (long long)sms * occ;
            const long long need = (w.nwork + wpb - 1) / wpb;
            if (blocks > need) blocks = need;
            k_ensemble_warp<P, ORDER><<<(unsigned)blocks, wpb * 32, smem, st>>>(w);
            return cudaGetLastError();
        } else {
            return cudaErrorNotSupported;
        }
    }
    static EnsembleOps make() {
        using LY = EnsLayout<P, ORDER>;
        return EnsembleOps{P::n, P::np, LY::slots_per_node, LY::oMESH, LY::oY, &run, kWarp ? &run_warp : nullptr, &warp_smem_bytes};
    }
};
#define ENS2(P)                                                  \
    { static const EnsembleOps o4 = EnsImpl<P, 4>::make();       \
      static const EnsembleOps o6 = EnsImpl<P, 6>::make();       \
      return order == 4 ? &o4 : order == 6 ? &o6 : nullptr; }

}  // namespace mirk
