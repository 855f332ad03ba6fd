// ensemble (one thread per trajectory) instantiations, part b: Linear2TP, Lotka, Layer
#include "ensemble.cuh"
#include "ops.cuh"
namespace mirk {
using namespace problems;
template <class P, int ORDER> struct EnsImpl {
    static void run(cudaStream_t st, const EnsArgs& a) {
        // n <= 2: cap registers at 96 (10 CTAs of 64 threads per SM): the kernel is memory-latency bound and
        // the extra resident warps buy 17 % (experiments/exp_ens.cu); larger n needs the registers
        constexpr int MINB = P::n <= 2 ? 10 : 1;
        // small shards (e.g. 262 144 trajectories strong-scaled over 8 GPUs): one warp per CTA so the CTAs
        // spread evenly over the SMs instead of leaving some SMs with a whole CTA more than others
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int tpb = (a.ntraj + 63) / 64 < 8LL * sms ? 32 : 64;
        const long long blocks = (a.ntraj + tpb - 1) / tpb;
        k_ensemble_solve<P, ORDER, MINB><<<(unsigned)blocks, tpb, 0, st>>>(a);
    }
    static EnsembleOps make() {
        using LY = EnsLayout<P, ORDER>;
        return EnsembleOps{P::n, P::np, LY::slots_per_node, LY::oMESH, LY::oY, &run};
    }
};
#define ENS2(P)                                                  \
    { static const EnsembleOps o4 = EnsImpl<P, 4>::make();       \
      static const EnsembleOps o6 = EnsImpl<P, 6>::make();       \
      return order == 4 ? &o4 : order == 6 ? &o6 : nullptr; }
const EnsembleOps* ensemble_ops_part_b(int id, int order) {
    switch (id) {
    case kLinear2TP: ENS2(Linear2TP)
    case kLotka: ENS2(Lotka)
    case kLayer: ENS2(Layer)
    default: return nullptr;
    }
}
}  // namespace mirk
