// ensemble (one thread per trajectory) instantiations, part a: Pendulum, Linear2
#include "ensemble.cuh"
#include "ops.cuh"
namespace mirk {
using namespace problems;
template <class P, int ORDER> struct EnsImpl {
    static void run(cudaStream_t st, const EnsArgs& a) {
        const long long blocks = (a.ntraj + 63) / 64;
        // n <= 2: cap registers at 96 (10 CTAs of 64 threads per SM): the kernel is memory-latency bound and
        // the extra resident warps buy 17 % (experiments/exp_ens.cu); larger n needs the registers
        constexpr int MINB = P::n <= 2 ? 10 : 1;
        k_ensemble_solve<P, ORDER, MINB><<<(unsigned)blocks, 64, 0, st>>>(a);
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
const EnsembleOps* ensemble_ops_part_a(int id, int order) {
    switch (id) {
    case kPendulum: ENS2(Pendulum)
    case kLinear2: ENS2(Linear2)
    default: return nullptr;
    }
}
}  // namespace mirk
