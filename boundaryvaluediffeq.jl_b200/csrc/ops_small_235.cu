// built-in problems with n <= 6 (ids 0..6), MIRK2, MIRK3 and MIRK5
#include "ops.cuh"
namespace mirk {
using namespace problems;
#define OPS2(P, NAME)                                                        \
    { static const ProblemOps o2 = OpsImpl<P, 2>::make(NAME);               \
      static const ProblemOps o3 = OpsImpl<P, 3>::make(NAME);               \
      static const ProblemOps o5 = OpsImpl<P, 5>::make(NAME);               \
      return order == 2 ? &o2 : order == 3 ? &o3 : order == 5 ? &o5 : nullptr; }
const ProblemOps* ops_small_235(int id, int order) {
    switch (id) {
    case kPendulum: OPS2(Pendulum, "pendulum")
    case kLinear2: OPS2(Linear2, "linear2")
    case kLinear2TP: OPS2(Linear2TP, "linear2_tp")
    case kSwirling: OPS2(Swirling, "swirling")
    case kLotka: OPS2(Lotka, "lotka")
    case kTorus: OPS2(Torus, "torus")
    case kLayer: OPS2(Layer, "layer")
    case kLaneEmden: OPS2(LaneEmden, "lane_emden")
    default: return nullptr;
    }
}
}  // namespace mirk
