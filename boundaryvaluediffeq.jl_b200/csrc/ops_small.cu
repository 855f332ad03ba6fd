// built-in problems with n <= 6 (ids 0..6, 10, 11), MIRK4 and MIRK6
#include "ops.cuh"
namespace mirk {
using namespace problems;
#define OPS2(P, NAME)                                                        \
    { static const ProblemOps o4 = OpsImpl<P, 4>::make(NAME);               \
      static const ProblemOps o6 = OpsImpl<P, 6>::make(NAME);               \
      return order == 4 ? &o4 : order == 6 ? &o6 : nullptr; }
const ProblemOps* ops_small(int id, int order) {
    switch (id) {
    case kPendulum: OPS2(Pendulum, "pendulum")
    case kLinear2: OPS2(Linear2, "linear2")
    case kLinear2TP: OPS2(Linear2TP, "linear2_tp")
    case kSwirling: OPS2(Swirling, "swirling")
    case kLotka: OPS2(Lotka, "lotka")
    case kTorus: OPS2(Torus, "torus")
    case kLayer: OPS2(Layer, "layer")
    case kLaneEmden: OPS2(LaneEmden, "lane_emden")
    case kRobinSine: OPS2(RobinSine, "robin_sine")
    default: return nullptr;
    }
}
}  // namespace mirk
