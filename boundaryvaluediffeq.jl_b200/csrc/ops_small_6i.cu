// built-in problems with n <= 6 (ids 0..6), MIRK6I (`order` code 7: the irrational 6th-order tableau)
#include "ops.cuh"
namespace mirk {
using namespace problems;
#define OPS6I(P, NAME)                                                           \
    { static const ProblemOps o = OpsImpl<P, kMIRK6I>::make(NAME);              \
      return order == kMIRK6I ? &o : nullptr; }
const ProblemOps* ops_small_6i(int id, int order) {
    switch (id) {
    case kPendulum: OPS6I(Pendulum, "pendulum")
    case kLinear2: OPS6I(Linear2, "linear2")
    case kLinear2TP: OPS6I(Linear2TP, "linear2_tp")
    case kSwirling: OPS6I(Swirling, "swirling")
    case kLotka: OPS6I(Lotka, "lotka")
    case kTorus: OPS6I(Torus, "torus")
    case kLayer: OPS6I(Layer, "layer")
    case kLaneEmden: OPS6I(LaneEmden, "lane_emden")
    default: return nullptr;
    }
}
}  // namespace mirk
