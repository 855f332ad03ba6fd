// built-in problem "chain16" (Chain<16>), MIRK4 and MIRK6
#include "ops.cuh"
namespace mirk {
const ProblemOps* ops_chain16(int order) {
    static const ProblemOps o4 = OpsImpl<problems::Chain<16>, 4>::make("chain16");
    static const ProblemOps o6 = OpsImpl<problems::Chain<16>, 6>::make("chain16");
    return order == 4 ? &o4 : order == 6 ? &o6 : nullptr;
}
}  // namespace mirk
