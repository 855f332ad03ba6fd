// built-in problem "chain8" (Chain<8>), MIRK4 and MIRK6
#include "ops.cuh"
namespace mirk {
const ProblemOps* ops_chain8(int order) {
    static const ProblemOps o4 = OpsImpl<problems::Chain<8>, 4>::make("chain8");
    static const ProblemOps o6 = OpsImpl<problems::Chain<8>, 6>::make("chain8");
    return order == 4 ? &o4 : order == 6 ? &o6 : nullptr;
}
}  // namespace mirk
