// built-in problem "bratu64" (BratuMOL<64>), MIRK4 and MIRK6
#include "ops.cuh"
namespace mirk {
const ProblemOps* ops_bratu64(int order) {
    static const ProblemOps o4 = OpsImpl<problems::BratuMOL<64>, 4>::make("bratu64");
    static const ProblemOps o6 = OpsImpl<problems::BratuMOL<64>, 6>::make("bratu64");
    return order == 4 ? &o4 : order == 6 ? &o6 : nullptr;
}
}  // namespace mirk
