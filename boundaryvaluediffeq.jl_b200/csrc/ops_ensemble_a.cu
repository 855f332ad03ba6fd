// ensemble (one thread per trajectory) instantiations, part a: Pendulum, Linear2
#define MIRK_OUTLINE_ELEMENTARY 1  // whole-solve kernels are bounded by code size (see dual.cuh)
#include "ensemble_ops.cuh"
namespace mirk {
const EnsembleOps* ensemble_ops_part_a(int id, int order) {
    switch (id) {
    case kPendulum: ENS2(Pendulum)
    case kLinear2: ENS2(Linear2)
    default: return nullptr;
    }
}
}  // namespace mirk
