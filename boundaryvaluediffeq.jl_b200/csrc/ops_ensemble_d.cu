// ensemble (one thread per trajectory) instantiations, part d: Swirling
#define MIRK_OUTLINE_ELEMENTARY 1  // whole-solve kernels are bounded by code size (see dual.cuh)
#include "ensemble_ops.cuh"
namespace mirk {
const EnsembleOps* ensemble_ops_part_d(int id, int order) {
    switch (id) {
    case kSwirling: ENS2(Swirling)
    default: return nullptr;
    }
}
}  // namespace mirk
