// ensemble (one thread per trajectory) instantiations, part c: Torus
#define MIRK_OUTLINE_ELEMENTARY 1  // whole-solve kernels are bounded by code size (see dual.cuh)
#include "ensemble_ops.cuh"
namespace mirk {
const EnsembleOps* ensemble_ops_part_c(int id, int order) {
    switch (id) {
    case kTorus: ENS2(Torus)
    default: return nullptr;
    }
}
}  // namespace mirk
