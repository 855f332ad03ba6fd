// ensemble (one thread per trajectory) instantiations, part b: Linear2TP, Lotka, Layer
#define MIRK_OUTLINE_ELEMENTARY 1  // whole-solve kernels are bounded by code size (see dual.cuh)
#include "ensemble_ops.cuh"
namespace mirk {
const EnsembleOps* ensemble_ops_part_b(int id, int order) {
    switch (id) {
    case kLinear2TP: ENS2(Linear2TP)
    case kLotka: ENS2(Lotka)
    case kLayer: ENS2(Layer)
    case kLaneEmden: ENS2(LaneEmden)
    default: return nullptr;
    }
}
}  // namespace mirk
