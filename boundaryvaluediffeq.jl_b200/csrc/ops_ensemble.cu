// dispatcher over the ensemble instantiation parts (split so they compile in parallel)
#include "ops.cuh"
namespace mirk {
const EnsembleOps* ensemble_ops_part_a(int id, int order);
const EnsembleOps* ensemble_ops_part_b(int id, int order);
const EnsembleOps* ensemble_ops_part_c(int id, int order);
const EnsembleOps* ensemble_ops_part_d(int id, int order);
const EnsembleOps* ensemble_ops_small(int id, int order) {
    if (const EnsembleOps* o = ensemble_ops_part_a(id, order)) return o;
    if (const EnsembleOps* o = ensemble_ops_part_b(id, order)) return o;
    if (const EnsembleOps* o = ensemble_ops_part_c(id, order)) return o;
    return ensemble_ops_part_d(id, order);
}
}  // namespace mirk
