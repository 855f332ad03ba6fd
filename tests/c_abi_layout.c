/* Prints sizeof / offsetof of every struct include/mirk_b200.h passes across the C ABI, one "name value" per line.
 * tests/test_host_logic.py compiles this with gcc against the header and compares the numbers with the ctypes
 * mirror (boundaryvaluediffeq.jl_b200/_lib.py) and with the Julia struct definitions (parsed from the .jl glue),
 * so a field added on one side only is caught without a Julia runtime. */
#include <stddef.h>
#include <stdio.h>

#include "mirk_b200.h"

#define SZ(T) printf("sizeof." #T " %zu\n", sizeof(T))
#define OFF(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))

int main(void) {
    SZ(mirk_desc);
    OFF(mirk_desc, problem_id); OFF(mirk_desc, order); OFF(mirk_desc, abstol); OFF(mirk_desc, adaptive);
    OFF(mirk_desc, defect_threshold); OFF(mirk_desc, max_num_subintervals); OFF(mirk_desc, maxiters);
    OFF(mirk_desc, reinterp_inplace); OFF(mirk_desc, chunk); OFF(mirk_desc, device); OFF(mirk_desc, n_params);
    OFF(mirk_desc, params); OFF(mirk_desc, nlsolve);
    OFF(mirk_desc, controller); OFF(mirk_desc, ge_method); OFF(mirk_desc, DE); OFF(mirk_desc, GE);
    SZ(mirk_problem_info);
    OFF(mirk_problem_info, n); OFF(mirk_problem_info, n_params); OFF(mirk_problem_info, problem_type);
    OFF(mirk_problem_info, n_bc); OFF(mirk_problem_info, n_bca); OFF(mirk_problem_info, max_bc_pts);
    SZ(mirk_result);
    OFF(mirk_result, retcode); OFF(mirk_result, n_mesh); OFF(mirk_result, outer_iters); OFF(mirk_result, newton_iters);
    OFF(mirk_result, resid_norm); OFF(mirk_result, defect_norm); OFF(mirk_result, n_hist); OFF(mirk_result, hist_n_mesh);
    OFF(mirk_result, hist_newton); OFF(mirk_result, hist_defect);
    SZ(mirk_ensemble_desc);
    OFF(mirk_ensemble_desc, problem_id); OFF(mirk_ensemble_desc, order); OFF(mirk_ensemble_desc, abstol);
    OFF(mirk_ensemble_desc, adaptive); OFF(mirk_ensemble_desc, defect_threshold);
    OFF(mirk_ensemble_desc, max_num_subintervals); OFF(mirk_ensemble_desc, maxiters);
    OFF(mirk_ensemble_desc, reinterp_inplace); OFF(mirk_ensemble_desc, device); OFF(mirk_ensemble_desc, node_cap);
    OFF(mirk_ensemble_desc, t0); OFF(mirk_ensemble_desc, t1); OFF(mirk_ensemble_desc, dt); OFF(mirk_ensemble_desc, nlsolve);
    return 0;
}
