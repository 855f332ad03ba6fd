import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import mirk_b200 as M
from boundaryvaluediffeq_jl_b200 import configs
for maker, nint in (("c4_bratu64", 3999), ("c5_chain16", 99999)):
    c = getattr(configs, maker)(nint)
    alg = M.MIRK6() if c.order == 6 else M.MIRK4()
    cache = M.init(M.BVProblem(c.problem, c.y0, c.tspan, p=c.p, mesh=c.mesh), alg, adaptive=False)
    t = time.time(); ret, it, nrm = cache.newton_solve(); dt = time.time() - t
    st, ms, ph, launches = cache.bench_newton_steps(3)
    print(json.dumps({"config": c.key, "n": c.n, "N": c.N, "newton": [ret, it, nrm], "solve_s": dt, "ms_per_step": ms / 3,
                      "phases_us": [round(1e3 * p / 3, 1) for p in ph[:7]]}), flush=True)
    cache.close()
