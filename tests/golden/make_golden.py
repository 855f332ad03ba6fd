#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the CPU oracle (run from the repo root).

The reference (Julia) cannot run in this image and ships no golden vectors of its own,
so these pin the CUDA path against the oracle at sizes too large to recompute inside every test run.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import mirk_b200  # noqa: E402,F401
from boundaryvaluediffeq_jl_b200 import configs  # noqa: E402

out = {}
# C2 at full size: |F|_inf after each of the Newton steps from the linear guess, and a solution checksum
c = configs.c2_chain8()
ws = O.Workspace(O.builtin(c.problem), c.order, c.p, c.mesh, c.y0)
norms = [float(np.max(np.abs(ws.loss())))]
for k in range(3):
    ws2 = O.Workspace(O.builtin(c.problem), c.order, c.p, c.mesh, c.y0)
    ret, it, nrm = ws2.newton(abstol=0.0, maxiters=k + 1)
    norms.append(float(nrm))
out["c2_newton_norms"] = norms
out["c2_solution_checksum"] = {"sum": float(ws2.y.sum()), "sum_abs": float(np.abs(ws2.y).sum()),
                               "y_mid": ws2.y[c.N // 2].tolist()}
# C5's problem (n = 32) and C4's (n = 128) on short meshes
for key, cfg in (("c5_short", configs.c5_chain16(999)), ("c4_short", configs.c4_bratu64(99))):
    w = O.Workspace(O.builtin(cfg.problem), cfg.order, cfg.p, cfg.mesh, cfg.y0)
    ret, it, nrm = w.newton()
    out[key] = {"retcode": ret, "iters": it, "resid_norm": float(nrm), "sum": float(w.y.sum()),
                "sum_abs": float(np.abs(w.y).sum()), "y_mid": w.y[cfg.N // 2].tolist()}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "newton_golden.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "y_mid"}) for k, v in out.items()}, indent=1))
