#!/usr/bin/env python
"""Golden vectors at the FULL sizes of BASELINE configs C4 (n = 128, N = 4000, MIRK4) and of one GPU's slice of
C5 (n = 32, 250 000 nodes = 2 000 000 / 8, MIRK6), from the CPU oracle (run from the repo root; minutes of CPU).

Writes tests/golden/newton_golden_large.json: the |F|_inf sequence over the Newton steps from the stated guess
and solution checksums after the last step.  Like make_golden.py these pin the CUDA path against the oracle
(the Julia reference cannot run in this image).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import mirk_b200  # noqa: E402,F401
from boundaryvaluediffeq_jl_b200 import configs  # noqa: E402

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "newton_golden_large.json")
out = json.load(open(path)) if os.path.exists(path) else {}
which = sys.argv[1:] or ["c4_full", "c5_slice"]
cases = {"c4_full": (configs.c4_bratu64(3999), 2), "c5_slice": (configs.c5_chain16(249999), 3)}
for key in which:
    cfg, steps = cases[key]
    t0 = time.time()
    ws = O.Workspace(O.builtin(cfg.problem), cfg.order, cfg.p, cfg.mesh, cfg.y0)
    norms = [float(np.max(np.abs(ws.loss())))]
    ret, it, nrm = ws.newton(abstol=0.0, maxiters=steps)
    norms.append(float(nrm))
    N = cfg.N
    out[key] = {"nint": cfg.nint, "steps": int(it), "norm_first": norms[0], "norm_last": norms[1],
                "sum": float(ws.y.sum()), "sum_abs": float(np.abs(ws.y).sum()),
                "y_mid": ws.y[N // 2].tolist(), "y_q1": ws.y[N // 4].tolist(), "y_last": ws.y[N - 2].tolist(),
                "oracle_seconds": time.time() - t0}
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(key, {k: v for k, v in out[key].items() if not isinstance(v, list)}, flush=True)
