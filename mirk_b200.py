"""Import alias for the package directory `boundaryvaluediffeq.jl_b200/` (a dot is not importable):

    import mirk_b200 as bvp
    sol = bvp.solve(bvp.BVProblem("pendulum", [1.57, 1.57], (0, 1.57), p=[9.81]), bvp.MIRK4(), dt=0.05)
"""
import importlib.util
import os
import sys

_NAME = "boundaryvaluediffeq_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "boundaryvaluediffeq.jl_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
package = _pkg
