"""B200-native MIRK4/MIRK6 collocation Newton path behind the SciML `solve(BVProblem, MIRK4(); dt)` /
`EnsembleProblem` interface.  All numerics run in libmirkb200.so (hand-written sm_100a CUDA behind the C ABI
of include/mirk_b200.h); this package is the thin host mirror of the reference interface.

The directory name contains a dot, so import it through the repo-root alias: `import mirk_b200`.
"""
from ._lib import LIB_PATH, MirkError, build, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import mesh_uniform  # noqa: F401
