"""The BASELINE.json configurations as concrete (problem, parameters, mesh, guess) tuples.

Constants and seeds are the ones frozen in BASELINE.md / SURVEY.md §8(d).  Used by
bench.py, the tests and the oracle-side CPU baseline alike, so both sides see identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


_MESH_PROVIDER = None


def set_mesh_provider(fn) -> None:
    """Use `fn(t0, t1, nint) -> ndarray` for the uniform meshes instead of the C ABI's host helper.  bench.py's
    reference arm passes the oracle's twin (oracle.mesh_uniform, same correctly rounded values) so that the CPU
    arm never maps libmirkb200.so."""
    global _MESH_PROVIDER
    _MESH_PROVIDER = fn


@dataclass
class Config:
    key: str            # C1..C5
    problem: str        # device-function / oracle built-in name
    order: int          # 4 = MIRK4, 6 = MIRK6
    n: int
    p: np.ndarray
    tspan: tuple
    nint: int           # mesh intervals (N - 1)
    y0: np.ndarray      # (N, n) guess at the mesh nodes
    adaptive: bool
    desc: str

    @property
    def N(self) -> int:
        return self.nint + 1

    @property
    def mesh(self) -> np.ndarray:
        # correctly rounded uniform mesh (Julia `range` is twice-precision): the C ABI's
        # host helper mirk_mesh_uniform (binary128; needs no GPU)
        if _MESH_PROVIDER is not None:
            return np.ascontiguousarray(_MESH_PROVIDER(float(self.tspan[0]), float(self.tspan[1]), int(self.nint)))
        from . import _lib as B
        m = np.zeros(self.nint + 1)
        B.check(B.lib().mirk_mesh_uniform(float(self.tspan[0]), float(self.tspan[1]), int(self.nint),
                                          m.ctypes.data_as(B.dp)))
        return m


def _chain(npend: int, seed: int, nint: int, order: int, key: str) -> Config:
    g, kappa, T = 9.81, 4.0, 0.5
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1.0, 1.0, npend)
    b = rng.uniform(-1.0, 1.0, npend)
    p = np.concatenate([[g, kappa], a, b])
    cfg = Config(key, f"chain{npend}", order, 2 * npend, p, (0.0, T), nint, np.zeros((1, 1)), False,
                 f"chain of {npend} torsionally coupled pendula, two-point, MIRK{order}, N={nint + 1} nodes")
    t = cfg.mesh
    theta = a[None, :] + (b - a)[None, :] * (t[:, None] / T)
    omega = np.broadcast_to(((b - a) / T)[None, :], theta.shape)
    cfg.y0 = np.ascontiguousarray(np.concatenate([theta, omega], axis=1))
    return cfg


def c1_pendulum() -> Config:
    """benchmark/simple_pendulum.jl:5-19,32 — MIRK4, dt = 0.05 => 32 intervals"""
    tspan = (0.0, math.pi / 2)
    nint = int(math.ceil((tspan[1] - tspan[0]) / 0.05))
    y0 = np.tile(np.array([math.pi / 2, math.pi / 2]), (nint + 1, 1))
    return Config("C1", "pendulum", 4, 2, np.array([9.81]), tspan, nint, y0, True,
                  "simple pendulum, MIRK4, dt=0.05, adaptive")


def c2_chain8(nint: int = 19999) -> Config:
    """MIRK6, n = 16, 20 000 mesh nodes, fixed mesh (the metric's headline configuration)"""
    return _chain(8, 0, nint, 6, "C2")


def c3_ensemble_params(ntraj: int = 262144) -> np.ndarray:
    """g/L ~ U(8, 12) from default_rng(1), one parameter per trajectory"""
    return np.random.default_rng(1).uniform(8.0, 12.0, (ntraj, 1))


def c4_bratu64(nint: int = 3999) -> Config:
    n = 128
    return Config("C4", "bratu64", 4, n, np.array([1.0]), (0.0, 1.0), nint, np.zeros((nint + 1, n)), False,
                  "2-D Bratu by the method of lines (64 lines), two-point, MIRK4, zero guess")


def c5_chain16(nint: int = 1999999) -> Config:
    return _chain(16, 2, nint, 6, "C5")
