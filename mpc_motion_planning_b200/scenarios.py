"""Seeded synthetic scenario batches for the benchmark configs.

The reference has no scenario generator; its mains hard-code one scenario each
(PKG/main_cbf_kin_c_sim.py:45-55, PKG/main_cbf_kin_c_sim_pre.py:45-56,
PKG/main_cbf_dyn_c_sim.py:44-51, PKG/main_kin_c_sim.py:42-46).  These generators draw
batches around those scenarios; the distributions are the ones SURVEY.md fixes so that
the GPU path, the CPU oracle and the benchmark all see identical inputs.
"""
from __future__ import annotations

import numpy as np

from .Obs_prediction import obs_prediction_batch

BASE_SEED = 20261018


def kin_cbf_static(B: int, N: int = 50, seed: int = BASE_SEED + 2, dt: float = 0.1):
    """config 2: kinematic CBF MPC, one static obstacle per scenario.

    returns x0 (B,4), xs (B,4), obs (B,1,N+1,6) float64 (obstacle rows [x,y,theta,v,l,w]).
    """
    rng = np.random.default_rng(seed)
    x0 = np.stack(
        [rng.uniform(0, 20, B), rng.uniform(0, 4.5, B), rng.uniform(-0.05, 0.05, B), rng.uniform(10, 25, B)], axis=1
    )
    xs = np.tile(np.array([400.0, 3.5, 0.0, 30.0]), (B, 1))
    ob = np.stack(
        [x0[:, 0] + rng.uniform(30, 80, B), rng.uniform(0, 4, B), np.zeros(B), np.zeros(B), np.full(B, 4.8), np.full(B, 1.8)],
        axis=1,
    )
    obs = np.repeat(ob[:, None, None, :], N + 1, axis=2)
    return x0, xs, np.ascontiguousarray(obs)


def kin_cbf_moving(B: int, N: int = 50, seed: int = BASE_SEED + 3, dt: float = 0.1):
    """config 3 / 5: kinematic CBF MPC with a predicted moving obstacle."""
    rng = np.random.default_rng(seed)
    x0 = np.stack(
        [rng.uniform(0, 20, B), rng.uniform(0, 4.5, B), rng.uniform(-0.05, 0.05, B), rng.uniform(10, 25, B)], axis=1
    )
    xs = np.tile(np.array([400.0, 3.5, 0.0, 30.0]), (B, 1))
    ob = np.stack(
        [
            x0[:, 0] + rng.uniform(30, 80, B),
            rng.uniform(0, 4, B),
            rng.uniform(-0.05, 0.05, B),
            rng.uniform(5, 12, B),
            np.full(B, 4.8),
            np.full(B, 1.8),
        ],
        axis=1,
    )
    obs = obs_prediction_batch(ob[:, None, :], dt, N)  # (B,1,N+1,6)
    return x0, xs, obs


def kin_nocbf(B: int, N: int = 50, seed: int = BASE_SEED + 1):
    """config 1 family: no-CBF kinematic tracking around PKG/main_kin_c_sim.py:42-46."""
    rng = np.random.default_rng(seed)
    x0 = np.stack(
        [rng.uniform(0, 20, B), rng.uniform(0, 4.5, B), rng.uniform(-0.05, 0.05, B), rng.uniform(10, 25, B)], axis=1
    )
    xs = np.tile(np.array([500.0, 3.5, 0.0, 30.0]), (B, 1))
    return x0, xs, np.zeros((B, 0, N + 1, 6))


def dyn_static(B: int, N: int = 50, seed: int = BASE_SEED + 4):
    """config 4: dynamic bicycle, one static obstacle centre (x,y)."""
    rng = np.random.default_rng(seed)
    z = np.zeros(B)
    x0 = np.stack([rng.uniform(0, 20, B), rng.uniform(-0.5, 4.5, B), rng.uniform(-0.05, 0.05, B), rng.uniform(8, 20, B), z, z], axis=1)
    xs = np.tile(np.array([600.0, 3.5, 0.0, 15.0, 0.0, 0.0]), (B, 1))
    oc = np.stack([x0[:, 0] + rng.uniform(60, 120, B), rng.uniform(-4, 4, B)], axis=1)
    obs = np.zeros((B, 1, N + 1, 6))
    obs[:, 0, :, 0:2] = oc[:, None, :]
    return x0, xs, obs
