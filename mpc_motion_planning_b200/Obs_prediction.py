"""Obstacle trajectory prediction — host mirror of PKG/Obs_prediction.py:3-40.

Same call surface (`obs_prediction(obs_list, dt, N_p)` -> list of (N_p+1, 6) arrays,
rows [x, y, theta, v, l, w]) plus a batched variant used by the scenario generators and
the batched solver.  The constant-velocity / constant-heading update is accumulated step
by step in the same operation order as the reference so that results are bit-identical.
"""
from __future__ import annotations

import numpy as np


def obs_prediction_batch(obs0: np.ndarray, dt: float, N_p: int) -> np.ndarray:
    """obs0 (..., 6) -> (..., N_p+1, 6)."""
    obs0 = np.asarray(obs0, dtype=np.float64)
    out = np.empty(obs0.shape[:-1] + (N_p + 1, 6), dtype=np.float64)
    x = obs0[..., 0].copy()
    y = obs0[..., 1].copy()
    th, v = obs0[..., 2], obs0[..., 3]
    # (v * cos(theta)) * dt, as the reference evaluates it at PKG/Obs_prediction.py:27-28
    dx = v * np.cos(th) * dt
    dy = v * np.sin(th) * dt
    for k in range(N_p + 1):
        out[..., k, 0] = x
        out[..., k, 1] = y
        x = x + dx
        y = y + dy
    out[..., 2] = th[..., None]
    out[..., 3] = v[..., None]
    out[..., 4] = obs0[..., 4][..., None]
    out[..., 5] = obs0[..., 5][..., None]
    return out


def obs_prediction(obs_list, dt, N_p):
    """List of (1,6) obstacle states -> list of (N_p+1,6) predicted trajectories."""
    return [obs_prediction_batch(np.asarray(o, dtype=np.float64).reshape(6), dt, N_p) for o in obs_list]
