"""Reference path generation — host mirror of PKG/RefPathGenerator.py:9-59.

Same class / method names and return shapes.  The reference trajectory has zero weight in
the cost (`aa = 0.0`, PKG/MPC_CBF_optimize_kin.py:194-197) but the mains still build it and
index it, so the surface is kept.
"""
from __future__ import annotations

import numpy as np


class RefPathGenerator:
    def __init__(self):
        self.ref_global = None
        self.step_x = None
        self.ref_len = None

    def define_ref_path(self, x0, xs, dt):
        """Straight global path from x0.x to xs.x at 1 m spacing, y/phi/v of the target (:9-24)."""
        x_start = float(np.asarray(x0).reshape(-1)[0])
        tgt = np.asarray(xs, dtype=float).reshape(-1)
        self.step_x = 1
        if tgt[0] > x_start:
            gx = np.arange(x_start, tgt[0] + self.step_x, self.step_x)
        else:
            gx = np.arange(x_start, tgt[0] - self.step_x, -self.step_x)
        self.ref_global = np.stack([gx, np.full_like(gx, tgt[1]), np.full_like(gx, tgt[2]), np.full_like(gx, tgt[3])], axis=1)
        self.ref_len = len(self.ref_global)
        return self.ref_global

    def find_ref_traj(self, x0, xs, T_horizon, dt, last_idx):
        """Nearest point search from `last_idx` and preview resampling to N_p+1 rows (:27-59)."""
        ego = np.asarray(x0, dtype=float).reshape(-1)
        tgt = np.asarray(xs, dtype=float).reshape(-1)
        N_p = int(T_horizon / dt)
        preview_v = 0.5 * ego[3] + 0.5 * tgt[3]
        preview_idx = int(preview_v * T_horizon / self.step_x)
        lo = max(0, last_idx - 5)
        hi = min(self.ref_len, last_idx + preview_idx)
        window = self.ref_global[lo:hi, :]
        dist = np.sqrt((window[:, 0] - ego[0]) ** 2 + (window[:, 1] - ego[1]) ** 2)
        # the reference walks forward and stops at the first non-decreasing distance
        min_idx, best = lo, np.inf
        for i, d in enumerate(dist):
            if d < best:
                best, min_idx = d, lo + i
            else:
                break
        idx = np.linspace(min_idx, min_idx + preview_idx, N_p + 1)
        idx = np.clip(idx, 0, self.ref_len - 1).astype(int)
        return self.ref_global[idx, :], min_idx
