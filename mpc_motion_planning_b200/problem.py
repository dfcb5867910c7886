"""Problem definitions: the constants the reference hard-codes in `optimize_problem` /
`initialize_constraints`, collected per NLP kind and turned into an `mpcb_cfg`.

kinds
  kin_nocbf    MPC_optimize_kin (source missing, recovered from the .pyc)
  kin_cbf      PKG/MPC_CBF_optimize_kin.py:136-255      static obstacle rows
  kin_cbf_pre  PKG/MPC_CBF_optimize_kin_pre.py:136-261  per-step obstacle trajectories
  dyn          PKG/MPC_CBF_optimize_dyn.py:137-250      dynamic bicycle (bounds as intended, "aligned")
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib

INF = float("inf")
KINDS = ("kin_nocbf", "kin_cbf", "kin_cbf_pre", "dyn")


@dataclass
class Weights:
    Q: tuple
    R: tuple
    DR: tuple
    du0_cost: bool


REFERENCE_WEIGHTS = {
    # pyc line table L140-156, L167-170
    "kin_nocbf": Weights((10.0, 100.0, 10.0, 100.0), (100.0, 100.0), (1e4, 1e4), False),
    # PKG/MPC_CBF_optimize_kin.py:168-184 ; i=0 rate cost vs Ulast=0 at :203-204
    "kin_cbf": Weights((1e1, 1e5, 3e5, 1e4), (1e4, 1e4), (1e5, 1e2), True),
    "kin_cbf_pre": Weights((1e1, 1e5, 3e5, 1e4), (1e4, 1e4), (1e5, 1e2), True),
    # PKG/MPC_CBF_optimize_dyn.py:189-209 ; obj_dU = 0 at i=0 (:221-224)
    "dyn": Weights((10.0, 1e5, 1e3, 1e3, 1.0, 1.0), (1e3, 1e3), (5e3, 5e2), False),
}


def horizon_steps(config: dict) -> int:
    """N_p of the uniform grid branch, PKG/MPC_CBF_optimize_kin.py:32-33."""
    mp = config["mpc_params"]
    return len(np.arange(0, mp["horizon"] + mp["T_S"], mp["T_S"], dtype=float)) - 1


def make_cfg(kind: str, config: dict, N: int | None = None, M: int = 1, weights: Weights | None = None,
             init_mode: int = _lib.INIT_ROLLOUT, mu_init: float = 30.0, max_iter: int = 100, tol: float = 1e-8,
             bounds: dict | None = None, obs_input: int = _lib.OBS_TRAJECTORY, cbf_gamma: float | None = None,
             ref_mode: int = _lib.REF_TERMINAL, dyn_rows: int = _lib.DYN_ROWS_ALIGNED, restoration: bool = False,
             resto_max_calls: int = 1) -> _lib.MpcbCfg:
    """Fill an mpcb_cfg from the YAML dict with the reference's hard-coded constants.

    `bounds` optionally overrides {'u_lo','u_hi','x_lo','x_hi','rate_lo','rate_hi'} (used by the
    reference-surface shim, which receives lbx/ubx/lbg/ubg lists at call time).
    `cbf_gamma` switches the obstacle rows of the kin-CBF kinds to the discrete-time CBF form the
    reference carries commented out (PKG/MPC_CBF_optimize_kin.py:244-248); `ref_mode` =
    REF_TRAJECTORY makes `xs` a per-stage target array (the `aa` blend of :194-199)."""
    if kind not in KINDS:
        raise ValueError(f"unknown kind {kind!r}")
    mp, vp = config["mpc_params"], config["vehicle_params"]
    tp, kc, dc = config["tire_params"], config["kinematics_constraints"], config["dynamics_constraints"]
    w = weights or REFERENCE_WEIGHTS[kind]
    c = _lib.MpcbCfg()
    nx = 6 if kind == "dyn" else 4
    c.model = _lib.MODEL_DYN if kind == "dyn" else _lib.MODEL_KIN
    c.N = N if N is not None else horizon_steps(config)
    c.M = 0 if kind == "kin_nocbf" else M
    c.obs_mode = {"kin_nocbf": _lib.OBS_NONE, "kin_cbf": _lib.OBS_ELLIPSE, "kin_cbf_pre": _lib.OBS_ELLIPSE, "dyn": _lib.OBS_SQRT}[kind]
    c.du0_cost = int(w.du0_cost)
    c.T = float(mp["T_S"])
    for i in range(nx):
        c.Q[i] = w.Q[i]
        c.x_lo[i], c.x_hi[i] = -INF, INF
    for i in range(2):
        c.R[i], c.DR[i] = w.R[i], w.DR[i]
    rad = math.pi / 180  # degrees -> radians at load, PKG/MPC_CBF_optimize_kin.py:62-63,70-71
    ddf = (dc["df_dot_min"] * np.pi / 180 * c.T, dc["df_dot_max"] * np.pi / 180 * c.T)
    if kind in ("kin_cbf", "kin_cbf_pre"):  # steering-rate rows only, :211-216, bounds :119-121
        c.n_rate = 1
        c.rate_ctrl[0] = 0
        c.rate_lo[0], c.rate_hi[0] = ddf
    elif kind == "dyn":  # both rate rows, PKG/MPC_CBF_optimize_dyn.py:229-231, bounds :126-129
        c.n_rate = 2
        c.rate_ctrl[0], c.rate_ctrl[1] = 0, 1
        c.rate_lo[0], c.rate_hi[0] = ddf
        c.rate_lo[1], c.rate_hi[1] = dc["jerk_min"] * c.T, dc["jerk_max"] * c.T
    c.u_lo[0], c.u_hi[0] = kc["df_min"] * np.pi / 180, kc["df_max"] * np.pi / 180
    c.u_lo[1], c.u_hi[1] = kc["ax_min"], kc["ax_max"]
    c.x_lo[1], c.x_hi[1] = kc["Y_min"], kc["Y_max"]
    c.x_lo[3], c.x_hi[3] = kc["vx_min"], kc["vx_max"]
    if nx == 6:
        c.x_lo[4], c.x_hi[4] = dc["vy_min"], dc["vy_max"]
    if bounds:
        for key, arr in bounds.items():
            tgt = getattr(c, key)
            for i, v in enumerate(arr):
                tgt[i] = float(v)
    c.obs_lo = 1.0 if kind == "dyn" else 0.0
    veh_w = vp["Veh_W"] if "Veh_W" in vp else vp["Veh_w"]  # the dyn module reads the lower-case key
    c.ego_hl, c.ego_hw = vp["Veh_L"] / 2, veh_w / 2
    c.safe_l, c.safe_w = 1.0, 0.5  # safe_disl, safe_disw, PKG/MPC_CBF_optimize_kin.py:224-225
    c.dyn_sx, c.dyn_sy = 4.0, 1.0  # PKG/MPC_CBF_optimize_dyn.py:240-241
    c.Veh_l, c.Veh_lf, c.Veh_lr = vp["Veh_l"], vp["Veh_lf"], vp["Veh_lr"]
    c.Veh_m, c.Veh_Iz = vp["Veh_m"], vp["Veh_Iz"]
    c.aopt_f, c.aopt_r = tp["aopt_f"], tp["aopt_r"]
    c.Fymax_f = tp["Cf_0"] * tp["aopt_f"] / 2  # PKG/MPC_CBF_optimize_kin.py:55-56
    c.Fymax_r = tp["Cr_0"] * tp["aopt_r"] / 2
    c.init_mode = init_mode
    c.max_iter = max_iter
    c.tol, c.mu_init, c.bound_relax = tol, mu_init, 1e-8
    c.obs_input = obs_input
    c.ref_mode = ref_mode
    c.dyn_rows = dyn_rows  # dyn: bound lists aligned with g, or paired exactly as shipped
    # restoration phase after a failed line search (kinematic families with rows); resto_max_calls = 0 is IPOPT's
    # behaviour (no cap on how often the phase is entered)
    c.restoration = int(bool(restoration) and kind in ("kin_cbf", "kin_cbf_pre"))
    c.resto_max_calls = int(resto_max_calls)
    if dyn_rows != _lib.DYN_ROWS_ALIGNED and kind != "dyn":
        raise ValueError("dyn_rows applies to the dyn kind only")
    if cbf_gamma is not None:
        if kind not in ("kin_cbf", "kin_cbf_pre"):
            raise ValueError("cbf_gamma applies to the kinematic CBF kinds only")
        c.obs_mode, c.cbf_gamma = _lib.OBS_DCBF, float(cbf_gamma)
    if ref_mode == _lib.REF_TRAJECTORY and kind == "dyn":
        raise ValueError("per-stage references are implemented for the kinematic kinds only")
    return c
