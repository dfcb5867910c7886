"""ctypes binding of the C oracle (oracle/mpc_oracle.c) — TEST INFRASTRUCTURE.

Builds liboracle.so with gcc on first use (see oracle/Makefile).  PARITY UNPINNED, see
oracle/nlp.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .nlp import Params, reference_weights

HERE = os.path.dirname(os.path.abspath(__file__))
INF = float("inf")


class OrcCfg(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("N", C.c_int32), ("M", C.c_int32), ("obs_mode", C.c_int32),
        ("du0_cost", C.c_int32), ("n_rate", C.c_int32), ("rate_ctrl", C.c_int32 * 2),
        ("init_mode", C.c_int32), ("max_iter", C.c_int32),
        ("T", C.c_double), ("Q", C.c_double * 6), ("R", C.c_double * 2), ("DR", C.c_double * 2),
        ("rate_lo", C.c_double * 2), ("rate_hi", C.c_double * 2),
        ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2),
        ("x_lo", C.c_double * 6), ("x_hi", C.c_double * 6),
        ("obs_lo", C.c_double),
        ("ego_hl", C.c_double), ("ego_hw", C.c_double), ("safe_l", C.c_double), ("safe_w", C.c_double),
        ("dyn_sx", C.c_double), ("dyn_sy", C.c_double),
        ("Veh_l", C.c_double), ("Veh_lf", C.c_double), ("Veh_lr", C.c_double), ("Veh_m", C.c_double),
        ("Veh_Iz", C.c_double), ("aopt_f", C.c_double), ("aopt_r", C.c_double),
        ("Fymax_f", C.c_double), ("Fymax_r", C.c_double),
        ("tol", C.c_double), ("mu_init", C.c_double), ("bound_relax", C.c_double),
        ("cbf_gamma", C.c_double), ("ref_mode", C.c_int32), ("rows_as_shipped", C.c_int32),
        ("restoration", C.c_int32), ("resto_max_calls", C.c_int32), ("integrator", C.c_int32), ("reserved", C.c_int32),
    ]


class OrcInfo(C.Structure):
    _fields_ = [("f", C.c_double), ("err", C.c_double), ("mu", C.c_double), ("obj_scale", C.c_double),
                ("status", C.c_int32), ("iters", C.c_int32), ("n_reg", C.c_int32), ("n_backtrack", C.c_int32),
                ("n_resto", C.c_int32), ("n_resto_iter", C.c_int32)]


def _host_tag() -> str:
    """fingerprint of the host CPU's instruction-set flags: the library is built with -march=native and travels
    with the repo snapshot, so it must be rebuilt on a machine with a different CPU"""
    import hashlib

    try:
        with open("/proc/cpuinfo") as fh:
            flags = next((line for line in fh if line.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(flags.encode()).hexdigest()


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    tag_file = so + ".host"
    srcs = [os.path.join(HERE, f) for f in ("mpc_oracle.c", "mpc_oracle.h", "dyn_model_gen.h")]
    tag = _host_tag()
    same_host = os.path.exists(tag_file) and open(tag_file).read().strip() == tag
    if force or not os.path.exists(so) or not same_host or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(tag_file, "w") as fh:
            fh.write(tag + "\n")
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        _lib.orc_solve.argtypes = [C.POINTER(OrcCfg), dp, dp, dp, dp, dp, dp, C.POINTER(OrcInfo)]
        _lib.orc_solve_batch.argtypes = [C.POINTER(OrcCfg), C.c_int, dp, dp, dp, dp, dp, dp, ip, ip, dp, C.c_int]
        _lib.orc_newton_step.argtypes = [C.POINTER(OrcCfg), dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
    return _lib


def make_cfg(kind: str, N: int | None = None, M: int = 1, params: Params | None = None, init_mode: int = 1,
             mu_init: float = 30.0, max_iter: int = 100, tol: float = 1e-8, cbf_gamma: float | None = None,
             ref_trajectory: bool = False, rows_as_shipped: bool = False, restoration: bool = False,
             resto_max_calls: int = 1, integrator: str = "euler") -> OrcCfg:
    p = params or Params()
    w = reference_weights(kind)
    c = OrcCfg()
    c.model = 1 if kind == "dyn" else 0
    c.N = N if N is not None else p.N_p
    c.M = 0 if kind == "kin_nocbf" else M
    c.obs_mode = {"kin_nocbf": 0, "kin_cbf": 1, "kin_cbf_pre": 1, "dyn": 2}[kind]
    c.du0_cost = int(w.du0_cost)
    nx = 6 if kind == "dyn" else 4
    for i in range(nx):
        c.Q[i] = w.Q[i]
        c.x_lo[i], c.x_hi[i] = -INF, INF
    for i in range(2):
        c.R[i], c.DR[i] = w.R[i], w.DR[i]
    c.T = p.T_S
    if kind in ("kin_cbf", "kin_cbf_pre"):
        c.n_rate = 1
        c.rate_ctrl[0] = 0
        c.rate_lo[0], c.rate_hi[0] = p.df_dot_min * p.T_S, p.df_dot_max * p.T_S
    elif kind == "dyn":
        c.n_rate = 2
        c.rate_ctrl[0], c.rate_ctrl[1] = 0, 1
        c.rate_lo[0], c.rate_hi[0] = p.df_dot_min * p.T_S, p.df_dot_max * p.T_S
        c.rate_lo[1], c.rate_hi[1] = p.jerk_min * p.T_S, p.jerk_max * p.T_S
    c.u_lo[0], c.u_hi[0] = p.df_min, p.df_max
    c.u_lo[1], c.u_hi[1] = p.ax_min, p.ax_max
    c.x_lo[1], c.x_hi[1] = p.Y_min, p.Y_max
    c.x_lo[3], c.x_hi[3] = p.vx_min, p.vx_max
    if nx == 6:
        c.x_lo[4], c.x_hi[4] = p.vy_min, p.vy_max
    c.obs_lo = 1.0 if kind == "dyn" else 0.0
    c.ego_hl, c.ego_hw, c.safe_l, c.safe_w = p.Veh_L / 2, p.Veh_W / 2, 1.0, 0.5
    c.dyn_sx, c.dyn_sy = 4.0, 1.0
    c.Veh_l, c.Veh_lf, c.Veh_lr, c.Veh_m, c.Veh_Iz = p.Veh_l, p.Veh_lf, p.Veh_lr, p.Veh_m, p.Veh_Iz
    c.aopt_f, c.aopt_r, c.Fymax_f, c.Fymax_r = p.aopt_f, p.aopt_r, p.Fymax_f, p.Fymax_r
    c.init_mode = init_mode
    c.max_iter = max_iter
    c.tol, c.mu_init, c.bound_relax = tol, mu_init, 1e-8
    if cbf_gamma is not None:  # the commented row of PKG/MPC_CBF_optimize_kin.py:244-248
        assert kind in ("kin_cbf", "kin_cbf_pre")
        c.obs_mode, c.cbf_gamma = 3, cbf_gamma
    c.ref_mode = int(ref_trajectory)  # xs is (N, nx) per-stage cost targets
    c.rows_as_shipped = int(rows_as_shipped)  # dyn: bound lists as PKG/MPC_CBF_optimize_dyn.py:112-133 ships them
    c.restoration = int(restoration and kind in ("kin_cbf", "kin_cbf_pre"))
    c.resto_max_calls = int(resto_max_calls)
    assert integrator in ("euler", "rk4") and (integrator == "euler" or kind != "dyn")
    c.integrator = int(integrator == "rk4")
    return c


def _xs(cfg, xs, B=None):
    nx = 6 if cfg.model == 1 else 4
    shape = (cfg.N, nx) if cfg.ref_mode else (nx,)
    return np.ascontiguousarray(xs, dtype=np.float64).reshape(shape if B is None else (B,) + shape)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def solve(cfg: OrcCfg, x0, xs, obs, z_init=None):
    nx = 6 if cfg.model == 1 else 4
    N = cfg.N
    nv = 2 * N + nx * (N + 1)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(nx)
    xs = _xs(cfg, xs)
    obs = None if obs is None or cfg.obs_mode == 0 else np.ascontiguousarray(obs, dtype=np.float64).reshape(cfg.M, N + 1, 6)
    zi = None if z_init is None else np.ascontiguousarray(z_init, dtype=np.float64).reshape(nv)
    z = np.zeros(nv)
    lam = np.zeros(nx * (N + 1))
    info = OrcInfo()
    rc = lib().orc_solve(C.byref(cfg), _dp(x0), _dp(xs), _dp(obs), _dp(zi), _dp(z), _dp(lam), C.byref(info))
    assert rc == 0, rc
    return z, lam, info


def solve_batch(cfg: OrcCfg, x0, xs, obs, z_init=None, want_z=False, nthreads=1):
    nx = 6 if cfg.model == 1 else 4
    N = cfg.N
    nv = 2 * N + nx * (N + 1)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(-1, nx)
    B = x0.shape[0]
    xs = _xs(cfg, xs, B)
    obs = None if obs is None or cfg.obs_mode == 0 else np.ascontiguousarray(obs, dtype=np.float64).reshape(B, cfg.M, N + 1, 6)
    zi = None if z_init is None else np.ascontiguousarray(z_init, dtype=np.float64).reshape(B, nv)
    u0 = np.zeros((B, 2))
    cost = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    z = np.zeros((B, nv)) if want_z else None
    ip = C.POINTER(C.c_int32)
    rc = lib().orc_solve_batch(C.byref(cfg), B, _dp(x0), _dp(xs), _dp(obs), _dp(zi), _dp(u0), _dp(cost),
                               status.ctypes.data_as(ip), iters.ctypes.data_as(ip), _dp(z), nthreads)
    assert rc == 0, rc
    return u0, cost, status, iters, z


def newton_step(cfg: OrcCfg, x0, xs, obs, z, mu, dw, obj_scale):
    nx = 6 if cfg.model == 1 else 4
    N = cfg.N
    nv = 2 * N + nx * (N + 1)
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(nx)
    xs = _xs(cfg, xs)
    obs = None if obs is None or cfg.obs_mode == 0 else np.ascontiguousarray(obs, dtype=np.float64).reshape(cfg.M, N + 1, 6)
    z = np.ascontiguousarray(z, dtype=np.float64).reshape(nv)
    dz = np.zeros(nv)
    lamp = np.zeros(nx * (N + 1))
    rc = lib().orc_newton_step(C.byref(cfg), _dp(x0), _dp(xs), _dp(obs), _dp(z), mu, dw, obj_scale, _dp(dz), _dp(lamp))
    return rc, dz, lamp
