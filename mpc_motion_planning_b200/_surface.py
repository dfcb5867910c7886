"""Shared implementation of the reference-surface classes (`MPC_optimize` in
MPC_CBF_optimize_kin / _kin_pre / _dyn / MPC_optimize_kin).

Mirrors PKG/MPC_CBF_optimize_kin.py:10-255: same constructor side effects (reads
mpc_parameters.yaml, computes N_p, converts degrees), same `initialize_constraints` lists,
and `optimize_problem(...)` returning a callable with CasADi's call convention
`solver(x0=, p=, lbx=, ubx=, lbg=, ubg=) -> {'x','f','g','lam_g','lam_x'}` plus `.stats()`.
The solve itself goes through the CUDA library (no CasADi, no CPU fallback).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib
from .helpers import load_config
from .problem import make_cfg

PARAMS_FILE = "mpc_parameters.yaml"


class _DM:
    """Minimal stand-in for casadi.DM: `.full()` and array conversion."""

    def __init__(self, a):
        self._a = np.array(a, dtype=np.float64)
        if self._a.ndim == 1:
            self._a = self._a.reshape(-1, 1)

    def full(self):
        return self._a.copy()

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __float__(self):
        return float(self._a.reshape(-1)[0])

    @property
    def shape(self):
        return self._a.shape


class _ModelFunction:
    """`mpc_solver.f(x, u)` -> object with `.full()` (PKG/MPC_CBF_optimize_kin.py:159, used by
    shift_movement at PKG/main_cbf_kin_c_sim.py:17-18)."""

    def __init__(self, owner):
        self._o = owner

    def __call__(self, x, u):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        return _DM(self._o._rhs(x, u))


class _Solver:
    """Callable returned by optimize_problem; replaces the CasADi `Function` of `nlpsol`."""

    def __init__(self, owner, obs_array, ref_state=None):
        self._o = owner
        self._obs = obs_array
        self._ref = None if ref_state is None else np.asarray(ref_state, dtype=np.float64)
        self._stats = {"success": False, "return_status": "not_run", "iter_count": 0}

    def stats(self):
        return dict(self._stats)

    def __call__(self, x0=None, p=None, lbx=None, ubx=None, lbg=None, ubg=None, lam_x0=None, lam_g0=None):
        o = self._o
        nx, N = o.num_states, o.N_p
        nv = 2 * N + nx * (N + 1)
        p = np.asarray(p, dtype=np.float64).reshape(-1)
        if p.size != 2 * nx:
            raise ValueError(f"p must have {2 * nx} entries [x0; xs], got {p.size}")
        z0 = np.zeros(nv) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1)
        if z0.size != nv:
            raise ValueError(f"x0 must have {nv} entries, got {z0.size}")
        if o.KIND == "dyn" and o.zero_guess == "reintegrate" and not np.any(z0[2 * N:]):
            # The reference's literal first guess (all states zero, PKG/main_cbf_dyn_c_sim.py:47-50) puts vx at its
            # lower bound 0, where the tire model's slip angles divide by vx: after the bound push the linearised Euler
            # map has entries of 630 per stage and the value function of the stage-wise recursion grows by 4e5 per stage -
            # beyond what FP64 can condense, whatever the scaling.  With `zero_guess =
            # "reintegrate"` (the default, announced by a warning) the controls of the guess are kept and the states are
            # re-integrated from p[:nx]; `zero_guess = "as_given"` passes the caller's x0= through untouched.
            warnings.warn("dyn drop-in: the all-zero state guess is replaced by the Euler roll-out of its controls from p[:nx] "
                          "(set MPC_optimize.zero_guess = 'as_given' to pass x0= through unchanged)", RuntimeWarning, stacklevel=2)
            z0 = z0.copy()
            X = np.zeros((N + 1, nx))
            X[0] = p[:nx]
            U = z0[: 2 * N].reshape(N, 2)
            for k in range(N):
                X[k + 1] = X[k] + o.T_S * o._rhs(X[k], U[k])
            z0[2 * N:] = X.reshape(-1)
        bs = o._batch_solver(lbx, ubx, lbg, ubg, self._obs)
        obs = self._obs[None] if self._obs is not None else None
        xs = p[None, nx:].copy()
        if o._stage_reference():
            # ref_X = aa*ref_state[i+1,:] + (1-aa)*P[n:2n], PKG/MPC_CBF_optimize_kin.py:196
            if self._ref is None or self._ref.shape[0] < N + 1:
                raise ValueError("aa != 0 needs a ref_state of N_p+1 rows")
            xs = (o.aa * self._ref[1: N + 1, :nx] + (1 - o.aa) * p[None, nx:])[None]
        out = bs.solve(p[None, :nx].copy(), xs, obs, z0[None], return_z=True, return_lam=True, return_duals=True)
        st = int(out["status"][0])
        self._stats = {"success": st in (_lib.ST_CONVERGED, _lib.ST_ACCEPTABLE), "return_status": _lib.RETURN_STATUS[st],
                       "iter_count": int(out["iters"][0])}
        z = out["z"][0]
        # CasADi's result keys; 'lam_g_eq' = the [init; defects] head of lam_g, kept for callers of
        # the first release
        res = {"x": _DM(z), "f": _DM([out["cost"][0]]), "lam_g": _DM(out["lam_g"][0]), "lam_x": _DM(out["lam_x"][0]),
               "lam_p": _DM(self._lam_p(out["lam"][0], nx)), "lam_g_eq": _DM(out["lam"][0])}
        res["g"] = _DM(o._g_of(z, p, self._obs))
        return res

    @staticmethod
    def _lam_p(lam_eq, nx):
        """dL/dp for p = [x0; xs]: x0 enters only the row X_0 - P[:nx] (PKG/MPC_CBF_optimize_kin.py:191), so its part is
        minus that row's multipliers; the xs part (cost gradient wrt the target) is not formed and reported as NaN."""
        return np.concatenate([-np.asarray(lam_eq[:nx]), np.full(nx, np.nan)])


class MPCOptimizeBase:
    KIND = None  # set by the concrete modules

    def __init__(self):
        # PKG/MPC_CBF_optimize_kin.py:11-36
        self.config = load_config(PARAMS_FILE)
        mp = self.config["mpc_params"]
        self.T_horizon = mp["horizon"]
        self.T_S = mp["T_S"]
        self.pre_time = mp["pre_time"]
        self.T_L = mp["T_L"]
        self.t_ratio = mp["t_ratio"]
        self.is_variable_time = mp["is_variable_time"]
        if self.is_variable_time == True:  # noqa: E712  (the YAML value 'Flase' never equals True, :19)
            t1 = np.arange(0, self.T_horizon * self.t_ratio, self.T_S, dtype=float)
            t2 = np.arange(t1[-1] + self.T_L, t1[-1] + self.T_L + self.T_horizon * (1 - self.t_ratio), self.T_L)
            # the two-rate grid only changes N_p and t_vector: the NLP still steps every stage with T_S
            # (`x_next = f*T_S + X`, :207), so the same kernels serve it
            self.N_p = len(t1) + len(t2)
            self.t_vector = np.concatenate((t1, t2))
        else:
            self.t_vector = np.arange(0, self.T_horizon + self.T_S, self.T_S, dtype=float)
            self.N_p = len(self.t_vector) - 1
        vp, dc, tp = self.config["vehicle_params"], self.config["dynamics_constraints"], self.config["tire_params"]
        kc = self.config["kinematics_constraints"]
        self.Veh_l, self.Veh_L = vp["Veh_l"], vp["Veh_L"]
        self.Veh_W = vp["Veh_W"] if "Veh_W" in vp else vp["Veh_w"]
        self.Veh_w = self.Veh_W
        self.Veh_m, self.Veh_lf, self.Veh_lr, self.Veh_Iz = vp["Veh_m"], vp["Veh_lf"], vp["Veh_lr"], vp["Veh_Iz"]
        self.aopt_f, self.aopt_r, self.Cf_0, self.Cr_0 = tp["aopt_f"], tp["aopt_r"], tp["Cf_0"], tp["Cr_0"]
        self.Fymax_f = self.Cf_0 * self.aopt_f / 2
        self.Fymax_r = self.Cr_0 * self.aopt_r / 2
        self.vy_max, self.vy_min = dc["vy_max"], dc["vy_min"]
        self.jerk_min, self.jerk_max = dc["jerk_min"], dc["jerk_max"]
        self.df_dot_min = dc["df_dot_min"] * np.pi / 180
        self.df_dot_max = dc["df_dot_max"] * np.pi / 180
        self.vx_max, self.vx_min = kc["vx_max"], kc["vx_min"]
        self.ax_max, self.ax_min = kc["ax_max"], kc["ax_min"]
        self.df_max = kc["df_max"] * np.pi / 180
        self.df_min = kc["df_min"] * np.pi / 180
        self.Y_max, self.Y_min = kc["Y_max"], kc["Y_min"]
        self.model_type = self.config["model_type"]
        self.num_states = 6 if self.KIND == "dyn" else 4
        self.num_controls = 2
        self.f = _ModelFunction(self)
        self._solvers = {}
        # solver options: IPOPT's as passed by the reference (:252-253) and this library's own
        self.max_iter = 100
        self.tol = 1e-8
        self.mu_init = 30.0
        self.init = "as_given"  # the CasADi call starts IPOPT at x0= exactly
        # dyn module only: what the solver call does with an all-zero STATE guess (see _Solver.__call__)
        self.zero_guess = "reintegrate"
        # IPOPT answers a failed line search with its restoration phase, as often as it takes (kin-CBF modules)
        self.restoration = True
        self.resto_max_calls = 0
        # shooting defects: "euler" is what the reference builds (:207) and the only mode with reference parity; "rk4" is the
        # classical Runge-Kutta option (kinematic modules; served by the lane engine, without the restoration phase)
        self.integrator = "euler"
        # Two switches the reference keeps as locals of optimize_problem (kin-CBF modules):
        #   aa = 0.0      weight of ref_state in the stage cost target          (:194-197)
        #   gamma = 1.00  with `g.append(h_func)` live and `gamma*h_func + h_dot` commented out (:235-248)
        # `cbf_rows = "dcbf"` activates the commented row with this gamma.
        self.aa = 0.0
        self.gamma = 1.00
        self.cbf_rows = "h"
        # dyn module: `initialize_constraints()` returns the bound lists exactly as the reference ships them
        # ("as_shipped", the default: drop-in behaviour) or in the order of the rows of g ("aligned"); the solver
        # call recognises which of the two it was given from lbg itself.
        self.dyn_bounds = "as_shipped"

    def _stage_reference(self):
        return self.aa != 0.0 and self.KIND != "dyn"

    def _dcbf(self):
        return self.cbf_rows == "dcbf" and self.KIND in ("kin_cbf", "kin_cbf_pre")

    def generate_ref_path(self, x0, xs):
        """Quintic lane-change reference, PKG/MPC_CBF_optimize_kin.py:258-308 (unused by the mains):
        position/velocity/acceleration boundary conditions over T = 3 s (end point xs moved to
        x0.x + xs.v*T), heading in DEGREES from the polynomial's velocity, speed xs.v, then a straight
        extension to the horizon.  Returns (N1 + N2 + 1, 4) rows [x, y, phi_deg, v]."""
        x0 = np.asarray(x0, dtype=np.float64).reshape(-1)
        xs = np.asarray(xs, dtype=np.float64).reshape(-1)
        T, dt = 3.0, 0.1
        n1, n2 = int(T / dt), int((self.T_horizon - T) / dt)
        t = np.linspace(0, T, n1)
        t2 = np.linspace(T, self.T_horizon, n2 + 1)
        # rows: value, first and second derivative of sum c_k t^k at t = 0 and t = T
        pw = np.arange(6)
        rows = []
        for tt in (0.0, T):
            for d in range(3):
                coef = np.array([np.prod(np.arange(k, k - d, -1)) if k >= d else 0.0 for k in pw], dtype=float)
                rows.append(coef * np.array([tt ** (k - d) if k >= d else 0.0 for k in pw]))
        A = np.array(rows)
        x_end = xs[3] * T + x0[0]
        cx = np.linalg.solve(A, np.array([x0[0], x0[3], 0, x_end, xs[3], 0]))
        cy = np.linalg.solve(A, np.array([x0[1], 0, 0, xs[1], 0, 0]))
        poly = lambda c: c[0] + c[1] * t + c[2] * t**2 + c[3] * t**3 + c[4] * t**4 + c[5] * t**5
        dpoly = lambda c: c[1] + 2 * c[2] * t + 3 * c[3] * t**2 + 4 * c[4] * t**3 + 5 * c[5] * t**4
        xt, yt = poly(cx), poly(cy)
        phi = np.arctan2(dpoly(cy), dpoly(cx)) * 180 / np.pi
        v = np.full(n1, xs[3])
        return np.column_stack((np.concatenate((xt, xt[-1] + v[-1] * (t2 - t[-1]))),
                                np.concatenate((yt, np.full_like(t2, yt[-1]))),
                                np.concatenate((phi, np.full_like(t2, phi[-1]))),
                                np.concatenate((v, np.full_like(t2, v[-1])))))

    # ---- ODE right-hand side (host, float64) ---------------------------------------------
    def _rhs(self, x, u):
        if self.KIND == "dyn":  # PKG/MPC_CBF_optimize_dyn.py:156-170
            _, _, phi, vx, vy, r = x
            df, ax = u
            af = df - (vy + self.Veh_lf * r) / vx
            ar = -(vy - self.Veh_lr * r) / vx
            Cf = self.Fymax_f * 2 * self.aopt_f / (self.aopt_f**2 + af**2)
            Cr = self.Fymax_r * 2 * self.aopt_r / (self.aopt_r**2 + ar**2)
            Fcf, Fcr = -Cf * af, -Cr * ar
            return np.array([vx * np.cos(phi) - vy * np.sin(phi), vx * np.sin(phi) + vy * np.cos(phi), r, ax + r * vy,
                             -r * vx + 2 / self.Veh_m * (Fcf * np.cos(df) + Fcr),
                             2 / self.Veh_Iz * (self.Veh_lf * Fcf - self.Veh_lr * Fcr)])
        _, _, phi, vx = x  # PKG/MPC_CBF_optimize_kin.py:153-156
        df, ax = u
        return np.array([vx * np.cos(phi), vx * np.sin(phi), vx * np.tan(df) / self.Veh_l, ax])

    # ---- bound lists (exact lengths / order of the reference) -----------------------------
    def _n_obs(self, obstacle):
        if self.KIND == "kin_cbf":
            return np.asarray(obstacle).shape[0]
        if self.KIND == "kin_cbf_pre":
            return len(obstacle)
        return 0

    def _initialize_constraints(self, obstacle=None):
        N, nx = self.N_p, self.num_states
        lbx, ubx, lbg, ubg = [], [], [], []
        for _ in range(N):
            lbx += [self.df_min, self.ax_min]
            ubx += [self.df_max, self.ax_max]
        for _ in range(N + 1):
            lo = [-np.inf, self.Y_min, -np.inf, self.vx_min]
            hi = [np.inf, self.Y_max, np.inf, self.vx_max]
            if nx == 6:
                lo += [self.vy_min, -np.inf]
                hi += [self.vy_max, np.inf]
            lbx += lo
            ubx += hi
        if self.KIND == "kin_nocbf":  # pyc L85-86: scalar zero bounds
            return 0.0, 0.0, lbx, ubx
        if self.KIND == "dyn":
            rate_lo = [self.df_dot_min * self.T_S, self.jerk_min * self.T_S]
            rate_hi = [self.df_dot_max * self.T_S, self.jerk_max * self.T_S]
            if self.dyn_bounds == "as_shipped":
                # The reference's own list (PKG/MPC_CBF_optimize_dyn.py:112-129): a block of six zeros per i = 0..N
                # with the rate pair appended for 0 < i < N.  Against the rows of g (:215,227-231) this puts the rate
                # bounds on the x / y defects of stage i and zeros on the rate rows - the
                # problem the reference's dyn main actually hands to IPOPT, solved as such.
                for i in range(N + 1):
                    lbg += [0.0] * 6
                    ubg += [0.0] * 6
                    if 0 < i < N:
                        lbg += rate_lo
                        ubg += rate_hi
            else:
                # bounds in the order of the rows of g: [init, d0, d1, (ddf1, dax1), d2, (ddf2, dax2), ...] -
                # what the reference's comments intend (`mpc_solver.dyn_bounds = "aligned"`)
                lbg += [0.0] * 6
                ubg += [0.0] * 6
                for i in range(N):
                    lbg += [0.0] * 6
                    ubg += [0.0] * 6
                    if i > 0:
                        lbg += rate_lo
                        ubg += rate_hi
            for _ in range(N + 1):
                lbg.append(1)
                ubg.append(np.inf)
            return lbg, ubg, lbx, ubx
        for _ in range(N + 1):
            lbg += [0.0] * 4
            ubg += [0.0] * 4
        for i in range(1, N):
            lbg.append(self.df_dot_min * self.T_S)
            ubg.append(self.df_dot_max * self.T_S)
        for _ in range(N):
            for _ in range(self._n_obs(obstacle)):
                lbg.append(0.0)
                ubg.append(np.inf)
        return lbg, ubg, lbx, ubx

    # ---- solver cache keyed by the bounds actually passed at call time --------------------
    def _batch_solver(self, lbx, ubx, lbg, ubg, obs_array):
        from .solver import BatchSolver

        N, nx = self.N_p, self.num_states
        bounds = {}
        if lbx is not None and ubx is not None:
            lbx = np.asarray(lbx, dtype=np.float64).reshape(-1)
            ubx = np.asarray(ubx, dtype=np.float64).reshape(-1)
            ul, uh = lbx[: 2 * N].reshape(N, 2), ubx[: 2 * N].reshape(N, 2)
            xl, xh = lbx[2 * N:].reshape(N + 1, nx), ubx[2 * N:].reshape(N + 1, nx)
            if not (np.all(ul == ul[0]) and np.all(uh == uh[0]) and np.all(xl == xl[0]) and np.all(xh == xh[0])):
                raise NotImplementedError("stage-varying lbx/ubx are not supported by the CUDA path")
            bounds.update(u_lo=ul[0], u_hi=uh[0], x_lo=xl[0], x_hi=xh[0])
        n_rate = {"kin_nocbf": 0, "kin_cbf": 1, "kin_cbf_pre": 1, "dyn": 2}[self.KIND]
        dyn_rows = "aligned"
        if lbg is not None and ubg is not None and n_rate and not np.isscalar(lbg):
            lg = np.asarray(lbg, dtype=np.float64).reshape(-1)
            ug = np.asarray(ubg, dtype=np.float64).reshape(-1)
            if self.KIND == "dyn":
                # aligned: init, d0, d1 then the first rate pair (rows 18,19); as shipped: the pair sits at 12,13
                shipped = bool(lg[12] != 0.0 or ug[12] != 0.0)
                r0 = 12 if shipped else 18
                bounds.update(rate_lo=lg[r0: r0 + 2], rate_hi=ug[r0: r0 + 2])
                dyn_rows = "as_shipped" if shipped else "aligned"
            else:
                r0 = nx * (N + 1)
                bounds.update(rate_lo=lg[r0: r0 + 1], rate_hi=ug[r0: r0 + 1])
            # everything else the caller passes in lbg/ubg must be what the kernels assume: zeros on the equality
            # rows, ONE rate interval for all stages, [obs_lo, inf) on the obstacle rows - in the reference's row order
            self._check_row_bounds(lg, ug, bounds, dyn_rows, 0 if obs_array is None else obs_array.shape[0])
        M = 0 if obs_array is None else obs_array.shape[0]
        gamma = float(self.gamma) if self._dcbf() else None
        ref = "trajectory" if self._stage_reference() else "terminal"
        rk4 = self.integrator == "rk4"
        key = (M, self.max_iter, self.tol, self.mu_init, self.init, gamma, ref, dyn_rows, self.restoration, self.resto_max_calls, self.integrator) + tuple(np.concatenate([np.ravel(v) for v in bounds.values()]).tolist() if bounds else ())
        if key not in self._solvers:
            self._solvers[key] = BatchSolver(self.KIND, config=self.config, N=N, M=max(M, 1), init=self.init, mu_init=self.mu_init,
                                             max_iter=self.max_iter, tol=self.tol, bounds=bounds or None, cbf_gamma=gamma, ref=ref,
                                             dyn_bounds=dyn_rows, restoration=self.restoration and not rk4, resto_max_calls=self.resto_max_calls,
                                             integrator=self.integrator)
        return self._solvers[key]

    def _check_row_bounds(self, lg, ug, bounds, dyn_rows, M):
        N, nx = self.N_p, self.num_states
        rl, rh = np.ravel(bounds["rate_lo"]), np.ravel(bounds["rate_hi"])
        elo, ehi = [], []
        if self.KIND == "dyn":
            if dyn_rows == "as_shipped":
                for i in range(N + 1):
                    elo += [0.0] * 6; ehi += [0.0] * 6
                    if 0 < i < N:
                        elo += list(rl); ehi += list(rh)
            else:
                elo += [0.0] * 6; ehi += [0.0] * 6
                for i in range(N):
                    elo += [0.0] * 6; ehi += [0.0] * 6
                    if i > 0:
                        elo += list(rl); ehi += list(rh)
            elo += [1.0] * (N + 1); ehi += [np.inf] * (N + 1)
        else:
            elo += [0.0] * (nx * (N + 1)); ehi += [0.0] * (nx * (N + 1))
            elo += [rl[0]] * (N - 1); ehi += [rh[0]] * (N - 1)
            elo += [0.0] * (N * M); ehi += [np.inf] * (N * M)
        if lg.size != len(elo) or not (np.array_equal(lg, elo) and np.array_equal(ug, ehi)):
            raise NotImplementedError("lbg/ubg differ from the pattern of initialize_constraints (zeros on the equality rows, one rate "
                                      "interval for every stage, [lb, inf) on the obstacle rows): row-varying bounds are not supported "
                                      "by the CUDA path")

    # ---- g(z) in the reference's row order (host evaluation for res['g']) -----------------
    def _g_of(self, z, p, obs):
        N, nx = self.N_p, self.num_states
        U = z[: 2 * N].reshape(N, 2)
        X = z[2 * N:].reshape(N + 1, nx)
        g = [X[0] - p[:nx]]
        rate = []
        for i in range(N):
            g.append(X[i + 1] - (X[i] + self.T_S * self._rhs(X[i], U[i])))
            if self.KIND == "dyn" and i > 0:
                g.append(U[i] - U[i - 1])
            elif self.KIND in ("kin_cbf", "kin_cbf_pre") and i > 0:
                rate.append(U[i, 0:1] - U[i - 1, 0:1])
        g += rate
        if obs is not None:
            stages = range(N + 1) if self.KIND == "dyn" else range(N)
            for i in stages:
                for j in range(obs.shape[0]):
                    o = obs[j, i]
                    if self.KIND == "dyn":
                        sx, sy = 4.0, 1.0
                    else:
                        sx = self.Veh_L / 2 + o[4] / 2 + 1.0
                        sy = self.Veh_W / 2 + o[5] / 2 + 0.5
                    e = (X[i, 0] - o[0]) ** 2 / sx**2 + (X[i, 1] - o[1]) ** 2 / sy**2 - 1
                    if self._dcbf():  # gamma*h_func + h_dot, h_func_next at the same obstacle row (:245-248)
                        en = (X[i + 1, 0] - o[0]) ** 2 / sx**2 + (X[i + 1, 1] - o[1]) ** 2 / sy**2 - 1
                        e = self.gamma * e + (en - e)
                    g.append(np.array([np.sqrt(e) if self.KIND == "dyn" else e]))
        return np.concatenate(g)

    # ---- obstacle argument -> (M, N+1, 6) array ---------------------------------------------
    def _obs_array(self, obstacle):
        N = self.N_p
        if self.KIND == "kin_nocbf":
            return None
        if self.KIND == "kin_cbf":  # static (M,6) rows, PKG/MPC_CBF_optimize_kin.py:236-243
            ob = np.asarray(obstacle, dtype=np.float64).reshape(-1, 6)
            return np.repeat(ob[:, None, :], N + 1, axis=1)
        if self.KIND == "kin_cbf_pre":  # list of (>=N,6) trajectories, _kin_pre.py:239-247
            out = np.zeros((len(obstacle), N + 1, 6))
            for j, tr in enumerate(obstacle):
                tr = np.asarray(tr, dtype=np.float64)
                n = min(tr.shape[0], N + 1)
                out[j, :n] = tr[:n]
                if n < N + 1:
                    out[j, n:] = tr[n - 1]
            return out
        ob = np.asarray(obstacle, dtype=np.float64).reshape(-1)  # dyn: centre (x,y), _dyn.py:238-239
        out = np.zeros((1, N + 1, 6))
        out[0, :, 0], out[0, :, 1] = ob[0], ob[1]
        return out
