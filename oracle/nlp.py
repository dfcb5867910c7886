"""CPU restatement of the reference NLPs (TEST INFRASTRUCTURE — not product code).

PARITY: the NLP DEFINITION restated here is pinned - tests/test_reference_vectors.py checks
objective, constraint rows (reference order) and bound lists against vectors the reference's own
`optimize_problem` / `initialize_constraints` produced when run unmodified on a sympy-backed
`casadi` stand-in (tests/golden/make_reference_vectors.py -> tests/golden/reference_nlp.npz;
f to 1e-14 relative, g to 1e-11).  The SOLVE is UNPINNED: the reference ships no tests or golden
vectors and CasADi + IPOPT (unpinned PyPI `casadi`) cannot be installed here (no network), so
the interior-point method in oracle/ipm_dense.py / mpc_oracle.c follows IPOPT's published
algorithm, not IPOPT's outputs.  (The no-CBF module exists only as a CPython-3.7 .pyc; the
generator executes its bytecode with tests/golden/pyc37.py, so it is pinned the same way.)

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.

What is restated (reference file:line, PKG = CasaDi_MPC_Optimize_Multishoot):

* decision vector  z = [vec(U) ; vec(X)] column-major, U (2,N), X (nx,N+1)
  -> [d0,a0,d1,a1,..., X_0(nx), X_1(nx), ...]       PKG/MPC_CBF_optimize_kin.py:250
* parameter        P = [x0 ; xs]                     PKG/MPC_CBF_optimize_kin.py:165
* cost             sum_{i<N} (X_i-xs)'Q(X_i-xs) + U_i'R U_i + dU_i' DR dU_i
                                                      PKG/MPC_CBF_optimize_kin.py:195-205
* Euler defects    X_{i+1} - (X_i + T_S f(X_i,U_i))  PKG/MPC_CBF_optimize_kin.py:207-208
* rate rows        U[0,i]-U[0,i-1], i=1..N-1         PKG/MPC_CBF_optimize_kin.py:211-216
* obstacle rows    (x-ox)^2/sX^2+(y-oy)^2/sY^2-1     PKG/MPC_CBF_optimize_kin.py:236-247
                   per-step centres                   PKG/MPC_CBF_optimize_kin_pre.py:239-253
* bounds           initialize_constraints             PKG/MPC_CBF_optimize_kin.py:84-134
* dyn model        tire model + rhs + rows            PKG/MPC_CBF_optimize_dyn.py:156-243
* no-CBF kin       recovered from PKG/__pycache__/MPC_optimize_kin.cpython-37.pyc
                   

Two options the reference carries but ships switched off:

* `cbf_gamma`      the commented row `gamma*h_func + h_dot` with
                   h_dot = h(X_{i+1}; obs_i) - h(X_i; obs_i)       PKG/MPC_CBF_optimize_kin.py:244-248
                   i.e. h(X_{i+1}; obs_i) - (1-gamma) h(X_i; obs_i) >= 0, i = 0..N-1
                   (both terms use the step-i obstacle, PKG/MPC_CBF_optimize_kin_pre.py:250-254)
* `xref`           per-stage cost targets ref_X = aa*ref_state[i+1] + (1-aa)*xs
                                                      PKG/MPC_CBF_optimize_kin.py:194-199
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

INF = float("inf")

# --------------------------------------------------------------------------
# parameters (values of PKG/mpc_parameters.yaml; restated, file is not read)
# --------------------------------------------------------------------------


@dataclass
class Params:
    T_S: float = 0.1
    horizon: float = 5.0
    Veh_m: float = 1575.0
    Veh_lf: float = 1.2
    Veh_lr: float = 1.6
    Veh_Iz: float = 2875.0
    Veh_l: float = 2.6
    Veh_W: float = 1.8
    Veh_L: float = 4.8
    aopt_f: float = 0.3490658503988659
    aopt_r: float = 0.19198621771937624
    Cf_0: float = -50000.0
    Cr_0: float = -50000.0
    vx_max: float = 40.0
    vx_min: float = 0.0
    ax_max: float = 3.0
    ax_min: float = -3.0
    df_max_deg: float = 35.0
    df_min_deg: float = -35.0
    Y_max: float = 5.0
    Y_min: float = -1.0
    vy_max: float = 5.0
    vy_min: float = -5.0
    jerk_min: float = -3.0
    jerk_max: float = 1.5
    df_dot_min_deg: float = -5.0
    df_dot_max_deg: float = 5.0

    @property
    def N_p(self) -> int:
        # PKG/MPC_CBF_optimize_kin.py:32-33
        return len(np.arange(0, self.horizon + self.T_S, self.T_S, dtype=float)) - 1

    @property
    def df_max(self):
        return self.df_max_deg * np.pi / 180

    @property
    def df_min(self):
        return self.df_min_deg * np.pi / 180

    @property
    def df_dot_max(self):
        return self.df_dot_max_deg * np.pi / 180

    @property
    def df_dot_min(self):
        return self.df_dot_min_deg * np.pi / 180

    @property
    def Fymax_f(self):
        return self.Cf_0 * self.aopt_f / 2  # PKG/MPC_CBF_optimize_kin.py:55

    @property
    def Fymax_r(self):
        return self.Cr_0 * self.aopt_r / 2


# --------------------------------------------------------------------------
# vehicle models: f(x,u), d f/d(x,u), sum_i lam_i d2 f_i / d(x,u)^2
# --------------------------------------------------------------------------


class KinModel:
    """x=[x,y,phi,vx], u=[df,ax]; PKG/MPC_CBF_optimize_kin.py:153-156."""

    nx = 4
    nu = 2

    def __init__(self, p: Params):
        self.L = p.Veh_l

    def f(self, x, u):
        _, _, phi, v = x
        df, ax = u
        return np.array([v * math.cos(phi), v * math.sin(phi), v * math.tan(df) / self.L, ax])

    def jac(self, x, u):
        """(nx, nx+nu) Jacobian wrt [x;u]."""
        _, _, phi, v = x
        df, _ = u
        c, s, t = math.cos(phi), math.sin(phi), math.tan(df)
        sec2 = 1.0 + t * t
        J = np.zeros((4, 6))
        J[0, 2] = -v * s
        J[0, 3] = c
        J[1, 2] = v * c
        J[1, 3] = s
        J[2, 3] = t / self.L
        J[2, 4] = v * sec2 / self.L
        J[3, 5] = 1.0
        return J

    def hess(self, x, u, lam):
        """sum_i lam_i * d2 f_i / d[x;u]^2, (6,6)."""
        _, _, phi, v = x
        df, _ = u
        c, s, t = math.cos(phi), math.sin(phi), math.tan(df)
        sec2 = 1.0 + t * t
        H = np.zeros((6, 6))
        # f0 = v cos(phi)
        H[2, 2] += lam[0] * (-v * c)
        H[2, 3] += lam[0] * (-s)
        H[3, 2] += lam[0] * (-s)
        # f1 = v sin(phi)
        H[2, 2] += lam[1] * (-v * s)
        H[2, 3] += lam[1] * c
        H[3, 2] += lam[1] * c
        # f2 = v tan(df)/L
        H[3, 4] += lam[2] * sec2 / self.L
        H[4, 3] += lam[2] * sec2 / self.L
        H[4, 4] += lam[2] * 2.0 * v * sec2 * t / self.L
        return H


class Rk4KinModel(KinModel):
    """The kinematic bicycle integrated with the classical Runge-Kutta scheme over one step T - the OPTION `north_star`
    names ("RK4 multiple-shooting rollout"; the reference itself integrates with explicit Euler, PKG/MPC_CBF_optimize_kin.py:207,
    and the only Runge-Kutta in its tree is the unfinished tutorial Reference/MPC/sim_test.py:35-37).

    Exposed as an INCREMENT function, so that every use of the model in NLP stays as it is:
        f(x,u) = (Phi(x,u) - x) / T,   Phi = x + T/6 (k1 + 2 k2 + 2 k3 + k4),
        k1 = g(x,u), k2 = g(x + T/2 k1, u), k3 = g(x + T/2 k2, u), k4 = g(x + T k3, u),   g = the Euler right-hand side.
    `jac` is the forward-mode Jacobian through the four stages; `hess` the second-order adjoint
        sum_s Z_s' [nu_s . g''(z_s)] Z_s,   nu_4 = b_4 lam, nu_s = b_s lam + a_{s+1} g_x(z_{s+1})' nu_{s+1},   Z_s = d z_s / d z.
    A = I + T f_x keeps the sparsity of the Euler step (x, y never enter g); B = T f_u gains five entries."""

    A_COEF = (0.0, 0.5, 0.5, 1.0)
    B_COEF = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)

    def __init__(self, p: Params, T: float):
        super().__init__(p)
        self.T = T

    def _stages(self, x, u):
        """stage points z_s = (x + a_s T k_{s-1}, u) and slopes k_s"""
        x = np.asarray(x, float)
        ks, xs_ = [], []
        for a in self.A_COEF:
            xs = x if not ks else x + a * self.T * ks[-1]
            xs_.append(xs)
            ks.append(KinModel.f(self, xs, u))
        return xs_, ks

    def f(self, x, u):
        _, ks = self._stages(x, u)
        return sum(b * k for b, k in zip(self.B_COEF, ks))

    def jac(self, x, u):
        xs_, _ = self._stages(x, u)
        E = np.eye(6)
        J = np.zeros((4, 6))
        K_prev = None
        for a, b, xs in zip(self.A_COEF, self.B_COEF, xs_):
            Z = E.copy()
            if K_prev is not None:
                Z[:4, :] += a * self.T * K_prev
            K_prev = KinModel.jac(self, xs, u) @ Z
            J += b * K_prev
        return J

    def hess(self, x, u, lam):
        lam = np.asarray(lam, float)
        xs_, _ = self._stages(x, u)
        E = np.eye(6)
        Zs, Gs, K_prev = [], [], None
        for a, xs in zip(self.A_COEF, xs_):
            Z = E.copy()
            if K_prev is not None:
                Z[:4, :] += a * self.T * K_prev
            G = KinModel.jac(self, xs, u)
            K_prev = G @ Z
            Zs.append(Z)
            Gs.append(G)
        H = np.zeros((6, 6))
        nu = None
        for s in (3, 2, 1, 0):
            nu = self.B_COEF[s] * lam if nu is None else self.B_COEF[s] * lam + self.A_COEF[s + 1] * self.T * (Gs[s + 1][:, :4].T @ nu)
            H += Zs[s].T @ KinModel.hess(self, xs_[s], u, nu) @ Zs[s]
        return H


class DynModel:
    """x=[x,y,phi,vx,vy,r], u=[df,ax]; PKG/MPC_CBF_optimize_dyn.py:156-170.

    Derivatives come from sympy (generated once per process); the C oracle and
    the CUDA kernel use independently generated/hand-checked code.
    """

    nx = 6
    nu = 2

    def __init__(self, p: Params):
        import sympy as sp

        X, Y, phi, vx, vy, r, df, ax = sp.symbols("x y phi vx vy r df ax", real=True)
        lf, lr, m, Iz = p.Veh_lf, p.Veh_lr, p.Veh_m, p.Veh_Iz
        alpha_f = df - (vy + lf * r) / vx
        alpha_r = -(vy - lr * r) / vx
        Cf = p.Fymax_f * 2 * p.aopt_f / (p.aopt_f**2 + alpha_f**2)
        Cr = p.Fymax_r * 2 * p.aopt_r / (p.aopt_r**2 + alpha_r**2)
        Fcf = -Cf * alpha_f
        Fcr = -Cr * alpha_r
        rhs = sp.Matrix(
            [
                vx * sp.cos(phi) - vy * sp.sin(phi),
                vx * sp.sin(phi) + vy * sp.cos(phi),
                r,
                ax + r * vy,
                -r * vx + 2 / m * (Fcf * sp.cos(df) + Fcr),
                2 / Iz * (lf * Fcf - lr * Fcr),
            ]
        )
        w = sp.Matrix([X, Y, phi, vx, vy, r, df, ax])
        self._f = sp.lambdify(list(w), list(rhs), "math", cse=True)
        self._J = sp.lambdify(list(w), list(rhs.jacobian(w)), "math", cse=True)
        lam = sp.symbols("l0:6", real=True)
        Lsum = sum(lam[i] * rhs[i] for i in range(6))
        self._H = sp.lambdify(list(w) + list(lam), list(sp.hessian(Lsum, list(w))), "math", cse=True)

    def f(self, x, u):
        return np.array(self._f(*x, *u), dtype=float).reshape(6)

    def jac(self, x, u):
        return np.array(self._J(*x, *u), dtype=float).reshape(6, 8)

    def hess(self, x, u, lam):
        return np.array(self._H(*x, *u, *lam), dtype=float).reshape(8, 8)


# --------------------------------------------------------------------------
# NLP
# --------------------------------------------------------------------------

KINDS = ("kin_nocbf", "kin_cbf", "kin_cbf_pre", "dyn")


@dataclass
class Weights:
    Q: np.ndarray
    R: np.ndarray
    DR: np.ndarray
    du0_cost: bool  # is the i=0 control-rate cost (vs Ulast=0) in the objective?


def reference_weights(kind: str) -> Weights:
    if kind == "kin_nocbf":  # pyc facts, SURVEY.md section 8 A0
        return Weights(np.array([10.0, 100.0, 10.0, 100.0]), np.array([100.0, 100.0]), np.array([1e4, 1e4]), False)
    if kind in ("kin_cbf", "kin_cbf_pre"):  # PKG/MPC_CBF_optimize_kin.py:168-184,201-204
        return Weights(np.array([1e1, 1e5, 3e5, 1e4]), np.array([1e4, 1e4]), np.array([1e5, 1e2]), True)
    if kind == "dyn":  # PKG/MPC_CBF_optimize_dyn.py:189-209,221-224
        return Weights(np.array([10.0, 1e5, 1e3, 1e3, 1.0, 1.0]), np.array([1e3, 1e3]), np.array([5e3, 5e2]), False)
    raise ValueError(kind)


class NLP:
    """One scenario of one of the four reference NLPs.

    Internal split used by the solvers:
      eq(z)   = [X_0 - x0 ; defects]                 (nx*(N+1) rows)
      ineq(z) = [rate rows ; obstacle rows]          with [dL, dU]
      zL <= z <= zU
    `g_ref(z)`, `lbg`, `ubg` give the reference's own row order.
    """

    def __init__(self, kind: str, x0, xs, obstacles=None, params: Params | None = None, N: int | None = None,
                 weights: Weights | None = None, cbf_gamma: float | None = None, xref=None, integrator: str = "euler"):
        assert kind in KINDS
        assert integrator in ("euler", "rk4") and (integrator == "euler" or kind != "dyn")
        assert cbf_gamma is None or kind in ("kin_cbf", "kin_cbf_pre")
        self.cbf_gamma = cbf_gamma
        self.kind = kind
        self.p = params or Params()
        self.N = N if N is not None else self.p.N_p
        self.T = self.p.T_S
        self.integrator = integrator
        self.model = DynModel(self.p) if kind == "dyn" else (Rk4KinModel(self.p, self.T) if integrator == "rk4" else KinModel(self.p))
        self.nx, self.nu = self.model.nx, 2
        self.x0 = np.asarray(x0, dtype=float).reshape(self.nx)
        self.xs = np.asarray(xs, dtype=float).reshape(self.nx)
        self.w = weights or reference_weights(kind)
        N, nx = self.N, self.nx
        # per-stage cost target (rows 0..N-1); the reference's aa = 0 makes every row xs
        self.xr = np.repeat(self.xs[None, :], N, axis=0) if xref is None else np.asarray(xref, dtype=float).reshape(N, nx)
        self.nv = 2 * N + nx * (N + 1)
        self.n_eq = nx * (N + 1)
        p = self.p

        # ---- obstacle data -------------------------------------------------
        if kind == "kin_nocbf":
            self.M = 0
            self.obs_stages = []
        elif kind == "kin_cbf":
            ob = np.asarray(obstacles, dtype=float).reshape(-1, 6)  # (M,6) [x,y,th,v,l,w]
            self.M = ob.shape[0]
            self.obs_stages = list(range(N))
            self.oc = np.repeat(ob[:, None, 0:2], N + 1, axis=1)  # (M,N+1,2) static
            self.osx = np.repeat((p.Veh_L / 2 + ob[:, 4] / 2 + 1.0)[:, None], N + 1, axis=1)
            self.osy = np.repeat((p.Veh_W / 2 + ob[:, 5] / 2 + 0.5)[:, None], N + 1, axis=1)
        elif kind == "kin_cbf_pre":
            tr = [np.asarray(t, dtype=float) for t in obstacles]  # list of (>=N,6)
            self.M = len(tr)
            self.obs_stages = list(range(N))
            self.oc = np.zeros((self.M, N + 1, 2))
            self.osx = np.ones((self.M, N + 1))
            self.osy = np.ones((self.M, N + 1))
            for j, t in enumerate(tr):
                n = min(t.shape[0], N + 1)
                self.oc[j, :n] = t[:n, 0:2]
                self.osx[j, :n] = p.Veh_L / 2 + t[:n, 4] / 2 + 1.0
                self.osy[j, :n] = p.Veh_W / 2 + t[:n, 5] / 2 + 0.5
        else:  # dyn: single static obstacle, sqrt form, stages 0..N  (PKG/MPC_CBF_optimize_dyn.py:238-243)
            ob = np.asarray(obstacles, dtype=float).reshape(-1)
            self.M = 1
            self.obs_stages = list(range(N + 1))
            self.oc = np.repeat(ob[None, None, 0:2], N + 1, axis=1)
            self.osx = np.full((1, N + 1), 4.0)
            self.osy = np.full((1, N + 1), 1.0)

        # ---- rate rows -------------------------------------------------------
        if kind in ("kin_cbf", "kin_cbf_pre"):
            self.rate_ctrl = [0]
            self.rate_lo = [p.df_dot_min * self.T]
            self.rate_hi = [p.df_dot_max * self.T]
        elif kind == "dyn":
            self.rate_ctrl = [0, 1]
            self.rate_lo = [p.df_dot_min * self.T, p.jerk_min * self.T]
            self.rate_hi = [p.df_dot_max * self.T, p.jerk_max * self.T]
        else:
            self.rate_ctrl, self.rate_lo, self.rate_hi = [], [], []

        # ---- inequality row table: (type, stage, idx) -------------------------
        rows = []
        for i in range(1, N):
            for k, c in enumerate(self.rate_ctrl):
                rows.append(("rate", i, k))
        for i in self.obs_stages:
            for j in range(self.M):
                rows.append(("obs", i, j))
        self.ineq_rows = rows
        self.n_ineq = len(rows)
        dL, dU = [], []
        for typ, i, k in rows:
            if typ == "rate":
                dL.append(self.rate_lo[k])
                dU.append(self.rate_hi[k])
            else:
                dL.append(1.0 if kind == "dyn" else 0.0)
                dU.append(INF)
        self.dL, self.dU = np.array(dL), np.array(dU)

        # ---- variable bounds (PKG/MPC_CBF_optimize_kin.py:90-105; dyn :91-110) ---
        zL = np.full(self.nv, -INF)
        zU = np.full(self.nv, INF)
        for i in range(N):
            zL[2 * i], zU[2 * i] = p.df_min, p.df_max
            zL[2 * i + 1], zU[2 * i + 1] = p.ax_min, p.ax_max
        for i in range(N + 1):
            b = 2 * N + nx * i
            zL[b + 1], zU[b + 1] = p.Y_min, p.Y_max
            zL[b + 3], zU[b + 3] = p.vx_min, p.vx_max
            if nx == 6:
                zL[b + 4], zU[b + 4] = p.vy_min, p.vy_max
        self.zL, self.zU = zL, zU

    # ---- index helpers ------------------------------------------------------
    def iu(self, i):
        return 2 * i

    def ix(self, i):
        return 2 * self.N + self.nx * i

    def split(self, z):
        N, nx = self.N, self.nx
        U = z[: 2 * N].reshape(N, 2)
        X = z[2 * N:].reshape(N + 1, nx)
        return U, X

    # ---- objective -------------------------------------------------------------
    def objective(self, z):
        U, X = self.split(z)
        w = self.w
        dX = X[: self.N] - self.xr
        f = float(np.sum(w.Q * dX * dX) + np.sum(w.R * U * U))
        dU = np.diff(U, axis=0)
        f += float(np.sum(w.DR * dU * dU))
        if w.du0_cost:
            f += float(np.sum(w.DR * U[0] * U[0]))
        return f

    def grad(self, z):
        U, X = self.split(z)
        w = self.w
        N = self.N
        g = np.zeros(self.nv)
        gU = 2 * w.R * U
        dU = np.diff(U, axis=0)
        gU[1:] += 2 * w.DR * dU
        gU[:-1] -= 2 * w.DR * dU
        if w.du0_cost:
            gU[0] += 2 * w.DR * U[0]
        g[: 2 * N] = gU.reshape(-1)
        gX = np.zeros_like(X)
        gX[:N] = 2 * w.Q * (X[:N] - self.xr)
        g[2 * N:] = gX.reshape(-1)
        return g

    # ---- equality rows -------------------------------------------------------
    def eq(self, z):
        U, X = self.split(z)
        N, nx = self.N, self.nx
        c = np.zeros(self.n_eq)
        c[:nx] = X[0] - self.x0
        for i in range(N):
            c[nx * (i + 1): nx * (i + 2)] = X[i + 1] - (X[i] + self.T * self.model.f(X[i], U[i]))
        return c

    def jac_eq(self, z):
        U, X = self.split(z)
        N, nx = self.N, self.nx
        J = np.zeros((self.n_eq, self.nv))
        J[:nx, self.ix(0): self.ix(0) + nx] = np.eye(nx)
        for i in range(N):
            Jf = self.model.jac(X[i], U[i])
            r = nx * (i + 1)
            J[r: r + nx, self.ix(i + 1): self.ix(i + 1) + nx] = np.eye(nx)
            J[r: r + nx, self.ix(i): self.ix(i) + nx] = -np.eye(nx) - self.T * Jf[:, :nx]
            J[r: r + nx, self.iu(i): self.iu(i) + 2] = -self.T * Jf[:, nx:]
        return J

    # ---- inequality rows --------------------------------------------------------
    def _ell(self, X, i, j, at=None):
        """ellipse function of obstacle j at its step-i position, evaluated at X[at] (default X[i])."""
        at = i if at is None else at
        dx = X[at, 0] - self.oc[j, i, 0]
        dy = X[at, 1] - self.oc[j, i, 1]
        return dx, dy, dx * dx / self.osx[j, i] ** 2 + dy * dy / self.osy[j, i] ** 2 - 1.0

    def ineq(self, z):
        U, X = self.split(z)
        d = np.zeros(self.n_ineq)
        for r, (typ, i, k) in enumerate(self.ineq_rows):
            if typ == "rate":
                c = self.rate_ctrl[k]
                d[r] = U[i, c] - U[i - 1, c]
            else:
                _, _, e = self._ell(X, i, k)
                if self.kind == "dyn":
                    d[r] = math.sqrt(e) if e >= 0 else float("nan")
                elif self.cbf_gamma is not None:
                    _, _, en = self._ell(X, i, k, at=i + 1)
                    d[r] = en - (1.0 - self.cbf_gamma) * e
                else:
                    d[r] = e
        return d

    def jac_ineq(self, z):
        U, X = self.split(z)
        J = np.zeros((self.n_ineq, self.nv))
        for r, (typ, i, k) in enumerate(self.ineq_rows):
            if typ == "rate":
                c = self.rate_ctrl[k]
                J[r, self.iu(i) + c] = 1.0
                J[r, self.iu(i - 1) + c] = -1.0
            else:
                dx, dy, e = self._ell(X, i, k)
                gx = 2 * dx / self.osx[k, i] ** 2
                gy = 2 * dy / self.osy[k, i] ** 2
                if self.kind == "dyn":
                    q = math.sqrt(e) if e > 0 else float("nan")
                    gx, gy = gx / (2 * q), gy / (2 * q)
                if self.cbf_gamma is not None:
                    dxn, dyn_, _ = self._ell(X, i, k, at=i + 1)
                    J[r, self.ix(i + 1) + 0] = 2 * dxn / self.osx[k, i] ** 2
                    J[r, self.ix(i + 1) + 1] = 2 * dyn_ / self.osy[k, i] ** 2
                    gx, gy = -(1.0 - self.cbf_gamma) * gx, -(1.0 - self.cbf_gamma) * gy
                J[r, self.ix(i) + 0] = gx
                J[r, self.ix(i) + 1] = gy
        return J

    # ---- Hessian of the Lagrangian  sigma*f + lam_eq'c + lam_in'd ------------------
    def hess_lag(self, z, lam_eq, lam_in, sigma=1.0):
        U, X = self.split(z)
        N, nx = self.N, self.nx
        w = self.w
        H = np.zeros((self.nv, self.nv))
        for i in range(N):
            for c in range(2):
                H[self.iu(i) + c, self.iu(i) + c] += sigma * 2 * w.R[c]
            for c in range(nx):
                H[self.ix(i) + c, self.ix(i) + c] += sigma * 2 * w.Q[c]
            if i > 0 or w.du0_cost:
                for c in range(2):
                    H[self.iu(i) + c, self.iu(i) + c] += sigma * 2 * w.DR[c]
            if i > 0:
                for c in range(2):
                    H[self.iu(i - 1) + c, self.iu(i - 1) + c] += sigma * 2 * w.DR[c]
                    H[self.iu(i) + c, self.iu(i - 1) + c] -= sigma * 2 * w.DR[c]
                    H[self.iu(i - 1) + c, self.iu(i) + c] -= sigma * 2 * w.DR[c]
            # defects: c_{i+1} = X_{i+1} - X_i - T f(X_i,U_i)
            lam = lam_eq[nx * (i + 1): nx * (i + 2)]
            Hf = -self.T * self.model.hess(X[i], U[i], lam)
            idx = list(range(self.ix(i), self.ix(i) + nx)) + [self.iu(i), self.iu(i) + 1]
            H[np.ix_(idx, idx)] += Hf
        for r, (typ, i, k) in enumerate(self.ineq_rows):
            if typ != "obs" or lam_in[r] == 0.0:
                continue
            a = 2.0 / self.osx[k, i] ** 2
            b = 2.0 / self.osy[k, i] ** 2
            ixx = self.ix(i)
            if self.kind == "dyn":
                dx, dy, e = self._ell(X, i, k)
                q = math.sqrt(e)
                gx, gy = a * dx, b * dy  # gradient of e
                # d = sqrt(e): d'' = e''/(2q) - grad e grad e' /(4 q^3)
                H[ixx, ixx] += lam_in[r] * (a / (2 * q) - gx * gx / (4 * q**3))
                H[ixx + 1, ixx + 1] += lam_in[r] * (b / (2 * q) - gy * gy / (4 * q**3))
                H[ixx, ixx + 1] += lam_in[r] * (-gx * gy / (4 * q**3))
                H[ixx + 1, ixx] += lam_in[r] * (-gx * gy / (4 * q**3))
            elif self.cbf_gamma is not None:
                ixn = self.ix(i + 1)
                H[ixn, ixn] += lam_in[r] * a
                H[ixn + 1, ixn + 1] += lam_in[r] * b
                H[ixx, ixx] -= lam_in[r] * (1.0 - self.cbf_gamma) * a
                H[ixx + 1, ixx + 1] -= lam_in[r] * (1.0 - self.cbf_gamma) * b
            else:
                H[ixx, ixx] += lam_in[r] * a
                H[ixx + 1, ixx + 1] += lam_in[r] * b
        return H

    # ---- the reference's own g ordering and bound lists -------------------------------
    def g_ref(self, z):
        c, d = self.eq(z), self.ineq(z)
        return np.concatenate([c, d])[self.g_perm()]

    def g_perm(self):
        """indices into [eq ; ineq] giving the reference's row order."""
        N, nx = self.N, self.nx
        if self.kind != "dyn":
            return np.arange(self.n_eq + self.n_ineq)  # init, defects, rate, obstacle
        perm = list(range(nx))  # init
        n_rate = len(self.rate_ctrl)
        for i in range(N):
            perm += list(range(nx * (i + 1), nx * (i + 2)))
            if i > 0:
                base = self.n_eq + (i - 1) * n_rate
                perm += list(range(base, base + n_rate))
        base = self.n_eq + (N - 1) * n_rate
        perm += list(range(base, base + (N + 1)))
        return np.array(perm)

    def lbg_ubg_aligned(self):
        lo = np.concatenate([np.zeros(self.n_eq), self.dL])[self.g_perm()]
        hi = np.concatenate([np.zeros(self.n_eq), self.dU])[self.g_perm()]
        return lo, hi

    # ---- start points -----------------------------------------------------------------
    def zero_start(self):
        return np.zeros(self.nv)  # PKG/main_cbf_kin_c_sim.py:47-50,92

    def rollout_start(self, U=None):
        N, nx = self.N, self.nx
        U = np.zeros((N, 2)) if U is None else np.asarray(U, float).reshape(N, 2)
        X = np.zeros((N + 1, nx))
        X[0] = self.x0
        for i in range(N):
            X[i + 1] = X[i] + self.T * self.model.f(X[i], U[i])
        return np.concatenate([U.reshape(-1), X.reshape(-1)])


# --------------------------------------------------------------------------
# reference default scenarios (inputs only)
# --------------------------------------------------------------------------


def default_scenario(kind: str, **kw) -> NLP:
    if kind == "kin_cbf":  # PKG/main_cbf_kin_c_sim.py:45-55
        return NLP(kind, [0, 3, 0, 15], [400, 3.5, 0, 30], [[50, 3.5, 0, 8, 4.8, 1.8]], **kw)
    if kind == "kin_cbf_pre":  # PKG/main_cbf_kin_c_sim_pre.py:45-56
        p = kw.get("params") or Params()
        N = kw.get("N") or p.N_p
        tr = predict_obstacles([np.array([[50, 3.5, 0, 10, 4.8, 1.8]])], p.T_S, N)
        return NLP(kind, [0, 3, 0, 15], [400, 3.5, 0, 30], tr, **kw)
    if kind == "kin_nocbf":  # PKG/main_kin_c_sim.py:42-46
        return NLP(kind, [0, 0, 0, 20], [500, 3.5, 0, 30], None, **kw)
    if kind == "dyn":  # PKG/main_cbf_dyn_c_sim.py:44-51
        return NLP(kind, [0, 0, 0, 10, 0, 0], [600, 3.5, 0, 15, 0, 0], [100, -3.5], **kw)
    raise ValueError(kind)


def predict_obstacles(obs_list, dt, N_p):
    """Constant-velocity roll-out, restating PKG/Obs_prediction.py:3-40 (vectorised)."""
    out = []
    k = np.arange(N_p + 1, dtype=float)
    for ob in obs_list:
        x, y, th, v, l, w = np.asarray(ob, float).reshape(6)
        tr = np.zeros((N_p + 1, 6))
        # the reference accumulates x += v cos(th) dt step by step; do the same to stay bit-identical
        xs_, ys_ = [x], [y]
        for _ in range(N_p):
            xs_.append(xs_[-1] + v * np.cos(th) * dt)
            ys_.append(ys_[-1] + v * np.sin(th) * dt)
        tr[:, 0], tr[:, 1] = xs_, ys_
        tr[:, 2], tr[:, 3], tr[:, 4], tr[:, 5] = th, v, l, w
        out.append(tr)
    return out


class ShippedDynNLP(NLP):
    """The dyn NLP with the bound lists exactly as `initialize_constraints` ships them
    (PKG/MPC_CBF_optimize_dyn.py:112-133), which do not line up with the rows of g (:215,227-243):
    the rate-bound pair of block i lands on the x and y defects of stage i, and the rate rows
    themselves meet the zeros of the next block.  The problem IPOPT would be given is therefore

        U_i = U_{i-1}, i = 1..N-1          (every rate row is an equality: one control pair for the horizon)
        x, y defects into stages 2..N within [df_dot_min*T, df_dot_max*T], [jerk_min*T, jerk_max*T]
        everything else as in the aligned problem.

    Rows are re-partitioned by their shipped bounds: lo == hi -> equality row, else inequality row.
    Oracle-level only (numpy NLP + dense interior point): the CUDA kernels and the C oracle solve the
    aligned problem."""

    def __init__(self, x0, xs, obstacles, **kw):
        super().__init__("dyn", x0, xs, obstacles, **kw)
        N, p = self.N, self.p
        lo, hi = [], []
        for i in range(N + 1):  # the reference's own loop, :112-129
            lo += [0.0] * 6
            hi += [0.0] * 6
            if 0 < i < N:
                lo += [p.df_dot_min * self.T, p.jerk_min * self.T]
                hi += [p.df_dot_max * self.T, p.jerk_max * self.T]
        lo += [1.0] * (N + 1)  # :131-133
        hi += [INF] * (N + 1)
        self.lbg_shipped, self.ubg_shipped = np.array(lo), np.array(hi)
        self._perm = super().g_perm()                 # reference row r = base row perm[r]
        self._is_eq = self.lbg_shipped == self.ubg_shipped
        self._rows_eq = np.where(self._is_eq)[0]
        self._rows_in = np.where(~self._is_eq)[0]
        self._base_n_eq, self._base_n_ineq = self.n_eq, self.n_ineq
        self.n_eq, self.n_ineq = len(self._rows_eq), len(self._rows_in)
        self._eq_target = self.lbg_shipped[self._rows_eq]
        self.dL, self.dU = self.lbg_shipped[self._rows_in], self.ubg_shipped[self._rows_in]

    def _g_base(self, z):
        return np.concatenate([NLP.eq(self, z), NLP.ineq(self, z)])[self._perm]

    def _J_base(self, z):
        return np.vstack([NLP.jac_eq(self, z), NLP.jac_ineq(self, z)])[self._perm]

    def eq(self, z):
        return self._g_base(z)[self._rows_eq] - self._eq_target

    def ineq(self, z):
        return self._g_base(z)[self._rows_in]

    def jac_eq(self, z):
        return self._J_base(z)[self._rows_eq]

    def jac_ineq(self, z):
        return self._J_base(z)[self._rows_in]

    def hess_lag(self, z, lam_eq, lam_in, sigma=1.0):
        lam_ref = np.zeros(len(self._perm))
        lam_ref[self._rows_eq] = lam_eq
        lam_ref[self._rows_in] = lam_in
        lam_base = np.zeros(len(self._perm))
        lam_base[self._perm] = lam_ref
        return NLP.hess_lag(self, z, lam_base[: self._base_n_eq], lam_base[self._base_n_eq:], sigma)

    def g_ref(self, z):
        return self._g_base(z)

    def g_perm(self):
        raise NotImplementedError("rows are partitioned by the shipped bounds; use lbg_shipped / ubg_shipped")
