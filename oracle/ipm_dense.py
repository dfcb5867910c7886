"""Dense primal-dual interior-point solver for the restated NLPs (TEST INFRASTRUCTURE).

PARITY UNPINNED (see oracle/nlp.py).  The reference delegates the solve to
CasADi's `nlpsol('ipopt')` (PKG/MPC_CBF_optimize_kin.py:251-254, invoked at
PKG/main_cbf_kin_c_sim.py:100).  IPOPT is not available here, so this file
restates IPOPT's *published* algorithm (Waechter & Biegler, Math. Prog. 106,
2006: slack form for inequality rows, monotone Fiacco-McCormick barrier update,
fraction-to-boundary, filter line search with second-order correction, inertia
correction of the primal block, gradient-based objective scaling) with dense
linear algebra (`scipy.linalg.ldl` for solve + inertia).  It is the *specification*
of the algorithm the C oracle (oracle/mpc_oracle.c) and the CUDA path implement with
a stage-wise Riccati recursion instead of the dense factorisation.

Options passed by the reference: max_iter 100, acceptable_tol 1e-8,
acceptable_obj_change_tol 1e-6; everything else IPOPT defaults.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import scipy.linalg as sla

from .nlp import NLP

INF = float("inf")

# status codes shared with include/mpcb200.h
ST_CONVERGED, ST_ACCEPTABLE, ST_MAXITER, ST_INFEASIBLE, ST_NAN, ST_RESTO_FAILED = 0, 1, 2, 3, 4, 5


@dataclass
class IpmOptions:
    tol: float = 1e-8
    max_iter: int = 100
    mu_init: float = 30.0  # IPOPT default is 0.1; see DESIGN.md (algorithm choices)
    kappa_eps: float = 10.0
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    tau_min: float = 0.99
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    bound_relax: float = 1e-8
    s_max: float = 100.0
    kappa_sigma: float = 1e10
    kappa_d: float = 1e-5
    obj_scale_max_grad: float = 100.0
    # filter
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8  # IPOPT default gamma_phi=1e-8
    delta: float = 1.0
    s_theta: float = 1.1
    s_phi: float = 2.3
    eta_phi: float = 1e-8
    gamma_alpha: float = 0.05
    max_soc: int = 0  # IPOPT default 4; never accepted on these NLPs, dropped from the spec
    kappa_soc: float = 0.99
    # inertia correction
    dw_first: float = 1e-4
    dw_min: float = 1e-20
    dw_max: float = 1e40
    kw_minus: float = 1.0 / 3.0
    kw_plus: float = 8.0
    kw_plus_first: float = 100.0
    dual_inf_tol: float = 1.0
    constr_viol_tol: float = 1e-4
    compl_inf_tol: float = 1e-4
    lam_init_max: float = 1e3
    acceptable_tol: float = 1e-6
    centered_mult_init: bool = False
    verbose: bool = False
    # restoration phase (see _psi and the `resto` branches of solve)
    restoration: bool = False
    resto_rho: float = 1000.0          # IPOPT resto_penalty_parameter
    resto_kappa: float = 0.9           # IPOPT required_infeasibility_reduction
    bound_mult_reset: float = 1000.0   # IPOPT bound_mult_reset_threshold
    resto_max_calls: int = 1           # not IPOPT: the (n+1)-th entry ends with ST_INFEASIBLE; 0 = no cap


@dataclass
class IpmResult:
    z: np.ndarray
    f: float
    status: int
    iters: int
    lam_eq: np.ndarray
    lam_in: np.ndarray
    zl: np.ndarray
    zu: np.ndarray
    err: float
    mu: float
    obj_scale: float
    n_reg: int = 0
    n_soc: int = 0
    n_backtrack: int = 0
    n_resto: int = 0        # times the restoration phase was entered
    n_resto_iter: int = 0   # iterations spent in it (counted in `iters`)


def _push(v, lo, hi, k1, k2):
    v = v.copy()
    both = np.isfinite(lo) & np.isfinite(hi)
    only_lo = np.isfinite(lo) & ~np.isfinite(hi)
    only_hi = ~np.isfinite(lo) & np.isfinite(hi)
    pl = np.where(both, np.minimum(k1 * np.maximum(1, np.abs(lo)), k2 * (hi - lo)), k1 * np.maximum(1, np.abs(lo)))
    pu = np.where(both, np.minimum(k1 * np.maximum(1, np.abs(hi)), k2 * (hi - lo)), k1 * np.maximum(1, np.abs(hi)))
    m = both | only_lo
    v[m] = np.maximum(v[m], lo[m] + pl[m])
    m = both | only_hi
    v[m] = np.minimum(v[m], hi[m] - pu[m])
    return v


def _relax(lo, hi, fac):
    lo2 = np.where(np.isfinite(lo), lo - fac * np.maximum(1, np.abs(lo)), lo)
    hi2 = np.where(np.isfinite(hi), hi + fac * np.maximum(1, np.abs(hi)), hi)
    return lo2, hi2


def _ftb(v, dv, lo, hi, tau):
    """largest a in (0,1] with v+a dv >= lo + (1-tau)(v-lo), same for hi."""
    a = 1.0
    m = np.isfinite(lo) & (dv < 0)
    if m.any():
        a = min(a, float(np.min(-tau * (v[m] - lo[m]) / dv[m])))
    m = np.isfinite(hi) & (dv > 0)
    if m.any():
        a = min(a, float(np.min(tau * (hi[m] - v[m]) / dv[m])))
    return a


def _ftb_dual(z, dz, tau):
    m = dz < 0
    if m.any():
        return min(1.0, float(np.min(-tau * z[m] / dz[m])))
    return 1.0


def _inertia_from_ldl(D):
    n = D.shape[0]
    pos = neg = zero = 0
    i = 0
    while i < n:
        if i + 1 < n and D[i + 1, i] != 0.0:
            ev = np.linalg.eigvalsh(D[i: i + 2, i: i + 2])
            for e in ev:
                if e > 0:
                    pos += 1
                elif e < 0:
                    neg += 1
                else:
                    zero += 1
            i += 2
        else:
            e = D[i, i]
            if e > 0:
                pos += 1
            elif e < 0:
                neg += 1
            else:
                zero += 1
            i += 1
    return pos, neg, zero


def _psi(r, mu, rho):
    """The l1 penalty rho*(p + n) on a row residual r = p - n with p, n >= 0 kept on their central path:
    (p, n) = argmin rho (p + n) - mu (log p + log n) s.t. p - n = r  -- IPOPT's closed form for the start of its
    restoration phase (Waechter & Biegler 2006, eq. 33), used here at EVERY restoration iterate, so p and n never
    become iterates of their own.  Returns psi, psi' (= the row's multiplier rho - mu/p), psi'' (= mu / (p^2 + n^2))."""
    r = np.asarray(r, float)
    b = mu * r / (2 * rho)
    q = np.sqrt(mu * mu + (rho * r) ** 2) / (2 * rho)

    def n_of(a, b_):  # a + q without cancellation when a < 0
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(a >= 0, a + q, b_ / np.where(a >= 0, 1.0, q - a))

    n = n_of((mu - rho * r) / (2 * rho), b)
    p = n_of((mu + rho * r) / (2 * rho), -b)  # p(r) = n(-r)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = rho * (p + n) - mu * (np.log(p) + np.log(n))
    return val, rho - mu / p, mu / (p * p + n * n)


def solve(nlp: NLP, z_init=None, opt: IpmOptions | None = None) -> IpmResult:
    o = opt or IpmOptions()
    nv, ne, ni = nlp.nv, nlp.n_eq, nlp.n_ineq
    zL, zU = _relax(nlp.zL, nlp.zU, o.bound_relax)
    dL, dU = _relax(nlp.dL, nlp.dU, o.bound_relax) if ni else (nlp.dL, nlp.dU)
    hzL, hzU = np.isfinite(zL), np.isfinite(zU)
    hdL, hdU = (np.isfinite(dL), np.isfinite(dU)) if ni else (np.zeros(0, bool), np.zeros(0, bool))

    z = nlp.zero_start() if z_init is None else np.asarray(z_init, float).copy()
    z = _push(z, zL, zU, o.bound_push, o.bound_frac)
    d0 = nlp.ineq(z)
    if not np.all(np.isfinite(d0)):
        return IpmResult(z, float("nan"), ST_NAN, 0, np.zeros(ne), np.zeros(ni), np.zeros(nv), np.zeros(nv), INF, o.mu_init, 1.0)
    s = _push(d0, dL, dU, o.bound_push, o.bound_frac) if ni else np.zeros(0)

    # gradient-based objective scaling at the (pushed) start point
    g0 = nlp.grad(z)
    gmax = float(np.max(np.abs(g0))) if nv else 0.0
    sigma = o.obj_scale_max_grad / gmax if gmax > o.obj_scale_max_grad else 1.0
    sigma = max(sigma, 1e-8)

    zl = np.where(hzL, 1.0, 0.0)
    zu = np.where(hzU, 1.0, 0.0)
    vl = np.where(hdL, 1.0, 0.0)
    vu = np.where(hdU, 1.0, 0.0)
    lam_c = np.zeros(ne)
    lam_d = np.zeros(ni)

    mu = o.mu_init
    tau = max(o.tau_min, 1 - mu)
    if o.centered_mult_init:
        zl = np.where(hzL, mu / np.where(hzL, z - zL, 1.0), 0.0)
        zu = np.where(hzU, mu / np.where(hzU, zU - z, 1.0), 0.0)
        vl = np.where(hdL, mu / np.where(hdL, s - dL, 1.0), 0.0)
        vu = np.where(hdU, mu / np.where(hdU, dU - s, 1.0), 0.0)
    n_bm = int(hzL.sum() + hzU.sum() + hdL.sum() + hdU.sum())
    only_lo_z = hzL & ~hzU
    only_hi_z = ~hzL & hzU
    only_lo_s = hdL & ~hdU
    only_hi_s = ~hdL & hdU

    def barrier(zz, ss, mu_):
        v = sigma * nlp.objective(zz)
        v -= mu_ * (np.sum(np.log(zz[hzL] - zL[hzL])) + np.sum(np.log(zU[hzU] - zz[hzU])))
        v -= mu_ * (np.sum(np.log(ss[hdL] - dL[hdL])) + np.sum(np.log(dU[hdU] - ss[hdU])))
        v += o.kappa_d * mu_ * (np.sum(zz[only_lo_z] - zL[only_lo_z]) + np.sum(zU[only_hi_z] - zz[only_hi_z]))
        v += o.kappa_d * mu_ * (np.sum(ss[only_lo_s] - dL[only_lo_s]) + np.sum(dU[only_hi_s] - ss[only_hi_s]))
        return float(v)

    def theta_of(zz, ss):
        c = nlp.eq(zz)
        dd = nlp.ineq(zz) - ss if ni else np.zeros(0)
        return float(np.sum(np.abs(c)) + np.sum(np.abs(dd))), c, dd

    def kkt_error(mu_, gz, Jc, Jd, c, dd):
        rz = gz + Jc.T @ lam_c + (Jd.T @ lam_d if ni else 0) - zl + zu
        rs = -lam_d - vl + vu
        dual = max(float(np.max(np.abs(rz))), float(np.max(np.abs(rs))) if ni else 0.0)
        prim = max(float(np.max(np.abs(c))), float(np.max(np.abs(dd))) if ni else 0.0)
        comp = 0.0
        for arr in (
            (z[hzL] - zL[hzL]) * zl[hzL],
            (zU[hzU] - z[hzU]) * zu[hzU],
            (s[hdL] - dL[hdL]) * vl[hdL],
            (dU[hdU] - s[hdU]) * vu[hdU],
        ):
            if arr.size:
                comp = max(comp, float(np.max(np.abs(arr - mu_))))
        sum_lam = float(np.sum(np.abs(lam_c)) + np.sum(np.abs(lam_d)))
        sum_z = float(np.sum(zl) + np.sum(zu) + np.sum(vl) + np.sum(vu))
        s_d = max(o.s_max, (sum_lam + sum_z) / max(1, ne + ni + n_bm)) / o.s_max
        s_c = max(o.s_max, sum_z / max(1, n_bm)) / o.s_max
        return max(dual / s_d, prim, comp / s_c), dual, prim, comp

    filt: list[tuple[float, float]] = []
    th0, _, _ = theta_of(z, s)
    theta_min = 1e-4 * max(1.0, th0)
    theta_max = 1e4 * max(1.0, th0)
    dw_last = 0.0
    n_reg = n_soc = n_bt = 0
    status = ST_MAXITER
    err0 = INF
    it = 0

    # ---- restoration phase state.  Entered when the filter line search (or the inertia correction) fails at a point
    # that is not acceptable.  IPOPT minimises rho*||c(x) - p + n||_1-style infeasibility plus a proximity term over all
    # rows; here (DEVIATION, stated in DESIGN.md) the shooting defects stay EQUALITIES - they can always be met, the free
    # states x and phi absorb them - and only the inequality rows d(z) - s are relaxed:
    #     min  zeta/2 ||D_R (z - z_R)||^2 + sum_rows psi_mu(d(z) - s)   s.t.  c(z) = 0, bounds on z and s
    # with psi_mu the l1 penalty whose p, n are eliminated on their central path (_psi), zeta = sqrt(mu),
    # D_R = 1/max(1,|z_R|), rho = 1000 (IPOPT's defaults).  The same interior-point loop runs on it (own filter, own mu).
    # It returns to the regular phase when the original infeasibility has dropped to kappa_resto * theta_R at a point the
    # original filter accepts; it ends with ST_INFEASIBLE when it converges itself (a stationary point of the
    # infeasibility: IPOPT's "Converged to a point of local infeasibility") and ST_RESTO_FAILED when its own line search
    # fails.
    resto = False
    n_resto = n_resto_it = 0
    z_R = DR2 = None
    zeta = 0.0
    mu_o = tau_o = th_R = 0.0
    filt_o: list[tuple[float, float]] = []
    theta_min_o = theta_max_o = 0.0
    it_resto0 = 0
    rho = o.resto_rho

    def resto_obj(zz, ss, mu_):
        """objective of the restoration problem (without the bound barriers)"""
        v = 0.5 * zeta * float(np.sum(DR2 * (zz - z_R) ** 2))
        if ni:
            pv, _, _ = _psi(nlp.ineq(zz) - ss, mu_, rho)
            v += float(np.sum(pv))
        return v

    def bound_barrier(zz, ss, mu_):
        v = -mu_ * (np.sum(np.log(zz[hzL] - zL[hzL])) + np.sum(np.log(zU[hzU] - zz[hzU])))
        v -= mu_ * (np.sum(np.log(ss[hdL] - dL[hdL])) + np.sum(np.log(dU[hdU] - ss[hdU])))
        v += o.kappa_d * mu_ * (np.sum(zz[only_lo_z] - zL[only_lo_z]) + np.sum(zU[only_hi_z] - zz[only_hi_z]))
        v += o.kappa_d * mu_ * (np.sum(ss[only_lo_s] - dL[only_lo_s]) + np.sum(dU[only_hi_s] - ss[only_hi_s]))
        return float(v)

    def in_filter_of(flt, tmax, t_, p_):
        if t_ >= tmax:
            return True
        for ft, fp in flt:
            if t_ >= ft and p_ >= fp:
                return True
        return False

    while True:
        Jc = nlp.jac_eq(z)
        Jd = nlp.jac_ineq(z) if ni else np.zeros((0, nv))
        th, c, dd = theta_of(z, s)  # original measure: ||c||_1 + ||d - s||_1
        if resto:
            # ---- restoration: leave it?  (needs one restoration step; IPOPT: required_infeasibility_reduction 0.9
            # and acceptability to the original filter, to which the point of entry was added)
            if it > it_resto0 and th <= o.resto_kappa * th_R:
                with np.errstate(all="ignore"):
                    phi_o = barrier(z, s, mu_o)
                if np.isfinite(phi_o) and not in_filter_of(filt_o, theta_max_o, th, phi_o):
                    resto = False
                    mu, tau = mu_o, tau_o
                    filt = filt_o
                    theta_min, theta_max = theta_min_o, theta_max_o
                    lam_c = np.zeros(ne)   # constr_mult_reset_threshold = 0: start again from zero multipliers
                    lam_d = np.zeros(ni)
                    if max(zl.max(initial=0), zu.max(initial=0), vl.max(initial=0), vu.max(initial=0)) > o.bound_mult_reset:
                        zl, zu = np.where(hzL, 1.0, 0.0), np.where(hzU, 1.0, 0.0)
                        vl, vu = np.where(hdL, 1.0, 0.0), np.where(hdU, 1.0, 0.0)
                    if o.verbose:
                        print(f"   leaving restoration at it {it}: theta {th:.3e} <= {o.resto_kappa}*{th_R:.3e}")
        if resto:
            th_full = th
            th = float(np.sum(np.abs(c)))  # restoration's own infeasibility: the equalities only
            _, lam_d, psi2 = _psi(dd, mu, rho) if ni else (0, np.zeros(0), np.zeros(0))
            gz = zeta * DR2 * (z - z_R)
        else:
            gz = sigma * nlp.grad(z)
        if resto:
            # kkt_error() with lam_d = psi'(r) and the rows not counted as primal infeasibility
            rzv = gz + Jc.T @ lam_c + (Jd.T @ lam_d if ni else 0) - zl + zu
            rsv = -lam_d - vl + vu

            def kkt_error_r(mu_):
                dual = max(float(np.max(np.abs(rzv))), float(np.max(np.abs(rsv))) if ni else 0.0)
                prim = float(np.max(np.abs(c)))
                comp = 0.0
                for arr in ((z[hzL] - zL[hzL]) * zl[hzL], (zU[hzU] - z[hzU]) * zu[hzU], (s[hdL] - dL[hdL]) * vl[hdL], (dU[hdU] - s[hdU]) * vu[hdU]):
                    if arr.size:
                        comp = max(comp, float(np.max(np.abs(arr - mu_))))
                sum_lam = float(np.sum(np.abs(lam_c)) + np.sum(np.abs(lam_d)))
                sum_z = float(np.sum(zl) + np.sum(zu) + np.sum(vl) + np.sum(vu))
                s_d = max(o.s_max, (sum_lam + sum_z) / max(1, ne + ni + n_bm)) / o.s_max
                s_c = max(o.s_max, sum_z / max(1, n_bm)) / o.s_max
                return max(dual / s_d, prim, comp / s_c), dual, prim, comp

            err0, du0, pr0, co0 = kkt_error_r(0.0)
        else:
            err0, du0, pr0, co0 = kkt_error(0.0, gz, Jc, Jd, c, dd)
        if o.verbose:
            print(f"it {it:3d}{' R' if resto else '  '} f={nlp.objective(z):.10e} th={th:.2e} mu={mu:.1e} err0={err0:.2e} (du {du0:.1e} pr {pr0:.1e} co {co0:.1e})")
        if err0 <= o.tol and du0 <= o.dual_inf_tol and pr0 <= o.constr_viol_tol and co0 <= o.compl_inf_tol:
            # in restoration: a stationary point of the infeasibility that the original filter / reduction test rejected
            status = ST_INFEASIBLE if resto else ST_CONVERGED
            break
        if it >= o.max_iter:
            status = ST_MAXITER
            break
        # barrier parameter update (possibly several decrements)
        while True:
            emu = kkt_error_r(mu)[0] if resto else kkt_error(mu, gz, Jc, Jd, c, dd)[0]
            if emu <= o.kappa_eps * mu and mu > o.tol / 10:
                mu = max(o.tol / 10, min(o.kappa_mu * mu, mu**o.theta_mu))
                tau = max(o.tau_min, 1 - mu)
                filt = []
                if resto:  # psi_mu and zeta move with mu
                    zeta = math.sqrt(mu)
                    _, lam_d, psi2 = _psi(dd, mu, rho) if ni else (0, np.zeros(0), np.zeros(0))
                    gz = zeta * DR2 * (z - z_R)
                    rzv = gz + Jc.T @ lam_c + (Jd.T @ lam_d if ni else 0) - zl + zu
                    rsv = -lam_d - vl + vu
            else:
                break

        # ---- Newton system ---------------------------------------------------------
        sl_z = np.where(hzL, z - zL, 1.0)
        su_z = np.where(hzU, zU - z, 1.0)
        sl_s = np.where(hdL, s - dL, 1.0)
        su_s = np.where(hdU, dU - s, 1.0)
        Sig_z = zl / sl_z + zu / su_z
        Sig_s = vl / sl_s + vu / su_s
        if resto:
            W = nlp.hess_lag(z, lam_c, lam_d, 0.0) + np.diag(zeta * DR2)
        else:
            W = nlp.hess_lag(z, lam_c, lam_d, sigma)
        gphi_z = gz - np.where(hzL, mu / sl_z, 0) + np.where(hzU, mu / su_z, 0)
        gphi_z = gphi_z + o.kappa_d * mu * (only_lo_z.astype(float) - only_hi_z.astype(float))
        gphi_s = -np.where(hdL, mu / sl_s, 0) + np.where(hdU, mu / su_s, 0)
        gphi_s = gphi_s + o.kappa_d * mu * (only_lo_s.astype(float) - only_hi_s.astype(float))
        rz = gphi_z + Jc.T @ lam_c + (Jd.T @ lam_d if ni else 0)
        rs = gphi_s - lam_d

        def kkt_solve(dw, c_rhs, dd_rhs):
            Ds = Sig_s + dw
            if resto:
                # rows enter through the penalty: new row multiplier lam+ = D_eff J dz + t_eff with
                # D_eff = 1/(1/Ds + 1/psi''), t_eff = D_eff (gphi_s/Ds + psi'/psi'');  ds = (lam+ - gphi_s)/Ds
                D = Ds * psi2 / (Ds + psi2)
                t_eff = D * (gphi_s / Ds + lam_d / psi2)
            else:
                D = Ds
                t_eff = D * dd_rhs + gphi_s
            Hc = W + np.diag(Sig_z + dw) + (Jd.T * D) @ Jd
            K = np.zeros((nv + ne, nv + ne))
            K[:nv, :nv] = Hc
            K[:nv, nv:] = Jc.T
            K[nv:, :nv] = Jc
            rhs = np.concatenate([-(gphi_z + Jc.T @ lam_c) - (Jd.T @ t_eff if ni else 0), -c_rhs])
            L, Dm, perm = sla.ldl(K, lower=True)
            pos, neg, zero = _inertia_from_ldl(Dm)
            if not (pos == nv and neg == ne and zero == 0):
                return None
            sol = np.linalg.solve(K, rhs)
            dz = sol[:nv]
            dlc = sol[nv:]
            if resto:
                lam_new = D * (Jd @ dz) + t_eff if ni else np.zeros(0)
                ds = (lam_new - gphi_s) / Ds if ni else np.zeros(0)
                dld = np.zeros(ni)  # the row multipliers are not iterates in restoration (lam_d = psi'(r))
            else:
                ds = Jd @ dz + dd_rhs if ni else np.zeros(0)
                dld = D * ds + rs if ni else np.zeros(0)
            return dz, ds, dlc, dld

        dw = 0.0
        sol = kkt_solve(0.0, c, dd)
        if sol is None:
            n_reg += 1
            dw = o.dw_first if dw_last == 0.0 else max(o.dw_min, o.kw_minus * dw_last)
            while True:
                sol = kkt_solve(dw, c, dd)
                if sol is not None:
                    break
                dw *= o.kw_plus_first if dw_last == 0.0 else o.kw_plus
                if dw > o.dw_max:
                    break
            if sol is not None:
                dw_last = dw
        accepted = False
        if sol is not None:
            dz, ds, dlc, dld = sol
            dzl = np.where(hzL, -zl + (mu - zl * dz) / sl_z, 0.0)
            dzu = np.where(hzU, -zu + (mu + zu * dz) / su_z, 0.0)
            dvl = np.where(hdL, -vl + (mu - vl * ds) / sl_s, 0.0)
            dvu = np.where(hdU, -vu + (mu + vu * ds) / su_s, 0.0)

            a_max = min(_ftb(z, dz, zL, zU, tau), _ftb(s, ds, dL, dU, tau) if ni else 1.0)
            a_dual = min(_ftb_dual(zl, dzl, tau), _ftb_dual(zu, dzu, tau), _ftb_dual(vl, dvl, tau), _ftb_dual(vu, dvu, tau))

            # ---- filter line search ----------------------------------------------------
            if resto:
                phi = resto_obj(z, s, mu) + bound_barrier(z, s, mu)
                gd = float(gphi_z @ dz + gphi_s @ ds + (lam_d @ (Jd @ dz - ds) if ni else 0.0))
            else:
                phi = barrier(z, s, mu)
                gd = float(gphi_z @ dz + (gphi_s @ ds if ni else 0.0))

            def in_filter(t_, p_):
                return in_filter_of(filt, theta_max, t_, p_)

            def acceptable(alpha, t_, p_):
                """returns (ok, armijo_case)"""
                if not (np.isfinite(t_) and np.isfinite(p_)):
                    return False, False
                if in_filter(t_, p_):
                    return False, False
                sw = gd < 0 and alpha * (-gd) ** o.s_phi > o.delta * th**o.s_theta
                if th <= theta_min and sw:
                    ok = p_ <= phi + o.eta_phi * alpha * gd + 10 * np.finfo(float).eps * abs(phi)
                    return ok, True
                ok = (t_ <= (1 - o.gamma_theta) * th) or (p_ <= phi - o.gamma_phi * th + 10 * np.finfo(float).eps * abs(phi))
                return ok, False

            if gd < 0 and th <= theta_min:
                a_min = o.gamma_alpha * min(o.gamma_theta, o.gamma_phi * th / (-gd) if th > 0 else INF,
                                            o.delta * th**o.s_theta / (-gd) ** o.s_phi if th > 0 else INF)
            elif gd < 0:
                a_min = o.gamma_alpha * min(o.gamma_theta, o.gamma_phi * th / (-gd))
            else:
                a_min = o.gamma_alpha * o.gamma_theta
            a_min = max(a_min, 1e-14)

            alpha = a_max
            armijo = False
            first = True
            z_new = s_new = None
            while alpha >= a_min:
                zt, st = z + alpha * dz, s + alpha * ds
                with np.errstate(all="ignore"):
                    tt, ct, ddt = theta_of(zt, st)
                    if resto:
                        tt = float(np.sum(np.abs(ct)))
                        pt = resto_obj(zt, st, mu) + bound_barrier(zt, st, mu) if np.isfinite(tt) else INF
                    else:
                        pt = barrier(zt, st, mu) if np.isfinite(tt) else INF
                ok, arm = acceptable(alpha, tt, pt)
                if ok:
                    accepted, armijo, z_new, s_new = True, arm, zt, st
                    break
                if first and not resto and np.isfinite(tt) and tt >= th and o.max_soc > 0:
                    # second-order correction (IPOPT max_soc; off by default in this specification: with max_soc = 4 it is
                    # never accepted on these NLPs, so the Riccati implementations do not carry it)
                    c_soc, d_soc = alpha * c + ct, alpha * dd + ddt
                    th_old = th
                    for _ in range(o.max_soc):
                        sol2 = kkt_solve(dw, c_soc, d_soc)
                        if sol2 is None:
                            break
                        dz2, ds2, dlc2, dld2 = sol2
                        a2 = min(_ftb(z, dz2, zL, zU, tau), _ftb(s, ds2, dL, dU, tau) if ni else 1.0)
                        zt2, st2 = z + a2 * dz2, s + a2 * ds2
                        with np.errstate(all="ignore"):
                            tt2, ct2, ddt2 = theta_of(zt2, st2)
                            pt2 = barrier(zt2, st2, mu) if np.isfinite(tt2) else INF
                        ok2, arm2 = acceptable(alpha, tt2, pt2)
                        if ok2:
                            accepted, armijo, z_new, s_new = True, arm2, zt2, st2
                            dlc, dld = dlc2, dld2
                            n_soc += 1
                            break
                        if not np.isfinite(tt2) or tt2 > o.kappa_soc * th_old:
                            break
                        th_old = tt2
                        c_soc, d_soc = a2 * c_soc + ct2, a2 * d_soc + ddt2
                    if accepted:
                        break
                first = False
                alpha *= 0.5
                n_bt += 1
        if not accepted:
            # IPOPT returns Solved_To_Acceptable_Level when the line search fails at an acceptable point; the
            # reference's acceptable_tol (1e-8 = tol) can never trigger, IPOPT's default 1e-6 is used for this exit only.
            if not resto and err0 <= o.acceptable_tol:
                status = ST_ACCEPTABLE
                break
            if resto or not o.restoration:
                status = ST_RESTO_FAILED if resto else ST_INFEASIBLE
                break
            if o.resto_max_calls > 0 and n_resto >= o.resto_max_calls:
                status = ST_INFEASIBLE
                break
            # ---- enter the restoration phase at (z, s)
            resto = True
            n_resto += 1
            it_resto0 = it
            phi_e = barrier(z, s, mu)
            filt_o = filt + [((1 - o.gamma_theta) * th, phi_e - o.gamma_phi * th)]
            theta_min_o, theta_max_o = theta_min, theta_max
            mu_o, tau_o, th_R = mu, tau, th
            z_R = z.copy()
            DR2 = 1.0 / np.maximum(1.0, np.abs(z_R)) ** 2
            c_inf = max(float(np.max(np.abs(c))), float(np.max(np.abs(dd))) if ni else 0.0)
            mu = max(mu_o, c_inf)
            tau = max(o.tau_min, 1 - mu)
            zeta = math.sqrt(mu)
            lam_c = np.zeros(ne)
            zl, zu, vl, vu = (np.minimum(rho, a_) for a_ in (zl, zu, vl, vu))
            filt = []
            thc = float(np.sum(np.abs(c)))
            theta_min = 1e-4 * max(1.0, thc)
            theta_max = 1e4 * max(1.0, thc)
            dw_last = 0.0
            if o.verbose:
                print(f"   entering restoration at it {it}: theta_R {th_R:.3e}, mu {mu:.2e}")
            continue
        if not armijo:
            filt.append(((1 - o.gamma_theta) * th, phi - o.gamma_phi * th))
        if o.verbose:
            print(f"       step: a_max={a_max:.2e} alpha={alpha:.2e} a_dual={a_dual:.2e} dw={dw:.1e} |dz|={np.max(np.abs(dz)):.2e}")
        # IPOPT: equality multipliers move with the primal step size
        z, s = z_new, s_new
        lam_c = lam_c + alpha * dlc
        lam_d = lam_d + alpha * dld
        zl = zl + a_dual * dzl
        zu = zu + a_dual * dzu
        vl = vl + a_dual * dvl
        vu = vu + a_dual * dvu
        # kappa_sigma safeguard
        for arr, gap, has in ((zl, z - zL, hzL), (zu, zU - z, hzU), (vl, s - dL, hdL), (vu, dU - s, hdU)):
            if has.any():
                arr[has] = np.maximum(np.minimum(arr[has], o.kappa_sigma * mu / gap[has]), mu / (o.kappa_sigma * gap[has]))
        it += 1
        if resto:
            n_resto_it += 1

    return IpmResult(z, nlp.objective(z), status, it, lam_c / sigma, lam_d / sigma, zl / sigma, zu / sigma, err0, mu, sigma,
                     n_reg, n_soc, n_bt, n_resto, n_resto_it)
