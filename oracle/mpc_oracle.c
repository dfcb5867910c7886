/* CPU oracle for the MPC solve path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * PARITY UNPINNED (see mpc_oracle.h).  Scalar C restatement of
 *   - the reference NLPs       PKG/MPC_CBF_optimize_kin.py:136-255, _kin_pre.py:136-261,
 *                              _dyn.py:137-250, pyc facts of MPC_optimize_kin
 *   - the bounds               PKG/MPC_CBF_optimize_kin.py:84-134, _dyn.py:85-135 (aligned)
 *   - IPOPT's published primal-dual filter interior-point method with the reference's
 *     options (PKG/MPC_CBF_optimize_kin.py:252-253); dense specification in ipm_dense.py.
 * Newton systems are solved with a stage-wise Riccati recursion on the state augmented by
 * the previous control (the control-rate cost/rows couple neighbouring stages).
 */
#include "mpc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <stdatomic.h>
#include "dyn_model_gen.h"

#define NXM 6
#define NST (ORC_NMAX + 1)

/* ---- IPOPT default constants used by the restatement ------------------------------ */
static const double KAPPA_EPS = 10.0, KAPPA_MU = 0.2, THETA_MU = 1.5, TAU_MIN = 0.99;
static const double BOUND_PUSH = 1e-2, BOUND_FRAC = 1e-2, S_MAX = 100.0, KAPPA_SIGMA = 1e10;
static const double KAPPA_D = 1e-5, OBJ_SCALE_MAX_GRAD = 100.0;
static const double GAMMA_THETA = 1e-5, GAMMA_PHI = 1e-8, DELTA_SW = 1.0, S_THETA = 1.1, S_PHI = 2.3;
static const double ETA_PHI = 1e-8, GAMMA_ALPHA = 0.05;
static const double DW_FIRST = 1e-4, DW_MIN = 1e-20, DW_MAX = 1e40, KW_MINUS = 1.0 / 3.0, KW_PLUS = 8.0,
                    KW_PLUS_FIRST = 100.0;
static const double ACCEPTABLE_TOL = 1e-6; /* line-search-failure exit only, see ipm_dense.py */
static const double DUAL_INF_TOL = 1.0, CONSTR_VIOL_TOL = 1e-4, COMPL_INF_TOL = 1e-4;
#define FILTER_CAP 128
/* restoration phase (specification: the `resto` branches of ipm_dense.solve) */
static const double RESTO_RHO = 1000.0, RESTO_KAPPA = 0.9, BOUND_MULT_RESET = 1000.0;

typedef struct {
  /* primal-dual iterate */
  double x[NST][NXM], u[NST][2];
  double lam[NST][NXM];                 /* multipliers of [X0-x0; defects] */
  double zlx[NST][NXM], zux[NST][NXM];  /* state-bound multipliers */
  double zlu[NST][2], zuu[NST][2];      /* control-bound multipliers */
  double sr[NST][2], vlr[NST][2], vur[NST][2], lr[NST][2];                  /* rate rows (stage k>=1) */
  double so[NST][ORC_MMAX], vlo[NST][ORC_MMAX], lo[NST][ORC_MMAX];          /* obstacle rows */
} iterate_t;

typedef struct {
  const orc_cfg *c;
  int nx, N, M;
  double x0[NXM], xr[NST][NXM]; /* start state, per-stage cost target */
  double ocx[NST][ORC_MMAX], ocy[NST][ORC_MMAX], isx2[NST][ORC_MMAX], isy2[NST][ORC_MMAX]; /* centre, 1/sX^2, 1/sY^2 */
  double xlo[NXM], xhi[NXM], ulo[2], uhi[2], rlo[2], rhi[2], olo; /* relaxed bounds */
  int xbl[NXM], xbu[NXM];
  double sigma; /* objective scaling */
  double pdyn[8];
  iterate_t it, tr; /* current and trial */
  /* evaluation at an iterate */
  double cdef[NST][NXM];                /* c_0 = X0-x0, c_k = defect into stage k */
  double A[NST][NXM][NXM], B[NST][NXM][2];
  double dobs[NST][ORC_MMAX], gox[NST][ORC_MMAX], goy[NST][ORC_MMAX]; /* row value and gradient */
  double hoxx[NST][ORC_MMAX], hoxy[NST][ORC_MMAX], hoyy[NST][ORC_MMAX]; /* row Hessian */
  /* ORC_OBS_DCBF: row k also depends on x_{k+1}: gradient and (diagonal) Hessian wrt its (x,y) */
  double gnx[NST][ORC_MMAX], gny[NST][ORC_MMAX], hnxx[NST][ORC_MMAX], hnyy[NST][ORC_MMAX];
  double reso[NST][ORC_MMAX], resr[NST][2];
  /* QP data */
  double Hxx[NST][NXM][NXM], Hux[NST][2][NXM], Huu[NST][2][2], E[NST][2];
  double gx[NST][NXM], gu[NST][2], tk[NST][2];
  double Dr[NST][2], Do[NST][ORC_MMAX], gsr[NST][2], gso[NST][ORC_MMAX];
  /* Riccati */
  double Kx[NST][2][NXM], Kw[NST][2][2], kk[NST][2];
  /* direction */
  double dx[NST][NXM], du[NST][2], lamp[NST][NXM];
  double dsr[NST][2], dso[NST][ORC_MMAX], lrp[NST][2], lop[NST][ORC_MMAX];
  double dzlx[NST][NXM], dzux[NST][NXM], dzlu[NST][2], dzuu[NST][2], dvlr[NST][2], dvur[NST][2], dvlo[NST][ORC_MMAX];
  double filt_t[FILTER_CAP], filt_p[FILTER_CAP];
  int nfilt;
  /* restoration phase: min zeta/2 ||D_R (z - z_R)||^2 + sum_rows psi_mu(row - s)  s.t. the dynamics, bounds */
  int resto;
  double zeta, mu_psi;
  double xR[NST][NXM], uR[NST][2];       /* z_R */
  double p2r[NST][2], p2o[NST][ORC_MMAX]; /* psi''(r) per row (psi' lives in it.lr / it.lo while in restoration) */
  double pe_thc, pe_thr, pe_bar, pe_lin, pe_f; /* pieces of the last eval_primal: |equalities|_1, |rows|_1, log sum, damping sum, objective */
  double filt_ot[FILTER_CAP], filt_op[FILTER_CAP];
  int nfilt_o;
} ws_t;

/* l1 penalty on a row residual r = p - n with (p, n) on their central path for the barrier parameter mu
 * (ipm_dense._psi): value, first and second derivative */
static double psi_n(double a, double b, double q) { return a >= 0 ? a + q : b / (q - a); }
static void psi_eval(double r, double mu, double *val, double *d1, double *d2) {
  const double rho = RESTO_RHO;
  double b = mu * r / (2 * rho), q = sqrt(mu * mu + (rho * r) * (rho * r)) / (2 * rho);
  double n = psi_n((mu - rho * r) / (2 * rho), b, q), p = psi_n((mu + rho * r) / (2 * rho), -b, q);
  if (val) *val = rho * (p + n) - mu * (log(p) + log(n));
  if (d1) { *d1 = rho - mu / p; *d2 = mu / (p * p + n * n); }
}
static double dr2(double v) { double m = fmax(1.0, fabs(v)); return 1.0 / (m * m); }

int orc_nx(const orc_cfg *c) { return c->model == ORC_MODEL_DYN ? 6 : 4; }

/* ------------------------------------------------------------------ vehicle models */
/* kinematic right-hand side g(x,u), its Jacobian wrt z = [x;u] (4x6) and sum_i nu_i g_i'' (6x6)
 * (PKG/MPC_CBF_optimize_kin.py:153-156) */
static void kin_g(double L, const double *x, const double *u, double *f) {
  f[0] = x[3] * cos(x[2]);
  f[1] = x[3] * sin(x[2]);
  f[2] = x[3] * tan(u[0]) / L;
  f[3] = u[1];
}
static void kin_gjac(double L, const double *x, const double *u, double G[4][6]) {
  memset(G, 0, sizeof(double) * 24);
  double c = cos(x[2]), s = sin(x[2]), t = tan(u[0]);
  G[0][2] = -x[3] * s; G[0][3] = c;
  G[1][2] = x[3] * c;  G[1][3] = s;
  G[2][3] = t / L;     G[2][4] = x[3] * (1.0 + t * t) / L;
  G[3][5] = 1.0;
}
static void kin_ghess(double L, const double *x, const double *u, const double *nu, double M[6][6]) {
  memset(M, 0, sizeof(double) * 36);
  double c = cos(x[2]), s = sin(x[2]), t = tan(u[0]), v = x[3], sec2 = 1.0 + t * t;
  M[2][2] = nu[0] * (-v * c) + nu[1] * (-v * s);
  M[2][3] = M[3][2] = nu[0] * (-s) + nu[1] * c;
  M[3][4] = M[4][3] = nu[2] * sec2 / L;
  M[4][4] = nu[2] * 2.0 * v * sec2 * t / L;
}
/* Runge-Kutta increment function (Phi(x,u) - x)/T with its forward-mode Jacobian and second-order adjoint
 * (oracle/nlp.py Rk4KinModel) */
static const double RK_A[4] = {0.0, 0.5, 0.5, 1.0}, RK_B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
static void rk4_all(double L, double T, const double *x, const double *u, const double *lam, double *f, double J[4][6], double H[6][6]) {
  double xs[4][4], k[4][4], Z[4][6][6], G[4][4][6], K[4][6];
  for (int s = 0; s < 4; s++) {
    for (int i = 0; i < 4; i++) xs[s][i] = x[i] + (s ? RK_A[s] * T * k[s - 1][i] : 0.0);
    kin_g(L, xs[s], u, k[s]);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) Z[s][i][j] = (i == j ? 1.0 : 0.0) + ((s && i < 4) ? RK_A[s] * T * K[i][j] : 0.0);
    kin_gjac(L, xs[s], u, G[s]);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 6; j++) { double a = 0; for (int m = 0; m < 6; m++) a += G[s][i][m] * Z[s][m][j]; K[i][j] = a; }
    for (int i = 0; i < 4; i++) {
      if (f) f[i] = (s ? f[i] : 0.0) + RK_B[s] * k[s][i];
      if (J) for (int j = 0; j < 6; j++) J[i][j] = (s ? J[i][j] : 0.0) + RK_B[s] * K[i][j];
    }
  }
  if (H) {
    double nu[4], nn[4];
    memset(H, 0, sizeof(double) * 36);
    for (int s = 3; s >= 0; s--) {
      for (int i = 0; i < 4; i++) {
        double a = RK_B[s] * lam[i];
        if (s < 3) for (int m = 0; m < 4; m++) a += RK_A[s + 1] * T * G[s + 1][m][i] * nu[m];
        nn[i] = a;
      }
      memcpy(nu, nn, sizeof nu);
      double M[6][6], MZ[6][6];
      kin_ghess(L, xs[s], u, nu, M);
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { double a = 0; for (int m = 0; m < 6; m++) a += M[i][m] * Z[s][m][j]; MZ[i][j] = a; }
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { double a = 0; for (int m = 0; m < 6; m++) a += Z[s][m][i] * MZ[m][j]; H[i][j] += a; }
    }
  }
}
#define IS_RK4(w) ((w)->c->integrator == 1 && (w)->c->model == ORC_MODEL_KIN)

static void model_f(const ws_t *w, const double *x, const double *u, double *f) {
  if (IS_RK4(w)) {
    rk4_all(w->c->Veh_l, w->c->T, x, u, 0, f, 0, 0);
  } else if (w->c->model == ORC_MODEL_KIN) {
    kin_g(w->c->Veh_l, x, u, f);
  } else {
    dyn_f(x, u, w->pdyn, f);
  }
}

static void model_jac(const ws_t *w, const double *x, const double *u, double Jx[NXM][NXM], double Ju[NXM][2]) {
  memset(Jx, 0, sizeof(double) * NXM * NXM);
  memset(Ju, 0, sizeof(double) * NXM * 2);
  if (IS_RK4(w)) {
    double J[4][6];
    rk4_all(w->c->Veh_l, w->c->T, x, u, 0, 0, J, 0);
    for (int i = 0; i < 4; i++) {
      for (int j = 0; j < 4; j++) Jx[i][j] = J[i][j];
      Ju[i][0] = J[i][4]; Ju[i][1] = J[i][5];
    }
  } else if (w->c->model == ORC_MODEL_KIN) {
    double c = cos(x[2]), s = sin(x[2]), t = tan(u[0]), L = w->c->Veh_l;
    Jx[0][2] = -x[3] * s;
    Jx[0][3] = c;
    Jx[1][2] = x[3] * c;
    Jx[1][3] = s;
    Jx[2][3] = t / L;
    Ju[2][0] = x[3] * (1.0 + t * t) / L;
    Ju[3][1] = 1.0;
  } else {
    double jx[36] = {0}, ju[12] = {0};
    dyn_jac(x, u, w->pdyn, jx, ju);
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) Jx[i][j] = jx[i * 6 + j];
      Ju[i][0] = ju[i * 2];
      Ju[i][1] = ju[i * 2 + 1];
    }
  }
}

/* H = sum_i lam_i d2 f_i / d[x;u]^2, (nx+2)^2 stored in 8x8 */
static void model_hess(const ws_t *w, const double *x, const double *u, const double *lam, double H[8][8]) {
  memset(H, 0, sizeof(double) * 64);
  if (IS_RK4(w)) {
    double H6[6][6];
    rk4_all(w->c->Veh_l, w->c->T, x, u, lam, 0, 0, H6);
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) H[i][j] = H6[i][j];
  } else if (w->c->model == ORC_MODEL_KIN) {
    double c = cos(x[2]), s = sin(x[2]), t = tan(u[0]), L = w->c->Veh_l, v = x[3];
    double sec2 = 1.0 + t * t;
    H[2][2] = lam[0] * (-v * c) + lam[1] * (-v * s);
    H[2][3] = H[3][2] = lam[0] * (-s) + lam[1] * c;
    H[3][4] = H[4][3] = lam[2] * sec2 / L;
    H[4][4] = lam[2] * 2.0 * v * sec2 * t / L;
  } else {
    double h[64] = {0};
    dyn_hess(x, u, w->pdyn, lam, h);
    /* generated order is [x(6); u(2)] = 8x8 already */
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) H[i][j] = h[i * 8 + j];
  }
}

/* ------------------------------------------------------------------ helpers */
static double relax_lo(double b, double f) { return isfinite(b) ? b - f * fmax(1.0, fabs(b)) : b; }
static double relax_hi(double b, double f) { return isfinite(b) ? b + f * fmax(1.0, fabs(b)) : b; }

static double push_in(double v, double lo, double hi) {
  int bl = isfinite(lo), bu = isfinite(hi);
  if (bl && bu) {
    double pl = fmin(BOUND_PUSH * fmax(1.0, fabs(lo)), BOUND_FRAC * (hi - lo));
    double pu = fmin(BOUND_PUSH * fmax(1.0, fabs(hi)), BOUND_FRAC * (hi - lo));
    v = fmax(v, lo + pl);
    v = fmin(v, hi - pu);
  } else if (bl) {
    v = fmax(v, lo + BOUND_PUSH * fmax(1.0, fabs(lo)));
  } else if (bu) {
    v = fmin(v, hi - BOUND_PUSH * fmax(1.0, fabs(hi)));
  }
  return v;
}

static int has_rate(const ws_t *w, int k) { return k >= 1 && k <= w->N - 1; }
/* dyn with the bound lists as shipped: at stages k = 1..N-1 the slack sr[k][c] (bounds = the rate bounds)
 * relaxes component c = 0,1 (x, y) of the defect into stage k+1, its row multiplier is lam[k+1][c]; the
 * rate rows U_k - U_{k-1} are equalities with multipliers lr[k][c]. */
#define SHIPPED(w) ((w)->c->rows_as_shipped && (w)->c->model == ORC_MODEL_DYN && (w)->c->n_rate == 2)
static int has_obs(const ws_t *w, int k) {
  if (w->c->obs_mode == ORC_OBS_ELLIPSE || w->c->obs_mode == ORC_OBS_DCBF) return k <= w->N - 1;
  if (w->c->obs_mode == ORC_OBS_SQRT) return k <= w->N;
  return 0;
}

/* obstacle row value/gradient/Hessian at (x,y) for stage k, obstacle j; returns 0 on NaN */
static int obs_row(const ws_t *w, int k, int j, double px, double py, double *d, double *gx, double *gy, double *hxx,
                   double *hxy, double *hyy) {
  double dx = px - w->ocx[k][j], dy = py - w->ocy[k][j];
  double a = w->isx2[k][j], b = w->isy2[k][j];
  double e = dx * dx * a + dy * dy * b - 1.0;
  if (w->c->obs_mode != ORC_OBS_SQRT) { /* PKG/MPC_CBF_optimize_kin.py:244,247 */
    *d = e;
    if (gx) { *gx = 2 * dx * a; *gy = 2 * dy * b; *hxx = 2 * a; *hxy = 0; *hyy = 2 * b; }
    return isfinite(e);
  }
  /* PKG/MPC_CBF_optimize_dyn.py:243: sqrt(e) */
  if (!(e > 0.0)) { *d = NAN; return 0; }
  double q = sqrt(e);
  *d = q;
  if (gx) {
    double ex = 2 * dx * a, ey = 2 * dy * b;
    *gx = ex / (2 * q);
    *gy = ey / (2 * q);
    double q3 = 4 * q * q * q;
    *hxx = a / q - ex * ex / q3;
    *hxy = -ex * ey / q3;
    *hyy = b / q - ey * ey / q3;
  }
  return 1;
}

/* ORC_OBS_DCBF row k: h(X_{k+1}; obs_k) - (1-gamma) h(X_k; obs_k), the commented
 * `gamma*h_func + h_dot` of PKG/MPC_CBF_optimize_kin.py:244-248 (both terms at the step-k obstacle,
 * PKG/MPC_CBF_optimize_kin_pre.py:250-254).  With store != 0 the derivatives go to the workspace. */
static int dcbf_row(ws_t *w, const iterate_t *q, int k, int j, double *d, int store) {
  double g1 = 1.0 - w->c->cbf_gamma;
  double e0, e1, gx, gy, hxx, hxy, hyy, nx_, ny_, nxx, nxy, nyy;
  obs_row(w, k, j, q->x[k][0], q->x[k][1], &e0, &gx, &gy, &hxx, &hxy, &hyy);
  /* the obstacle of step k evaluated at the next state */
  double dx = q->x[k + 1][0] - w->ocx[k][j], dy = q->x[k + 1][1] - w->ocy[k][j];
  double a = w->isx2[k][j], b = w->isy2[k][j];
  e1 = dx * dx * a + dy * dy * b - 1.0;
  nx_ = 2 * dx * a; ny_ = 2 * dy * b; nxx = 2 * a; nxy = 0; nyy = 2 * b;
  (void)nxy;
  *d = e1 - g1 * e0;
  if (store) {
    w->gox[k][j] = -g1 * gx; w->goy[k][j] = -g1 * gy;
    w->hoxx[k][j] = -g1 * hxx; w->hoxy[k][j] = -g1 * hxy; w->hoyy[k][j] = -g1 * hyy;
    w->gnx[k][j] = nx_; w->gny[k][j] = ny_; w->hnxx[k][j] = nxx; w->hnyy[k][j] = nyy;
  }
  return isfinite(*d);
}
#define IS_DCBF(w) ((w)->c->obs_mode == ORC_OBS_DCBF)

/* unscaled objective of the primal part of an iterate (PKG/MPC_CBF_optimize_kin.py:195-205) */
static double objective(const ws_t *w, const iterate_t *q) {
  const orc_cfg *c = w->c;
  double f = 0;
  for (int k = 0; k < w->N; k++) {
    for (int i = 0; i < w->nx; i++) { double e = q->x[k][i] - w->xr[k][i]; f += c->Q[i] * e * e; }
    for (int i = 0; i < 2; i++) {
      f += c->R[i] * q->u[k][i] * q->u[k][i];
      if (k > 0) { double e = q->u[k][i] - q->u[k - 1][i]; f += c->DR[i] * e * e; }
      else if (c->du0_cost) f += c->DR[i] * q->u[0][i] * q->u[0][i];
    }
  }
  return f;
}

/* constraint residuals, theta (1-norm), barrier function; returns 0 if not finite */
static int eval_primal(ws_t *w, const iterate_t *q, double mu, double *theta, double *phi, double *fobj,
                       int store) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  double th = 0, thr = 0, bar = 0, lin = 0, fr = 0;
  double cd[NXM];
  for (int i = 0; i < nx; i++) { cd[i] = q->x[0][i] - w->x0[i]; th += fabs(cd[i]); if (store) w->cdef[0][i] = cd[i]; }
  for (int k = 0; k <= N; k++) {
    if (k < N) {
      double f[NXM];
      model_f(w, q->x[k], q->u[k], f);
      for (int i = 0; i < nx; i++) {
        double d = q->x[k + 1][i] - (q->x[k][i] + c->T * f[i]);
        if (SHIPPED(w) && i < 2 && has_rate(w, k)) d -= q->sr[k][i]; /* relaxed row: defect - slack */
        th += fabs(d);
        if (store) w->cdef[k + 1][i] = d;
      }
      for (int i = 0; i < 2; i++) bar += log(q->u[k][i] - w->ulo[i]) + log(w->uhi[i] - q->u[k][i]);
    }
    for (int i = 0; i < nx; i++) {
      if (w->xbl[i]) bar += log(q->x[k][i] - w->xlo[i]);
      if (w->xbu[i]) bar += log(w->xhi[i] - q->x[k][i]);
    }
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        int ci = c->rate_ctrl[r];
        double d = q->u[k][ci] - q->u[k - 1][ci] - (SHIPPED(w) ? 0.0 : q->sr[k][r]);
        if (SHIPPED(w)) th += fabs(d); else thr += fabs(d);
        if (w->resto) { double v; psi_eval(d, mu, &v, 0, 0); fr += v; }
        if (store) w->resr[k][r] = d;
        bar += log(q->sr[k][r] - w->rlo[r]) + log(w->rhi[r] - q->sr[k][r]);
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double d;
        if (IS_DCBF(w)) { if (!dcbf_row(w, q, k, j, &d, 0)) return 0; }
        else if (!obs_row(w, k, j, q->x[k][0], q->x[k][1], &d, 0, 0, 0, 0, 0)) return 0;
        double r_ = d - q->so[k][j];
        thr += fabs(r_);
        if (w->resto) { double v; psi_eval(r_, mu, &v, 0, 0); fr += v; }
        if (store) { w->reso[k][j] = r_; w->dobs[k][j] = d; }
        bar += log(q->so[k][j] - w->olo);
        lin += q->so[k][j] - w->olo;
      }
  }
  double f = objective(w, q);
  *fobj = f;
  w->pe_thc = th; w->pe_thr = thr; w->pe_bar = bar; w->pe_lin = lin; w->pe_f = f;
  if (w->resto) {
    /* restoration: the rows are penalised, not constrained; objective = proximity term + penalties */
    for (int k = 0; k <= N; k++) {
      for (int i = 0; i < nx; i++) { double e = q->x[k][i] - w->xR[k][i]; fr += 0.5 * w->zeta * dr2(w->xR[k][i]) * e * e; }
      if (k < N) for (int i = 0; i < 2; i++) { double e = q->u[k][i] - w->uR[k][i]; fr += 0.5 * w->zeta * dr2(w->uR[k][i]) * e * e; }
    }
    *theta = th;
    *phi = fr - mu * bar + KAPPA_D * mu * lin;
  } else {
    *theta = th + thr;
    *phi = w->sigma * f - mu * bar + KAPPA_D * mu * lin;
  }
  return isfinite(th + thr) && isfinite(*phi);
}

/* Jacobians of the dynamics and the obstacle rows at the current iterate */
static void eval_lin(ws_t *w) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N;
  iterate_t *q = &w->it;
  for (int k = 0; k <= N; k++) {
    if (k < N) {
      double Jx[NXM][NXM], Ju[NXM][2];
      model_jac(w, q->x[k], q->u[k], Jx, Ju);
      for (int i = 0; i < nx; i++) {
        for (int j = 0; j < nx; j++) w->A[k][i][j] = (i == j ? 1.0 : 0.0) + c->T * Jx[i][j];
        w->B[k][i][0] = c->T * Ju[i][0];
        w->B[k][i][1] = c->T * Ju[i][1];
      }
    }
    if (has_obs(w, k))
      for (int j = 0; j < w->M; j++) {
        if (IS_DCBF(w)) dcbf_row(w, q, k, j, &w->dobs[k][j], 1);
        else
          obs_row(w, k, j, q->x[k][0], q->x[k][1], &w->dobs[k][j], &w->gox[k][j], &w->goy[k][j], &w->hoxx[k][j],
                  &w->hoxy[k][j], &w->hoyy[k][j]);
      }
  }
}

/* gradient of the unscaled objective wrt u_k */
static void grad_u(const ws_t *w, const iterate_t *q, int k, double *g) {
  const orc_cfg *c = w->c;
  for (int i = 0; i < 2; i++) {
    double v = 2 * c->R[i] * q->u[k][i];
    if (k > 0) v += 2 * c->DR[i] * (q->u[k][i] - q->u[k - 1][i]);
    else if (c->du0_cost) v += 2 * c->DR[i] * q->u[0][i];
    if (k + 1 <= w->N - 1) v -= 2 * c->DR[i] * (q->u[k + 1][i] - q->u[k][i]);
    g[i] = v;
  }
}

typedef struct { double err, dual, prim, compl_min, compl_max, sum_lam, sum_z; int n_bm, n_eq; } kkt_t;

/* KKT residual pieces at the current iterate (needs eval_primal(store) + eval_lin) */
static void kkt_pieces(ws_t *w, kkt_t *o) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  iterate_t *q = &w->it;
  double dual = 0, prim = 0, cmin = INFINITY, cmax = -INFINITY, sl = 0, sz = 0;
  int nbm = 0, neq = 0;
#define COMPL(gap, mult) do { double p_ = (gap) * (mult); if (p_ < cmin) cmin = p_; if (p_ > cmax) cmax = p_; sz += (mult); nbm++; } while (0)
  for (int k = 0; k <= N; k++) {
    /* stationarity wrt x_k */
    for (int i = 0; i < nx; i++) {
      double r = q->lam[k][i];
      if (w->resto) r += w->zeta * dr2(w->xR[k][i]) * (q->x[k][i] - w->xR[k][i]);
      if (k < N) {
        if (!w->resto) r += w->sigma * 2 * c->Q[i] * (q->x[k][i] - w->xr[k][i]);
        for (int a = 0; a < nx; a++) r -= w->A[k][a][i] * q->lam[k + 1][a];
      }
      if (w->xbl[i]) { r -= q->zlx[k][i]; COMPL(q->x[k][i] - w->xlo[i], q->zlx[k][i]); }
      if (w->xbu[i]) { r += q->zux[k][i]; COMPL(w->xhi[i] - q->x[k][i], q->zux[k][i]); }
      if (i < 2 && has_obs(w, k))
        for (int j = 0; j < M; j++) r += q->lo[k][j] * (i == 0 ? w->gox[k][j] : w->goy[k][j]);
      if (i < 2 && IS_DCBF(w) && k >= 1) /* row k-1 also depends on x_k */
        for (int j = 0; j < M; j++) r += q->lo[k - 1][j] * (i == 0 ? w->gnx[k - 1][j] : w->gny[k - 1][j]);
      dual = fmax(dual, fabs(r));
      prim = fmax(prim, fabs(w->cdef[k][i]));
      sl += fabs(q->lam[k][i]);
      neq++;
    }
    if (k < N) {
      double g[2];
      grad_u(w, q, k, g);
      for (int i = 0; i < 2; i++) {
        double r = (w->resto ? w->zeta * dr2(w->uR[k][i]) * (q->u[k][i] - w->uR[k][i]) : w->sigma * g[i]) - q->zlu[k][i] + q->zuu[k][i];
        for (int a = 0; a < nx; a++) r -= w->B[k][a][i] * q->lam[k + 1][a];
        for (int rr = 0; rr < c->n_rate; rr++)
          if (c->rate_ctrl[rr] == i) {
            if (has_rate(w, k)) r += q->lr[k][rr];
            if (has_rate(w, k + 1)) r -= q->lr[k + 1][rr];
          }
        dual = fmax(dual, fabs(r));
        COMPL(q->u[k][i] - w->ulo[i], q->zlu[k][i]);
        COMPL(w->uhi[i] - q->u[k][i], q->zuu[k][i]);
      }
    }
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        dual = fmax(dual, fabs(-(SHIPPED(w) ? q->lam[k + 1][r] : q->lr[k][r]) - q->vlr[k][r] + q->vur[k][r]));
        if (!w->resto) prim = fmax(prim, fabs(w->resr[k][r]));
        COMPL(q->sr[k][r] - w->rlo[r], q->vlr[k][r]);
        COMPL(w->rhi[r] - q->sr[k][r], q->vur[k][r]);
        sl += fabs(q->lr[k][r]);
        neq++;
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        dual = fmax(dual, fabs(-q->lo[k][j] - q->vlo[k][j]));
        if (!w->resto) prim = fmax(prim, fabs(w->reso[k][j]));
        COMPL(q->so[k][j] - w->olo, q->vlo[k][j]);
        sl += fabs(q->lo[k][j]);
        neq++;
      }
  }
#undef COMPL
  o->dual = dual; o->prim = prim; o->compl_min = cmin; o->compl_max = cmax;
  o->sum_lam = sl; o->sum_z = sz; o->n_bm = nbm; o->n_eq = neq;
}

static double kkt_error(const kkt_t *o, double mu, double *co_out) {
  double co = fmax(fabs(o->compl_max - mu), fabs(o->compl_min - mu));
  if (o->n_bm == 0) co = 0;
  double s_d = fmax(S_MAX, (o->sum_lam + o->sum_z) / fmax(1, o->n_eq + o->n_bm)) / S_MAX;
  double s_c = fmax(S_MAX, o->sum_z / fmax(1, o->n_bm)) / S_MAX;
  if (co_out) *co_out = co;
  return fmax(fmax(o->dual / s_d, o->prim), co / s_c);
}

/* ---- condensed QP data: everything that does not depend on dw ----------------------- */
static void build_qp(ws_t *w, double mu) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  iterate_t *q = &w->it;
  for (int k = 0; k <= N; k++) {
    double H[8][8];
    memset(H, 0, sizeof H);
    if (k < N) {
      double Hf[8][8];
      model_hess(w, q->x[k], q->u[k], q->lam[k + 1], Hf);
      for (int a = 0; a < nx + 2; a++)
        for (int b = 0; b < nx + 2; b++) {
          int ia = a < nx ? a : 6 + (a - nx), ib = b < nx ? b : 6 + (b - nx); /* model order is [x(nx);u] */
          H[ia][ib] = -c->T * Hf[a][b];
        }
    }
    /* state block */
    for (int i = 0; i < nx; i++) {
      double g = 0;
      if (w->resto) { double d2_ = w->zeta * dr2(w->xR[k][i]); H[i][i] += d2_; g += d2_ * (q->x[k][i] - w->xR[k][i]); }
      else if (k < N) { H[i][i] += w->sigma * 2 * c->Q[i]; g += w->sigma * 2 * c->Q[i] * (q->x[k][i] - w->xr[k][i]); }
      if (w->xbl[i]) { double gap = q->x[k][i] - w->xlo[i]; H[i][i] += q->zlx[k][i] / gap; g -= mu / gap; }
      if (w->xbu[i]) { double gap = w->xhi[i] - q->x[k][i]; H[i][i] += q->zux[k][i] / gap; g += mu / gap; }
      w->gx[k][i] = g;
    }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double gap = q->so[k][j] - w->olo;
        w->Do[k][j] = q->vlo[k][j] / gap;               /* Sigma_s (dw added later) */
        w->gso[k][j] = -mu / gap + KAPPA_D * mu;          /* barrier gradient wrt the slack */
        H[0][0] += q->lo[k][j] * w->hoxx[k][j];
        H[0][1] += q->lo[k][j] * w->hoxy[k][j];
        H[1][0] += q->lo[k][j] * w->hoxy[k][j];
        H[1][1] += q->lo[k][j] * w->hoyy[k][j];
      }
    if (IS_DCBF(w) && k >= 1)
      for (int j = 0; j < M; j++) {
        H[0][0] += q->lo[k - 1][j] * w->hnxx[k - 1][j];
        H[1][1] += q->lo[k - 1][j] * w->hnyy[k - 1][j];
      }
    for (int i = 0; i < nx; i++)
      for (int j = 0; j < nx; j++) w->Hxx[k][i][j] = H[i][j];
    if (k < N) {
      for (int i = 0; i < 2; i++) {
        double g = w->sigma * 2 * c->R[i] * q->u[k][i];
        double hd = w->sigma * 2 * c->R[i];
        if (k == 0 && c->du0_cost) { hd += w->sigma * 2 * c->DR[i]; g += w->sigma * 2 * c->DR[i] * q->u[0][i]; }
        if (w->resto) { hd = w->zeta * dr2(w->uR[k][i]); g = hd * (q->u[k][i] - w->uR[k][i]); }
        double gl = q->u[k][i] - w->ulo[i], gh = w->uhi[i] - q->u[k][i];
        hd += q->zlu[k][i] / gl + q->zuu[k][i] / gh;
        g += -mu / gl + mu / gh;
        H[6 + i][6 + i] += hd;
        w->gu[k][i] = g;
      }
      for (int i = 0; i < 2; i++) {
        for (int j = 0; j < nx; j++) w->Hux[k][i][j] = H[6 + i][j];
        for (int j = 0; j < 2; j++) w->Huu[k][i][j] = H[6 + i][6 + j];
      }
      /* coupling with the previous control: cost DR plus rate rows */
      for (int i = 0; i < 2; i++) { w->E[k][i] = 0; w->tk[k][i] = 0; }
      if (k >= 1 && !w->resto)
        for (int i = 0; i < 2; i++) {
          w->E[k][i] = w->sigma * 2 * c->DR[i];
          w->tk[k][i] = w->sigma * 2 * c->DR[i] * (q->u[k][i] - q->u[k - 1][i]);
        }
      if (has_rate(w, k))
        for (int r = 0; r < c->n_rate; r++) {
          double gl = q->sr[k][r] - w->rlo[r], gh = w->rhi[r] - q->sr[k][r];
          w->Dr[k][r] = q->vlr[k][r] / gl + q->vur[k][r] / gh;
          w->gsr[k][r] = -mu / gl + mu / gh;
        }
    }
  }
}

/* 2x2 symmetric solve helpers */
static int inv2(const double F[2][2], double Fi[2][2]) {
  double det = F[0][0] * F[1][1] - F[0][1] * F[1][0];
  if (!(F[0][0] > 0.0) || !(det > 0.0) || !isfinite(det)) return 0;
  double id = 1.0 / det;
  Fi[0][0] = F[1][1] * id; Fi[1][1] = F[0][0] * id; Fi[0][1] = -F[0][1] * id; Fi[1][0] = -F[1][0] * id;
  return 1;
}

/* Riccati factor+solve of the condensed QP with primal regularisation dw and constraint
 * right-hand side (cdef,resr,reso).  Returns 0 when some Fuu is not positive definite
 * (wrong inertia). */
/* A row `row(z) - s (= 0)` condensed into the stage: new row multiplier lam+ = D * (J dz) + tt.
 * regular phase: D = Sigma_s + dw, tt = D * residual + barrier gradient of the slack;
 * restoration:   the row is penalised by psi: D = 1/(1/Ds + 1/psi''), tt = D (gs/Ds + psi'/psi'') */
static void row_cond(const ws_t *w, double Ds, double gs, double res, double lam, double p2, double *D, double *tt) {
  if (w->resto) { *D = Ds * p2 / (Ds + p2); *tt = *D * (gs / Ds + lam / p2); }
  else { *D = Ds; *tt = Ds * res + gs; }
}
/* slack step and new row multiplier once J dz is known */
static void row_step(const ws_t *w, double Ds, double gs, double res, double lam, double p2, double Jdz, double *ds, double *lnew) {
  if (w->resto) {
    double D, tt;
    row_cond(w, Ds, gs, res, lam, p2, &D, &tt);
    *lnew = D * Jdz + tt;
    *ds = (*lnew - gs) / Ds;
  } else {
    *ds = Jdz + res;
    *lnew = Ds * *ds + gs;
  }
}

static int riccati(ws_t *w, double dw) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  double Pxx[NXM][NXM], Pxw[NXM][2], Pww[2][2], px[NXM], pw[2];
  /* terminal stage */
  {
    int k = N;
    double gxk[NXM];
    for (int i = 0; i < nx; i++) { gxk[i] = w->gx[k][i]; for (int j = 0; j < nx; j++) Pxx[i][j] = w->Hxx[k][i][j]; Pxx[i][i] += dw; }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double D, t, gx_ = w->gox[k][j], gy_ = w->goy[k][j];
        row_cond(w, w->Do[k][j] + dw, w->gso[k][j], w->reso[k][j], w->it.lo[k][j], w->p2o[k][j], &D, &t);
        Pxx[0][0] += D * gx_ * gx_; Pxx[0][1] += D * gx_ * gy_; Pxx[1][0] += D * gx_ * gy_; Pxx[1][1] += D * gy_ * gy_;
        gxk[0] += gx_ * t; gxk[1] += gy_ * t;
      }
    for (int i = 0; i < nx; i++) { px[i] = gxk[i]; Pxw[i][0] = Pxw[i][1] = 0; }
    Pww[0][0] = Pww[0][1] = Pww[1][0] = Pww[1][1] = 0; pw[0] = pw[1] = 0;
  }
  for (int k = N - 1; k >= 0; k--) {
    double (*A)[NXM] = w->A[k];
    double (*B)[2] = w->B[k];
    double Hxx[NXM][NXM], gxk[NXM], guk[2], E[2], t[2];
    double dHux[2][NXM] = {{0}}, dHuu[2][2] = {{0}}, dgu[2] = {0, 0}; /* DCBF rows reach u_k through x_{k+1} */
    for (int i = 0; i < nx; i++) { gxk[i] = w->gx[k][i]; for (int j = 0; j < nx; j++) Hxx[i][j] = w->Hxx[k][i][j]; Hxx[i][i] += dw; }
    if (has_obs(w, k) && !IS_DCBF(w))
      for (int j = 0; j < M; j++) {
        double D, tt, gx_ = w->gox[k][j], gy_ = w->goy[k][j];
        row_cond(w, w->Do[k][j] + dw, w->gso[k][j], w->reso[k][j], w->it.lo[k][j], w->p2o[k][j], &D, &tt);
        Hxx[0][0] += D * gx_ * gx_; Hxx[0][1] += D * gx_ * gy_; Hxx[1][0] += D * gx_ * gy_; Hxx[1][1] += D * gy_ * gy_;
        gxk[0] += gx_ * tt; gxk[1] += gy_ * tt;
      }
    if (has_obs(w, k) && IS_DCBF(w))
      for (int j = 0; j < M; j++) {
        /* row k is linear in (dx_k, dx_{k+1}); with dx_{k+1} = A dx_k + B du_k + b it becomes a row
         * in (dx_k, du_k): v = [g_k + A' g_next ; B' g_next], residual + g_next' b */
        double D, tt;
        row_cond(w, w->Do[k][j] + dw, w->gso[k][j], w->reso[k][j], w->it.lo[k][j], w->p2o[k][j], &D, &tt);
        double gn[2] = {w->gnx[k][j], w->gny[k][j]};
        double v[NXM + 2];
        for (int i = 0; i < nx; i++) v[i] = (i == 0 ? w->gox[k][j] : i == 1 ? w->goy[k][j] : 0.0) + gn[0] * A[0][i] + gn[1] * A[1][i];
        for (int i = 0; i < 2; i++) v[nx + i] = gn[0] * B[0][i] + gn[1] * B[1][i];
        tt -= D * (gn[0] * w->cdef[k + 1][0] + gn[1] * w->cdef[k + 1][1]);  /* J dz = v'[dx_k; du_k] + g_next' b, b = -c_{k+1} */
        for (int i = 0; i < nx; i++) {
          for (int l = 0; l < nx; l++) Hxx[i][l] += D * v[i] * v[l];
          gxk[i] += v[i] * tt;
        }
        for (int i = 0; i < 2; i++) {
          for (int l = 0; l < nx; l++) dHux[i][l] += D * v[nx + i] * v[l];
          for (int l = 0; l < 2; l++) dHuu[i][l] += D * v[nx + i] * v[nx + l];
          dgu[i] += v[nx + i] * tt;
        }
      }
    for (int i = 0; i < 2; i++) { E[i] = w->E[k][i]; t[i] = w->tk[k][i]; guk[i] = w->gu[k][i]; }
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        int ci = c->rate_ctrl[r];
        double D, tt;
        row_cond(w, w->Dr[k][r] + dw, w->gsr[k][r], w->resr[k][r], w->it.lr[k][r], w->p2r[k][r], &D, &tt);
        E[ci] += D;
        t[ci] += tt;
      }
    /* b = -c_{k+1} */
    double b[NXM], Pb[NXM];
    for (int i = 0; i < nx; i++) b[i] = -w->cdef[k + 1][i];
    for (int i = 0; i < nx; i++) { double s = px[i]; for (int j = 0; j < nx; j++) s += Pxx[i][j] * b[j]; Pb[i] = s; }
    /* PA = Pxx A, PB = Pxx B */
    double PA[NXM][NXM], PB[NXM][2];
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) { double s = 0; for (int a = 0; a < nx; a++) s += Pxx[i][a] * A[a][j]; PA[i][j] = s; }
      for (int j = 0; j < 2; j++) { double s = 0; for (int a = 0; a < nx; a++) s += Pxx[i][a] * B[a][j]; PB[i][j] = s; }
    }
    double Fxx[NXM][NXM], Fux[2][NXM], Fuu[2][2], fx[NXM], fu[2];
    for (int i = 0; i < nx; i++)
      for (int j = 0; j < nx; j++) { double s = Hxx[i][j]; for (int a = 0; a < nx; a++) s += A[a][i] * PA[a][j]; Fxx[i][j] = s; }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < nx; j++) {
        double s = w->Hux[k][i][j] + dHux[i][j];
        for (int a = 0; a < nx; a++) s += B[a][i] * PA[a][j] + Pxw[a][i] * A[a][j];
        Fux[i][j] = s;
      }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        double s = w->Huu[k][i][j] + dHuu[i][j] + Pww[i][j];
        if (i == j) s += dw + E[i];
        for (int a = 0; a < nx; a++) s += B[a][i] * PB[a][j] + B[a][i] * Pxw[a][j] + Pxw[a][i] * B[a][j];
        Fuu[i][j] = s;
      }
    for (int i = 0; i < nx; i++) { double s = gxk[i]; for (int a = 0; a < nx; a++) s += A[a][i] * Pb[a]; fx[i] = s; }
    for (int i = 0; i < 2; i++) {
      double s = guk[i] + dgu[i] + t[i] + pw[i];
      for (int a = 0; a < nx; a++) s += B[a][i] * Pb[a] + Pxw[a][i] * b[a];
      fu[i] = s;
    }
    double Fi[2][2];
    if (!inv2(Fuu, Fi)) return 0;
    /* gains */
    for (int i = 0; i < 2; i++) {
      for (int j = 0; j < nx; j++) w->Kx[k][i][j] = -(Fi[i][0] * Fux[0][j] + Fi[i][1] * Fux[1][j]);
      for (int j = 0; j < 2; j++) w->Kw[k][i][j] = Fi[i][j] * E[j]; /* -Fi * Fuw, Fuw = -diag(E) */
      w->kk[k][i] = -(Fi[i][0] * fu[0] + Fi[i][1] * fu[1]);
    }
    /* value function of stage k */
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) Pxx[i][j] = Fxx[i][j] + Fux[0][i] * w->Kx[k][0][j] + Fux[1][i] * w->Kx[k][1][j];
      for (int j = 0; j < 2; j++) Pxw[i][j] = Fux[0][i] * w->Kw[k][0][j] + Fux[1][i] * w->Kw[k][1][j];
      px[i] = fx[i] + Fux[0][i] * w->kk[k][0] + Fux[1][i] * w->kk[k][1];
    }
    for (int i = 0; i < nx; i++)
      for (int j = i + 1; j < nx; j++) { double m = 0.5 * (Pxx[i][j] + Pxx[j][i]); Pxx[i][j] = Pxx[j][i] = m; }
    for (int i = 0; i < 2; i++) {
      for (int j = 0; j < 2; j++) Pww[i][j] = (i == j ? E[i] : 0.0) - E[i] * w->Kw[k][i][j];
      pw[i] = -t[i] - E[i] * w->kk[k][i];
    }
    { double m = 0.5 * (Pww[0][1] + Pww[1][0]); Pww[0][1] = Pww[1][0] = m; }
  }
  /* forward */
  for (int i = 0; i < nx; i++) w->dx[0][i] = -w->cdef[0][i];
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < 2; i++) {
      double s = w->kk[k][i];
      for (int j = 0; j < nx; j++) s += w->Kx[k][i][j] * w->dx[k][j];
      if (k > 0) s += w->Kw[k][i][0] * w->du[k - 1][0] + w->Kw[k][i][1] * w->du[k - 1][1];
      w->du[k][i] = s;
    }
    for (int i = 0; i < nx; i++) {
      double s = -w->cdef[k + 1][i];
      for (int j = 0; j < nx; j++) s += w->A[k][i][j] * w->dx[k][j];
      s += w->B[k][i][0] * w->du[k][0] + w->B[k][i][1] * w->du[k][1];
      w->dx[k + 1][i] = s;
    }
  }
  /* slack steps and new row multipliers */
  for (int k = 0; k <= N; k++) {
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        int ci = c->rate_ctrl[r];
        row_step(w, w->Dr[k][r] + dw, w->gsr[k][r], w->resr[k][r], w->it.lr[k][r], w->p2r[k][r], w->du[k][ci] - w->du[k - 1][ci],
                 &w->dsr[k][r], &w->lrp[k][r]);
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double Jdz = w->gox[k][j] * w->dx[k][0] + w->goy[k][j] * w->dx[k][1];
        if (IS_DCBF(w)) Jdz += w->gnx[k][j] * w->dx[k + 1][0] + w->gny[k][j] * w->dx[k + 1][1];
        row_step(w, w->Do[k][j] + dw, w->gso[k][j], w->reso[k][j], w->it.lo[k][j], w->p2o[k][j], Jdz, &w->dso[k][j], &w->lop[k][j]);
      }
  }
  /* adjoint recursion for the new dynamics multipliers */
  for (int k = N; k >= 0; k--) {
    for (int i = 0; i < nx; i++) {
      double s = w->gx[k][i] + dw * w->dx[k][i];
      for (int j = 0; j < nx; j++) s += w->Hxx[k][i][j] * w->dx[k][j];
      if (k < N) s += w->Hux[k][0][i] * w->du[k][0] + w->Hux[k][1][i] * w->du[k][1];
      if (i < 2 && has_obs(w, k))
        for (int j = 0; j < M; j++) s += (i == 0 ? w->gox[k][j] : w->goy[k][j]) * w->lop[k][j];
      if (i < 2 && IS_DCBF(w) && k >= 1)
        for (int j = 0; j < M; j++) s += (i == 0 ? w->gnx[k - 1][j] : w->gny[k - 1][j]) * w->lop[k - 1][j];
      s = -s;
      if (k < N) for (int a = 0; a < nx; a++) s += w->A[k][a][i] * w->lamp[k + 1][a];
      w->lamp[k][i] = s;
    }
  }
  return 1;
}

/* Newton system of the dyn problem with the bound lists as shipped.  Stage 0 has the free control U_0;
 * at stages k >= 1 the control is tied to the previous one (dU_k = dU_{k-1} - e_k, e_k = U_k - U_{k-1}) and the
 * free inputs are the two slacks that relax the x/y defects, entering x_{k+1} additively.  The value function
 * keeps the same (x, w = previous control) form as in riccati(). */
static int riccati_shipped(ws_t *w, double dw) {
  const int nx = w->nx, N = w->N, M = w->M;
  double Pxx[NXM][NXM], Pxw[NXM][2], Pww[2][2], px[NXM], pw[2];
  {
    int k = N;
    double gxk[NXM];
    for (int i = 0; i < nx; i++) { gxk[i] = w->gx[k][i]; for (int j = 0; j < nx; j++) Pxx[i][j] = w->Hxx[k][i][j]; Pxx[i][i] += dw; }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double D = w->Do[k][j] + dw, gx_ = w->gox[k][j], gy_ = w->goy[k][j];
        double t = D * w->reso[k][j] + w->gso[k][j];
        Pxx[0][0] += D * gx_ * gx_; Pxx[0][1] += D * gx_ * gy_; Pxx[1][0] += D * gx_ * gy_; Pxx[1][1] += D * gy_ * gy_;
        gxk[0] += gx_ * t; gxk[1] += gy_ * t;
      }
    for (int i = 0; i < nx; i++) { px[i] = gxk[i]; Pxw[i][0] = Pxw[i][1] = 0; }
    Pww[0][0] = Pww[0][1] = Pww[1][0] = Pww[1][1] = 0; pw[0] = pw[1] = 0;
  }
  for (int k = N - 1; k >= 0; k--) {
    double (*A)[NXM] = w->A[k];
    double (*B)[2] = w->B[k];
    double Hxx[NXM][NXM], gxk[NXM];
    for (int i = 0; i < nx; i++) { gxk[i] = w->gx[k][i]; for (int j = 0; j < nx; j++) Hxx[i][j] = w->Hxx[k][i][j]; Hxx[i][i] += dw; }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double D = w->Do[k][j] + dw, gx_ = w->gox[k][j], gy_ = w->goy[k][j];
        double tt = D * w->reso[k][j] + w->gso[k][j];
        Hxx[0][0] += D * gx_ * gx_; Hxx[0][1] += D * gx_ * gy_; Hxx[1][0] += D * gx_ * gy_; Hxx[1][1] += D * gy_ * gy_;
        gxk[0] += gx_ * tt; gxk[1] += gy_ * tt;
      }
    double b[NXM], Pb[NXM];
    for (int i = 0; i < nx; i++) b[i] = -w->cdef[k + 1][i];
    for (int i = 0; i < nx; i++) { double s = px[i]; for (int j = 0; j < nx; j++) s += Pxx[i][j] * b[j]; Pb[i] = s; }
    double PA[NXM][NXM], PB[NXM][2];
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) { double s = 0; for (int a = 0; a < nx; a++) s += Pxx[i][a] * A[a][j]; PA[i][j] = s; }
      for (int j = 0; j < 2; j++) { double s = 0; for (int a = 0; a < nx; a++) s += Pxx[i][a] * B[a][j]; PB[i][j] = s; }
    }
    /* quadratic over (dx_k, dU_k) of the stage cost plus V_{k+1}(A dx + B dU + b, w+ = dU); the DR cost
     * couples dU_k with dU_{k-1} only through dU_k - dU_{k-1} = -e_k, a constant on this manifold */
    double Fxx[NXM][NXM], Fux[2][NXM], Fuu[2][2], fx[NXM], fu[2];
    for (int i = 0; i < nx; i++)
      for (int j = 0; j < nx; j++) { double s = Hxx[i][j]; for (int a = 0; a < nx; a++) s += A[a][i] * PA[a][j]; Fxx[i][j] = s; }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < nx; j++) {
        double s = w->Hux[k][i][j];
        for (int a = 0; a < nx; a++) s += B[a][i] * PA[a][j] + Pxw[a][i] * A[a][j];
        Fux[i][j] = s;
      }
    for (int i = 0; i < 2; i++)
      for (int j = 0; j < 2; j++) {
        double s = w->Huu[k][i][j] + Pww[i][j];
        if (i == j) s += dw;
        for (int a = 0; a < nx; a++) s += B[a][i] * PB[a][j] + B[a][i] * Pxw[a][j] + Pxw[a][i] * B[a][j];
        Fuu[i][j] = s;
      }
    for (int i = 0; i < nx; i++) { double s = gxk[i]; for (int a = 0; a < nx; a++) s += A[a][i] * Pb[a]; fx[i] = s; }
    for (int i = 0; i < 2; i++) {
      double s = w->gu[k][i] + pw[i];
      for (int a = 0; a < nx; a++) s += B[a][i] * Pb[a] + Pxw[a][i] * b[a];
      fu[i] = s;
    }
    if (k == 0) {
      double Fi[2][2];
      if (!inv2(Fuu, Fi)) return 0;
      for (int i = 0; i < 2; i++) {
        for (int j = 0; j < nx; j++) w->Kx[k][i][j] = -(Fi[i][0] * Fux[0][j] + Fi[i][1] * Fux[1][j]);
        w->Kw[k][i][0] = w->Kw[k][i][1] = 0.0;
        w->kk[k][i] = -(Fi[i][0] * fu[0] + Fi[i][1] * fu[1]);
      }
      break;
    }
    /* k >= 1: inputs are the two slacks; they reach x_{k+1} rows 0,1 */
    const double *e = w->resr[k];
    double Hss[2][2], Hi[2][2], Cm[2][2], c0[2];
    for (int a = 0; a < 2; a++) {
      for (int j = 0; j < 2; j++) { Cm[a][j] = PB[a][j] + Pxw[a][j]; Hss[a][j] = Pxx[a][j]; }
      Hss[a][a] += w->Dr[k][a] + dw;
    }
    for (int a = 0; a < 2; a++) c0[a] = Pb[a] + w->gsr[k][a] - (Cm[a][0] * e[0] + Cm[a][1] * e[1]);
    if (!inv2(Hss, Hi)) return 0;
    for (int a = 0; a < 2; a++) {
      for (int j = 0; j < nx; j++) w->Kx[k][a][j] = -(Hi[a][0] * PA[0][j] + Hi[a][1] * PA[1][j]);
      for (int j = 0; j < 2; j++) w->Kw[k][a][j] = -(Hi[a][0] * Cm[0][j] + Hi[a][1] * Cm[1][j]);
      w->kk[k][a] = -(Hi[a][0] * c0[0] + Hi[a][1] * c0[1]);
    }
    double nPxx[NXM][NXM], nPxw[NXM][2], nPww[2][2], npx[NXM], npw[2];
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) nPxx[i][j] = Fxx[i][j] + PA[0][i] * w->Kx[k][0][j] + PA[1][i] * w->Kx[k][1][j];
      for (int j = 0; j < 2; j++) nPxw[i][j] = Fux[j][i] + PA[0][i] * w->Kw[k][0][j] + PA[1][i] * w->Kw[k][1][j];
      npx[i] = fx[i] - (Fux[0][i] * e[0] + Fux[1][i] * e[1]) + PA[0][i] * w->kk[k][0] + PA[1][i] * w->kk[k][1];
    }
    for (int i = 0; i < 2; i++) {
      for (int j = 0; j < 2; j++) nPww[i][j] = Fuu[i][j] + Cm[0][i] * w->Kw[k][0][j] + Cm[1][i] * w->Kw[k][1][j];
      npw[i] = fu[i] - (Fuu[i][0] * e[0] + Fuu[i][1] * e[1]) + Cm[0][i] * w->kk[k][0] + Cm[1][i] * w->kk[k][1];
    }
    for (int i = 0; i < nx; i++) {
      for (int j = 0; j < nx; j++) Pxx[i][j] = 0.5 * (nPxx[i][j] + nPxx[j][i]);
      Pxw[i][0] = nPxw[i][0]; Pxw[i][1] = nPxw[i][1];
      px[i] = npx[i];
    }
    Pww[0][0] = nPww[0][0]; Pww[1][1] = nPww[1][1]; Pww[0][1] = Pww[1][0] = 0.5 * (nPww[0][1] + nPww[1][0]);
    pw[0] = npw[0]; pw[1] = npw[1];
  }
  /* forward */
  for (int i = 0; i < nx; i++) w->dx[0][i] = -w->cdef[0][i];
  for (int k = 0; k < N; k++) {
    double ds[2] = {0, 0};
    if (k == 0) {
      for (int i = 0; i < 2; i++) {
        double s = w->kk[0][i];
        for (int j = 0; j < nx; j++) s += w->Kx[0][i][j] * w->dx[0][j];
        w->du[0][i] = s;
      }
    } else {
      for (int a = 0; a < 2; a++) {
        double s = w->kk[k][a];
        for (int j = 0; j < nx; j++) s += w->Kx[k][a][j] * w->dx[k][j];
        s += w->Kw[k][a][0] * w->du[k - 1][0] + w->Kw[k][a][1] * w->du[k - 1][1];
        ds[a] = s;
        w->dsr[k][a] = s;
      }
      for (int i = 0; i < 2; i++) w->du[k][i] = w->du[k - 1][i] - w->resr[k][i];
    }
    for (int i = 0; i < nx; i++) {
      double s = -w->cdef[k + 1][i];
      for (int j = 0; j < nx; j++) s += w->A[k][i][j] * w->dx[k][j];
      s += w->B[k][i][0] * w->du[k][0] + w->B[k][i][1] * w->du[k][1];
      if (i < 2) s += ds[i];
      w->dx[k + 1][i] = s;
    }
  }
  for (int k = 0; k <= N; k++)
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double D = w->Do[k][j] + dw;
        w->dso[k][j] = w->gox[k][j] * w->dx[k][0] + w->goy[k][j] * w->dx[k][1] + w->reso[k][j];
        w->lop[k][j] = D * w->dso[k][j] + w->gso[k][j];
      }
  /* adjoint recursion: multipliers of every defect row (relaxed ones included) */
  for (int k = N; k >= 0; k--)
    for (int i = 0; i < nx; i++) {
      double s = w->gx[k][i] + dw * w->dx[k][i];
      for (int j = 0; j < nx; j++) s += w->Hxx[k][i][j] * w->dx[k][j];
      if (k < N) s += w->Hux[k][0][i] * w->du[k][0] + w->Hux[k][1][i] * w->du[k][1];
      if (i < 2 && has_obs(w, k))
        for (int j = 0; j < M; j++) s += (i == 0 ? w->gox[k][j] : w->goy[k][j]) * w->lop[k][j];
      s = -s;
      if (k < N) for (int a = 0; a < nx; a++) s += w->A[k][a][i] * w->lamp[k + 1][a];
      w->lamp[k][i] = s;
    }
  /* multipliers of the rate equalities from stationarity in U_k, k = N-1..1 */
  double nu[2] = {0, 0};
  for (int k = N - 1; k >= 1; k--)
    for (int i = 0; i < 2; i++) {
      double s = w->gu[k][i] + dw * w->du[k][i] + w->Huu[k][i][0] * w->du[k][0] + w->Huu[k][i][1] * w->du[k][1];
      for (int j = 0; j < nx; j++) s += w->Hux[k][i][j] * w->dx[k][j];
      for (int a = 0; a < nx; a++) s -= w->B[k][a][i] * w->lamp[k + 1][a];
      s += w->E[k][i] * (w->du[k][i] - w->du[k - 1][i]) + w->tk[k][i];
      if (k + 1 <= N - 1) s -= w->E[k + 1][i] * (w->du[k + 1][i] - w->du[k][i]) + w->tk[k + 1][i];
      nu[i] = nu[i] - s;
      w->lrp[k][i] = nu[i];
    }
  return 1;
}

/* bound-multiplier steps; returns primal and dual fraction-to-boundary step sizes */
static void mult_steps(ws_t *w, double mu, double tau, double *a_pr, double *a_du) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  iterate_t *q = &w->it;
  double ap = 1.0, ad = 1.0;
#define LOWER(v, dv, lo, z, dzout) do { double gap = (v) - (lo); double dz_ = -(z) + (mu - (z) * (dv)) / gap; dzout = dz_; \
    if ((dv) < 0) ap = fmin(ap, -tau * gap / (dv)); if (dz_ < 0) ad = fmin(ad, -tau * (z) / dz_); } while (0)
#define UPPER(v, dv, hi, z, dzout) do { double gap = (hi) - (v); double dz_ = -(z) + (mu + (z) * (dv)) / gap; dzout = dz_; \
    if ((dv) > 0) ap = fmin(ap, tau * gap / (dv)); if (dz_ < 0) ad = fmin(ad, -tau * (z) / dz_); } while (0)
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < nx; i++) {
      if (w->xbl[i]) LOWER(q->x[k][i], w->dx[k][i], w->xlo[i], q->zlx[k][i], w->dzlx[k][i]);
      if (w->xbu[i]) UPPER(q->x[k][i], w->dx[k][i], w->xhi[i], q->zux[k][i], w->dzux[k][i]);
    }
    if (k < N)
      for (int i = 0; i < 2; i++) {
        LOWER(q->u[k][i], w->du[k][i], w->ulo[i], q->zlu[k][i], w->dzlu[k][i]);
        UPPER(q->u[k][i], w->du[k][i], w->uhi[i], q->zuu[k][i], w->dzuu[k][i]);
      }
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        LOWER(q->sr[k][r], w->dsr[k][r], w->rlo[r], q->vlr[k][r], w->dvlr[k][r]);
        UPPER(q->sr[k][r], w->dsr[k][r], w->rhi[r], q->vur[k][r], w->dvur[k][r]);
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) LOWER(q->so[k][j], w->dso[k][j], w->olo, q->vlo[k][j], w->dvlo[k][j]);
  }
#undef LOWER
#undef UPPER
  *a_pr = ap; *a_du = ad;
}

/* directional derivative of the barrier function along the primal step */
static double barrier_dir(ws_t *w, double mu) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  iterate_t *q = &w->it;
  double gd = 0;
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < nx; i++) {
      double g = 0;
      if (w->resto) g += w->zeta * dr2(w->xR[k][i]) * (q->x[k][i] - w->xR[k][i]);
      else if (k < N) g += w->sigma * 2 * c->Q[i] * (q->x[k][i] - w->xr[k][i]);
      if (w->xbl[i]) g -= mu / (q->x[k][i] - w->xlo[i]);
      if (w->xbu[i]) g += mu / (w->xhi[i] - q->x[k][i]);
      gd += g * w->dx[k][i];
    }
    if (k < N) {
      double g[2];
      grad_u(w, q, k, g);
      for (int i = 0; i < 2; i++)
        gd += ((w->resto ? w->zeta * dr2(w->uR[k][i]) * (q->u[k][i] - w->uR[k][i]) : w->sigma * g[i]) - mu / (q->u[k][i] - w->ulo[i]) +
               mu / (w->uhi[i] - q->u[k][i])) * w->du[k][i];
    }
    /* restoration: + psi'(r) * d(row - s) for every penalised row */
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        gd += w->gsr[k][r] * w->dsr[k][r];
        if (w->resto) { int ci = c->rate_ctrl[r]; gd += q->lr[k][r] * (w->du[k][ci] - w->du[k - 1][ci] - w->dsr[k][r]); }
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        gd += w->gso[k][j] * w->dso[k][j];
        if (w->resto) {
          double Jdz = w->gox[k][j] * w->dx[k][0] + w->goy[k][j] * w->dx[k][1];
          if (IS_DCBF(w)) Jdz += w->gnx[k][j] * w->dx[k + 1][0] + w->gny[k][j] * w->dx[k + 1][1];
          gd += q->lo[k][j] * (Jdz - w->dso[k][j]);
        }
      }
  }
  return gd;
}

static void make_trial(ws_t *w, double a) {
  int nx = w->nx, N = w->N, M = w->M;
  const orc_cfg *c = w->c;
  iterate_t *q = &w->it, *t = &w->tr;
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < nx; i++) t->x[k][i] = q->x[k][i] + a * w->dx[k][i];
    if (k < N) for (int i = 0; i < 2; i++) t->u[k][i] = q->u[k][i] + a * w->du[k][i];
    if (has_rate(w, k)) for (int r = 0; r < c->n_rate; r++) t->sr[k][r] = q->sr[k][r] + a * w->dsr[k][r];
    if (has_obs(w, k)) for (int j = 0; j < M; j++) t->so[k][j] = q->so[k][j] + a * w->dso[k][j];
  }
}

static double clampz(double z, double mu, double gap) {
  return fmax(fmin(z, KAPPA_SIGMA * mu / gap), mu / (KAPPA_SIGMA * gap));
}

static void accept_step(ws_t *w, double a, double ad, double mu) {
  int nx = w->nx, N = w->N, M = w->M;
  const orc_cfg *c = w->c;
  iterate_t *q = &w->it, *t = &w->tr;
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < nx; i++) {
      q->x[k][i] = t->x[k][i];
      q->lam[k][i] += a * (w->lamp[k][i] - q->lam[k][i]);
      if (w->xbl[i]) q->zlx[k][i] = clampz(q->zlx[k][i] + ad * w->dzlx[k][i], mu, q->x[k][i] - w->xlo[i]);
      if (w->xbu[i]) q->zux[k][i] = clampz(q->zux[k][i] + ad * w->dzux[k][i], mu, w->xhi[i] - q->x[k][i]);
    }
    if (k < N)
      for (int i = 0; i < 2; i++) {
        q->u[k][i] = t->u[k][i];
        q->zlu[k][i] = clampz(q->zlu[k][i] + ad * w->dzlu[k][i], mu, q->u[k][i] - w->ulo[i]);
        q->zuu[k][i] = clampz(q->zuu[k][i] + ad * w->dzuu[k][i], mu, w->uhi[i] - q->u[k][i]);
      }
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        q->sr[k][r] = t->sr[k][r];
        q->lr[k][r] += a * (w->lrp[k][r] - q->lr[k][r]);
        q->vlr[k][r] = clampz(q->vlr[k][r] + ad * w->dvlr[k][r], mu, q->sr[k][r] - w->rlo[r]);
        q->vur[k][r] = clampz(q->vur[k][r] + ad * w->dvur[k][r], mu, w->rhi[r] - q->sr[k][r]);
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        q->so[k][j] = t->so[k][j];
        q->lo[k][j] += a * (w->lop[k][j] - q->lo[k][j]);
        q->vlo[k][j] = clampz(q->vlo[k][j] + ad * w->dvlo[k][j], mu, q->so[k][j] - w->olo);
      }
  }
}

/* ---- problem setup ------------------------------------------------------------------- */
static void setup(ws_t *w, const orc_cfg *c, const double *x0, const double *xs, const double *obs) {
  memset(w, 0, sizeof *w);
  w->c = c;
  w->nx = orc_nx(c);
  w->N = c->N;
  w->M = c->obs_mode == ORC_OBS_NONE ? 0 : c->M;
  for (int i = 0; i < w->nx; i++) {
    w->x0[i] = x0[i];
    for (int k = 0; k <= c->N; k++) /* row N is never read (the cost runs over stages 0..N-1) */
      w->xr[k][i] = c->ref_mode == ORC_REF_TRAJECTORY ? xs[(size_t)(k < c->N ? k : c->N - 1) * w->nx + i] : xs[i];
    w->xlo[i] = relax_lo(c->x_lo[i], c->bound_relax);
    w->xhi[i] = relax_hi(c->x_hi[i], c->bound_relax);
    w->xbl[i] = isfinite(c->x_lo[i]);
    w->xbu[i] = isfinite(c->x_hi[i]);
  }
  for (int i = 0; i < 2; i++) {
    w->ulo[i] = relax_lo(c->u_lo[i], c->bound_relax);
    w->uhi[i] = relax_hi(c->u_hi[i], c->bound_relax);
    w->rlo[i] = relax_lo(c->rate_lo[i], c->bound_relax);
    w->rhi[i] = relax_hi(c->rate_hi[i], c->bound_relax);
  }
  w->olo = relax_lo(c->obs_lo, c->bound_relax);
  for (int j = 0; j < w->M; j++)
    for (int k = 0; k <= w->N; k++) {
      const double *o = obs + ((size_t)j * (w->N + 1) + k) * 6;
      w->ocx[k][j] = o[0];
      w->ocy[k][j] = o[1];
      double sx, sy;
      if (c->obs_mode != ORC_OBS_SQRT) { /* PKG/MPC_CBF_optimize_kin_pre.py:246-249 */
        sx = c->ego_hl + o[4] / 2 + c->safe_l;
        sy = c->ego_hw + o[5] / 2 + c->safe_w;
      } else {
        sx = c->dyn_sx;
        sy = c->dyn_sy;
      }
      w->isx2[k][j] = 1.0 / (sx * sx);
      w->isy2[k][j] = 1.0 / (sy * sy);
    }
  double p[8] = {c->Veh_lf, c->Veh_lr, c->Veh_m, c->Veh_Iz, c->aopt_f, c->aopt_r, c->Fymax_f, c->Fymax_r};
  memcpy(w->pdyn, p, sizeof p);
}

/* start point: push into bounds, slacks, multipliers (IPOPT defaults), objective scaling */
static int init_iterate(ws_t *w, const double *z_init) {
  const orc_cfg *c = w->c;
  int nx = w->nx, N = w->N, M = w->M;
  iterate_t *q = &w->it;
  for (int k = 0; k < N; k++)
    for (int i = 0; i < 2; i++) q->u[k][i] = push_in(z_init ? z_init[2 * k + i] : 0.0, w->ulo[i], w->uhi[i]);
  if (c->init_mode == ORC_INIT_ROLLOUT) {
    for (int i = 0; i < nx; i++) q->x[0][i] = w->x0[i];
    for (int k = 0; k < N; k++) {
      double f[NXM];
      model_f(w, q->x[k], q->u[k], f);
      for (int i = 0; i < nx; i++) q->x[k + 1][i] = q->x[k][i] + c->T * f[i];
    }
  } else {
    for (int k = 0; k <= N; k++)
      for (int i = 0; i < nx; i++) q->x[k][i] = z_init ? z_init[2 * N + nx * k + i] : 0.0;
  }
  for (int k = 0; k <= N; k++) {
    for (int i = 0; i < nx; i++) {
      q->x[k][i] = push_in(q->x[k][i], w->xlo[i], w->xhi[i]);
      q->zlx[k][i] = w->xbl[i] ? 1.0 : 0.0;
      q->zux[k][i] = w->xbu[i] ? 1.0 : 0.0;
    }
    if (k < N) for (int i = 0; i < 2; i++) { q->zlu[k][i] = 1.0; q->zuu[k][i] = 1.0; }
  }
  for (int k = 0; k <= N; k++) {
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) {
        int ci = c->rate_ctrl[r];
        double v0 = q->u[k][ci] - q->u[k - 1][ci];
        if (SHIPPED(w)) { /* the slack of the relaxed defect component r */
          double f[NXM];
          model_f(w, q->x[k], q->u[k], f);
          v0 = q->x[k + 1][r] - (q->x[k][r] + c->T * f[r]);
        }
        q->sr[k][r] = push_in(v0, w->rlo[r], w->rhi[r]);
        q->vlr[k][r] = q->vur[k][r] = 1.0;
      }
    if (has_obs(w, k))
      for (int j = 0; j < M; j++) {
        double d;
        if (IS_DCBF(w)) { if (!dcbf_row(w, q, k, j, &d, 0)) return 0; }
        else if (!obs_row(w, k, j, q->x[k][0], q->x[k][1], &d, 0, 0, 0, 0, 0)) return 0;
        q->so[k][j] = push_in(d, w->olo, INFINITY);
        q->vlo[k][j] = 1.0;
      }
  }
  /* gradient-based objective scaling (IPOPT nlp_scaling_max_gradient = 100) */
  double gmax = 0;
  for (int k = 0; k < N; k++) {
    double g[2];
    grad_u(w, q, k, g);
    gmax = fmax(gmax, fmax(fabs(g[0]), fabs(g[1])));
    for (int i = 0; i < nx; i++) gmax = fmax(gmax, fabs(2 * c->Q[i] * (q->x[k][i] - w->xr[k][i])));
  }
  w->sigma = gmax > OBJ_SCALE_MAX_GRAD ? OBJ_SCALE_MAX_GRAD / gmax : 1.0;
  if (w->sigma < 1e-8) w->sigma = 1e-8;
  return 1;
}

static int in_filter(const ws_t *w, double th, double ph, double theta_max) {
  if (th >= theta_max) return 1;
  for (int i = 0; i < w->nfilt; i++)
    if (th >= w->filt_t[i] && ph >= w->filt_p[i]) return 1;
  return 0;
}

static void filter_add(ws_t *w, double th, double ph) {
  /* drop entries dominated by the new one */
  int n = 0;
  for (int i = 0; i < w->nfilt; i++)
    if (!(w->filt_t[i] >= th && w->filt_p[i] >= ph)) { w->filt_t[n] = w->filt_t[i]; w->filt_p[n] = w->filt_p[i]; n++; }
  if (n == FILTER_CAP) { /* overwrite the oldest */
    for (int i = 1; i < n; i++) { w->filt_t[i - 1] = w->filt_t[i]; w->filt_p[i - 1] = w->filt_p[i]; }
    n--;
  }
  w->filt_t[n] = th; w->filt_p[n] = ph;
  w->nfilt = n + 1;
}

static void write_out(const ws_t *w, double *z_out, double *lam_eq_out) {
  int nx = w->nx, N = w->N;
  if (z_out) {
    for (int k = 0; k < N; k++) { z_out[2 * k] = w->it.u[k][0]; z_out[2 * k + 1] = w->it.u[k][1]; }
    for (int k = 0; k <= N; k++) for (int i = 0; i < nx; i++) z_out[2 * N + nx * k + i] = w->it.x[k][i];
  }
  if (lam_eq_out)
    for (int k = 0; k <= N; k++) for (int i = 0; i < nx; i++) lam_eq_out[nx * k + i] = w->it.lam[k][i] / w->sigma;
}

/* restoration: the row multipliers are not iterates but psi'(residual); psi'' goes into the condensation */
static void resto_multipliers(ws_t *w, double mu) {
  const orc_cfg *c = w->c;
  for (int k = 0; k <= w->N; k++) {
    if (has_rate(w, k))
      for (int r = 0; r < c->n_rate; r++) psi_eval(w->resr[k][r], mu, 0, &w->it.lr[k][r], &w->p2r[k][r]);
    if (has_obs(w, k))
      for (int j = 0; j < w->M; j++) psi_eval(w->reso[k][j], mu, 0, &w->it.lo[k][j], &w->p2o[k][j]);
  }
}

static int solve_ws(ws_t *w, const double *z_init, double *z_out, double *lam_eq_out, orc_info *info) {
  const orc_cfg *c = w->c;
  double mu = c->mu_init, tau = fmax(TAU_MIN, 1 - mu);
  double tol = c->tol;
  int status = ORC_MAXITER, it = 0, n_reg = 0, n_bt = 0;
  /* restoration phase state (see ipm_dense.solve) */
  int n_resto = 0, n_resto_it = 0, it_resto0 = 0;
  double mu_o = 0, tau_o = 0, th_R = 0, theta_min_o = 0, theta_max_o = 0;
  const int nx = w->nx, N = w->N, M = w->M;
  double err0 = INFINITY, dw_last = 0.0;
  if (!init_iterate(w, z_init)) { status = ORC_NAN; goto done; }
  double theta, phi, fobj;
  if (!eval_primal(w, &w->it, mu, &theta, &phi, &fobj, 1)) { status = ORC_NAN; goto done; }
  double theta_min = 1e-4 * fmax(1.0, theta), theta_max = 1e4 * fmax(1.0, theta);
  w->nfilt = 0;
  for (;;) {
    if (w->resto && it > it_resto0) {
      /* leave the restoration phase?  original infeasibility reduced to kappa_resto * theta_R at a point the original
       * filter (with the point of entry added) accepts */
      double th_o = w->pe_thc + w->pe_thr;
      if (th_o <= RESTO_KAPPA * th_R) {
        double phi_o = w->sigma * w->pe_f - mu_o * w->pe_bar + KAPPA_D * mu_o * w->pe_lin;
        int blocked = th_o >= theta_max_o;
        for (int i = 0; i < w->nfilt_o && !blocked; i++) blocked = th_o >= w->filt_ot[i] && phi_o >= w->filt_op[i];
        if (isfinite(phi_o) && !blocked) {
          w->resto = 0;
          mu = mu_o; tau = tau_o; theta_min = theta_min_o; theta_max = theta_max_o;
          w->nfilt = w->nfilt_o;
          memcpy(w->filt_t, w->filt_ot, sizeof w->filt_t);
          memcpy(w->filt_p, w->filt_op, sizeof w->filt_p);
          iterate_t *q = &w->it;
          double zmax = 0;
          for (int k = 0; k <= N; k++) {
            for (int i = 0; i < nx; i++) { q->lam[k][i] = 0; zmax = fmax(zmax, fmax(q->zlx[k][i], q->zux[k][i])); }
            for (int i = 0; i < 2; i++) { q->lr[k][i] = 0; if (k < N) zmax = fmax(zmax, fmax(q->zlu[k][i], q->zuu[k][i])); if (has_rate(w, k) && i < c->n_rate) zmax = fmax(zmax, fmax(q->vlr[k][i], q->vur[k][i])); }
            for (int j = 0; j < M; j++) { q->lo[k][j] = 0; if (has_obs(w, k)) zmax = fmax(zmax, q->vlo[k][j]); }
          }
          if (zmax > BOUND_MULT_RESET)
            for (int k = 0; k <= N; k++) {
              for (int i = 0; i < nx; i++) { q->zlx[k][i] = w->xbl[i] ? 1.0 : 0.0; q->zux[k][i] = w->xbu[i] ? 1.0 : 0.0; }
              if (k < N) for (int i = 0; i < 2; i++) q->zlu[k][i] = q->zuu[k][i] = 1.0;
              if (has_rate(w, k)) for (int r = 0; r < c->n_rate; r++) q->vlr[k][r] = q->vur[k][r] = 1.0;
              if (has_obs(w, k)) for (int j = 0; j < M; j++) q->vlo[k][j] = 1.0;
            }
          eval_primal(w, &w->it, mu, &theta, &phi, &fobj, 1);
        }
      }
    }
    eval_lin(w);
    if (w->resto) resto_multipliers(w, mu);
    kkt_t kk;
    kkt_pieces(w, &kk);
    double co0;
    err0 = kkt_error(&kk, 0.0, &co0);
    if (err0 <= tol && kk.dual <= DUAL_INF_TOL && kk.prim <= CONSTR_VIOL_TOL && co0 <= COMPL_INF_TOL) {
      /* in restoration: a stationary point of the infeasibility (IPOPT: "converged to a point of local infeasibility") */
      status = w->resto ? ORC_INFEASIBLE : ORC_CONVERGED;
      break;
    }
    if (it >= c->max_iter) { status = ORC_MAXITER; break; }
    int mu_changed = 0;
    while (kkt_error(&kk, mu, 0) <= KAPPA_EPS * mu && mu > tol / 10) {
      mu = fmax(tol / 10, fmin(KAPPA_MU * mu, pow(mu, THETA_MU)));
      tau = fmax(TAU_MIN, 1 - mu);
      w->nfilt = 0;
      mu_changed = 1;
      if (w->resto) { /* psi_mu and zeta move with mu */
        w->zeta = sqrt(mu);
        resto_multipliers(w, mu);
        kkt_pieces(w, &kk);
      }
    }
    if (mu_changed) eval_primal(w, &w->it, mu, &theta, &phi, &fobj, 1);
    build_qp(w, mu);
    double dw = 0.0;
    int ok = SHIPPED(w) ? riccati_shipped(w, 0.0) : riccati(w, 0.0);
    if (!ok) {
      n_reg++;
      dw = dw_last == 0.0 ? DW_FIRST : fmax(DW_MIN, KW_MINUS * dw_last);
      for (;;) {
        ok = SHIPPED(w) ? riccati_shipped(w, dw) : riccati(w, dw);
        if (ok) break;
        dw *= dw_last == 0.0 ? KW_PLUS_FIRST : KW_PLUS;
        if (dw > DW_MAX) break;
      }
      if (ok) dw_last = dw;
    }
    int accepted = 0, armijo = 0;
    double a = 0, a_dual = 1;
    if (ok) {
    double a_max;
    mult_steps(w, mu, tau, &a_max, &a_dual);
    double gd = barrier_dir(w, mu);
    double a_min;
    if (gd < 0 && theta <= theta_min) {
      a_min = GAMMA_THETA;
      if (theta > 0) {
        a_min = fmin(a_min, GAMMA_PHI * theta / (-gd));
        a_min = fmin(a_min, DELTA_SW * pow(theta, S_THETA) / pow(-gd, S_PHI));
      }
    } else if (gd < 0) {
      a_min = fmin(GAMMA_THETA, GAMMA_PHI * theta / (-gd));
    } else {
      a_min = GAMMA_THETA;
    }
    a_min = fmax(GAMMA_ALPHA * a_min, 1e-14);
    a = a_max;
    double th_t = 0, ph_t = 0, f_t = 0;
    while (a >= a_min) {
      make_trial(w, a);
      int fin = eval_primal(w, &w->tr, mu, &th_t, &ph_t, &f_t, 0);
      if (fin && !in_filter(w, th_t, ph_t, theta_max)) {
        int sw = gd < 0 && a * pow(-gd, S_PHI) > DELTA_SW * pow(theta, S_THETA);
        if (theta <= theta_min && sw) {
          if (ph_t <= phi + ETA_PHI * a * gd + 10 * DBL_EPSILON * fabs(phi)) { accepted = 1; armijo = 1; }
        } else if (th_t <= (1 - GAMMA_THETA) * theta || ph_t <= phi - GAMMA_PHI * theta + 10 * DBL_EPSILON * fabs(phi)) {
          accepted = 1;
        }
      }
      if (accepted) break;
      a *= 0.5;
      n_bt++;
    }
    } /* ok */
    if (!accepted) {
      if (!w->resto && err0 <= ACCEPTABLE_TOL) { status = ORC_ACCEPTABLE; break; }
      if (w->resto || !c->restoration || SHIPPED(w)) { status = w->resto ? ORC_RESTO_FAILED : ORC_INFEASIBLE; break; }
      if (c->resto_max_calls > 0 && n_resto >= c->resto_max_calls) { status = ORC_INFEASIBLE; break; }
      /* ---- enter the restoration phase at the current iterate */
      iterate_t *q = &w->it;
      n_resto++;
      it_resto0 = it;
      filter_add(w, (1 - GAMMA_THETA) * theta, phi - GAMMA_PHI * theta); /* the point of entry joins the original filter */
      memcpy(w->filt_ot, w->filt_t, sizeof w->filt_t);
      memcpy(w->filt_op, w->filt_p, sizeof w->filt_p);
      w->nfilt_o = w->nfilt;
      theta_min_o = theta_min; theta_max_o = theta_max;
      mu_o = mu; tau_o = tau; th_R = theta;
      double c_inf = 0;
      for (int k = 0; k <= N; k++) {
        for (int i = 0; i < nx; i++) { w->xR[k][i] = q->x[k][i]; c_inf = fmax(c_inf, fabs(w->cdef[k][i])); q->lam[k][i] = 0; }
        for (int i = 0; i < 2; i++) w->uR[k][i] = k < N ? q->u[k][i] : 0.0;
        if (has_rate(w, k)) for (int r = 0; r < c->n_rate; r++) c_inf = fmax(c_inf, fabs(w->resr[k][r]));
        if (has_obs(w, k)) for (int j = 0; j < M; j++) c_inf = fmax(c_inf, fabs(w->reso[k][j]));
      }
      mu = fmax(mu_o, c_inf);
      tau = fmax(TAU_MIN, 1 - mu);
      w->zeta = sqrt(mu);
      for (int k = 0; k <= N; k++) {
        for (int i = 0; i < nx; i++) { q->zlx[k][i] = fmin(RESTO_RHO, q->zlx[k][i]); q->zux[k][i] = fmin(RESTO_RHO, q->zux[k][i]); }
        for (int i = 0; i < 2; i++) {
          q->zlu[k][i] = fmin(RESTO_RHO, q->zlu[k][i]); q->zuu[k][i] = fmin(RESTO_RHO, q->zuu[k][i]);
          q->vlr[k][i] = fmin(RESTO_RHO, q->vlr[k][i]); q->vur[k][i] = fmin(RESTO_RHO, q->vur[k][i]);
        }
        for (int j = 0; j < M; j++) q->vlo[k][j] = fmin(RESTO_RHO, q->vlo[k][j]);
      }
      w->nfilt = 0;
      w->resto = 1;
      eval_primal(w, &w->it, mu, &theta, &phi, &fobj, 1);
      theta_min = 1e-4 * fmax(1.0, theta);
      theta_max = 1e4 * fmax(1.0, theta);
      dw_last = 0.0;
      continue;
    }
    if (!armijo) filter_add(w, (1 - GAMMA_THETA) * theta, phi - GAMMA_PHI * theta);
    accept_step(w, a, a_dual, mu);
    eval_primal(w, &w->it, mu, &theta, &phi, &fobj, 1);
    it++;
    if (w->resto) n_resto_it++;
  }
done:
  write_out(w, z_out, lam_eq_out);
  if (info) {
    info->f = status == ORC_NAN ? NAN : objective(w, &w->it);
    info->err = err0; info->mu = mu; info->obj_scale = w->sigma;
    info->status = status; info->iters = it; info->n_reg = n_reg; info->n_backtrack = n_bt;
    info->n_resto = n_resto; info->n_resto_iter = n_resto_it;
  }
  return 0;
}

int orc_solve(const orc_cfg *cfg, const double *x0, const double *xs, const double *obs, const double *z_init,
              double *z_out, double *lam_eq_out, orc_info *info) {
  if (cfg->N < 1 || cfg->N > ORC_NMAX || cfg->M > ORC_MMAX) return -1;
  ws_t *w = (ws_t *)malloc(sizeof(ws_t));
  if (!w) return -2;
  setup(w, cfg, x0, xs, obs);
  int rc = solve_ws(w, z_init, z_out, lam_eq_out, info);
  free(w);
  return rc;
}

typedef struct {
  const orc_cfg *cfg;
  int B;
  const double *x0, *xs, *obs, *z_init;
  double *u0, *cost, *z_out;
  int32_t *status, *iters;
  atomic_int next;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const orc_cfg *cfg = j->cfg;
  int nx = orc_nx(cfg), N = cfg->N;
  int M = cfg->obs_mode == ORC_OBS_NONE ? 0 : cfg->M;
  size_t nv = 2 * (size_t)N + (size_t)nx * (N + 1);
  size_t so = (size_t)M * (N + 1) * 6;
  ws_t *w = (ws_t *)malloc(sizeof(ws_t));
  double *z = (double *)malloc(sizeof(double) * nv);
  for (;;) {
    int b = atomic_fetch_add(&j->next, 1);
    if (b >= j->B) break;
    orc_info info;
    size_t sxs = cfg->ref_mode == ORC_REF_TRAJECTORY ? (size_t)nx * cfg->N : (size_t)nx;
    setup(w, cfg, j->x0 + (size_t)b * nx, j->xs + (size_t)b * sxs, j->obs ? j->obs + (size_t)b * so : 0);
    solve_ws(w, j->z_init ? j->z_init + (size_t)b * nv : 0, z, 0, &info);
    j->u0[2 * b] = z[0];
    j->u0[2 * b + 1] = z[1];
    j->cost[b] = info.f;
    j->status[b] = info.status;
    j->iters[b] = info.iters;
    if (j->z_out) memcpy(j->z_out + (size_t)b * nv, z, sizeof(double) * nv);
  }
  free(w);
  free(z);
  return 0;
}

int orc_solve_batch(const orc_cfg *cfg, int B, const double *x0, const double *xs, const double *obs,
                    const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters, double *z_out,
                    int nthreads) {
  if (cfg->N < 1 || cfg->N > ORC_NMAX || cfg->M > ORC_MMAX) return -1;
  batch_job job = {cfg, B, x0, xs, obs, z_init, u0, cost, z_out, status, iters, 0};
  atomic_init(&job.next, 0);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], 0, batch_worker, &job);
  batch_worker(&job);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
  return 0;
}

int orc_newton_step(const orc_cfg *cfg, const double *x0, const double *xs, const double *obs, const double *z,
                    double mu, double dw, double obj_scale, double *dz, double *lam_plus) {
  ws_t *w = (ws_t *)malloc(sizeof(ws_t));
  if (!w) return -2;
  setup(w, cfg, x0, xs, obs);
  orc_cfg c2 = *cfg;
  c2.init_mode = ORC_INIT_AS_GIVEN;
  w->c = &c2;
  int rc = 0;
  if (!init_iterate(w, z)) rc = 1;
  w->sigma = obj_scale;
  double th, ph, f;
  if (!rc && !eval_primal(w, &w->it, mu, &th, &ph, &f, 1)) rc = 1;
  if (!rc) {
    eval_lin(w);
    build_qp(w, mu);
    if (!(SHIPPED(w) ? riccati_shipped(w, dw) : riccati(w, dw))) rc = 2;
  }
  if (!rc) {
    int nx = w->nx, N = w->N;
    for (int k = 0; k < N; k++) { dz[2 * k] = w->du[k][0]; dz[2 * k + 1] = w->du[k][1]; }
    for (int k = 0; k <= N; k++) for (int i = 0; i < nx; i++) { dz[2 * N + nx * k + i] = w->dx[k][i]; if (lam_plus) lam_plus[nx * k + i] = w->lamp[k][i]; }
  }
  free(w);
  return rc;
}
