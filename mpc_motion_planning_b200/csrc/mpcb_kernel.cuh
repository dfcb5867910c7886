// mpcb200 device code: one scenario per warp, FP64 primal-dual interior point with a
// stage-wise Riccati recursion.  sm_100a only.
//
// What this replaces in the reference (PKG = CasaDi_MPC_Optimize_Multishoot):
//   * CasADi SX graph + AD of the NLP built in MPC_optimize.optimize_problem
//     (PKG/MPC_CBF_optimize_kin.py:136-255, _kin_pre.py:136-261, _dyn.py:137-250)
//       -> Model::f / jac / hess below (analytic derivatives), stage_* functions
//   * IPOPT + MUMPS behind ca.nlpsol / solver(...)  (PKG/MPC_CBF_optimize_kin.py:251-254,
//     PKG/main_cbf_kin_c_sim.py:100)
//       -> solve_kernel: barrier loop, filter line search, inertia correction;
//          riccati_backward/forward replace the sparse LDL^T of the KKT matrix.
//
// Data layout: every per-stage quantity of one scenario is a row [S = N+1] of doubles in
// shared memory (struct-of-arrays, stage index fastest), so the stage-parallel phases
// (lane = stage) are bank-conflict free and the serial recursions read broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mpcb {

// ---- IPOPT default constants used by the algorithm (same values as the specification
// in oracle/ipm_dense.py; restated here because product code never includes oracle/) ----
#define MPCB_KAPPA_EPS 10.0
#define MPCB_KAPPA_MU 0.2
#define MPCB_THETA_MU 1.5
#define MPCB_TAU_MIN 0.99
#define MPCB_BOUND_PUSH 1e-2
#define MPCB_BOUND_FRAC 1e-2
#define MPCB_S_MAX 100.0
#define MPCB_KAPPA_SIGMA 1e10
#define MPCB_KAPPA_D 1e-5
#define MPCB_OBJ_SCALE_MAX_GRAD 100.0
#define MPCB_GAMMA_THETA 1e-5
#define MPCB_GAMMA_PHI 1e-8
#define MPCB_S_THETA 1.1
#define MPCB_S_PHI 2.3
#define MPCB_ETA_PHI 1e-8
#define MPCB_GAMMA_ALPHA 0.05
#define MPCB_DW_FIRST 1e-4
#define MPCB_DW_MIN 1e-20
#define MPCB_DW_MAX 1e40
#define MPCB_KW_MINUS (1.0 / 3.0)
#define MPCB_KW_PLUS 8.0
#define MPCB_KW_PLUS_FIRST 100.0
#define MPCB_DUAL_INF_TOL 1.0
#define MPCB_CONSTR_VIOL_TOL 1e-4
#define MPCB_COMPL_INF_TOL 1e-4
#define MPCB_DBL_EPS 2.220446049250313e-16
#define MPCB_FILTER_SLOTS 4 /* 4 x 32 lanes = 128 filter entries */

struct KParams {
  int B, N, obs_mode, du0_cost, init_mode, max_iter;
  int rate_ctrl[2];
  int n_eq, n_bm;  // counts used by the IPOPT error scaling
  double T;
  double Q[6], R[2], DR[2];
  double rate_lo[2], rate_hi[2];  // relaxed
  double u_lo[2], u_hi[2];        // relaxed
  double x_lo[6], x_hi[6];        // relaxed (only the model's bounded components are read)
  double obs_lo;                  // relaxed
  double ego_hl, ego_hw, safe_l, safe_w, dyn_sx, dyn_sy;
  double Veh_l, lf, lr, m, Iz, aopt_f, aopt_r, Fymax_f, Fymax_r;
  double tol, mu_init;
  const double *x0, *xs, *obs, *z_init;
  double *u0, *cost, *z_out, *lam_out;
  int32_t *status, *iters;
};

// ------------------------------------------------------------------------------------
// warp reductions (all lanes end with the same value)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------
// vehicle models
// ------------------------------------------------------------------------------------
// Kinematic bicycle, x=[x,y,phi,vx], u=[df,ax]   (PKG/MPC_CBF_optimize_kin.py:153-156)
struct KinModel {
  static constexpr int NX = 4;
  static constexpr int NBX = 2;  // bounded state components: y, vx  (PKG/..._kin.py:97-105)
  static constexpr int NJ = 6;   // stored Jacobian entries
  __device__ static __forceinline__ constexpr int bx(int i) { return i == 0 ? 1 : 3; }

  __device__ static __forceinline__ void f(const double *x, const double *u, const KParams &p, double *f) {
    double s, c;
    sincos(x[2], &s, &c);
    f[0] = x[3] * c;
    f[1] = x[3] * s;
    f[2] = x[3] * tan(u[0]) / p.Veh_l;
    f[3] = u[1];
  }
  // f and the nonzero entries of df/d[x;u]
  __device__ static __forceinline__ void fjac(const double *x, const double *u, const KParams &p, double *f, double *J) {
    double s, c;
    sincos(x[2], &s, &c);
    double t = tan(u[0]);
    f[0] = x[3] * c;
    f[1] = x[3] * s;
    f[2] = x[3] * t / p.Veh_l;
    f[3] = u[1];
    J[0] = -x[3] * s;                       // d f0 / d phi
    J[1] = c;                               // d f0 / d v
    J[2] = x[3] * c;                        // d f1 / d phi
    J[3] = s;                               // d f1 / d v
    J[4] = t / p.Veh_l;                     // d f2 / d v
    J[5] = x[3] * (1.0 + t * t) / p.Veh_l;  // d f2 / d df
  }
  // A = I + T df/dx, B = T df/du (dense, structural zeros written as literals)
  __device__ static __forceinline__ void expand(const double *J, double T, double A[4][4], double B[4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) A[i][j] = (i == j) ? 1.0 : 0.0;
      B[i][0] = 0.0;
      B[i][1] = 0.0;
    }
    A[0][2] = T * J[0];
    A[0][3] = T * J[1];
    A[1][2] = T * J[2];
    A[1][3] = T * J[3];
    A[2][3] = T * J[4];
    B[2][0] = T * J[5];
    B[3][1] = T;
  }
  // H += -T * sum_i lam_i d2 f_i / d[x;u]^2 (Hxx symmetric full, Hux 2xNX, Huu 2x2)
  __device__ static __forceinline__ void add_hess(const double *x, const double *u, const KParams &p, const double *lam,
                                                  double Hxx[4][4], double Hux[2][4], double Huu[2][2]) {
    double s, c;
    sincos(x[2], &s, &c);
    double t = tan(u[0]), v = x[3];
    double sec2 = 1.0 + t * t;
    double h22 = lam[0] * (-v * c) + lam[1] * (-v * s);
    double h23 = lam[0] * (-s) + lam[1] * c;
    double h3d = lam[2] * sec2 / p.Veh_l;
    double hdd = lam[2] * 2.0 * v * sec2 * t / p.Veh_l;
    Hxx[2][2] += -p.T * h22;
    Hxx[2][3] += -p.T * h23;
    Hxx[3][2] += -p.T * h23;
    Hux[0][3] += -p.T * h3d;
    Huu[0][0] += -p.T * hdd;
  }
};

// ------------------------------------------------------------------------------------
// shared-memory layout of one scenario
// ------------------------------------------------------------------------------------
template <class Mdl, int NR, int MO>
struct Layout {
  static constexpr int NX = Mdl::NX, NBX = Mdl::NBX, NJ = Mdl::NJ;
  static constexpr int NP = NX * (NX + 1) / 2;
  // iterate
  static constexpr int X = 0;
  static constexpr int U = X + NX;
  static constexpr int LAM = U + 2;
  static constexpr int ZLX = LAM + NX;
  static constexpr int ZUX = ZLX + NBX;
  static constexpr int ZLU = ZUX + NBX;
  static constexpr int ZUU = ZLU + 2;
  static constexpr int SR = ZUU + 2;
  static constexpr int VLR = SR + NR;
  static constexpr int VUR = VLR + NR;
  static constexpr int LR = VUR + NR;
  static constexpr int SO = LR + NR;
  static constexpr int VLO = SO + MO;
  static constexpr int LO = VLO + MO;
  // obstacle data
  static constexpr int OCX = LO + MO;
  static constexpr int OCY = OCX + MO;
  static constexpr int ISX = OCY + MO;
  static constexpr int ISY = ISX + MO;
  // evaluation
  static constexpr int CDEF = ISY + MO;
  static constexpr int JAC = CDEF + NX;
  // condensed QP
  static constexpr int HXX = JAC + NJ;
  static constexpr int HUX = HXX + NP;
  static constexpr int HUU = HUX + 2 * NX;
  static constexpr int EE = HUU + 3;
  static constexpr int GX = EE + 2;
  static constexpr int GU = GX + NX;
  static constexpr int TK = GU + 2;
  // Riccati gains
  static constexpr int KX = TK + 2;
  static constexpr int KW = KX + 2 * NX;
  static constexpr int KK = KW + 4;
  // direction
  static constexpr int DX = KK + 2;
  static constexpr int DU = DX + NX;
  static constexpr int LAMP = DU + 2;
  static constexpr int DSR = LAMP + NX;
  static constexpr int LRP = DSR + NR;
  static constexpr int DSO = LRP + NR;
  static constexpr int LOP = DSO + MO;
  static constexpr int NFIELDS = LOP + MO;
  __host__ __device__ static constexpr size_t bytes(int N) { return sizeof(double) * (size_t)NFIELDS * (size_t)(N + 1); }
};

__device__ __forceinline__ double push_in(double v, double lo, double hi) {
  // IPOPT bound_push / bound_frac for a two-sided interval
  double pl = fmin(MPCB_BOUND_PUSH * fmax(1.0, fabs(lo)), MPCB_BOUND_FRAC * (hi - lo));
  double pu = fmin(MPCB_BOUND_PUSH * fmax(1.0, fabs(hi)), MPCB_BOUND_FRAC * (hi - lo));
  v = fmax(v, lo + pl);
  v = fmin(v, hi - pu);
  return v;
}
__device__ __forceinline__ double push_lo(double v, double lo) { return fmax(v, lo + MPCB_BOUND_PUSH * fmax(1.0, fabs(lo))); }
__device__ __forceinline__ double clampz(double z, double mu, double gap) {
  return fmax(fmin(z, MPCB_KAPPA_SIGMA * mu / gap), mu / (MPCB_KAPPA_SIGMA * gap));
}
__device__ __forceinline__ int pidx(int i, int j, int n) {  // packed upper-triangular index, i<=j
  return i * n - i * (i - 1) / 2 + (j - i);
}

// ------------------------------------------------------------------------------------
// the solver: one warp = one scenario
// ------------------------------------------------------------------------------------
template <class Mdl, int NR, int MO, int OBS_MODE>
struct Solver {
  using L = Layout<Mdl, NR, MO>;
  static constexpr int NX = Mdl::NX, NBX = Mdl::NBX, NJ = Mdl::NJ, NP = L::NP;

  const KParams &p;
  double *sm;
  int S, N, lane;
  double sigma;
  double x0[NX], xs[NX];

  __device__ Solver(const KParams &p_, double *sm_, int lane_) : p(p_), sm(sm_), S(p_.N + 1), N(p_.N), lane(lane_) {}

  __device__ __forceinline__ double &at(int field, int k) { return sm[field * S + k]; }
  __device__ __forceinline__ bool has_rate(int k) const { return NR > 0 && k >= 1 && k <= N - 1; }
  __device__ __forceinline__ bool has_obs(int k) const {
    if (OBS_MODE == 1) return k <= N - 1;
    if (OBS_MODE == 2) return k <= N;
    return false;
  }

  // obstacle row value (+ gradient, Hessian) at (px,py) for stage k, obstacle j
  __device__ __forceinline__ double obs_val(int k, int j, double px, double py) {
    double dx = px - at(L::OCX + j, k), dy = py - at(L::OCY + j, k);
    double e = dx * dx * at(L::ISX + j, k) + dy * dy * at(L::ISY + j, k) - 1.0;
    if (OBS_MODE == 1) return e;              // PKG/MPC_CBF_optimize_kin.py:244,247
    return e > 0.0 ? sqrt(e) : nan("");       // PKG/MPC_CBF_optimize_dyn.py:243
  }
  __device__ __forceinline__ void obs_grad(int k, int j, double px, double py, double &d, double &gx, double &gy, double &hxx,
                                           double &hxy, double &hyy) {
    double dx = px - at(L::OCX + j, k), dy = py - at(L::OCY + j, k);
    double a = at(L::ISX + j, k), b = at(L::ISY + j, k);
    double e = dx * dx * a + dy * dy * b - 1.0;
    if (OBS_MODE == 1) {
      d = e; gx = 2 * dx * a; gy = 2 * dy * b; hxx = 2 * a; hxy = 0.0; hyy = 2 * b;
    } else {
      double q = sqrt(e);
      double ex = 2 * dx * a, ey = 2 * dy * b;
      d = q; gx = ex / (2 * q); gy = ey / (2 * q);
      double q3 = 4 * q * q * q;
      hxx = a / q - ex * ex / q3; hxy = -ex * ey / q3; hyy = b / q - ey * ey / q3;
    }
  }

  // gradient of the unscaled objective wrt u_k[i]  (PKG/MPC_CBF_optimize_kin.py:199-205)
  __device__ __forceinline__ double grad_u(int k, int i, double uk, double ukm1, double ukp1) const {
    double v = 2 * p.R[i] * uk;
    if (k > 0) v += 2 * p.DR[i] * (uk - ukm1);
    else if (p.du0_cost) v += 2 * p.DR[i] * uk;
    if (k + 1 <= N - 1) v -= 2 * p.DR[i] * (ukp1 - uk);
    return v;
  }

  // ---------------------------------------------------------------- primal evaluation
  // constraint residual 1-norm, objective, barrier pieces at z + alpha*dz (TRIAL) or at z
  // (storing the defects).  Returns lane-uniform sums.
  template <bool TRIAL>
  __device__ void eval_primal(double alpha, double &theta, double &fobj, double &bar, double &lin) {
    double th = 0, fo = 0, br = 0, ln = 0;
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], uk[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k) + (TRIAL ? alpha * at(L::DX + i, k) : 0.0);
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double c0 = xk[i] - x0[i];
          th += fabs(c0);
          if (!TRIAL) at(L::CDEF + i, 0) = c0;
        }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) uk[i] = at(L::U + i, k) + (TRIAL ? alpha * at(L::DU + i, k) : 0.0);
        double f[NX];
        Mdl::f(xk, uk, p, f);
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double xn = at(L::X + i, k + 1) + (TRIAL ? alpha * at(L::DX + i, k + 1) : 0.0);
          double d = xn - (xk[i] + p.T * f[i]);
          th += fabs(d);
          if (!TRIAL) at(L::CDEF + i, k + 1) = d;
          double e = xk[i] - xs[i];
          fo += p.Q[i] * e * e;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
          br += log(uk[i] - p.u_lo[i]) + log(p.u_hi[i] - uk[i]);
          fo += p.R[i] * uk[i] * uk[i];
          if (k > 0) {
            double um = at(L::U + i, k - 1) + (TRIAL ? alpha * at(L::DU + i, k - 1) : 0.0);
            double e = uk[i] - um;
            fo += p.DR[i] * e * e;
          } else if (p.du0_cost) {
            fo += p.DR[i] * uk[i] * uk[i];
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = Mdl::bx(b);
        br += log(xk[i] - p.x_lo[i]) + log(p.x_hi[i] - xk[i]);
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          double um = at(L::U + ci, k - 1) + (TRIAL ? alpha * at(L::DU + ci, k - 1) : 0.0);
          double ukc = ci == 0 ? uk[0] : uk[1];
          double s = at(L::SR + r, k) + (TRIAL ? alpha * at(L::DSR + r, k) : 0.0);
          th += fabs(ukc - um - s);
          br += log(s - p.rate_lo[r]) + log(p.rate_hi[r] - s);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double d = obs_val(k, j, xk[0], xk[1]);
          double s = at(L::SO + j, k) + (TRIAL ? alpha * at(L::DSO + j, k) : 0.0);
          th += fabs(d - s);
          br += log(s - p.obs_lo);
          ln += s - p.obs_lo;
        }
      }
    }
    theta = warp_sum(th);
    fobj = warp_sum(fo);
    bar = warp_sum(br);
    lin = warp_sum(ln);
  }

  // ---------------------------------------------------------------- Jacobians + KKT error pieces
  struct Kkt { double dual, prim, cmin, cmax, sum_lam, sum_z; };

  __device__ void eval_lin_kkt(Kkt &o) {
    // pass 1: Jacobians of the dynamics (stored compactly)
    for (int k = lane; k < N; k += 32) {
      double xk[NX], uk[2], f[NX], J[NJ];
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k);
      uk[0] = at(L::U + 0, k);
      uk[1] = at(L::U + 1, k);
      Mdl::fjac(xk, uk, p, f, J);
#pragma unroll
      for (int i = 0; i < NJ; i++) at(L::JAC + i, k) = J[i];
    }
    __syncwarp();
    double dual = 0, prim = 0, cmin = INFINITY, cmax = -INFINITY, sl = 0, sz = 0;
#define MPCB_COMPL(gap, mult) do { double p_ = (gap) * (mult); cmin = fmin(cmin, p_); cmax = fmax(cmax, p_); sz += (mult); } while (0)
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], lam[NX], lam1[NX], A[NX][NX], B[NX][2];
#pragma unroll
      for (int i = 0; i < NX; i++) { xk[i] = at(L::X + i, k); lam[i] = at(L::LAM + i, k); lam1[i] = 0; }
      if (k < N) {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        Mdl::expand(J, p.T, A, B);
#pragma unroll
        for (int i = 0; i < NX; i++) lam1[i] = at(L::LAM + i, k + 1);
      }
      double rx[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double r = lam[i];
        if (k < N) {
          r += sigma * 2 * p.Q[i] * (xk[i] - xs[i]);
#pragma unroll
          for (int a = 0; a < NX; a++) r -= A[a][i] * lam1[a];
        }
        rx[i] = r;
        prim = fmax(prim, fabs(at(L::CDEF + i, k)));
        sl += fabs(lam[i]);
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = Mdl::bx(b);
        double zl = at(L::ZLX + b, k), zu = at(L::ZUX + b, k);
        rx[i] += -zl + zu;
        MPCB_COMPL(xk[i] - p.x_lo[i], zl);
        MPCB_COMPL(p.x_hi[i] - xk[i], zu);
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double d, gx, gy, hxx, hxy, hyy;
          obs_grad(k, j, xk[0], xk[1], d, gx, gy, hxx, hxy, hyy);
          double lo = at(L::LO + j, k), vl = at(L::VLO + j, k), s = at(L::SO + j, k);
          rx[0] += lo * gx;
          rx[1] += lo * gy;
          dual = fmax(dual, fabs(-lo - vl));
          prim = fmax(prim, fabs(d - s));
          MPCB_COMPL(s - p.obs_lo, vl);
          sl += fabs(lo);
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) dual = fmax(dual, fabs(rx[i]));
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k);
          double um = k > 0 ? at(L::U + i, k - 1) : 0.0;
          double up = k + 1 <= N - 1 ? at(L::U + i, k + 1) : 0.0;
          double zl = at(L::ZLU + i, k), zu = at(L::ZUU + i, k);
          double r = sigma * grad_u(k, i, uk, um, up) - zl + zu;
#pragma unroll
          for (int a = 0; a < NX; a++) r -= B[a][i] * lam1[a];
#pragma unroll
          for (int rr = 0; rr < NR; rr++)
            if (p.rate_ctrl[rr] == i) {
              if (has_rate(k)) r += at(L::LR + rr, k);
              if (has_rate(k + 1)) r -= at(L::LR + rr, k + 1);
            }
          dual = fmax(dual, fabs(r));
          MPCB_COMPL(uk - p.u_lo[i], zl);
          MPCB_COMPL(p.u_hi[i] - uk, zu);
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          double s = at(L::SR + r, k), vl = at(L::VLR + r, k), vu = at(L::VUR + r, k), lr = at(L::LR + r, k);
          dual = fmax(dual, fabs(-lr - vl + vu));
          prim = fmax(prim, fabs(at(L::U + ci, k) - at(L::U + ci, k - 1) - s));
          MPCB_COMPL(s - p.rate_lo[r], vl);
          MPCB_COMPL(p.rate_hi[r] - s, vu);
          sl += fabs(lr);
        }
      }
    }
#undef MPCB_COMPL
    o.dual = warp_max(dual);
    o.prim = warp_max(prim);
    o.cmin = warp_min(cmin);
    o.cmax = warp_max(cmax);
    o.sum_lam = warp_sum(sl);
    o.sum_z = warp_sum(sz);
  }

  __device__ __forceinline__ double kkt_error(const Kkt &o, double mu, double &co) const {
    co = p.n_bm > 0 ? fmax(fabs(o.cmax - mu), fabs(o.cmin - mu)) : 0.0;
    double s_d = fmax(MPCB_S_MAX, (o.sum_lam + o.sum_z) / fmax(1.0, (double)(p.n_eq + p.n_bm))) / MPCB_S_MAX;
    double s_c = fmax(MPCB_S_MAX, o.sum_z / fmax(1.0, (double)p.n_bm)) / MPCB_S_MAX;
    return fmax(fmax(o.dual / s_d, o.prim), co / s_c);
  }

  // ---------------------------------------------------------------- condensed QP (stage parallel)
  // Effective stage Hessian/gradient with the slack rows eliminated and the primal
  // regularisation dw applied.
  __device__ void build_qp(double mu, double dw) {
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], uk[2] = {0, 0};
      double Hxx[NX][NX], Hux[2][NX], Huu[2][2], gx[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
#pragma unroll
        for (int j = 0; j < NX; j++) Hxx[i][j] = 0;
        Hux[0][i] = 0;
        Hux[1][i] = 0;
      }
      Huu[0][0] = Huu[0][1] = Huu[1][0] = Huu[1][1] = 0;
      if (k < N) {
        uk[0] = at(L::U + 0, k);
        uk[1] = at(L::U + 1, k);
        double lam1[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) lam1[i] = at(L::LAM + i, k + 1);
        Mdl::add_hess(xk, uk, p, lam1, Hxx, Hux, Huu);
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double g = 0;
        if (k < N) {
          Hxx[i][i] += sigma * 2 * p.Q[i];
          g += sigma * 2 * p.Q[i] * (xk[i] - xs[i]);
        }
        Hxx[i][i] += dw;
        gx[i] = g;
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = Mdl::bx(b);
        double gl = xk[i] - p.x_lo[i], gh = p.x_hi[i] - xk[i];
        Hxx[i][i] += at(L::ZLX + b, k) / gl + at(L::ZUX + b, k) / gh;
        gx[i] += -mu / gl + mu / gh;
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double d, ox, oy, hxx, hxy, hyy;
          obs_grad(k, j, xk[0], xk[1], d, ox, oy, hxx, hxy, hyy);
          double s = at(L::SO + j, k), gap = s - p.obs_lo;
          double D = at(L::VLO + j, k) / gap + dw;
          double gs = -mu / gap + MPCB_KAPPA_D * mu;
          double lo = at(L::LO + j, k);
          double t = D * (d - s) + gs;
          Hxx[0][0] += lo * hxx + D * ox * ox;
          Hxx[0][1] += lo * hxy + D * ox * oy;
          Hxx[1][0] += lo * hxy + D * ox * oy;
          Hxx[1][1] += lo * hyy + D * oy * oy;
          gx[0] += ox * t;
          gx[1] += oy * t;
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        at(L::GX + i, k) = gx[i];
#pragma unroll
        for (int j = i; j < NX; j++) at(L::HXX + pidx(i, j, NX), k) = Hxx[i][j];
      }
      if (k < N) {
        double E[2] = {0, 0}, t[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double g = sigma * 2 * p.R[i] * uk[i];
          double hd = sigma * 2 * p.R[i] + dw;
          if (k == 0 && p.du0_cost) {
            hd += sigma * 2 * p.DR[i];
            g += sigma * 2 * p.DR[i] * uk[i];
          }
          double gl = uk[i] - p.u_lo[i], gh = p.u_hi[i] - uk[i];
          hd += at(L::ZLU + i, k) / gl + at(L::ZUU + i, k) / gh;
          g += -mu / gl + mu / gh;
          Huu[i][i] += hd;
          at(L::GU + i, k) = g;
          if (k >= 1) {
            E[i] = sigma * 2 * p.DR[i];
            t[i] = sigma * 2 * p.DR[i] * (uk[i] - at(L::U + i, k - 1));
          }
        }
        if (has_rate(k)) {
#pragma unroll
          for (int r = 0; r < NR; r++) {
            int ci = p.rate_ctrl[r];
            double s = at(L::SR + r, k);
            double gl = s - p.rate_lo[r], gh = p.rate_hi[r] - s;
            double D = at(L::VLR + r, k) / gl + at(L::VUR + r, k) / gh + dw;
            double gs = -mu / gl + mu / gh;
            double res = at(L::U + ci, k) - at(L::U + ci, k - 1) - s;
            double tt = D * res + gs;
            if (ci == 0) { E[0] += D; t[0] += tt; } else { E[1] += D; t[1] += tt; }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
          at(L::EE + i, k) = E[i];
          at(L::TK + i, k) = t[i];
#pragma unroll
          for (int j = 0; j < NX; j++) at(L::HUX + i * NX + j, k) = Hux[i][j];
        }
        at(L::HUU + 0, k) = Huu[0][0];
        at(L::HUU + 1, k) = Huu[0][1];
        at(L::HUU + 2, k) = Huu[1][1];
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- Riccati (serial over stages)
  // Every lane runs the same recursion on broadcast reads; lane 0 stores the gains.
  // Returns false when some F_uu is not positive definite (wrong inertia).
  __device__ bool riccati_backward() {
    double Pxx[NX][NX], Pxw[NX][2], Pww[2][2], px[NX], pw[2];
#pragma unroll
    for (int i = 0; i < NX; i++) {
#pragma unroll
      for (int j = 0; j < NX; j++) Pxx[i][j] = at(L::HXX + (i <= j ? pidx(i, j, NX) : pidx(j, i, NX)), N);
      px[i] = at(L::GX + i, N);
      Pxw[i][0] = Pxw[i][1] = 0;
    }
    Pww[0][0] = Pww[0][1] = Pww[1][0] = Pww[1][1] = 0;
    pw[0] = pw[1] = 0;
    bool ok = true;
    for (int k = N - 1; k >= 0; k--) {
      double A[NX][NX], B[NX][2];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        Mdl::expand(J, p.T, A, B);
      }
      double E[2] = {at(L::EE + 0, k), at(L::EE + 1, k)};
      double t[2] = {at(L::TK + 0, k), at(L::TK + 1, k)};
      double b[NX], Pb[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) b[i] = -at(L::CDEF + i, k + 1);
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = px[i];
#pragma unroll
        for (int j = 0; j < NX; j++) s += Pxx[i][j] * b[j];
        Pb[i] = s;
      }
      double PA[NX][NX], PB[NX][2];
#pragma unroll
      for (int i = 0; i < NX; i++) {
#pragma unroll
        for (int j = 0; j < NX; j++) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NX; a++) s += Pxx[i][a] * A[a][j];
          PA[i][j] = s;
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double s = 0;
#pragma unroll
          for (int a = 0; a < NX; a++) s += Pxx[i][a] * B[a][j];
          PB[i][j] = s;
        }
      }
      double Fxx[NX][NX], Fux[2][NX], Fuu[2][2], fx[NX], fu[2];
#pragma unroll
      for (int i = 0; i < NX; i++)
#pragma unroll
        for (int j = 0; j < NX; j++) {
          double s = at(L::HXX + (i <= j ? pidx(i, j, NX) : pidx(j, i, NX)), k);
#pragma unroll
          for (int a = 0; a < NX; a++) s += A[a][i] * PA[a][j];
          Fxx[i][j] = s;
        }
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < NX; j++) {
          double s = at(L::HUX + i * NX + j, k);
#pragma unroll
          for (int a = 0; a < NX; a++) s += B[a][i] * PA[a][j] + Pxw[a][i] * A[a][j];
          Fux[i][j] = s;
        }
#pragma unroll
      for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
          double s = at(L::HUU + (i + j), k) + Pww[i][j];  // packed [00,01,11]
          if (i == j) s += E[i];
#pragma unroll
          for (int a = 0; a < NX; a++) s += B[a][i] * PB[a][j] + B[a][i] * Pxw[a][j] + Pxw[a][i] * B[a][j];
          Fuu[i][j] = s;
        }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = at(L::GX + i, k);
#pragma unroll
        for (int a = 0; a < NX; a++) s += A[a][i] * Pb[a];
        fx[i] = s;
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        double s = at(L::GU + i, k) + t[i] + pw[i];
#pragma unroll
        for (int a = 0; a < NX; a++) s += B[a][i] * Pb[a] + Pxw[a][i] * b[a];
        fu[i] = s;
      }
      double det = Fuu[0][0] * Fuu[1][1] - Fuu[0][1] * Fuu[1][0];
      if (!(Fuu[0][0] > 0.0) || !(det > 0.0) || !isfinite(det)) { ok = false; break; }
      double id = 1.0 / det;
      double Fi[2][2] = {{Fuu[1][1] * id, -Fuu[0][1] * id}, {-Fuu[1][0] * id, Fuu[0][0] * id}};
      double Kx[2][NX], Kw[2][2], kk[2];
#pragma unroll
      for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < NX; j++) Kx[i][j] = -(Fi[i][0] * Fux[0][j] + Fi[i][1] * Fux[1][j]);
#pragma unroll
        for (int j = 0; j < 2; j++) Kw[i][j] = Fi[i][j] * E[j];
        kk[i] = -(Fi[i][0] * fu[0] + Fi[i][1] * fu[1]);
      }
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
          for (int j = 0; j < NX; j++) at(L::KX + i * NX + j, k) = Kx[i][j];
          at(L::KW + i * 2 + 0, k) = Kw[i][0];
          at(L::KW + i * 2 + 1, k) = Kw[i][1];
          at(L::KK + i, k) = kk[i];
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
#pragma unroll
        for (int j = 0; j < NX; j++) Pxx[i][j] = Fxx[i][j] + Fux[0][i] * Kx[0][j] + Fux[1][i] * Kx[1][j];
#pragma unroll
        for (int j = 0; j < 2; j++) Pxw[i][j] = Fux[0][i] * Kw[0][j] + Fux[1][i] * Kw[1][j];
        px[i] = fx[i] + Fux[0][i] * kk[0] + Fux[1][i] * kk[1];
      }
#pragma unroll
      for (int i = 0; i < NX; i++)
#pragma unroll
        for (int j = i + 1; j < NX; j++) {
          double m = 0.5 * (Pxx[i][j] + Pxx[j][i]);
          Pxx[i][j] = m;
          Pxx[j][i] = m;
        }
#pragma unroll
      for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) Pww[i][j] = (i == j ? E[i] : 0.0) - E[i] * Kw[i][j];
        pw[i] = -t[i] - E[i] * kk[i];
      }
      double m = 0.5 * (Pww[0][1] + Pww[1][0]);
      Pww[0][1] = m;
      Pww[1][0] = m;
    }
    __syncwarp();
    return ok;
  }

  __device__ void riccati_forward() {
    double dx[NX], dum[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < NX; i++) dx[i] = -at(L::CDEF + i, 0);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::DX + i, 0) = dx[i];
    }
    for (int k = 0; k < N; k++) {
      double A[NX][NX], B[NX][2], du[2];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        Mdl::expand(J, p.T, A, B);
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        double s = at(L::KK + i, k);
#pragma unroll
        for (int j = 0; j < NX; j++) s += at(L::KX + i * NX + j, k) * dx[j];
        if (k > 0) s += at(L::KW + i * 2 + 0, k) * dum[0] + at(L::KW + i * 2 + 1, k) * dum[1];
        du[i] = s;
      }
      double dn[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = -at(L::CDEF + i, k + 1);
#pragma unroll
        for (int j = 0; j < NX; j++) s += A[i][j] * dx[j];
        s += B[i][0] * du[0] + B[i][1] * du[1];
        dn[i] = s;
      }
      if (lane == 0) {
        at(L::DU + 0, k) = du[0];
        at(L::DU + 1, k) = du[1];
#pragma unroll
        for (int i = 0; i < NX; i++) at(L::DX + i, k + 1) = dn[i];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) dx[i] = dn[i];
      dum[0] = du[0];
      dum[1] = du[1];
    }
    __syncwarp();
  }

  // new dynamics multipliers: parallel r_k = Hxx_eff dx + Hux' du + gx_eff, then the serial
  // adjoint recursion lam+_k = A_k' lam+_{k+1} - r_k
  __device__ void adjoint() {
    for (int k = lane; k <= N; k += 32) {
      double dx[NX], du[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < NX; i++) dx[i] = at(L::DX + i, k);
      if (k < N) { du[0] = at(L::DU + 0, k); du[1] = at(L::DU + 1, k); }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = at(L::GX + i, k);
#pragma unroll
        for (int j = 0; j < NX; j++) s += at(L::HXX + (i <= j ? pidx(i, j, NX) : pidx(j, i, NX)), k) * dx[j];
        if (k < N) s += at(L::HUX + 0 * NX + i, k) * du[0] + at(L::HUX + 1 * NX + i, k) * du[1];
        at(L::LAMP + i, k) = s;
      }
    }
    __syncwarp();
    double lp[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) lp[i] = -at(L::LAMP + i, N);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::LAMP + i, N) = lp[i];
    }
    for (int k = N - 1; k >= 0; k--) {
      double A[NX][NX], B[NX][2], ln[NX];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        Mdl::expand(J, p.T, A, B);
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = -at(L::LAMP + i, k);
#pragma unroll
        for (int a = 0; a < NX; a++) s += A[a][i] * lp[a];
        ln[i] = s;
      }
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) at(L::LAMP + i, k) = ln[i];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) lp[i] = ln[i];
    }
    __syncwarp();
  }

  // slack steps, new row multipliers, step sizes (fraction to boundary), barrier slope
  __device__ void slack_and_steps(double mu, double dw, double tau, double &a_pr, double &a_du, double &gd_out) {
    double ap = 1.0, ad = 1.0, gd = 0.0;
#define MPCB_LOWER(gap, dv, z) do { double dz_ = -(z) + (mu - (z) * (dv)) / (gap); \
    if ((dv) < 0) ap = fmin(ap, -tau * (gap) / (dv)); if (dz_ < 0) ad = fmin(ad, -tau * (z) / dz_); } while (0)
#define MPCB_UPPER(gap, dv, z) do { double dz_ = -(z) + (mu + (z) * (dv)) / (gap); \
    if ((dv) > 0) ap = fmin(ap, tau * (gap) / (dv)); if (dz_ < 0) ad = fmin(ad, -tau * (z) / dz_); } while (0)
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], dx[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        dx[i] = at(L::DX + i, k);
        double g = (k < N) ? sigma * 2 * p.Q[i] * (xk[i] - xs[i]) : 0.0;
        gd += g * dx[i];
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = Mdl::bx(b);
        double gl = xk[i] - p.x_lo[i], gh = p.x_hi[i] - xk[i];
        gd += (-mu / gl + mu / gh) * dx[i];
        MPCB_LOWER(gl, dx[i], at(L::ZLX + b, k));
        MPCB_UPPER(gh, dx[i], at(L::ZUX + b, k));
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k), du = at(L::DU + i, k);
          double um = k > 0 ? at(L::U + i, k - 1) : 0.0;
          double up = k + 1 <= N - 1 ? at(L::U + i, k + 1) : 0.0;
          double gl = uk - p.u_lo[i], gh = p.u_hi[i] - uk;
          gd += (sigma * grad_u(k, i, uk, um, up) - mu / gl + mu / gh) * du;
          MPCB_LOWER(gl, du, at(L::ZLU + i, k));
          MPCB_UPPER(gh, du, at(L::ZUU + i, k));
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          double s = at(L::SR + r, k);
          double gl = s - p.rate_lo[r], gh = p.rate_hi[r] - s;
          double vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          double D = vl / gl + vu / gh + dw;
          double gs = -mu / gl + mu / gh;
          double res = at(L::U + ci, k) - at(L::U + ci, k - 1) - s;
          double ds = at(L::DU + ci, k) - at(L::DU + ci, k - 1) + res;
          at(L::DSR + r, k) = ds;
          at(L::LRP + r, k) = D * ds + gs;
          gd += gs * ds;
          MPCB_LOWER(gl, ds, vl);
          MPCB_UPPER(gh, ds, vu);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double d, ox, oy, hxx, hxy, hyy;
          obs_grad(k, j, xk[0], xk[1], d, ox, oy, hxx, hxy, hyy);
          double s = at(L::SO + j, k), gap = s - p.obs_lo, vl = at(L::VLO + j, k);
          double D = vl / gap + dw;
          double gs = -mu / gap + MPCB_KAPPA_D * mu;
          double ds = ox * dx[0] + oy * dx[1] + (d - s);
          at(L::DSO + j, k) = ds;
          at(L::LOP + j, k) = D * ds + gs;
          gd += gs * ds;
          MPCB_LOWER(gap, ds, vl);
        }
      }
    }
    a_pr = warp_min(ap);
    a_du = warp_min(ad);
    gd_out = warp_sum(gd);
    __syncwarp();
  }

  // accept the step: primal a, duals a_du (IPOPT: equality multipliers move with a)
  __device__ void accept_step(double a, double ad, double mu) {
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double l = at(L::LAM + i, k);
        at(L::LAM + i, k) = l + a * (at(L::LAMP + i, k) - l);
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = Mdl::bx(b);
        double x = at(L::X + i, k), dx = at(L::DX + i, k);
        double gl = x - p.x_lo[i], gh = p.x_hi[i] - x;
        double zl = at(L::ZLX + b, k), zu = at(L::ZUX + b, k);
        double dzl = -zl + (mu - zl * dx) / gl, dzu = -zu + (mu + zu * dx) / gh;
        double xn = x + a * dx;
        at(L::ZLX + b, k) = clampz(zl + ad * dzl, mu, xn - p.x_lo[i]);
        at(L::ZUX + b, k) = clampz(zu + ad * dzu, mu, p.x_hi[i] - xn);
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::X + i, k) += a * at(L::DX + i, k);
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double u = at(L::U + i, k), du = at(L::DU + i, k);
          double gl = u - p.u_lo[i], gh = p.u_hi[i] - u;
          double zl = at(L::ZLU + i, k), zu = at(L::ZUU + i, k);
          double dzl = -zl + (mu - zl * du) / gl, dzu = -zu + (mu + zu * du) / gh;
          double un = u + a * du;
          at(L::ZLU + i, k) = clampz(zl + ad * dzl, mu, un - p.u_lo[i]);
          at(L::ZUU + i, k) = clampz(zu + ad * dzu, mu, p.u_hi[i] - un);
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k), ds = at(L::DSR + r, k);
          double gl = s - p.rate_lo[r], gh = p.rate_hi[r] - s;
          double vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          double dvl = -vl + (mu - vl * ds) / gl, dvu = -vu + (mu + vu * ds) / gh;
          double sn = s + a * ds;
          at(L::SR + r, k) = sn;
          at(L::VLR + r, k) = clampz(vl + ad * dvl, mu, sn - p.rate_lo[r]);
          at(L::VUR + r, k) = clampz(vu + ad * dvu, mu, p.rate_hi[r] - sn);
          double l = at(L::LR + r, k);
          at(L::LR + r, k) = l + a * (at(L::LRP + r, k) - l);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double s = at(L::SO + j, k), ds = at(L::DSO + j, k);
          double gap = s - p.obs_lo, vl = at(L::VLO + j, k);
          double dvl = -vl + (mu - vl * ds) / gap;
          double sn = s + a * ds;
          at(L::SO + j, k) = sn;
          at(L::VLO + j, k) = clampz(vl + ad * dvl, mu, sn - p.obs_lo);
          double l = at(L::LO + j, k);
          at(L::LO + j, k) = l + a * (at(L::LOP + j, k) - l);
        }
      }
    }
    __syncwarp();
    // the controls are read by neighbouring stages (rate rows), update them last
    for (int k = lane; k < N; k += 32) {
      at(L::U + 0, k) += a * at(L::DU + 0, k);
      at(L::U + 1, k) += a * at(L::DU + 1, k);
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- start point
  __device__ bool init_iterate(int b) {
    const int nv = 2 * N + NX * (N + 1);
    const double *zi = p.z_init ? p.z_init + (size_t)b * nv : nullptr;
    const double *ob = p.obs ? p.obs + (size_t)b * MO * (N + 1) * 6 : nullptr;
    // obstacle trajectory staged in shared memory (centre and 1/semi-axis^2 per step)
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int j = 0; j < MO; j++) {
        const double *o = ob + ((size_t)j * (N + 1) + k) * 6;
        double sx, sy;
        if (OBS_MODE == 1) {  // PKG/MPC_CBF_optimize_kin_pre.py:246-249
          sx = p.ego_hl + o[4] / 2 + p.safe_l;
          sy = p.ego_hw + o[5] / 2 + p.safe_w;
        } else {
          sx = p.dyn_sx;
          sy = p.dyn_sy;
        }
        at(L::OCX + j, k) = o[0];
        at(L::OCY + j, k) = o[1];
        at(L::ISX + j, k) = 1.0 / (sx * sx);
        at(L::ISY + j, k) = 1.0 / (sy * sy);
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          at(L::U + i, k) = push_in(zi ? zi[2 * k + i] : 0.0, p.u_lo[i], p.u_hi[i]);
          at(L::ZLU + i, k) = 1.0;
          at(L::ZUU + i, k) = 1.0;
        }
      }
      if (p.init_mode == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) at(L::X + i, k) = zi ? zi[2 * N + NX * k + i] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::LAM + i, k) = 0.0;
    }
    __syncwarp();
    if (p.init_mode == 1) {  // Euler roll-out of the guessed controls (PKG/MPC_CBF_optimize_kin.py:207)
      double x[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) x[i] = x0[i];
      for (int k = 0; k <= N; k++) {
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NX; i++) at(L::X + i, k) = x[i];
        }
        if (k < N) {
          double u[2] = {at(L::U + 0, k), at(L::U + 1, k)}, f[NX];
          Mdl::f(x, u, p, f);
#pragma unroll
          for (int i = 0; i < NX; i++) x[i] = x[i] + p.T * f[i];
        }
      }
      __syncwarp();
    }
    bool fin = true;
    double gmax = 0;
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int b2 = 0; b2 < NBX; b2++) {
        int i = Mdl::bx(b2);
        at(L::X + i, k) = push_in(at(L::X + i, k), p.x_lo[i], p.x_hi[i]);
        at(L::ZLX + b2, k) = 1.0;
        at(L::ZUX + b2, k) = 1.0;
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          at(L::SR + r, k) = push_in(at(L::U + ci, k) - at(L::U + ci, k - 1), p.rate_lo[r], p.rate_hi[r]);
          at(L::VLR + r, k) = 1.0;
          at(L::VUR + r, k) = 1.0;
          at(L::LR + r, k) = 0.0;
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double d = obs_val(k, j, at(L::X + 0, k), at(L::X + 1, k));
          if (!isfinite(d)) fin = false;
          at(L::SO + j, k) = push_lo(d, p.obs_lo);
          at(L::VLO + j, k) = 1.0;
          at(L::LO + j, k) = 0.0;
        }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k);
          double um = k > 0 ? at(L::U + i, k - 1) : 0.0;
          double up = k + 1 <= N - 1 ? at(L::U + i, k + 1) : 0.0;
          gmax = fmax(gmax, fabs(grad_u(k, i, uk, um, up)));
        }
#pragma unroll
        for (int i = 0; i < NX; i++) gmax = fmax(gmax, fabs(2 * p.Q[i] * (at(L::X + i, k) - xs[i])));
      }
    }
    gmax = warp_max(gmax);
    sigma = gmax > MPCB_OBJ_SCALE_MAX_GRAD ? MPCB_OBJ_SCALE_MAX_GRAD / gmax : 1.0;
    if (sigma < 1e-8) sigma = 1e-8;
    __syncwarp();
    return __all_sync(0xffffffffu, fin);
  }

  // ---------------------------------------------------------------- main loop
  __device__ void run(int b) {
#pragma unroll
    for (int i = 0; i < NX; i++) {
      x0[i] = p.x0[(size_t)b * NX + i];
      xs[i] = p.xs[(size_t)b * NX + i];
    }
    double mu = p.mu_init, tau = fmax(MPCB_TAU_MIN, 1 - mu);
    const double tol = p.tol;
    int status = 2, it = 0;
    double dw_last = 0.0;
    double theta = 0, fobj = 0, bar = 0, lin = 0, phi = 0;
    // filter entries distributed over the lanes' registers
    double ft[MPCB_FILTER_SLOTS], fp[MPCB_FILTER_SLOTS];
    int nfilt = 0;
#pragma unroll
    for (int s = 0; s < MPCB_FILTER_SLOTS; s++) { ft[s] = INFINITY; fp[s] = INFINITY; }

    bool okinit = init_iterate(b);
    if (okinit) eval_primal<false>(0.0, theta, fobj, bar, lin);
    if (!okinit || !isfinite(theta) || !isfinite(bar)) {
      status = 4;
    } else {
      __syncwarp();
      const double theta_min = 1e-4 * fmax(1.0, theta), theta_max = 1e4 * fmax(1.0, theta);
      for (;;) {
        Kkt kk;
        eval_lin_kkt(kk);
        double co0;
        double err0 = kkt_error(kk, 0.0, co0);
        if (err0 <= tol && kk.dual <= MPCB_DUAL_INF_TOL && kk.prim <= MPCB_CONSTR_VIOL_TOL && co0 <= MPCB_COMPL_INF_TOL) { status = 0; break; }
        if (it >= p.max_iter) { status = 2; break; }
        {
          double co;
          while (kkt_error(kk, mu, co) <= MPCB_KAPPA_EPS * mu && mu > tol / 10) {
            mu = fmax(tol / 10, fmin(MPCB_KAPPA_MU * mu, pow(mu, MPCB_THETA_MU)));
            tau = fmax(MPCB_TAU_MIN, 1 - mu);
            nfilt = 0;
#pragma unroll
            for (int s = 0; s < MPCB_FILTER_SLOTS; s++) { ft[s] = INFINITY; fp[s] = INFINITY; }
          }
        }
        phi = sigma * fobj - mu * bar + MPCB_KAPPA_D * mu * lin;
        double dw = 0.0;
        build_qp(mu, 0.0);
        bool ok = riccati_backward();
        if (!ok) {
          dw = dw_last == 0.0 ? MPCB_DW_FIRST : fmax(MPCB_DW_MIN, MPCB_KW_MINUS * dw_last);
          for (;;) {
            build_qp(mu, dw);
            ok = riccati_backward();
            if (ok) break;
            dw *= dw_last == 0.0 ? MPCB_KW_PLUS_FIRST : MPCB_KW_PLUS;
            if (dw > MPCB_DW_MAX) break;
          }
          if (!ok) { status = 3; break; }
          dw_last = dw;
        }
        riccati_forward();
        adjoint();
        double a_max, a_dual, gd;
        slack_and_steps(mu, dw, tau, a_max, a_dual, gd);
        double a_min;
        if (gd < 0 && theta <= theta_min) {
          a_min = MPCB_GAMMA_THETA;
          if (theta > 0) {
            a_min = fmin(a_min, MPCB_GAMMA_PHI * theta / (-gd));
            a_min = fmin(a_min, pow(theta, MPCB_S_THETA) / pow(-gd, MPCB_S_PHI));
          }
        } else if (gd < 0) {
          a_min = fmin(MPCB_GAMMA_THETA, MPCB_GAMMA_PHI * theta / (-gd));
        } else {
          a_min = MPCB_GAMMA_THETA;
        }
        a_min = fmax(MPCB_GAMMA_ALPHA * a_min, 1e-14);
        double a = a_max;
        bool accepted = false, armijo = false;
        double th_t = 0, f_t = 0, bar_t = 0, lin_t = 0;
        while (a >= a_min) {
          eval_primal<true>(a, th_t, f_t, bar_t, lin_t);
          double ph_t = sigma * f_t - mu * bar_t + MPCB_KAPPA_D * mu * lin_t;
          bool fin = isfinite(th_t) && isfinite(ph_t);
          bool blocked = th_t >= theta_max;
#pragma unroll
          for (int s = 0; s < MPCB_FILTER_SLOTS; s++) blocked = blocked || (th_t >= ft[s] && ph_t >= fp[s]);
          blocked = __any_sync(0xffffffffu, blocked);
          if (fin && !blocked) {
            bool sw = gd < 0 && a * pow(-gd, MPCB_S_PHI) > pow(theta, MPCB_S_THETA);
            if (theta <= theta_min && sw) {
              if (ph_t <= phi + MPCB_ETA_PHI * a * gd + 10 * MPCB_DBL_EPS * fabs(phi)) { accepted = true; armijo = true; }
            } else if (th_t <= (1 - MPCB_GAMMA_THETA) * theta || ph_t <= phi - MPCB_GAMMA_PHI * theta + 10 * MPCB_DBL_EPS * fabs(phi)) {
              accepted = true;
            }
          }
          if (accepted) break;
          a *= 0.5;
        }
        if (!accepted) { status = 3; break; }
        if (!armijo) {
          int slot = nfilt % (32 * MPCB_FILTER_SLOTS);
          if ((slot & 31) == lane) {
#pragma unroll
            for (int s = 0; s < MPCB_FILTER_SLOTS; s++)
              if (s == (slot >> 5)) { ft[s] = (1 - MPCB_GAMMA_THETA) * theta; fp[s] = phi - MPCB_GAMMA_PHI * theta; }
          }
          nfilt++;
        }
        accept_step(a, a_dual, mu);
        eval_primal<false>(0.0, theta, fobj, bar, lin);
        __syncwarp();
        it++;
      }
    }
    // ---- results
    const int nv = 2 * N + NX * (N + 1);
    if (lane == 0) {
      p.u0[2 * (size_t)b + 0] = at(L::U + 0, 0);
      p.u0[2 * (size_t)b + 1] = at(L::U + 1, 0);
      p.cost[b] = status == 4 ? nan("") : fobj;
      p.status[b] = status;
      p.iters[b] = it;
    }
    if (p.z_out) {
      double *z = p.z_out + (size_t)b * nv;
      for (int idx = lane; idx < 2 * N; idx += 32) z[idx] = at(L::U + (idx & 1), idx >> 1);
      for (int idx = lane; idx < NX * (N + 1); idx += 32) z[2 * N + idx] = at(L::X + (idx % NX), idx / NX);
    }
    if (p.lam_out) {
      double *l = p.lam_out + (size_t)b * NX * (N + 1);
      for (int idx = lane; idx < NX * (N + 1); idx += 32) l[idx] = at(L::LAM + (idx % NX), idx / NX) / sigma;
    }
  }
};

template <class Mdl, int NR, int MO, int OBS_MODE>
__global__ void __launch_bounds__(32) solve_kernel(const __grid_constant__ KParams p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    Solver<Mdl, NR, MO, OBS_MODE> s(p, smem, lane);
    s.run(b);
    __syncwarp();
  }
}

}  // namespace mpcb
