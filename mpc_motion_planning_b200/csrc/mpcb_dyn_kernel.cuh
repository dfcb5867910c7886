// mpcb200 device code for the dynamic bicycle NLP (PKG/MPC_CBF_optimize_dyn.py:137-250):
// x = [x, y, phi, vx, vy, r], u = [df, ax], saturating tire forces, both control-rate rows,
// obstacle row sqrt(ellipse - 1) >= 1 at stages 0..N.  Same interior-point algorithm and
// warp-per-scenario organisation as the kinematic kernel (mpcb_kernel.cuh); the Riccati sweep
// is the generic dense one (nx = 6) and the model derivatives come from dyn_model.cuh.
#pragma once
#include "mpcb_kernel.cuh"
#include "dyn_model.cuh"
#include "dyn_riccati.cuh"

namespace mpcb {

struct DynLayout {
  static constexpr int NX = 6, NBX = 3, NR = 2, MO = 1, NJ = DYN_NJ;
  static constexpr int NR_ROWS = NR, MO_ROWS = MO;
  static constexpr int FTO = 0, FPO = 0;  // (names of the shared run loop's restoration branches, unused here)
  // ---- shared memory: working set of the serial sweeps (stage-major records, see KinLayout)
  static constexpr int CDEF = 0;
  static constexpr int LAMP = CDEF;  // alias
  static constexpr int JAC = CDEF + NX;
  // structurally nonzero entries of the stage Hessian only: the (x,y) block (obstacle row + cost),
  // the (phi,vx,vy,r) block (model + cost) - 3 + 10 of the 21 - and d2L/(d delta d{vx,vy,r})
  static constexpr int HXX = JAC + NJ;
  static constexpr int HUX = HXX + DYN_NHX;
  static constexpr int GX = HUX + DYN_NHU;
  static constexpr int R18 = GX + NX;      // [HUU(2) EE(2) GU(2) TK(2) ...] -> gains [KX(12) KW(4) KK(2)] -> slack steps
  static constexpr int HUU = R18, EE = R18 + 2, GU = R18 + 4, TK = R18 + 6;
  static constexpr int KX = R18, KW = R18 + 12, KK = R18 + 16;
  static constexpr int DSR = R18, LRP = DSR + NR, DSO = LRP + NR, LOP = DSO + MO;
  static constexpr int CDEFT = R18 + 8;  // defects of the trial point (6 slots)
  static constexpr int NSH = R18 + 18;
  // 63 doubles per stage = 25.7 KB per scenario at N = 50: two blocks of four warps per SM
  static constexpr int NF = NSH | 1;
  // ---- global-memory slab: the primal-dual iterate and the obstacle centre (stage-parallel phases only)
  static constexpr int G0 = 96;
  static_assert(NF <= G0, "field id spaces overlap");
  static constexpr int X = G0;
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
  static constexpr int OCX = LO + MO;
  static constexpr int OCY = OCX + MO;
  // the step: written by lane 0 in the forward sweep, read by the stage-parallel phases only
  static constexpr int DX = OCY + MO;
  static constexpr int DU = DX + NX;
  // rows-as-shipped mode only: Huu diagonal and gu of the stage (the shared copies are overwritten by the
  // gains before the multipliers of the tied controls are recovered), and that stationarity residual
  static constexpr int HUG = DU + 2;
  static constexpr int GUG = HUG + 2;
  static constexpr int RUG = GUG + 2;
  static constexpr int NG = RUG + 2 - G0;
  static constexpr int ER = R18 + 8;  // rows-as-shipped: e_k = U_k - U_{k-1} for the backward sweep
  static constexpr int SG = 132;
  __host__ __device__ static constexpr size_t bytes(int N) { return sizeof(double) * (size_t)NF * (size_t)(N + 1); }
  __host__ __device__ static constexpr size_t slab_doubles() { return (size_t)NG * SG; }
};

__device__ __forceinline__ int pidx6(int i, int j) { return i * 6 - i * (i - 1) / 2 + (j - i); }  // i <= j
__device__ __forceinline__ int sidx6(int i, int j) { return i <= j ? pidx6(i, j) : pidx6(j, i); }
// slot of Hxx(i,j), i <= j, in the compact storage (order of dyn_riccati_F's hx argument); -1 = structural zero
__device__ __forceinline__ constexpr int hslot(int i, int j) {
  return j < 2 ? i + j : (i < 2 ? -1 : 3 + (i - 2) * 4 - (i - 2) * (i - 3) / 2 + (j - i));
}
static_assert(hslot(0, 0) == 0 && hslot(0, 1) == 1 && hslot(1, 1) == 2 && hslot(2, 2) == 3 && hslot(2, 5) == 6 && hslot(3, 3) == 7 &&
              hslot(4, 4) == 10 && hslot(5, 5) == 12 && hslot(1, 2) == -1, "compact Hessian order");

// SHP: pair the rows of g with the bound lists exactly as PKG/MPC_CBF_optimize_dyn.py:112-133 ships them
// (SURVEY.md section 0.4).  Then, for k = 1..N-1, the slack SR(r,k) with the rate bounds relaxes component
// r = 0,1 (x, y) of the defect into stage k+1 (its row multiplier is LAM(r,k+1)) and the rate rows
// U_k - U_{k-1} are equalities with multipliers LR(r,k): one control pair for the whole horizon.
template <bool SHP>
struct DynSolver {
  using L = DynLayout;
  static constexpr int NX = 6, NBX = 3, NR = 2, MO = 1, NJ = DYN_NJ;
  static constexpr bool RS = false;  // no restoration-capable sibling for this family: a failed line search ends with status 3
  static constexpr bool ROWS_INTERLEAVED = true;  // g = init, then per stage: defect, rate rows; obstacle rows last
  __device__ static __forceinline__ constexpr int bx(int i) { return i == 0 ? 1 : (i == 1 ? 3 : 4); }  // y, vx, vy

  const KParams &p;
  double *gs;  // this warp's slab in global memory
  int woff;    // this warp's offset into the block's shared memory (in doubles)
  int &tick;
  int N, lane;
  double sigma;
  double x0[NX], xs[NX];

  // names the shared run loop mentions inside its `if constexpr (RS)` branches (never executed here)
  bool resto = false;
  double zeta = 0.0, rmu = 0.0, o_thr = 0.0, o_f = 0.0;
  __device__ __forceinline__ void resto_switch(bool) {}
  __device__ __forceinline__ void resto_multipliers(double) {}

  __device__ DynSolver(const KParams &p_, double *gs_, int woff_, int &tick_, int lane_)
      : p(p_), gs(gs_), woff(woff_), tick(tick_), N(p_.N), lane(lane_) {}

  __device__ __forceinline__ double &at(int field, int k) {
    return field >= L::G0 ? gs[(field - L::G0) * L::SG + k] : g_smem[woff + k * L::NF + field];
  }
  __device__ __forceinline__ bool has_rate(int k) const { return k >= 1 && k <= N - 1; }

  __device__ __forceinline__ double grad_u(int k, int i, double uk, double ukm1, double ukp1) const {
    double v = 2 * p.R[i] * uk;                    // PKG/MPC_CBF_optimize_dyn.py:218-225
    if (k > 0) v += 2 * p.DR[i] * (uk - ukm1);
    else if (p.du0_cost) v += 2 * p.DR[i] * uk;
    if (k + 1 <= N - 1) v -= 2 * p.DR[i] * (ukp1 - uk);
    return v;
  }

  // obstacle row d = sqrt(e), e = (x-ox)^2/sX^2 + (y-oy)^2/sY^2 - 1   (PKG/..._dyn.py:238-243)
  __device__ __forceinline__ void obs_row(int k, double px, double py, double &d, double &gx, double &gy, double &hxx, double &hxy,
                                          double &hyy) {
    double dx = px - at(L::OCX, k), dy = py - at(L::OCY, k);
    double a = 1.0 / (p.dyn_sx * p.dyn_sx), b = 1.0 / (p.dyn_sy * p.dyn_sy);  // exact for the reference's 4 and 1
    double e = dx * dx * a + dy * dy * b - 1.0;
    double q = e > 0.0 ? sqrt(e) : nan("");
    double ex = 2 * dx * a, ey = 2 * dy * b;
    d = q;
    const double rq = fast_rcp(q);  // q > 0 or NaN (inside the ellipse), which propagates
    gx = 0.5 * ex * rq;
    gy = 0.5 * ey * rq;
    double q3 = 4 * q * q * q;
    const double rq3 = fast_rcp(q3);
    hxx = a * rq - ex * ex * rq3;
    hxy = -ex * ey * rq3;
    hyy = b * rq - ey * ey * rq3;
  }

  __device__ __forceinline__ void eval_point(double alpha, bool fresh, double &theta, double &fobj, double &bar, double &lin) {
    const int cdst = fresh ? L::CDEF : L::CDEFT;
    double th = 0, fo = 0, br = 0, ln = 0;
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], uk[2] = {0, 0};
      double gp = 1.0;
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k) + alpha * at(L::DX + i, k);
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double c0 = xk[i] - x0[i];
          th += fabs(c0);
          at(cdst + i, 0) = c0;
        }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) uk[i] = at(L::U + i, k) + alpha * at(L::DU + i, k);
        double f[NX];
        {
          double J[NJ];
          dyn_fjac(xk, uk, p, f, J);
#pragma unroll
          for (int i = 0; i < NJ; i++) at(L::JAC + i, k) = J[i];
        }
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double xn = at(L::X + i, k + 1) + alpha * at(L::DX + i, k + 1);
          double d = xn - (xk[i] + p.T * f[i]);
          if (SHP && i < 2 && has_rate(k)) d -= at(L::SR + i, k) + (alpha != 0.0 ? alpha * at(L::DSR + i, k) : 0.0);
          th += fabs(d);
          at(cdst + i, k + 1) = d;
          double e = xk[i] - xs[i];
          fo += p.Q[i] * e * e;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
          gp *= (uk[i] - p.u_lo[i]) * (p.u_hi[i] - uk[i]);
          fo += p.R[i] * uk[i] * uk[i];
          if (k > 0) {
            double um = at(L::U + i, k - 1) + alpha * at(L::DU + i, k - 1);
            double e = uk[i] - um;
            fo += p.DR[i] * e * e;
          } else if (p.du0_cost) {
            fo += p.DR[i] * uk[i] * uk[i];
          }
        }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = bx(b);
        gp *= (xk[i] - p.x_lo[i]) * (p.x_hi[i] - xk[i]);
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double um = at(L::U + r, k - 1) + alpha * at(L::DU + r, k - 1);
          double s = at(L::SR + r, k) + (alpha != 0.0 ? alpha * at(L::DSR + r, k) : 0.0);
          th += fabs(SHP ? uk[r] - um : uk[r] - um - s);
          gp *= (s - p.rate_lo[r]) * (p.rate_hi[r] - s);
        }
      }
      {
        double dx = xk[0] - at(L::OCX, k), dy = xk[1] - at(L::OCY, k);
        double e = dx * dx / (p.dyn_sx * p.dyn_sx) + dy * dy / (p.dyn_sy * p.dyn_sy) - 1.0;
        double d = e > 0.0 ? sqrt(e) : nan("");
        double s = at(L::SO, k) + (alpha != 0.0 ? alpha * at(L::DSO, k) : 0.0);
        th += fabs(d - s);
        gp *= s - p.obs_lo;
        ln += s - p.obs_lo;
      }
      br += d_log(gp);
    }
    theta = warp_sum(th);
    fobj = warp_sum(fo);
    bar = warp_sum(br);
    lin = warp_sum(ln);
    __syncwarp();
  }

  struct Kkt { double dual, prim, cmin, cmax, sum_lam, sum_z; };

  __device__ __forceinline__ void kkt_pieces(Kkt &o) {
    double dual = 0, prim = 0, cmin = INFINITY, cmax = -INFINITY, sl = 0, sz = 0;
#define MPCB_COMPL(gap, mult) do { double p_ = (gap) * (mult); cmin = fmin(cmin, p_); cmax = fmax(cmax, p_); sz += (mult); } while (0)
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], rx[NX], l1[NX], A[NX][NX], B[NX][2];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        rx[i] = at(L::LAM + i, k);
        l1[i] = 0;
        prim = fmax(prim, fabs(at(L::CDEF + i, k)));
        sl += fabs(rx[i]);
      }
      if (k < N) {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        dyn_expand(J, p.T, A, B);
#pragma unroll
        for (int i = 0; i < NX; i++) l1[i] = at(L::LAM + i, k + 1);
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double r = sigma * 2 * p.Q[i] * (xk[i] - xs[i]);
#pragma unroll
          for (int a = 0; a < NX; a++) r -= A[a][i] * l1[a];
          rx[i] += r;
        }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = bx(b);
        double zl = at(L::ZLX + b, k), zu = at(L::ZUX + b, k);
        rx[i] += -zl + zu;
        MPCB_COMPL(xk[i] - p.x_lo[i], zl);
        MPCB_COMPL(p.x_hi[i] - xk[i], zu);
      }
      {
        double d, gx, gy, hxx, hxy, hyy;
        obs_row(k, xk[0], xk[1], d, gx, gy, hxx, hxy, hyy);
        double lo = at(L::LO, k), vl = at(L::VLO, k), s = at(L::SO, k);
        rx[0] += lo * gx;
        rx[1] += lo * gy;
        dual = fmax(dual, fabs(-lo - vl));
        prim = fmax(prim, fabs(d - s));
        MPCB_COMPL(s - p.obs_lo, vl);
        sl += fabs(lo);
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
          for (int a = 0; a < NX; a++) r -= B[a][i] * l1[a];
          if (has_rate(k)) r += at(L::LR + i, k);
          if (has_rate(k + 1)) r -= at(L::LR + i, k + 1);
          dual = fmax(dual, fabs(r));
          MPCB_COMPL(uk - p.u_lo[i], zl);
          MPCB_COMPL(p.u_hi[i] - uk, zu);
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k), vl = at(L::VLR + r, k), vu = at(L::VUR + r, k), lr = at(L::LR + r, k);
          dual = fmax(dual, fabs(-(SHP ? at(L::LAM + r, k + 1) : lr) - vl + vu));
          prim = fmax(prim, fabs(SHP ? at(L::U + r, k) - at(L::U + r, k - 1) : at(L::U + r, k) - at(L::U + r, k - 1) - s));
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
    // 1/s_d and 1/s_c of IPOPT's error scaling; the denominators are positive, fast_rcp is exact to ~1 ulp
    double rs_d = MPCB_S_MAX * fast_rcp(fmax(MPCB_S_MAX, (o.sum_lam + o.sum_z) * fast_rcp(fmax(1.0, (double)(p.n_eq + p.n_bm)))));
    double rs_c = MPCB_S_MAX * fast_rcp(fmax(MPCB_S_MAX, o.sum_z * fast_rcp(fmax(1.0, (double)p.n_bm))));
    return fmax(fmax(o.dual * rs_d, o.prim), co * rs_c);
  }

  __device__ __forceinline__ void build_qp(double mu, double dw) {
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], uk[2] = {0, 0}, gx[NX];
      double Hxx[NX][NX];
      double hud[NX];  // d2L / (d delta d x_j)
      double hdd_f = 0;
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        hud[i] = 0;
        gx[i] = 0;
#pragma unroll
        for (int j = 0; j < NX; j++) Hxx[i][j] = 0;
        Hxx[i][i] = dw;
      }
      if (k < N) {
        uk[0] = at(L::U + 0, k);
        uk[1] = at(L::U + 1, k);
        double l1[NX], H[DYN_NH];
#pragma unroll
        for (int i = 0; i < NX; i++) l1[i] = at(L::LAM + i, k + 1);
        dyn_hess(xk, uk, p, l1, H);
        // packed upper triangle over (phi, vx, vy, r, df) = state indices 2..5 and the steering input
        int q = 0;
#pragma unroll
        for (int a = 0; a < 5; a++)
#pragma unroll
          for (int b = a; b < 5; b++) {
            double v = -p.T * H[q++];
            if (b < 4) { Hxx[2 + a][2 + b] += v; if (a != b) Hxx[2 + b][2 + a] += v; }
            else if (a < 4) hud[2 + a] += v;
            else hdd_f = v;
          }
#pragma unroll
        for (int i = 0; i < NX; i++) {
          Hxx[i][i] += sigma * 2 * p.Q[i];
          gx[i] = sigma * 2 * p.Q[i] * (xk[i] - xs[i]);
        }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = bx(b);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        Hxx[i][i] += at(L::ZLX + b, k) * rl + at(L::ZUX + b, k) * rh;
        gx[i] += mu * (rh - rl);
      }
      {
        double d, ox, oy, hxx, hxy, hyy;
        obs_row(k, xk[0], xk[1], d, ox, oy, hxx, hxy, hyy);
        double s = at(L::SO, k), rg = fast_rcp(s - p.obs_lo);
        double D = at(L::VLO, k) * rg + dw;
        double gs = -mu * rg + MPCB_KAPPA_D * mu;
        double lo = at(L::LO, k);
        double t = D * (d - s) + gs;
        Hxx[0][0] += lo * hxx + D * ox * ox;
        Hxx[0][1] += lo * hxy + D * ox * oy;
        Hxx[1][0] += lo * hxy + D * ox * oy;
        Hxx[1][1] += lo * hyy + D * oy * oy;
        gx[0] += ox * t;
        gx[1] += oy * t;
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        at(L::GX + i, k) = gx[i];
        if (i >= 3) at(L::HUX + i - 3, k) = hud[i];
#pragma unroll
        for (int j = i; j < NX; j++)
          if (hslot(i, j) >= 0) at(L::HXX + hslot(i, j), k) = Hxx[i][j];
      }
      if (k < N) {
        double E[2] = {0, 0}, t[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double g = sigma * 2 * p.R[i] * uk[i];
          double hd = sigma * 2 * p.R[i] + dw + (i == 0 ? hdd_f : 0.0);
          if (k == 0 && p.du0_cost) {
            hd += sigma * 2 * p.DR[i];
            g += sigma * 2 * p.DR[i] * uk[i];
          }
          double rl = fast_rcp(uk[i] - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk[i]);
          hd += at(L::ZLU + i, k) * rl + at(L::ZUU + i, k) * rh;
          g += mu * (rh - rl);
          at(L::HUU + i, k) = hd;
          at(L::GU + i, k) = g;
          if (SHP) {
            at(L::HUG + i, k) = hd;
            at(L::GUG + i, k) = g;
          } else if (k >= 1) {
            E[i] = sigma * 2 * p.DR[i];
            t[i] = sigma * 2 * p.DR[i] * (uk[i] - at(L::U + i, k - 1));
          }
        }
        if (has_rate(k)) {
#pragma unroll
          for (int r = 0; r < NR; r++) {
            double s = at(L::SR + r, k);
            double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
            double D = at(L::VLR + r, k) * rl + at(L::VUR + r, k) * rh + dw;
            double gs = mu * (rh - rl);
            if (SHP) {  // curvature and gradient of the relaxation slack (the stage input), and e_k
              E[r] = D;
              t[r] = gs;
              at(L::ER + r, k) = uk[r] - at(L::U + r, k - 1);
            } else {
              double res = uk[r] - at(L::U + r, k - 1) - s;
              E[r] += D;
              t[r] += D * res + gs;
            }
          }
        }
        at(L::EE + 0, k) = E[0];
        at(L::EE + 1, k) = E[1];
        at(L::TK + 0, k) = t[0];
        at(L::TK + 1, k) = t[1];
      }
    }
    __syncwarp();
  }

  // Riccati on the state augmented with the previous control (see KinSolver).  The first half of
  // a stage (F = H + [A B]' V [A B], ~300 FMAs) is generated for the sparsity of this model's
  // A and B (dyn_riccati.cuh); the pivot, gains and value-function update are generic.
  __device__ __forceinline__ bool riccati_backward() {
    double P[21], W[12], Q[3] = {0, 0, 0}, px[NX], pw[2] = {0, 0};
#pragma unroll
    for (int q = 0; q < 21; q++) P[q] = 0.0;
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
      for (int j = i; j < NX; j++)
        if (hslot(i, j) >= 0) P[pidx6(i, j)] = at(L::HXX + hslot(i, j), N);
#pragma unroll
    for (int i = 0; i < NX; i++) {
      px[i] = at(L::GX + i, N);
      W[2 * i] = W[2 * i + 1] = 0;
    }
    bool ok = true;
#pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      double a_[NJ], hx[DYN_NHX], hu[DYN_NHU], gx[NX], b[NX];
#pragma unroll
      for (int i = 0; i < NJ; i++) a_[i] = p.T * at(L::JAC + i, k);
      // structurally nonzero Hessian entries: (x,y) block, (phi..r) block, steering row vs (vx,vy,r)
#pragma unroll
      for (int q = 0; q < DYN_NHX; q++) hx[q] = at(L::HXX + q, k);
#pragma unroll
      for (int q = 0; q < DYN_NHU; q++) hu[q] = at(L::HUX + q, k);
#pragma unroll
      for (int i = 0; i < NX; i++) {
        gx[i] = at(L::GX + i, k);
        b[i] = -at(L::CDEF + i, k + 1);
      }
      const double Ed = at(L::EE + 0, k), Ea = at(L::EE + 1, k), td = at(L::TK + 0, k), ta = at(L::TK + 1, k);
      double Fxx[21], Fux[12], Fuu[3], fx[NX], fu[2];
      const bool tied = SHP && k >= 1;
      dyn_riccati_F(P, W, Q, px, pw, a_, p.T, hx, hu, at(L::HUU + 0, k), at(L::HUU + 1, k), tied ? 0.0 : Ed, tied ? 0.0 : Ea, tied ? 0.0 : td,
                    tied ? 0.0 : ta, gx, at(L::GU + 0, k), at(L::GU + 1, k), b, Fxx, Fux, Fuu, fx, fu);
      if (SHP && k >= 1) {
        // Rows as shipped: the control of this stage is tied to the previous one (dU_k = dw - e_k) and the
        // free inputs are the two slacks that relax the x/y defects, entering x_{k+1} rows 0,1.  F above is the
        // quadratic over (dx, dU) (EE/TK hold the slack curvature and gradient here, so E = t = 0 went in);
        // eliminate the slack inputs and rename dU -> dw.
        const double e0 = at(L::ER + 0, k), e1 = at(L::ER + 1, k);
        double A[NX][NX], B[NX][2];
        {
          double J[NJ];
#pragma unroll
          for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
          dyn_expand(J, p.T, A, B);
        }
        double M2[2][NX], Cm[2][2], c0[2];
#pragma unroll
        for (int a = 0; a < 2; a++) {
#pragma unroll
          for (int j = 0; j < NX; j++) {
            double v = 0;
#pragma unroll
            for (int i = 0; i < NX; i++) v += P[sidx6(a, i)] * A[i][j];
            M2[a][j] = v;
          }
          double pb = px[a];
#pragma unroll
          for (int i = 0; i < NX; i++) pb += P[sidx6(a, i)] * b[i];
#pragma unroll
          for (int j = 0; j < 2; j++) {
            double v = W[2 * a + j];
#pragma unroll
            for (int i = 0; i < NX; i++) v += P[sidx6(a, i)] * B[i][j];
            Cm[a][j] = v;
          }
          c0[a] = pb + (a == 0 ? td : ta) - (Cm[a][0] * e0 + Cm[a][1] * e1);
        }
        const double h00 = P[pidx6(0, 0)] + Ed, h01 = P[pidx6(0, 1)], h11 = P[pidx6(1, 1)] + Ea;
        const double dets = h00 * h11 - h01 * h01;
        if (!(h00 > 0.0) || !(dets > 0.0) || !isfinite(dets)) { ok = false; break; }
        const double ids = fast_rcp(dets);
        const double i00 = h11 * ids, i01 = -h01 * ids, i11 = h00 * ids;
        double Ks[2][NX], Kw2[2][2], ks[2];
#pragma unroll
        for (int j = 0; j < NX; j++) {
          Ks[0][j] = -(i00 * M2[0][j] + i01 * M2[1][j]);
          Ks[1][j] = -(i01 * M2[0][j] + i11 * M2[1][j]);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
          Kw2[0][j] = -(i00 * Cm[0][j] + i01 * Cm[1][j]);
          Kw2[1][j] = -(i01 * Cm[0][j] + i11 * Cm[1][j]);
        }
        ks[0] = -(i00 * c0[0] + i01 * c0[1]);
        ks[1] = -(i01 * c0[0] + i11 * c0[1]);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < NX; j++) {
            at(L::KX + j, k) = Ks[0][j];
            at(L::KX + NX + j, k) = Ks[1][j];
          }
          at(L::KW + 0, k) = Kw2[0][0]; at(L::KW + 1, k) = Kw2[0][1]; at(L::KW + 2, k) = Kw2[1][0]; at(L::KW + 3, k) = Kw2[1][1];
          at(L::KK + 0, k) = ks[0]; at(L::KK + 1, k) = ks[1];
        }
        {
          int q = 0;
#pragma unroll
          for (int i = 0; i < NX; i++)
#pragma unroll
            for (int j = i; j < NX; j++, q++) P[q] = Fxx[q] + 0.5 * ((M2[0][i] * Ks[0][j] + M2[1][i] * Ks[1][j]) + (M2[0][j] * Ks[0][i] + M2[1][j] * Ks[1][i]));
        }
#pragma unroll
        for (int i = 0; i < NX; i++) {
          W[2 * i] = Fux[i] + M2[0][i] * Kw2[0][0] + M2[1][i] * Kw2[1][0];
          W[2 * i + 1] = Fux[6 + i] + M2[0][i] * Kw2[0][1] + M2[1][i] * Kw2[1][1];
          px[i] = fx[i] - (Fux[i] * e0 + Fux[6 + i] * e1) + M2[0][i] * ks[0] + M2[1][i] * ks[1];
        }
        Q[0] = Fuu[0] + Cm[0][0] * Kw2[0][0] + Cm[1][0] * Kw2[1][0];
        Q[1] = Fuu[1] + 0.5 * ((Cm[0][0] * Kw2[0][1] + Cm[1][0] * Kw2[1][1]) + (Cm[0][1] * Kw2[0][0] + Cm[1][1] * Kw2[1][0]));
        Q[2] = Fuu[2] + Cm[0][1] * Kw2[0][1] + Cm[1][1] * Kw2[1][1];
        pw[0] = fu[0] - (Fuu[0] * e0 + Fuu[1] * e1) + Cm[0][0] * ks[0] + Cm[1][0] * ks[1];
        pw[1] = fu[1] - (Fuu[1] * e0 + Fuu[2] * e1) + Cm[0][1] * ks[0] + Cm[1][1] * ks[1];
        continue;
      }
      const double det = Fuu[0] * Fuu[2] - Fuu[1] * Fuu[1];
      if (!(Fuu[0] > 0.0) || !(det > 0.0) || !isfinite(det)) { ok = false; break; }
      const double id = fast_rcp(det);
      const double idd = Fuu[2] * id, ida = -Fuu[1] * id, iaa = Fuu[0] * id;
      double Kd[NX], Ka[NX];
#pragma unroll
      for (int j = 0; j < NX; j++) {
        Kd[j] = -(idd * Fux[j] + ida * Fux[6 + j]);
        Ka[j] = -(ida * Fux[j] + iaa * Fux[6 + j]);
      }
      const double wdd = idd * Ed, wda = ida * Ea, wad = ida * Ed, waa = iaa * Ea;
      const double kkd = -(idd * fu[0] + ida * fu[1]), kka = -(ida * fu[0] + iaa * fu[1]);
      __syncwarp();  // every lane has consumed the QP slots of this stage before lane 0 reuses them
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NX; j++) {
          at(L::KX + j, k) = Kd[j];
          at(L::KX + NX + j, k) = Ka[j];
        }
        at(L::KW + 0, k) = wdd; at(L::KW + 1, k) = wda; at(L::KW + 2, k) = wad; at(L::KW + 3, k) = waa;
        at(L::KK + 0, k) = kkd; at(L::KK + 1, k) = kka;
      }
      {
        int q = 0;
#pragma unroll
        for (int i = 0; i < NX; i++)
#pragma unroll
          for (int j = i; j < NX; j++, q++) P[q] = Fxx[q] + Fux[i] * Kd[j] + Fux[6 + i] * Ka[j];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        W[2 * i] = Fux[i] * wdd + Fux[6 + i] * wad;
        W[2 * i + 1] = Fux[i] * wda + Fux[6 + i] * waa;
        px[i] = fx[i] + Fux[i] * kkd + Fux[6 + i] * kka;
      }
      Q[0] = Ed - Ed * wdd;
      Q[1] = -0.5 * (Ed * wda + Ea * wad);
      Q[2] = Ea - Ea * waa;
      pw[0] = -td - Ed * kkd;
      pw[1] = -ta - Ea * kka;
    }
    __syncwarp();
    return ok;
  }

  __device__ __forceinline__ void riccati_forward() {
    double dx[NX], dum[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < NX; i++) dx[i] = -at(L::CDEF + i, 0);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::DX + i, 0) = dx[i];
    }
#pragma unroll 1
    for (int k = 0; k < N; k++) {
      double A[NX][NX], B[NX][2], du[2];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        dyn_expand(J, p.T, A, B);
      }
      double ds[2] = {0, 0};  // rows as shipped: the relaxation slacks' step (the stage input for k >= 1)
#pragma unroll
      for (int i = 0; i < 2; i++) {
        double s = at(L::KK + i, k);
#pragma unroll
        for (int j = 0; j < NX; j++) s += at(L::KX + i * NX + j, k) * dx[j];
        s += at(L::KW + i * 2 + 0, k) * dum[0] + at(L::KW + i * 2 + 1, k) * dum[1];
        if (SHP && k >= 1) {
          ds[i] = s;
          du[i] = dum[i] - (at(L::U + i, k) - at(L::U + i, k - 1));
        } else {
          du[i] = s;
        }
      }
      double dn[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = -at(L::CDEF + i, k + 1);
#pragma unroll
        for (int j = 0; j < NX; j++) s += A[i][j] * dx[j];
        s += B[i][0] * du[0] + B[i][1] * du[1];
        if (i < 2) s += ds[i];
        dn[i] = s;
      }
      if (SHP) __syncwarp();  // every lane has read the gains of this stage before lane 0 reuses two of their slots
      if (lane == 0) {
        if (SHP && k >= 1) { at(L::DSR + 0, k) = ds[0]; at(L::DSR + 1, k) = ds[1]; }
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
    if (lane == 0) { at(L::DU + 0, N) = 0.0; at(L::DU + 1, N) = 0.0; }
    __syncwarp();
  }

  __device__ __forceinline__ void adjoint() {
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double dx[NX], ud = k < N ? at(L::DU + 0, k) : 0.0, r[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) dx[i] = at(L::DX + i, k);
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double s = at(L::GX + i, k);
        if (i >= 3) s += at(L::HUX + i - 3, k) * ud;
#pragma unroll
        for (int j = 0; j < NX; j++) {
          const int q = i <= j ? hslot(i, j) : hslot(j, i);
          if (q >= 0) s += at(L::HXX + q, k) * dx[j];
        }
        r[i] = s;
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::LAMP + i, k) = r[i];
    }
    __syncwarp();
    double lp[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) lp[i] = -at(L::LAMP + i, N);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::LAMP + i, N) = lp[i];
    }
#pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      double A[NX][NX], B[NX][2], ln[NX];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        dyn_expand(J, p.T, A, B);
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

  // Rows as shipped: new multipliers of the tied-control equalities U_k - U_{k-1} = 0 from stationarity in U_k,
  // nu_k = nu_{k+1} - [Huu dU_k + Hux dx_k + gu - B' lam+_{k+1}]  (the DR cost terms cancel: dU_k - dU_{k-1} = -e_k).
  __device__ __forceinline__ void tied_control_multipliers() {
#pragma unroll 1
    for (int k = lane; k <= N - 1; k += 32) {
      if (k < 1) continue;
      double A[NX][NX], B[NX][2], dx[NX], l1[NX];
      {
        double J[NJ];
#pragma unroll
        for (int i = 0; i < NJ; i++) J[i] = at(L::JAC + i, k);
        dyn_expand(J, p.T, A, B);
      }
#pragma unroll
      for (int i = 0; i < NX; i++) { dx[i] = at(L::DX + i, k); l1[i] = at(L::LAMP + i, k + 1); }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        double r = at(L::GUG + i, k) + at(L::HUG + i, k) * at(L::DU + i, k);
        if (i == 0) r += at(L::HUX + 0, k) * dx[3] + at(L::HUX + 1, k) * dx[4] + at(L::HUX + 2, k) * dx[5];
#pragma unroll
        for (int a = 0; a < NX; a++) r -= B[a][i] * l1[a];
        at(L::RUG + i, k) = r;
      }
    }
    __syncwarp();
    if (lane == 0) {
      double n0 = 0.0, n1 = 0.0;
      for (int k = N - 1; k >= 1; k--) {
        n0 -= at(L::RUG + 0, k);
        n1 -= at(L::RUG + 1, k);
        at(L::LRP + 0, k) = n0;
        at(L::LRP + 1, k) = n1;
      }
    }
    __syncwarp();
  }

  __device__ __forceinline__ void slack_and_steps(double mu, double dw, double tau, double &a_pr, double &a_du, double &gd_out) {
    if (SHP) tied_control_multipliers();
    double rp = 0.0, rd = 0.0, gd = 0.0;
#define MPCB_LOWER(rgap, dv, z) do { double dz_ = -(z) + (mu - (z) * (dv)) * (rgap); \
    rp = fmax(rp, -(dv) * (rgap)); rd = fmax(rd, -dz_ * fast_rcp(z)); } while (0)
#define MPCB_UPPER(rgap, dv, z) do { double dz_ = -(z) + (mu + (z) * (dv)) * (rgap); \
    rp = fmax(rp, (dv) * (rgap)); rd = fmax(rd, -dz_ * fast_rcp(z)); } while (0)
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], dx[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        dx[i] = at(L::DX + i, k);
        if (k < N) gd += sigma * 2 * p.Q[i] * (xk[i] - xs[i]) * dx[i];
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = bx(b);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        gd += mu * (rh - rl) * dx[i];
        MPCB_LOWER(rl, dx[i], at(L::ZLX + b, k));
        MPCB_UPPER(rh, dx[i], at(L::ZUX + b, k));
      }
      double dsr[NR] = {0, 0}, lrp[NR] = {0, 0};
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k), du = at(L::DU + i, k);
          double um = k > 0 ? at(L::U + i, k - 1) : 0.0;
          double up = k + 1 <= N - 1 ? at(L::U + i, k + 1) : 0.0;
          double rl = fast_rcp(uk - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk);
          gd += (sigma * grad_u(k, i, uk, um, up) + mu * (rh - rl)) * du;
          MPCB_LOWER(rl, du, at(L::ZLU + i, k));
          MPCB_UPPER(rh, du, at(L::ZUU + i, k));
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k);
          double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
          double vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          double D = vl * rl + vu * rh + dw;
          double gs = mu * (rh - rl);
          double ds;
          if (SHP) {  // the slack step came out of the forward sweep, the equality multiplier out of the pass above
            ds = at(L::DSR + r, k);
            lrp[r] = at(L::LRP + r, k);
          } else {
            double res = at(L::U + r, k) - at(L::U + r, k - 1) - s;
            ds = at(L::DU + r, k) - at(L::DU + r, k - 1) + res;
            lrp[r] = D * ds + gs;
          }
          dsr[r] = ds;
          gd += gs * ds;
          MPCB_LOWER(rl, ds, vl);
          MPCB_UPPER(rh, ds, vu);
        }
      }
      double dso, lop;
      {
        double d, ox, oy, hxx, hxy, hyy;
        obs_row(k, xk[0], xk[1], d, ox, oy, hxx, hxy, hyy);
        double s = at(L::SO, k), rg = fast_rcp(s - p.obs_lo), vl = at(L::VLO, k);
        double D = vl * rg + dw;
        double gs = -mu * rg + MPCB_KAPPA_D * mu;
        double ds = ox * dx[0] + oy * dx[1] + (d - s);
        dso = ds;
        lop = D * ds + gs;
        gd += gs * ds;
        MPCB_LOWER(rg, ds, vl);
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) { at(L::DSR + r, k) = dsr[r]; at(L::LRP + r, k) = lrp[r]; }
      }
      at(L::DSO, k) = dso;
      at(L::LOP, k) = lop;
    }
#undef MPCB_LOWER
#undef MPCB_UPPER
    rp = warp_max(rp);
    rd = warp_max(rd);
    a_pr = rp > tau ? tau * fast_rcp(rp) : 1.0;
    a_du = rd > tau ? tau * fast_rcp(rd) : 1.0;
    gd_out = warp_sum(gd);
    __syncwarp();
  }

  __device__ __forceinline__ void accept_step(double a, double ad, double mu) {
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double l = at(L::LAM + i, k);
        at(L::LAM + i, k) = l + a * (at(L::LAMP + i, k) - l);
        at(L::CDEF + i, k) = at(L::CDEFT + i, k);
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = bx(b);
        double x = at(L::X + i, k), dx = at(L::DX + i, k);
        double rl = fast_rcp(x - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - x);
        double zl = at(L::ZLX + b, k), zu = at(L::ZUX + b, k);
        double dzl = -zl + (mu - zl * dx) * rl, dzu = -zu + (mu + zu * dx) * rh;
        double xn = x + a * dx;
        at(L::ZLX + b, k) = clampz(zl + ad * dzl, mu, fast_rcp(xn - p.x_lo[i]));
        at(L::ZUX + b, k) = clampz(zu + ad * dzu, mu, fast_rcp(p.x_hi[i] - xn));
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::X + i, k) += a * at(L::DX + i, k);
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double u = at(L::U + i, k), du = at(L::DU + i, k);
          double rl = fast_rcp(u - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - u);
          double zl = at(L::ZLU + i, k), zu = at(L::ZUU + i, k);
          double dzl = -zl + (mu - zl * du) * rl, dzu = -zu + (mu + zu * du) * rh;
          double un = u + a * du;
          at(L::U + i, k) = un;
          at(L::ZLU + i, k) = clampz(zl + ad * dzl, mu, fast_rcp(un - p.u_lo[i]));
          at(L::ZUU + i, k) = clampz(zu + ad * dzu, mu, fast_rcp(p.u_hi[i] - un));
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k), ds = at(L::DSR + r, k);
          double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
          double vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          double dvl = -vl + (mu - vl * ds) * rl, dvu = -vu + (mu + vu * ds) * rh;
          double sn = s + a * ds;
          at(L::SR + r, k) = sn;
          at(L::VLR + r, k) = clampz(vl + ad * dvl, mu, fast_rcp(sn - p.rate_lo[r]));
          at(L::VUR + r, k) = clampz(vu + ad * dvu, mu, fast_rcp(p.rate_hi[r] - sn));
          double l = at(L::LR + r, k);
          at(L::LR + r, k) = l + a * (at(L::LRP + r, k) - l);
        }
      }
      {
        double s = at(L::SO, k), ds = at(L::DSO, k);
        double rg = fast_rcp(s - p.obs_lo), vl = at(L::VLO, k);
        double dvl = -vl + (mu - vl * ds) * rg;
        double sn = s + a * ds;
        at(L::SO, k) = sn;
        at(L::VLO, k) = clampz(vl + ad * dvl, mu, fast_rcp(sn - p.obs_lo));
        double l = at(L::LO, k);
        at(L::LO, k) = l + a * (at(L::LOP, k) - l);
      }
    }
    __syncwarp();
  }

  __device__ __forceinline__ bool init_iterate(int b) {
    const int nv = 2 * N + NX * (N + 1);
    const double *zi = p.z_init ? p.z_init + (size_t)b * nv : nullptr;
    const double *ob = p.obs + (size_t)b * MO * (p.obs_input ? 1 : (N + 1)) * 6;
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      at(L::OCX, k) = ob[p.obs_input ? 0 : (size_t)k * 6 + 0];  // static centre (PKG/MPC_CBF_optimize_dyn.py:238-239)
      at(L::OCY, k) = ob[p.obs_input ? 1 : (size_t)k * 6 + 1];
#pragma unroll
      for (int i = 0; i < 2; i++) {
        at(L::U + i, k) = k < N ? push_in(zi ? zi[2 * k + i] : 0.0, p.u_lo[i], p.u_hi[i]) : 0.0;
        at(L::ZLU + i, k) = 1.0;
        at(L::ZUU + i, k) = 1.0;
        at(L::DU + i, k) = 0.0;
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (p.init_mode == 0) at(L::X + i, k) = zi ? zi[2 * N + NX * k + i] : 0.0;
        at(L::LAM + i, k) = 0.0;
        at(L::DX + i, k) = 0.0;
      }
    }
    __syncwarp();
    if (p.init_mode == 1) {
      double x[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) x[i] = x0[i];
#pragma unroll 1
      for (int k = 0; k <= N; k++) {
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NX; i++) at(L::X + i, k) = x[i];
        }
        if (k < N) {
          double u[2] = {at(L::U + 0, k), at(L::U + 1, k)}, f[NX];
          dyn_f(x, u, p, f);
#pragma unroll
          for (int i = 0; i < NX; i++) x[i] = x[i] + p.T * f[i];
        }
      }
      __syncwarp();
    }
    bool fin = true;
    double gmax = 0;
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int b2 = 0; b2 < NBX; b2++) {
        int i = bx(b2);
        at(L::X + i, k) = push_in(at(L::X + i, k), p.x_lo[i], p.x_hi[i]);
        at(L::ZLX + b2, k) = 1.0;
        at(L::ZUX + b2, k) = 1.0;
      }
    }
    if (SHP) __syncwarp();  // the relaxed-defect slacks below read the pushed state of stage k+1 (another lane's)
#pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      if (has_rate(k)) {
        double dfx[2] = {0, 0};
        if (SHP) {  // the slack relaxes the x / y defect into stage k+1
          double xk[NX], uk[2] = {at(L::U + 0, k), at(L::U + 1, k)}, f[NX];
#pragma unroll
          for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k);
          dyn_f(xk, uk, p, f);
          dfx[0] = at(L::X + 0, k + 1) - (xk[0] + p.T * f[0]);
          dfx[1] = at(L::X + 1, k + 1) - (xk[1] + p.T * f[1]);
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
          at(L::SR + r, k) = push_in(SHP ? dfx[r] : at(L::U + r, k) - at(L::U + r, k - 1), p.rate_lo[r], p.rate_hi[r]);
          at(L::VLR + r, k) = 1.0;
          at(L::VUR + r, k) = 1.0;
          at(L::LR + r, k) = 0.0;
        }
      }
      {
        double dx = at(L::X + 0, k) - at(L::OCX, k), dy = at(L::X + 1, k) - at(L::OCY, k);
        double e = dx * dx / (p.dyn_sx * p.dyn_sx) + dy * dy / (p.dyn_sy * p.dyn_sy) - 1.0;
        double d = e > 0.0 ? sqrt(e) : nan("");
        if (!isfinite(d)) fin = false;
        at(L::SO, k) = push_lo(d, p.obs_lo);
        at(L::VLO, k) = 1.0;
        at(L::LO, k) = 0.0;
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

#define MPCB_ITER_SYNC() do { if (((++tick) & (MPCB_SYNC_EVERY - 1)) == 0) block_iteration_vote(0); } while (0)
#include "mpcb_run_loop.inc"
#undef MPCB_ITER_SYNC
};

// persistent, W warps per block in step, see kin_solve_kernel
template <int W, bool SHP>
__global__ void __launch_bounds__(32 * W) dyn_solve_kernel(const __grid_constant__ KParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *gs = p.slab + ((size_t)blockIdx.x * W + warp) * DynLayout::slab_doubles();
  const int woff = warp * DynLayout::NF * (p.N + 1);
  int tick = 0;
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(p.counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= p.B) break;
    if (p.order) b = p.order[b];  // caller-supplied processing order (longest expected first)
    DynSolver<SHP> s(p, gs, woff, tick, lane);
    s.run(b);
    __syncwarp();
  }
  while (!block_iteration_vote(1)) {
  }
}

}  // namespace mpcb
