// mpcb200 device code, second engine: ONE SCENARIO PER LANE (kinematic family, plain obstacle rows).
//
// Why a second engine.  In the warp-per-scenario kernels (mpcb_kernel.cuh) the three serial sweeps of an interior-point
// iteration - backward Riccati, forward roll-out of the step, adjoint recursion - are executed by all 32 lanes on the same
// values: 73 % of the executed warp instructions, 1 useful lane of 32, and the FP64 pipe is the limiter
// (profiles/r01_final3_*).  With one scenario per lane every lane does useful FP64 work; the per-stage records then do not
// fit shared memory, so they live structure-of-arrays across scenarios in global memory ([stage][field][slot], slot
// fastest): a warp's access to one field of one stage is one 256-byte line, and the sweeps run at the HBM roofline
// (scripts/micro/lane_sweep.cu: 57.8 M scenario-stages/s/SM against 11 M for the redundant-lane sweeps).
//
// Two rules the stage loops follow, both measured: every load of a stage comes before its first
// store - all fields sit behind one pointer, so the compiler keeps each load behind every earlier store, and a store
// whose value waits for a load would serialise the stage into several HBM round trips - and the recursions that run
// backwards load stage k-1 before they store stage k.
//
// Same algorithm, same formulas as KinSolver (what replaces IPOPT + MUMPS behind PKG/MPC_CBF_optimize_kin.py:251-254 and
// the CasADi derivatives of :136-255); only the mapping to the machine differs:
//   * every phase is a loop over the stages of the lane's own scenario;
//   * lanes of a warp are at different points of their solves (iteration counts differ 5x, line searches backtrack,
//     inertia corrections repeat the factorisation), so the warp runs a fixed ROUND of phases and every lane takes part in
//     the phases its state asks for:  evaluate (first point / line-search trial) -> accept -> KKT error, barrier update ->
//     Newton step (condense, factorise, roll out, adjoint, step sizes).  A lane whose line search backtracks sits out the
//     rest of the round; a lane whose factorisation has the wrong inertia repeats only that phase in the next round;
//     a lane that is done writes its results and takes the next scenario from the queue.
#pragma once
#include "mpcb_kernel.cuh"

namespace mpcb {

// RK4: the shooting defects are X_{k+1} - Phi(X_k, U_k) with Phi the classical Runge-Kutta step (cfg.integrator): the stage
// then keeps the 11 entries of d Phi / d(x,u) that differ from the identity instead of sin/cos/tan, and d2L/(du dx) has
// four entries instead of one.
template <int NR, int MO, bool RK4 = false>
struct LaneLayout {
  static constexpr int NX = 4, NBX = 2;
  // field ids of the per-slot workspace, rows of S = N + 1 stages
  static constexpr int X = 0, U = X + 4, LAM = U + 2, ZLX = LAM + 4, ZUX = ZLX + 2, ZLU = ZUX + 2, ZUU = ZLU + 2;
  static constexpr int SR = ZUU + 2, VLR = SR + NR, VUR = VLR + NR, LR = VUR + NR;
  static constexpr int SO = LR + NR, VLO = SO + MO, LO = VLO + MO;
  static constexpr int OCX = LO + MO, OCY = OCX + MO, ISX = OCY + MO, ISY = ISX + MO;  // obstacle trajectory, 1/semi-axis^2
  static constexpr int DX = ISY + MO, DU = DX + 4, DSR = DU + 2, DSO = DSR + NR;       // the step
  static constexpr int LAMP = DSO + MO, LRP = LAMP + 4, LOP = LRP + NR;                // new multipliers
  static constexpr int CDEF = LOP + MO;  // two buffers of 4: defects of the iterate / of the trial point (flipped on acceptance)
  static constexpr int NJ = RK4 ? 11 : 3;  // Euler: sin(phi), cos(phi), tan(delta); RK4: a02 a03 a12 a13 a23 b00 b01 b10 b11 b20 b21
  static constexpr int TRG = CDEF + 8;   // two buffers of NJ
  static constexpr int HXX = TRG + 2 * NJ;  // condensed stage Hessian h00 h01 h11 h22 h23 h33, d2L/(du dx), gradient
  static constexpr int NHUX = RK4 ? 4 : 1;  // Euler: (delta, v); RK4: (delta, phi) (delta, v) (a, phi) (a, v)
  static constexpr int HUX = HXX + 6, GX = HUX + NHUX;
  // (the Euler families evaluate the condensed Hessian in both passes, LaneSolver::stage_h, and keep no copy: the ids
  // above then alias the gains and are never used)
  static constexpr int KX = RK4 ? GX + 4 : HXX, KW = KX + 8, KK = KW + 4;  // Riccati gains
  static constexpr int FLT = KK + 2;     // filter: 2 rows of theta entries, 2 rows of phi entries
  static constexpr int NFIELD = FLT + 4;
  __host__ __device__ static constexpr size_t slot_doubles(int N) { return (size_t)NFIELD * (size_t)(N + 1); }
};

enum { LANE_IDLE = 0, LANE_EVAL = 1, LANE_ACCEPT = 2, LANE_KKT = 3, LANE_NEWTON = 4, LANE_DONE = 5 };

// The classical Runge-Kutta step of the kinematic bicycle and its derivatives (oracle/nlp.py Rk4KinModel states the
// same thing densely).  Only phi and v of the stage points matter (x, y never enter the right-hand side), so everything
// is carried in the coordinates q = (phi, v, delta, a).
struct Rk4Stages {
  double v[4], sn[4], cs[4];  // stage speeds, sin / cos of the stage headings
  double dphi[4][4], dva[4];  // d phi_s / dq; d v_s / dq = (0, 1, 0, dva[s])
  double t, sec2;
  __device__ __forceinline__ void point(const double *x, const double *u, double T, double rL) {
    const double al[4] = {0.0, 0.5, 0.5, 1.0};
    t = tan(u[0]);
    sec2 = 1.0 + t * t;
    double ph = x[2];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      if (s > 0) ph = x[2] + al[s] * T * (v[s - 1] * t * rL);
      v[s] = x[3] + al[s] * T * u[1];
      dva[s] = al[s] * T;
      sincos(ph, &sn[s], &cs[s]);
      // d phi_s = e_phi + al T rL (t dv_{s-1} + v_{s-1} sec2 e_delta)
      dphi[s][0] = 1.0;
      dphi[s][1] = s > 0 ? al[s] * T * rL * t : 0.0;
      dphi[s][2] = s > 0 ? al[s] * T * rL * v[s - 1] * sec2 : 0.0;
      dphi[s][3] = s > 0 ? al[s] * T * rL * t * dva[s - 1] : 0.0;
    }
  }
  // Phi(x,u)
  __device__ __forceinline__ void step(const double *x, const double *u, double T, double rL, double *xn) const {
    const double b[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    double k0 = 0, k1 = 0, k2 = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) { k0 += b[s] * v[s] * cs[s]; k1 += b[s] * v[s] * sn[s]; k2 += b[s] * v[s]; }
    xn[0] = x[0] + T * k0; xn[1] = x[1] + T * k1; xn[2] = x[2] + T * (k2 * t * rL); xn[3] = x[3] + T * u[1];
  }
  // the 11 entries of d Phi / d(x,u) that are not those of the identity (b31 = T)
  __device__ __forceinline__ void jac(double T, double rL, double *J) const {
    const double b[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const double dv[4] = {0.0, 1.0, 0.0, dva[s]};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        d0[q] += b[s] * (cs[s] * dv[q] - v[s] * sn[s] * dphi[s][q]);
        d1[q] += b[s] * (sn[s] * dv[q] + v[s] * cs[s] * dphi[s][q]);
        d2[q] += b[s] * rL * (t * dv[q] + (q == 2 ? v[s] * sec2 : 0.0));
      }
    }
    J[0] = T * d0[0]; J[1] = T * d0[1]; J[2] = T * d1[0]; J[3] = T * d1[1]; J[4] = T * d2[1];  // a02 a03 a12 a13 a23
    J[5] = T * d0[2]; J[6] = T * d0[3]; J[7] = T * d1[2]; J[8] = T * d1[3]; J[9] = T * d2[2]; J[10] = T * d2[3];  // b00 b01 b10 b11 b20 b21
  }
  // sum_i lam_i d2 Phi_i / dq2, upper triangle in the order (phi phi, phi v, phi delta, phi a, v v, v delta, v a, delta delta, delta a, a a):
  // second-order adjoint  sum_s Z_s' [nu_s . g''(z_s)] Z_s,  nu_s = T b_s lam + al_{s+1} T g_x(z_{s+1})' nu_{s+1}
  __device__ __forceinline__ void hess(const double *lam, double T, double rL, double *H) const {
    const double al[4] = {0.0, 0.5, 0.5, 1.0}, b[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
#pragma unroll
    for (int i = 0; i < 10; i++) H[i] = 0.0;
    double n0 = 0, n1 = 0, n2 = 0;
#pragma unroll
    for (int s = 3; s >= 0; s--) {
      double m2 = T * b[s] * lam[2];
      if (s < 3) m2 += al[s + 1] * T * (-v[s + 1] * sn[s + 1] * n0 + v[s + 1] * cs[s + 1] * n1);
      // components 0, 1 of nu never pick up anything (x, y do not enter g): nu_s0 = T b_s lam0, nu_s1 = T b_s lam1
      n0 = T * b[s] * lam[0]; n1 = T * b[s] * lam[1]; n2 = m2;
      // NOTE: the recursion above needs nu_{s+1,0..1} of the NEXT stage, which are T b_{s+1} lam: recomputed here
      const double Mpp = -v[s] * (n0 * cs[s] + n1 * sn[s]), Mpv = -n0 * sn[s] + n1 * cs[s];
      const double Mvd = n2 * sec2 * rL, Mdd = n2 * 2.0 * v[s] * sec2 * t * rL;
      const double r0[4] = {dphi[s][0], dphi[s][1], dphi[s][2], dphi[s][3]}, r1[4] = {0.0, 1.0, 0.0, dva[s]};
      int e = 0;
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i; j < 4; j++, e++) {
          double h = Mpp * r0[i] * r0[j] + Mpv * (r0[i] * r1[j] + r1[i] * r0[j]);
          h += Mvd * (r1[i] * (j == 2 ? 1.0 : 0.0) + (i == 2 ? 1.0 : 0.0) * r1[j]);
          if (i == 2 && j == 2) h += Mdd;
          H[e] += h;
        }
    }
  }
};

template <int NR, int MO, bool RK4 = false>
struct LaneSolver {
  using L = LaneLayout<NR, MO, RK4>;
  static constexpr int NX = 4, NBX = 2;
  const KParams &p;
  double *ws;
  const size_t nslot, slot;
  const int N, S;

  // ---- per-lane solver state
  int state = LANE_IDLE, b = -1, status = 2, it = 0, nfilt = 0, cur = 0;
  bool trial = false, first = true, tried0 = false, to_resto = false;
  double sigma = 1.0, mu = 0, tau = 0, dw = 0, dw_last = 0;
  double theta = 0, fobj = 0, bar = 0, lin = 0, phi = 0, theta_min = 0, theta_max = 0;
  double a = 0, a_min = 0, a_dual = 1, gd = 0, pgd = 0, pth = 0, err_last = INFINITY;
  double th_e = 0, f_e = 0, bar_e = 0, lin_e = 0;  // last evaluation
  double xs[NX];

  __device__ LaneSolver(const KParams &p_, double *ws_, size_t nslot_, size_t slot_)
      : p(p_), ws(ws_), nslot(nslot_), slot(slot_), N(p_.N), S(p_.N + 1) {}

  // Workspace layout [stage][field][slot]: one 256-byte line per field, stage and warp, the ~40 fields a stage touches in
  // consecutive rows.  Measured against [field][stage][slot] (the fields of a stage 23 MB apart: a DRAM page and a TLB
  // entry each) +18 %, against a warp-private [warp][stage][field][lane] +4 %.
  __device__ __forceinline__ double &at(int f, int k) { return ws[((size_t)k * L::NFIELD + f) * nslot + slot]; }
  // Software prefetch: the stage loops carry their recursions in registers, but nothing lets the hardware see the next
  // stage's loads early; every phase therefore asks for the lines of the stage PF_DIST ahead (one 256-byte line per
  // field and warp) while it works on the current one.  The kernel uses no shared memory, so L1 holds them.
#ifndef MPCB_LANE_PF_DIST
#define MPCB_LANE_PF_DIST 0  // measured: 2 stages ahead costs 9 % at full occupancy and buys 20 % at 4 warps per SM (profiles/experiments)
#endif
#ifndef MPCB_LANE_PF_INSTR
#define MPCB_LANE_PF_INSTR "prefetch.global.L2"
#endif
  template <int F0, int CNT>
  __device__ __forceinline__ void pf(int k) {
    if (MPCB_LANE_PF_DIST > 0 && k >= 0 && k <= N) {
#pragma unroll
      for (int f = 0; f < CNT; f++) asm volatile(MPCB_LANE_PF_INSTR " [%0];" ::"l"(&at(F0 + f, k)));
    }
  }
  __device__ __forceinline__ bool has_rate(int k) const { return NR > 0 && k >= 1 && k <= N - 1; }
  __device__ __forceinline__ bool has_obs(int k) const { return MO > 0 && k <= N - 1; }
  __device__ static __forceinline__ constexpr int bx(int i) { return i == 0 ? 1 : 3; }  // bounded states: y, vx
  __device__ __forceinline__ int cdef(int buf) const { return L::CDEF + 4 * buf; }

  __device__ __forceinline__ int trg(int buf) const { return L::TRG + L::NJ * buf; }
  // A = d Phi/dx (entries a02 a03 a12 a13 a23 beside the unit diagonal) and B = d Phi/du of stage k at the current iterate
  struct AB { double a02, a03, a12, a13, a23, b00, b01, b10, b11, b20, b21, b31; };
  __device__ __forceinline__ AB load_ab(int tg, int k, double v) {
    AB o;
    if (RK4) {
      o.a02 = at(tg + 0, k); o.a03 = at(tg + 1, k); o.a12 = at(tg + 2, k); o.a13 = at(tg + 3, k); o.a23 = at(tg + 4, k);
      o.b00 = at(tg + 5, k); o.b01 = at(tg + 6, k); o.b10 = at(tg + 7, k); o.b11 = at(tg + 8, k); o.b20 = at(tg + 9, k); o.b21 = at(tg + 10, k);
    } else {
      const double rL = 1.0 / p.Veh_l, s = at(tg + 0, k), c = at(tg + 1, k), t = at(tg + 2, k);
      o.a02 = p.T * (-v * s); o.a03 = p.T * c; o.a12 = p.T * (v * c); o.a13 = p.T * s; o.a23 = p.T * (t * rL);
      o.b00 = 0; o.b01 = 0; o.b10 = 0; o.b11 = 0; o.b20 = p.T * (v * (1.0 + t * t) * rL); o.b21 = 0;
    }
    o.b31 = p.T;
    return o;
  }

  __device__ __forceinline__ double grad_u(int k, int i, double uk, double ukm1, double ukp1) const {
    double v = 2 * p.R[i] * uk;
    if (k > 0) v += 2 * p.DR[i] * (uk - ukm1);
    else if (p.du0_cost) v += 2 * p.DR[i] * uk;
    if (k + 1 <= N - 1) v -= 2 * p.DR[i] * (ukp1 - uk);
    return v;
  }

  // ---------------------------------------------------------------- start point (KinSolver::init_iterate)
  __device__ void init_iterate() {
    const int nv = 2 * N + NX * (N + 1);
    const double *zi = p.z_init ? p.z_init + (size_t)b * nv : nullptr;
    const double *ob = p.obs ? p.obs + (size_t)b * MO * (p.obs_input ? 1 : (N + 1)) * 6 : nullptr;
    double x0[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) { x0[i] = p.x0[(size_t)b * NX + i]; xs[i] = p.xs[(size_t)b * NX + i]; }
    if (MO > 0 && p.obs_input) {  // obstacle states / static rows: PKG/Obs_prediction.py:27-28 step by step
#pragma unroll
      for (int j = 0; j < MO; j++) {
        const double *o = ob + (size_t)j * 6;
        double x = o[0], y = o[1];
        const bool moving = p.obs_input == 1;
        const double dx = moving ? o[3] * cos(o[2]) * p.T : 0.0, dy = moving ? o[3] * sin(o[2]) * p.T : 0.0;
        const double sx = p.ego_hl + o[4] / 2 + p.safe_l, sy = p.ego_hw + o[5] / 2 + p.safe_w;
        const double isx = 1.0 / (sx * sx), isy = 1.0 / (sy * sy);
#pragma unroll 1
        for (int k = 0; k <= N; k++) {
          at(L::OCX + j, k) = x; at(L::OCY + j, k) = y; at(L::ISX + j, k) = isx; at(L::ISY + j, k) = isy;
          x = x + dx;
          y = y + dy;
        }
      }
    }
    // controls, rolled-out or given states, pushed into their bounds; unit bound multipliers; slacks
    double x[NX], um = 0.0, gmax = 0.0;
#pragma unroll
    for (int i = 0; i < NX; i++) x[i] = x0[i];
#pragma unroll 1
    for (int k = 0; k <= N; k++) {
      if (MO > 0 && !p.obs_input) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          const double *o = ob + ((size_t)j * (N + 1) + k) * 6;
          double sx = p.ego_hl + o[4] / 2 + p.safe_l, sy = p.ego_hw + o[5] / 2 + p.safe_w;
          at(L::OCX + j, k) = o[0]; at(L::OCY + j, k) = o[1];
          at(L::ISX + j, k) = 1.0 / (sx * sx); at(L::ISY + j, k) = 1.0 / (sy * sy);
        }
      }
      double uk[2] = {0, 0};
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) uk[i] = push_in(zi ? zi[2 * k + i] : 0.0, p.u_lo[i], p.u_hi[i]);
      }
      double xk[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = p.init_mode == 1 ? x[i] : (zi ? zi[2 * N + NX * k + i] : 0.0);
      if (p.init_mode == 1 && k < N) {  // roll-out of the guessed controls with the step of the defects (PKG/MPC_CBF_optimize_kin.py:207)
        if (RK4) {
          Rk4Stages rk;
          double xn[NX];
          rk.point(x, uk, p.T, 1.0 / p.Veh_l);
          rk.step(x, uk, p.T, 1.0 / p.Veh_l, xn);
#pragma unroll
          for (int i = 0; i < NX; i++) x[i] = xn[i];
        } else {
          double s, c, t;
          d_trig(x[2], uk[0], &s, &c, &t);
          double f0 = x[3] * c, f1 = x[3] * s, f2 = x[3] * t * (1.0 / p.Veh_l);
          x[0] = x[0] + p.T * f0; x[1] = x[1] + p.T * f1; x[2] = x[2] + p.T * f2; x[3] = x[3] + p.T * uk[1];
        }
      }
#pragma unroll
      for (int b2 = 0; b2 < NBX; b2++) { int i = bx(b2); xk[i] = push_in(xk[i], p.x_lo[i], p.x_hi[i]); at(L::ZLX + b2, k) = 1.0; at(L::ZUX + b2, k) = 1.0; }
#pragma unroll
      for (int i = 0; i < NX; i++) { at(L::X + i, k) = xk[i]; at(L::LAM + i, k) = 0.0; at(L::DX + i, k) = 0.0; }
#pragma unroll
      for (int i = 0; i < 2; i++) { at(L::U + i, k) = uk[i]; at(L::ZLU + i, k) = 1.0; at(L::ZUU + i, k) = 1.0; at(L::DU + i, k) = 0.0; }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          at(L::SR + r, k) = push_in(uk[0] - um, p.rate_lo[r], p.rate_hi[r]);  // rate row on the steering angle (rate_ctrl[0] = 0)
          at(L::VLR + r, k) = 1.0; at(L::VUR + r, k) = 1.0; at(L::LR + r, k) = 0.0; at(L::DSR + r, k) = 0.0;
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double d = dx * dx * at(L::ISX + j, k) + dy * dy * at(L::ISY + j, k) - 1.0;
          at(L::SO + j, k) = push_lo(d, p.obs_lo);
          at(L::VLO + j, k) = 1.0; at(L::LO + j, k) = 0.0; at(L::DSO + j, k) = 0.0;
        }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < NX; i++) gmax = fmax(gmax, fabs(2 * p.Q[i] * (xk[i] - xs[i])));
      }
      um = uk[0];
    }
    // gradient-based objective scaling needs u_{k-1}, u_k, u_{k+1}: a second pass over the controls
    {
      double u_m[2] = {0, 0}, u_c[2] = {at(L::U + 0, 0), at(L::U + 1, 0)};
#pragma unroll 1
      for (int k = 0; k < N; k++) {
        double u_n[2] = {0, 0};
        if (k + 1 <= N - 1) { u_n[0] = at(L::U + 0, k + 1); u_n[1] = at(L::U + 1, k + 1); }
#pragma unroll
        for (int i = 0; i < 2; i++) gmax = fmax(gmax, fabs(grad_u(k, i, u_c[i], u_m[i], u_n[i])));
        u_m[0] = u_c[0]; u_m[1] = u_c[1]; u_c[0] = u_n[0]; u_c[1] = u_n[1];
      }
    }
    sigma = gmax > MPCB_OBJ_SCALE_MAX_GRAD ? MPCB_OBJ_SCALE_MAX_GRAD / gmax : 1.0;
    if (sigma < 1e-8) sigma = 1e-8;
    mu = p.mu_init;
    tau = fmax(MPCB_TAU_MIN, 1 - mu);
    status = 2; it = 0; nfilt = 0; cur = 0; dw_last = 0.0; err_last = INFINITY;
    trial = false; first = true; a = 0.0; to_resto = false;
  }

  // ---------------------------------------------------------------- point evaluation (KinSolver::eval_point)
  // at z + alpha dz: infeasibility (1-norm), objective, log-barrier sum and damping sum; stores the point's defects and
  // sin/cos/tan in buffer `buf`
  __device__ void eval_point(double alpha, int buf) {
    const int cd = cdef(buf), tg = trg(buf);
    const double rL = 1.0 / p.Veh_l;
    double th = 0, fo = 0, br = 0, ln = 0;
    double xk[NX], um[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < NX; i++) {
      xk[i] = at(L::X + i, 0) + alpha * at(L::DX + i, 0);
      double c0 = xk[i] - p.x0[(size_t)b * NX + i];
      th += fabs(c0);
      at(cd + i, 0) = c0;
    }
#pragma unroll 1
    for (int k = 0; k <= N; k++) {
      pf<L::X, 6>(k + 1 + MPCB_LANE_PF_DIST);           // X of stage k+1 is read at stage k
      pf<L::SR, 1>(k + MPCB_LANE_PF_DIST);
      pf<L::SO, 1>(k + MPCB_LANE_PF_DIST);
      pf<L::OCX, 4 * MO>(k + MPCB_LANE_PF_DIST);
      pf<L::DX, 6 + NR + MO>(k + 1 + MPCB_LANE_PF_DIST);
      double gp = 1.0, uk[2] = {0, 0};
      double xn_[NX] = {0, 0, 0, 0}, dxn_[NX] = {0, 0, 0, 0};  // (every load of the stage before its first store, see LaneSolver)
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) uk[i] = at(L::U + i, k) + alpha * at(L::DU + i, k);
#pragma unroll
        for (int i = 0; i < NX; i++) { xn_[i] = at(L::X + i, k + 1); dxn_[i] = at(L::DX + i, k + 1); }
      }
#pragma unroll
      for (int b2 = 0; b2 < NBX; b2++) { int i = bx(b2); gp *= (xk[i] - p.x_lo[i]) * (p.x_hi[i] - xk[i]); }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k) + alpha * at(L::DSR + r, k);
          th += fabs(uk[0] - um[0] - s);
          gp *= (s - p.rate_lo[r]) * (p.rate_hi[r] - s);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double d = dx * dx * at(L::ISX + j, k) + dy * dy * at(L::ISY + j, k) - 1.0;  // PKG/..._kin.py:244,247
          double s = at(L::SO + j, k) + alpha * at(L::DSO + j, k);
          th += fabs(d - s);
          gp *= s - p.obs_lo;
          ln += s - p.obs_lo;
        }
      }
      if (k < N) {
        double step[NX];  // Phi(x_k, u_k)
        if (RK4) {
          Rk4Stages rk;
          double J[11];
          rk.point(xk, uk, p.T, rL);
          rk.step(xk, uk, p.T, rL, step);
          rk.jac(p.T, rL, J);
#pragma unroll
          for (int i = 0; i < 11; i++) at(tg + i, k) = J[i];
        } else {
          double s, c, t;
          d_trig(xk[2], uk[0], &s, &c, &t);
          at(tg + 0, k) = s; at(tg + 1, k) = c; at(tg + 2, k) = t;
          const double f[NX] = {xk[3] * c, xk[3] * s, xk[3] * t * rL, uk[1]};  // PKG/MPC_CBF_optimize_kin.py:153-156
#pragma unroll
          for (int i = 0; i < NX; i++) step[i] = xk[i] + p.T * f[i];
        }
        double xn[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) {
          xn[i] = xn_[i] + alpha * dxn_[i];
          double d = xn[i] - step[i];
          th += fabs(d);
          at(cd + i, k + 1) = d;
          double e = xk[i] - xs[i];
          fo += p.Q[i] * e * e;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
          gp *= (uk[i] - p.u_lo[i]) * (p.u_hi[i] - uk[i]);
          fo += p.R[i] * uk[i] * uk[i];
          if (k > 0) { double e = uk[i] - um[i]; fo += p.DR[i] * e * e; }
          else if (p.du0_cost) fo += p.DR[i] * uk[i] * uk[i];
        }
#pragma unroll
        for (int i = 0; i < NX; i++) xk[i] = xn[i];
        um[0] = uk[0]; um[1] = uk[1];
      }
      br += d_log(gp);
    }
    th_e = th; f_e = fo; bar_e = br; lin_e = ln;
  }

  // ---------------------------------------------------------------- KKT error pieces (KinSolver::kkt_pieces)
  struct Kkt { double dual, prim, cmin, cmax, sum_lam, sum_z; };

  __device__ void kkt_pieces(Kkt &o) {
    const int cd = cdef(cur), tg = trg(cur);
    double dual = 0, prim = 0, cmin = INFINITY, cmax = -INFINITY, sl = 0, sz = 0;
#define MPCB_COMPL(gap, mult) do { double p_ = (gap) * (mult); cmin = fmin(cmin, p_); cmax = fmax(cmax, p_); sz += (mult); } while (0)
    double lam[NX], um[2] = {0, 0}, uc[2] = {at(L::U + 0, 0), at(L::U + 1, 0)};
    double lr_c = 0.0;  // multiplier of the rate row of stage k
#pragma unroll
    for (int i = 0; i < NX; i++) lam[i] = at(L::LAM + i, 0);
#pragma unroll 1  // (the pass stores nothing, so unrolling would put the loads of several stages in flight: measured -3 % at 2)
    for (int k = 0; k <= N; k++) {
      pf<L::X, L::OCX - L::X + 4 * MO>(k + 1 + MPCB_LANE_PF_DIST);  // the whole iterate and the obstacle row
      pf<L::CDEF, 8>(k + 1 + MPCB_LANE_PF_DIST);
      pf<L::TRG, 2 * L::NJ>(k + 1 + MPCB_LANE_PF_DIST);
      double xk[NX], l1[NX] = {0, 0, 0, 0}, un[2] = {0, 0};
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        prim = fmax(prim, fabs(at(cd + i, k)));
        sl += fabs(lam[i]);
      }
      double rx[NX] = {lam[0], lam[1], lam[2], lam[3]};
      AB ab = {};
      if (k < N) {
        ab = load_ab(tg, k, xk[3]);
        const double a02 = ab.a02, a03 = ab.a03, a12 = ab.a12, a13 = ab.a13, a23 = ab.a23;
#pragma unroll
        for (int i = 0; i < NX; i++) {
          l1[i] = at(L::LAM + i, k + 1);
          rx[i] += sigma * 2 * p.Q[i] * (xk[i] - xs[i]);
        }
        rx[0] -= l1[0];
        rx[1] -= l1[1];
        rx[2] -= l1[2] + a02 * l1[0] + a12 * l1[1];
        rx[3] -= l1[3] + a03 * l1[0] + a13 * l1[1] + a23 * l1[2];
        if (k + 1 <= N - 1) { un[0] = at(L::U + 0, k + 1); un[1] = at(L::U + 1, k + 1); }
      }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) {
        int i = bx(b_);
        double zl = at(L::ZLX + b_, k), zu = at(L::ZUX + b_, k);
        rx[i] += -zl + zu;
        MPCB_COMPL(xk[i] - p.x_lo[i], zl);
        MPCB_COMPL(p.x_hi[i] - xk[i], zu);
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double a_ = at(L::ISX + j, k), b_ = at(L::ISY + j, k);
          double d = dx * dx * a_ + dy * dy * b_ - 1.0;
          double lo = at(L::LO + j, k), vl = at(L::VLO + j, k), s = at(L::SO + j, k);
          rx[0] += lo * (2 * dx * a_);
          rx[1] += lo * (2 * dy * b_);
          dual = fmax(dual, fabs(-lo - vl));
          prim = fmax(prim, fabs(d - s));
          MPCB_COMPL(s - p.obs_lo, vl);
          sl += fabs(lo);
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) dual = fmax(dual, fabs(rx[i]));
      double lr_n = 0.0;  // multiplier of the rate row of stage k+1
      if (NR > 0 && has_rate(k + 1)) lr_n = at(L::LR, k + 1);
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double zl = at(L::ZLU + i, k), zu = at(L::ZUU + i, k);
          double r = sigma * grad_u(k, i, uc[i], um[i], un[i]) - zl + zu;
          r -= (i == 0) ? ab.b00 * l1[0] + ab.b10 * l1[1] + ab.b20 * l1[2] : ab.b01 * l1[0] + ab.b11 * l1[1] + ab.b21 * l1[2] + ab.b31 * l1[3];  // - B' lam_{k+1}
          if (NR > 0 && i == 0) {
            if (has_rate(k)) r += lr_c;
            if (has_rate(k + 1)) r -= lr_n;
          }
          dual = fmax(dual, fabs(r));
          MPCB_COMPL(uc[i] - p.u_lo[i], zl);
          MPCB_COMPL(p.u_hi[i] - uc[i], zu);
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = at(L::SR + r, k), vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          dual = fmax(dual, fabs(-lr_c - vl + vu));
          prim = fmax(prim, fabs(uc[0] - um[0] - s));
          MPCB_COMPL(s - p.rate_lo[r], vl);
          MPCB_COMPL(p.rate_hi[r] - s, vu);
          sl += fabs(lr_c);
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) lam[i] = l1[i];
      um[0] = uc[0]; um[1] = uc[1]; uc[0] = un[0]; uc[1] = un[1];
      lr_c = lr_n;
    }
#undef MPCB_COMPL
    o.dual = dual; o.prim = prim; o.cmin = cmin; o.cmax = cmax; o.sum_lam = sl; o.sum_z = sz;
  }

  __device__ __forceinline__ double kkt_error(const Kkt &o, double mu_, double &co) const {
    co = p.n_bm > 0 ? fmax(fabs(o.cmax - mu_), fabs(o.cmin - mu_)) : 0.0;
    double rs_d = MPCB_S_MAX * fast_rcp(fmax(MPCB_S_MAX, (o.sum_lam + o.sum_z) * fast_rcp(fmax(1.0, (double)(p.n_eq + p.n_bm)))));
    double rs_c = MPCB_S_MAX * fast_rcp(fmax(MPCB_S_MAX, o.sum_z * fast_rcp(fmax(1.0, (double)p.n_bm))));
    return fmax(fmax(o.dual * rs_d, o.prim), co * rs_c);
  }

  // ---------------------------------------------------------------- Newton step, part 1: condense + factorise
  // KinSolver::build_qp and riccati_backward in one backward pass; the condensed stage Hessian / gradient is kept for the
  // adjoint pass, the gains for the forward pass.  Returns false when some F_uu is not positive definite.
  // Condensed stage Hessian and gradient of the Euler families (KinSolver::build_qp): h = diag + (0,1) (2,3) entries,
  // d2L/(d delta d v), the model part of d2L/d delta^2 and gx.  The factorisation pass and the roll-out both evaluate it
  // from the iterate (recomputing it costs a few dozen flops; keeping it in the workspace cost 22 doubles of traffic per
  // scenario-stage in a kernel that is bandwidth-bound).
  struct StageH { double h[NX], h01, h23, hdv, hdd_f, gx[NX], a02, a03, a12, a13, a23, b2; };
  __device__ __forceinline__ void stage_h(int k, const double *xk, const double *tg_, const double *l_next, const double *zlx_, const double *zux_,
                                          bool obst, const double *ocx_, const double *ocy_, const double *isx_, const double *isy_,
                                          const double *so_, const double *vlo_, const double *lo_, StageH &o) const {
    const double T = p.T, rL = 1.0 / p.Veh_l;
    double h[NX] = {dw, dw, dw, dw}, h01 = 0, h23 = 0, hdv = 0, hdd_f = 0, gx[NX] = {0, 0, 0, 0};
    double a02 = 0, a03 = 0, a12 = 0, a13 = 0, a23 = 0, b2 = 0;
    if (k < N) {
      const double s = tg_[0], c = tg_[1], t = tg_[2];
      a02 = T * (-xk[3] * s); a03 = T * c; a12 = T * (xk[3] * c); a13 = T * s; a23 = T * (t * rL);
      b2 = T * (xk[3] * (1.0 + t * t) * rL);
      const double jd = (T * rL) * (1.0 + t * t);
      h[2] += l_next[0] * a12 - l_next[1] * a02;
      h23 = l_next[0] * a13 - l_next[1] * a03;
      hdv = -l_next[2] * jd;
      hdd_f = -2.0 * l_next[2] * b2 * t;
#pragma unroll
      for (int i = 0; i < NX; i++) { h[i] += sigma * 2 * p.Q[i]; gx[i] = sigma * 2 * p.Q[i] * (xk[i] - xs[i]); }
    }
#pragma unroll
    for (int b_ = 0; b_ < NBX; b_++) {
      int i = bx(b_);
      double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
      h[i] += zlx_[b_] * rl + zux_[b_] * rh;
      gx[i] += mu * (rh - rl);
    }
    if (obst) {
#pragma unroll
      for (int j = 0; j < MO; j++) {
        double dx = xk[0] - ocx_[j], dy = xk[1] - ocy_[j];
        double a_ = isx_[j], b_ = isy_[j];
        double d = dx * dx * a_ + dy * dy * b_ - 1.0;
        double ox = 2 * dx * a_, oy = 2 * dy * b_;
        double s = so_[j], rg = fast_rcp(s - p.obs_lo);
        double D = vlo_[j] * rg + dw;
        double gs = -mu * rg + MPCB_KAPPA_D * mu;
        double lo = lo_[j];
        double t = D * (d - s) + gs;
        h[0] += lo * (2 * a_) + D * ox * ox;
        h01 += D * ox * oy;
        h[1] += lo * (2 * b_) + D * oy * oy;
        gx[0] += ox * t;
        gx[1] += oy * t;
      }
    }
#pragma unroll
    for (int i = 0; i < NX; i++) { o.h[i] = h[i]; o.gx[i] = gx[i]; }
    o.h01 = h01; o.h23 = h23; o.hdv = hdv; o.hdd_f = hdd_f;
    o.a02 = a02; o.a03 = a03; o.a12 = a12; o.a13 = a13; o.a23 = a23; o.b2 = b2;
  }

  __device__ bool backward() {
    const int cd = cdef(cur), tg = trg(cur);
    const double T = p.T;
    double p00, p01, p11, p22, p23, p33, p02 = 0, p03 = 0, p12 = 0, p13 = 0, px0, px1, px2, px3;
    double w0d = 0, w0a = 0, w1d = 0, w1a = 0, w2d = 0, w2a = 0, w3d = 0, w3a = 0;
    double qdd = 0, qda = 0, qaa = 0, pwd = 0, pwa = 0;
    double l_next[NX] = {0, 0, 0, 0}, c_next[NX] = {0, 0, 0, 0};  // multipliers / defects of stage k+1
    double u_k[2] = {0, 0};                                       // controls of stage k (read one stage ahead)
    if (N >= 1) { u_k[0] = at(L::U + 0, N - 1); u_k[1] = at(L::U + 1, N - 1); }
#pragma unroll 1
    for (int k = N; k >= 0; k--) {
      pf<L::X, L::OCX - L::X + 4 * MO>(k - 1 - MPCB_LANE_PF_DIST);
      pf<L::CDEF, 8>(k - 1 - MPCB_LANE_PF_DIST);
      pf<L::TRG, 2 * L::NJ>(k - 1 - MPCB_LANE_PF_DIST);
      double xk[NX], h[NX] = {dw, dw, dw, dw}, h01 = 0, h23 = 0, hdv = 0, hdd_f = 0, gx[NX] = {0, 0, 0, 0};
      double a02 = 0, a03 = 0, a12 = 0, a13 = 0, a23 = 0, b2 = 0;
      // every load of the stage first (a store whose value waits for a load would hold back the loads behind it)
      constexpr int NR1 = NR > 0 ? NR : 1, MO1 = MO > 0 ? MO : 1;
      double tg_[3] = {0, 0, 0}, zlx_[NBX], zux_[NBX], c_k[NX], lam_k[NX], u_m[2] = {0, 0}, zlu_[2] = {0, 0}, zuu_[2] = {0, 0};
      double sr_[NR1], vlr_[NR1], vur_[NR1], ocx_[MO1], ocy_[MO1], isx_[MO1], isy_[MO1], so_[MO1], vlo_[MO1], lo_[MO1];
      const bool rate = has_rate(k), obst = has_obs(k);
#pragma unroll
      for (int i = 0; i < NX; i++) { xk[i] = at(L::X + i, k); c_k[i] = at(cd + i, k); lam_k[i] = at(L::LAM + i, k); }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) { zlx_[b_] = at(L::ZLX + b_, k); zux_[b_] = at(L::ZUX + b_, k); }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 3; i++) tg_[i] = at(tg + i, k);
#pragma unroll
        for (int i = 0; i < 2; i++) { zlu_[i] = at(L::ZLU + i, k); zuu_[i] = at(L::ZUU + i, k); }
        if (k >= 1) { u_m[0] = at(L::U + 0, k - 1); u_m[1] = at(L::U + 1, k - 1); }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) { sr_[r] = at(L::SR + r, k); vlr_[r] = at(L::VLR + r, k); vur_[r] = at(L::VUR + r, k); }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          ocx_[j] = at(L::OCX + j, k); ocy_[j] = at(L::OCY + j, k); isx_[j] = at(L::ISX + j, k); isy_[j] = at(L::ISY + j, k);
          so_[j] = at(L::SO + j, k); vlo_[j] = at(L::VLO + j, k); lo_[j] = at(L::LO + j, k);
        }
      }
      StageH sh;
      stage_h(k, xk, tg_, l_next, zlx_, zux_, obst, ocx_, ocy_, isx_, isy_, so_, vlo_, lo_, sh);
#pragma unroll
      for (int i = 0; i < NX; i++) { h[i] = sh.h[i]; gx[i] = sh.gx[i]; }
      h01 = sh.h01; h23 = sh.h23; hdv = sh.hdv; hdd_f = sh.hdd_f;
      a02 = sh.a02; a03 = sh.a03; a12 = sh.a12; a13 = sh.a13; a23 = sh.a23; b2 = sh.b2;
      if (k == N) {
        p00 = h[0]; p01 = h01; p11 = h[1]; p22 = h[2]; p23 = h23; p33 = h[3];
        px0 = gx[0]; px1 = gx[1]; px2 = gx[2]; px3 = gx[3];
      } else {
        // control block of the stage: cost, bounds, coupling with the previous control (rate cost and rate row)
        double Huu[2], gu[2], E[2] = {0, 0}, tk[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = u_k[i];
          double g = sigma * 2 * p.R[i] * uk;
          double hd = sigma * 2 * p.R[i] + dw + (i == 0 ? hdd_f : 0.0);
          if (k == 0 && p.du0_cost) { hd += sigma * 2 * p.DR[i]; g += sigma * 2 * p.DR[i] * uk; }
          double rl = fast_rcp(uk - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk);
          hd += zlu_[i] * rl + zuu_[i] * rh;
          g += mu * (rh - rl);
          Huu[i] = hd; gu[i] = g;
          if (k >= 1) { E[i] = sigma * 2 * p.DR[i]; tk[i] = sigma * 2 * p.DR[i] * (uk - u_m[i]); }
        }
        if (rate) {
#pragma unroll
          for (int r = 0; r < NR; r++) {
            double s = sr_[r];
            double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
            double D = vlr_[r] * rl + vur_[r] * rh + dw;
            double gs = mu * (rh - rl);
            double res = u_k[0] - u_m[0] - s;
            E[0] += D; tk[0] += D * res + gs;
          }
        }
        const double Ed = E[0], Ea = E[1], td = tk[0], ta = tk[1];
        const double b0 = -c_next[0], b1 = -c_next[1], b2_ = -c_next[2], b3 = -c_next[3];
        const double m02 = p02 + a02 * p00 + a12 * p01;
        const double m12 = p12 + a02 * p01 + a12 * p11;
        const double m22 = p22 + a02 * p02 + a12 * p12;
        const double m32 = p23 + a02 * p03 + a12 * p13;
        const double m03 = p03 + a03 * p00 + a13 * p01 + a23 * p02;
        const double m13 = p13 + a03 * p01 + a13 * p11 + a23 * p12;
        const double m23 = p23 + a03 * p02 + a13 * p12 + a23 * p22;
        const double m33 = p33 + a03 * p03 + a13 * p13 + a23 * p23;
        const double f00 = h[0] + p00, f01 = h01 + p01, f11 = h[1] + p11;
        const double f02 = m02, f03 = m03, f12 = m12, f13 = m13;
        const double f22 = h[2] + m22 + a02 * m02 + a12 * m12;
        const double f23 = h23 + m23 + a02 * m03 + a12 * m13;
        const double f33 = h[3] + m33 + a03 * m03 + a13 * m13 + a23 * m23;
        const double ud0 = b2 * p02 + w0d, ud1 = b2 * p12 + w1d;
        const double ud2 = b2 * m22 + w2d + a02 * w0d + a12 * w1d;
        const double ud3 = hdv + b2 * m23 + w3d + a03 * w0d + a13 * w1d + a23 * w2d;
        const double ua0 = T * p03 + w0a, ua1 = T * p13 + w1a;
        const double ua2 = T * m32 + w2a + a02 * w0a + a12 * w1a;
        const double ua3 = T * m33 + w3a + a03 * w0a + a13 * w1a + a23 * w2a;
        const double Fdd = Huu[0] + Ed + qdd + b2 * (b2 * p22 + 2.0 * w2d);
        const double Fda = qda + b2 * (T * p23) + b2 * w2a + T * w3d;
        const double Faa = Huu[1] + Ea + qaa + T * (T * p33 + 2.0 * w3a);
        const double Pb0 = px0 + p00 * b0 + p01 * b1 + p02 * b2_ + p03 * b3;
        const double Pb1 = px1 + p01 * b0 + p11 * b1 + p12 * b2_ + p13 * b3;
        const double Pb2 = px2 + p02 * b0 + p12 * b1 + p22 * b2_ + p23 * b3;
        const double Pb3 = px3 + p03 * b0 + p13 * b1 + p23 * b2_ + p33 * b3;
        const double fx0 = gx[0] + Pb0, fx1 = gx[1] + Pb1;
        const double fx2 = gx[2] + Pb2 + a02 * Pb0 + a12 * Pb1;
        const double fx3 = gx[3] + Pb3 + a03 * Pb0 + a13 * Pb1 + a23 * Pb2;
        const double fud = gu[0] + td + pwd + b2 * Pb2 + w0d * b0 + w1d * b1 + w2d * b2_ + w3d * b3;
        const double fua = gu[1] + ta + pwa + T * Pb3 + w0a * b0 + w1a * b1 + w2a * b2_ + w3a * b3;
        const double det = Fdd * Faa - Fda * Fda;
        if (!(Fdd > 0.0) || !(det > 0.0) || !isfinite(det)) return false;
        const double id = fast_rcp(det);
        const double idd = Faa * id, ida = -Fda * id, iaa = Fdd * id;
        const double kd0 = -(idd * ud0 + ida * ua0), kd1 = -(idd * ud1 + ida * ua1), kd2 = -(idd * ud2 + ida * ua2), kd3 = -(idd * ud3 + ida * ua3);
        const double ka0 = -(ida * ud0 + iaa * ua0), ka1 = -(ida * ud1 + iaa * ua1), ka2 = -(ida * ud2 + iaa * ua2), ka3 = -(ida * ud3 + iaa * ua3);
        const double wdd = idd * Ed, wda = ida * Ea, wad = ida * Ed, waa = iaa * Ea;
        const double kkd = -(idd * fud + ida * fua), kka = -(ida * fud + iaa * fua);
        at(L::KX + 0, k) = kd0; at(L::KX + 1, k) = kd1; at(L::KX + 2, k) = kd2; at(L::KX + 3, k) = kd3;
        at(L::KX + 4, k) = ka0; at(L::KX + 5, k) = ka1; at(L::KX + 6, k) = ka2; at(L::KX + 7, k) = ka3;
        at(L::KW + 0, k) = wdd; at(L::KW + 1, k) = wda; at(L::KW + 2, k) = wad; at(L::KW + 3, k) = waa;
        at(L::KK + 0, k) = kkd; at(L::KK + 1, k) = kka;
        p00 = f00 + ud0 * kd0 + ua0 * ka0;
        p11 = f11 + ud1 * kd1 + ua1 * ka1;
        p22 = f22 + ud2 * kd2 + ua2 * ka2;
        p33 = f33 + ud3 * kd3 + ua3 * ka3;
        p01 = f01 + ud0 * kd1 + ua0 * ka1;
        p02 = f02 + ud0 * kd2 + ua0 * ka2;
        p03 = f03 + ud0 * kd3 + ua0 * ka3;
        p12 = f12 + ud1 * kd2 + ua1 * ka2;
        p13 = f13 + ud1 * kd3 + ua1 * ka3;
        p23 = f23 + ud2 * kd3 + ua2 * ka3;
        w0d = ud0 * wdd + ua0 * wad; w0a = ud0 * wda + ua0 * waa;
        w1d = ud1 * wdd + ua1 * wad; w1a = ud1 * wda + ua1 * waa;
        w2d = ud2 * wdd + ua2 * wad; w2a = ud2 * wda + ua2 * waa;
        w3d = ud3 * wdd + ua3 * wad; w3a = ud3 * wda + ua3 * waa;
        px0 = fx0 + ud0 * kkd + ua0 * kka;
        px1 = fx1 + ud1 * kkd + ua1 * kka;
        px2 = fx2 + ud2 * kkd + ua2 * kka;
        px3 = fx3 + ud3 * kkd + ua3 * kka;
        qdd = Ed - Ed * wdd;
        qda = -0.5 * (Ed * wda + Ea * wad);
        qaa = Ea - Ea * waa;
        pwd = -td - Ed * kkd;
        pwa = -ta - Ea * kka;
        u_k[0] = u_m[0]; u_k[1] = u_m[1];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) { l_next[i] = lam_k[i]; c_next[i] = c_k[i]; }
    }
    return true;
  }

  // The same pass for the Runge-Kutta defects: B = d Phi/du is full, d2L/(du dx) has four entries, Huu three; the value
  // function recursion is written with small dense arrays (the compiler unrolls them).
  __device__ bool backward_rk4() {
    const int cd = cdef(cur), tg = trg(cur);
    const double T = p.T, rL = 1.0 / p.Veh_l;
    double P[4][4], W[4][2], Qw[2][2] = {{0, 0}, {0, 0}}, px[4], pw[2] = {0, 0};
    double l_next[NX] = {0, 0, 0, 0}, c_next[NX] = {0, 0, 0, 0};
    double u_k[2] = {0, 0};
    if (N >= 1) { u_k[0] = at(L::U + 0, N - 1); u_k[1] = at(L::U + 1, N - 1); }
#pragma unroll 1
    for (int k = N; k >= 0; k--) {
      double xk[NX], h[NX] = {dw, dw, dw, dw}, h01 = 0, h23 = 0, gx[NX] = {0, 0, 0, 0};
      double Hux[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}}, Hm[10];
#pragma unroll
      for (int i = 0; i < 10; i++) Hm[i] = 0.0;
      // every load of the stage first (see backward)
      constexpr int NR1 = NR > 0 ? NR : 1, MO1 = MO > 0 ? MO : 1;
      double zlx_[NBX], zux_[NBX], c_k[NX], lam_k[NX], u_m[2] = {0, 0}, zlu_[2] = {0, 0}, zuu_[2] = {0, 0};
      double sr_[NR1], vlr_[NR1], vur_[NR1], ocx_[MO1], ocy_[MO1], isx_[MO1], isy_[MO1], so_[MO1], vlo_[MO1], lo_[MO1];
      const bool rate = has_rate(k), obst = has_obs(k);
#pragma unroll
      for (int i = 0; i < NX; i++) { xk[i] = at(L::X + i, k); c_k[i] = at(cd + i, k); lam_k[i] = at(L::LAM + i, k); }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) { zlx_[b_] = at(L::ZLX + b_, k); zux_[b_] = at(L::ZUX + b_, k); }
      AB ab = {};
      if (k < N) {
        ab = load_ab(tg, k, xk[3]);
#pragma unroll
        for (int i = 0; i < 2; i++) { zlu_[i] = at(L::ZLU + i, k); zuu_[i] = at(L::ZUU + i, k); }
        if (k >= 1) { u_m[0] = at(L::U + 0, k - 1); u_m[1] = at(L::U + 1, k - 1); }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) { sr_[r] = at(L::SR + r, k); vlr_[r] = at(L::VLR + r, k); vur_[r] = at(L::VUR + r, k); }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          ocx_[j] = at(L::OCX + j, k); ocy_[j] = at(L::OCY + j, k); isx_[j] = at(L::ISX + j, k); isy_[j] = at(L::ISY + j, k);
          so_[j] = at(L::SO + j, k); vlo_[j] = at(L::VLO + j, k); lo_[j] = at(L::LO + j, k);
        }
      }
      if (k < N) {
        Rk4Stages rk;
        rk.point(xk, u_k, T, rL);
        rk.hess(l_next, T, rL, Hm);  // defect = X_{k+1} - Phi: the Lagrangian takes MINUS lam' Phi''
        h[2] -= Hm[0]; h23 = -Hm[1]; h[3] -= Hm[4];
        Hux[0][2] = -Hm[2]; Hux[0][3] = -Hm[5]; Hux[1][2] = -Hm[3]; Hux[1][3] = -Hm[6];
#pragma unroll
        for (int i = 0; i < NX; i++) { h[i] += sigma * 2 * p.Q[i]; gx[i] = sigma * 2 * p.Q[i] * (xk[i] - xs[i]); }
      }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) {
        int i = bx(b_);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        h[i] += zlx_[b_] * rl + zux_[b_] * rh;
        gx[i] += mu * (rh - rl);
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - ocx_[j], dy = xk[1] - ocy_[j];
          double a_ = isx_[j], b_ = isy_[j];
          double d = dx * dx * a_ + dy * dy * b_ - 1.0;
          double ox = 2 * dx * a_, oy = 2 * dy * b_;
          double s = so_[j], rg = fast_rcp(s - p.obs_lo);
          double D = vlo_[j] * rg + dw;
          double gs = -mu * rg + MPCB_KAPPA_D * mu;
          double lo = lo_[j];
          double t = D * (d - s) + gs;
          h[0] += lo * (2 * a_) + D * ox * ox;
          h01 += D * ox * oy;
          h[1] += lo * (2 * b_) + D * oy * oy;
          gx[0] += ox * t;
          gx[1] += oy * t;
        }
      }
      at(L::HXX + 0, k) = h[0]; at(L::HXX + 1, k) = h01; at(L::HXX + 2, k) = h[1];
      at(L::HXX + 3, k) = h[2]; at(L::HXX + 4, k) = h23; at(L::HXX + 5, k) = h[3];
      at(L::HUX + 0, k) = Hux[0][2]; at(L::HUX + 1, k) = Hux[0][3]; at(L::HUX + 2, k) = Hux[1][2]; at(L::HUX + 3, k) = Hux[1][3];
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::GX + i, k) = gx[i];
      double Hxx[4][4] = {{h[0], h01, 0, 0}, {h01, h[1], 0, 0}, {0, 0, h[2], h23}, {0, 0, h23, h[3]}};
      if (k == N) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          px[i] = gx[i]; W[i][0] = 0; W[i][1] = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) P[i][j] = Hxx[i][j];
        }
      } else {
        double Huu[2][2] = {{0, -Hm[8]}, {-Hm[8], 0}}, gu[2], E[2] = {0, 0}, tk[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = u_k[i];
          double g = sigma * 2 * p.R[i] * uk;
          double hd = sigma * 2 * p.R[i] + dw - (i == 0 ? Hm[7] : Hm[9]);
          if (k == 0 && p.du0_cost) { hd += sigma * 2 * p.DR[i]; g += sigma * 2 * p.DR[i] * uk; }
          double rl = fast_rcp(uk - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk);
          hd += zlu_[i] * rl + zuu_[i] * rh;
          g += mu * (rh - rl);
          Huu[i][i] = hd; gu[i] = g;
          if (k >= 1) { E[i] = sigma * 2 * p.DR[i]; tk[i] = sigma * 2 * p.DR[i] * (uk - u_m[i]); }
        }
        if (rate) {
#pragma unroll
          for (int r = 0; r < NR; r++) {
            double s = sr_[r];
            double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
            double D = vlr_[r] * rl + vur_[r] * rh + dw;
            double gs = mu * (rh - rl);
            double res = u_k[0] - u_m[0] - s;
            E[0] += D; tk[0] += D * res + gs;
          }
        }
        const double A[4][4] = {{1, 0, ab.a02, ab.a03}, {0, 1, ab.a12, ab.a13}, {0, 0, 1, ab.a23}, {0, 0, 0, 1}};
        const double Bm[4][2] = {{ab.b00, ab.b01}, {ab.b10, ab.b11}, {ab.b20, ab.b21}, {0, ab.b31}};
        double b[4], Pb[4], PA[4][4], PB[4][2];
#pragma unroll
        for (int i = 0; i < 4; i++) b[i] = -c_next[i];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          double a = px[i];
#pragma unroll
          for (int j = 0; j < 4; j++) a += P[i][j] * b[j];
          Pb[i] = a;
#pragma unroll
          for (int j = 0; j < 4; j++) { double m = 0; for (int q = 0; q < 4; q++) m += P[i][q] * A[q][j]; PA[i][j] = m; }
#pragma unroll
          for (int j = 0; j < 2; j++) { double m = 0; for (int q = 0; q < 4; q++) m += P[i][q] * Bm[q][j]; PB[i][j] = m; }
        }
        double Fxx[4][4], Fux[2][4], Fuu[2][2], fx[4], fu[2];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
          for (int j = 0; j < 4; j++) { double m = Hxx[i][j]; for (int q = 0; q < 4; q++) m += A[q][i] * PA[q][j]; Fxx[i][j] = m; }
          double m = gx[i];
#pragma unroll
          for (int q = 0; q < 4; q++) m += A[q][i] * Pb[q];
          fx[i] = m;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
          for (int j = 0; j < 4; j++) { double m = Hux[i][j]; for (int q = 0; q < 4; q++) m += Bm[q][i] * PA[q][j] + W[q][i] * A[q][j]; Fux[i][j] = m; }
#pragma unroll
          for (int j = 0; j < 2; j++) {
            double m = Huu[i][j] + Qw[i][j] + (i == j ? E[i] : 0.0);
#pragma unroll
            for (int q = 0; q < 4; q++) m += Bm[q][i] * PB[q][j] + Bm[q][i] * W[q][j] + W[q][i] * Bm[q][j];
            Fuu[i][j] = m;
          }
          double m = gu[i] + tk[i] + pw[i];
#pragma unroll
          for (int q = 0; q < 4; q++) m += Bm[q][i] * Pb[q] + W[q][i] * b[q];
          fu[i] = m;
        }
        const double Fda = 0.5 * (Fuu[0][1] + Fuu[1][0]);
        const double det = Fuu[0][0] * Fuu[1][1] - Fda * Fda;
        if (!(Fuu[0][0] > 0.0) || !(det > 0.0) || !isfinite(det)) return false;
        const double id = fast_rcp(det);
        const double Fi[2][2] = {{Fuu[1][1] * id, -Fda * id}, {-Fda * id, Fuu[0][0] * id}};
        double Kx[2][4], Kw[2][2], kk[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
          for (int j = 0; j < 4; j++) Kx[i][j] = -(Fi[i][0] * Fux[0][j] + Fi[i][1] * Fux[1][j]);
#pragma unroll
          for (int j = 0; j < 2; j++) Kw[i][j] = Fi[i][j] * E[j];
          kk[i] = -(Fi[i][0] * fu[0] + Fi[i][1] * fu[1]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) { at(L::KX + j, k) = Kx[0][j]; at(L::KX + 4 + j, k) = Kx[1][j]; }
        at(L::KW + 0, k) = Kw[0][0]; at(L::KW + 1, k) = Kw[0][1]; at(L::KW + 2, k) = Kw[1][0]; at(L::KW + 3, k) = Kw[1][1];
        at(L::KK + 0, k) = kk[0]; at(L::KK + 1, k) = kk[1];
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
          for (int j = 0; j < 4; j++) P[i][j] = Fxx[i][j] + Fux[0][i] * Kx[0][j] + Fux[1][i] * Kx[1][j];
#pragma unroll
          for (int j = 0; j < 2; j++) W[i][j] = Fux[0][i] * Kw[0][j] + Fux[1][i] * Kw[1][j];
          px[i] = fx[i] + Fux[0][i] * kk[0] + Fux[1][i] * kk[1];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = i + 1; j < 4; j++) { double m = 0.5 * (P[i][j] + P[j][i]); P[i][j] = m; P[j][i] = m; }
#pragma unroll
        for (int i = 0; i < 2; i++) {
#pragma unroll
          for (int j = 0; j < 2; j++) Qw[i][j] = (i == j ? E[i] : 0.0) - E[i] * Kw[i][j];
          pw[i] = -tk[i] - E[i] * kk[i];
        }
        { double m = 0.5 * (Qw[0][1] + Qw[1][0]); Qw[0][1] = m; Qw[1][0] = m; }
        u_k[0] = u_m[0]; u_k[1] = u_m[1];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) { l_next[i] = lam_k[i]; c_next[i] = c_k[i]; }
    }
    return true;
  }

  // ---------------------------------------------------------------- Newton step, part 2: roll the step out
  // KinSolver::riccati_forward + the stage residual r_k of the adjoint pass + slack_and_steps
  __device__ void forward(double &a_pr, double &a_du, double &gd_out) {
    const int cd = cdef(cur), tg = trg(cur);
    const double T = p.T;
    double rp = 0.0, rd = 0.0, g_d = 0.0;
#define MPCB_LOWER(rgap, dv, z) do { double dz_ = -(z) + (mu - (z) * (dv)) * (rgap); \
    rp = fmax(rp, -(dv) * (rgap)); rd = fmax(rd, -dz_ * fast_rcp(z)); } while (0)
#define MPCB_UPPER(rgap, dv, z) do { double dz_ = -(z) + (mu + (z) * (dv)) * (rgap); \
    rp = fmax(rp, (dv) * (rgap)); rd = fmax(rd, -dz_ * fast_rcp(z)); } while (0)
    double d[NX], vd = 0, va = 0;  // dx_k, du_{k-1}
    double um[2] = {0, 0}, uc[2] = {at(L::U + 0, 0), at(L::U + 1, 0)};
#pragma unroll
    for (int i = 0; i < NX; i++) d[i] = -at(cd + i, 0);
#pragma unroll 1
    for (int k = 0; k <= N; k++) {
      pf<L::X, L::OCX - L::X + 4 * MO>(k + 1 + MPCB_LANE_PF_DIST);
      pf<L::CDEF, 8>(k + 1 + MPCB_LANE_PF_DIST);
      pf<L::TRG, 2 * L::NJ>(k + 1 + MPCB_LANE_PF_DIST);
      pf<L::HXX, (RK4 ? 10 + L::NHUX : 0) + 14>(k + 1 + MPCB_LANE_PF_DIST);   // (condensed Hessian, gradient and) the gains
      constexpr int NR1 = NR > 0 ? NR : 1, MO1 = MO > 0 ? MO : 1, NHUX1 = RK4 ? 4 : 1;
      double xk[NX], un[2] = {0, 0}, ud = 0, ua = 0, n[NX] = {0, 0, 0, 0};
      // every load of the stage first (a store whose value waits for a load would hold back the loads behind it)
      double kk_[2] = {0, 0}, kx_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, kw_[4] = {0, 0, 0, 0}, cn_[NX] = {0, 0, 0, 0};
      double hx_[6], gx_[NX], hux_[NHUX1], zlx_[NBX], zux_[NBX], zlu_[2] = {0, 0}, zuu_[2] = {0, 0};
      double sr_[NR1], vlr_[NR1], vur_[NR1], ocx_[MO1], ocy_[MO1], isx_[MO1], isy_[MO1], so_[MO1], vlo_[MO1], lo_[MO1];
      double tg_[3] = {0, 0, 0}, ln_[NX] = {0, 0, 0, 0};  // Euler families: the stage Hessian is evaluated again (stage_h)
      const bool rate = has_rate(k), obst = has_obs(k);
      AB ab = {};
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k);
      if (RK4) {
#pragma unroll
        for (int i = 0; i < NX; i++) gx_[i] = at(L::GX + i, k);
#pragma unroll
        for (int i = 0; i < 6; i++) hx_[i] = at(L::HXX + i, k);
#pragma unroll
        for (int i = 0; i < NHUX1; i++) hux_[i] = at(L::HUX + i, k);
      } else if (k < N) {
#pragma unroll
        for (int i = 0; i < 3; i++) tg_[i] = at(tg + i, k);
#pragma unroll
        for (int i = 0; i < NX; i++) ln_[i] = at(L::LAM + i, k + 1);
      }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) { zlx_[b_] = at(L::ZLX + b_, k); zux_[b_] = at(L::ZUX + b_, k); }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) { kk_[i] = at(L::KK + i, k); zlu_[i] = at(L::ZLU + i, k); zuu_[i] = at(L::ZUU + i, k); }
#pragma unroll
        for (int i = 0; i < 8; i++) kx_[i] = at(L::KX + i, k);
#pragma unroll
        for (int i = 0; i < 4; i++) kw_[i] = at(L::KW + i, k);
#pragma unroll
        for (int i = 0; i < NX; i++) cn_[i] = at(cd + i, k + 1);
        if (RK4) ab = load_ab(tg, k, xk[3]);
        if (k + 1 <= N - 1) { un[0] = at(L::U + 0, k + 1); un[1] = at(L::U + 1, k + 1); }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) { sr_[r] = at(L::SR + r, k); vlr_[r] = at(L::VLR + r, k); vur_[r] = at(L::VUR + r, k); }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          ocx_[j] = at(L::OCX + j, k); ocy_[j] = at(L::OCY + j, k); isx_[j] = at(L::ISX + j, k); isy_[j] = at(L::ISY + j, k);
          so_[j] = at(L::SO + j, k); vlo_[j] = at(L::VLO + j, k);
          if (!RK4) lo_[j] = at(L::LO + j, k);
        }
      }
      if (!RK4) {
        StageH sh;
        stage_h(k, xk, tg_, ln_, zlx_, zux_, obst, ocx_, ocy_, isx_, isy_, so_, vlo_, lo_, sh);
        hx_[0] = sh.h[0]; hx_[1] = sh.h01; hx_[2] = sh.h[1]; hx_[3] = sh.h[2]; hx_[4] = sh.h23; hx_[5] = sh.h[3];
        hux_[0] = sh.hdv;
#pragma unroll
        for (int i = 0; i < NX; i++) gx_[i] = sh.gx[i];
        ab.a02 = sh.a02; ab.a03 = sh.a03; ab.a12 = sh.a12; ab.a13 = sh.a13; ab.a23 = sh.a23; ab.b20 = sh.b2; ab.b31 = T;
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::DX + i, k) = d[i];
      if (k < N) {
        ud = kk_[0] + kx_[0] * d[0] + kx_[1] * d[1] + kx_[2] * d[2] + kx_[3] * d[3] + kw_[0] * vd + kw_[1] * va;
        ua = kk_[1] + kx_[4] * d[0] + kx_[5] * d[1] + kx_[6] * d[2] + kx_[7] * d[3] + kw_[2] * vd + kw_[3] * va;
        n[0] = d[0] + ab.a02 * d[2] + ab.a03 * d[3] - cn_[0];
        n[1] = d[1] + ab.a12 * d[2] + ab.a13 * d[3] - cn_[1];
        n[2] = d[2] + ab.a23 * d[3] + ab.b20 * ud - cn_[2];
        n[3] = d[3] + T * ua - cn_[3];
        if (RK4) { n[0] += ab.b00 * ud + ab.b01 * ua; n[1] += ab.b10 * ud + ab.b11 * ua; n[2] += ab.b21 * ua; }
      }
      at(L::DU + 0, k) = ud; at(L::DU + 1, k) = ua;
      // stage residual of the adjoint recursion: r_k = Hxx_eff dx + Hux' du + gx_eff (written to the LAMP row)
      {
        const double h00 = hx_[0], h01 = hx_[1], h11 = hx_[2], h22 = hx_[3];
        const double h23 = hx_[4], h33 = hx_[5];
        double r2 = gx_[2] + h22 * d[2] + h23 * d[3], r3 = gx_[3] + h23 * d[2] + h33 * d[3];
        if (RK4) {  // Hux' du: (delta, phi) (delta, v) (a, phi) (a, v)
          r2 += hux_[0] * ud + hux_[NHUX1 > 2 ? 2 : 0] * ua;
          r3 += hux_[NHUX1 > 1 ? 1 : 0] * ud + hux_[NHUX1 > 3 ? 3 : 0] * ua;
        } else {
          r3 += hux_[0] * ud;
        }
        at(L::LAMP + 0, k) = gx_[0] + h00 * d[0] + h01 * d[1];
        at(L::LAMP + 1, k) = gx_[1] + h01 * d[0] + h11 * d[1];
        at(L::LAMP + 2, k) = r2;
        at(L::LAMP + 3, k) = r3;
      }
      // slack steps, new row multipliers, fraction to the boundary, barrier slope
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (k < N) g_d += sigma * 2 * p.Q[i] * (xk[i] - xs[i]) * d[i];
      }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) {
        int i = bx(b_);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        g_d += mu * (rh - rl) * d[i];
        MPCB_LOWER(rl, d[i], zlx_[b_]);
        MPCB_UPPER(rh, d[i], zux_[b_]);
      }
      if (k < N) {
        const double du_[2] = {ud, ua};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double rl = fast_rcp(uc[i] - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uc[i]);
          g_d += (sigma * grad_u(k, i, uc[i], um[i], un[i]) + mu * (rh - rl)) * du_[i];
          MPCB_LOWER(rl, du_[i], zlu_[i]);
          MPCB_UPPER(rh, du_[i], zuu_[i]);
        }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = sr_[r];
          double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
          double vl = vlr_[r], vu = vur_[r];
          double D = vl * rl + vu * rh + dw;
          double gs = mu * (rh - rl);
          double res = uc[0] - um[0] - s;
          double ds = ud - vd + res;
          at(L::DSR + r, k) = ds;
          at(L::LRP + r, k) = D * ds + gs;
          g_d += gs * ds;
          MPCB_LOWER(rl, ds, vl);
          MPCB_UPPER(rh, ds, vu);
        }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double ex = xk[0] - ocx_[j], ey = xk[1] - ocy_[j];
          double a_ = isx_[j], b_ = isy_[j];
          double dd = ex * ex * a_ + ey * ey * b_ - 1.0;
          double s = so_[j], rg = fast_rcp(s - p.obs_lo), vl = vlo_[j];
          double D = vl * rg + dw;
          double gs = -mu * rg + MPCB_KAPPA_D * mu;
          double ds = (2 * ex * a_) * d[0] + (2 * ey * b_) * d[1] + (dd - s);
          at(L::DSO + j, k) = ds;
          at(L::LOP + j, k) = D * ds + gs;
          g_d += gs * ds;
          MPCB_LOWER(rg, ds, vl);
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) d[i] = n[i];
      vd = ud; va = ua;
      um[0] = uc[0]; um[1] = uc[1]; uc[0] = un[0]; uc[1] = un[1];
    }
#undef MPCB_LOWER
#undef MPCB_UPPER
    a_pr = rp > tau ? tau * fast_rcp(rp) : 1.0;
    a_du = rd > tau ? tau * fast_rcp(rd) : 1.0;
    gd_out = g_d;
  }

  // new dynamics multipliers: lam+_k = A_k' lam+_{k+1} - r_k  (KinSolver::adjoint)
  __device__ void adjoint() {
    const int tg = trg(cur);
    constexpr int NT = RK4 ? 5 : 3;
    double l0 = -at(L::LAMP + 0, N), l1 = -at(L::LAMP + 1, N), l2 = -at(L::LAMP + 2, N), l3 = -at(L::LAMP + 3, N);
    at(L::LAMP + 0, N) = l0; at(L::LAMP + 1, N) = l1; at(L::LAMP + 2, N) = l2; at(L::LAMP + 3, N) = l3;
    // the inputs of stage k-1 are loaded before stage k stores (one stage in flight ahead of the recursion)
    double v_n = 0, t_n[NT], r_n[NX];
    if (N >= 1) {
      v_n = RK4 ? 0.0 : at(L::X + 3, N - 1);
#pragma unroll
      for (int i = 0; i < NT; i++) t_n[i] = at(tg + i, N - 1);
#pragma unroll
      for (int i = 0; i < NX; i++) r_n[i] = at(L::LAMP + i, N - 1);
    }
#pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      pf<L::X + 3, 1>(k - 1 - MPCB_LANE_PF_DIST);
      pf<L::TRG, 2 * L::NJ>(k - 1 - MPCB_LANE_PF_DIST);
      pf<L::LAMP, 4>(k - 1 - MPCB_LANE_PF_DIST);
      const double v = v_n, r0 = r_n[0], r1 = r_n[1], r2 = r_n[2], r3 = r_n[3];
      double a02, a03, a12, a13, a23;
      if (RK4) { a02 = t_n[0]; a03 = t_n[1]; a12 = t_n[2]; a13 = t_n[NT > 3 ? 3 : 0]; a23 = t_n[NT > 4 ? 4 : 0]; }
      else {
        const double rL = 1.0 / p.Veh_l, s_ = t_n[0], c_ = t_n[1], t_ = t_n[2];
        a02 = p.T * (-v * s_); a03 = p.T * c_; a12 = p.T * (v * c_); a13 = p.T * s_; a23 = p.T * (t_ * rL);
      }
      if (k >= 1) {
        v_n = RK4 ? 0.0 : at(L::X + 3, k - 1);
#pragma unroll
        for (int i = 0; i < NT; i++) t_n[i] = at(tg + i, k - 1);
#pragma unroll
        for (int i = 0; i < NX; i++) r_n[i] = at(L::LAMP + i, k - 1);
      }
      const double n0 = l0 - r0;
      const double n1 = l1 - r1;
      const double n2 = l2 + a02 * l0 + a12 * l1 - r2;
      const double n3 = l3 + a03 * l0 + a13 * l1 + a23 * l2 - r3;
      at(L::LAMP + 0, k) = n0; at(L::LAMP + 1, k) = n1; at(L::LAMP + 2, k) = n2; at(L::LAMP + 3, k) = n3;
      l0 = n0; l1 = n1; l2 = n2; l3 = n3;
    }
  }

  // accept the step: primal a, duals ad (KinSolver::accept_step); the trial point's buffer becomes the current one
  __device__ void accept_step(double a_, double ad) {
    constexpr int NR1 = NR > 0 ? NR : 1, MO1 = MO > 0 ? MO : 1;
#pragma unroll 1
    for (int k = 0; k <= N; k++) {
      pf<L::X, L::OCX - L::X>(k + 1 + MPCB_LANE_PF_DIST);
      pf<L::DX, L::CDEF - L::DX>(k + 1 + MPCB_LANE_PF_DIST);
      // every load of the stage first: a store whose value waits for a load would otherwise hold back the loads behind it
      double lam[NX], lamp[NX], xv[NX], dxv[NX], zlx[NBX], zux[NBX];
      double uv[2] = {0, 0}, duv[2] = {0, 0}, zlu[2] = {0, 0}, zuu[2] = {0, 0};
      double sr[NR1], dsr[NR1], vlr[NR1], vur[NR1], lr[NR1], lrp[NR1];
      double so[MO1], dso[MO1], vlo[MO1], lo[MO1], lop[MO1];
      const bool rate = has_rate(k), obst = has_obs(k);
#pragma unroll
      for (int i = 0; i < NX; i++) { lam[i] = at(L::LAM + i, k); lamp[i] = at(L::LAMP + i, k); xv[i] = at(L::X + i, k); dxv[i] = at(L::DX + i, k); }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) { zlx[b_] = at(L::ZLX + b_, k); zux[b_] = at(L::ZUX + b_, k); }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) { uv[i] = at(L::U + i, k); duv[i] = at(L::DU + i, k); zlu[i] = at(L::ZLU + i, k); zuu[i] = at(L::ZUU + i, k); }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          sr[r] = at(L::SR + r, k); dsr[r] = at(L::DSR + r, k); vlr[r] = at(L::VLR + r, k); vur[r] = at(L::VUR + r, k);
          lr[r] = at(L::LR + r, k); lrp[r] = at(L::LRP + r, k);
        }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          so[j] = at(L::SO + j, k); dso[j] = at(L::DSO + j, k); vlo[j] = at(L::VLO + j, k); lo[j] = at(L::LO + j, k); lop[j] = at(L::LOP + j, k);
        }
      }
#pragma unroll
      for (int i = 0; i < NX; i++) { double l = lam[i]; at(L::LAM + i, k) = l + a_ * (lamp[i] - l); }
#pragma unroll
      for (int b_ = 0; b_ < NBX; b_++) {
        int i = bx(b_);
        double x = xv[i], dx = dxv[i];
        double rl = fast_rcp(x - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - x);
        double zl = zlx[b_], zu = zux[b_];
        double dzl = -zl + (mu - zl * dx) * rl, dzu = -zu + (mu + zu * dx) * rh;
        double xn = x + a_ * dx;
        at(L::ZLX + b_, k) = clampz(zl + ad * dzl, mu, fast_rcp(xn - p.x_lo[i]));
        at(L::ZUX + b_, k) = clampz(zu + ad * dzu, mu, fast_rcp(p.x_hi[i] - xn));
      }
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::X + i, k) = xv[i] + a_ * dxv[i];
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double u = uv[i], du = duv[i];
          double rl = fast_rcp(u - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - u);
          double zl = zlu[i], zu = zuu[i];
          double dzl = -zl + (mu - zl * du) * rl, dzu = -zu + (mu + zu * du) * rh;
          double un = u + a_ * du;
          at(L::U + i, k) = un;
          at(L::ZLU + i, k) = clampz(zl + ad * dzl, mu, fast_rcp(un - p.u_lo[i]));
          at(L::ZUU + i, k) = clampz(zu + ad * dzu, mu, fast_rcp(p.u_hi[i] - un));
        }
      }
      if (rate) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          double s = sr[r], ds = dsr[r];
          double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
          double vl = vlr[r], vu = vur[r];
          double dvl = -vl + (mu - vl * ds) * rl, dvu = -vu + (mu + vu * ds) * rh;
          double sn = s + a_ * ds;
          at(L::SR + r, k) = sn;
          at(L::VLR + r, k) = clampz(vl + ad * dvl, mu, fast_rcp(sn - p.rate_lo[r]));
          at(L::VUR + r, k) = clampz(vu + ad * dvu, mu, fast_rcp(p.rate_hi[r] - sn));
          double l = lr[r];
          at(L::LR + r, k) = l + a_ * (lrp[r] - l);
        }
      }
      if (obst) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double s = so[j], ds = dso[j];
          double rg = fast_rcp(s - p.obs_lo), vl = vlo[j];
          double dvl = -vl + (mu - vl * ds) * rg;
          double sn = s + a_ * ds;
          at(L::SO + j, k) = sn;
          at(L::VLO + j, k) = clampz(vl + ad * dvl, mu, fast_rcp(sn - p.obs_lo));
          double l = lo[j];
          at(L::LO + j, k) = l + a_ * (lop[j] - l);
        }
      }
    }
    cur ^= 1;
  }

  // ---------------------------------------------------------------- filter (entries live in the workspace)
  __device__ __forceinline__ bool filter_blocks(double th, double ph) {
    if (th >= theta_max) return true;
    const int n = nfilt < 2 * S ? nfilt : 2 * S;
    for (int e = 0; e < n; e++) {
      double ft = at(L::FLT + (e >= S ? 1 : 0), e >= S ? e - S : e), fp = at(L::FLT + 2 + (e >= S ? 1 : 0), e >= S ? e - S : e);
      if (th >= ft && ph >= fp) return true;
    }
    return false;
  }
  __device__ __forceinline__ void filter_add(double th, double ph) {
    const int e = nfilt % (2 * S);
    at(L::FLT + (e >= S ? 1 : 0), e >= S ? e - S : e) = th;
    at(L::FLT + 2 + (e >= S ? 1 : 0), e >= S ? e - S : e) = ph;
    nfilt++;
  }

  // ---------------------------------------------------------------- results
  __device__ void write_results() {
    const int nv = 2 * N + NX * (N + 1);
    p.u0[2 * (size_t)b + 0] = at(L::U + 0, 0);
    p.u0[2 * (size_t)b + 1] = at(L::U + 1, 0);
    p.cost[b] = status == 4 ? nan("") : fobj;
    p.status[b] = status;
    p.iters[b] = it;
    if (p.z_out) {
      double *z = p.z_out + (size_t)b * nv;
      for (int k = 0; k < N; k++) { z[2 * k] = at(L::U + 0, k); z[2 * k + 1] = at(L::U + 1, k); }
      for (int k = 0; k <= N; k++)
        for (int i = 0; i < NX; i++) z[2 * N + NX * k + i] = at(L::X + i, k);
    }
    if (p.lam_out) {
      double *l = p.lam_out + (size_t)b * NX * (N + 1);
      for (int k = 0; k <= N; k++)
        for (int i = 0; i < NX; i++) l[NX * k + i] = at(L::LAM + i, k) / sigma;
    }
    if (p.lam_g_out) {
      const int n_obs_st = MO > 0 ? N : 0;
      double *g = p.lam_g_out + (size_t)b * (NX * (N + 1) + NR * (N - 1) + MO * n_obs_st);
      const double rs = 1.0 / sigma;
      for (int k = 0; k <= N; k++) {
        for (int i = 0; i < NX; i++) g[NX * k + i] = at(L::LAM + i, k) / sigma;
        if (has_rate(k))
          for (int r = 0; r < NR; r++) g[NX * (N + 1) + NR * (k - 1) + r] = at(L::LR + r, k) * rs;
        if (has_obs(k))
          for (int j = 0; j < MO; j++) g[NX * (N + 1) + NR * (N - 1) + MO * k + j] = at(L::LO + j, k) * rs;
      }
    }
    if (p.lam_x_out) {
      double *lx = p.lam_x_out + (size_t)b * nv;
      const double rs = 1.0 / sigma;
      for (int k = 0; k <= N; k++) {
        if (k < N)
          for (int i = 0; i < 2; i++) lx[2 * k + i] = (at(L::ZUU + i, k) - at(L::ZLU + i, k)) * rs;
        for (int i = 0; i < NX; i++) lx[2 * N + NX * k + i] = 0.0;
        for (int bb = 0; bb < NBX; bb++) lx[2 * N + NX * k + bx(bb)] = (at(L::ZUX + bb, k) - at(L::ZLX + bb, k)) * rs;
      }
    }
    if (to_resto && p.resto_list) {  // published after the (provisional) results, which the restoration pass overwrites
      __threadfence();
      p.resto_list[atomicAdd(p.resto_sync, 1)] = b;
    }
  }

  __device__ __forceinline__ void give_up() {
    status = err_last <= MPCB_ACCEPTABLE_TOL ? 1 : 3;
    to_resto = status == 3;
    state = LANE_DONE;
  }
};

// Persistent kernel: every lane owns one workspace slot and pulls scenarios from the queue until it is empty.
#ifndef MPCB_LANE_BLOCK
#define MPCB_LANE_BLOCK 128
#endif
#ifndef MPCB_LANE_MIN_BLOCKS
#define MPCB_LANE_MIN_BLOCKS 3  // register budget 65536 / (128 * blocks) per thread: 2 -> 254 registers, 3 -> 168 (measured best), 4 -> 128
#endif
#ifndef MPCB_LANE_TRIALS_PER_ROUND
#define MPCB_LANE_TRIALS_PER_ROUND 1  // trial points a backtracking lane evaluates before the round moves on (4 measured slower: the warp waits)
#endif
template <int NR, int MO, bool RK4 = false>
__global__ void __launch_bounds__(MPCB_LANE_BLOCK, RK4 ? 2 : MPCB_LANE_MIN_BLOCKS) kin_lane_kernel(const __grid_constant__ KParams p, double *ws, size_t nslot) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  LaneSolver<NR, MO, RK4> s(p, ws, nslot, slot);
  using Kkt = typename LaneSolver<NR, MO, RK4>::Kkt;
  const double tol = p.tol;
  bool queue_empty = false;
  for (;;) {
    // ---- refill: a finished (or fresh) lane takes the next scenario
    if (s.state == LANE_DONE) { s.write_results(); s.state = LANE_IDLE; }
    if (s.state == LANE_IDLE && !queue_empty) {
      int q = atomicAdd(p.counter, 1);
      if (q >= p.B) queue_empty = true;
      else {
        s.b = p.order ? p.order[q] : q;
        s.init_iterate();
        s.state = LANE_EVAL;
      }
    }
    if (__all_sync(0xffffffffu, s.state == LANE_IDLE)) break;
    // ---- evaluate: the first point of a solve, or a line-search trial point (a rejected trial is followed by the next
    // one right away, a few times: a failing line search halves its step ~45 times before it gives up)
    for (int tr = 0; tr < MPCB_LANE_TRIALS_PER_ROUND; tr++) {
    if (!__any_sync(0xffffffffu, s.state == LANE_EVAL)) break;
    if (s.state == LANE_EVAL) {
      s.eval_point(s.a, s.trial ? s.cur ^ 1 : s.cur);
      if (s.trial) {
        double ph_t = s.sigma * s.f_e - s.mu * s.bar_e + MPCB_KAPPA_D * s.mu * s.lin_e;
        bool fin = isfinite(s.th_e) && isfinite(ph_t);
        bool accepted = false, armijo = false;
        if (fin && !s.filter_blocks(s.th_e, ph_t)) {
          bool sw = s.gd < 0 && s.a * s.pgd > s.pth;
          if (s.theta <= s.theta_min && sw) {
            if (ph_t <= s.phi + MPCB_ETA_PHI * s.a * s.gd + 10 * MPCB_DBL_EPS * fabs(s.phi)) { accepted = true; armijo = true; }
          } else if (s.th_e <= (1 - MPCB_GAMMA_THETA) * s.theta || ph_t <= s.phi - MPCB_GAMMA_PHI * s.theta + 10 * MPCB_DBL_EPS * fabs(s.phi)) {
            accepted = true;
          }
        }
        if (!accepted) {
          s.a *= 0.5;
          if (s.a < s.a_min) s.give_up();
        } else {
          if (!armijo) s.filter_add((1 - MPCB_GAMMA_THETA) * s.theta, s.phi - MPCB_GAMMA_PHI * s.theta);
          s.state = LANE_ACCEPT;
        }
      } else {
        s.theta = s.th_e; s.fobj = s.f_e; s.bar = s.bar_e; s.lin = s.lin_e;
        if (!isfinite(s.theta) || !isfinite(s.bar)) { s.status = 4; s.state = LANE_DONE; }
        else {
          s.theta_min = 1e-4 * fmax(1.0, s.theta);
          s.theta_max = 1e4 * fmax(1.0, s.theta);
          s.state = LANE_KKT;
        }
      }
    }
    }
    // ---- accept the trial point
    if (s.state == LANE_ACCEPT) {
      s.accept_step(s.a, s.a_dual);
      s.it++;
      s.trial = false;
      s.theta = s.th_e; s.fobj = s.f_e; s.bar = s.bar_e; s.lin = s.lin_e;
      s.state = LANE_KKT;
    }
    // ---- KKT error, convergence test, barrier update
    if (s.state == LANE_KKT) {
      Kkt kk;
      s.kkt_pieces(kk);
      double co0;
      double err0 = s.kkt_error(kk, 0.0, co0);
      s.err_last = err0;
      if (err0 <= tol && kk.dual <= MPCB_DUAL_INF_TOL && kk.prim <= MPCB_CONSTR_VIOL_TOL && co0 <= MPCB_COMPL_INF_TOL) { s.status = 0; s.state = LANE_DONE; }
      else if (s.it >= p.max_iter) { s.status = 2; s.state = LANE_DONE; }
      else {
        double co;
        while (s.kkt_error(kk, s.mu, co) <= MPCB_KAPPA_EPS * s.mu && s.mu > tol / 10) {
          s.mu = fmax(tol / 10, fmin(MPCB_KAPPA_MU * s.mu, s.mu * sqrt(s.mu)));
          s.tau = fmax(MPCB_TAU_MIN, 1 - s.mu);
          s.nfilt = 0;
        }
        s.phi = s.sigma * s.fobj - s.mu * s.bar + MPCB_KAPPA_D * s.mu * s.lin;
        s.dw = 0.0;
        s.tried0 = false;
        s.state = LANE_NEWTON;
      }
    }
    // ---- Newton step with IPOPT's inertia-correction schedule (one factorisation attempt per round)
    if (s.state == LANE_NEWTON) {
      if (RK4 ? s.backward_rk4() : s.backward()) {
        if (s.dw > 0.0) s.dw_last = s.dw;
        double a_max;
        s.forward(a_max, s.a_dual, s.gd);
        s.adjoint();
        s.pgd = 0.0; s.pth = 0.0;
        if (s.gd < 0 && s.theta <= s.theta_min) { s.pgd = d_pow(-s.gd, MPCB_S_PHI); s.pth = d_pow(s.theta, MPCB_S_THETA); }
        if (s.gd < 0 && s.theta <= s.theta_min) {
          s.a_min = MPCB_GAMMA_THETA;
          if (s.theta > 0) {
            s.a_min = fmin(s.a_min, MPCB_GAMMA_PHI * s.theta / (-s.gd));
            s.a_min = fmin(s.a_min, s.pth / s.pgd);
          }
        } else if (s.gd < 0) {
          s.a_min = fmin(MPCB_GAMMA_THETA, MPCB_GAMMA_PHI * s.theta / (-s.gd));
        } else {
          s.a_min = MPCB_GAMMA_THETA;
        }
        s.a_min = fmax(MPCB_GAMMA_ALPHA * s.a_min, 1e-14);
        s.a = a_max;
        s.trial = true;
        s.state = LANE_EVAL;
        if (s.a < s.a_min) s.give_up();
      } else {
        if (!s.tried0) {
          s.tried0 = true;
          s.dw = s.dw_last == 0.0 ? MPCB_DW_FIRST : fmax(MPCB_DW_MIN, MPCB_KW_MINUS * s.dw_last);
        } else {
          s.dw *= s.dw_last == 0.0 ? MPCB_KW_PLUS_FIRST : MPCB_KW_PLUS;
          if (s.dw > MPCB_DW_MAX) s.give_up();
        }
      }
    }
  }
  if (p.resto_sync) {  // producer side of the restoration hand-over: this warp adds no more entries
    __threadfence();
    if ((threadIdx.x & 31) == 0) atomicAdd(p.resto_sync + 2, 1);
  }
}

}  // namespace mpcb
