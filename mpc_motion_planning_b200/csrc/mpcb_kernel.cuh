// mpcb200 device code: one scenario per warp, FP64 primal-dual interior point with a
// stage-wise Riccati recursion.  sm_100a only.
//
// What this replaces in the reference (PKG = CasaDi_MPC_Optimize_Multishoot):
//   * CasADi SX graph + AD of the NLP built in MPC_optimize.optimize_problem
//     (PKG/MPC_CBF_optimize_kin.py:136-255, _kin_pre.py:136-261)
//       -> kin_point / build_qp below (hand-derived first and second derivatives)
//   * IPOPT + MUMPS behind ca.nlpsol / solver(...)  (PKG/MPC_CBF_optimize_kin.py:251-254,
//     PKG/main_cbf_kin_c_sim.py:100)
//       -> Solver::run: barrier loop, filter line search, inertia correction;
//          riccati_backward/forward/adjoint replace the sparse LDL^T of the KKT matrix.
//
// Data layout: one record of NF doubles per stage in shared memory (stage-major, NF odd):
// field offsets are compile-time immediates of LDS/STS, the stage-parallel phases (lane =
// stage) are bank-conflict free and the serial recursions read broadcasts.
//
// Code-size discipline: the v1 kernel was 17k SASS instructions and spent most of its time
// in instruction-cache misses (profiles/r01_v1_*).  Transcendentals are therefore called
// through __noinline__ wrappers, and the Riccati sweep is written out for the sparsity of
// the kinematic model (A = I + T df/dx has 5 off-diagonal entries, B has 2) so that the whole
// serial loop body fits the L0/L1 instruction caches.
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
/* IPOPT returns Solved_To_Acceptable_Level when the line search fails at a point that meets the
 * acceptable tolerances.  The reference sets acceptable_tol = tol = 1e-8, which can never trigger;
 * IPOPT's default 1e-6 is used for this exit only (status MPCB_ACCEPTABLE). */
#define MPCB_ACCEPTABLE_TOL 1e-6
#define MPCB_FILTER_SLOTS 4 /* 4 x 32 lanes = 128 filter entries */
/* restoration phase (IPOPT: resto_penalty_parameter, required_infeasibility_reduction, bound_mult_reset_threshold) */
#define MPCB_RESTO_RHO 1000.0
#define MPCB_RESTO_KAPPA 0.9
#define MPCB_BOUND_MULT_RESET 1000.0

struct KParams {
  int B, N, obs_mode, du0_cost, init_mode, max_iter, obs_input;
  int ref_mode;   // 1: xs holds [B][N][nx] per-stage cost targets (kinematic kernels)
  double cbf_g1;  // 1 - gamma of the discrete-time CBF rows (obs_mode 3)
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
  double *lam_g_out, *lam_x_out;  // optional: res['lam_g'] [B][rows of g], res['lam_x'] [B][nv]
  int32_t *status, *iters;
  double *slab;    // per-resident-block scratch in global memory (kinematic kernels)
  int *counter;    // work queue head (persistent kernels)
  const int32_t *order;  // optional [B] permutation: queue position -> scenario index
  double *trace;   // optional [B][trace_rows][8] per-iteration log (mu, theta, err, dual, prim, compl, alpha, dw)
  int trace_rows;
  // Restoration: the main kernels carry no restoration code.  A scenario whose line search fails is appended to
  // resto_list (producer side, RS = false) and solved again, from its start point, by the restoration-capable sibling
  // kernel (consumer side, RS = true).  The sibling is launched right behind the main kernel as a programmatic
  // dependent launch: its blocks become resident as the main kernel's last wave releases the SMs, take tickets, and
  // wait for list entries until every main warp has reported that its queue is empty.
  //   resto_sync[0] = entries reserved, [1] = consumer tickets handed out, [2] = main warps that have finished
  int32_t *resto_list;   // [B], preset to -1
  int *resto_sync;
  int resto_consumer;    // this launch is the restoration pass
  int main_warps;        // warps of the main launch (consumer side: when resto_sync[2] reaches it the list is complete)
  int resto_max_calls;
  int restoration;       // RS kernels: enter the restoration phase in place (0: end with status 3 like the kernels without it)
  int *debug_errors;     // -DMPCB_DEBUG_SLOTS builds: count of shared-memory slots read while another phase's data was in them
  int debug_selftest;    // ... make one expectation wrong on purpose (the checker must trip)
};

// ------------------------------------------------------------------------------------
// out-of-line math (one copy each in the instruction stream)
// ------------------------------------------------------------------------------------
extern __shared__ double g_smem[];

// Reciprocal from the MUFU seed (about 20 bits) and two Newton steps: ~1 ulp, 5 instructions,
// no slow-path call (IEEE division costs a 77-instruction subroutine per use, see
// profiles/r01_v3_*).  Arguments here are bound gaps and pivots: positive normal numbers.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// l1 penalty rho*(p+n) on a row residual r = p - n, with p, n >= 0 held on their central path for the barrier
// parameter mu (IPOPT's closed form for the start of its restoration phase, Waechter & Biegler 2006 eq. 33, applied
// at every restoration iterate): psi' = rho - mu/p is the row's multiplier, psi'' = mu/(p^2+n^2) its curvature.
struct Psi { double v, d1, d2; };
static __device__ __noinline__ Psi d_psi(double r, double mu, bool want_value) {
  const double rho = MPCB_RESTO_RHO;
  const double b = mu * r / (2 * rho), q = sqrt(mu * mu + (rho * r) * (rho * r)) / (2 * rho);
  const double an = (mu - rho * r) / (2 * rho), ap = (mu + rho * r) / (2 * rho);
  const double n = an >= 0 ? an + q : b / (q - an);
  const double pp = ap >= 0 ? ap + q : -b / (q - ap);
  Psi o;
  o.v = want_value ? rho * (pp + n) - mu * log(pp * n) : 0.0;
  o.d1 = rho - mu / pp;
  o.d2 = mu / (pp * pp + n * n);
  return o;
}

static __device__ __noinline__ double d_log(double x) { return log(x); }
static __device__ __noinline__ double d_pow(double x, double y) { return pow(x, y); }
static __device__ __noinline__ double3 d_trig3(double phi, double delta) {  // (sin phi, cos phi, tan delta)
  double s, c;
  sincos(phi, &s, &c);
  return make_double3(s, c, tan(delta));
}
__device__ __forceinline__ void d_trig(double phi, double delta, double *s, double *c, double *t) {
  double3 r = d_trig3(phi, delta);
  *s = r.x; *c = r.y; *t = r.z;
}

// The ONE block barrier of the solve kernels.  Protocol: every warp of a block arrives here once per round, whatever it is
// doing - a working warp at the top of each interior-point iteration with done = 0 (from Solver::run), a warp whose queue
// is empty with done = 1 (from the drain loop at the end of the kernel) - and the block retires when a round's votes are
// all 1.  Both call sites go through this single out-of-line function, so all warps execute the same barrier instruction
// (bar.red.and on barrier 0 counts arriving warps; per-warp arrival counts differ only by WHICH site called).
static __device__ __noinline__ int block_iteration_vote(int done) { return __syncthreads_and(done); }

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

// Kinematic bicycle, x=[x,y,phi,vx], u=[df,ax]   (PKG/MPC_CBF_optimize_kin.py:153-156)
struct KinModel {
  static constexpr int NX = 4;
  static constexpr int NBX = 2;  // bounded state components: y, vx  (PKG/..._kin.py:97-105)
  __device__ static __forceinline__ constexpr int bx(int i) { return i == 0 ? 1 : 3; }
};

// ------------------------------------------------------------------------------------
// shared-memory layout of one scenario (rows of S = N+1 doubles)
// ------------------------------------------------------------------------------------
// FH: the stage Hessian Hxx is stored full (10 entries) - needed by the discrete-time CBF rows,
// whose rank-one barrier term couples all four states of a stage (see build_qp).
// GS: the step (dx, du) lives in the global slab instead of shared memory.  It costs ~2 % at N = 50
// (the stage-parallel phases read it through L2) but frees 6 doubles per stage, which buys resident
// warps at long horizons (N = 100: 8 warps per SM instead of 6, +12 %); the host picks per horizon.
// AS: everything in shared memory, no slab (the small-batch variant: with a handful of scenarios per SM
// occupancy is irrelevant and the L2 round trips of the stage-parallel phases are the latency).
// RS: the restoration-capable sibling keeps the reference point z_R of the proximity term in the slab as well.
template <int NR, int MO, bool FH = false, bool GS = false, bool AS = false, bool RS = false>
struct KinLayout {
  static_assert(!(GS && AS), "the step lives in the slab or everything lives in shared memory");
  static constexpr int NX = 4, NBX = 2;
  static constexpr int NR_ROWS = NR, MO_ROWS = MO;
  // ---- shared memory: the working set of the serial sweeps (one record of NF doubles per stage)
  static constexpr int CDEF = 0;         // c_0 = X0 - x0, c_k = defect into stage k
  static constexpr int LAMP = CDEF;      // alias: new dynamics multipliers (written after the forward sweep)
  static constexpr int JAC = CDEF + NX;  // a02 a03 a12 a13 a23 b2  (A = I + T df/dx, B = T df/du)
  static constexpr int HXX = JAC + 6;    // h00 h01 h11 h22 h23 h33 [h02 h03 h12 h13 when FH]
  static constexpr int NH = FH ? 10 : 6;
  static constexpr int HUX = HXX + NH;   // d2L/(d delta d v)
  static constexpr int GX = HUX + 1;
  // 14-slot region: [HUU(2) EE(2) GU(2) TK(2) -(6)] before the backward sweep of a stage,
  // the Riccati gains [KX(8) KW(4) KK(2)] after it, the slack steps after the forward sweep
  static constexpr int R14 = GX + NX;
  static constexpr int HUU = R14, EE = R14 + 2, GU = R14 + 4, TK = R14 + 6;
  static constexpr int KX = R14, KW = R14 + 8, KK = R14 + 12;
  static constexpr int DSR = R14, LRP = DSR + NR, DSO = LRP + NR, LOP = DSO + MO;
  static constexpr int CDEFT = R14 + 10;  // defects of the line-search trial point (moved to CDEF on acceptance)
  static_assert(2 * NR + 2 * MO <= 10, "slack steps and trial defects must fit the gain region");
  static constexpr int NSH = R14 + 14 + (GS ? 0 : NX + 2);
  // stage-major storage: element (field, k) lives at k*NF + field.  NF is odd so that the
  // stage-parallel phases (lane = stage, stride NF doubles) touch 16 distinct even banks per
  // half-warp: conflict free; the serial sweeps read broadcasts.
  static constexpr int NF = NSH | 1;
  // ---- global memory (L2-resident slab of the resident block): everything only the
  // stage-parallel phases touch - the primal-dual iterate and the staged obstacle trajectory.
  // Rows of SG doubles per field, stage index fastest: lane = stage accesses are coalesced.
  static constexpr int G0 = 64;
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
  static constexpr int OCX = LO + MO;   // obstacle trajectory: centre and 1/semi-axis^2 per step
  static constexpr int OCY = OCX + MO;
  static constexpr int ISX = OCY + MO;
  static constexpr int ISY = ISX + MO;
  static constexpr int XR = ISY + MO;   // per-stage cost target (only with ref_mode)
  // the step: written by lane 0 in the forward sweep, read by the stage-parallel phases only
  static constexpr int DX = GS ? XR + NX : R14 + 14;
  static constexpr int DU = DX + NX;
  static constexpr int NG0 = XR + NX + (GS ? NX + 2 : 0) - G0;
  static constexpr int XRS = G0 + NG0;  // restoration: z_R (states, controls)
  static constexpr int URS = XRS + NX;
  static constexpr int FTO = URS + 2;   // restoration: the original filter (MPCB_FILTER_SLOTS rows, lane-indexed)
  static constexpr int FPO = FTO + 4;
  static constexpr int NG = NG0 + (RS ? NX + 2 + 8 : 0);
  static constexpr int SG = 132;        // row stride (>= MPCB_NMAX + 1)
  static constexpr int NFA = AS ? ((NF + NG) | 1) : NF;  // record length in shared memory
#ifdef MPCB_DEBUG_SLOTS
  // debug build: one owner tag per slot of the aliased regions, kept behind the records of the warp
  __host__ __device__ static constexpr size_t bytes(int N) { return (sizeof(double) + sizeof(double)) * (size_t)NFA * (size_t)(N + 1); }
#else
  __host__ __device__ static constexpr size_t bytes(int N) { return sizeof(double) * (size_t)NFA * (size_t)(N + 1); }
#endif
  __host__ __device__ static constexpr size_t slab_doubles() { return AS ? 0 : (size_t)NG * SG; }
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
__device__ __forceinline__ double clampz(double z, double mu, double rgap) {
  // kappa_sigma safeguard; rgap = 1/gap
  return fmax(fmin(z, MPCB_KAPPA_SIGMA * mu * rgap), mu * rgap * (1.0 / MPCB_KAPPA_SIGMA));
}

// ------------------------------------------------------------------------------------
// the solver: one warp = one scenario (kinematic model family)
// ------------------------------------------------------------------------------------
template <int NR, int MO, int OBS_MODE, bool GS = false, bool AS = false, bool RS_ = false>
struct KinSolver {
  static constexpr bool DCBF = OBS_MODE == 3;  // rows h(X_{k+1};obs_k) - (1-gamma) h(X_k;obs_k) >= 0
  static constexpr bool RS = RS_;              // restoration-capable build (the second-pass kernel)
  using L = KinLayout<NR, MO, DCBF, GS, AS, RS>;
  static constexpr int NX = 4, NBX = 2;
  static constexpr bool ROWS_INTERLEAVED = false;  // g = [init; defects][rate rows][obstacle rows]
  __device__ static __forceinline__ constexpr int bx(int i) { return KinModel::bx(i); }

  const KParams &p;
  double *gs;  // this warp's slab in global memory
  int woff;    // this warp's offset into the block's shared memory (in doubles)
  int &tick;   // this warp's running iteration count (across scenarios) for the block barrier
  int N, lane;
  double sigma;
  double x0[NX], xs[NX];
  // restoration phase (RS builds): min zeta/2 ||D_R (z - z_R)||^2 + sum_rows psi_mu(row - s) s.t. the dynamics and the
  // bounds; `resto` switches every phase below between the two problems (uniform per warp)
  bool resto = false;
  double zeta = 0.0;
  double o_thr = 0.0, o_f = 0.0;  // restoration: row part of the original infeasibility, original objective

  __device__ KinSolver(const KParams &p_, double *gs_, int woff_, int &tick_, int lane_)
      : p(p_), gs(gs_), woff(woff_), tick(tick_), N(p_.N), lane(lane_) {}

  // ---- sanitizer substitute (compute-sanitizer is closed on this pool): the shared-memory record reuses slots between
  // phases - LAMP aliases CDEF; the 14-slot region holds [HUU EE GU TK] -> the Riccati gains -> the slack steps and the
  // trial point's defects.  With -DMPCB_DEBUG_SLOTS every store into those slots records WHICH data it is and every load
  // states which data it expects; a mismatch is counted in KParams.debug_errors (tests/test_gpu_parity.py runs one
  // solve per kernel family on this build and asserts zero).  In the regular build both calls compile to nothing.
  enum { TAG_NONE = 0, TAG_DEFECT = 1, TAG_NEWLAM = 2, TAG_QP = 3, TAG_GAINS = 4, TAG_STEPS = 5, TAG_TRIALDEF = 6 };
#ifdef MPCB_DEBUG_SLOTS
  __device__ __forceinline__ double &tag_at(int field, int k) {
    const int idx = woff + k * L::NFA + (AS && field >= L::G0 ? L::NF + (field - L::G0) : field);
    return g_smem[idx + (int)(L::bytes(N) / (2 * sizeof(double))) * (int)(blockDim.x >> 5)];
  }
  __device__ __forceinline__ void own(int field, int n, int k, int tag) {
    for (int i = 0; i < n; i++) tag_at(field + i, k) = (double)tag;
  }
  __device__ __forceinline__ void expect(int field, int n, int k, int tag) {
    for (int i = 0; i < n; i++)
      if (tag_at(field + i, k) != (double)tag && p.debug_errors) atomicAdd(p.debug_errors, 1);
  }
#else
  __device__ __forceinline__ void own(int, int, int, int) {}
  __device__ __forceinline__ void expect(int, int, int, int) {}
#endif
  __device__ __forceinline__ bool in_resto() const { return RS && resto; }
  __device__ static __forceinline__ double dr2(double v) { double m = fmax(1.0, fabs(v)); return 1.0 / (m * m); }
  // A row `row(z) - s` condensed into its stage: new row multiplier lam+ = D (J dz) + tt.
  // regular: D = Sigma_s + dw, tt = D residual + slack barrier gradient;
  // restoration: the row is penalised by psi: D = 1/(1/Ds + 1/psi''), tt = D (gs/Ds + psi'/psi'')
  __device__ __forceinline__ void row_cond(double Ds, double gsl, double res, double mu, double &D, double &tt) const {
    if (in_resto()) {
      Psi ps = d_psi(res, mu, false);
      D = Ds * ps.d2 / (Ds + ps.d2);
      tt = D * (gsl / Ds + ps.d1 / ps.d2);
    } else {
      D = Ds;
      tt = Ds * res + gsl;
    }
  }
  __device__ __forceinline__ void row_step(double Ds, double gsl, double res, double mu, double Jdz, double &ds, double &lnew) const {
    if (in_resto()) {
      double D, tt;
      row_cond(Ds, gsl, res, mu, D, tt);
      lnew = D * Jdz + tt;
      ds = (lnew - gsl) / Ds;
    } else {
      ds = Jdz + res;
      lnew = Ds * ds + gsl;
    }
  }

  // field ids are compile-time constants at (almost) every use, so the space test folds away
  __device__ __forceinline__ double &at(int field, int k) {
    if (AS) return g_smem[woff + k * L::NFA + (field >= L::G0 ? L::NF + (field - L::G0) : field)];
    return field >= L::G0 ? gs[(field - L::G0) * L::SG + k] : g_smem[woff + k * L::NF + field];
  }
  __device__ __forceinline__ bool has_rate(int k) const { return NR > 0 && k >= 1 && k <= N - 1; }
  __device__ __forceinline__ bool has_obs(int k) const { return (OBS_MODE == 1 || DCBF) && k <= N - 1; }
  // cost target of stage k: xs, or the per-stage reference ref_X = aa*ref_state[k+1] + (1-aa)*xs
  // (PKG/MPC_CBF_optimize_kin.py:194-199) that the host passes as [B][N][nx]
  __device__ __forceinline__ double xref(int i, int k) { return p.ref_mode ? at(L::XR + i, k) : xs[i]; }

  // gradient of the unscaled objective wrt u_k[i]  (PKG/MPC_CBF_optimize_kin.py:199-205)
  __device__ __forceinline__ double grad_u(int k, int i, double uk, double ukm1, double ukp1) const {
    double v = 2 * p.R[i] * uk;
    if (k > 0) v += 2 * p.DR[i] * (uk - ukm1);
    else if (p.du0_cost) v += 2 * p.DR[i] * uk;
    if (k + 1 <= N - 1) v -= 2 * p.DR[i] * (ukp1 - uk);
    return v;
  }

  // ---------------------------------------------------------------- point evaluation
  // constraint residual 1-norm, objective and barrier pieces at z + alpha*dz.  The dynamics
  // Jacobian of the point is always kept (the old one is dead once the step is known) and its
  // defects go to CDEF (first iterate) or CDEFT (trial point; accept_step moves them), so an
  // accepted trial point needs no second evaluation.
  double rmu = 0.0;  // barrier parameter seen by psi (set by the run loop before every phase that needs it)
  __device__ __forceinline__ void eval_point(double alpha, bool fresh, double &theta, double &fobj, double &bar, double &lin) {
    const int cdst = fresh ? L::CDEF : L::CDEFT;
    const double rL = 1.0 / p.Veh_l;  // one division per call instead of three per stage
    double th = 0, fo = 0, br = 0, ln = 0;
    double thr = 0, fr = 0;  // RS builds: row part of theta kept apart, restoration objective
    const double mu_psi = rmu;
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], uk[2] = {0, 0}, xnext[2] = {0, 0};
      double gp = 1.0;  // product of the bound gaps of this stage: sum of logs = log of the product
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k) + alpha * at(L::DX + i, k);
      if (k == 0) {
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double c0 = xk[i] - x0[i];
          th += fabs(c0);
          at(cdst + i, 0) = c0;
        }
        own(cdst, NX, 0, fresh ? TAG_DEFECT : TAG_TRIALDEF);
      }
      if (in_resto()) {
#pragma unroll
        for (int i = 0; i < NX; i++) { double xr_ = at(L::XRS + i, k), e = xk[i] - xr_; fr += 0.5 * zeta * dr2(xr_) * e * e; }
      }
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) uk[i] = at(L::U + i, k) + alpha * at(L::DU + i, k);
        if (in_resto()) {
#pragma unroll
          for (int i = 0; i < 2; i++) { double ur_ = at(L::URS + i, k), e = uk[i] - ur_; fr += 0.5 * zeta * dr2(ur_) * e * e; }
        }
        double s, c, t;
        d_trig(xk[2], uk[0], &s, &c, &t);
        double f[NX];
        f[0] = xk[3] * c;                 // PKG/MPC_CBF_optimize_kin.py:153-156
        f[1] = xk[3] * s;
        f[2] = xk[3] * t * rL;
        f[3] = uk[1];
        {
          at(L::JAC + 0, k) = p.T * (-xk[3] * s);                       // a02 = T d f0/d phi
          at(L::JAC + 1, k) = p.T * c;                                  // a03 = T d f0/d v
          at(L::JAC + 2, k) = p.T * (xk[3] * c);                        // a12
          at(L::JAC + 3, k) = p.T * s;                                  // a13
          at(L::JAC + 4, k) = p.T * (t * rL);                      // a23
          at(L::JAC + 5, k) = p.T * (xk[3] * (1.0 + t * t) * rL);  // b2  = T d f2/d delta
        }
#pragma unroll
        for (int i = 0; i < NX; i++) {
          double xn = at(L::X + i, k + 1) + alpha * at(L::DX + i, k + 1);
          if (i < 2) xnext[i] = xn;
          double d = xn - (xk[i] + p.T * f[i]);
          th += fabs(d);
          at(cdst + i, k + 1) = d;
          own(cdst + i, 1, k + 1, fresh ? TAG_DEFECT : TAG_TRIALDEF);
          double e = xk[i] - xref(i, k);
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
        int i = KinModel::bx(b);
        gp *= (xk[i] - p.x_lo[i]) * (p.x_hi[i] - xk[i]);
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          double um = at(L::U + ci, k - 1) + alpha * at(L::DU + ci, k - 1);
          double ukc = ci == 0 ? uk[0] : uk[1];
          // the slack step lives in the gain region, which holds gains while alpha == 0
          if (alpha != 0.0) expect(L::DSR + r, 1, k, TAG_STEPS);
          double s = at(L::SR + r, k) + (alpha != 0.0 ? alpha * at(L::DSR + r, k) : 0.0);
          if (RS) thr += fabs(ukc - um - s); else th += fabs(ukc - um - s);
          if (in_resto()) fr += d_psi(ukc - um - s, mu_psi, true).v;
          gp *= (s - p.rate_lo[r]) * (p.rate_hi[r] - s);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double d = dx * dx * at(L::ISX + j, k) + dy * dy * at(L::ISY + j, k) - 1.0;  // PKG/..._kin.py:244,247
          if (DCBF) {  // gamma*h_func + h_dot, PKG/..._kin.py:245-248
            double ex = xnext[0] - at(L::OCX + j, k), ey = xnext[1] - at(L::OCY + j, k);
            d = (ex * ex * at(L::ISX + j, k) + ey * ey * at(L::ISY + j, k) - 1.0) - p.cbf_g1 * d;
          }
          if (alpha != 0.0) expect(L::DSO + j, 1, k, TAG_STEPS);
          double s = at(L::SO + j, k) + (alpha != 0.0 ? alpha * at(L::DSO + j, k) : 0.0);
          if (RS) thr += fabs(d - s); else th += fabs(d - s);
          if (in_resto()) fr += d_psi(d - s, mu_psi, true).v;
          gp *= s - p.obs_lo;
          ln += s - p.obs_lo;
        }
      }
      br += d_log(gp);
    }
    theta = warp_sum(th);
    fobj = warp_sum(fo);
    bar = warp_sum(br);
    lin = warp_sum(ln);
    if (RS) {
      thr = warp_sum(thr);
      if (resto) {  // the rows are penalised, not constrained: theta = the equalities, objective = proximity + penalties
        o_thr = thr;
        o_f = fobj;
        fobj = warp_sum(fr);
      } else {
        theta += thr;
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- KKT error pieces
  struct Kkt { double dual, prim, cmin, cmax, sum_lam, sum_z; };

  __device__ __forceinline__ void kkt_pieces(Kkt &o) {
    double dual = 0, prim = 0, cmin = INFINITY, cmax = -INFINITY, sl = 0, sz = 0;
#define MPCB_COMPL(gap, mult) do { double p_ = (gap) * (mult); cmin = fmin(cmin, p_); cmax = fmax(cmax, p_); sz += (mult); } while (0)
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX], lam[NX], l1[NX] = {0, 0, 0, 0};
#pragma unroll
      for (int i = 0; i < NX; i++) {
        xk[i] = at(L::X + i, k);
        lam[i] = at(L::LAM + i, k);
        expect(L::CDEF + i, 1, k, TAG_DEFECT);
        prim = fmax(prim, fabs(at(L::CDEF + i, k)));
        sl += fabs(lam[i]);
      }
      double rx[NX] = {lam[0], lam[1], lam[2], lam[3]};
      double a02 = 0, a03 = 0, a12 = 0, a13 = 0, a23 = 0, b2 = 0;
      if (k < N) {
        a02 = at(L::JAC + 0, k); a03 = at(L::JAC + 1, k); a12 = at(L::JAC + 2, k);
        a13 = at(L::JAC + 3, k); a23 = at(L::JAC + 4, k); b2 = at(L::JAC + 5, k);
#pragma unroll
        for (int i = 0; i < NX; i++) {
          l1[i] = at(L::LAM + i, k + 1);
          if (!in_resto()) rx[i] += sigma * 2 * p.Q[i] * (xk[i] - xref(i, k));
        }
        // - A' lam_{k+1}
        rx[0] -= l1[0];
        rx[1] -= l1[1];
        rx[2] -= l1[2] + a02 * l1[0] + a12 * l1[1];
        rx[3] -= l1[3] + a03 * l1[0] + a13 * l1[1] + a23 * l1[2];
      }
      if (in_resto()) {
#pragma unroll
        for (int i = 0; i < NX; i++) { double xr_ = at(L::XRS + i, k); rx[i] += zeta * dr2(xr_) * (xk[i] - xr_); }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = KinModel::bx(b);
        double zl = at(L::ZLX + b, k), zu = at(L::ZUX + b, k);
        rx[i] += -zl + zu;
        MPCB_COMPL(xk[i] - p.x_lo[i], zl);
        MPCB_COMPL(p.x_hi[i] - xk[i], zu);
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double a = at(L::ISX + j, k), b = at(L::ISY + j, k);
          double d = dx * dx * a + dy * dy * b - 1.0;
          double lo = at(L::LO + j, k), vl = at(L::VLO + j, k), s = at(L::SO + j, k);
          if (DCBF) {
            double ex = at(L::X + 0, k + 1) - at(L::OCX + j, k), ey = at(L::X + 1, k + 1) - at(L::OCY + j, k);
            d = (ex * ex * a + ey * ey * b - 1.0) - p.cbf_g1 * d;
            rx[0] -= p.cbf_g1 * lo * (2 * dx * a);
            rx[1] -= p.cbf_g1 * lo * (2 * dy * b);
          } else {
            rx[0] += lo * (2 * dx * a);
            rx[1] += lo * (2 * dy * b);
          }
          dual = fmax(dual, fabs(-lo - vl));
          if (!in_resto()) prim = fmax(prim, fabs(d - s));
          MPCB_COMPL(s - p.obs_lo, vl);
          sl += fabs(lo);
        }
      }
      if (DCBF && k >= 1) {  // row k-1 also depends on X_k
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double lo = at(L::LO + j, k - 1);
          rx[0] += lo * (2 * (xk[0] - at(L::OCX + j, k - 1)) * at(L::ISX + j, k - 1));
          rx[1] += lo * (2 * (xk[1] - at(L::OCY + j, k - 1)) * at(L::ISY + j, k - 1));
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
          if (in_resto()) { double ur_ = at(L::URS + i, k); r = zeta * dr2(ur_) * (uk - ur_) - zl + zu; }
          r -= (i == 0) ? b2 * l1[2] : p.T * l1[3];  // - B' lam_{k+1}
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
          if (!in_resto()) prim = fmax(prim, fabs(at(L::U + ci, k) - at(L::U + ci, k - 1) - s));
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

  // ---------------------------------------------------------------- condensed QP (stage parallel)
  // Effective stage Hessian/gradient with the slack rows eliminated and the primal
  // regularisation dw applied.
  __device__ __forceinline__ void build_qp(double mu, double dw) {
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double xk[NX];
#pragma unroll
      for (int i = 0; i < NX; i++) xk[i] = at(L::X + i, k);
      double h[NX] = {dw, dw, dw, dw};  // diagonal of Hxx
      double h01 = 0, h23 = 0, hdv = 0, hdd_f = 0;
      double h02 = 0, h03 = 0, h12 = 0, h13 = 0;  // only the discrete-time CBF rows fill these
      double gx[NX] = {0, 0, 0, 0};
      double a02 = 0, a03 = 0, a12 = 0, a13 = 0;
      if (k < N) {
        // - T sum_i lam_i d2 f_i, written with the stored Jacobian entries
        double l0 = at(L::LAM + 0, k + 1), l1 = at(L::LAM + 1, k + 1), l2 = at(L::LAM + 2, k + 1);
        a02 = at(L::JAC + 0, k); a03 = at(L::JAC + 1, k); a12 = at(L::JAC + 2, k); a13 = at(L::JAC + 3, k);
        double a23 = at(L::JAC + 4, k), b2 = at(L::JAC + 5, k);
        double t = a23 * (p.Veh_l / p.T);               // tan(delta)
        double jd = (p.T / p.Veh_l) * (1.0 + t * t);    // T sec^2(delta) / L
        h[2] += l0 * a12 - l1 * a02;                    // -T (-l0 v cos - l1 v sin)
        h23 = l0 * a13 - l1 * a03;                      // -T (-l0 sin + l1 cos)
        hdv = -l2 * jd;
        hdd_f = -2.0 * l2 * b2 * t;
        if (!in_resto()) {
#pragma unroll
          for (int i = 0; i < NX; i++) {
            h[i] += sigma * 2 * p.Q[i];
            gx[i] = sigma * 2 * p.Q[i] * (xk[i] - xref(i, k));
          }
        }
      }
      if (in_resto()) {
#pragma unroll
        for (int i = 0; i < NX; i++) { double xr_ = at(L::XRS + i, k), w_ = zeta * dr2(xr_); h[i] += w_; gx[i] = w_ * (xk[i] - xr_); }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = KinModel::bx(b);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        h[i] += at(L::ZLX + b, k) * rl + at(L::ZUX + b, k) * rh;
        gx[i] += mu * (rh - rl);
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = xk[0] - at(L::OCX + j, k), dy = xk[1] - at(L::OCY + j, k);
          double a = at(L::ISX + j, k), b = at(L::ISY + j, k);
          double d = dx * dx * a + dy * dy * b - 1.0;
          double ox = 2 * dx * a, oy = 2 * dy * b;
          double s = at(L::SO + j, k), rg = fast_rcp(s - p.obs_lo);
          double D = at(L::VLO + j, k) * rg + dw;
          double gs = -mu * rg + MPCB_KAPPA_D * mu;
          double lo = at(L::LO + j, k);
          double tcond = 0.0;
          if (RS) {
            if (DCBF) {
              double ex = at(L::X + 0, k + 1) - at(L::OCX + j, k), ey = at(L::X + 1, k + 1) - at(L::OCY + j, k);
              row_cond(D, gs, ((ex * ex * a + ey * ey * b - 1.0) - p.cbf_g1 * d) - s, mu, D, tcond);
            } else {
              row_cond(D, gs, d - s, mu, D, tcond);
            }
          }
          if (DCBF) {
            // Row k is linear in (dx_k, dx_{k+1}).  The Euler step moves the position by states
            // only (B has no entries in rows 0,1), so with dx_{k+1} = A dx_k + B du_k + b the row
            // becomes v'dx_k + res with v = g_k + A' g_next: a rank-one term on the stage's own
            // states, which keeps the Riccati structure.
            double ex = at(L::X + 0, k + 1) - at(L::OCX + j, k), ey = at(L::X + 1, k + 1) - at(L::OCY + j, k);
            double nx_ = 2 * ex * a, ny_ = 2 * ey * b;
            d = (ex * ex * a + ey * ey * b - 1.0) - p.cbf_g1 * d;
            double v0 = nx_ - p.cbf_g1 * ox, v1 = ny_ - p.cbf_g1 * oy;
            double v2 = nx_ * a02 + ny_ * a12, v3 = nx_ * a03 + ny_ * a13;
            double res = (d - s) - nx_ * at(L::CDEF + 0, k + 1) - ny_ * at(L::CDEF + 1, k + 1);
            double t = D * res + gs;
            if (RS) t = tcond - D * (nx_ * at(L::CDEF + 0, k + 1) + ny_ * at(L::CDEF + 1, k + 1));  // J dz = v'dx_k + g_next'b
            h[0] += D * v0 * v0 - p.cbf_g1 * lo * (2 * a);
            h[1] += D * v1 * v1 - p.cbf_g1 * lo * (2 * b);
            h[2] += D * v2 * v2;
            h[3] += D * v3 * v3;
            h01 += D * v0 * v1; h02 += D * v0 * v2; h03 += D * v0 * v3;
            h12 += D * v1 * v2; h13 += D * v1 * v3; h23 += D * v2 * v3;
            gx[0] += v0 * t; gx[1] += v1 * t; gx[2] += v2 * t; gx[3] += v3 * t;
          } else {
            double t = RS ? tcond : D * (d - s) + gs;
            h[0] += lo * (2 * a) + D * ox * ox;
            h01 += D * ox * oy;
            h[1] += lo * (2 * b) + D * oy * oy;
            gx[0] += ox * t;
            gx[1] += oy * t;
          }
        }
      }
      if (DCBF && k >= 1) {  // curvature of row k-1 in X_k
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double lo = at(L::LO + j, k - 1);
          h[0] += lo * (2 * at(L::ISX + j, k - 1));
          h[1] += lo * (2 * at(L::ISY + j, k - 1));
        }
      }
      if (L::NH == 10) {
        at(L::HXX + 6, k) = h02; at(L::HXX + 7, k) = h03; at(L::HXX + 8, k) = h12; at(L::HXX + 9, k) = h13;
      }
      at(L::HXX + 0, k) = h[0];
      at(L::HXX + 1, k) = h01;
      at(L::HXX + 2, k) = h[1];
      at(L::HXX + 3, k) = h[2];
      at(L::HXX + 4, k) = h23;
      at(L::HXX + 5, k) = h[3];
      at(L::HUX, k) = hdv;
#pragma unroll
      for (int i = 0; i < NX; i++) at(L::GX + i, k) = gx[i];
      if (k < N) {
        double E[2] = {0, 0}, t[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k);
          double g = sigma * 2 * p.R[i] * uk;
          double hd = sigma * 2 * p.R[i] + dw + (i == 0 ? hdd_f : 0.0);
          if (k == 0 && p.du0_cost) {
            hd += sigma * 2 * p.DR[i];
            g += sigma * 2 * p.DR[i] * uk;
          }
          if (in_resto()) {
            double ur_ = at(L::URS + i, k), w_ = zeta * dr2(ur_);
            hd = w_ + dw + (i == 0 ? hdd_f : 0.0);
            g = w_ * (uk - ur_);
          }
          double rl = fast_rcp(uk - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk);
          hd += at(L::ZLU + i, k) * rl + at(L::ZUU + i, k) * rh;
          g += mu * (rh - rl);
          at(L::HUU + i, k) = hd;
          at(L::GU + i, k) = g;
          own(L::HUU + i, 1, k, TAG_QP);
          own(L::GU + i, 1, k, TAG_QP);
          if (k >= 1 && !in_resto()) {
            E[i] = sigma * 2 * p.DR[i];
            t[i] = sigma * 2 * p.DR[i] * (uk - at(L::U + i, k - 1));
          }
        }
        if (has_rate(k)) {
#pragma unroll
          for (int r = 0; r < NR; r++) {
            int ci = p.rate_ctrl[r];
            double s = at(L::SR + r, k);
            double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
            double D = at(L::VLR + r, k) * rl + at(L::VUR + r, k) * rh + dw;
            double gs = mu * (rh - rl);
            double res = at(L::U + ci, k) - at(L::U + ci, k - 1) - s;
            double tt = D * res + gs;
            if (RS) row_cond(D, gs, res, mu, D, tt);
            if (ci == 0) { E[0] += D; t[0] += tt; } else { E[1] += D; t[1] += tt; }
          }
        }
        at(L::EE + 0, k) = E[0];
        at(L::EE + 1, k) = E[1];
        at(L::TK + 0, k) = t[0];
        at(L::TK + 1, k) = t[1];
        own(L::EE, 2, k, TAG_QP);
        own(L::TK, 2, k, TAG_QP);
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- Riccati (serial over stages)
  // Value function V_k(x, w) = 1/2 [x;w]' [P W; W' Q] [x;w] + [px;pw]' [x;w] on the state
  // augmented with the previous control w = u_{k-1}.  Every lane runs the same recursion on
  // broadcast reads; lane 0 stores the gains over the consumed QP slots of the stage.
  // Returns false when some F_uu is not positive definite (wrong inertia).
  __device__ __forceinline__ bool riccati_backward() {
    const double T = p.T;
    double p00 = at(L::HXX + 0, N), p01 = at(L::HXX + 1, N), p11 = at(L::HXX + 2, N), p22 = at(L::HXX + 3, N);
    double p23 = at(L::HXX + 4, N), p33 = at(L::HXX + 5, N), p02 = 0, p03 = 0, p12 = 0, p13 = 0;
    double px0 = at(L::GX + 0, N), px1 = at(L::GX + 1, N), px2 = at(L::GX + 2, N), px3 = at(L::GX + 3, N);
    double w0d = 0, w0a = 0, w1d = 0, w1a = 0, w2d = 0, w2a = 0, w3d = 0, w3a = 0;  // Pxw
    double qdd = 0, qda = 0, qaa = 0, pwd = 0, pwa = 0;                             // Pww, pw
    bool ok = true;
    #pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      const double a02 = at(L::JAC + 0, k), a03 = at(L::JAC + 1, k), a12 = at(L::JAC + 2, k), a13 = at(L::JAC + 3, k);
      const double a23 = at(L::JAC + 4, k), b2 = at(L::JAC + 5, k);
      expect(L::HUU, 8, k, TAG_QP);
      expect(L::CDEF, NX, k + 1, TAG_DEFECT);
      const double Ed = at(L::EE + 0, k), Ea = at(L::EE + 1, k), td = at(L::TK + 0, k), ta = at(L::TK + 1, k);
      const double b0 = -at(L::CDEF + 0, k + 1), b1 = -at(L::CDEF + 1, k + 1), b2_ = -at(L::CDEF + 2, k + 1), b3 = -at(L::CDEF + 3, k + 1);
      // M = P A (columns 0,1 are those of P)
      const double m02 = p02 + a02 * p00 + a12 * p01;
      const double m12 = p12 + a02 * p01 + a12 * p11;
      const double m22 = p22 + a02 * p02 + a12 * p12;
      const double m32 = p23 + a02 * p03 + a12 * p13;
      const double m03 = p03 + a03 * p00 + a13 * p01 + a23 * p02;
      const double m13 = p13 + a03 * p01 + a13 * p11 + a23 * p12;
      const double m23 = p23 + a03 * p02 + a13 * p12 + a23 * p22;
      const double m33 = p33 + a03 * p03 + a13 * p13 + a23 * p23;
      // Fxx = Hxx + A' M (upper triangle)
      const double f00 = at(L::HXX + 0, k) + p00, f01 = at(L::HXX + 1, k) + p01, f11 = at(L::HXX + 2, k) + p11;
      const double f02 = L::NH == 10 ? m02 + at(L::HXX + 6, k) : m02, f03 = L::NH == 10 ? m03 + at(L::HXX + 7, k) : m03;
      const double f12 = L::NH == 10 ? m12 + at(L::HXX + 8, k) : m12, f13 = L::NH == 10 ? m13 + at(L::HXX + 9, k) : m13;
      const double f22 = at(L::HXX + 3, k) + m22 + a02 * m02 + a12 * m12;
      const double f23 = at(L::HXX + 4, k) + m23 + a02 * m03 + a12 * m13;
      const double f33 = at(L::HXX + 5, k) + m33 + a03 * m03 + a13 * m13 + a23 * m23;
      // Fux = Hux + B' M + Pwx A
      const double ud0 = b2 * p02 + w0d, ud1 = b2 * p12 + w1d;
      const double ud2 = b2 * m22 + w2d + a02 * w0d + a12 * w1d;
      const double ud3 = at(L::HUX, k) + b2 * m23 + w3d + a03 * w0d + a13 * w1d + a23 * w2d;
      const double ua0 = T * p03 + w0a, ua1 = T * p13 + w1a;
      const double ua2 = T * m32 + w2a + a02 * w0a + a12 * w1a;
      const double ua3 = T * m33 + w3a + a03 * w0a + a13 * w1a + a23 * w2a;
      // Fuu = Huu + E + Pww + B'PB + B'Pxw + (B'Pxw)'
      const double Fdd = at(L::HUU + 0, k) + Ed + qdd + b2 * (b2 * p22 + 2.0 * w2d);
      const double Fda = qda + b2 * (T * p23) + b2 * w2a + T * w3d;
      const double Faa = at(L::HUU + 1, k) + Ea + qaa + T * (T * p33 + 2.0 * w3a);
      // vectors
      const double Pb0 = px0 + p00 * b0 + p01 * b1 + p02 * b2_ + p03 * b3;
      const double Pb1 = px1 + p01 * b0 + p11 * b1 + p12 * b2_ + p13 * b3;
      const double Pb2 = px2 + p02 * b0 + p12 * b1 + p22 * b2_ + p23 * b3;
      const double Pb3 = px3 + p03 * b0 + p13 * b1 + p23 * b2_ + p33 * b3;
      const double fx0 = at(L::GX + 0, k) + Pb0, fx1 = at(L::GX + 1, k) + Pb1;
      const double fx2 = at(L::GX + 2, k) + Pb2 + a02 * Pb0 + a12 * Pb1;
      const double fx3 = at(L::GX + 3, k) + Pb3 + a03 * Pb0 + a13 * Pb1 + a23 * Pb2;
      const double fud = at(L::GU + 0, k) + td + pwd + b2 * Pb2 + w0d * b0 + w1d * b1 + w2d * b2_ + w3d * b3;
      const double fua = at(L::GU + 1, k) + ta + pwa + T * Pb3 + w0a * b0 + w1a * b1 + w2a * b2_ + w3a * b3;
      const double det = Fdd * Faa - Fda * Fda;
      if (!(Fdd > 0.0) || !(det > 0.0) || !isfinite(det)) { ok = false; break; }
      const double id = fast_rcp(det);
      const double idd = Faa * id, ida = -Fda * id, iaa = Fdd * id;  // Fuu^{-1}
      // gains: u = Kx x + Kw w + kk
      const double kd0 = -(idd * ud0 + ida * ua0), kd1 = -(idd * ud1 + ida * ua1), kd2 = -(idd * ud2 + ida * ua2), kd3 = -(idd * ud3 + ida * ua3);
      const double ka0 = -(ida * ud0 + iaa * ua0), ka1 = -(ida * ud1 + iaa * ua1), ka2 = -(ida * ud2 + iaa * ua2), ka3 = -(ida * ud3 + iaa * ua3);
      const double wdd = idd * Ed, wda = ida * Ea, wad = ida * Ed, waa = iaa * Ea;  // Kw = Fuu^{-1} diag(E)
      const double kkd = -(idd * fud + ida * fua), kka = -(ida * fud + iaa * fua);
      __syncwarp();  // every lane has consumed the QP slots of this stage before lane 0 reuses them
      if (lane == 0) {
        at(L::KX + 0, k) = kd0; at(L::KX + 1, k) = kd1; at(L::KX + 2, k) = kd2; at(L::KX + 3, k) = kd3;
        at(L::KX + 4, k) = ka0; at(L::KX + 5, k) = ka1; at(L::KX + 6, k) = ka2; at(L::KX + 7, k) = ka3;
        at(L::KW + 0, k) = wdd; at(L::KW + 1, k) = wda; at(L::KW + 2, k) = wad; at(L::KW + 3, k) = waa;
        at(L::KK + 0, k) = kkd; at(L::KK + 1, k) = kka;
        own(L::KX, 14, k, TAG_GAINS);
      }
      // value function of stage k
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
    }
    __syncwarp();
    return ok;
  }

  __device__ __forceinline__ void riccati_forward() {
    const double T = p.T;
    double d0 = -at(L::CDEF + 0, 0), d1 = -at(L::CDEF + 1, 0), d2 = -at(L::CDEF + 2, 0), d3 = -at(L::CDEF + 3, 0);
    double vd = 0, va = 0;  // previous control step
    if (lane == 0) { at(L::DX + 0, 0) = d0; at(L::DX + 1, 0) = d1; at(L::DX + 2, 0) = d2; at(L::DX + 3, 0) = d3; }
    #pragma unroll 1
    for (int k = 0; k < N; k++) {
      expect(L::KX, 14, k, TAG_GAINS);
      expect(L::CDEF, NX, k + 1, TAG_DEFECT);
      const double ud = at(L::KK + 0, k) + at(L::KX + 0, k) * d0 + at(L::KX + 1, k) * d1 + at(L::KX + 2, k) * d2 + at(L::KX + 3, k) * d3 +
                        at(L::KW + 0, k) * vd + at(L::KW + 1, k) * va;
      const double ua = at(L::KK + 1, k) + at(L::KX + 4, k) * d0 + at(L::KX + 5, k) * d1 + at(L::KX + 6, k) * d2 + at(L::KX + 7, k) * d3 +
                        at(L::KW + 2, k) * vd + at(L::KW + 3, k) * va;
      const double n0 = d0 + at(L::JAC + 0, k) * d2 + at(L::JAC + 1, k) * d3 - at(L::CDEF + 0, k + 1);
      const double n1 = d1 + at(L::JAC + 2, k) * d2 + at(L::JAC + 3, k) * d3 - at(L::CDEF + 1, k + 1);
      const double n2 = d2 + at(L::JAC + 4, k) * d3 + at(L::JAC + 5, k) * ud - at(L::CDEF + 2, k + 1);
      const double n3 = d3 + T * ua - at(L::CDEF + 3, k + 1);
      if (lane == 0) {
        at(L::DU + 0, k) = ud; at(L::DU + 1, k) = ua;
        at(L::DX + 0, k + 1) = n0; at(L::DX + 1, k + 1) = n1; at(L::DX + 2, k + 1) = n2; at(L::DX + 3, k + 1) = n3;
      }
      d0 = n0; d1 = n1; d2 = n2; d3 = n3; vd = ud; va = ua;
    }
    if (lane == 0) { at(L::DU + 0, N) = 0.0; at(L::DU + 1, N) = 0.0; }
    __syncwarp();
  }

  // new dynamics multipliers: parallel r_k = Hxx_eff dx + Hux' du + gx_eff (written over the
  // consumed defects), then the serial adjoint recursion lam+_k = A_k' lam+_{k+1} - r_k
  __device__ __forceinline__ void adjoint() {
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      double d0 = at(L::DX + 0, k), d1 = at(L::DX + 1, k), d2 = at(L::DX + 2, k), d3 = at(L::DX + 3, k);
      double ud = k < N ? at(L::DU + 0, k) : 0.0;
      double h00 = at(L::HXX + 0, k), h01 = at(L::HXX + 1, k), h11 = at(L::HXX + 2, k), h22 = at(L::HXX + 3, k);
      double h23 = at(L::HXX + 4, k), h33 = at(L::HXX + 5, k), hdv = at(L::HUX, k);
      double r0 = at(L::GX + 0, k) + h00 * d0 + h01 * d1;
      double r1 = at(L::GX + 1, k) + h01 * d0 + h11 * d1;
      double r2 = at(L::GX + 2, k) + h22 * d2 + h23 * d3;
      double r3 = at(L::GX + 3, k) + h23 * d2 + h33 * d3 + hdv * ud;
      if (L::NH == 10) {
        double h02 = at(L::HXX + 6, k), h03 = at(L::HXX + 7, k), h12 = at(L::HXX + 8, k), h13 = at(L::HXX + 9, k);
        r0 += h02 * d2 + h03 * d3;
        r1 += h12 * d2 + h13 * d3;
        r2 += h02 * d0 + h12 * d1;
        r3 += h03 * d0 + h13 * d1;
      }
      at(L::LAMP + 0, k) = r0; at(L::LAMP + 1, k) = r1; at(L::LAMP + 2, k) = r2; at(L::LAMP + 3, k) = r3;
      own(L::LAMP, NX, k, TAG_NEWLAM);
    }
    __syncwarp();
    double l0 = -at(L::LAMP + 0, N), l1 = -at(L::LAMP + 1, N), l2 = -at(L::LAMP + 2, N), l3 = -at(L::LAMP + 3, N);
    if (lane == 0) { at(L::LAMP + 0, N) = l0; at(L::LAMP + 1, N) = l1; at(L::LAMP + 2, N) = l2; at(L::LAMP + 3, N) = l3; }
    #pragma unroll 1
    for (int k = N - 1; k >= 0; k--) {
      expect(L::LAMP, NX, k, TAG_NEWLAM);
      const double n0 = l0 - at(L::LAMP + 0, k);
      const double n1 = l1 - at(L::LAMP + 1, k);
      const double n2 = l2 + at(L::JAC + 0, k) * l0 + at(L::JAC + 2, k) * l1 - at(L::LAMP + 2, k);
      const double n3 = l3 + at(L::JAC + 1, k) * l0 + at(L::JAC + 3, k) * l1 + at(L::JAC + 4, k) * l2 - at(L::LAMP + 3, k);
      __syncwarp();
      if (lane == 0) { at(L::LAMP + 0, k) = n0; at(L::LAMP + 1, k) = n1; at(L::LAMP + 2, k) = n2; at(L::LAMP + 3, k) = n3; }
      l0 = n0; l1 = n1; l2 = n2; l3 = n3;
    }
    __syncwarp();
  }

  // slack steps, new row multipliers, step sizes (fraction to boundary), barrier slope.
  // The largest step a with v + a dv >= (1-tau) v-bound for every bounded quantity is
  // a = min(1, tau / max_i(-dv_i / gap_i)).
  __device__ __forceinline__ void slack_and_steps(double mu, double dw, double tau, double &a_pr, double &a_du, double &gd_out) {
    double rp = 0.0, rd = 0.0, gd = 0.0;  // largest primal / dual shrink ratios
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
        if (in_resto()) { double xr_ = at(L::XRS + i, k); gd += zeta * dr2(xr_) * (xk[i] - xr_) * dx[i]; }
        else if (k < N) gd += sigma * 2 * p.Q[i] * (xk[i] - xref(i, k)) * dx[i];
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = KinModel::bx(b);
        double rl = fast_rcp(xk[i] - p.x_lo[i]), rh = fast_rcp(p.x_hi[i] - xk[i]);
        gd += mu * (rh - rl) * dx[i];
        MPCB_LOWER(rl, dx[i], at(L::ZLX + b, k));
        MPCB_UPPER(rh, dx[i], at(L::ZUX + b, k));
      }
      double dsr[NR > 0 ? NR : 1], lrp[NR > 0 ? NR : 1], dso[MO > 0 ? MO : 1], lop[MO > 0 ? MO : 1];
      if (k < N) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
          double uk = at(L::U + i, k), du = at(L::DU + i, k);
          double um = k > 0 ? at(L::U + i, k - 1) : 0.0;
          double up = k + 1 <= N - 1 ? at(L::U + i, k + 1) : 0.0;
          double rl = fast_rcp(uk - p.u_lo[i]), rh = fast_rcp(p.u_hi[i] - uk);
          if (in_resto()) { double ur_ = at(L::URS + i, k); gd += (zeta * dr2(ur_) * (uk - ur_) + mu * (rh - rl)) * du; }
          else gd += (sigma * grad_u(k, i, uk, um, up) + mu * (rh - rl)) * du;
          MPCB_LOWER(rl, du, at(L::ZLU + i, k));
          MPCB_UPPER(rh, du, at(L::ZUU + i, k));
        }
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          double s = at(L::SR + r, k);
          double rl = fast_rcp(s - p.rate_lo[r]), rh = fast_rcp(p.rate_hi[r] - s);
          double vl = at(L::VLR + r, k), vu = at(L::VUR + r, k);
          double D = vl * rl + vu * rh + dw;
          double gs = mu * (rh - rl);
          double res = at(L::U + ci, k) - at(L::U + ci, k - 1) - s;
          double ds = at(L::DU + ci, k) - at(L::DU + ci, k - 1) + res;
          dsr[r] = ds;
          lrp[r] = D * ds + gs;
          if (RS) {
            const double Jdz = at(L::DU + ci, k) - at(L::DU + ci, k - 1);
            row_step(D, gs, res, mu, Jdz, ds, lrp[r]);
            dsr[r] = ds;
            if (resto) gd += at(L::LR + r, k) * (Jdz - ds);  // psi'(r) d(row - s)
          }
          gd += gs * ds;
          MPCB_LOWER(rl, ds, vl);
          MPCB_UPPER(rh, ds, vu);
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double ex = xk[0] - at(L::OCX + j, k), ey = xk[1] - at(L::OCY + j, k);
          double a = at(L::ISX + j, k), b = at(L::ISY + j, k);
          double d = ex * ex * a + ey * ey * b - 1.0;
          double s = at(L::SO + j, k), rg = fast_rcp(s - p.obs_lo), vl = at(L::VLO + j, k);
          double D = vl * rg + dw;
          double gs = -mu * rg + MPCB_KAPPA_D * mu;
          double ds;
          if (DCBF) {
            double fx = at(L::X + 0, k + 1) - at(L::OCX + j, k), fy = at(L::X + 1, k + 1) - at(L::OCY + j, k);
            double nx_ = 2 * fx * a, ny_ = 2 * fy * b;
            d = (fx * fx * a + fy * fy * b - 1.0) - p.cbf_g1 * d;
            const double Jdz = nx_ * at(L::DX + 0, k + 1) + ny_ * at(L::DX + 1, k + 1) - p.cbf_g1 * ((2 * ex * a) * dx[0] + (2 * ey * b) * dx[1]);
            ds = Jdz + (d - s);
            // The adjoint sweep ran on the condensed stage costs (row k folded into stage k); the
            // multipliers of the original problem differ by the row's pull on X_{k+1}.
            double l = D * ds + gs;
            if (RS) {
              row_step(D, gs, d - s, mu, Jdz, ds, l);
              if (resto) gd += at(L::LO + j, k) * (Jdz - ds);
            }
            at(L::LAMP + 0, k + 1) -= nx_ * l;
            at(L::LAMP + 1, k + 1) -= ny_ * l;
            lop[j] = l;
          } else {
            const double Jdz = (2 * ex * a) * dx[0] + (2 * ey * b) * dx[1];
            ds = Jdz + (d - s);
            lop[j] = D * ds + gs;
            if (RS) {
              row_step(D, gs, d - s, mu, Jdz, ds, lop[j]);
              if (resto) gd += at(L::LO + j, k) * (Jdz - ds);
            }
          }
          dso[j] = ds;
          gd += gs * ds;
          MPCB_LOWER(rg, ds, vl);
        }
      }
      // the gains of this stage are dead (forward sweep done): reuse their slots
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) { at(L::DSR + r, k) = dsr[r]; at(L::LRP + r, k) = lrp[r]; own(L::DSR + r, 1, k, TAG_STEPS); own(L::LRP + r, 1, k, TAG_STEPS); }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) { at(L::DSO + j, k) = dso[j]; at(L::LOP + j, k) = lop[j]; own(L::DSO + j, 1, k, TAG_STEPS); own(L::LOP + j, 1, k, TAG_STEPS); }
      }
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

  // accept the step: primal a, duals a_du (IPOPT: equality multipliers move with a)
  __device__ __forceinline__ void accept_step(double a, double ad, double mu) {
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      expect(L::LAMP, NX, k, p.debug_selftest ? TAG_GAINS : TAG_NEWLAM);
      expect(L::CDEFT, NX, k, TAG_TRIALDEF);
#pragma unroll
      for (int i = 0; i < NX; i++) {
        double l = at(L::LAM + i, k);
        at(L::LAM + i, k) = l + a * (at(L::LAMP + i, k) - l);
        at(L::CDEF + i, k) = at(L::CDEFT + i, k);  // LAMP aliases CDEF: consumed just above
      }
      own(L::CDEF, NX, k, TAG_DEFECT);
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        int i = KinModel::bx(b);
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
          expect(L::DSR + r, 1, k, TAG_STEPS);
          expect(L::LRP + r, 1, k, TAG_STEPS);
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
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          expect(L::DSO + j, 1, k, TAG_STEPS);
          expect(L::LOP + j, 1, k, TAG_STEPS);
          double s = at(L::SO + j, k), ds = at(L::DSO + j, k);
          double rg = fast_rcp(s - p.obs_lo), vl = at(L::VLO + j, k);
          double dvl = -vl + (mu - vl * ds) * rg;
          double sn = s + a * ds;
          at(L::SO + j, k) = sn;
          at(L::VLO + j, k) = clampz(vl + ad * dvl, mu, fast_rcp(sn - p.obs_lo));
          double l = at(L::LO + j, k);
          at(L::LO + j, k) = l + a * (at(L::LOP + j, k) - l);
        }
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- restoration passes (RS builds)
  // the row multipliers are not iterates in restoration: LR / LO = psi'(row residual)
  __device__ __forceinline__ void resto_multipliers(double mu) {
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          int ci = p.rate_ctrl[r];
          at(L::LR + r, k) = d_psi(at(L::U + ci, k) - at(L::U + ci, k - 1) - at(L::SR + r, k), mu, false).d1;
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          double dx = at(L::X + 0, k) - at(L::OCX + j, k), dy = at(L::X + 1, k) - at(L::OCY + j, k);
          double a = at(L::ISX + j, k), b = at(L::ISY + j, k);
          double d = dx * dx * a + dy * dy * b - 1.0;
          if (DCBF) {
            double ex = at(L::X + 0, k + 1) - at(L::OCX + j, k), ey = at(L::X + 1, k + 1) - at(L::OCY + j, k);
            d = (ex * ex * a + ey * ey * b - 1.0) - p.cbf_g1 * d;
          }
          at(L::LO + j, k) = d_psi(d - at(L::SO + j, k), mu, false).d1;
        }
      }
    }
    __syncwarp();
  }
  // entering: z_R = z, dynamics multipliers to zero, bound multipliers capped at rho
  // leaving (enter = false): every constraint multiplier to zero (IPOPT constr_mult_reset_threshold = 0), bound
  // multipliers back to one when any of them exceeds bound_mult_reset_threshold
  __device__ __forceinline__ void resto_switch(bool enter) {
    double zmax = 0.0;
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (enter) at(L::XRS + i, k) = at(L::X + i, k);
        at(L::LAM + i, k) = 0.0;
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        if (enter) at(L::URS + i, k) = k < N ? at(L::U + i, k) : 0.0;
        if (k < N) {
          if (enter) { at(L::ZLU + i, k) = fmin(MPCB_RESTO_RHO, at(L::ZLU + i, k)); at(L::ZUU + i, k) = fmin(MPCB_RESTO_RHO, at(L::ZUU + i, k)); }
          zmax = fmax(zmax, fmax(at(L::ZLU + i, k), at(L::ZUU + i, k)));
        }
      }
#pragma unroll
      for (int b = 0; b < NBX; b++) {
        if (enter) { at(L::ZLX + b, k) = fmin(MPCB_RESTO_RHO, at(L::ZLX + b, k)); at(L::ZUX + b, k) = fmin(MPCB_RESTO_RHO, at(L::ZUX + b, k)); }
        zmax = fmax(zmax, fmax(at(L::ZLX + b, k), at(L::ZUX + b, k)));
      }
      if (has_rate(k)) {
#pragma unroll
        for (int r = 0; r < NR; r++) {
          if (enter) { at(L::VLR + r, k) = fmin(MPCB_RESTO_RHO, at(L::VLR + r, k)); at(L::VUR + r, k) = fmin(MPCB_RESTO_RHO, at(L::VUR + r, k)); }
          else at(L::LR + r, k) = 0.0;
          zmax = fmax(zmax, fmax(at(L::VLR + r, k), at(L::VUR + r, k)));
        }
      }
      if (has_obs(k)) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          if (enter) at(L::VLO + j, k) = fmin(MPCB_RESTO_RHO, at(L::VLO + j, k));
          else at(L::LO + j, k) = 0.0;
          zmax = fmax(zmax, at(L::VLO + j, k));
        }
      }
    }
    zmax = warp_max(zmax);
    if (!enter && zmax > MPCB_BOUND_MULT_RESET) {
      #pragma unroll 1
      for (int k = lane; k <= N; k += 32) {
#pragma unroll
        for (int i = 0; i < 2; i++) { at(L::ZLU + i, k) = 1.0; at(L::ZUU + i, k) = 1.0; }
#pragma unroll
        for (int b = 0; b < NBX; b++) { at(L::ZLX + b, k) = 1.0; at(L::ZUX + b, k) = 1.0; }
        if (has_rate(k)) {
#pragma unroll
          for (int r = 0; r < NR; r++) { at(L::VLR + r, k) = 1.0; at(L::VUR + r, k) = 1.0; }
        }
        if (has_obs(k)) {
#pragma unroll
          for (int j = 0; j < MO; j++) at(L::VLO + j, k) = 1.0;
        }
      }
    }
    __syncwarp();
  }

  // ---------------------------------------------------------------- start point
  __device__ __forceinline__ bool init_iterate(int b) {
    const int nv = 2 * N + NX * (N + 1);
    const double *zi = p.z_init ? p.z_init + (size_t)b * nv : nullptr;
    // obstacle trajectory staged per step (centre and 1/semi-axis^2)
    const double *ob = p.obs ? p.obs + (size_t)b * MO * (p.obs_input ? 1 : (N + 1)) * 6 : nullptr;
    if (MO > 0 && p.obs_input) {
      // obstacle states given: the constant-velocity recursion of PKG/Obs_prediction.py:27-28,
      // accumulated step by step in the reference's operation order (device cos/sin may differ from
      // libm in the last ulp, nothing else does)
#pragma unroll
      for (int j = 0; j < MO; j++) {
        const double *o = ob + (size_t)j * 6;
        double x = o[0], y = o[1];
        // MPCB_OBS_STATIC: the row is the obstacle at every step (PKG/MPC_CBF_optimize_kin.py:236-243 reads
        // columns 0,1,4,5 only); adding +0.0 per step leaves x, y bit-identical to a repeated row
        const bool moving = p.obs_input == 1;
        const double dx = moving ? o[3] * cos(o[2]) * p.T : 0.0, dy = moving ? o[3] * sin(o[2]) * p.T : 0.0;
        const double sx = p.ego_hl + o[4] / 2 + p.safe_l, sy = p.ego_hw + o[5] / 2 + p.safe_w;
        const double isx = 1.0 / (sx * sx), isy = 1.0 / (sy * sy);
#pragma unroll 1
        for (int k = 0; k <= N; k++) {
          if ((k & 31) == lane) { at(L::OCX + j, k) = x; at(L::OCY + j, k) = y; at(L::ISX + j, k) = isx; at(L::ISY + j, k) = isy; }
          x = x + dx;
          y = y + dy;
        }
      }
    }
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
      if (!p.obs_input) {
#pragma unroll
        for (int j = 0; j < MO; j++) {
          const double *o = ob + ((size_t)j * (N + 1) + k) * 6;
          double sx = p.ego_hl + o[4] / 2 + p.safe_l;  // PKG/MPC_CBF_optimize_kin_pre.py:246-249
          double sy = p.ego_hw + o[5] / 2 + p.safe_w;
          at(L::OCX + j, k) = o[0];
          at(L::OCY + j, k) = o[1];
          at(L::ISX + j, k) = 1.0 / (sx * sx);
          at(L::ISY + j, k) = 1.0 / (sy * sy);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        at(L::U + i, k) = k < N ? push_in(zi ? zi[2 * k + i] : 0.0, p.u_lo[i], p.u_hi[i]) : 0.0;
        at(L::ZLU + i, k) = 1.0;
        at(L::ZUU + i, k) = 1.0;
        at(L::DU + i, k) = 0.0;
      }
      if (p.ref_mode) {
        const double *xr = p.xs + ((size_t)b * N + (k < N ? k : N - 1)) * NX;
#pragma unroll
        for (int i = 0; i < NX; i++) at(L::XR + i, k) = xr[i];
      }
#pragma unroll
      for (int i = 0; i < NX; i++) {
        if (p.init_mode == 0) at(L::X + i, k) = zi ? zi[2 * N + NX * k + i] : 0.0;
        at(L::LAM + i, k) = 0.0;
        at(L::DX + i, k) = 0.0;
      }
    }
    __syncwarp();
    if (p.init_mode == 1) {  // Euler roll-out of the guessed controls (PKG/MPC_CBF_optimize_kin.py:207)
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
          double u0_ = at(L::U + 0, k), u1_ = at(L::U + 1, k), s, c, t;
          d_trig(x[2], u0_, &s, &c, &t);
          double f0 = x[3] * c, f1 = x[3] * s, f2 = x[3] * t * (1.0 / p.Veh_l);  // same expression as eval_point
          x[0] = x[0] + p.T * f0;
          x[1] = x[1] + p.T * f1;
          x[2] = x[2] + p.T * f2;
          x[3] = x[3] + p.T * u1_;
        }
      }
      __syncwarp();
    }
    double gmax = 0;
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
#pragma unroll
      for (int b2 = 0; b2 < NBX; b2++) {
        int i = KinModel::bx(b2);
        at(L::X + i, k) = push_in(at(L::X + i, k), p.x_lo[i], p.x_hi[i]);
        at(L::ZLX + b2, k) = 1.0;
        at(L::ZUX + b2, k) = 1.0;
      }
    }
    if (DCBF) __syncwarp();  // the rows below read the pushed state of stage k+1 (another lane's)
    #pragma unroll 1
    for (int k = lane; k <= N; k += 32) {
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
          double dx = at(L::X + 0, k) - at(L::OCX + j, k), dy = at(L::X + 1, k) - at(L::OCY + j, k);
          double d = dx * dx * at(L::ISX + j, k) + dy * dy * at(L::ISY + j, k) - 1.0;
          if (DCBF) {
            double ex = at(L::X + 0, k + 1) - at(L::OCX + j, k), ey = at(L::X + 1, k + 1) - at(L::OCY + j, k);
            d = (ex * ex * at(L::ISX + j, k) + ey * ey * at(L::ISY + j, k) - 1.0) - p.cbf_g1 * d;
          }
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
        for (int i = 0; i < NX; i++) gmax = fmax(gmax, fabs(2 * p.Q[i] * (at(L::X + i, k) - xref(i, k))));
      }
    }
    gmax = warp_max(gmax);
    sigma = gmax > MPCB_OBJ_SCALE_MAX_GRAD ? MPCB_OBJ_SCALE_MAX_GRAD / gmax : 1.0;
    if (sigma < 1e-8) sigma = 1e-8;
    __syncwarp();
    return true;
  }

// All warps of a block start every interior-point iteration together (see kin_solve_kernel).
#ifndef MPCB_SYNC_EVERY
#define MPCB_SYNC_EVERY 1  // barrier every k-th iteration (power of two)
#endif
#define MPCB_ITER_SYNC() do { if (((++tick) & (MPCB_SYNC_EVERY - 1)) == 0) block_iteration_vote(0); } while (0)
#include "mpcb_run_loop.inc"
#undef MPCB_ITER_SYNC
};

// Persistent kernel: every resident warp pulls scenarios from a global queue until it is empty
// (iteration counts differ by 5x between scenarios: dynamic assignment keeps the SMs busy).
// W warps per block, one scenario per warp.  The warps of a block are kept in step at iteration
// granularity by one block barrier per interior-point iteration: the stage-parallel phases are
// ~4k instructions of straight-line code that stream through the instruction cache once per
// iteration, and unsynchronised warps thrash it (profiles/r01_*: `no_instruction` was the top
// stall and throughput did not scale with resident warps).  Warps in step fetch each line once.
// A warp whose queue is empty keeps answering the barrier until every warp of the block is done.
// W = 4 (one warp per SM sub-partition) measured best at N = 50; long horizons need more shared
// memory per scenario, the host picks the W in {4, 2, 1} that keeps most warps resident.
#ifndef MPCB_KIN_RESIDENT_WARPS
#define MPCB_KIN_RESIDENT_WARPS 12  // register budget: 65536 / (12 * 32) = 170 registers per thread
#endif
#ifndef MPCB_RS_INLINE_BOUNDS
#define MPCB_RS_INLINE_BOUNDS 0  // 1 with -DMPCB_RS_INLINE=1: the restoration-capable kernels ARE the main kernels and keep the register cap
#endif
template <int NR, int MO, int OBS_MODE, int W, bool GS, bool AS = false, bool RS = false>
__global__ void __launch_bounds__(32 * W, (AS || (RS && !MPCB_RS_INLINE_BOUNDS)) ? 1 : MPCB_KIN_RESIDENT_WARPS / W) kin_solve_kernel(const __grid_constant__ KParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  using L = KinLayout<NR, MO, OBS_MODE == 3, GS, AS, RS>;
  double *gs = AS ? nullptr : p.slab + ((size_t)blockIdx.x * W + warp) * L::slab_doubles();
  const int woff = warp * L::NFA * (p.N + 1);
  int tick = 0;
  // lets a dependent launch (the restoration pass) become resident as soon as every block of this grid has started
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (RS && p.resto_consumer) {
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(p.resto_sync + 1, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= p.B) break;
      int b = -1;
      if (lane == 0) {
        volatile int32_t *list = p.resto_list;
        volatile int *sync = p.resto_sync;
        for (;;) {
          b = list[t];
          if (b >= 0) break;
          if (sync[2] >= p.main_warps) {  // every producer is done and its entries are visible: look once more
            __threadfence();
            b = list[t];
            break;
          }
          __nanosleep(2000);
        }
      }
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b < 0) break;
      KinSolver<NR, MO, OBS_MODE, GS, AS, RS> s(p, gs, woff, tick, lane);
      s.run(b);
      __syncwarp();
    }
  } else {
    for (;;) {
      int b = 0;
      if (lane == 0) b = atomicAdd(p.counter, 1);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b >= p.B) break;
      if (p.order) b = p.order[b];  // caller-supplied processing order (longest expected first)
      KinSolver<NR, MO, OBS_MODE, GS, AS, RS> s(p, gs, woff, tick, lane);
      s.run(b);
      __syncwarp();
    }
    if (p.resto_sync) {  // producer side: this warp will add no more entries
      __threadfence();
      if (lane == 0) atomicAdd(p.resto_sync + 2, 1);
    }
  }
  while (!block_iteration_vote(1)) {
  }
}

}  // namespace mpcb
