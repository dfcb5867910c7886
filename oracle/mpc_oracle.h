/* CPU oracle for the MPC solve path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * PARITY UNPINNED (solver): the reference's solve is CasADi nlpsol('ipopt')
 * (PKG/MPC_CBF_optimize_kin.py:251-254, called at PKG/main_cbf_kin_c_sim.py:100);
 * neither CasADi nor IPOPT exists in this image and the reference holds no golden
 * vectors, so this oracle restates the NLP exactly (oracle/nlp.py has the per-line
 * citations and IS pinned against the reference's own expressions, see
 * tests/test_reference_vectors.py) and IPOPT's published algorithm approximately
 * (oracle/ipm_dense.py is the dense specification; this file solves the same Newton
 * systems with a scalar stage-wise Riccati recursion so that it is fast enough to be
 * the CPU baseline).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 */
#ifndef MPC_ORACLE_H
#define MPC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NMAX 128 /* max horizon */
#define ORC_MMAX 4   /* max obstacles */

enum { ORC_MODEL_KIN = 0, ORC_MODEL_DYN = 1 };
enum { ORC_OBS_NONE = 0, ORC_OBS_ELLIPSE = 1, ORC_OBS_SQRT = 2, ORC_OBS_DCBF = 3 };
enum { ORC_REF_TERMINAL = 0, ORC_REF_TRAJECTORY = 1 };
enum { ORC_INIT_AS_GIVEN = 0, ORC_INIT_ROLLOUT = 1 };
enum { ORC_CONVERGED = 0, ORC_ACCEPTABLE = 1, ORC_MAXITER = 2, ORC_INFEASIBLE = 3, ORC_NAN = 4, ORC_RESTO_FAILED = 5 };

typedef struct {
  int32_t model;      /* ORC_MODEL_* */
  int32_t N;          /* horizon steps */
  int32_t M;          /* obstacles per scenario */
  int32_t obs_mode;   /* ORC_OBS_* */
  int32_t du0_cost;   /* i=0 control-rate cost vs Ulast=0 (PKG/MPC_CBF_optimize_kin.py:203-204) */
  int32_t n_rate;     /* number of rate-limited controls (rows for i=1..N-1) */
  int32_t rate_ctrl[2];
  int32_t init_mode;  /* ORC_INIT_* */
  int32_t max_iter;
  double T;           /* T_S */
  double Q[6], R[2], DR[2];
  double rate_lo[2], rate_hi[2];
  double u_lo[2], u_hi[2];
  double x_lo[6], x_hi[6]; /* +-inf for free */
  double obs_lo;      /* lower bound of the obstacle rows (0 kin, 1 dyn) */
  double ego_hl, ego_hw, safe_l, safe_w; /* sX = ego_hl + l/2 + safe_l (kin rows) */
  double dyn_sx, dyn_sy; /* fixed semi-axes of the dyn row (4, 1) */
  /* vehicle */
  double Veh_l, Veh_lf, Veh_lr, Veh_m, Veh_Iz, aopt_f, aopt_r, Fymax_f, Fymax_r;
  /* interior-point options */
  double tol, mu_init, bound_relax;
  /* options the reference ships switched off */
  double cbf_gamma;   /* ORC_OBS_DCBF: rows h(X_{k+1};obs_k) - (1-gamma) h(X_k;obs_k) >= 0, k=0..N-1
                         (the commented row of PKG/MPC_CBF_optimize_kin.py:244-248) */
  int32_t ref_mode;   /* ORC_REF_TRAJECTORY: xs is (N, nx) per-stage cost targets
                         (ref_X of PKG/MPC_CBF_optimize_kin.py:194-199 with aa != 0) */
  int32_t rows_as_shipped; /* dyn only: pair the rows of g with the bound lists exactly as
                         PKG/MPC_CBF_optimize_dyn.py:112-133 ships them: every rate row
                         is an equality U_i = U_{i-1}, the x/y defects of stages 2..N are range rows with the
                         rate bounds.  0 = aligned (what the code's comments intend). */
  int32_t restoration; /* 1: a failed line search enters the restoration phase (ipm_dense.py) instead of ending with
                         ORC_INFEASIBLE; not available with rows_as_shipped */
  int32_t resto_max_calls; /* > 0: entering the restoration phase for the (resto_max_calls+1)-th time ends the solve with
                         ORC_INFEASIBLE (restoration keeps reducing the infeasibility by the required 10 % without ever reaching a
                         point from which the regular phase converges: a locally infeasible start).  IPOPT has no such cap and
                         would use up max_iter; 0 = IPOPT's behaviour. */
  int32_t integrator;  /* 0: explicit Euler defects X_{k+1} - X_k - T f(X_k,U_k) (the reference, PKG/MPC_CBF_optimize_kin.py:207);
                         1: classical Runge-Kutta (kinematic model) - the option BASELINE.json's north_star names.  Implemented
                         as the increment function (Phi_RK4(x,u) - x)/T in place of f, see oracle/nlp.py Rk4KinModel */
  int32_t reserved;
} orc_cfg;

typedef struct {
  double f;       /* unscaled objective */
  double err;     /* final scaled KKT error */
  double mu;
  double obj_scale;
  int32_t status, iters, n_reg, n_backtrack;
  int32_t n_resto, n_resto_iter; /* restoration phases entered, iterations spent in them (part of iters) */
} orc_info;

/* one scenario.  xs: nx, or (N, nx) with ORC_REF_TRAJECTORY; obs: (M, N+1, 6) rows [x,y,theta,v,l,w];
 * z_init: nv or NULL (zeros);
 * z_out: nv = 2N + nx(N+1) in the reference order [vec(U); vec(X)]; lam_g_out: nullable,
 * multipliers of [X0-x0 ; defects] (nx(N+1)), unscaled. */
int orc_solve(const orc_cfg *cfg, const double *x0, const double *xs, const double *obs,
              const double *z_init, double *z_out, double *lam_eq_out, orc_info *info);

/* batch with OpenMP threads; arrays are row-major [B][...]; z_init/z_out nullable */
int orc_solve_batch(const orc_cfg *cfg, int B, const double *x0, const double *xs, const double *obs,
                    const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                    double *z_out, int nthreads);

/* one Newton step at a given primal point with unit bound multipliers and zero equality
 * multipliers (used by the Riccati-vs-dense test): returns dz (nv) */
int orc_newton_step(const orc_cfg *cfg, const double *x0, const double *xs, const double *obs,
                    const double *z, double mu, double dw, double obj_scale, double *dz, double *lam_plus);

int orc_nx(const orc_cfg *cfg);

#ifdef __cplusplus
}
#endif
#endif
