/* mpcb200 — C ABI of the B200-native batched nonlinear-MPC solve path.
 *
 * Drop-in boundary for the one hot path of ZhuorenLi/MPC_motion_planning: the per-control-
 * step "build the multiple-shooting NLP and solve it" pair
 *
 *     solver = mpc_solver.optimize_problem(ego_state, ref_state, obstacle)   PKG/main_cbf_kin_c_sim.py:99
 *     res    = solver(x0=init_control, p=c_p, lbg=, lbx=, ubg=, ubx=)        PKG/main_cbf_kin_c_sim.py:100
 *
 * (PKG = CasaDi_MPC_Optimize_Multishoot; same pair at main_cbf_kin_c_sim_pre.py:99-100,
 * main_cbf_dyn_c_sim.py:89-90, main_kin_c_sim.py:83-84).  The reference has no FFI layer —
 * the seam is CasADi's functional interface — so these entry points are what a ctypes
 * binding placed behind `MPC_optimize.optimize_problem` calls.
 *
 * All entry points are plain C: pointers + sizes, no C++/torch types, no exceptions.
 * Buffers are caller-owned.  `*_batch` takes DEVICE pointers and is asynchronous on the
 * given stream; `*_batch_host` takes HOST pointers and is synchronous (copies inside).
 */
#ifndef MPCB200_H
#define MPCB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCB_VERSION 200 /* 0.2.0: mpcb_set_order takes the length, mpcb_reserve, mpcb_set_dual_outputs, mpcb_n_g;
                            0.1.3: mpcb_submit_batch_host / mpcb_wait; 0.1.2: cfg gained dyn_rows; 0.1.1: ref_mode, cbf_gamma */
#define MPCB_NMAX 128    /* maximum horizon N */
#define MPCB_MMAX 4      /* maximum obstacles per scenario in this build (the mains carry a commented 3-obstacle list) */

/* vehicle model: replaces the CasADi `rhs` of PKG/MPC_CBF_optimize_kin.py:153-156 (KIN) and
 * PKG/MPC_CBF_optimize_dyn.py:156-170 (DYN) */
enum { MPCB_MODEL_KIN = 0, MPCB_MODEL_DYN = 1 };
/* obstacle rows: NONE = MPC_optimize_kin (no CBF); ELLIPSE = h(X_i) >= 0, i=0..N-1
 * (PKG/MPC_CBF_optimize_kin.py:236-247, _kin_pre.py:236-253); SQRT = sqrt(ellipse-1) >= 1,
 * i=0..N (PKG/MPC_CBF_optimize_dyn.py:238-243); DCBF = the discrete-time control-barrier row the
 * reference carries commented out, `gamma*h_func + h_dot` with h_dot = h(X_{i+1}) - h(X_i), both at
 * the step-i obstacle (PKG/MPC_CBF_optimize_kin.py:244-248, _kin_pre.py:250-254), i.e.
 * h(X_{i+1}; obs_i) - (1-gamma) h(X_i; obs_i) >= 0, i=0..N-1 (kinematic model) */
enum { MPCB_OBS_NONE = 0, MPCB_OBS_ELLIPSE = 1, MPCB_OBS_SQRT = 2, MPCB_OBS_DCBF = 3 };
/* cost target: TERMINAL = xs [B][nx], every stage tracks xs (the reference's aa = 0,
 * PKG/MPC_CBF_optimize_kin.py:194-199); TRAJECTORY = xs [B][N][nx], stage i tracks row i =
 * aa*ref_state[i+1] + (1-aa)*xs, blended by the host (kinematic model) */
enum { MPCB_REF_TERMINAL = 0, MPCB_REF_TRAJECTORY = 1 };
/* how `z_init` (the `x0=` argument of the CasADi call, PKG/main_cbf_kin_c_sim.py:92,100) is used:
 * AS_GIVEN = controls and states as passed; ROLLOUT = controls as passed, states re-integrated
 * from the parameter x0 with the Euler step of PKG/MPC_CBF_optimize_kin.py:207 */
enum { MPCB_INIT_AS_GIVEN = 0, MPCB_INIT_ROLLOUT = 1 };
/* layout of the `obs` argument: TRAJECTORY = [B][M][N+1][6], the output of obs_prediction
 * (PKG/Obs_prediction.py:3-40); INITIAL = [B][M][6], the obstacle states themselves - the library
 * then runs the same constant-velocity recursion (x += v cos(theta) dt, step by step, :27-28) on the
 * device while staging the trajectory, so nothing per-step crosses PCIe or HBM */
/* STATIC = [B][M][6], obstacle rows that hold at every step - what the static-obstacle module's
 * `optimize_problem(ego_state, ref_state, obstacle)` takes (PKG/MPC_CBF_optimize_kin.py:136,236-243;
 * PKG/main_cbf_kin_c_sim.py:55,99); identical results to TRAJECTORY with the row repeated */
enum { MPCB_OBS_TRAJECTORY = 0, MPCB_OBS_INITIAL = 1, MPCB_OBS_STATIC = 2 };

/* dyn model: how the rows of g are paired with bounds.  ALIGNED = as the code's comments intend (defects = 0,
 * rate rows within the rate bounds).  AS_SHIPPED = exactly as the lists of PKG/MPC_CBF_optimize_dyn.py:112-133
 * line up with g (:215,227-243): every rate row is an equality U_i = U_{i-1} (one control pair for the whole
 * horizon) and the x / y defects into stages 2..N are range rows with the rate bounds - what IPOPT is actually
 * given by the reference's dyn main. */
enum { MPCB_DYN_ROWS_ALIGNED = 0, MPCB_DYN_ROWS_AS_SHIPPED = 1 };
enum { MPCB_ENGINE_AUTO = 0, MPCB_ENGINE_WARP = 1, MPCB_ENGINE_LANE = 2 };
/* shooting defects: EULER = X_{k+1} - X_k - T f(X_k,U_k), what the reference builds (PKG/MPC_CBF_optimize_kin.py:207) and the
 * only mode with reference parity; RK4 = X_{k+1} - Phi(X_k,U_k) with the classical Runge-Kutta step, forward-mode Jacobians
 * through its four stages and the exact second-order adjoint for the Lagrangian Hessian - the option BASELINE.json's north_star
 * names (the reference's tree has Runge-Kutta only in the tutorial Reference/MPC/sim_test.py:35-37).  RK4 is served by the
 * lane engine: kinematic model, plain rows, at most two obstacles */
enum { MPCB_INTEGRATOR_EULER = 0, MPCB_INTEGRATOR_RK4 = 1 };

/* per-scenario outcome, mapped to IPOPT return_status strings by the Python shim */
enum {
  MPCB_CONVERGED = 0,  /* Solve_Succeeded */
  MPCB_ACCEPTABLE = 1, /* Solved_To_Acceptable_Level: the line search failed at a point whose scaled KKT error is
                          already <= 1e-6 (IPOPT's default acceptable_tol; the reference's 1e-8 equals tol) */
  MPCB_MAXITER = 2,    /* Maximum_Iterations_Exceeded */
  MPCB_INFEASIBLE = 3, /* Infeasible_Problem_Detected: the restoration phase converged to a stationary point of the
                          infeasibility, or was entered more than cfg.resto_max_calls times without the regular phase
                          converging from where it left off (a locally infeasible start); with cfg.restoration = 0, and
                          for the dyn family: the line search failed and no restoration was attempted */
  MPCB_NAN = 4,        /* Invalid_Number_Detected */
  MPCB_RESTO_FAILED = 5 /* Restoration_Failed: the restoration phase's own line search / inertia correction failed */
};

/* return codes of the entry points */
enum {
  MPCB_OK = 0,
  MPCB_E_ARG = -1,      /* bad argument / unsupported configuration */
  MPCB_E_CUDA = -2,     /* CUDA runtime error (see mpcb_last_cuda_error) */
  MPCB_E_NOMEM = -3,
  MPCB_E_NODEVICE = -4  /* no CUDA device: there is no CPU fallback */
};

/* Problem definition.  Everything the reference hard-codes in `optimize_problem` /
 * `initialize_constraints` / mpc_parameters.yaml is a field here; the Python host fills it
 * with the reference's values. */
typedef struct mpcb_cfg {
  int32_t model;        /* MPCB_MODEL_* */
  int32_t N;            /* N_p (PKG/MPC_CBF_optimize_kin.py:32-33) */
  int32_t M;            /* obstacles per scenario */
  int32_t obs_mode;     /* MPCB_OBS_* */
  int32_t du0_cost;     /* include (U_0-Ulast)'DR(U_0-Ulast), Ulast=0 (kin-CBF: yes, :203-204; dyn/no-CBF: no) */
  int32_t n_rate;       /* number of rate-limited controls, rows for i=1..N-1 (:211-216; dyn :229-231) */
  int32_t rate_ctrl[2]; /* which control each rate row constrains (0 = df, 1 = ax) */
  int32_t init_mode;    /* MPCB_INIT_* */
  int32_t max_iter;     /* ipopt.max_iter (:252) */
  double T;             /* T_S */
  double Q[6], R[2], DR[2];           /* weights (:168-184) */
  double rate_lo[2], rate_hi[2];      /* df_dot_min*T_S .. (:119-121); dyn also jerk*T_S */
  double u_lo[2], u_hi[2];            /* lbx/ubx of the controls (:90-95) */
  double x_lo[6], x_hi[6];            /* lbx/ubx of the states, +-inf when free (:97-105) */
  double obs_lo;                      /* lower bound of the obstacle rows (0 kin, 1 dyn) */
  double ego_hl, ego_hw, safe_l, safe_w; /* Veh_L/2, Veh_W/2, safe_disl, safe_disw (:220-225) */
  double dyn_sx, dyn_sy;              /* safe_X, safe_Y of the dyn row (_dyn.py:240-241) */
  double Veh_l, Veh_lf, Veh_lr, Veh_m, Veh_Iz, aopt_f, aopt_r, Fymax_f, Fymax_r;
  double tol;           /* ipopt.tol (default 1e-8) */
  double mu_init;       /* initial barrier parameter (IPOPT default 0.1; this library's default 30) */
  double bound_relax;   /* ipopt.bound_relax_factor (1e-8) */
  int32_t obs_input;    /* MPCB_OBS_TRAJECTORY (default), MPCB_OBS_INITIAL or MPCB_OBS_STATIC */
  int32_t ref_mode;     /* MPCB_REF_TERMINAL (default) or MPCB_REF_TRAJECTORY */
  double cbf_gamma;     /* gamma in (0,1] of the MPCB_OBS_DCBF rows (the reference's `gamma = 1.00`, :235) */
  int32_t dyn_rows;     /* MPCB_DYN_ROWS_ALIGNED (default) or MPCB_DYN_ROWS_AS_SHIPPED */
  int32_t restoration;  /* 1: a failed line search / inertia correction enters the restoration phase (what IPOPT does behind
                           PKG/MPC_CBF_optimize_kin.py:252-254) instead of ending the solve; kinematic families with rows.
                           The scenarios concerned are solved again, from their start point, by a restoration-capable
                           sibling kernel launched right after the main one on the same stream */
  int32_t resto_max_calls; /* > 0: the (n+1)-th entry into the restoration phase ends the solve with MPCB_INFEASIBLE; 0 = no cap
                           (IPOPT: the phases alternate until max_iter).  Library default 1, drop-in classes 0 */
  int32_t engine;       /* MPCB_ENGINE_AUTO (default), _WARP: one scenario per warp (every family), _LANE: one scenario per
                           lane (kinematic model, plain rows, one target per scenario).  AUTO uses the
                           lane engine where it measures faster: the families without obstacle rows from 20,480 scenarios up */
  int32_t integrator;   /* MPCB_INTEGRATOR_EULER (default) or MPCB_INTEGRATOR_RK4 */
  int32_t reserved;
} mpcb_cfg;

typedef struct mpcb_handle mpcb_handle;

int mpcb_version(void);
const char *mpcb_strerror(int code);
const char *mpcb_last_cuda_error(void);

/* state dimension implied by cfg->model (4 or 6); nv = 2N + nx(N+1) */
int mpcb_nx(const mpcb_cfg *cfg);
int mpcb_nv(const mpcb_cfg *cfg);

/* validates cfg, selects the device-resident kernel variant; no device allocation beyond
 * a few bytes.  One handle per GPU (current device at creation); not re-entrant. */
int mpcb_create(const mpcb_cfg *cfg, mpcb_handle **out);
void mpcb_destroy(mpcb_handle *h);

/* upper bound of the device scratch mpcb_create allocates for this configuration (independent of
 * B: a slab per RESIDENT warp of the persistent kernel) */
int mpcb_workspace_bytes(const mpcb_cfg *cfg, int B, size_t *bytes);

/* One solve at a time per handle (it owns one work queue and one iterate slab): a call on a different stream than the
 * previous call on the same handle first waits, on the device, for that one to finish.
 *
 * Replaces optimize_problem + solver(...) for B independent scenarios.  DEVICE pointers,
 * row-major, float64:
 *   x0 [B][nx], xs [B][nx]        = the parameter vector p=[x0;xs] (PKG/main_cbf_kin_c_sim.py:89);
 *                                   xs [B][N][nx] with cfg.ref_mode = MPCB_REF_TRAJECTORY
 *   obs [B][M][N+1][6]            = obs_prediction rows [x,y,theta,v,l,w] (PKG/Obs_prediction.py:3-40);
 *                                   static obstacles repeat the row; dyn uses columns 0,1 only.
 *                                   With cfg.obs_input = MPCB_OBS_INITIAL or MPCB_OBS_STATIC: [B][M][6] (see above)
 *   z_init [B][nv] or NULL        = `x0=` warm start [vec(U);vec(X)] (NULL = zeros, :47-50)
 * outputs:
 *   u0 [B][2]                     = first control, res['x'][0:2]
 *   cost [B]                      = res['f']
 *   status [B], iters [B]         = per-scenario outcome (MPCB_CONVERGED ...) and iteration count
 *   z_out [B][nv] or NULL         = res['x']
 *   lam_out [B][nx(N+1)] or NULL  = multipliers of [X_0-x0 ; defects] (head of res['lam_g'])
 * Asynchronous on `stream` (a cudaStream_t passed as void*). */
int mpcb_solve_batch(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                     const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                     double *z_out, double *lam_out, void *stream);

/* Same with HOST pointers: allocates/reuses device buffers inside the handle, copies in,
 * solves, copies out, synchronises.  This is the call a ctypes/cgo/JNI stub binds. */
int mpcb_solve_batch_host(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                          const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                          double *z_out, double *lam_out);

/* The same call split in two, for callers with several independent batches per control period (the reference
 * has no counterpart: its solver(...) call, PKG/main_cbf_kin_c_sim.py:100, blocks):
 * mpcb_submit_batch_host enqueues copy-in, solve and copy-out on the handle's own stream and returns;
 * mpcb_wait blocks until everything submitted on the handle has finished.  The host buffers must stay valid and
 * the outputs unread until mpcb_wait returns; use page-locked host memory for the copies to overlap.  A handle
 * still holds one batch at a time -- a second submit on the same handle queues behind the first and reuses its
 * device buffers -- so a pipeline uses two (or more) handles and alternates: while one batch's long scenarios
 * finish on a few warps, the other handle's batch fills the rest of the GPU. */
int mpcb_submit_batch_host(mpcb_handle *h, int B, const double *x0, const double *xs, const double *obs,
                           const double *z_init, double *u0, double *cost, int32_t *status, int32_t *iters,
                           double *z_out, double *lam_out);
int mpcb_wait(mpcb_handle *h);

/* Sizes the device staging buffers of the *_host entry points for batches of up to B scenarios, so that no later
 * mpcb_solve_batch_host / mpcb_submit_batch_host with a batch <= B allocates ("no allocation
 * in solve_batch"; mpcb_solve_batch with device pointers never allocates).  Without it the first call that sees a
 * larger batch grows them. */
int mpcb_reserve(mpcb_handle *h, int B);

/* Number of rows of g (= length of CasADi's res['g'] / res['lam_g']) for this configuration. */
int mpcb_n_g(const mpcb_cfg *cfg);

/* The rest of the CasADi result dict (`{'x','f','g','lam_x','lam_g','lam_p'}`; the mains read
 * only 'x', PKG/main_cbf_kin_c_sim.py:100-102).  When set, every later solve on the handle also writes
 *   lam_g [B][mpcb_n_g]  multipliers of ALL rows of g in the reference's row order - kin: X_0 - x0, defects,
 *                        steering-rate rows, obstacle rows (PKG/MPC_CBF_optimize_kin.py:191,207-216,236-247); dyn: X_0 - x0,
 *                        then per stage the defect and (i > 0) the two rate rows, then the obstacle rows
 *                        (PKG/MPC_CBF_optimize_dyn.py:215,227-231,242-243)
 *   lam_x [B][nv]        bound multipliers z_U - z_L in the order of 'x' (CasADi's sign: > 0 at an active upper bound)
 * in the units of the unscaled objective.  on_host = 0: DEVICE buffers, written by mpcb_solve_batch; on_host = 1: HOST
 * buffers, filled by the *_host entry points.  Either pointer may be NULL; (NULL, NULL) switches the outputs off. */
int mpcb_set_dual_outputs(mpcb_handle *h, double *lam_g, double *lam_x, int on_host);

/* Closed-loop helper (row N1 of SURVEY.md section 8f; PKG/main_cbf_kin_c_sim.py:16-26):
 * plant Euler step x0 <- x0 + T f(x0, U_0) and warm-start shift of z (drop first row, repeat last),
 * in place on DEVICE buffers x0 [B][nx], z [B][nv]. */
int mpcb_shift_batch(mpcb_handle *h, int B, double *x0, double *z, void *stream);

/* Batched RefPathGenerator (row N2 of SURVEY.md section 8f; PKG/RefPathGenerator.py:9-59) on DEVICE
 * buffers, kinematic model.  The straight global path of define_ref_path is implicit: points
 * path_x0[b] + i*(+-1) towards xs.x with y/phi/v of xs (:9-24).  Per scenario: nearest-point walk
 * from last_idx (updated in place) and preview resampling to N+1 rows (find_ref_traj, :27-59), with
 * numpy's exact index arithmetic.  Outputs (either may be NULL):
 *   ref [B][N+1][4]           = the `ref_traj` the mains pass as ref_state
 *   stage_targets [B][N][4]   = aa*ref[i+1] + (1-aa)*xs, the `xs` argument of mpcb_solve_batch under
 *                               MPCB_REF_TRAJECTORY (PKG/MPC_CBF_optimize_kin.py:196)
 * T_horizon is the YAML `horizon` (int(T_horizon/T_S) must equal cfg.N). */
int mpcb_ref_traj_batch(mpcb_handle *h, int B, double T_horizon, const double *x0, const double *xs, const double *path_x0,
                        int32_t *last_idx, double aa, double *ref, double *stage_targets, void *stream);

/* Batched obs_prediction (PKG/Obs_prediction.py:3-40) and the mains' obstacle update on DEVICE buffers:
 *   obs_state [n_obstacles][6]      rows [x, y, theta, v, l, w] (a [B][M][6] array has n_obstacles = B*M)
 *   traj      [n_obstacles][N+1][6] or NULL: the constant-velocity roll-out, x += v cos(theta) dt accumulated step by step
 *                                   in the reference's operation order (:27-28) - the `obs` argument of mpcb_solve_batch
 *                                   under MPCB_OBS_TRAJECTORY
 *   advance != 0                    afterwards moves every obstacle one step, in place (PKG/main_cbf_kin_c_sim_pre.py:106:
 *                                   `obs = [obs_trajectories_dyn[0][1]]`)
 * With cfg.obs_input = MPCB_OBS_INITIAL the solve kernels run the same recursion themselves; this entry is for callers
 * that want the trajectories, and for closed loops that only need the update. */
int mpcb_obs_prediction_batch(mpcb_handle *h, int n_obstacles, double *obs_state, double *traj, int advance, void *stream);

/* Scheduling hint.  The resident warps pull scenarios from a queue; iteration counts differ by 5x
 * between scenarios, so at small batches (a few scenarios per resident warp) the makespan is set by
 * long scenarios that start late.  `order` (DEVICE, [n] permutation of 0..n-1, read by every later
 * solve until reset with NULL; a solve whose batch size differs from n fails with MPCB_E_ARG) makes queue position q process scenario order[q]: pass the scenarios
 * sorted by expected work, longest first - in a closed loop the previous step's `iters` is a good
 * predictor.  Results do not depend on the order. */
int mpcb_set_order(mpcb_handle *h, const int32_t *order, int n);

/* Diagnostics (the reference only has IPOPT's print_level log, PKG/MPC_CBF_optimize_kin.py:252):
 * when set, every later solve writes one row per interior-point iteration and scenario into the
 * DEVICE buffer trace[B][rows][8] = (mu, theta, kkt_error, dual_inf, primal_inf, compl_inf,
 * accepted step size of the previous iteration, last inertia regularisation).  rows = 0 disables. */
int mpcb_set_trace_buffer(mpcb_handle *h, double *trace, int rows);

/* Sanitizer substitute.  The warp-per-scenario kernels reuse shared-memory slots between the phases of an iteration; the
 * build with -DMPCB_DEBUG_SLOTS (python -m mpc_motion_planning_b200.build --debug-slots -> libmpcb200_debug.so) tags every
 * store into those slots with the kind of data and checks the tag at every load.  *count = violations seen so far by the
 * solves on this handle, or -1 when the library is not that build. */
int mpcb_debug_slot_errors(mpcb_handle *h, int *count);

/* kernel launch statistics of the last solve on this handle */
typedef struct mpcb_launch_info {
  int32_t grid, block, smem_bytes, regs_per_thread, blocks_per_sm, num_sms;
  int64_t launches; /* kernels launched by this handle so far */
} mpcb_launch_info;
int mpcb_get_launch_info(mpcb_handle *h, mpcb_launch_info *out);

/* Measured FP64 FMA-pipe throughput of the current device (TFLOP/s, FMA = 2 flops): the
 * roofline denominator bench.py reports the solve kernel against. */
int mpcb_fp64_peak_tflops(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MPCB200_H */
