/* Plain-C caller of libmpcb200: the reference's kin-CBF main, first MPC step
 * (PKG/main_cbf_kin_c_sim.py:45-55,99-104), through the host-pointer entry.
 *
 *   gcc -std=c11 -Iinclude examples/batch_solve.c -Lmpc_motion_planning_b200 -lmpcb200 -lm -o batch_solve
 *   LD_LIBRARY_PATH=mpc_motion_planning_b200 ./batch_solve
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpcb200.h"

int main(void) {
  const double PI = 3.14159265358979323846;
  mpcb_cfg c;
  memset(&c, 0, sizeof c);
  c.model = MPCB_MODEL_KIN;
  c.N = 50;
  c.M = 1;
  c.obs_mode = MPCB_OBS_ELLIPSE;
  c.obs_input = MPCB_OBS_STATIC;          /* obstacle rows (M,6) as optimize_problem takes them */
  c.du0_cost = 1;                         /* PKG/MPC_CBF_optimize_kin.py:203-204 */
  c.n_rate = 1;                           /* steering-rate rows, :211-216 */
  c.rate_ctrl[0] = 0;
  c.init_mode = MPCB_INIT_ROLLOUT;
  c.max_iter = 100;                       /* :252 */
  c.T = 0.1;
  const double Q[4] = {1e1, 1e5, 3e5, 1e4}, R[2] = {1e4, 1e4}, DR[2] = {1e5, 1e2}; /* :168-184 */
  memcpy(c.Q, Q, sizeof Q);
  memcpy(c.R, R, sizeof R);
  memcpy(c.DR, DR, sizeof DR);
  c.rate_lo[0] = -5.0 * PI / 180 * c.T;   /* df_dot_min * T_S, :119-121 */
  c.rate_hi[0] = 5.0 * PI / 180 * c.T;
  c.u_lo[0] = -35.0 * PI / 180; c.u_hi[0] = 35.0 * PI / 180;   /* :90-95 */
  c.u_lo[1] = -3.0;             c.u_hi[1] = 3.0;
  for (int i = 0; i < 6; i++) { c.x_lo[i] = -INFINITY; c.x_hi[i] = INFINITY; }
  c.x_lo[1] = -1.0; c.x_hi[1] = 5.0;      /* Y, :97-105 */
  c.x_lo[3] = 0.0;  c.x_hi[3] = 40.0;     /* vx */
  c.obs_lo = 0.0;
  c.ego_hl = 4.8 / 2; c.ego_hw = 1.8 / 2; c.safe_l = 1.0; c.safe_w = 0.5; /* :220-225 */
  c.Veh_l = 2.6;
  c.tol = 1e-8; c.mu_init = 30.0; c.bound_relax = 1e-8;

  mpcb_handle *h = NULL;
  int rc = mpcb_create(&c, &h);
  if (rc != MPCB_OK) { fprintf(stderr, "mpcb_create: %s (%s)\n", mpcb_strerror(rc), mpcb_last_cuda_error()); return 2; }

  enum { B = 3 };
  double x0[B][4] = {{0, 3, 0, 15}, {5, 1.0, 0.01, 20}, {0, 3, 0, 15}};
  double xs[B][4] = {{400, 3.5, 0, 30}, {400, 3.5, 0, 30}, {400, 3.5, 0, 30}};
  double obs[B][1][6] = {{{50, 3.5, 0, 8, 4.8, 1.8}}, {{70, 0.5, 0, 0, 4.8, 1.8}}, {{50, 3.5, 0, 8, 4.8, 1.8}}};
  double u0[B][2], cost[B];
  int32_t status[B], iters[B];
  rc = mpcb_solve_batch_host(h, B, &x0[0][0], &xs[0][0], &obs[0][0][0], NULL, &u0[0][0], cost, status, iters, NULL, NULL);
  if (rc != MPCB_OK) { fprintf(stderr, "mpcb_solve_batch_host: %s (%s)\n", mpcb_strerror(rc), mpcb_last_cuda_error()); return 3; }
  for (int b = 0; b < B; b++)
    printf("scenario %d: status %d, %d iterations, cost %.10e, u0 = (%.8f, %.8f)\n", b, status[b], iters[b], cost[b], u0[b][0], u0[b][1]);
  /* the same batch twice more, in flight on two handles (mpcb_submit_batch_host / mpcb_wait) */
  mpcb_handle *h2 = NULL;
  rc = mpcb_create(&c, &h2);
  if (rc != MPCB_OK) { fprintf(stderr, "mpcb_create (second handle): %s\n", mpcb_strerror(rc)); return 4; }
  double u0a[B][2], costa[B], u0b[B][2], costb[B];
  int32_t sta[B], ita[B], stb[B], itb[B];
  rc = mpcb_submit_batch_host(h, B, &x0[0][0], &xs[0][0], &obs[0][0][0], NULL, &u0a[0][0], costa, sta, ita, NULL, NULL);
  if (rc == MPCB_OK) rc = mpcb_submit_batch_host(h2, B, &x0[0][0], &xs[0][0], &obs[0][0][0], NULL, &u0b[0][0], costb, stb, itb, NULL, NULL);
  if (rc == MPCB_OK) rc = mpcb_wait(h);
  if (rc == MPCB_OK) rc = mpcb_wait(h2);
  if (rc != MPCB_OK) { fprintf(stderr, "submit/wait: %s (%s)\n", mpcb_strerror(rc), mpcb_last_cuda_error()); return 5; }
  int same = memcmp(u0a, u0, sizeof u0) == 0 && memcmp(u0b, u0, sizeof u0) == 0 && memcmp(costa, cost, sizeof cost) == 0 &&
             memcmp(costb, cost, sizeof cost) == 0 && memcmp(sta, status, sizeof status) == 0 && memcmp(itb, iters, sizeof iters) == 0;
  printf("two handles in flight: %s\n", same ? "identical" : "DIFFERENT");
  mpcb_destroy(h2);
  mpcb_destroy(h);
  /* the reference's default scenario: cost 1.0947508e8, u0 = (0.0356462, 3.0) */
  int ok = status[0] == MPCB_CONVERGED && fabs(cost[0] - 1.0947508274e8) <= 1e-6 * 1.1e8 && fabs(u0[0][0] - 0.03564619) <= 1e-6 &&
           cost[2] == cost[0] && u0[2][0] == u0[0][0] && same;
  printf("%s\n", ok ? "OK" : "MISMATCH");
  return ok ? 0 : 1;
}
