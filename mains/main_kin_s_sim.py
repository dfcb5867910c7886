"""PKG/main_kin_s_sim.py with the plotting stripped: ONE solve of the kinematic tracking MPC without obstacle rows
(module MPC_optimize_kin) from the reference's all-zero guess, then one plant step (PKG/main_kin_s_sim.py:40-100)."""
import time

import numpy as np
from _common import PARAMS_FILE, shift_movement, summary

import MPC_optimize_kin
from helpers import load_config

if __name__ == "__main__":
    config = load_config(PARAMS_FILE)
    T_horizon, T_S = config["mpc_params"]["horizon"], config["mpc_params"]["T_S"]
    N_p = len(np.arange(0, T_horizon + T_S, T_S, dtype=float)) - 1
    mpc_solver = MPC_optimize_kin.MPC_optimize()
    n_states, n_controls = mpc_solver.num_states, mpc_solver.num_controls
    t0 = 0.0
    x0 = np.array([0, 0, 0, 20]).reshape(-1, 1).astype(float)   # :42
    xs = np.array([500, 3.5, 0, 30]).reshape(-1, 1).astype(float)  # :46
    next_states = np.zeros((N_p + 1, n_states))                 # x_m.copy().T, :44-45
    u0 = np.zeros((n_controls, N_p))                            # np.array([0, 0]*N_p).reshape(-1, 2).T, :47
    lbg, ubg, lbx, ubx = mpc_solver.initialize_constraints()    # :63
    xh, uh, caltimeh, stats = [x0], [], [0], []
    start_time = time.time()
    c_p = np.concatenate((x0, xs))
    init_control = np.concatenate((u0.reshape(-1, 1), next_states.reshape(-1, 1)))
    solver = mpc_solver.optimize_problem(ego_state=x0, ref_state=xs)                 # :83
    res = solver(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)         # :84
    stats.append(solver.stats())
    solve_opt = res["x"].full()
    u0 = solve_opt[: N_p * n_controls].reshape(N_p, n_controls)
    x_m = solve_opt[N_p * n_controls:].reshape(N_p + 1, n_states)
    uh.append(u0[0, :])
    t0, x0, u0, next_states = shift_movement(T_S, t0, x0, u0, x_m, mpc_solver.f)
    xh.append(np.reshape(x0, (-1, 1)))
    caltimeh.append((time.time() - start_time) * 1000)
    stats.append(stats[0])  # summary() averages the warm-started steps; there are none in the single-shot main
    summary("main_kin_s_sim", xh, uh, caltimeh, stats)
    print(f"predicted end of horizon: {np.round(x_m[-1], 3)}; f = {float(res['f']):.6e}; "
          f"max |lam_g| {np.abs(res['lam_g'].full()).max():.3e}, max |lam_x| {np.abs(res['lam_x'].full()).max():.3e}")
