"""PKG/main_kin_c_sim.py with the plotting stripped: kinematic tracking MPC without obstacle rows
(module MPC_optimize_kin), up to 100 closed-loop steps."""
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
    x0 = np.array([0, 0, 0, 20]).reshape(-1, 1).astype(float)
    xs = np.array([500, 3.5, 0, 30]).reshape(-1, 1).astype(float)
    u0 = np.zeros((N_p, n_controls))
    next_states = np.zeros((N_p + 1, n_states))  # the reference's literal first guess: all zeros (x_m.copy().T)
    lbg, ubg, lbx, ubx = mpc_solver.initialize_constraints()
    xh, uh, caltimeh, stats = [x0], [], [0], []
    sim_time, mpciter = 10, 0
    while np.linalg.norm(x0 - xs) > 1e-2 and mpciter - sim_time / T_S < 0.0:
        start_time = time.time()
        c_p = np.concatenate((x0, xs))
        init_control = np.concatenate((u0.reshape(-1, 1), next_states.reshape(-1, 1)))
        solver = mpc_solver.optimize_problem(ego_state=x0, ref_state=xs)
        res = solver(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        stats.append(solver.stats())
        solve_opt = res["x"].full()
        u0 = solve_opt[: N_p * n_controls].reshape(N_p, n_controls)
        x_m = solve_opt[N_p * n_controls:].reshape(N_p + 1, n_states)
        uh.append(u0[0, :])
        t0, x0, u0, next_states = shift_movement(T_S, t0, x0, u0, x_m, mpc_solver.f)
        x0 = np.reshape(x0, (-1, 1))
        xh.append(x0)
        mpciter += 1
        caltimeh.append((time.time() - start_time) * 1000)
    summary("main_kin_c_sim", xh, uh, caltimeh, stats)
