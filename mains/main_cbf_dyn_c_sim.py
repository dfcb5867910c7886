"""PKG/main_cbf_dyn_c_sim.py with the plotting stripped: dynamic bicycle CBF MPC, static obstacle,
100 closed-loop steps, applied control zeroed at step 10 (the reference's disturbance, :98-100)."""
import time

import numpy as np
from _common import PARAMS_FILE, shift_movement, summary

import MPC_CBF_optimize_dyn
from helpers import load_config

if __name__ == "__main__":
    config = load_config(PARAMS_FILE)
    T_horizon, T_S = config["mpc_params"]["horizon"], config["mpc_params"]["T_S"]
    N_p = len(np.arange(0, T_horizon + T_S, T_S, dtype=float)) - 1
    mpc_solver = MPC_CBF_optimize_dyn.MPC_optimize()
    n_states, n_controls = mpc_solver.num_states, mpc_solver.num_controls
    t0 = 0.0
    x0 = np.array([0, 0, 0, 10, 0, 0]).reshape(-1, 1).astype(float)
    xs = np.array([600, 3.5, 0, 15, 0, 0]).reshape(-1, 1).astype(float)
    u0 = np.zeros((N_p, n_controls))
    next_states = np.zeros((N_p + 1, n_states))  # the reference's literal first guess (the dyn shim re-integrates all-zero states)
    obs = np.array([100, -3.5])
    lbg, ubg, lbx, ubx = mpc_solver.initialize_constraints()
    xh, uh, caltimeh, stats = [x0], [], [0], []
    sim_time, mpciter = 10, 0
    while mpciter - sim_time / T_S < 0.0:
        start_time = time.time()
        c_p = np.concatenate((x0, xs))
        init_control = np.concatenate((u0.reshape(-1, 1), next_states.reshape(-1, 1)))
        solver = mpc_solver.optimize_problem(ego_state=x0, ref_state=xs, obstacle=obs)
        res = solver(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        stats.append(solver.stats())
        solve_opt = res["x"].full()
        u0 = solve_opt[: N_p * n_controls].reshape(N_p, n_controls)
        x_m = solve_opt[N_p * n_controls:].reshape(N_p + 1, n_states)
        if mpciter == 10:
            u0[0, 0] = 0
            u0[0, 1] = 0
        uh.append(u0[0, :].copy())
        t0, x0, u0, next_states = shift_movement(T_S, t0, x0, u0, x_m, mpc_solver.f)
        x0 = np.reshape(x0, (-1, 1))
        xh.append(x0)
        mpciter += 1
        caltimeh.append((time.time() - start_time) * 1000)
    summary("main_cbf_dyn_c_sim", xh, uh, caltimeh, stats)
