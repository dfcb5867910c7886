"""PKG/main_cbf_kin_c_sim_pre.py with the plotting stripped: kinematic CBF MPC, one moving
obstacle re-predicted every step, 80 closed-loop steps.  Same module names, same call protocol;
the solve runs on the GPU through libmpcb200.  Run from any directory."""
import time

import numpy as np
from _common import PARAMS_FILE, shift_movement, summary

import MPC_CBF_optimize_kin_pre
import RefPathGenerator
from helpers import load_config
from Obs_prediction import obs_prediction

if __name__ == "__main__":
    config = load_config(PARAMS_FILE)
    T_horizon, T_S = config["mpc_params"]["horizon"], config["mpc_params"]["T_S"]
    N_p = len(np.arange(0, T_horizon + T_S, T_S, dtype=float)) - 1
    mpc_solver = MPC_CBF_optimize_kin_pre.MPC_optimize()
    n_states, n_controls = mpc_solver.num_states, mpc_solver.num_controls
    t0 = 0.0
    x0 = np.array([0, 3, 0, 15]).reshape(-1, 1).astype(float)
    xs = np.array([400, 3.5, 0, 30]).reshape(-1, 1).astype(float)
    u0 = np.zeros((N_p, n_controls))
    next_states = np.zeros((N_p + 1, n_states))  # the reference's literal first guess: all zeros (x_m.copy().T)
    ref = RefPathGenerator.RefPathGenerator()
    ref.define_ref_path(x0, xs, T_S)
    obs = [np.array([[50, 3.5, 0, 10, 4.8, 1.8]])]
    lbg, ubg, lbx, ubx = mpc_solver.initialize_constraints(obs)
    xh, uh, caltimeh, stats = [x0], [], [0], []
    sim_time, mpciter, last_idx = 8, 0, 0
    while mpciter - sim_time / T_S < 0.0:
        start_time = time.time()
        c_p = np.concatenate((x0, xs))
        init_control = np.concatenate((u0.reshape(-1, 1), next_states.reshape(-1, 1)))
        ref_traj, last_idx = ref.find_ref_traj(x0, xs, T_horizon, T_S, last_idx)
        obs_trajectories_dyn = obs_prediction(obs, T_S, N_p)
        solver = mpc_solver.optimize_problem(ego_state=x0, ref_state=ref_traj, obs_trajectories=obs_trajectories_dyn)
        res = solver(x0=init_control, p=c_p, lbg=lbg, lbx=lbx, ubg=ubg, ubx=ubx)
        stats.append(solver.stats())
        solve_opt = res["x"].full()
        u0 = solve_opt[: N_p * n_controls].reshape(N_p, n_controls)
        x_m = solve_opt[N_p * n_controls:].reshape(N_p + 1, n_states)
        obs = [obs_trajectories_dyn[0][1].reshape(1, -1)]
        uh.append(u0[0, :])
        t0, x0, u0, next_states = shift_movement(T_S, t0, x0, u0, x_m, mpc_solver.f)
        x0 = np.reshape(x0, (-1, 1))
        xh.append(x0)
        mpciter += 1
        caltimeh.append((time.time() - start_time) * 1000)
    summary("main_cbf_kin_c_sim_pre", xh, uh, caltimeh, stats)
