"""Shared pieces of the closed-loop mains: the reference's `shift_movement`
(PKG/main_cbf_kin_c_sim.py:16-26) and a text summary in place of the matplotlib figures."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpc_motion_planning_b200  # noqa: E402

mpc_motion_planning_b200.install_reference_names()  # `import MPC_CBF_optimize_kin` as in the reference

PARAMS_FILE = "mpc_parameters.yaml"


def shift_movement(T, t0, x0, u, x_f, f):
    f_value = f(x0, u[0, :])
    st = x0 + T * f_value.full()
    t = t0 + T
    u_end = np.concatenate((u[1:], u[-1:]))
    x_f = np.concatenate((x_f[1:], x_f[-1:]), axis=0)
    return t, st, u_end, x_f


def summary(name, xh, uh, caltimeh, stats):
    np.set_printoptions(linewidth=200)
    xh = np.array(xh).reshape(len(xh), -1)
    uh = np.array(uh)
    ok = sum(1 for s in stats if s["success"])
    print(f"{name}: {len(uh)} MPC steps, {ok} solved; final state {np.round(xh[-1], 3)}; "
          f"|df| max {np.abs(uh[:, 0]).max():.4f} rad, ax in [{uh[:, 1].min():.2f}, {uh[:, 1].max():.2f}]; "
          f"step time mean {np.mean(caltimeh[1:]):.2f} ms, max {np.max(caltimeh[1:]):.2f} ms; "
          f"IPM iterations first/rest {stats[0]['iter_count']}/{np.mean([s['iter_count'] for s in stats[1:]]):.1f}")
