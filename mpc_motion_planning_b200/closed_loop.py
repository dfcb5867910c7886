"""Batched closed-loop simulation on the device — the `while` loops of the reference mains
(PKG/main_cbf_kin_c_sim.py:87-123, main_cbf_kin_c_sim_pre.py:86-126, main_kin_c_sim.py:70-97,
main_cbf_dyn_c_sim.py:75-108) for B vehicles at once: solve, apply the first control with the
plant Euler step, shift the warm start, advance the obstacle one step.  Nothing leaves the GPU
between steps except what the caller asks to record.
"""
from __future__ import annotations

import torch


def predict_obstacles(obs_state: torch.Tensor, dt: float, N: int) -> torch.Tensor:
    """obs_state (B,M,6) rows [x,y,theta,v,l,w] -> (B,M,N+1,6), the constant-velocity roll-out of
    PKG/Obs_prediction.py:3-40 accumulated step by step (same operation order)."""
    B, M, _ = obs_state.shape
    out = obs_state[:, :, None, :].repeat(1, 1, N + 1, 1)
    dx = obs_state[..., 3] * torch.cos(obs_state[..., 2]) * dt
    dy = obs_state[..., 3] * torch.sin(obs_state[..., 2]) * dt
    x = obs_state[..., 0].clone()
    y = obs_state[..., 1].clone()
    for k in range(N + 1):
        out[:, :, k, 0] = x
        out[:, :, k, 1] = y
        x = x + dx
        y = y + dy
    return out.contiguous()


def run_closed_loop(solver, x0: torch.Tensor, xs: torch.Tensor, obs_state: torch.Tensor | None, steps: int,
                    moving: bool = True, disturbance_step: int | None = None, aa: float | None = None,
                    T_horizon: float | None = None, longest_first: bool = False):
    """Returns dict(x (steps+1,B,nx), u (steps,B,2), status (steps,B), iters (steps,B)).

    longest_first: order each step's work queue by the previous step's iteration counts (descending),
    which shortens the last wave at small batches (`mpcb_set_order`).
    aa: with a solver built with ref="trajectory", every step first runs the batched
    `find_ref_traj` on the device (the mains call it before `optimize_problem`,
    PKG/main_cbf_kin_c_sim_pre.py:97) and tracks aa*ref_traj[i+1] + (1-aa)*xs per stage.

    obs_state: (B,M,6) obstacle states (kin kinds; `moving=False` keeps them fixed as in
    main_cbf_kin_c_sim.py), (B,1,6) with only columns 0,1 used for dyn, or None for the no-CBF NLP.
    disturbance_step: zero the applied control at that step (PKG/main_cbf_dyn_c_sim.py:98-100)."""
    gen = _closed_loop_steps(solver, x0, xs, obs_state, steps, moving, disturbance_step, aa, T_horizon, longest_first)
    while True:
        try:
            next(gen)
        except StopIteration as done:
            return done.value


def run_closed_loop_lanes(solvers, x0: torch.Tensor, xs: torch.Tensor, obs_state: torch.Tensor | None, steps: int, **kw):
    """The same closed loop with the fleet split into len(solvers) contiguous groups, each on its own handle and
    stream: the groups' control steps are independent, so one group's last wave of long solves overlaps the other
    group's next step (batches in flight).  Same arguments and the same result, bit for bit, as
    `run_closed_loop` on the whole fleet."""
    L = len(solvers)
    dev = x0.device
    B = x0.shape[0]
    cur = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(L)]
    cuts = [(B * i) // L for i in range(L + 1)]
    gens = []
    for i, s in enumerate(solvers):
        lo, hi = cuts[i], cuts[i + 1]
        streams[i].wait_stream(cur)
        gens.append(_closed_loop_steps(s, x0[lo:hi], xs[lo:hi], None if obs_state is None else obs_state[lo:hi], steps,
                                       kw.get("moving", True), kw.get("disturbance_step"), kw.get("aa"), kw.get("T_horizon"),
                                       kw.get("longest_first", False)))
    parts = [None] * L
    for _ in range(steps + 1):           # `steps` yields, then the return
        for i, g in enumerate(gens):
            if parts[i] is None:
                with torch.cuda.stream(streams[i]):
                    try:
                        next(g)
                    except StopIteration as done:
                        parts[i] = done.value
    for st in streams:
        cur.wait_stream(st)
    for part in parts:
        for t in part.values():
            t.record_stream(cur)
    return {k: torch.cat([part[k] for part in parts], dim=1) for k in parts[0]}


def _closed_loop_steps(solver, x0, xs, obs_state, steps, moving, disturbance_step, aa, T_horizon, longest_first):
    """Generator behind both drivers: enqueues one control step per `next`, returns the histories."""
    dev = x0.device
    B = x0.shape[0]
    N, nv = solver.N, solver.nv
    dt = float(solver.cfg.T)
    x = x0.clone().to(torch.float64).contiguous()
    z = torch.zeros((B, nv), dtype=torch.float64, device=dev)  # zero first guess (PKG/main_cbf_kin_c_sim.py:47-50)
    xs = xs.to(torch.float64).contiguous()
    obs = None if obs_state is None else obs_state.clone().to(torch.float64).contiguous()
    xh = [x.clone()]
    uh, sth, ith = [], [], []
    stage_ref = aa is not None and solver.ref_trajectory
    if stage_ref:
        path_x0 = x[:, 0].clone()  # define_ref_path(x0, xs, T_S) before the loop (:52)
        last_idx = torch.zeros(B, dtype=torch.int32, device=dev)
        T_h = float(T_horizon if T_horizon is not None else solver.config["mpc_params"]["horizon"])
    try:
        for step in range(steps):
            target = xs
            if stage_ref:
                _, target = solver.ref_traj(x, xs, path_x0, last_idx, T_h, aa)
            traj = None
            if obs is not None and solver.obs_initial:
                traj = obs if moving else torch.cat([obs[..., :3], torch.zeros_like(obs[..., 3:4]), obs[..., 4:]], dim=-1)  # prediction in-kernel
            elif obs is not None:  # one kernel: the reference's obs_prediction for the whole fleet
                traj = solver.predict_obstacles(obs) if moving else obs[:, :, None, :].repeat(1, 1, N + 1, 1).contiguous()
            out = solver.solve(x, target, traj, z, return_z=True)
            z = out["z"]
            if disturbance_step is not None and step == disturbance_step:
                z[:, 0:2] = 0.0
            uh.append(z[:, 0:2].clone())
            sth.append(out["status"])
            ith.append(out["iters"])
            if longest_first:
                solver.set_order(torch.argsort(out["iters"], descending=True, stable=True).to(torch.int32))
            solver.shift(x, z)  # plant Euler step with U_0 and warm-start shift, in place
            if obs is not None and moving:
                solver.predict_obstacles(obs, want_traj=False, advance=True)  # main_cbf_kin_c_sim_pre.py:106, in place
            xh.append(x.clone())
            yield step
    finally:
        # an abandoned or failing loop must not leave its order installed on the handle
        if longest_first:
            solver.set_order(None)
    return {"x": torch.stack(xh), "u": torch.stack(uh), "status": torch.stack(sth), "iters": torch.stack(ith)}
