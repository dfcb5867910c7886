"""Batched solve entry the reference lacks: B independent scenarios per call.

`BatchSolver.solve` is the batched counterpart of

    solver = mpc_solver.optimize_problem(ego_state, ref_state, obstacle)    PKG/main_cbf_kin_c_sim.py:99
    res = solver(x0=init_control, p=c_p, lbg=, lbx=, ubg=, ubx=)            PKG/main_cbf_kin_c_sim.py:100

It calls the C ABI of include/mpcb200.h.  torch is used only for device memory and
streams; numpy inputs take the host-pointer entry (copies inside the library).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .helpers import PACKAGE_PARAMS, load_config
from .problem import make_cfg


def _nx(kind):
    return 6 if kind == "dyn" else 4


class BatchSolver:
    """One handle on the current CUDA device for one NLP kind / horizon / obstacle count.

    `restoration=True` (kin-CBF kinds): a failed line search enters the restoration phase, as IPOPT does behind the
    reference's `nlpsol` call, instead of ending the solve with status 3; `resto_max_calls` caps how often (0 = no cap,
    IPOPT's behaviour).  Off by default here: on the benchmark batches it turns 1-2.5 % more scenarios into successes
    for 40-60 % more time.  The drop-in `MPC_optimize` classes switch it on without a cap."""

    def __init__(self, kind: str = "kin_cbf_pre", config: dict | None = None, N: int | None = None, M: int = 1,
                 init: str = "rollout", mu_init: float = 30.0, max_iter: int = 100, tol: float = 1e-8,
                 weights=None, bounds: dict | None = None, obs_input: str = "trajectory", cbf_gamma: float | None = None,
                 ref: str = "terminal", cfg_overrides: dict | None = None, dyn_bounds: str = "aligned",
                 restoration: bool = False, resto_max_calls: int = 1, engine: str = "auto",
                 integrator: str = "euler"):
        self.lib = _lib.load()
        self.kind = kind
        self.config = config if config is not None else load_config(PACKAGE_PARAMS)
        init_mode = {"rollout": _lib.INIT_ROLLOUT, "as_given": _lib.INIT_AS_GIVEN}[init]
        # "initial": obstacle states, rolled out in the kernel; "static": rows that hold at every step
        self.obs_initial = {"trajectory": False, "initial": True, "static": True}[obs_input]
        obs_code = {"trajectory": _lib.OBS_TRAJECTORY, "initial": _lib.OBS_INITIAL, "static": _lib.OBS_STATIC}[obs_input]
        self.ref_trajectory = {"terminal": False, "trajectory": True}[ref]
        self.cfg = make_cfg(kind, self.config, N=N, M=M, weights=weights, init_mode=init_mode, mu_init=mu_init,
                            max_iter=max_iter, tol=tol, bounds=bounds,
                            obs_input=obs_code,
                            cbf_gamma=cbf_gamma,
                            ref_mode=_lib.REF_TRAJECTORY if self.ref_trajectory else _lib.REF_TERMINAL,
                            dyn_rows={"aligned": _lib.DYN_ROWS_ALIGNED, "as_shipped": _lib.DYN_ROWS_AS_SHIPPED}[dyn_bounds],
                            restoration=restoration, resto_max_calls=resto_max_calls)
        # "warp": one scenario per warp (every kind); "lane": one scenario per lane (kinematic kinds, plain rows); "auto"
        # picks per batch size what measures faster
        self.cfg.engine = {"auto": _lib.ENGINE_AUTO, "warp": _lib.ENGINE_WARP, "lane": _lib.ENGINE_LANE}[engine]
        # "euler": the reference's defects (parity mode); "rk4": classical Runge-Kutta shooting defects (kinematic kinds, lane engine)
        self.cfg.integrator = {"euler": _lib.INTEGRATOR_EULER, "rk4": _lib.INTEGRATOR_RK4}[integrator]
        for key, val in (cfg_overrides or {}).items():  # any mpcb_cfg field, e.g. {"safe_l": 1.5, "T": 0.08, "Q": [...]}
            cur = getattr(self.cfg, key)
            if hasattr(cur, "__len__"):
                for i, v in enumerate(val):
                    cur[i] = v
            else:
                setattr(self.cfg, key, val)
        self.N, self.M = int(self.cfg.N), int(self.cfg.M)
        self.nx = _nx(kind)
        self.nv = 2 * self.N + self.nx * (self.N + 1)
        self.n_g = int(self.lib.mpcb_n_g(C.byref(self.cfg)))  # rows of g = len(res['g']) = len(res['lam_g'])
        # obs argument: (B,M,N+1,6) obs_prediction rows, or (B,M,6) obstacle states when obs_input="initial"
        self.obs_shape = (self.M, 6) if self.obs_initial else (self.M, self.N + 1, 6)
        # xs argument: (B,nx) target, or (B,N,nx) per-stage targets when ref="trajectory"
        self.xs_shape = (self.N, self.nx) if self.ref_trajectory else (self.nx,)
        self._h = C.c_void_p()
        _lib.check(self.lib.mpcb_create(C.byref(self.cfg), C.byref(self._h)), "mpcb_create")
        # the handle (slab, work queue) lives on the CUDA device that was current at creation
        self.device_index = None
        try:
            import torch

            if torch.cuda.is_available():
                self.device_index = torch.cuda.current_device()
        except ImportError:
            pass

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.mpcb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ device path
    def reserve(self, B: int):
        """Size the staging buffers of the host-pointer entry points for batches up to B (mpcb_reserve): no later host
        solve with a batch <= B allocates."""
        _lib.check(self.lib.mpcb_reserve(self._h, int(B)), "mpcb_reserve")

    def solve(self, x0, xs, obs=None, z_init=None, return_z: bool = False, return_lam: bool = False, return_duals: bool = False):
        """x0 (B,nx); xs (B,nx) [(B,N,nx) with ref="trajectory"]; obs (B,M,N+1,6) or None; z_init (B,nv) or None.

        torch CUDA tensors -> asynchronous on the current stream, returns torch tensors;
        numpy arrays -> synchronous host entry, returns numpy arrays.
        dict keys: u0 (B,2), cost (B,), status (B,) int32, iters (B,) int32 [, z (B,nv)] [, lam (B,nx(N+1))]
        [, lam_g (B,n_g), lam_x (B,nv) with return_duals: CasADi's res['lam_g'] / res['lam_x']]."""
        if isinstance(x0, np.ndarray):
            return self._solve_host(x0, xs, obs, z_init, return_z, return_lam, return_duals)
        import torch

        B = x0.shape[0]
        dev = x0.device
        if dev.type != "cuda":
            raise _lib.MpcbError("solve() needs CUDA tensors or numpy arrays; there is no CPU fallback")
        if self.device_index is not None and dev.index is not None and dev.index != self.device_index:
            raise _lib.MpcbError(f"this handle was created on cuda:{self.device_index}, the tensors are on cuda:{dev.index}")

        def prep(t, shape):
            if t is None:
                return None
            t = t.to(device=dev, dtype=torch.float64).contiguous()
            if tuple(t.shape) != shape:
                raise ValueError(f"expected shape {shape}, got {tuple(t.shape)}")
            return t

        x0 = prep(x0, (B, self.nx))
        xs = prep(xs, (B,) + self.xs_shape)
        obs = prep(obs, (B,) + self.obs_shape) if self.M > 0 else None
        z_init = prep(z_init, (B, self.nv))
        u0 = torch.empty((B, 2), dtype=torch.float64, device=dev)
        cost = torch.empty((B,), dtype=torch.float64, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        iters = torch.empty((B,), dtype=torch.int32, device=dev)
        z = torch.empty((B, self.nv), dtype=torch.float64, device=dev) if return_z else None
        lam = torch.empty((B, self.nx * (self.N + 1)), dtype=torch.float64, device=dev) if return_lam else None
        lam_g = torch.empty((B, self.n_g), dtype=torch.float64, device=dev) if return_duals else None
        lam_x = torch.empty((B, self.nv), dtype=torch.float64, device=dev) if return_duals else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            if return_duals:
                _lib.check(self.lib.mpcb_set_dual_outputs(self._h, ptr(lam_g), ptr(lam_x), 0), "mpcb_set_dual_outputs")
            try:
                rc = self.lib.mpcb_solve_batch(self._h, B, ptr(x0), ptr(xs), ptr(obs), ptr(z_init), ptr(u0), ptr(cost),
                                               ptr(status), ptr(iters), ptr(z), ptr(lam), stream)
            finally:
                if return_duals:
                    self.lib.mpcb_set_dual_outputs(self._h, None, None, 0)
        _lib.check(rc, "mpcb_solve_batch")
        out = {"u0": u0, "cost": cost, "status": status, "iters": iters}
        if return_duals:
            out["lam_g"], out["lam_x"] = lam_g, lam_x
        if return_z:
            out["z"] = z
        if return_lam:
            out["lam"] = lam
        return out

    # ------------------------------------------------------------------ host path
    def _solve_host(self, x0, xs, obs, z_init, return_z, return_lam, return_duals=False):
        B = x0.shape[0]
        f64 = lambda a, shape: None if a is None else np.ascontiguousarray(a, dtype=np.float64).reshape(shape)
        x0 = f64(x0, (B, self.nx))
        xs = f64(xs, (B,) + self.xs_shape)
        obs = f64(obs, (B,) + self.obs_shape) if self.M > 0 else None
        z_init = f64(z_init, (B, self.nv))
        u0 = np.empty((B, 2))
        cost = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        iters = np.empty(B, dtype=np.int32)
        z = np.empty((B, self.nv)) if return_z else None
        lam = np.empty((B, self.nx * (self.N + 1))) if return_lam else None
        lam_g = np.empty((B, self.n_g)) if return_duals else None
        lam_x = np.empty((B, self.nv)) if return_duals else None
        ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
        if return_duals:
            _lib.check(self.lib.mpcb_set_dual_outputs(self._h, ptr(lam_g), ptr(lam_x), 1), "mpcb_set_dual_outputs")
        try:
            rc = self.lib.mpcb_solve_batch_host(self._h, B, ptr(x0), ptr(xs), ptr(obs), ptr(z_init), ptr(u0), ptr(cost),
                                                ptr(status), ptr(iters), ptr(z), ptr(lam))
        finally:
            if return_duals:
                self.lib.mpcb_set_dual_outputs(self._h, None, None, 0)
        _lib.check(rc, "mpcb_solve_batch_host")
        out = {"u0": u0, "cost": cost, "status": status, "iters": iters}
        if return_duals:
            out["lam_g"], out["lam_x"] = lam_g, lam_x
        if return_z:
            out["z"] = z
        if return_lam:
            out["lam"] = lam
        return out

    def solve_host_ptrs(self, B, x0, xs, obs, z_init, u0, cost, status, iters):
        """Raw host-pointer call (pinned torch CPU tensors): used by bench.py's e2e timing."""
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = self.lib.mpcb_solve_batch_host(self._h, B, ptr(x0), ptr(xs), ptr(obs), ptr(z_init), ptr(u0), ptr(cost),
                                            ptr(status), ptr(iters), None, None)
        _lib.check(rc, "mpcb_solve_batch_host")

    def submit_host_ptrs(self, B, x0, xs, obs, z_init, u0, cost, status, iters):
        """Non-blocking half of `solve_host_ptrs` (mpcb_submit_batch_host): copy-in, solve and copy-out are queued on
        the handle's stream; the tensors (page-locked CPU tensors) must stay alive and unread until `wait()`."""
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        rc = self.lib.mpcb_submit_batch_host(self._h, B, ptr(x0), ptr(xs), ptr(obs), ptr(z_init), ptr(u0), ptr(cost),
                                             ptr(status), ptr(iters), None, None)
        _lib.check(rc, "mpcb_submit_batch_host")

    def wait(self):
        """Block until everything submitted on this handle has finished (mpcb_wait)."""
        _lib.check(self.lib.mpcb_wait(self._h), "mpcb_wait")

    def shift(self, x0, z):
        """In-place plant Euler step + warm-start shift on CUDA tensors (PKG/main_cbf_kin_c_sim.py:16-26)."""
        import torch

        B = x0.shape[0]
        stream = C.c_void_p(torch.cuda.current_stream(x0.device).cuda_stream)
        with torch.cuda.device(x0.device):
            rc = self.lib.mpcb_shift_batch(self._h, B, C.c_void_p(x0.data_ptr()), C.c_void_p(z.data_ptr()), stream)
        _lib.check(rc, "mpcb_shift_batch")

    def predict_obstacles(self, obs_state, want_traj: bool = True, advance: bool = False):
        """Batched `obs_prediction` on the device (mpcb_obs_prediction_batch; PKG/Obs_prediction.py:3-40): obs_state (B,M,6)
        float64 CUDA tensor -> (B,M,N+1,6); `advance` then moves the obstacles one step in place, the mains' update
        (PKG/main_cbf_kin_c_sim_pre.py:106)."""
        import torch

        assert obs_state.is_cuda and obs_state.dtype == torch.float64 and obs_state.is_contiguous() and obs_state.shape[-1] == 6
        n = obs_state.numel() // 6
        traj = torch.empty(tuple(obs_state.shape[:-1]) + (self.N + 1, 6), dtype=torch.float64, device=obs_state.device) if want_traj else None
        stream = C.c_void_p(torch.cuda.current_stream(obs_state.device).cuda_stream)
        with torch.cuda.device(obs_state.device):
            rc = self.lib.mpcb_obs_prediction_batch(self._h, n, C.c_void_p(obs_state.data_ptr()),
                                                    C.c_void_p(traj.data_ptr()) if traj is not None else None, int(advance), stream)
        _lib.check(rc, "mpcb_obs_prediction_batch")
        return traj

    def ref_traj(self, x0, xs, path_x0, last_idx, T_horizon: float, aa: float | None = None):
        """Batched `RefPathGenerator.find_ref_traj` on the device (PKG/RefPathGenerator.py:27-59) over the
        implicit straight path of `define_ref_path` (:9-24) that starts at x = path_x0[b].

        x0, xs (B,4) and path_x0 (B,) float64 CUDA tensors; last_idx (B,) int32 CUDA tensor, updated in
        place.  Returns ref (B,N+1,4); with `aa` also the stage targets aa*ref[i+1] + (1-aa)*xs (B,N,4),
        the `xs` argument of a solver built with ref="trajectory"."""
        import torch

        B = x0.shape[0]
        dev = x0.device
        assert self.nx == 4 and last_idx.dtype == torch.int32 and last_idx.is_contiguous()
        x0, xs, path_x0 = (t.to(torch.float64).contiguous() for t in (x0, xs, path_x0))
        ref = torch.empty((B, self.N + 1, 4), dtype=torch.float64, device=dev)
        stage = torch.empty((B, self.N, 4), dtype=torch.float64, device=dev) if aa is not None else None
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            rc = self.lib.mpcb_ref_traj_batch(self._h, B, float(T_horizon), C.c_void_p(x0.data_ptr()), C.c_void_p(xs.data_ptr()),
                                              C.c_void_p(path_x0.data_ptr()), C.c_void_p(last_idx.data_ptr()),
                                              float(aa if aa is not None else 0.0), C.c_void_p(ref.data_ptr()),
                                              C.c_void_p(stage.data_ptr()) if stage is not None else None, stream)
        _lib.check(rc, "mpcb_ref_traj_batch")
        return (ref, stage) if aa is not None else ref

    def set_order(self, order):
        """order: CUDA int32 tensor (B,), a permutation - queue position q processes scenario order[q]
        (longest expected first); None resets to arrival order.  Results do not depend on it."""
        if order is None:
            _lib.check(self.lib.mpcb_set_order(self._h, None, 0), "mpcb_set_order")
            self._order = None
        else:
            import torch

            assert order.is_cuda and order.dtype == torch.int32 and order.is_contiguous()
            _lib.check(self.lib.mpcb_set_order(self._h, C.c_void_p(order.data_ptr()), int(order.numel())), "mpcb_set_order")
            self._order = order  # keep it alive

    def set_trace(self, trace):
        """trace: CUDA float64 tensor (B, rows, 8) to receive the per-iteration log, or None."""
        if trace is None:
            _lib.check(self.lib.mpcb_set_trace_buffer(self._h, None, 0), "mpcb_set_trace_buffer")
            self._trace = None
        else:
            assert trace.is_cuda and trace.is_contiguous() and trace.shape[2] == 8
            _lib.check(self.lib.mpcb_set_trace_buffer(self._h, C.c_void_p(trace.data_ptr()), int(trace.shape[1])), "mpcb_set_trace_buffer")
            self._trace = trace  # keep it alive

    def debug_slot_errors(self) -> int:
        """Slot-ownership violations counted by a -DMPCB_DEBUG_SLOTS build (mpcb_debug_slot_errors); -1 otherwise."""
        n = C.c_int(0)
        _lib.check(self.lib.mpcb_debug_slot_errors(self._h, C.byref(n)), "mpcb_debug_slot_errors")
        return int(n.value)

    def launch_info(self) -> dict:
        info = _lib.MpcbLaunchInfo()
        _lib.check(self.lib.mpcb_get_launch_info(self._h, C.byref(info)), "mpcb_get_launch_info")
        return {k: int(getattr(info, k)) for k, _ in info._fields_}
