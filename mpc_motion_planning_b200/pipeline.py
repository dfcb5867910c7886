"""Several independent batches in flight: `PipelinedSolver`.

The reference solves one scenario per blocking `solver(...)` call (PKG/main_cbf_kin_c_sim.py:100).  A batch of
B scenarios on one handle ends with a last wave in which the few scenarios that need 80-100 iterations keep a
handful of warps busy while the rest of the GPU idles (about a quarter of the time at
B = 10,000).  When the caller has more than one batch per control period the idle part is recoverable without
touching the kernel: each lane owns a handle (its own stream, work queue and slab), and a batch submitted on one
lane fills the SMs the other lane's last wave has released.  Results are bit-identical to `BatchSolver`'s -- the
lanes run the same kernel on the same inputs.

    pipe = PipelinedSolver(2, "kin_cbf", obs_input="static")
    t0 = pipe.submit(x0_a, xs_a, obs_a)          # CUDA tensors; returns a ticket at once
    t1 = pipe.submit(x0_b, xs_b, obs_b)
    out_a = pipe.result(t0)                      # dict of tensors, valid on the caller's current stream

`submit_host` / `wait` are the same over page-locked host tensors through mpcb_submit_batch_host / mpcb_wait
(the C-ABI pair a non-Python host would use, include/mpcb200.h)."""
from __future__ import annotations

import torch

from .solver import BatchSolver


class PipelinedSolver:
    def __init__(self, lanes: int = 2, *solver_args, **solver_kw):
        if lanes < 1:
            raise ValueError("lanes must be >= 1")
        if not torch.cuda.is_available():
            raise RuntimeError("PipelinedSolver needs a CUDA device: there is no CPU fallback")
        self.solvers = [BatchSolver(*solver_args, **solver_kw) for _ in range(lanes)]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.streams = [torch.cuda.Stream(dev) for _ in range(lanes)]
        self._pending = [None] * lanes        # device path: (event, out) of the lane's batch in flight
        self._host_busy = [False] * lanes
        self._next = 0

    @property
    def lanes(self) -> int:
        return len(self.solvers)

    def _take_lane(self) -> int:
        lane = self._next
        self._next = (lane + 1) % self.lanes
        return lane

    # ---- device tensors
    def submit(self, x0, xs, obs=None, z_init=None, return_z: bool = False, return_lam: bool = False) -> int:
        """Queue one batch (CUDA tensors) on the next lane; the lane's stream first waits for the caller's current
        stream, so inputs produced there are seen.  Returns the ticket to pass to `result`."""
        lane = self._take_lane()
        st = self.streams[lane]
        st.wait_stream(torch.cuda.current_stream(x0.device))
        with torch.cuda.stream(st):
            out = self.solvers[lane].solve(x0, xs, obs, z_init, return_z=return_z, return_lam=return_lam)
            ev = torch.cuda.Event()
            ev.record(st)
        for t in (x0, xs, obs, z_init):
            if t is not None:
                t.record_stream(st)
        self._pending[lane] = (ev, out)
        return lane

    def result(self, ticket: int) -> dict:
        """Outputs of the batch behind `ticket`, ordered after the solve on the caller's current stream (no host sync)."""
        ev, out = self._pending[ticket]
        torch.cuda.current_stream().wait_event(ev)
        for t in out.values():
            t.record_stream(torch.cuda.current_stream())
        return out

    # ---- page-locked host tensors, through the C-ABI pair
    def submit_host(self, B, x0, xs, obs, z_init, u0, cost, status, iters) -> int:
        lane = self._take_lane()
        if self._host_busy[lane]:
            self.solvers[lane].wait()
        self.solvers[lane].submit_host_ptrs(B, x0, xs, obs, z_init, u0, cost, status, iters)
        self._host_busy[lane] = True
        return lane

    def wait(self, ticket: int | None = None):
        """Block until the batch behind `ticket` (or every batch in flight) has been copied back."""
        for lane in (range(self.lanes) if ticket is None else (ticket,)):
            if self._host_busy[lane]:
                self.solvers[lane].wait()
                self._host_busy[lane] = False

    def launch_info(self) -> dict:
        info = self.solvers[0].launch_info()
        info["launches"] = sum(s.launch_info()["launches"] for s in self.solvers)
        return info

    def close(self):
        for s in self.solvers:
            s.close()
