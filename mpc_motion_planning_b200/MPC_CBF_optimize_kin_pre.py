"""Drop-in for PKG/MPC_CBF_optimize_kin_pre.py: same module name, `MPC_optimize` class, constructor,
`initialize_constraints`, `optimize_problem` and `.f`; the solve runs on the B200 through libmpcb200."""
from __future__ import annotations

from ._surface import MPCOptimizeBase, PARAMS_FILE, _Solver  # noqa: F401


class MPC_optimize(MPCOptimizeBase):
    KIND = "kin_cbf_pre"

    def initialize_constraints(self, obs_trajectories):
        """-> (lbg, ubg, lbx, ubx) lists with the reference's lengths and order."""
        return self._initialize_constraints(obs_trajectories)

    def optimize_problem(self, ego_state, ref_state, obs_trajectories):
        """Returns the solver callable.  `ego_state` is unused; `ref_state` enters the stage cost
        with weight `self.aa` (0.0 as shipped, so it has no effect unless the attribute is changed)."""
        return _Solver(self, self._obs_array(obs_trajectories), ref_state)
