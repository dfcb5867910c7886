"""Drop-in for PKG/MPC_CBF_optimize_dyn.py: same module name, `MPC_optimize` class, constructor,
`initialize_constraints`, `optimize_problem` and `.f`; the solve runs on the B200 through libmpcb200."""
from __future__ import annotations

from ._surface import MPCOptimizeBase, PARAMS_FILE, _Solver  # noqa: F401


class MPC_optimize(MPCOptimizeBase):
    KIND = "dyn"

    def initialize_constraints(self):
        """-> (lbg, ubg, lbx, ubx) lists with the reference's lengths and order."""
        return self._initialize_constraints(None)

    def optimize_problem(self, ego_state, ref_state, obstacle):
        """Returns the solver callable.  `ego_state` is unused and `ref_state` has zero weight in
        the reference too (aa = 0.0)."""
        return _Solver(self, self._obs_array(obstacle))
