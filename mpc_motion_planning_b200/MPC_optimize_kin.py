"""Drop-in for the reference's missing-source module MPC_optimize_kin (imported by
PKG/main_kin_c_sim.py:2; recovered from PKG/__pycache__/MPC_optimize_kin.cpython-37.pyc):
kinematic tracking NLP without obstacle or rate rows."""
from __future__ import annotations

from ._surface import MPCOptimizeBase, PARAMS_FILE, _Solver  # noqa: F401


class MPC_optimize(MPCOptimizeBase):
    KIND = "kin_nocbf"

    def initialize_constraints(self):
        """-> (lbg, ubg, lbx, ubx); lbg = ubg = 0.0 scalars as in the reference."""
        return self._initialize_constraints(None)

    def optimize_problem(self, ego_state, ref_state):
        return _Solver(self, None)
