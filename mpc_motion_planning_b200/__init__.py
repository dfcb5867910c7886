"""B200-native batched nonlinear-MPC solver for the vehicle motion-planning NLPs of
ZhuorenLi/MPC_motion_planning (host mirror of the reference surface + CUDA solve path)."""
import importlib
import sys

__version__ = "0.1.0"

REFERENCE_MODULES = ("MPC_CBF_optimize_kin", "MPC_CBF_optimize_kin_pre", "MPC_CBF_optimize_dyn", "MPC_optimize_kin",
                     "RefPathGenerator", "Obs_prediction", "helpers")


def install_reference_names():
    """Register the drop-in modules under the reference's flat names, so that code written against
    the reference (`import MPC_CBF_optimize_kin`, `from Obs_prediction import obs_prediction`, ...)
    runs unchanged."""
    for name in REFERENCE_MODULES:
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")
