"""ctypes binding of libmpcb200.so (include/mpcb200.h).  Fails loudly when the CUDA
library is missing or no GPU is present: there is no CPU fallback in the product path."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MPCB200_LIB") or os.path.join(HERE, "libmpcb200.so")  # env override: A/B builds

MODEL_KIN, MODEL_DYN = 0, 1
OBS_NONE, OBS_ELLIPSE, OBS_SQRT, OBS_DCBF = 0, 1, 2, 3
INIT_AS_GIVEN, INIT_ROLLOUT = 0, 1
OBS_TRAJECTORY, OBS_INITIAL, OBS_STATIC = 0, 1, 2
REF_TERMINAL, REF_TRAJECTORY = 0, 1
DYN_ROWS_ALIGNED, DYN_ROWS_AS_SHIPPED = 0, 1
ENGINE_AUTO, ENGINE_WARP, ENGINE_LANE = 0, 1, 2
INTEGRATOR_EULER, INTEGRATOR_RK4 = 0, 1
ST_CONVERGED, ST_ACCEPTABLE, ST_MAXITER, ST_INFEASIBLE, ST_NAN, ST_RESTO_FAILED = 0, 1, 2, 3, 4, 5

# IPOPT return_status strings (CasADi `solver.stats()['return_status']`)
RETURN_STATUS = {
    ST_CONVERGED: "Solve_Succeeded",
    ST_ACCEPTABLE: "Solved_To_Acceptable_Level",
    ST_MAXITER: "Maximum_Iterations_Exceeded",
    ST_INFEASIBLE: "Infeasible_Problem_Detected",
    ST_NAN: "Invalid_Number_Detected",
    ST_RESTO_FAILED: "Restoration_Failed",
}


class MpcbCfg(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("N", C.c_int32), ("M", C.c_int32), ("obs_mode", C.c_int32),
        ("du0_cost", C.c_int32), ("n_rate", C.c_int32), ("rate_ctrl", C.c_int32 * 2),
        ("init_mode", C.c_int32), ("max_iter", C.c_int32),
        ("T", C.c_double), ("Q", C.c_double * 6), ("R", C.c_double * 2), ("DR", C.c_double * 2),
        ("rate_lo", C.c_double * 2), ("rate_hi", C.c_double * 2),
        ("u_lo", C.c_double * 2), ("u_hi", C.c_double * 2),
        ("x_lo", C.c_double * 6), ("x_hi", C.c_double * 6),
        ("obs_lo", C.c_double),
        ("ego_hl", C.c_double), ("ego_hw", C.c_double), ("safe_l", C.c_double), ("safe_w", C.c_double),
        ("dyn_sx", C.c_double), ("dyn_sy", C.c_double),
        ("Veh_l", C.c_double), ("Veh_lf", C.c_double), ("Veh_lr", C.c_double), ("Veh_m", C.c_double),
        ("Veh_Iz", C.c_double), ("aopt_f", C.c_double), ("aopt_r", C.c_double),
        ("Fymax_f", C.c_double), ("Fymax_r", C.c_double),
        ("tol", C.c_double), ("mu_init", C.c_double), ("bound_relax", C.c_double),
        ("obs_input", C.c_int32), ("ref_mode", C.c_int32), ("cbf_gamma", C.c_double),
        ("dyn_rows", C.c_int32), ("restoration", C.c_int32), ("resto_max_calls", C.c_int32), ("engine", C.c_int32), ("integrator", C.c_int32), ("reserved", C.c_int32),
    ]


class MpcbLaunchInfo(C.Structure):
    _fields_ = [("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32), ("regs_per_thread", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("num_sms", C.c_int32), ("launches", C.c_int64)]


EXPORTS = [
    "mpcb_version", "mpcb_strerror", "mpcb_last_cuda_error", "mpcb_nx", "mpcb_nv", "mpcb_create", "mpcb_destroy",
    "mpcb_workspace_bytes", "mpcb_solve_batch", "mpcb_solve_batch_host", "mpcb_submit_batch_host", "mpcb_wait", "mpcb_shift_batch", "mpcb_ref_traj_batch", "mpcb_get_launch_info", "mpcb_fp64_peak_tflops", "mpcb_set_trace_buffer", "mpcb_set_order",
    "mpcb_reserve", "mpcb_n_g", "mpcb_set_dual_outputs", "mpcb_debug_slot_errors", "mpcb_obs_prediction_batch",
]

_lib = None


class MpcbError(RuntimeError):
    pass


def load():
    """Load libmpcb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MpcbError(
            f"{SO_PATH} is missing: build the CUDA library first (python -m mpc_motion_planning_b200.build). "
            "This package has no CPU fallback."
        )
    lib = C.CDLL(SO_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses (device or host)
    lib.mpcb_version.restype = C.c_int
    lib.mpcb_strerror.restype = C.c_char_p
    lib.mpcb_strerror.argtypes = [C.c_int]
    lib.mpcb_last_cuda_error.restype = C.c_char_p
    lib.mpcb_nx.argtypes = [C.POINTER(MpcbCfg)]
    lib.mpcb_nv.argtypes = [C.POINTER(MpcbCfg)]
    lib.mpcb_create.argtypes = [C.POINTER(MpcbCfg), C.POINTER(vp)]
    lib.mpcb_destroy.argtypes = [vp]
    lib.mpcb_destroy.restype = None
    lib.mpcb_workspace_bytes.argtypes = [C.POINTER(MpcbCfg), C.c_int, C.POINTER(C.c_size_t)]
    lib.mpcb_solve_batch.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, ip, ip, dp, dp, vp]
    lib.mpcb_solve_batch_host.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, ip, ip, dp, dp]
    lib.mpcb_submit_batch_host.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp, ip, ip, dp, dp]
    lib.mpcb_wait.argtypes = [vp]
    lib.mpcb_shift_batch.argtypes = [vp, C.c_int, dp, dp, vp]
    lib.mpcb_ref_traj_batch.argtypes = [vp, C.c_int, C.c_double, dp, dp, dp, ip, C.c_double, dp, dp, vp]
    lib.mpcb_get_launch_info.argtypes = [vp, C.POINTER(MpcbLaunchInfo)]
    lib.mpcb_fp64_peak_tflops.argtypes = [C.POINTER(C.c_double)]
    lib.mpcb_set_trace_buffer.argtypes = [vp, dp, C.c_int]
    lib.mpcb_set_order.argtypes = [vp, ip, C.c_int]
    lib.mpcb_reserve.argtypes = [vp, C.c_int]
    lib.mpcb_n_g.argtypes = [C.POINTER(MpcbCfg)]
    lib.mpcb_set_dual_outputs.argtypes = [vp, dp, dp, C.c_int]
    lib.mpcb_debug_slot_errors.argtypes = [vp, C.POINTER(C.c_int)]
    lib.mpcb_obs_prediction_batch.argtypes = [vp, C.c_int, dp, dp, C.c_int, vp]
    _lib = lib
    return lib


def check(rc: int, what: str = "mpcb"):
    if rc != 0:
        lib = load()
        msg = lib.mpcb_strerror(rc).decode()
        cu = lib.mpcb_last_cuda_error().decode()
        raise MpcbError(f"{what} failed: {msg} (code {rc})" + (f" [{cu}]" if cu and rc == -2 else ""))
