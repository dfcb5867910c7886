"""B200-native batched nonlinear-MPC solver for the vehicle motion-planning NLPs of
ZhuorenLi/MPC_motion_planning (host mirror of the reference surface + CUDA solve path)."""
__version__ = "0.1.0"
