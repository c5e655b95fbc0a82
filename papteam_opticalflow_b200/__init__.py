"""B200-native drop-in for the `pyflow.coarse2fine_flow` hot path of
ElijahHyndman/PAPTeam_OpticalFlow (Ce Liu's coarse-to-fine variational optical flow).

The arithmetic lives in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/pyflow_b200.h (built into papteam_opticalflow_b200/libpyflow_b200.so); this package is the
host-side mirror of the reference's Cython module (Par/pyflow.pyx).  There is no CPU fallback.
"""
from .pyflow import (coarse2fine_flow, coarse2fine_flow_batch, coarse2fine_flow_multigpu, multi_solve, sequence_flow, flow_to_bgr, decode_flow_u16, save_flow_u16, load_flow_u16,  # noqa: F401
                     set_solver_variant, get_solver_variant, FlowPlan, MODES)

__all__ = ["coarse2fine_flow", "coarse2fine_flow_batch", "multi_solve", "FlowPlan", "MODES"]
