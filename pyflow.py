"""`import pyflow` drop-in: same module and function name as the reference's Cython extension
(Par/pyflow.pyx), backed by the B200 CUDA library.  See papteam_opticalflow_b200/pyflow.py."""
from papteam_opticalflow_b200.pyflow import *  # noqa: F401,F403
from papteam_opticalflow_b200.pyflow import (coarse2fine_flow, coarse2fine_flow_batch, coarse2fine_flow_multigpu,  # noqa: F401
                                             multi_solve, sequence_flow, flow_to_bgr, decode_flow_u16, save_flow_u16, load_flow_u16, set_solver_variant, get_solver_variant, batch_last_stats, FlowPlan, MODES)
