"""cafe_mpc_b200 — B200-native batched HS-DDP hot path of CAFE-MPC."""
from .api import (HKDProblem, MHPCProblem, LocoProblem, BarrelRollProblem, MultiPhaseDDP, load_hsddp_setting, measure_fp64_peak,  # noqa: F401
                  unpack_lcm_command, unpack_solution)
