"""Small fixed workload for ncu captures: one batched HKD solve (dev tool)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
prob = cm.HKDProblem(os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(R, "data/HKDMPC/settings/ddp_setting.info"))
opt.max_AL_iter = 1; opt.max_DDP_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x0 = workload.hkd_batch(prob, 64)
import numpy as np
x0 = np.tile(x0, ((B + 63) // 64, 1))[:B]
s = cm.MultiPhaseDDP(prob, 0, B); s.set_initial_condition(x0); s.solve(opt)
print("ok", s.solve_ms(), s.get_timing()["launches"])
