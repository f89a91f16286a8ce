"""Fixed workload for ncu captures: ONE batched solve of the bench workload through the C ABI (dev tool).
usage: profile_cmd.py [mhpc|hkd] [B] [max_AL] [max_DDP]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
kind = sys.argv[1] if len(sys.argv) > 1 else "mhpc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")
if kind == "hkd":
    prob = cm.HKDProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/HKDMPC/settings/ddp_setting.info")); x0 = workload.hkd_batch(prob, min(B, 256))
else:
    prob = cm.MHPCProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info")); x0 = workload.mhpc_batch(min(B, 256))
if len(sys.argv) > 3: opt.max_AL_iter = int(sys.argv[3])
if len(sys.argv) > 4: opt.max_DDP_iter = int(sys.argv[4])
x0 = np.tile(x0, ((B + len(x0) - 1) // len(x0), 1))[:B]
s = cm.MultiPhaseDDP(prob, 0, B); s.set_initial_condition(x0); s.solve(opt)
print("ok solve_ms", s.solve_ms(), s.get_timing()["launches"])
info = s.get_solver_info()
print("iters", sum(i["iter"] for i in info) / B, "sweeps(reg_iter_total)", sum(i["reg_iter_total"] for i in info) / B, "ls", sum(i["ls_iter_total"] for i in info) / B)
