"""Per-knot comparison GPU vs oracle after exactly one LQ + backward sweep + linear rollout (dev tool)."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_solve, oracle_get
kind = sys.argv[1] if len(sys.argv) > 1 else "hkd"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
csv = os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv")
if kind == "hkd":
    prob = cm.HKDProblem(csv, k0=k0); opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "HKDMPC/settings/ddp_setting.info")); x0 = workload.hkd_batch(prob, 4)
else:
    prob = cm.MHPCProblem(csv, k0=k0); opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "MHPC/settings/ddp_setting.info")); x0 = workload.mhpc_batch(4)
opt.max_DDP_iter = 1; opt.max_AL_iter = 1; opt.cost_thresh = 1e30; opt.dynamics_feas_thresh = 1e30
s = cm.MultiPhaseDDP(prob, 0, 4); s.set_initial_condition(x0); s.solve(opt)
oi, oh, ot, osol = oracle_solve(prob.deck, opt, x0[b])
print(s.get_solver_info()[b]); print(oi)
for ph in range(len(prob.phases())):
    for name in ["X","U","Y","Defect","l","lx","lu","ly","lxx","luu","lyy","A","B","C","D","Phix","Phixx","Px","Quu","Qux","Qu","K","dU","G","dX"]:
        g = s.debug_get(name, ph, b); o = oracle_get(name, ph)
        if len(g) != len(o): print(ph, name, "LEN", len(g), len(o)); continue
        if len(g) == 0: continue
        err = np.abs(g - o); i = int(np.argmax(err))
        print(ph, name, "maxabs %.3e at %d (g %.6g o %.6g) scale %.3g" % (err.max(), i, g[i], o[i], np.abs(o).max()))
