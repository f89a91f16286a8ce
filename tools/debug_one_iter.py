import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_solve, oracle_get
prob = cm.HKDProblem(os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "HKDMPC/settings/ddp_setting.info"))
opt.max_DDP_iter = 1; opt.max_AL_iter = 1; opt.cost_thresh = 1e30; opt.dynamics_feas_thresh = 1e30
x0 = workload.hkd_batch(prob, 4)
s = cm.MultiPhaseDDP(prob, 0, 4); s.set_initial_condition(x0); s.solve(opt)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
oi, oh, ot, osol = oracle_solve(prob.deck, opt, x0[b])
print(s.get_solver_info()[b]); print(oi)
for ph in range(3):
    for name in ["X","U","Defect","l","lx","lu","lxx","luu","A","B","Phix","Phixx","Px","Quu","Qux","Qu","K","dU","G","dX"]:
        g = s.debug_get(name, ph, b); o = oracle_get(name, ph)
        if len(g) != len(o): print(ph, name, "LEN", len(g), len(o)); continue
        if len(g) == 0: continue
        err = np.abs(g - o); i = int(np.argmax(err))
        print(ph, name, "maxabs %.3e at %d (g %.6g o %.6g) scale %.3g" % (err.max(), i, g[i], o[i], np.abs(o).max()))
