"""dev tool (GPU box): per-knot LQ / sweep parity of the in-place barrel-roll deck after n DDP iterations (GPU vs oracle)."""
import copy, os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_get, oracle_solve
def relerr(g, o, floor=1e-6):
    o = np.asarray(o); g = np.asarray(g)
    return 0.0 if o.size == 0 else float(np.max(np.abs(g - o)) / max(np.max(np.abs(o)), floor))
prob = cm.BarrelRollProblem()
opt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
x0 = workload.mhpc_batch(2)
guess = prob.initial_guess(x0)
for nit in (int(a) for a in sys.argv[1:] or ["1", "2", "3"]):
    o1 = copy.copy(opt)
    o1.max_DDP_iter = nit; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    s = cm.MultiPhaseDDP(prob, 0, 2)
    s.set_initial_condition(x0); s.set_initial_guess(guess); s.solve(o1)
    oi, oh, ot, osol = oracle_solve(prob.deck, o1, x0[0], cap=320, guess=guess[0])
    info = s.get_solver_info()[0]; hist = s.get_history(320)[0]
    print("iters", nit, "gpu", info, "\n oracle", oi)
    print(" hist gpu", hist[:info["n_hist"], 0], "\n hist ora", oh[:, 0])
    tr = s.get_trace(320)[0]
    print(" trace gpu", tr[:info["iter"]], "\n trace ora", ot)
    worst = {}
    for ph in range(6):
        for name in ("X", "U", "Y", "Defect", "l", "lx", "lu", "ly", "lxx", "luu", "lyy", "A", "B", "C", "D", "Phix", "Phixx", "Quu", "Qux", "Qu", "K", "dU", "G", "dX") + (("Px",) if ph < 5 else ()):
            try:
                e = relerr(s.debug_get(name, ph, 0), oracle_get(name, ph))
            except Exception as ex:
                e = str(ex)
            if isinstance(e, str) or e > 1e-11: worst[(ph, name)] = e
    print(" deviations > 1e-11:", worst)
    s.close()
