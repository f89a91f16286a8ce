"""Quick GPU-vs-oracle check on a small perturbed HKD batch (dev tool; the real tests are tests/)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle_bindings import oracle_solve

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prob = cm.HKDProblem(os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "HKDMPC/settings/ddp_setting.info"))
x0 = workload.hkd_batch(prob, B)
s = cm.MultiPhaseDDP(prob, 0, B)
s.set_initial_condition(x0)
s.set_profiling(True)
t = time.time(); s.solve(opt); print("gpu solve s", time.time() - t)
print(s.get_timing())
info = s.get_solver_info(); hist = s.get_history(256); trace = s.get_trace(256); sol = s.get_solution()
np.set_printoptions(linewidth=220, precision=6)
worst = 0
for b in range(min(B, 16)):
    oi, oh, ot, os_ = oracle_solve(prob.deck, opt, x0[b])
    gi = info[b]
    same = all(gi[k] == oi[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist"))
    nh = min(gi["n_hist"], oi["n_hist"])
    herr = np.max(np.abs(hist[b, :nh] - oh[:nh]) / (np.abs(oh[:nh]) + 1e-12)) if nh else 0
    serr = np.max(np.abs(sol[b] - os_) / (np.abs(os_) + 1e-9))
    worst = max(worst, herr)
    print(b, "counts_equal", same, {k: (gi[k], oi[k]) for k in ("iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")},
          "hist_relerr %.2e sol_relerr %.2e cost %.9g %.9g" % (herr, serr, gi["cost"], oi["cost"]))
    if not same or herr > 1e-6:
        n = min(gi["iter"], oi["iter"], 6)
        print("gpu trace\n", trace[b, :n]); print("oracle trace\n", ot[:n])
print("worst hist relerr", worst)
