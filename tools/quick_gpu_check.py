"""Quick GPU-vs-oracle check on a small perturbed batch (dev tool; the real tests are tests/)."""
import os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_solve

kind = sys.argv[1] if len(sys.argv) > 1 else "hkd"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
csv = os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv")
if kind == "hkd":
    prob = cm.HKDProblem(csv, k0=k0); opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "HKDMPC/settings/ddp_setting.info")); x0 = workload.hkd_batch(prob, B)
else:
    prob = cm.MHPCProblem(csv, k0=k0); opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "MHPC/settings/ddp_setting.info")); x0 = workload.mhpc_batch(B)
s = cm.MultiPhaseDDP(prob, 0, B)
s.set_initial_condition(x0)
s.set_profiling(True)
t = time.time(); s.solve(opt); print("gpu solve s", time.time() - t)
print(s.get_timing())
info = s.get_solver_info(); hist = s.get_history(256); trace = s.get_trace(256); sol = s.get_solution()
np.set_printoptions(linewidth=220, precision=6)
worst = 0; worst_s = 0
for b in range(min(B, 16)):
    oi, oh, ot, os_ = oracle_solve(prob.deck, opt, x0[b])
    gi = info[b]
    same = all(gi[k] == oi[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist"))
    nh = min(gi["n_hist"], oi["n_hist"])
    herr = np.max(np.abs(hist[b, :nh, 0] - oh[:nh, 0]) / (np.abs(oh[:nh, 0]) + 1e-300)) if nh else 0
    gp, op = cm.unpack_solution(prob.deck, sol[b]), cm.unpack_solution(prob.deck, os_)
    serr = max(np.max(np.abs(pg[n] - po[n])) / max(np.max(np.abs(po[n])), 1e-300) for pg, po in zip(gp, op) for n in ("Xbar", "Ubar", "K", "dU", "Quu", "Qux", "G") if po[n].size)
    worst = max(worst, herr); worst_s = max(worst_s, serr)
    print(b, "counts_equal", same, {k: (gi[k], oi[k]) for k in ("iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")},
          "cost_hist_relerr %.2e sol_normrelerr %.2e cost %.12g %.12g" % (herr, serr, gi["cost"], oi["cost"]))
    if not same:
        n = min(gi["iter"], oi["iter"], 8)
        print("gpu trace\n", trace[b, :n + 1]); print("oracle trace\n", ot[:n + 1])
print("worst cost-history relerr", worst, "worst solution norm-relerr", worst_s)
