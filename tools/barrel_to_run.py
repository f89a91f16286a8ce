"""dev tool (GPU box): the in-place barrel roll (BarrelRollTO.cpp) at its full 30 x 10 iteration caps on a perturbed batch; prints the solve
time, the counters, and the deviation of problem 0 from the CPU oracle. usage: barrel_to_run.py [B] [check_oracle 0|1]"""
import json, os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
check = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prob = cm.BarrelRollProblem()
opt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
x0 = workload.mhpc_batch(B)
guess = prob.initial_guess(x0)
s = cm.MultiPhaseDDP(prob, 0, B)
s.set_initial_condition(x0); s.set_initial_guess(guess)
s.solve(opt)
t = time.time(); s.solve(opt); wall = time.time() - t
info = s.get_solver_info()
out = {"B": B, "solve_s": wall, "solves_per_s": B / wall, "mean_iter": float(np.mean([i["iter"] for i in info])), "info0": info[0],
       "cost_min_max": [min(i["cost"] for i in info), max(i["cost"] for i in info)], "status_counts": {str(k): sum(1 for i in info if i["status"] == k) for k in (0, 1, 2)}}
if check:
    from oracle_bindings import oracle_solve
    t = time.time()
    oi, oh, ot, osol = oracle_solve(prob.deck, opt, x0[0], cap=320, guess=guess[0])
    out["oracle_s"] = time.time() - t
    out["oracle_info0"] = oi
    hist = s.get_history(320)
    n = min(oi["n_hist"], info[0]["n_hist"])
    out["hist_max_rel_dev"] = float(np.max(np.abs(hist[0, :n, 0] - oh[:n, 0]) / np.abs(oh[:n, 0])))
    out["counters_equal"] = all(info[0][k] == oi[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist"))
print(json.dumps(out))
