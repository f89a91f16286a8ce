"""dev tool (GPU box): per-iteration trace of GPU and oracle on the far-initial-state fault cases of tests/test_gpu_faults.py"""
import sys, os, copy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_solve
prob = cm.MHPCProblem(os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
opt = copy.copy(opt); opt.max_AL_iter = 2; opt.max_DDP_iter = 3
np.set_printoptions(linewidth=250, precision=9)
col = int(os.environ.get("COL", "0"))
for off in [float(a) for a in sys.argv[1:]]:
    x0 = workload.mhpc_batch(3); x0[:, col] += off
    s = cm.MultiPhaseDDP(prob, 0, 3); s.set_initial_condition(x0); s.solve(opt)
    info = s.get_solver_info(); trace = s.get_trace(64); hist = s.get_history(64)
    for b in range(2):
        oi, oh, ot, _ = oracle_solve(prob.deck, opt, x0[b])
        keys = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist", "cost")
        print("off", off, "b", b, "oracle", [oi[k] for k in keys], "gpu", [info[b][k] for k in keys])
        n = max(oi["iter"], info[b]["iter"])
        print("oracle trace\n", ot[:n]); print("gpu trace\n", trace[b, :info[b]["iter"]])
