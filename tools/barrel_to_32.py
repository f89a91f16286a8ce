"""dev tool (GPU box): the 32 in-place barrel-roll problems at full caps on the GPU next to the committed four roundings of the oracle."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")
prob = cm.BarrelRollProblem(); fopt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
g = json.load(open(os.path.join(R, "tests/golden/barrel_to_four_roundings.json")))["builds"]
x0 = workload.mhpc_batch(32)
s = cm.MultiPhaseDDP(prob, 0, 32); s.set_initial_condition(x0); s.set_initial_guess(prob.initial_guess(x0)); s.solve(fopt)
info = s.get_solver_info()
for b in range(32):
    got = [info[b][k] for k in COUNTS]
    refs = {k: v[b] for k, v in g.items()}
    agree = len(set(tuple(v[0]) for v in refs.values())) == 1
    same = [k for k, v in refs.items() if v[0] == got]
    print(json.dumps({"b": b, "oracle_builds_agree": agree, "gpu": got, "gpu_cost": info[b]["cost"], "max_t": info[b]["max_tconstr"], "max_p": info[b]["max_pconstr"], "feas": info[b]["feas"], "equal_to": same,
                      "oracle": {k: [v[0][1], v[0][2], v[0][4], round(v[1], 6)] for k, v in refs.items()}}))
