"""Per-launch kernel times against the batch size (dev tool): tells whether a kernel is bound by its own latency (time flat in B up to
one wave) or by a shared SM resource (time grows with the CTAs per SM). usage: bwd_scaling.py [mhpc|hkd] B1 B2 ..."""
import copy, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
kind = sys.argv[1] if len(sys.argv) > 1 else "mhpc"
Bs = [int(a) for a in sys.argv[2:]] or [148, 296, 592, 1184, 4096]
csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")
if kind == "hkd":
    prob = cm.HKDProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/HKDMPC/settings/ddp_setting.info")); base = workload.hkd_batch(prob, 256)
else:
    prob = cm.MHPCProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info")); base = workload.mhpc_batch(256)
opt = copy.copy(opt); opt.max_AL_iter = 1; opt.max_DDP_iter = 3   # three ticks with every problem active
for B in Bs:
    x0 = np.tile(base, ((B + 255) // 256, 1))[:B]
    s = cm.MultiPhaseDDP(prob, 0, B); s.set_initial_condition(x0); s.solve(opt)
    s.set_profiling(True); s.solve(opt); tm = s.get_timing(); s.close()
    print(json.dumps({"B": B, "ctas_per_sm": round(B / 148, 2), **{k: round(tm["ms"][k] / max(tm["launches"][k], 1), 3) for k in ("lq", "misc", "bwd", "roll")}}), flush=True)
