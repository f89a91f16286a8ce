"""Generates tests/golden/*.npz from the CPU oracle (run in this container; commit the outputs)."""
import os, sys
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
from oracle_bindings import oracle_solve

prob = cm.HKDProblem(os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "HKDMPC/settings/ddp_setting.info"))
x0 = workload.hkd_batch(prob, 8)
info, hist, trace, sol = oracle_solve(prob.deck, opt, x0[0])
counts = np.array([info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")])
np.savez_compressed(os.path.join(R, "tests/golden/hkd_trot_nominal.npz"), x0=x0[0], counts=counts, hist=hist, trace=trace, sol=sol,
                    final=np.array([info["cost"], info["feas"], info["max_tconstr"], info["max_pconstr"]]))
# small perturbed batch: counters and final costs only
rows = []
for b in range(8):
    i, h, t, s = oracle_solve(prob.deck, opt, x0[b])
    rows.append([i[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] + [i["cost"], i["feas"]])
np.savez_compressed(os.path.join(R, "tests/golden/hkd_trot_batch8.npz"), x0=x0, rows=np.array(rows))
print(counts, info)

# ---- MHPC (whole-body + SRB) trot, t0 = 0 and the impact-bearing start offset k0 = 20
opt2 = cm.load_hsddp_setting(os.path.join(cm.api.DATA, "MHPC/settings/ddp_setting.info"))
csv = os.path.join(cm.api.DATA, "Reference/Data/trot/heuristic/quad_reference.csv")
x0m = workload.mhpc_batch(4)
out = {"x0": x0m}
for key, k0 in (("k0", 0), ("k20", 20)):
    pm = cm.MHPCProblem(csv, k0=k0)
    for b in (0, 3):
        i, h, t, s = oracle_solve(pm.deck, opt2, x0m[b])
        out["%s_counts_%d" % (key, b)] = np.array([i[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")])
        out["%s_hist_%d" % (key, b)] = h
        out["%s_sol_%d" % (key, b)] = s
        out["%s_final_%d" % (key, b)] = np.array([i["cost"], i["feas"], i["max_tconstr"], i["max_pconstr"]])
        print(key, b, i)
np.savez_compressed(os.path.join(R, "tests/golden/mhpc_trot.npz"), **out)

# ---- MHPC running barrel roll (BASELINE config 4): k0 = 0 and the mid-roll offset with the 4-foot landing impact
out = {}
for key, k0 in (("k0", 0), ("k205", workload.BARREL_K0_IMPACT)):
    pm = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=k0)
    x0b = workload.barrel_batch(pm, 4)
    out[key + "_x0"] = x0b
    out[key + "_phases"] = np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td] for p in pm.phases()])
    for b in (0, 3):
        i, h, t, s = oracle_solve(pm.deck, opt2, x0b[b])
        out["%s_counts_%d" % (key, b)] = np.array([i[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")])
        out["%s_hist_%d" % (key, b)] = h
        out["%s_sol_%d" % (key, b)] = s
        out["%s_final_%d" % (key, b)] = np.array([i["cost"], i["feas"], i["max_tconstr"], i["max_pconstr"]])
        print("barrel", key, b, i)
np.savez_compressed(os.path.join(R, "tests/golden/mhpc_barrel.npz"), **out)

# ---- LocoProblem (Locomotion/LocoProblem.cpp, Loco_TO.cpp): whole-body-only 1 s flypace plan, torque + GRF barriers only
pl = cm.LocoProblem()
optl = cm.load_hsddp_setting(workload.LOCO_DDP_SETTING)
x0l = workload.mhpc_batch(4)   # Loco_TO.cpp:49-55 is the nominal state of this table
out = {"x0": x0l, "phases": np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td, p.no_joint_limit, p.no_min_height] for p in pl.phases()])}
for b in (0, 3):
    i, h, t, s = oracle_solve(pl.deck, optl, x0l[b], cap=320)
    out["counts_%d" % b] = np.array([i[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")])
    out["hist_%d" % b] = h
    out["sol_%d" % b] = s
    out["final_%d" % b] = np.array([i["cost"], i["feas"], i["max_tconstr"], i["max_pconstr"]])
    print("loco", b, i)
np.savez_compressed(os.path.join(R, "tests/golden/loco_flypace.npz"), **out)

# ---- in-place barrel roll (BarrelRoll/BarrelRollTO.cpp): six hand-scheduled WB phases, joint-speed barrier, interpolated initial states.
# The full 30 x 10 caps take ~30 s per problem on the CPU (the solve runs into them); the golden uses 5 x 10.
pb = cm.BarrelRollProblem()
optb = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
optb.max_AL_iter = 5
x0b = workload.mhpc_batch(4)
gb = pb.initial_guess(x0b)
out = {"x0": x0b, "guess_xbar_0": np.concatenate([r["Xbar"][0] for r in __import__("cafe_mpc_b200.mpc", fromlist=["x"]).unpack_batch(pb, gb)]),
       "phases": np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td, p.joint_speed_limit] for p in pb.phases()])}
for b in (0, 3):
    i, h, t, s = oracle_solve(pb.deck, optb, x0b[b], cap=320, guess=gb[b])
    out["counts_%d" % b] = np.array([i[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")])
    out["hist_%d" % b] = h
    out["sol_%d" % b] = s
    out["final_%d" % b] = np.array([i["cost"], i["feas"], i["max_tconstr"], i["max_pconstr"]])
    print("barrel_to", b, i)
np.savez_compressed(os.path.join(R, "tests/golden/barrel_to.npz"), **out)
