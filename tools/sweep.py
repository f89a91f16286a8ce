"""BASELINE config 5: sweep of horizon length x batch size for the HKD and the whole-body (MHPC) decks on the local GPU (dev tool;
prints one JSON line per point). Horizons: HKD plan_duration in {0.3, 0.6, 1.2} s; MHPC plan_dur_wb in {0.25, 0.5} s with
plan_dur_srb 0.5 s. With --gpus N the batch is cut over N GPUs of the box by cafe_gpu_create_multi (one process, one host thread per GPU; time = wall clock
around cafe_gpu_multi_solve_batch, i.e. incl. the H2D copy of the initial states).
usage: sweep.py [--batches 256,1024,4096,16384] [--quick] [--gpus N]"""
import argparse, json, os, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="256,1024,4096,16384")
ap.add_argument("--quick", action="store_true")
ap.add_argument("--gpus", type=int, default=1)
args = ap.parse_args()
batches = [int(b) for b in args.batches.split(",")]
csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")


def run(prob, opt, gen, B, tag):
    x0 = gen(min(B, 512))
    x0 = np.tile(x0, ((B + len(x0) - 1) // len(x0), 1))[:B]
    if args.gpus > 1:
        from cafe_mpc_b200 import api
        s = api.MultiGPUDDP(prob, args.gpus, B)
        s.solve(x0, opt)      # warm-up (page-in, clocks)
        ms = 1e30
        for _ in range(2):
            t0 = time.perf_counter(); s.solve(x0, opt); ms = min(ms, (time.perf_counter() - t0) * 1e3)
    else:
        s = cm.MultiPhaseDDP(prob, 0, B)
        s.set_initial_condition(x0)
        s.solve(opt)          # warm-up (page-in, clocks)
        s.solve(opt)
        ms = s.solve_ms()
    info = s.get_solver_info()
    knots = sum(p.horizon for p in prob.phases())
    print(json.dumps({"deck": tag, "knots": knots, "phases": [(p.model, p.horizon) for p in prob.phases()], "batch": B, "n_gpus": args.gpus, "solve_ms": round(ms, 2),
                      "solves_per_s": round(B / ms * 1e3, 1), "mean_iter": round(sum(i["iter"] for i in info) / B, 2),
                      "knot_iterations_per_s": round(sum(i["iter"] for i in info) * knots / ms * 1e3)}), flush=True)
    s.close()


opt_h = cm.load_hsddp_setting(os.path.join(R, "data/HKDMPC/settings/ddp_setting.info"))
opt_m = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info"))
for T in ((0.6,) if args.quick else (0.3, 0.6, 1.2)):
    prob = cm.HKDProblem(csv, plan_duration=T)
    for B in batches:
        run(prob, opt_h, lambda n: workload.hkd_batch(prob, n), B, "HKD trot plan_duration=%.1f" % T)
base = open(os.path.join(R, "data/MHPC/settings/mhpc_config.info")).read()
for Twb in ((0.25,) if args.quick else (0.25, 0.5)):
    with tempfile.NamedTemporaryFile("w", suffix=".info", delete=False) as f:
        f.write(base.replace("plan_dur_wb             0.25", "plan_dur_wb             %.2f" % Twb))
        cfg = f.name
    prob = cm.MHPCProblem(csv, mhpc_config=cfg)
    for B in batches:
        run(prob, opt_m, workload.mhpc_batch, B, "MHPC trot plan_dur_wb=%.2f srb=0.50" % Twb)
