#!/usr/bin/env python
"""The WHOLE headline batch through the reference's own solver build (oracle/_ref/ref_mhpc = MHPCProblem / WBM / HSDDPSolver compiled unchanged from
/root/reference against the stand-ins of oracle/refbuild; needs /root/reference for the binary, so it runs in the build container): all 4096 perturbed MHPC
trot problems of the SplitMix64 table (cafe_mpc_b200.workload.mhpc_batch), initial solve. Writes tests/golden/ref_mhpc_headline.npz = the reference's
counters (iterations, line-search trials, regularisation steps) and final cost of every problem, and checks the CPU oracle against all of them.

usage: python tools/ref_headline.py [n_problems] [n_processes]"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))
os.environ.setdefault("CAFE_HOST_ONLY", "1")
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")


def ref_slice(args):
    x0, yaw = args
    from ref_sweep import run_ref
    probs, t = run_ref("ref_mhpc", [CSV, yaw, "@in", "@out"], x0)
    return [list(p[0]["counters"]) + [p[0]["final"][0]] for p in probs]


def oracle_slice(x0):
    import cafe_mpc_b200 as cm
    from oracle_bindings import oracle_solve
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    out = []
    for x in x0:
        info, _, _, _ = oracle_solve(prob.deck, opt, x, cap=320)
        out.append([info["iter"], info["ls_iter_total"], info["reg_iter_total"], info["cost"]])
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    prob = cm.MHPCProblem(CSV)
    yaw = repr(float(prob.deck.contents.hip_yaw))
    x0 = w.mhpc_batch(n)
    cuts = np.array_split(np.arange(n), procs)
    t = time.time()
    with mp.get_context("spawn").Pool(procs) as pool:
        ref = np.array(sum(pool.map(ref_slice, [(x0[c], yaw) for c in cuts]), []))
        t_ref = time.time() - t
        t = time.time()
        ora = np.array(sum(pool.map(oracle_slice, [x0[c] for c in cuts]), []))
        t_ora = time.time() - t
    mism = int(np.sum(np.any(ref[:, :3].astype(np.int64) != ora[:, :3].astype(np.int64), axis=1)))
    worst = float(np.max(np.abs(ora[:, 3] - ref[:, 3]) / np.abs(ref[:, 3])))
    digest = hashlib.sha1(np.ascontiguousarray(x0).tobytes()).hexdigest()
    np.savez_compressed(os.path.join(REPO, "tests/golden/ref_mhpc_headline.npz"), counters=ref[:, :3].astype(np.int32), final_cost=ref[:, 3],
                        x0_sha1=np.array(digest), note=np.array("reference's own solver build (oracle/_ref/ref_mhpc), tools/ref_headline.py"))
    print(json.dumps({"workload": "mhpc_trot_headline_batch", "problems": n, "oracle_counter_mismatches": mism, "oracle_worst_rel_final_cost": worst,
                      "iterations_min_max": [int(ref[:, 0].min()), int(ref[:, 0].max())], "mean_iterations": float(ref[:, 0].mean()),
                      "line_search_trials_max": int(ref[:, 1].max()), "reference_wall_s": round(t_ref, 1), "oracle_wall_s": round(t_ora, 1), "processes": procs}))


if __name__ == "__main__":
    main()
