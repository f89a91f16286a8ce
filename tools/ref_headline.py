#!/usr/bin/env python
"""The WHOLE headline batch through the reference's own solver build (oracle/_ref/ref_mhpc = MHPCProblem / WBM / HSDDPSolver compiled unchanged from
/root/reference against the stand-ins of oracle/refbuild; needs /root/reference for the binary, so it runs in the build container): all 4096 perturbed MHPC
trot problems of the SplitMix64 table (cafe_mpc_b200.workload.mhpc_batch), initial solve. Writes tests/golden/ref_mhpc_headline.npz = the reference's
counters (iterations, line-search trials, regularisation steps) and final cost of every problem, and checks the CPU oracle against all of them.

The same for the other two tables the GPU was swept on (`hkd`: HKD trot, `barrel`: running barrel roll landing at offset 205) - those write only the JSON line,
no fixture.

usage: python tools/ref_headline.py [n_problems] [n_processes] [mhpc|hkd|barrel]"""
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


def setup(kind):
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    if kind == "hkd":
        return cm.HKDProblem(CSV), cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    if kind == "barrel":
        return cm.MHPCProblem(w.BARREL_CSV, mhpc_config=w.BARREL_CONFIG, k0=w.BARREL_K0_IMPACT), opt
    return cm.MHPCProblem(CSV), opt


def ref_slice(args):
    rows, yaw, kind = args
    import tempfile
    from ref_sweep import run_ref
    if kind == "hkd":
        probs, t = run_ref("ref_hkd", [CSV, "@in", "@out"], rows)
    elif kind == "barrel":
        from cafe_mpc_b200 import workload as w
        src = open(w.BARREL_CSV).read().split("\n")
        with tempfile.TemporaryDirectory() as td:   # the reference reads the record file from its first row: trim the start offset off
            fcsv = os.path.join(td, "quad_reference.csv")
            open(fcsv, "w").write("\n".join(src[:2] + src[2 + 18 * w.BARREL_K0_IMPACT:]))
            probs, t = run_ref("ref_mhpc", [fcsv, yaw, "@in", "@out", "../MHPC/settings/mhpc_config_barrel.info"], rows)
    else:
        probs, t = run_ref("ref_mhpc", [CSV, yaw, "@in", "@out"], rows)
    return [list(p[0]["counters"]) + [p[0]["final"][0]] for p in probs]


def oracle_slice(args):
    x0, kind = args
    from oracle_bindings import oracle_solve
    prob, opt = setup(kind)
    out = []
    for x in x0:
        info, _, _, _ = oracle_solve(prob.deck, opt, x, cap=320)
        out.append([info["iter"], info["ls_iter_total"], info["reg_iter_total"], info["cost"]])
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
    kind = sys.argv[3] if len(sys.argv) > 3 else "mhpc"
    from cafe_mpc_b200 import workload as w
    prob, _ = setup(kind)
    yaw = repr(float(prob.deck.contents.hip_yaw))
    if kind == "hkd":    # the reference program takes body state and joint angles and runs compute_hkd_state itself (tools/ref_sweep.py)
        body = np.tile(w.HKD_NOMINAL_BODY, (n, 1)); qJ = np.tile(w.HKD_NOMINAL_QJ, (n, 1))
        for b in range(1, n):
            for j in range(12):
                body[b, j] += w.HKD_BODY_SCALE[j] * (2 * w.uniform(b, j) - 1)
                qJ[b, j] += w.HKD_QJ_SCALE[j] * (2 * w.uniform(b, 12 + j) - 1)
        rows = np.hstack([body, qJ]); x0 = w.hkd_batch(prob, n)
    elif kind == "barrel":
        x0 = w.barrel_batch(prob, n); rows = x0
    else:
        x0 = w.mhpc_batch(n); rows = x0
    cuts = np.array_split(np.arange(n), procs)
    t = time.time()
    with mp.get_context("spawn").Pool(procs) as pool:
        ref = np.array(sum(pool.map(ref_slice, [(rows[c], yaw, kind) for c in cuts]), []))
        t_ref = time.time() - t
        t = time.time()
        ora = np.array(sum(pool.map(oracle_slice, [(x0[c], kind) for c in cuts]), []))
        t_ora = time.time() - t
    mism = int(np.sum(np.any(ref[:, :3].astype(np.int64) != ora[:, :3].astype(np.int64), axis=1)))
    worst = float(np.max(np.abs(ora[:, 3] - ref[:, 3]) / np.abs(ref[:, 3])))
    digest = hashlib.sha1(np.ascontiguousarray(x0).tobytes()).hexdigest()
    if kind == "mhpc" and n == 4096:
        np.savez_compressed(os.path.join(REPO, "tests/golden/ref_mhpc_headline.npz"), counters=ref[:, :3].astype(np.int32), final_cost=ref[:, 3],
                            x0_sha1=np.array(digest), note=np.array("reference's own solver build (oracle/_ref/ref_mhpc), tools/ref_headline.py"))
    print(json.dumps({"workload": {"mhpc": "mhpc_trot_headline_batch", "hkd": "hkd_trot", "barrel": "running_barrel_roll_k0_205"}[kind], "problems": n, "oracle_counter_mismatches": mism, "oracle_worst_rel_final_cost": worst,
                      "iterations_min_max": [int(ref[:, 0].min()), int(ref[:, 0].max())], "mean_iterations": float(ref[:, 0].mean()),
                      "line_search_trials_max": int(ref[:, 1].max()), "reference_wall_s": round(t_ref, 1), "oracle_wall_s": round(t_ora, 1), "processes": procs}))


if __name__ == "__main__":
    main()
