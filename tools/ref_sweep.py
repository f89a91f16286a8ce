#!/usr/bin/env python
"""Large-sample check of the CPU oracle against the REFERENCE's own solver build (oracle/_ref/ref_hkd, ref_mhpc; needs /root/reference for the
binaries, so it runs in the build container): N perturbed problems of the HKD trot, the MHPC trot and the running barrel roll at offset 205, initial solve,
every counter and the final cost compared. One JSON line per workload (profiles/r02_ref_sweep.jsonl).

usage: python tools/ref_sweep.py [n_hkd] [n_mhpc] [n_barrel]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "tools"))
os.environ.setdefault("CAFE_HOST_ONLY", "1")
import cafe_mpc_b200 as cm  # noqa: E402
from cafe_mpc_b200 import workload as w  # noqa: E402
from make_ref_golden import parse  # noqa: E402
from oracle_bindings import oracle_solve  # noqa: E402

CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
RUN = os.path.join(REPO, "data/_run")


def run_ref(exe, args, rows):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        with open(fin, "w") as f:
            f.write("%d 0\n" % len(rows))
            for r in rows:
                f.write(" ".join(repr(float(v)) for v in r) + "\n")
        a = [os.path.join(REPO, "oracle/_ref", exe)] + [x.replace("@in", fin).replace("@out", fout) for x in args]
        t = time.time()
        subprocess.check_call(a, cwd=RUN, stdout=subprocess.DEVNULL)
        return parse(fout), time.time() - t


def compare(name, prob, opt, x0, probs, t_ref):
    mism, worst = 0, 0.0
    t = time.time()
    for b in range(len(x0)):
        info, _, _, _ = oracle_solve(prob.deck, opt, x0[b], cap=320)
        rec = probs[b][0]
        if [info["iter"], info["ls_iter_total"], info["reg_iter_total"]] != list(rec["counters"]):
            mism += 1
        worst = max(worst, abs(info["cost"] - rec["final"][0]) / abs(rec["final"][0]))
    print(json.dumps({"workload": name, "problems": len(x0), "counter_mismatches": mism, "worst_rel_final_cost": worst,
                      "iterations_min_max": [int(min(p[0]["counters"][0] for p in probs)), int(max(p[0]["counters"][0] for p in probs))],
                      "line_search_trials_max": int(max(p[0]["counters"][1] for p in probs)),
                      "reference_build_s": round(t_ref, 1), "oracle_s": round(time.time() - t, 1)}), flush=True)


def main():
    n_hkd, n_mhpc, n_barrel = [int(v) for v in (sys.argv[1:4] + ["64", "64", "16"][len(sys.argv) - 1:])]
    # HKD trot
    prob = cm.HKDProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
    body = np.tile(w.HKD_NOMINAL_BODY, (n_hkd, 1)); qJ = np.tile(w.HKD_NOMINAL_QJ, (n_hkd, 1))
    for b in range(1, n_hkd):
        for j in range(12):
            body[b, j] += w.HKD_BODY_SCALE[j] * (2 * w.uniform(b, j) - 1)
            qJ[b, j] += w.HKD_QJ_SCALE[j] * (2 * w.uniform(b, 12 + j) - 1)
    probs, t_ref = run_ref("ref_hkd", [CSV, "@in", "@out"], np.hstack([body, qJ]))
    compare("hkd_trot", prob, opt, w.hkd_batch(prob, n_hkd), probs, t_ref)
    # MHPC trot
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    x0 = w.mhpc_batch(n_mhpc)
    yaw = repr(float(prob.deck.contents.hip_yaw))
    probs, t_ref = run_ref("ref_mhpc", [CSV, yaw, "@in", "@out"], x0)
    compare("mhpc_trot", prob, opt, x0, probs, t_ref)
    # running barrel roll, landing (offset 205)
    k0 = w.BARREL_K0_IMPACT
    prob = cm.MHPCProblem(w.BARREL_CSV, mhpc_config=w.BARREL_CONFIG, k0=k0)
    x0 = w.barrel_batch(prob, n_barrel)
    src = open(w.BARREL_CSV).read().split("\n")
    with tempfile.TemporaryDirectory() as td:
        fcsv = os.path.join(td, "quad_reference.csv")
        open(fcsv, "w").write("\n".join(src[:2] + src[2 + 18 * k0:]))
        probs, t_ref = run_ref("ref_mhpc", [fcsv, yaw, "@in", "@out", "../MHPC/settings/mhpc_config_barrel.info"], x0)
    compare("running_barrel_roll_k0_205", prob, opt, x0, probs, t_ref)


if __name__ == "__main__":
    main()
