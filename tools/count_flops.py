#!/usr/bin/env python
"""MEASUREMENT TOOL: algorithmic flops of every stage, COUNTED by the CPU oracle built with an instrumented
scalar type (tools/flopcount/counted.hpp: `double` re-read as a tallying one-double struct, oracle sources compiled unchanged).

Runs the first DDP iteration of the nominal problem of a deck stage by stage and phase by phase (tools/flopcount/count_entry.cpp) and prints, per
model and stage, the operations per knot: additions, multiplications, divisions, square roots, transcendental calls executed by the oracle's own
C++ plus the reference's CasADi functions it called, priced with the number of arithmetic statements of the generated C (counted from the sources
under /root/reference when present, else from tools/flopcount/casadi_ops.json). 1 flop = one add, mul, div, sqrt or transcendental call.

    python tools/count_flops.py [mhpc|hkd|barrel|loco] [--json profiles/r02_counted_flops.json]

Note what is counted: the ORACLE's algorithm, e.g. its dense O(n^3) products that skip exact zeros and its forward-mode (37-direction dual number)
RNEA derivatives, where the reference calls Pinocchio's analytic computeRNEADerivatives and the GPU kernels use structured symbolic derivatives.
The sweep's figure is the check on SURVEY's dense formula F_bwd; the model-stage figures replace SURVEY's estimates.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
FC = os.path.join(REPO, "tools", "flopcount")
OPS_JSON = os.path.join(FC, "casadi_ops.json")
STAGES = ["roll", "cost", "lq", "bwd", "lin"]
KINDS = ["add", "mul", "div", "sqrt", "trans", "cmp"]
MODEL_NAMES = {0: "HKD", 1: "WB", 2: "SRB"}


def build():
    out = os.path.join(FC, "build")
    os.makedirs(out, exist_ok=True)
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "ref"], check=True, capture_output=True)
    objs = []
    for src in ["oracle/hsddp_oracle.cpp", "oracle/hkd_oracle.cpp", "oracle/srb_oracle.cpp", "oracle/wb_oracle.cpp", "tools/flopcount/count_entry.cpp"]:
        o = os.path.join(out, os.path.basename(src)[:-4] + ".o")
        s = os.path.join(REPO, src)
        deps = [s, os.path.join(FC, "counted.hpp")] + [os.path.join(REPO, "oracle", h) for h in ("hsddp_oracle.hpp", "wb_dynamics.hpp", "casadi_ref.hpp")]
        if not os.path.exists(o) or any(os.path.getmtime(d) > os.path.getmtime(o) for d in deps):
            subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-w", "-include", os.path.join(FC, "counted.hpp"), "-c", s, "-o", o], check=True)
        objs.append(o)
    so = os.path.join(out, "libcafe_oracle_counted.so")
    subprocess.run(["g++", "-shared", "-o", so] + objs + ["-L" + os.path.join(REPO, "oracle", "_ref"), "-lcafe_ref_casadi", "-ldl",
                                                         "-Wl,-rpath," + os.path.join(REPO, "oracle", "_ref")], check=True)
    return so


def casadi_statement_counts():
    """arithmetic statements (one operation each: CasADi's generated C is three-address code) per exported function"""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        return json.load(open(OPS_JSON))
    files = [os.path.join(ref, "HKDMPC/HKD-TrajOpt/CasadiGen/source", f) for f in sorted(os.listdir(os.path.join(ref, "HKDMPC/HKD-TrajOpt/CasadiGen/source")))]
    files += [os.path.join(ref, "MHPC/MHPC-Trajopt/CasadiGen/source", f) for f in ("SRBDynamics.cpp", "MCKinematicsDerivativs.cpp")]
    counts = {}
    op = re.compile(r"^\s*(?:a\d+|w\d+)\s*=\s*(.*);\s*$")
    plain = re.compile(r"^(?:arg\[\d+\]\s*\?.*|[-+]?[\d.eE+-]+|a\d+|w\d+)$")
    for f in files:
        body, static_ops, cur = {}, {}, None
        for line in open(f, errors="replace"):
            m = re.match(r"^static int (casadi_f\d+)\(", line)
            if m:
                cur = m.group(1); static_ops[cur] = 0; continue
            m = re.search(r"CASADI_SYMBOL_EXPORT int (\w+)\(const casadi_real\*\* arg", line)
            if m:
                cur = "export:" + m.group(1); continue
            if cur is None:
                continue
            if cur.startswith("export:"):
                m = re.search(r"return (casadi_f\d+)\(", line)
                if m:
                    body[cur[7:]] = m.group(1); cur = None
                continue
            m = op.match(line)
            if m and not plain.match(m.group(1).strip()) and not m.group(1).strip().startswith("(-a") :
                static_ops[cur] += 1
            elif m and m.group(1).strip().startswith("(-a"):
                pass  # sign flip: free, like in counted.hpp
        for name, fn in body.items():
            counts[name] = static_ops[fn]
    json.dump(counts, open(OPS_JSON, "w"), indent=1, sort_keys=True)
    return counts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="mhpc")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    import bench
    from cafe_mpc_b200._ctypes_defs import Deck, Options
    prob, opt, batch, n0 = bench.make_problem(a.workload)
    x0 = np.ascontiguousarray(np.asarray(batch(1))[0], dtype=np.float64)   # problem 0 = the nominal problem
    assert x0.size == n0
    lib = C.CDLL(build())
    assert lib.cafe_count_set_casadi_lib(os.path.join(REPO, "oracle", "_ref", "libcafe_ref_casadi.so").encode()) == 0
    lib.cafe_count_casadi_name.restype = C.c_char_p
    ns, nk, nc = lib.cafe_count_n_stage(), lib.cafe_count_n_kind(), lib.cafe_count_n_casadi()
    names = [lib.cafe_count_casadi_name(i).decode() for i in range(nc)]
    deck = prob.deck
    np_ = deck.contents.n_phases
    ops = np.zeros((np_ + 1, ns, nk), dtype=np.int64)
    calls = np.zeros((np_ + 1, ns, nc), dtype=np.int64)
    lib.cafe_count_first_iteration.argtypes = [C.POINTER(Deck), C.POINTER(Options), C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.cafe_count_first_iteration(deck, C.byref(opt), x0.ctypes.data, ops.ctypes.data, calls.ctypes.data)
    assert rc == np_, rc
    cas = casadi_statement_counts()
    out = {"workload": a.workload, "what": "first DDP iteration of the unperturbed problem, regularisation 0, one trial; operations per knot of a phase = phase total / horizon",
           "flop": "add + mul + div + sqrt + transcendental calls of the oracle's C++ (instrumented scalar) + arithmetic statements of the CasADi functions it called",
           "casadi_statements": cas, "phases": []}
    print(f"{'phase':>5} {'model':>5} {'h':>3} {'stage':>5} {'flop/knot':>11} {'own C++':>10} {'CasADi':>9}   add / mul / div / sqrt / trans per knot; CasADi calls per knot")
    for i in range(np_):
        ph = deck.contents.phase[i]
        h = ph.horizon
        model = MODEL_NAMES.get(ph.model, str(ph.model))
        row = {"phase": i, "model": model, "horizon": h, "contacts": [int(ph.contact[l]) for l in range(4)], "stages": {}}
        for s in range(ns):
            own = int(ops[i, s, :5].sum())
            cc = {names[c]: int(calls[i, s, c]) for c in range(nc) if calls[i, s, c]}
            cflop = sum(n * cas.get(k, 0) for k, n in cc.items())
            row["stages"][STAGES[s]] = {"flop_per_knot": (own + cflop) / h, "own_per_knot": own / h, "casadi_per_knot": cflop / h,
                                        "ops": {KINDS[k]: int(ops[i, s, k]) for k in range(nk)}, "casadi_calls": cc}
            print(f"{i:5d} {model:>5} {h:3d} {STAGES[s]:>5} {(own + cflop) / h:11.0f} {own / h:10.0f} {cflop / h:9.0f}   "
                  + " / ".join(f"{ops[i, s, k] / h:.0f}" for k in range(5)) + "; " + ", ".join(f"{k} {n / h:.2f}" for k, n in cc.items()))
        out["phases"].append(row)
    # one whole solve of the same problem
    sops = np.zeros(nk, dtype=np.int64); scalls = np.zeros(nc, dtype=np.int64); cnt = (C.c_int * 3)()
    lib.cafe_count_solve.argtypes = [C.POINTER(Deck), C.POINTER(Options), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.cafe_count_solve(deck, C.byref(opt), x0.ctypes.data, sops.ctypes.data, scalls.ctypes.data, cnt) == 0
    own = int(sops[:5].sum()); cflop = int(sum(int(scalls[c]) * cas.get(names[c], 0) for c in range(nc)))
    out["solve"] = {"flop": own + cflop, "own": own, "casadi": cflop, "iterations": cnt[0], "line_search_trials": cnt[1], "regularisation_steps": cnt[2],
                    "ops": {KINDS[k]: int(sops[k]) for k in range(nk)}}
    print("whole solve of the nominal problem: %.1f MFLOP (%.1f own C++ + %.1f CasADi), %d iterations, %d line-search trials, %d sweeps"
          % ((own + cflop) / 1e6, own / 1e6, cflop / 1e6, cnt[0], cnt[1], cnt[2]))
    if a.json:
        json.dump(out, open(os.path.join(REPO, a.json) if not os.path.isabs(a.json) else a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
