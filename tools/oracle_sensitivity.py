"""How well-posed is "iteration counts bit-exact" for a workload? (dev tool, CPU only)  Runs N perturbed problems through the CPU oracle twice:
the regular build (-O3 -march=x86-64-v3: fused multiply-adds, AVX2 reductions) and a build of the SAME sources without FMA contraction
(-O2 -ffp-contract=off -march=x86-64), i.e. two legal roundings of one and the same algorithm, and counts the problems whose counters
(status, iterations, line-search trials, regularisation steps, outer iterations, history length) differ.
usage: oracle_sensitivity.py [mhpc|hkd|barrel|loco|barrel_to] [N] [variant.so] [dump.json]
dump.json receives, per problem, the counters and the final cost of both builds (tests/golden/barrel_to_four_roundings.json merges two such dumps: the kinematic-partial file of oracle/_ref compiled -O1 and -O3)."""
import ctypes as C, json, multiprocessing as mp, os, subprocess, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np

COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")


def worker(args):
    kind, x0, variant = args
    import oracle_bindings
    if variant:
        oracle_bindings._lib = None      # a pool process may have loaded the regular build already
        oracle_bindings.ORACLE_LIB = variant
        oracle_bindings.build_oracle = lambda: variant
    from parity_sweep import make
    cm, prob, opt, gen = make(kind)
    out = []
    for x in x0:
        oi, oh, ot, osol = oracle_bindings.oracle_solve(prob.deck, opt, x, cap=320, guess=prob.initial_guess(x)[0] if kind == "barrel_to" else None)
        out.append(([oi[k] for k in COUNTS], oi["cost"], osol))
    return out


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "barrel_to"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    variant = sys.argv[3] if len(sys.argv) > 3 else "/tmp/ovar/libcafe_oracle_nofma.so"
    if not os.path.exists(variant):
        os.makedirs(os.path.dirname(variant), exist_ok=True)
        src = ["hsddp_oracle.cpp", "hkd_oracle.cpp", "srb_oracle.cpp", "wb_oracle.cpp", "casadi_eval.cpp"]
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-march=x86-64", "-std=c++17", "-fPIC", "-shared", "-o", variant] + src +
                       ["-L_ref", "-lcafe_ref_casadi", "-Wl,-rpath," + os.path.join(R, "oracle/_ref")], cwd=os.path.join(R, "oracle"), check=True)
    from parity_sweep import make
    cm, prob, opt, gen = make(kind)
    x0 = gen(N)
    cores = os.cpu_count()
    chunks = [x0[i::cores] for i in range(cores)]
    t0 = time.time()
    with mp.Pool(cores) as pool:
        a = pool.map(worker, [(kind, c, None) for c in chunks])
    with mp.Pool(cores) as pool:
        b = pool.map(worker, [(kind, c, variant) for c in chunks])
    def in_order(chunked):   # chunk c holds problems c, c + cores, c + 2 cores, ...
        out = [None] * N
        for c, ch in enumerate(chunked):
            for j, r in enumerate(ch):
                out[c + cores * j] = r
        return out
    a = in_order(a); b = in_order(b)
    mism = sum(1 for ra, rb in zip(a, b) if ra[0] != rb[0])
    cost = max(abs(ra[1] - rb[1]) / max(abs(ra[1]), 1e-300) for ra, rb in zip(a, b))
    sol = max(np.abs(ra[2] - rb[2]).max() / np.abs(ra[2]).max() for ra, rb in zip(a, b))
    if len(sys.argv) > 4:
        json.dump({"workload": kind, "counters": list(COUNTS), "x0": "cafe_mpc_b200.workload.mhpc_batch(%d)" % N if kind != "hkd" else "hkd_batch", "fma": [[r[0], r[1]] for r in a], "no_fma": [[r[0], r[1]] for r in b]},
                  open(sys.argv[4], "w"), indent=0)
    print(json.dumps({"workload": kind, "problems": N, "what": "CPU oracle (-O3, FMA) vs the same sources built without FMA contraction", "counter_mismatches": mism,
                      "worst_cost_relerr": cost, "worst_solution_normrelerr": sol, "mean_iter": float(np.mean([r[0][1] for r in a])), "seconds": round(time.time() - t0, 1), "cores": cores}))
