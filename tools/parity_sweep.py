"""Large-sample parity check (dev tool, run on the GPU box): N perturbed problems of a workload through the CUDA path and through the
CPU oracle (one single-threaded instance per host core), reporting how many problems differ in any counter (status, iterations,
line-search trials, regularisation steps, outer iterations, history length) and the worst relative deviation of the final cost and of
the packed solution. usage: parity_sweep.py [mhpc|hkd|barrel|loco|barrel_to] [N]"""
import json, multiprocessing as mp, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np

COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")


def make(kind):
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload
    csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    if kind == "hkd":
        prob = cm.HKDProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/HKDMPC/settings/ddp_setting.info")); gen = lambda n: workload.hkd_batch(prob, n)
    elif kind == "barrel":
        prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=workload.BARREL_K0_IMPACT)
        opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info")); gen = lambda n: workload.barrel_batch(prob, n)
    elif kind == "loco":
        prob = cm.LocoProblem(); opt = cm.load_hsddp_setting(workload.LOCO_DDP_SETTING); gen = workload.mhpc_batch
    elif kind == "barrel_to":
        prob = cm.BarrelRollProblem(); opt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING); gen = workload.mhpc_batch
    else:
        prob = cm.MHPCProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info")); gen = workload.mhpc_batch
    return cm, prob, opt, gen


def worker(args):
    kind, x0 = args
    from oracle_bindings import oracle_solve
    cm, prob, opt, gen = make(kind)
    out = []
    for x in x0:
        oi, oh, ot, osol = oracle_solve(prob.deck, opt, x, cap=320, guess=prob.initial_guess(x)[0] if kind == "barrel_to" else None)
        out.append(([oi[k] for k in COUNTS], oi["cost"], osol))
    return out


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "mhpc"
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    cm, prob, opt, gen = make(kind)
    x0 = gen(N)
    s = cm.MultiPhaseDDP(prob, 0, N); s.set_initial_condition(x0)
    if kind == "barrel_to": s.set_initial_guess(prob.initial_guess(x0))
    s.solve(opt)
    info = s.get_solver_info(); sol = s.get_solution()
    cores = len(os.sched_getaffinity(0))
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(worker, [(kind, x0[i::cores]) for i in range(cores)])
    mism, wc, ws = 0, 0.0, 0.0
    for c, chunk in enumerate(res):
        for q, (cnt, cost, osol) in enumerate(chunk):
            b = c + q * cores
            if [info[b][k] for k in COUNTS] != cnt: mism += 1
            wc = max(wc, abs(info[b]["cost"] - cost) / max(abs(cost), 1e-300))
            ws = max(ws, float(np.max(np.abs(sol[b] - osol)) / max(np.max(np.abs(osol)), 1e-300)))
    print(json.dumps({"workload": kind, "problems": N, "counter_mismatches": mism, "worst_cost_relerr": wc, "worst_solution_normrelerr": ws,
                      "mean_iter": sum(i["iter"] for i in info) / N, "oracle_s": round(time.perf_counter() - t, 1), "cores": cores}))
