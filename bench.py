#!/usr/bin/env python3
"""bench.py — batched HS-DDP solves/sec on B200.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference ...                    the CPU restatement of the reference (oracle) on all host cores

A "step" = one complete solve of a batch of perturbed-initial-state problems (every problem to its own
termination). value = solves/s with x0 resident in HBM; e2e = same through the host-buffer C-ABI calls
(H2D of x0, solve, D2H of the command records; for N>1 also the final NCCL gather to rank 0)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

METRIC = "batched HS-DDP solves/sec"
UNIT = "solves/s"
F_BWD = {0: 233280.0, 1: 373392.0, 2: 30096.0}  # dense flop per knot of one backward-sweep pass
F_LIN = {0: 6960.0, 1: 8200.0, 2: 1800.0}        # linear rollout per knot
# algorithmic bytes of one LQ approximation per knot: inputs X, U, Y, Defect + the reference's outputs
# A, B, C, D, lx, lu, ly, lxx, luu, lyy (model ids 0 HKD (24,24,0), 1 WB (36,12,12), 2 SRB (12,12,0)), 8 bytes per double
def _lq_doubles(n, m, p):
    return (2 * n + m + p) + (n * n + n * m + p * n + p * m + n + m + p + n * n + m * m + p * p)
B_LQ = {0: 8.0 * _lq_doubles(24, 24, 0), 1: 8.0 * _lq_doubles(36, 12, 12), 2: 8.0 * _lq_doubles(12, 12, 0)}


def workload_name(args):
    if args.workload == "hkd":
        return "HKD trot (3 phases h=11/25/24, n=m=24), %d perturbed problems per GPU" % args.batch
    if args.workload == "barrel":
        return ("MHPC running barrel roll at the impact-bearing start offset k0=205 (WB flight h=22 -> 4-foot landing impact -> WB h=3; SRB h=10), "
                "%d perturbed problems per GPU" % args.batch)
    if args.workload == "barrel_to":
        return ("in-place barrel roll (BarrelRoll/BarrelRollTO.cpp): 6 WB phases / 125 knots, joint-speed barrier, two 4-foot landings, solve started from "
                "the interpolated state trajectory, %d perturbed problems per GPU" % args.batch)
    if args.workload == "loco":
        return ("LocoProblem (Locomotion/Loco_TO.cpp): whole-body-only 1.0 s flypace plan, 9 WB phases / 100 knots, three flight -> stance "
                "touchdowns, torque + GRF barriers, %d perturbed problems per GPU" % args.batch)
    return "MHPC trot (WB h=11 + WB h=14, n=36 m=12 p=12; SRB h=10, n=m=12), %d perturbed problems per GPU" % args.batch


def make_problem(workload):
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as wl
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    if workload == "hkd":
        prob = cm.HKDProblem(csv)
        opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
        return prob, opt, (lambda B: wl.hkd_batch(prob, B)), 24
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    if workload == "barrel":
        prob = cm.MHPCProblem(wl.BARREL_CSV, mhpc_config=wl.BARREL_CONFIG, k0=wl.BARREL_K0_IMPACT)
        return prob, opt, (lambda B: wl.barrel_batch(prob, B)), 36
    if workload == "barrel_to":
        prob = cm.BarrelRollProblem()
        return prob, cm.load_hsddp_setting(wl.BARREL_TO_DDP_SETTING), (lambda B: wl.mhpc_batch(B)), 36
    if workload == "loco":
        prob = cm.LocoProblem()
        return prob, cm.load_hsddp_setting(wl.LOCO_DDP_SETTING), (lambda B: wl.mhpc_batch(B)), 36
    prob = cm.MHPCProblem(csv)
    return prob, opt, (lambda B: wl.mhpc_batch(B)), 36


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def _cpu_worker(task):
    from oracle_bindings import oracle_solve
    x0s, workload = task
    prob, opt, _, _ = make_problem(workload)
    t = time.perf_counter()
    for x in x0s:
        oracle_solve(prob.deck, opt, x, cap=320, guess=prob.initial_guess(x)[0] if workload == "barrel_to" else None)
    return time.perf_counter() - t


def cpu_sample(x0, cores, per_core, workload):
    """One single-threaded oracle instance per host core over disjoint slices."""
    import multiprocessing as mp
    n = min(len(x0), cores * per_core)
    chunks = [x0[i:n:cores] for i in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(c[:1], workload) for c in chunks])  # warm-up: imports, page-in
        t = time.perf_counter()
        pool.map(_cpu_worker, [(c, workload) for c in chunks])
        wall = time.perf_counter() - t
    return n / wall, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    prob, opt, gen_x0, n0 = make_problem(args.workload)
    per_core = args.cpu_per_core
    x0 = gen_x0(min(args.batch, cores * per_core))
    vals = []
    for i in range(args.warmup + args.steps):
        v, n = cpu_sample(x0, cores, per_core, args.workload)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    sample = "%d problems of the workload (first of the SplitMix64 table), %d per core, one single-threaded oracle instance per core" % (len(x0), per_core)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * len(x0) / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(args), "note": "CPU restatement of the reference solver (oracle/), reference CasADi C linked from oracle/_ref; the reference itself needs Eigen/Boost/Pinocchio/LCM and cannot be built in this image"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="problems per GPU")
    ap.add_argument("--workload", default="mhpc", choices=["mhpc", "hkd", "barrel", "loco", "barrel_to"])
    ap.add_argument("--gain-knots", type=int, default=8)
    ap.add_argument("--cpu-per-core", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import distributed as cdist
    from cafe_mpc_b200 import workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    Bg = B * world
    prob, opt, gen_x0, n0 = make_problem(args.workload)
    x0_all = gen_x0(Bg) if Bg <= 8192 else np.tile(gen_x0(8192), ((Bg + 8191) // 8192, 1))[:Bg]
    lo, hi = cdist.shard_range(Bg, world, rank)
    x0 = np.ascontiguousarray(x0_all[lo:hi])
    solver = cm.MultiPhaseDDP(prob, local, B)
    if args.workload == "barrel_to":
        solver.set_initial_guess(prob.initial_guess(x0))   # BarrelRollTO.cpp:131-147: part of the problem set-up, stays in force
    # device-resident inputs: x0 as [n0][ldb]
    x0_dev = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
    x0_pin = torch.from_numpy(x0).pin_memory()
    rec = solver.command_size(args.gain_knots)
    cmd_dev = torch.empty((B, rec), dtype=torch.float64, device="cuda")
    cmd_pin = torch.empty((B, rec), dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        solver.solve_device(x0_dev.data_ptr(), B, B, opt)
        return solver.solve_ms()

    def step_e2e():
        solver.set_initial_condition(x0_pin.numpy())
        solver.solve(opt)                                            # H2D of x0 inside
        if world > 1:
            solver.get_commands_device(args.gain_knots, cmd_dev.data_ptr())
            out = cdist.gather_records(cmd_dev, Bg, world, rank)     # the one collective: NCCL gather to rank 0
            if rank == 0:
                cmd_pin.copy_(out[:B], non_blocking=False)           # rank 0 reads the result on the host
        else:
            solver.get_commands(args.gain_knots, out=cmd_pin.numpy())  # D2H of the command records
        return float(cmd_pin[0, 0])

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        dev_ms += step_resident()
    barrier()
    wall = time.perf_counter() - t0
    launches = sum(solver.get_timing()["launches"].values()) * args.steps
    info = solver.get_solver_info()
    for _ in range(1):
        step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    wall_e2e = time.perf_counter() - t1
    sampler.stop.set()
    sampler.join(timeout=2)
    tt = torch.tensor([wall, wall_e2e, dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_ms = [float(v) for v in tt.cpu()]

    # ---- roofline of the dominant kernel (k_bwd), measured live with per-launch CUDA events on its stream
    roof = None
    cpu = None
    if rank == 0:
        solver.set_profiling(True)
        step_resident()
        tm = solver.get_timing()
        solver.set_profiling(False)
        pinfo = solver.get_solver_info()
        phases = prob.phases()
        f_sweep = sum(F_BWD[p.model] * p.horizon for p in phases)
        f_lin = sum(F_LIN[p.model] * p.horizon for p in phases)
        flops = sum(i["reg_iter_total"] * f_sweep + i["iter"] * f_lin for i in pinfo)
        n_l = max(tm["launches"]["bwd"], 1)
        peak = cm.measure_fp64_peak(local)
        ach = flops / (tm["ms"]["bwd"] * 1e-3) / 1e12
        share = tm["ms"]["bwd"] / max(sum(tm["ms"].values()), 1e-9)
        traffic, traffic_src = _ncu_traffic("k_bwd2") if (args.workload == "mhpc" and B == 4096) else (None, None)
        roof_bwd = {"bound": "tensor", "pipe": "fp64 tensor pipe (DMMA m8n8k4) + fp64 FMA pipe", "kernel": "k_bwd2", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "flop_per_launch": flops / n_l, "avg_launch_ms": tm["ms"]["bwd"] / n_l, "share_of_step": share,
                "peak_source": "measured live: cafe_gpu_measure_fp64_peak = max(DFMA chains, DMMA m8n8k4 chains) microbenchmark; MEASURED_PEAKS.json has no fp64 entry"}
        # the LQ stage (k_lq + its cooperative whole-body part k_lq_wb_dense, timed in the "misc" slot): HBM-bound by design, every
        # active problem reads its iterate and writes the linearisation once per DDP iteration
        lq_ms = tm["ms"]["lq"] + tm["ms"]["misc"]
        n_lq = max(tm["launches"]["lq"], 1)
        b_pass = sum(B_LQ[p.model] * p.horizon for p in phases)
        lq_bytes = sum(i["iter"] * b_pass for i in pinfo)
        t_lq, _ = _ncu_traffic("k_lq") if (args.workload == "mhpc" and B == 4096) else (None, None)
        t_ds, _ = _ncu_traffic("k_lq_wb_dense") if (args.workload == "mhpc" and B == 4096) else (None, None)
        n_wb = sum(1 for p in phases if p.model == 1)
        roof_lq = {"bound": "hbm", "kernel": "k_lq (+ k_lq_wb_dense)", "achieved": lq_bytes / (lq_ms * 1e-3) / 1e9, "peak": _hbm_peak(), "unit": "GB/s",
                   "frac": lq_bytes / (lq_ms * 1e-3) / 1e9 / _hbm_peak(), "traffic": (t_lq + n_wb * t_ds) if (t_lq and t_ds) else t_lq,
                   "traffic_source": traffic_src, "bytes_per_launch": lq_bytes / n_lq, "avg_launch_ms": lq_ms / n_lq,
                   "share_of_step": lq_ms / max(sum(tm["ms"].values()), 1e-9),
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)",
                   "note": "algorithmic bytes = the reference's dense LQ inputs and outputs per knot (A,B,C,D, cost partials); whole-body decks: DRAM traffic "
                           "far above it = register spills and thread-local arrays of the generated routines; HKD decks: structural zeros of A, B, lxx, "
                           "luu are never rewritten, so fewer bytes move than this count"}
        # `roofline` is the stage with the larger share of the step; the other one rides along
        roof = dict(roof_lq if lq_ms > tm["ms"]["bwd"] else roof_bwd)
        roof["other"] = roof_bwd if lq_ms > tm["ms"]["bwd"] else roof_lq
        roof["kernel_ms"] = tm["ms"]; roof["hbm_peak_gbs"] = _hbm_peak()
        # secondary figures per kernel family: share of the step and, where a committed ncu capture exists, the HBM fraction
        # (DRAM bytes of one full-batch launch x launches / live kernel time; later ticks run fewer active problems, so this is an upper bound)
        per = {}
        tot_ms = max(sum(tm["ms"].values()), 1e-9)
        for fam, kname in (("roll", "k_roll"), ("lq", "k_lq"), ("bwd", "k_bwd2"), ("misc", None), ("select", None), ("accept", None)):
            e = {"ms": tm["ms"][fam], "launches": tm["launches"][fam], "share_of_step": tm["ms"][fam] / tot_ms}
            tb, _ = _ncu_traffic(kname) if (kname and args.workload == "mhpc" and B == 4096) else (None, None)
            if tb and tm["ms"][fam] > 0:
                gbs = tb * tm["launches"][fam] / (tm["ms"][fam] * 1e-3) / 1e9
                e["hbm_gbs_upper_bound"] = gbs; e["hbm_frac_upper_bound"] = gbs / _hbm_peak()
            per[fam] = e
        roof["per_kernel"] = per
        if not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            v, n = cpu_sample(x0_all[: cores * args.cpu_per_core], cores, args.cpu_per_core, args.workload)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d problems (first of the same SplitMix64 table), one single-threaded oracle instance per core" % n}
    if rank == 0:
        it = [i["iter"] for i in info]
        print(json.dumps({
            "metric": METRIC, "value": Bg * args.steps / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * wall / args.steps, "device_ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": Bg, "per_gpu_batch": B, "parallelism": "batch sharded over %d GPU(s), no collective in the solve" % world,
                       "settings": {"hkd": "HKDMPC/settings (10x5 iteration caps, alpha 0.1)",
                                    "mhpc": "MHPC/settings (10x20 iteration caps, alpha 0.5, BG_alpha 10, cost_weights_regular, constraint_params_regular)",
                                    "barrel": "MHPC/settings (10x20 iteration caps, alpha 0.5, BG_alpha 10, cost_weights_barrel, constraint_params_barrel)",
                                    "barrel_to": "BarrelRoll/setting (30x10 iteration caps, alpha 0.5, BG_alpha 10, br_cost_weights, br_constraint_params)",
                                    "loco": "Locomotion/settings (30x10 iteration caps, alpha 0.5, BG_alpha 10, loco_cost_weights, loco_constraint_params)"}[args.workload],
                       "l2": "working set per solve (GBs of per-problem arrays) exceeds the 126 MB L2; no flush needed",
                       "mean_ddp_iterations": sum(it) / len(it), "max_ddp_iterations": max(it)},
            "e2e": {"value": Bg * args.steps / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * n0 * 8), "d2h_bytes_per_step": int(B * rec * 8),
                    "what": "cafe_gpu_solve_batch(host x0) + cafe_gpu_get_commands (Xbar,Ubar,Y all knots; K,Qu,Quu,Qux first %d knots)%s" % (args.gain_knots, " + NCCL gather to rank 0" if world > 1 else "")},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full summary of the newest round (profiles/rNN_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(REPO, "profiles", "r*_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    k = d.get("kernels", {}).get(kernel)
    return (k["dram_bytes_per_launch"], os.path.relpath(files[-1], REPO) + ": " + d.get("source", "")) if k else (None, None)


def _hbm_peak():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()
