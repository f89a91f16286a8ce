#!/usr/bin/env python3
"""bench.py — batched HS-DDP solves/sec on B200.

  python bench.py --gpus N --steps K --warmup W          our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference ...                    the CPU restatement of the reference (oracle) on all host cores

A "step" = one complete solve of a batch of perturbed-initial-state problems (every problem to its own termination).
value = solves/s with x0 resident in HBM; e2e = the same through the host-buffer C-ABI calls (H2D of x0, solve, pack of the command
records, for N > 1 the NCCL gather to rank 0 through cafe_gpu_gather_commands, D2H of ALL records on rank 0).

Scaling: --scaling strong (default) keeps the GLOBAL batch at --batch (the headline: 4096 MHPC trot problems cut over N GPUs);
--scaling weak gives every GPU --batch problems. torch.distributed is plumbing only (rendezvous, barriers, max over ranks, handing the
NCCL id around); the solve and the gather go through the C ABI."""
import argparse
import glob
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
# operation counts of the generated straight-line routines (tools/gen_wb.py, tools/gen_wb_leg.py print them; one op = one flop)
OPS_TERMS = 197 + 4 * 1886
OPS_DERIVS = 851 + 4 * (1834 + 2453 + 2014 + 2691 + 1909 + 1502 + 1781 + 866 + 850)
TM_NNZ, DP_NNZ = 13 + 4 * 79, 18 + 4 * 198       # doubles a knot's rigid-body terms / derivative pieces occupy


def _chol_flop(n):
    return sum((n - j) * (2 * j + 1) for j in range(n))


def wb_fwd_flop(nr):
    """k_wb_fwd per (problem, knot, step size): chol(M), Y = L^-1 Jc^T (+ rhs), S = Y^T Y, chol(S), lambda, qdd; x+, defects, cost"""
    return _chol_flop(18) + (nr + 1) * 324 + nr * (nr + 1) // 2 * 36 + _chol_flop(nr) + 36 * nr + 2 * nr * nr + 36 * nr + 324 + 800


def wb_sens_flop(nr):
    """k_wb_sens per (problem, knot): factorisation + 48 columns of L^-1, Y^T, Ls^-T Ls^-1, Y, L^-T + assembly of R, a"""
    return _chol_flop(18) + nr * 324 + nr * (nr + 1) // 2 * 36 + _chol_flop(nr) + 48 * (2 * 324 + 4 * 18 * nr + 2 * nr * nr) + 2000


def workload_name(args):
    if args.workload == "hkd":
        return "HKD trot (3 phases h=11/25/24, n=m=24)"
    if args.workload == "barrel":
        return "MHPC running barrel roll at the impact-bearing start offset k0=205 (WB flight h=22 -> 4-foot landing impact -> WB h=3; SRB h=10)"
    if args.workload == "barrel_to":
        return ("in-place barrel roll (BarrelRoll/BarrelRollTO.cpp): 6 WB phases / 125 knots, joint-speed barrier, two 4-foot landings, solve started from "
                "the interpolated state trajectory")
    if args.workload == "loco":
        return "LocoProblem (Locomotion/Loco_TO.cpp): whole-body-only 1.0 s flypace plan, 9 WB phases / 100 knots, three flight -> stance touchdowns, torque + GRF barriers"
    return "MHPC trot (WB h=11 + WB h=14, n=36 m=12 p=12; SRB h=10, n=m=12)"


SETTINGS = {"hkd": "HKDMPC/settings (10x5 iteration caps, alpha 0.1)",
            "mhpc": "MHPC/settings (10x20 iteration caps, alpha 0.5, BG_alpha 10, cost_weights_regular, constraint_params_regular)",
            "barrel": "MHPC/settings (10x20 iteration caps, alpha 0.5, BG_alpha 10, cost_weights_barrel, constraint_params_barrel)",
            "barrel_to": "BarrelRoll/setting (30x10 iteration caps, alpha 0.5, BG_alpha 10, br_cost_weights, br_constraint_params)",
            "loco": "Locomotion/settings (30x10 iteration caps, alpha 0.5, BG_alpha 10, loco_cost_weights, loco_constraint_params)"}


def make_config(args, world):
    """the same dict in both arms (ours and --impl reference): the workload the metric is quoted on"""
    Bg = args.batch if args.scaling == "strong" else args.batch * world
    return {"workload": "%s, %d perturbed problems (SplitMix64 table, SURVEY.md section 8d)" % (workload_name(args), Bg),
            "global_batch": Bg, "scaling": args.scaling, "settings": SETTINGS[args.workload],
            "l2": "working set per solve (GBs of per-problem arrays) exceeds the 126 MB L2; no flush needed"}


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


_CPU_PROBLEM = {}


def _cpu_worker(task):
    from oracle_bindings import oracle_solve
    x0s, workload = task
    if workload not in _CPU_PROBLEM:   # one deck per worker process
        _CPU_PROBLEM[workload] = make_problem(workload)
    prob, opt, _, _ = _CPU_PROBLEM[workload]
    t = time.perf_counter()
    for x in x0s:
        oracle_solve(prob.deck, opt, x, cap=320, guess=prob.initial_guess(x)[0] if workload == "barrel_to" else None)
    return time.perf_counter() - t


def cpu_sample(x0, cores, per_core, workload, budget_s=0.0):
    """One single-threaded oracle instance per host core over disjoint slices. per_core > 0: that many problems per core;
    per_core <= 0: as many of x0's problems as `budget_s` seconds allow at the rate of a one-problem-per-core probe - the whole batch if it fits."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(x0[i:i + 1], workload) for i in range(min(cores, len(x0)))])  # warm-up: imports, page-in
        t = time.perf_counter()
        pool.map(_cpu_worker, [(x0[i:len(x0):cores][:2], workload) for i in range(cores)])  # probe: two problems per core
        probe = (time.perf_counter() - t) / 2
        if per_core > 0:
            n = min(len(x0), cores * per_core)
        else:
            n = int(min(len(x0), max(cores, budget_s / max(probe, 1e-3) * cores)))
            n -= n % cores if n >= cores else 0
        chunks = [x0[i:n:cores] for i in range(cores)]
        t = time.perf_counter()
        pool.map(_cpu_worker, [(c, workload) for c in chunks])
        wall = time.perf_counter() - t
    return n / wall, n


def run_reference(args):
    os.environ["CAFE_HOST_ONLY"] = "1"   # before cafe_mpc_b200 is imported: deck builders from libcafe_host.so, no CUDA library in this process
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    prob, opt, gen_x0, n0 = make_problem(args.workload)
    per_core = args.cpu_per_core
    cfg = make_config(args, world)
    Bg = cfg["global_batch"]
    x0 = gen_x0(Bg if Bg <= 8192 else 8192)
    vals, n, n_first = [], 0, 0
    for i in range(args.warmup + args.steps):
        # the first timed step covers the WHOLE batch when it fits about 20 s of wall time on these cores (MHPC trot: 4096 problems ~ 14 s on 16 cores),
        # warm-up and later steps a prefix of about 2 s / 5 s (a rate does not depend on the sample size; the run stays within a few minutes)
        budget = 2.0 if i < args.warmup else (20.0 if i == args.warmup else 5.0)
        v, n = cpu_sample(x0, cores, per_core, args.workload, budget_s=budget)
        if i >= args.warmup:
            vals.append(v)
        if i == args.warmup:
            n_first = n
    v = sum(vals) / len(vals)
    n = n_first
    sample = ("first timed step = %s %d problems of the workload's table (later steps: a prefix of about 5 s), one single-threaded instance of the CPU restatement "
              "of the reference solver (oracle/, reference CasADi C from oracle/_ref) per host core over disjoint slices; the reference itself needs Eigen / Boost / "
              "Pinocchio / LCM and cannot be built in this image" % ("all" if n >= Bg else "the first", n))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=4096, help="global batch (--scaling strong) or problems per GPU (--scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--workload", default="mhpc", choices=["mhpc", "hkd", "barrel", "loco", "barrel_to"])
    ap.add_argument("--gain-knots", type=int, default=8)
    ap.add_argument("--cpu-per-core", type=int, default=0, help="CPU baseline: problems per core (0: as many as ~20 s / ~10 s allow, the whole batch if it fits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import api as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = make_config(args, world)
    Bg = cfg["global_batch"]
    prob, opt, gen_x0, n0 = make_problem(args.workload)
    x0_all = gen_x0(Bg) if Bg <= 8192 else np.tile(gen_x0(8192), ((Bg + 8191) // 8192, 1))[:Bg]
    lo, hi = capi.shard_range(Bg, world, rank)
    per = (Bg + world - 1) // world
    B = hi - lo
    x0 = np.ascontiguousarray(x0_all[lo:hi])
    solver = cm.MultiPhaseDDP(prob, local, per)
    if world > 1:
        # the NCCL id of the solver's own communicator travels over torch.distributed (plumbing); the gather itself is the C ABI's
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        solver.comm_init_rank(world, rank, bytes(idt.cpu().numpy().tobytes()))
    if args.workload == "barrel_to":
        solver.set_initial_guess(prob.initial_guess(x0))   # BarrelRollTO.cpp:131-147: part of the problem set-up, stays in force
    # device-resident inputs: x0 as [n0][ldb]
    x0_dev = torch.from_numpy(np.ascontiguousarray(x0.T)).cuda()
    x0_pin = torch.from_numpy(x0).pin_memory()
    rec = solver.command_size(args.gain_knots)
    n_out = world * per if world > 1 else B
    cmd_dev = torch.empty((n_out if rank == 0 else 1, rec), dtype=torch.float64, device="cuda")
    cmd_pin = torch.empty((n_out if rank == 0 else 1, rec), dtype=torch.float64).pin_memory()
    # second slot of the two-deep collection pipeline (records of step i travel while step i + 1 is solved)
    cmd_dev2 = [cmd_dev, torch.empty_like(cmd_dev)]
    cmd_pin2 = [cmd_pin, torch.empty((n_out if rank == 0 else 1, rec), dtype=torch.float64).pin_memory()]

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
            solver.gather_commands(args.gain_knots, per, cmd_dev.data_ptr())   # pack + the one collective: NCCL send / recv to rank 0
            if rank == 0:
                cmd_pin.copy_(cmd_dev, non_blocking=False)           # rank 0 reads EVERY rank's records on the host
        else:
            solver.get_commands(args.gain_knots, out=cmd_pin.numpy())  # D2H of the command records
        return float(cmd_pin[0, 0])

    def run_e2e_pipelined(n):
        """n end-to-end steps as a user's serving loop would run them: every step copies its x0 from page-locked host memory, solves, and hands its
        records to the asynchronous collection (pack on the solver's stream; D2H - at N > 1 the NCCL send / recv to rank 0 and the D2H of ALL records
        on rank 0 - on a copy stream) which overlaps the NEXT step's solve; the records of step i - 1 are read on the host while step i's travel.
        Every byte of every step has landed when this returns."""
        sink = 0.0
        for i in range(n):
            slot = i & 1
            solver.set_initial_condition(x0_pin.numpy())
            solver.solve(opt)                                        # H2D of x0 inside
            if world > 1:
                solver.gather_commands_async(args.gain_knots, per, cmd_dev2[slot].data_ptr() if rank == 0 else 0, cmd_pin2[slot].data_ptr() if rank == 0 else 0, slot)
            else:
                solver.get_commands_async(args.gain_knots, cmd_pin2[slot].numpy(), slot)
            if i > 0:
                solver.commands_wait(slot ^ 1)
                sink += float(cmd_pin2[slot ^ 1][0, 0])
        solver.commands_wait((n - 1) & 1)
        sink += float(cmd_pin2[(n - 1) & 1][0, 0])
        return sink

    # N > 1, second way of collecting the records (reported beside `e2e` as `e2e_host_collect`): no inter-GPU traffic at all - every rank copies its
    # own shard over its own PCIe link straight into ONE page-locked host buffer that all ranks map (POSIX shared memory registered with
    # cudaHostRegister in every process), rank 0 reads the whole batch from it once every rank has finished
    shm, host_all = None, None
    if world > 1:
        from multiprocessing import shared_memory
        name = [None]
        if rank == 0:
            try:   # only where /dev/shm has room for the records (a short tmpfs would end the process with SIGBUS on first touch)
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize > 2 * n_out * rec * 8:
                    shm = shared_memory.SharedMemory(create=True, size=n_out * rec * 8)
                    name[0] = shm.name
            except Exception:
                shm, name[0] = None, None
        dist.broadcast_object_list(name, 0)
        if name[0] is not None:
            try:
                if rank != 0:
                    shm = shared_memory.SharedMemory(name=name[0])
                    try:   # the creator (rank 0) unlinks it; keep this process' resource tracker from doing so a second time at exit
                        from multiprocessing import resource_tracker
                        resource_tracker.unregister(shm._name, "shared_memory")
                    except Exception:
                        pass
                host_all = np.ndarray((n_out, rec), dtype=np.float64, buffer=shm.buf)
                if int(torch.cuda.cudart().cudaHostRegister(host_all.ctypes.data, host_all.nbytes, 0)) != 0:
                    host_all = None   # not page-lockable here: the figure is omitted
            except Exception:
                host_all = None
    my_rows = host_all[rank * per: rank * per + B] if host_all is not None else None

    def step_e2e_host():
        solver.set_initial_condition(x0_pin.numpy())
        solver.solve(opt)
        solver.get_commands(args.gain_knots, out=my_rows)           # D2H of this rank's records into the shared page-locked buffer
        dist.barrier()                                               # every shard has landed
        return float(host_all[0, 0])

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
    wall_e2e_blocking = time.perf_counter() - t1
    run_e2e_pipelined(2)
    barrier()
    t1 = time.perf_counter()
    run_e2e_pipelined(args.steps)
    barrier()
    wall_e2e = time.perf_counter() - t1
    wall_e2e_host = 0.0
    if world > 1:
        ok_all = torch.tensor([1.0 if host_all is not None else 0.0], device="cuda")
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        if float(ok_all.cpu()) > 0:
            step_e2e_host()
            barrier()
            t2 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e_host()
            barrier()
            wall_e2e_host = time.perf_counter() - t2
    sampler.stop.set()
    sampler.join(timeout=2)
    it_sum = float(sum(i["iter"] for i in info)); it_max = float(max(i["iter"] for i in info))
    tt = torch.tensor([wall, wall_e2e, dev_ms, it_max, wall_e2e_host, wall_e2e_blocking], dtype=torch.float64, device="cuda")
    ts = torch.tensor([it_sum, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    wall, wall_e2e, dev_ms, it_max, wall_e2e_host, wall_e2e_blocking = [float(v) for v in tt.cpu()]
    it_sum, launches = [float(v) for v in ts.cpu()]

    roof = None
    cpu = None
    if rank == 0:
        roof = roofline(args, solver, prob, step_resident, local, B, Bg)
        if not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            v, n = cpu_sample(x0_all, cores, args.cpu_per_core, args.workload, budget_s=10.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%s %d problems of the same SplitMix64 table (about 10 s of CPU work), one single-threaded oracle instance per core" % ("all" if n >= Bg else "the first", n)}
        cfg = dict(cfg)
        cfg.update({"per_gpu_batch": per, "parallelism": "batch sharded over %d GPU(s), no collective in the solve" % world,
                    "mean_ddp_iterations": it_sum / Bg, "max_ddp_iterations": it_max})
        print(json.dumps({
            "metric": METRIC, "value": Bg * args.steps / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * wall / args.steps, "device_ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": Bg * args.steps / wall_e2e, "unit": UNIT, "h2d_bytes_per_step": int(Bg * n0 * 8), "d2h_bytes_per_step": int(n_out * rec * 8),
                    "what": "per step: cafe_gpu_solve_batch(host x0) + %s (Xbar,Ubar,Y all knots; K,Qu,Quu,Qux first %d knots); the collection of step i "
                            "(copy stream) overlaps the solve of step i + 1, all records of all steps on the host before the clock stops; byte counts are whole-job totals per step"
                            % ("cafe_gpu_gather_commands_async (pack + NCCL send/recv to rank 0 + D2H of all %d records on rank 0)" % n_out if world > 1 else "cafe_gpu_get_commands_async", args.gain_knots)},
            "e2e_blocking": {"value": Bg * args.steps / wall_e2e_blocking, "unit": UNIT,
                             "what": "the same steps with the blocking calls (%s): every step waits for its own records before the next x0 is copied"
                                     % ("cafe_gpu_gather_commands + D2H on rank 0" if world > 1 else "cafe_gpu_get_commands")},
            "e2e_host_collect": ({"value": Bg * args.steps / wall_e2e_host, "unit": UNIT,
                                  "what": "cafe_gpu_solve_batch(host x0) + cafe_gpu_get_commands of every rank's shard over its own PCIe link into one page-locked host "
                                          "buffer shared by the ranks (POSIX shared memory + cudaHostRegister): no inter-GPU traffic, rank 0 holds all %d records" % n_out}
                                 if wall_e2e_host > 0 else None),
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "cpu_baseline": cpu}))
    if world > 1:
        if host_all is not None:
            torch.cuda.cudart().cudaHostUnregister(host_all.ctypes.data)
        del my_rows, host_all
        dist.barrier()
        if shm is not None:
            try:
                shm.close()
                if rank == 0:
                    shm.unlink()
            except Exception:
                pass
        dist.destroy_process_group()


def roofline(args, solver, prob, step_resident, local, B, Bg):
    """Per-kernel roofline figures of rank 0's shard, measured live with per-launch CUDA events on the solver's stream (profiling mode:
    one stream, a host sync per launch). For every kernel: algorithmic bytes and flops per work unit x the units it was
    launched on, its measured time, BOTH fractions (HBM and fp64), and the DRAM bytes of one full-batch launch from the committed
    ncu --set full capture where one exists. `roofline` itself is the kernel with the largest share of the step."""
    import cafe_mpc_b200 as cm
    solver.set_profiling(True)
    step_resident()
    tm = solver.get_timing()
    solver.set_profiling(False)
    pinfo = solver.get_solver_info()
    phases = prob.phases()
    peak64 = cm.measure_fp64_peak(local)
    hbm = _hbm_peak()
    tot_ms = max(sum(tm["ms"].values()), 1e-9)
    # per-unit algorithmic work of the whole-body kernels, averaged over the running knots of the deck (contact rows differ per phase)
    wb = [(p.horizon, 3 * sum(1 for c in p.contact if c > 0)) for p in phases if p.model == 1]
    n_wbk = max(sum(h for h, _ in wb), 1)
    avg = lambda f: sum(h * f(nr) for h, nr in wb) / n_wbk
    lxx_nnz = _lxx_nnz(prob, phases) if wb else 0.0
    model = {   # doubles in, doubles out, flop per unit
        "wb_terms": (36, TM_NNZ, OPS_TERMS),
        "wb_fwd": (TM_NNZ + 36 + 12 + 72, 36 + 12 + 18 + 3, avg(wb_fwd_flop)),
        "wb_derivs": (36 + 18 + 12, DP_NNZ, OPS_DERIVS),
        "wb_sens": (DP_NNZ + TM_NNZ, 18 * 48 + avg(lambda nr: nr * 48), avg(wb_sens_flop)),
        "wb_cost": (TM_NNZ + 72 + 60, 36 + 12 + 12 + 12 + 36 + lxx_nnz + 1, 4000.0),
    }
    names = {"roll": "k_roll", "lq": "k_lq", "bwd": "k_bwd2", "wb_terms": "k_wb_terms", "wb_fwd": "k_wb_fwd", "wb_derivs": "k_wb_derivs", "wb_sens": "k_wb_sens",
             "wb_cost": "k_wb_cost", "select": "k_ls_scan / k_select / k_compact", "accept": "k_accept", "misc": "k_init / k_unpack"}
    full = args.workload == "mhpc" and B == 4096
    per = {}
    for fam in tm["ms"]:
        ms, nl, units = tm["ms"][fam], tm["launches"][fam], tm["units"][fam]
        e = {"kernel": names[fam], "ms": ms, "launches": nl, "share_of_step": ms / tot_ms}
        if fam in model and ms > 0:
            din, dout, fl = model[fam]
            e.update({"units": units, "bytes_per_unit": 8.0 * (din + dout), "flop_per_unit": fl,
                      "hbm_gbs": 8.0 * (din + dout) * units / (ms * 1e-3) / 1e9, "fp64_tflops": fl * units / (ms * 1e-3) / 1e12})
        if fam == "bwd" and ms > 0:
            f_sweep = sum(F_BWD[p.model] * p.horizon for p in phases)
            f_lin = sum(F_LIN[p.model] * p.horizon for p in phases)
            flops = sum(i["reg_iter_total"] * f_sweep + i["iter"] * f_lin for i in pinfo)
            b_sweep = 8.0 * sum(_bwd_doubles(p.model, lxx_nnz) * p.horizon for p in phases)
            byts = sum(i["reg_iter_total"] * b_sweep for i in pinfo)
            e.update({"units": units, "flop_total": flops, "bytes_total": byts, "hbm_gbs": byts / (ms * 1e-3) / 1e9, "fp64_tflops": flops / (ms * 1e-3) / 1e12})
        if "hbm_gbs" in e:
            e["hbm_frac"] = e["hbm_gbs"] / hbm
            e["fp64_frac"] = e["fp64_tflops"] / peak64
        tb, src = _ncu_traffic(names[fam]) if full else (None, None)
        if tb:
            e["dram_bytes_full_batch_launch_ncu"] = tb
            e["traffic_source"] = src
        per[fam] = e
    top = max(per, key=lambda k: per[k]["ms"])
    t = per[top]
    bound = "tensor" if top == "bwd" else "hbm"
    ach = t.get("fp64_tflops") if bound == "tensor" else t.get("hbm_gbs")
    roof = {"bound": bound, "kernel": t["kernel"], "achieved": ach, "peak": peak64 if bound == "tensor" else hbm, "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
            "frac": (ach / (peak64 if bound == "tensor" else hbm)) if ach else None, "traffic": t.get("dram_bytes_full_batch_launch_ncu"),
            "traffic_source": t.get("traffic_source"), "avg_launch_ms": t["ms"] / max(t["launches"], 1), "share_of_step": t["share_of_step"],
            "pipe": "fp64 tensor pipe (DMMA m8n8k4) + fp64 FMA pipe" if bound == "tensor" else None,
            "peak_source": ("measured live: cafe_gpu_measure_fp64_peak = max(DFMA chains, DMMA m8n8k4 chains) microbenchmark; MEASURED_PEAKS.json has no fp64 entry"
                            if bound == "tensor" else "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"),
            "fp64_peak_tflops": peak64, "hbm_peak_gbs": hbm,
            "achieved_definition": "k_bwd2: dense flop counts of the reference's sweep formulas (SURVEY.md 8d: WB 373 392 / HKD 233 280 / SRB 30 096 per knot and sweep + "
                                   "linear rollout) x the bit-exact per-problem sweep counters / live kernel time; other kernels: algorithmic bytes (inputs + outputs of "
                                   "the kernel per work unit, DESIGN.md section 3) x launched units / live kernel time",
            "kernel_ms": tm["ms"], "per_kernel": per}
    return roof


def _bwd_doubles(model, lxx_nnz):
    """doubles one sweep pass + the linear rollout move per knot (inputs of the sweep, its outputs, inputs of the linear rollout)"""
    if model == 1:
        ab, cd, k = 18 * 48, 12 * 48, 12 * 36
        return (ab + cd + lxx_nnz + 36 + 12 + 36 + 12 + 12 + 36) + (2 * k + 144 + k + 12 + 36 + 12) + (k + ab + lxx_nnz + 12 + 48 + 36 + 12 + 36)
    n, m = (24, 24) if model == 0 else (12, 12)
    return (n * n + n * m + n * n + m * m + 3 * n + m) + (2 * m * n + m * m + m * n + m + n + m) + (m * n + n * n + n * m + n * n + m * m + 3 * n + 2 * m)


def _lxx_nnz(prob, phases):
    """mean number of structural non-zeros of lxx over the running whole-body knots (cafe_deck_lq_pattern)"""
    import ctypes as C
    from cafe_mpc_b200.lib import lib
    tot, cnt = 0, 0
    for pi, p in enumerate(phases):
        if p.model != 1:
            continue
        for k in range(p.horizon):
            w = (C.c_ulonglong * 21)()
            if lib.cafe_deck_lq_pattern(prob.deck, pi, k, 2, w) > 0:
                tot += sum(bin(x).count("1") for x in w); cnt += 1
    return tot / max(cnt, 1)


def _ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full summary of the newest round (profiles/rNN_traffic.json)."""
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
