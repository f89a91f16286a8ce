"""Receding-horizon batch (first slice; dev tool): B perturbed MHPC trot problems, one cold solve with the
initialisation caps, then `steps` MPC updates: shift the plan by dt_mpc / dt_wb = 2 knots, warm start from the previous solution
(cafe_mpc_b200/mpc.py), re-solve under the run-time caps (max_AL_iter_runtime x max_DDP_iter_runtime, MHPCLocomotion.cpp:86-87).
The "measured" state of the next step is the plan's own prediction two knots ahead plus a small disturbance (no simulator here).
Prints one JSON line per step. usage: mpc_loop.py [B] [steps] [k0] [update|host|device]
"update" (default): ONE solver for the whole loop, cafe_gpu_update_deck per step (deck replaced, device shift, arena re-used);
"device": the shift runs on the GPU (cafe_gpu_shift_guess from the previous solver's arrays, cafe_gpu_get_planned_state for the next
initial state); "host": D2H of the packed solutions, numpy shift (cafe_mpc_b200/mpc.py), H2D of the guess."""
import copy, json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import mpc, workload

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mode = sys.argv[4] if len(sys.argv) > 4 else "update"
csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")
opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info"))
ort = copy.copy(opt)
ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
x0 = workload.mhpc_batch(min(B, 512))
x0 = np.tile(x0, ((B + len(x0) - 1) // len(x0), 1))[:B]
noise = 2e-3 * (x0 - x0[0])
prob = cm.MHPCProblem(csv, k0=k0)
s = cm.MultiPhaseDDP(prob, 0, B)
s.set_initial_condition(x0)
s.solve(opt)
s.solve(opt)
info = s.get_solver_info()
print(json.dumps({"step": 0, "k0": k0, "phases": [p.horizon for p in prob.phases()], "solve_ms": round(s.solve_ms(), 2), "solves_per_s": round(B / s.solve_ms() * 1e3, 1),
                  "mean_iter": sum(i["iter"] for i in info) / B, "mean_cost": float(np.mean([i["cost"] for i in info])), "max_feas": max(i["feas"] for i in info)}), flush=True)
sol = s.get_solution() if mode == "host" else None
cmd_buf = None
if mode == "update":
    import torch   # plumbing only: a page-locked host buffer for the wire records
    cmd_buf = torch.empty((B, 1080 * 8), dtype=torch.float32, pin_memory=True).numpy()
for step in range(1, steps + 1):
    t0 = time.perf_counter()
    k1 = k0 + 2
    p1 = cm.MHPCProblem(csv, k0=k1, mpc_update_nsteps=2)   # deck after MHPCProblem::update: a freshly opened tail phase has no shooting states
    if mode == "host":
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol)
        x1 = mpc.state_at(prob, mpc.unpack_batch(prob, sol), 2) + noise
        t1 = time.perf_counter()
        s.close()
        s = cm.MultiPhaseDDP(p1, 0, B)
        s.set_initial_condition(x1)
        s.set_initial_guess(guess)
    elif mode == "update":
        x1 = s.planned_state(2) + noise
        t1 = time.perf_counter()
        s.update_deck(p1, 2)
        s.set_initial_condition(x1)
    else:
        x1 = s.planned_state(2) + noise
        s1 = cm.MultiPhaseDDP(p1, 0, B)
        s1.set_initial_condition(x1)
        s1.shift_guess_from(s, k0, k1)
        t1 = time.perf_counter()
        s.close()
        s = s1
    t2 = time.perf_counter()
    s.solve(ort)
    ms = s.solve_ms()
    t2b = time.perf_counter()
    if mode == "host":
        sol = s.get_solution()
    else:
        cmd = s.get_lcm_commands(8, out=cmd_buf)   # what the controller consumes (float32 MHPC_Command_lcmt fields)
    t3 = time.perf_counter()
    info = s.get_solver_info()        # diagnostics of this tool, outside the step's wall time
    print(json.dumps({"step": step, "mode": mode, "k0": k1, "phases": [p.horizon for p in p1.phases()], "solve_ms": round(ms, 2), "solves_per_s": round(B / ms * 1e3, 1),
                      "mean_iter": sum(i["iter"] for i in info) / B, "mean_cost": float(np.mean([i["cost"] for i in info])),
                      "max_feas": max(i["feas"] for i in info), "shift_ms": round(1e3 * (t1 - t0), 1), "handle_and_upload_ms": round(1e3 * (t2 - t1), 1),
                      "solve_wall_ms": round(1e3 * (t2b - t2), 1), "readback_ms": round(1e3 * (t3 - t2b), 1), "step_wall_ms": round(1e3 * (t3 - t0), 1)}), flush=True)
    prob, k0 = p1, k1
