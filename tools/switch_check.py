"""dev tool (GPU box): the tick-level switches change time, never bits. Runs the MHPC trot batch under every combination of
CAFE_LQ_OVERLAP (linearisation on two streams) and CAFE_BWD_SMALL (256-thread sweep for short lists) in a fresh process each and prints
solve time and a digest of the solution records. usage: switch_check.py [batch ...]"""
import hashlib, json, os, subprocess, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, R)
    import numpy as np
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload
    B = int(sys.argv[2])
    prob = cm.MHPCProblem(os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
    opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info"))
    x0 = workload.mhpc_batch(B)
    s = cm.MultiPhaseDDP(prob, 0, B); s.set_initial_condition(x0)
    ms = []
    for _ in range(4):
        s.solve(opt); ms.append(s.solve_ms())
    cmd = s.get_commands(8)
    print(json.dumps({"batch": B, "ms": round(min(ms), 3), "digest": hashlib.sha1(np.ascontiguousarray(cmd).tobytes()).hexdigest()[:16],
                      "iters": int(sum(i["iter"] for i in s.get_solver_info()))}))
    sys.exit(0)
for B in [int(a) for a in sys.argv[1:]] or [512, 4096]:
    seen = set()
    for ov in ("0", "1"):
        for sm in ("0", "296"):
            env = dict(os.environ, CAFE_LQ_OVERLAP=ov, CAFE_BWD_SMALL=sm)
            r = subprocess.run([sys.executable, __file__, "--child", str(B)], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED " + r.stderr[-400:]
            print("overlap", ov, "bwd_small", sm, line, flush=True)
            if r.returncode == 0: seen.add(json.loads(line)["digest"])
    print("batch", B, "identical records:", len(seen) == 1)
