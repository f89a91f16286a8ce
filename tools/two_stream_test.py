"""Experiment (dev tool): one 4096-problem handle against two 2048-problem handles driven by two host threads (their kernels overlap
on the device and fill each other's wave tails). usage: two_stream_test.py [B] [parts]"""
import os, sys, threading, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import workload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
parts = [int(a) for a in sys.argv[2:]] or [1, 2, 3, 4]
csv = os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv")
prob = cm.MHPCProblem(csv); opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info"))
x0 = workload.mhpc_batch(B)
for P in parts:
    cuts = [B * i // P for i in range(P + 1)]
    hs = [cm.MultiPhaseDDP(prob, 0, cuts[i + 1] - cuts[i]) for i in range(P)]
    xd = [torch.from_numpy(np.ascontiguousarray(x0[cuts[i]:cuts[i + 1]].T)).cuda() for i in range(P)]
    def run(i):
        n = cuts[i + 1] - cuts[i]
        hs[i].solve_device(xd[i].data_ptr(), n, n, opt)
    def step():
        th = [threading.Thread(target=run, args=(i,)) for i in range(P)]
        for t in th: t.start()
        for t in th: t.join()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print({"parts": P, "ms_per_batch": round(dt * 1e3, 2), "solves_per_s": round(B / dt, 1)}, flush=True)
    for h in hs: h.close()
