"""Run under torchrun (one rank per GPU): every rank solves its contiguous shard, the command records are gathered on rank 0 through
the C ABI's NCCL path (cafe_gpu_comm_init_rank + cafe_gpu_gather_commands) and compared, bit for bit, with the records of ONE GPU
solving the whole batch. Prints one JSON line on rank 0. Used by tests/test_gpu_multi.py and the multi-GPU profiles."""
import argparse, json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch
import torch.distributed as dist
import cafe_mpc_b200 as cm
from cafe_mpc_b200 import api, workload

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=70)
ap.add_argument("--out", default="")
a = ap.parse_args()
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prob = cm.MHPCProblem(os.path.join(R, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
opt = cm.load_hsddp_setting(os.path.join(R, "data/MHPC/settings/ddp_setting.info"))
x0 = workload.mhpc_batch(a.batch)
lo, hi = api.shard_range(a.batch, world, rank)
per = (a.batch + world - 1) // world
s = cm.MultiPhaseDDP(prob, local, per)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
s.comm_init_rank(world, rank, bytes(idt.cpu().numpy().tobytes()))
s.set_initial_condition(x0[lo:hi])
s.solve(opt)
rec = s.command_size(8)
buf = torch.zeros((world * per if rank == 0 else 1, rec), dtype=torch.float64, device="cuda")
s.gather_commands(8, per, buf.data_ptr())
# the asynchronous gather (pack on the solver's stream, NCCL + D2H on the copy stream) while another solve runs: same records on rank 0's host
buf2 = torch.zeros((world * per if rank == 0 else 1, rec), dtype=torch.float64, device="cuda")
pin2 = torch.zeros((world * per if rank == 0 else 1, rec), dtype=torch.float64).pin_memory()
s.gather_commands_async(8, per, buf2.data_ptr() if rank == 0 else 0, pin2.data_ptr() if rank == 0 else 0, 0)
s.set_initial_condition(x0[lo:hi][::-1].copy())     # other initial states: the solver's arrays are overwritten while the records travel
s.solve(opt)
s.commands_wait(0)
s.set_initial_condition(x0[lo:hi])
s.solve(opt)
infos = [None] * world
dist.all_gather_object(infos, s.get_solver_info())
if rank == 0:
    got = buf.cpu().numpy()[: a.batch]
    one = cm.MultiPhaseDDP(prob, 0, a.batch)
    one.set_initial_condition(x0)
    one.solve(opt)
    ref = one.get_commands(8)
    if a.out:
        np.save(a.out, got)
    print(json.dumps({"world": world, "batch": a.batch, "bitwise_equal": bool(np.array_equal(got, ref)), "async_bitwise_equal": bool(np.array_equal(pin2.numpy()[: a.batch], ref)),
                      "info_equal": sum(infos, []) == one.get_solver_info(),
                      "max_abs_diff": float(np.max(np.abs(got - ref)))}))
dist.barrier()
dist.destroy_process_group()
