"""Multi-GPU paths of the C ABI (include/cafe_gpu.h): the batch cut over several GPUs gives, bit for bit, the records
of one GPU. (a) one process, cafe_gpu_create_multi; (b) one process per GPU under torchrun with the solver's own NCCL communicator
(cafe_gpu_comm_init_rank + cafe_gpu_gather_commands). Both need at least two GPUs and skip otherwise (`gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch  # noqa: F401  (before the solver library: the process then shares PyTorch's NCCL build, see cafe_gpu.h)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pytestmark = pytest.mark.gpu
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _single(prob, opt, x0):
    import cafe_mpc_b200 as cm
    s = cm.MultiPhaseDDP(prob, 0, len(x0))
    s.set_initial_condition(x0)
    s.solve(opt)
    return s.get_solver_info(), s.get_commands(8)


def test_create_multi_on_one_gpu_equals_single_solver():
    """cafe_gpu_create_multi with one device is the single-GPU path (no NCCL involved): same counters, same records, bit for bit"""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import api, workload
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    x0 = workload.mhpc_batch(24)
    info1, cmd1 = _single(prob, opt, x0)
    m = api.MultiGPUDDP(prob, 1, 24)
    m.solve(x0, opt)
    assert m.get_solver_info() == info1
    assert np.array_equal(m.get_commands(8), cmd1)


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_mpc_update_equals_single_solver(ndev):
    """cafe_gpu_multi_update_deck: two MPC updates on every GPU's slice == the same updates on one solver, bit for bit"""
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    import copy
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import api, workload
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    ort = copy.copy(opt); ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    B, k0 = 22, 8
    x0 = workload.mhpc_batch(B)
    prob = cm.MHPCProblem(CSV, k0=k0)
    one = cm.MultiPhaseDDP(prob, 0, B); one.set_initial_condition(x0); one.solve(opt)
    m = api.MultiGPUDDP(prob, ndev, B); m.solve(x0, opt)
    for step in range(2):
        k0 += 2
        p1 = cm.MHPCProblem(CSV, k0=k0, mpc_update_nsteps=2)
        x1 = one.planned_state(2) + 1e-3 * (x0 - x0[0])
        one.update_deck(p1, 2); one.set_initial_condition(x1); one.solve(ort)
        m.update_deck(p1, 2); m.solve(x1, ort)
        assert m.get_solver_info() == one.get_solver_info(), step
        assert np.array_equal(m.get_commands(8), one.get_commands(8)), step


@pytest.mark.parametrize("ndev", [2, 4])
def test_create_multi_equals_single_gpu_bitwise(ndev):
    if _ngpu() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import api, workload
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    B = 70   # ragged shards: 35 + 35, 18 + 18 + 18 + 16
    x0 = workload.mhpc_batch(B)
    info1, cmd1 = _single(prob, opt, x0)
    m = api.MultiGPUDDP(prob, ndev, B)
    m.solve(x0, opt)
    assert m.get_solver_info() == info1
    assert np.array_equal(m.get_commands(8), cmd1)   # NCCL-gathered on GPU 0, then D2H


def test_torchrun_gather_equals_single_gpu_bitwise(tmp_path):
    """two ranks under torchrun, each with its own solver and the solver's NCCL communicator: rank 0's gathered records == one GPU's"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "gathered.npy"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(REPO, "tools", "multi_rank_check.py"), "--batch", "70", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["bitwise_equal"] and res["info_equal"] and res["async_bitwise_equal"], res
