"""include/hsddp_facade/: the reference's own C++ call sequence (MHPC-Trajopt/test/testMHPCProblem.cpp:9-89, MHPCLocomotion.cpp:91-150,
HKDMPC.cpp:20-140) compiled against the facade headers and run through the C ABI (examples/facade_mhpc.cpp), compared with the same
solves made from Python: the cold solve bit for bit, the MPC updates (host-side trajectory shift of MHPCProblem::update /
HKDProblem::update in the facade vs cafe_mpc_b200/mpc.py) bit for bit as well."""
import copy
import json
import os
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
EXE = os.path.join(REPO, "tests/_build/facade_mhpc")
RUN = os.path.join(REPO, "data/_run")   # the reference resolves its files relative to "../"


def build_example():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    os.makedirs(RUN, exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(REPO, "include/hsddp_facade"), "-I" + os.path.join(REPO, "include"),
                    os.path.join(REPO, "examples/facade_mhpc.cpp"), "-L" + os.path.join(REPO, "cafe_mpc_b200"), "-lcafe_gpu",
                    "-Wl,-rpath," + os.path.join(REPO, "cafe_mpc_b200"), "-o", EXE], check=True)


def run_example(*args):
    r = subprocess.run([EXE] + [str(a) for a in args], cwd=RUN, capture_output=True, text=True, timeout=600)
    return r, [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]


def test_facade_compiles_and_fails_loudly_without_a_gpu(built_lib):
    """CPU: the reference's call sequence compiles with only the include path changed; without a device the solve refuses (no fallback)"""
    import torch
    build_example()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the refusal path is not reachable")
    r, lines = run_example("mhpc", 0)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and not lines


def _compare(line, info, sol_phases, lead):
    assert [line["iter"], line["ls"], line["n_hist"]] == [info["iter"], info["ls_iter_total"], info["n_hist"]]
    assert line["cost"] == info["cost"] and line["feas"] == info["feas"]
    assert np.array_equal(np.asarray(line["u0"]), sol_phases[0]["Ubar"][0])
    assert np.array_equal(np.asarray(line["x_end"]), sol_phases[lead - 1]["Xbar"][-1])


@pytest.mark.gpu
def test_facade_mhpc_sequence_equals_the_c_abi_path(cm):
    from cafe_mpc_b200 import mpc
    build_example()
    k0, n_upd = 8, 3                                   # updates at 10, 12 (front phase removed, tail phase opened), 14
    r, lines = run_example("mhpc", n_upd, k0)
    assert r.returncode == 0, r.stderr[-2000:]
    assert [l["tag"] for l in lines] == ["mhpc_initial"] + ["mhpc_update_%d" % i for i in range(1, n_upd + 1)]
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    x0 = np.zeros((1, 36)); x0[0, 2] = 0.2486; x0[0, 6:18] = np.tile([0, -0.8, 1.6], 4)
    prob = cm.MHPCProblem(CSV, k0=k0)
    s = cm.MultiPhaseDDP(prob, 0, 1); s.set_initial_condition(x0); s.solve(opt)
    sol = s.get_solution(); al = s.get_al_params()
    n_wb = sum(1 for p in prob.phases() if p.model == 1)
    _compare(lines[0], s.get_solver_info()[0], cm.unpack_solution(prob.deck, sol[0]), n_wb)
    reg_total = s.get_solver_info()[0]["reg_iter_total"]
    assert lines[0]["reg_total"] == reg_total
    ort = copy.copy(opt); ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    for step in range(n_upd):
        k1 = k0 + 2
        p1 = cm.MHPCProblem(CSV, k0=k1, mpc_update_nsteps=2)
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol)
        x1 = mpc.state_at(prob, cm.unpack_solution(prob.deck, sol[0]), 2)[None]
        s1 = cm.MultiPhaseDDP(p1, 0, 1); s1.set_initial_condition(x1); s1.set_initial_guess(guess)
        s1.set_al_params(mpc.shift_al(prob, k0, p1, k1, al)); s1.solve(ort)     # the phases carry sigma / lambda over the update, like the reference's
        sol = s1.get_solution(); al = s1.get_al_params()
        n_wb = sum(1 for p in p1.phases() if p.model == 1)
        _compare(lines[1 + step], s1.get_solver_info()[0], cm.unpack_solution(p1.deck, sol[0]), n_wb)
        reg_total += s1.get_solver_info()[0]["reg_iter_total"]       # the reference never resets reg_iter_total_ (MultiPhaseDDP.h:113)
        assert lines[1 + step]["reg_total"] == reg_total
        prob, k0 = p1, k1


@pytest.mark.gpu
def test_facade_hkd_sequence_equals_the_c_abi_path(cm):
    from cafe_mpc_b200 import mpc
    build_example()
    k0, n_upd = 6, 4                                   # updates at 8, 10, 12 (front phase removed), 14
    r, lines = run_example("hkd", n_upd, k0)
    assert r.returncode == 0, r.stderr[-2000:]
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
    prob = cm.HKDProblem(CSV, k0=k0)
    x0 = prob.reference_records()[0, :24][None].copy(); x0[0, 5] += 0.01; x0[0, 9] += 0.05
    s = cm.MultiPhaseDDP(prob, 0, 1); s.set_initial_condition(x0); s.solve(opt)
    sol = s.get_solution(); al = s.get_al_params()
    _compare(lines[0], s.get_solver_info()[0], cm.unpack_solution(prob.deck, sol[0]), len(prob.phases()))
    ort = copy.copy(opt); ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    for step in range(n_upd):
        k1 = k0 + 2
        p1 = cm.HKDProblem(CSV, k0=k1, mpc_update=True)
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol)
        x1 = mpc.state_at(prob, cm.unpack_solution(prob.deck, sol[0]), 2)[None]
        s1 = cm.MultiPhaseDDP(p1, 0, 1); s1.set_initial_condition(x1); s1.set_initial_guess(guess)
        s1.set_al_params(mpc.shift_al(prob, k0, p1, k1, al)); s1.solve(ort)
        sol = s1.get_solution(); al = s1.get_al_params()
        _compare(lines[1 + step], s1.get_solver_info()[0], cm.unpack_solution(p1.deck, sol[0]), len(p1.phases()))
        prob, k0 = p1, k1
