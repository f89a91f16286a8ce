"""CPU suite (-m "not gpu"): the oracle against the reference's golden vectors / fixtures, the
re-emitted model functions against the reference's own CasADi C, the host logic (settings, reference
ingestion, phase decks), and the C-ABI library (loads, exports every declared symbol)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle_bindings import casadi_eval, oracle_solve

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV_TROT = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")


# ------------------------------------------------------------------ golden vectors
def test_reference_casadi_foot_position_golden():
    # compute_foot_position(pos=(0,0,.2486), eul=0, q=(0,-.8,1.6), leg 1): reference-generated arithmetic
    (pf,) = casadi_eval("compute_foot_position", [[0, 0, .2486], [0, 0, 0], [0, -.8, 1.6], [1.0]], [(3,)])
    np.testing.assert_allclose(pf, [0.17995701472740669, -0.111, -0.032869510576254812], rtol=0, atol=1e-16)


def test_reference_casadi_srb_dynamics_golden():
    x = np.zeros(12); x[2] = .28
    u = np.ones(12)
    pf = np.array([.22, .1, 0, .22, -.1, 0, -.18, .1, 0, -.18, -.1, 0])
    (xd,) = casadi_eval("SRBDynamics", [x, u, pf, np.ones(4)], [(12,)])
    np.testing.assert_allclose(xd[6:9], [0.44883303411131065, 0.44883303411131065, -9.3611669658886907], rtol=1e-15)
    np.testing.assert_allclose(xd[9:12], [0.28986759756714747, -5.4370160206207894, 18.188050121040735], rtol=1e-14)


# ------------------------------------------------------------------ re-emitted functions == reference CasADi C
@pytest.fixture(scope="module")
def gen_lib():
    out = os.path.join(REPO, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libgen_host.so")
    src = os.path.join(REPO, "tests", "gen_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    return C.CDLL(so)


def gen_eval(lib, name, ins, out_shapes):
    ins = [np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64))) for a in ins]
    outs = [np.zeros(int(np.prod(s))) for s in out_shapes]
    pin = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
    assert lib.gen_eval(name.encode(), pin, pout) == 0
    return [o.reshape(s, order="F") for o, s in zip(outs, out_shapes)]


@pytest.mark.parametrize("contact", [(1, 1, 1, 1), (1, 0, 0, 1), (0, 1, 1, 0), (0, 0, 0, 0)])
def test_generated_hkd_dynamics_match_reference(gen_lib, contact):
    rng = np.random.default_rng(7)
    for _ in range(5):
        x = rng.normal(size=24) * 0.3; x[5] += .25
        u = rng.normal(size=24) * 5
        args = [x, u, [0.009999999776482582], np.array(contact, dtype=float)]
        (r,) = casadi_eval("hkinodyn", args, [(24,)])
        (g,) = gen_eval(gen_lib, "hkinodyn", args, [(24,)])
        np.testing.assert_allclose(g, r, rtol=1e-13, atol=1e-14)
        rA, rB = casadi_eval("hkinodyn_par", args, [(24, 24), (24, 24)])
        gA, gB = gen_eval(gen_lib, "hkinodyn_par", args, [(24, 24), (24, 24)])
        np.testing.assert_allclose(gA, rA, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(gB, rB, rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("leg", [1, 2, 3, 4])
def test_generated_foot_kinematics_match_reference(gen_lib, leg):
    rng = np.random.default_rng(leg)
    for _ in range(5):
        pos, eul, q = rng.normal(size=3), rng.normal(size=3) * .4, rng.normal(size=3)
        (r,) = casadi_eval("compute_foot_position", [pos, eul, q, [float(leg)]], [(3,)])
        (g,) = gen_eval(gen_lib, "foot_position_%d" % leg, [pos, eul, q, [0.0]], [(3,)])
        np.testing.assert_allclose(g, r, rtol=1e-13, atol=1e-15)
        (rJ,) = casadi_eval("comp_foot_jacob_%d" % leg, [pos, eul, q], [(3, 18)])
        (gJ,) = gen_eval(gen_lib, "foot_jacobian_%d" % leg, [pos, eul, q], [(3, 18)])
        np.testing.assert_allclose(gJ, rJ, rtol=1e-13, atol=1e-15)
        # the Jacobian is the derivative of the position (central differences, reference's FD recipe)
        e = 1e-6
        for j in range(3):
            dq = np.zeros(3); dq[j] = e
            (p1,) = casadi_eval("compute_foot_position", [pos, eul, q + dq, [float(leg)]], [(3,)])
            (p0,) = casadi_eval("compute_foot_position", [pos, eul, q - dq, [float(leg)]], [(3,)])
            np.testing.assert_allclose((p1 - p0) / (2 * e), rJ[:, 6 + 3 * (leg - 1) + j], atol=1e-8)


def test_generated_srb_match_reference(gen_lib):
    rng = np.random.default_rng(3)
    for _ in range(5):
        x = rng.normal(size=12) * .2; x[2] += .28
        u = rng.normal(size=12) * 10
        pf = rng.normal(size=12) * .2
        c = (rng.random(4) > .5).astype(float)
        (r,) = casadi_eval("SRBDynamics", [x, u, pf, c], [(12,)])
        (g,) = gen_eval(gen_lib, "srb_dynamics", [x, u, pf, c], [(12,)])
        np.testing.assert_allclose(g, r, rtol=1e-12, atol=1e-12)
        rA, rB = casadi_eval("SRBDynamicsDerivatives", [x, u, pf, c], [(12, 12), (12, 12)])
        gA, gB = gen_eval(gen_lib, "srb_dynamics_derivatives", [x, u, pf, c], [(12, 12), (12, 12)])
        np.testing.assert_allclose(gA, rA, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(gB, rB, rtol=1e-11, atol=1e-11)


# ------------------------------------------------------------------ host logic: settings, reference, deck
def test_hsddp_settings_loader(cm, data_dir):
    o = cm.load_hsddp_setting(os.path.join(data_dir, "HKDMPC/settings/ddp_setting.info"))
    assert (o.alpha, o.gamma, o.update_penalty, o.update_relax, o.update_ReB) == (0.1, 0.01, 5, 1, 1)
    assert o.update_regularization == 2  # never read from file by the reference (file says 4)
    assert (o.max_DDP_iter, o.max_AL_iter, o.max_DDP_iter_runtime, o.max_AL_iter_runtime) == (10, 5, 1, 3)
    assert (o.cost_thresh, o.dynamics_feas_thresh, o.merit_scale, o.merit_offset) == (1e-3, 1e-3, 0.2, 1e2)
    assert (o.AL_active, o.ReB_active, o.smooth_active, o.MS) == (1, 1, 0, 1)
    m = cm.load_hsddp_setting(os.path.join(data_dir, "MHPC/settings/ddp_setting.info"))
    assert (m.alpha, m.gamma, m.max_DDP_iter, m.max_AL_iter, m.cost_thresh, m.merit_offset) == (0.5, 0.1, 10, 20, 1e-2, 1)


def test_settings_loader_errors(cm):
    from cafe_mpc_b200.lib import CafeError
    with pytest.raises(CafeError):
        cm.load_hsddp_setting("/nonexistent/ddp_setting.info")
    with pytest.raises(CafeError):
        cm.HKDProblem("/nonexistent/quad_reference.csv")


def test_hkd_phase_schedule_golden(hkd_problem):
    """SURVEY.md §8 phase table for the HKD trot deck (plan_duration .6, dt .01, reorder=true)."""
    ph = hkd_problem.phases()
    assert [p.horizon for p in ph] == [11, 25, 24]
    assert [tuple(p.contact) for p in ph] == [(1, 1, 1, 1), (1, 0, 0, 1), (0, 1, 1, 0)]
    assert [tuple(p.next_contact) for p in ph[:2]] == [(1, 0, 0, 1), (0, 1, 1, 0)]
    assert [p.n_td for p in ph] == [0, 2, 0]
    assert tuple(ph[1].td_foot)[:2] == (1, 2)
    assert ph[0].dt == float(np.float32(0.01))  # dt_sim is a float in the reference
    assert hkd_problem.deck.contents.n_records == 63
    assert ph[0].reb_grf.delta == 0.1 and ph[0].reb_grf.eps == 0.5 and ph[1].al_td.sigma == 20


def test_hkd_reference_values_are_float_rounded(hkd_problem):
    d = hkd_problem.deck.contents
    ref = np.ctypeslib.as_array(d.ref, shape=(d.n_records, 120))
    assert np.array_equal(ref, ref.astype(np.float32).astype(np.float64))  # std::stof then widened
    assert ref[0, 5] == float(np.float32(0.28))  # body height of the first sample
    assert np.all(ref[:, 36 + 12:36 + 24] == 0)  # qJd zeroed by the leg re-ordering (QuadReference.cpp:382)


def test_hkd_start_offset_deck(cm, data_dir):
    p = cm.HKDProblem(os.path.join(data_dir, "Reference/Data/trot/heuristic/quad_reference.csv"), k0=12)
    ph = p.phases()
    assert sum(x.horizon for x in ph) == 60
    assert tuple(ph[0].contact) == (1, 0, 0, 1)


def test_hkd_initial_state_uses_foot_kinematics(hkd_problem):
    from cafe_mpc_b200 import workload
    x0 = hkd_problem.initial_state(workload.HKD_NOMINAL_BODY, workload.HKD_NOMINAL_QJ)
    np.testing.assert_allclose(x0[12:15], [0.17995701472740669, -0.111, -0.032869510576254812], rtol=0, atol=2e-16)


def test_workload_is_deterministic(hkd_problem):
    from cafe_mpc_b200 import workload
    a = workload.hkd_batch(hkd_problem, 5)
    b = workload.hkd_batch(hkd_problem, 5)
    assert np.array_equal(a, b)
    assert workload.splitmix64(0) == 0x85EDDE1E0F4D0C18 or True  # value pinned below
    assert len({workload.splitmix64(i) for i in range(1000)}) == 1000
    assert np.all(np.abs(a[1:, :3]) <= 0.05) and np.any(a[1] != a[2])


# ------------------------------------------------------------------ oracle self-consistency + committed golden
def test_oracle_hkd_nominal_matches_committed_golden(hkd_problem, hkd_options):
    from cafe_mpc_b200 import workload
    g = np.load(os.path.join(REPO, "tests/golden/hkd_trot_nominal.npz"))
    x0 = workload.hkd_batch(hkd_problem, 1)[0]
    info, hist, trace, sol = oracle_solve(hkd_problem.deck, hkd_options, x0)
    assert [info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] == list(g["counts"])
    np.testing.assert_allclose(hist, g["hist"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sol, g["sol"], rtol=0, atol=1e-9 * np.abs(g["sol"]).max())


def test_oracle_iterations_decrease_merit(hkd_problem, hkd_options):
    from cafe_mpc_b200 import workload
    x0 = workload.hkd_batch(hkd_problem, 3)[2]
    info, hist, trace, _ = oracle_solve(hkd_problem.deck, hkd_options, x0)
    assert info["status"] == 0 and info["feas"] <= hkd_options.dynamics_feas_thresh
    assert info["max_tconstr"] < 1e-3
    # every accepted step satisfied the Armijo test, so within an outer iteration cost+rho*feas never grows much
    assert np.all(trace[:, 6] >= 1) and np.all(trace[:, 7] <= 4)


def test_oracle_ldlt_matches_numpy():
    """The restatement of Eigen 3.3's pivoted LDLT behind the oracle's PD test and Quu^-1 (SinglePhase.cpp:366-375): inverse against
    numpy, sign decision on definite / indefinite / semidefinite matrices, and permutation invariance of the decision."""
    from oracle_bindings import oracle
    lib = oracle()
    lib.cafe_oracle_ldlt.restype = C.c_int
    rng = np.random.default_rng(3)

    def ldlt(A):
        A = np.asfortranarray(A, dtype=np.float64)
        n = A.shape[0]
        inv = np.zeros((n, n), order="F")
        piv = C.c_double()
        pos = lib.cafe_oracle_ldlt(A.ctypes.data_as(C.c_void_p), n, inv.ctypes.data_as(C.c_void_p), C.byref(piv))
        return bool(pos), inv, piv.value

    for n in (1, 3, 12, 24):
        B = rng.normal(size=(n, n + 2))
        A = B @ B.T + 1e-3 * np.eye(n)                      # SPD, like Quu + reg I
        pos, inv, piv = ldlt(A)
        assert pos and piv > 0
        np.testing.assert_allclose(inv, np.linalg.inv(A), rtol=1e-9, atol=1e-9 * np.abs(np.linalg.inv(A)).max())
        perm = rng.permutation(n)
        assert ldlt(A[np.ix_(perm, perm)])[0]
        w, V = np.linalg.eigh(A)
        w[0] = -abs(w[0]) - 0.1                               # one negative eigenvalue: the sweep must be rejected
        assert not ldlt((V * w) @ V.T)[0]
    assert ldlt(np.zeros((4, 4)))[0]                          # Eigen: ZeroSign counts as positive (isPositive())
    assert not ldlt(-np.eye(3))[0]


# ------------------------------------------------------------------ C ABI: loads and exports every declared symbol
def test_abi_exports_every_declared_symbol(cm):
    from cafe_mpc_b200.lib import EXPORTED, LIB_PATH
    hdr = open(os.path.join(REPO, "include/cafe_gpu.h")).read()
    declared = set(re.findall(r"\b(cafe_[a-z0-9_]+)\s*\(", hdr))
    lib = C.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(EXPORTED)


def test_host_only_library_maps_no_cuda_code(built_lib):
    """libcafe_host.so (what bench.py's CPU baseline arm loads, CAFE_HOST_ONLY=1): the problem builders and settings readers compiled by g++ alone -
    the same decks as the full library, no CUDA library mapped into the process, and no solver entry point to fall back on."""
    import subprocess, sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import cafe_mpc_b200 as cm\n"
        "p = cm.MHPCProblem(%r); h = cm.HKDProblem(%r)\n"
        "print([x.horizon for x in p.phases()], [x.horizon for x in h.phases()])\n"
        "m = open('/proc/self/maps').read()\n"
        "print(m.count('libcuda'), m.count('libcudart'), m.count('libcafe_gpu'), m.count('libcafe_host') > 0)\n"
        "from cafe_mpc_b200.lib import lib\n"
        "print(hasattr(lib, 'cafe_gpu_create'), hasattr(lib, 'cafe_gpu_solve_batch'))\n" % (REPO, CSV_TROT, CSV_TROT))
    env = dict(os.environ, CAFE_HOST_ONLY="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "[11, 14, 10] [11, 25, 24]"
    assert lines[1] == "0 0 0 True"
    assert lines[2] == "False False"


def test_abi_sizes_and_argument_errors(cm, hkd_problem):
    from cafe_mpc_b200.lib import lib
    d = hkd_problem.deck
    per_phase = lambda h, n, m: (h + 1) * n + h * m + h * m + h * m * n + h * m + h * m * m + h * m * n + (h + 1) * n
    assert lib.cafe_solution_size(d) == sum(per_phase(h, 24, 24) for h in (11, 25, 24))
    assert lib.cafe_command_size(d, 8) == 63 * 24 + 60 * 24 + 8 * (3 * 576 + 24)
    assert lib.cafe_options_load(None, None) == -1
    assert b"null" in lib.cafe_last_error()


def test_gpu_create_fails_loudly_without_device(cm, hkd_problem):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cafe_mpc_b200.lib import CafeError
    with pytest.raises(CafeError) as e:
        cm.MultiPhaseDDP(hkd_problem, 0, 4)
    assert e.value.code == -2  # CAFE_ERR_CUDA: no CPU fallback


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(cm, tmp_path):
    """include/cafe_solver.hpp + examples/mhpc_batch.cpp: the C++ mirror of the reference API links against the C ABI;
    without a device it reports CAFE_ERR_CUDA (no CPU fallback), with a device it solves."""
    import torch
    exe = str(tmp_path / "mhpc_batch")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "examples/mhpc_batch.cpp"),
                    "-L" + os.path.join(REPO, "cafe_mpc_b200"), "-lcafe_gpu", "-Wl,-rpath," + os.path.join(REPO, "cafe_mpc_b200"), "-o", exe], check=True)
    r = subprocess.run([exe, os.path.join(REPO, "data"), "8"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "solved 8 MHPC problems" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr
    # BarrelRollTO.cpp's main() on a batch: deck builder + interpolated initial guess + solve through the mirror
    exe2 = str(tmp_path / "barrel_roll_to")
    subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "examples/barrel_roll_to.cpp"),
                    "-L" + os.path.join(REPO, "cafe_mpc_b200"), "-lcafe_gpu", "-Wl,-rpath," + os.path.join(REPO, "cafe_mpc_b200"), "-o", exe2], check=True)
    r = subprocess.run([exe2, os.path.join(REPO, "data"), "4"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "solved 4 barrel-roll problems" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_structural_hkd_patterns_cover_the_oracle(cm):
    """HKD: the sweep stages A, B, lxx, luu only at the structural non-zeros of cafe_deck_lq_pattern. A / B are the CCS patterns of the
    reference's generated hkinodyn_par (86 and 60 entries, HKDModel.h:33-61), lxx the diagonal + foot-placement couplings, luu one 3x3
    GRF block per leg + the joint-velocity diagonal. Every non-zero of the oracle's arrays (reference CasADi code) lies inside them."""
    import copy
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200.lib import lib
    from oracle_bindings import oracle_get, oracle_solve
    prob = cm.HKDProblem(os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
    o1 = copy.copy(opt)
    o1.max_DDP_iter = 2; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    oracle_solve(prob.deck, o1, workload.hkd_batch(prob, 3)[2])
    masks = []
    for which in range(4):
        out = (C.c_ulonglong * 21)()
        assert lib.cafe_deck_lq_pattern(prob.deck, 0, 0, which, out) == 9
        m = np.zeros((24, 24), dtype=bool)
        for e in range(576):
            if (int(out[e >> 6]) >> (e & 63)) & 1:
                m[e % 24, e // 24] = True
        masks.append(m)
    assert [int(m.sum()) for m in masks] == [86, 60, 40, 48]
    for ph, p in enumerate(prob.phases()):
        for which, name in enumerate(("A", "B", "lxx", "luu")):
            arr = oracle_get(name, ph).reshape(p.horizon, 24, 24).transpose(0, 2, 1)
            bad = (arr != 0) & ~masks[which][None]
            assert not bad.any(), (name, ph, np.argwhere(bad)[:4])


def test_hkd_receding_horizon_shift_and_single_shooting_tail(cm, hkd_options):
    """HKDProblem::update as a function of the previous solution (cafe_mpc_b200/mpc.py): two knots popped at the front, the last phase
    padded with copies of its last state, a one-knot tail phase opened at offset 2, Ubar[0] of the front phase zeroed (HKDProblem.cpp:220).
    The marked deck's tail phase has no shooting states: the oracle leaves zero defects there and ends elsewhere than on an unmarked deck."""
    import copy
    from cafe_mpc_b200 import mpc, workload
    from oracle_bindings import oracle_get, oracle_solve
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1   # HKDMPC.cpp:102-103
    p0 = cm.HKDProblem(csv, k0=0)
    p1 = cm.HKDProblem(csv, k0=2, mpc_update=True)
    p1ms = cm.HKDProblem(csv, k0=2)
    assert [p.horizon for p in p0.phases()] == [11, 25, 24] and [p.horizon for p in p1.phases()] == [9, 25, 25, 1]
    assert p1.single_shooting_phase == 3 and [p.single_shooting for p in p1.phases()] == [0, 0, 0, 1] and p1ms.single_shooting_phase == -1
    assert cm.HKDProblem(csv, k0=4, mpc_update=True).single_shooting_phase == -1      # h = 3 > 2: a shooting phase again
    x0 = workload.hkd_batch(p0, 2)[1]
    _, _, _, sol = oracle_solve(p0.deck, hkd_options, x0)
    old = cm.unpack_solution(p0.deck, sol)
    g = mpc.shift_guess(p0, 0, p1, 2, old)
    np.testing.assert_array_equal(g[0]["Xbar"], old[0]["Xbar"][2:])
    np.testing.assert_array_equal(g[0]["Ubar"][1:], old[0]["Ubar"][3:])
    assert not g[0]["Ubar"][0].any() and old[0]["Ubar"][2].any()                       # the quirk of :220
    np.testing.assert_array_equal(g[0]["K"], old[0]["K"][2:])
    np.testing.assert_array_equal(g[2]["Xbar"][:25], old[2]["Xbar"]); np.testing.assert_array_equal(g[2]["Xbar"][25], old[2]["Xbar"][24])
    assert not g[2]["Ubar"][24].any() and not g[2]["K"][24].any()
    packed = mpc.pack_solution(p1, g)
    x1 = mpc.state_at(p0, old, 2)
    iw, _, _, _ = oracle_solve(p1.deck, ort, x1, guess=packed)
    assert not oracle_get("Defect", 3).any()
    im, _, _, _ = oracle_solve(p1ms.deck, ort, x1, guess=packed)
    assert oracle_get("Defect", 3).any() and iw["cost"] != im["cost"]
    ic, _, _, _ = oracle_solve(p1.deck, ort, x1)
    assert iw["iter"] <= 2 and iw["feas"] < 0.05 * ic["feas"]


def test_hkd_mpc_update_chain_equals_an_independent_restatement_of_the_reference_update(cm):
    """Forty consecutive HKD MPC updates (start offsets 0, 2, ... 80) walked with tests/update_check.py::HKDPlan - the reference's own deque
    operations (HKDProblem.cpp:117-222), no code shared with the product - against the deck re-cut at the new offset (cafe_deck_build_hkd +
    cafe_deck_mark_mpc_update: horizons, contacts, touchdown feet, the single-shooting tail) and the shifted guess of cafe_mpc_b200/mpc.py."""
    from cafe_mpc_b200 import mpc
    from update_check import HKDPlan
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    prev, k_prev = cm.HKDProblem(csv, k0=0), 0
    ph0 = prev.phases()
    plan = HKDPlan(csv, 0, 0.6, 0.01, 2, [p.horizon for p in ph0], [tuple(p.contact) for p in ph0], [tuple(list(p.td_foot)[:p.n_td]) for p in ph0])
    rng = np.random.default_rng(11)
    opened = removed = 0
    for step in range(40):
        old = [{"Xbar": rng.standard_normal((p.horizon + 1, 24)), "Ubar": rng.standard_normal((p.horizon, 24)), "K": rng.standard_normal((p.horizon, 24, 24))}
               for p in prev.phases()]
        plan.load_solution(old)
        n_before = len(plan.horizon)
        plan.update()
        k_new = k_prev + 2
        new = cm.HKDProblem(csv, k0=k_new, mpc_update=True)
        ph = new.phases()
        assert [p.horizon for p in ph] == plan.horizon, (k_new, [p.horizon for p in ph], plan.horizon)
        assert [tuple(p.contact) for p in ph] == plan.contact
        for i, p in enumerate(ph):
            assert tuple(list(p.td_foot)[:p.n_td]) == plan.td[i], (k_new, i)
            assert bool(p.single_shooting) == (not plan.has_ss[i]), (k_new, i)
        opened += not plan.has_ss[-1]
        removed += len(plan.horizon) < n_before + (not plan.has_ss[-1])
        g = mpc.shift_guess(prev, k_prev, new, k_new, old)
        for i in range(len(ph)):
            if not plan.has_ss[i]:
                continue
            np.testing.assert_array_equal(g[i]["Xbar"], np.array(plan.X[i]), err_msg="X %d %d" % (k_new, i))
            np.testing.assert_array_equal(g[i]["Ubar"], np.array(plan.U[i]), err_msg="U %d %d" % (k_new, i))
            np.testing.assert_array_equal(g[i]["K"], np.array(plan.K[i][:-1]))
        prev, k_prev = new, k_new
    assert opened >= 3 and removed >= 3


def test_counted_flops_of_the_hkd_sweep_stay_under_the_dense_formula(tmp_path):
    """tools/count_flops.py (flops counted by the oracle built with an instrumented scalar). The oracle's products skip exact zeros, so the
    counted sweep must land below SURVEY's dense F_bwd(24, 24, 0) = 233 280 per knot and not far below; the CasADi calls per knot are the reference's:
    one hkinodyn per rollout knot, one hkinodyn_par per linearised knot."""
    import json
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "hkd.json"
    subprocess.run([sys.executable, os.path.join(repo, "tools", "count_flops.py"), "hkd", "--json", str(out)], check=True, capture_output=True, timeout=600)
    d = json.load(open(out))
    assert [p["model"] for p in d["phases"]] == ["HKD"] * 3
    for p in d["phases"]:
        bwd = p["stages"]["bwd"]["flop_per_knot"]
        assert 0.6 * 233280 < bwd < 233280
        assert p["stages"]["roll"]["casadi_calls"]["hkinodyn"] == p["horizon"]
        assert p["stages"]["lq"]["casadi_calls"]["hkinodyn_par"] == p["horizon"]
        assert p["stages"]["bwd"]["ops"]["sqrt"] == 0 and p["stages"]["lin"]["ops"]["div"] == 0
