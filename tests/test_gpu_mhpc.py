"""GPU suite for the MHPC (whole-body + single-rigid-body, cascaded fidelity) path through the C ABI:
BASELINE configs 2 (single MHPC trot solve, GPU == CPU) and 3 (batch 1024 MHPC trot on one B200), plus the
impact-bearing start offset that exercises the WB impact map, touchdown constraints and AL updates."""
import copy
import os

import numpy as np
import pytest

from oracle_bindings import oracle_get, oracle_solve

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")
RTOL = 1e-9


def relerr(g, o, floor=1e-6):
    o = np.asarray(o); g = np.asarray(g)
    return 0.0 if o.size == 0 else float(np.max(np.abs(g - o)) / max(np.max(np.abs(o)), floor))


# Qu and dU are cancellation residuals that vanish at convergence (|Qu| ~ 1e-7 against torque-scale terms ~ 1): their error is
# measured against max(|.|, 1e-3), everything else against its own largest entry
FLOOR = {"Qu": 1e-3, "dU": 1e-3}


@pytest.fixture(scope="module")
def opt(cm):
    return cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))


def solve_gpu(cm, prob, opt, x0):
    s = cm.MultiPhaseDDP(prob, 0, len(x0))
    s.set_initial_condition(x0)
    s.solve(opt)
    return s


def compare_with_oracle(cm, prob, opt, x0, s, which):
    info = s.get_solver_info(); hist = s.get_history(256); sol = s.get_solution()
    for b in which:
        oi, oh, ot, osol = oracle_solve(prob.deck, opt, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], b          # bit-exact counters
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=RTOL)  # per-iteration cost
        gp, op = cm.unpack_solution(prob.deck, sol[b]), cm.unpack_solution(prob.deck, osol)
        for pg, po in zip(gp, op):
            for name in ("Xbar", "Ubar", "Y", "K", "dU", "Qu", "Quu", "Qux", "G"):  # final trajectories and gains
                assert relerr(pg[name], po[name], FLOOR.get(name, 1e-6)) < RTOL, (b, name)


@pytest.mark.parametrize("k0", [0, 20])
def test_single_mhpc_solve_matches_oracle_and_golden(cm, opt, k0):
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV, k0=k0)
    x0 = workload.mhpc_batch(4)
    s = solve_gpu(cm, prob, opt, x0)
    compare_with_oracle(cm, prob, opt, x0, s, range(4))
    g = np.load(os.path.join(REPO, "tests/golden/mhpc_trot.npz"))
    info = s.get_solver_info()
    key = "k0" if k0 == 0 else "k20"
    for b in (0, 3):
        assert [info[b][k] for k in COUNTS] == list(g["%s_counts_%d" % (key, b)])
        assert abs(info[b]["cost"] - g["%s_final_%d" % (key, b)][0]) <= RTOL * abs(g["%s_final_%d" % (key, b)][0])
    if k0 == 20:
        assert all(i["outer_iter"] > 1 for i in info)  # the AL loop really ran


@pytest.mark.parametrize("k0", [0, 20])
def test_mhpc_one_iteration_per_knot_parity(cm, opt, k0):
    """Dynamics / cost / constraint partials (A, B, C, D, l**), reset-map Jacobian and every backward-sweep product per knot:
    codegen'd analytic derivatives on the GPU vs dual-number RNEA + reference CasADi partials + explicit KKT inverse in the oracle."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV, k0=k0)
    o1 = copy.copy(opt)
    o1.max_DDP_iter = 1; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    x0 = workload.mhpc_batch(4)
    s = solve_gpu(cm, prob, o1, x0)
    for b in (0, 2):
        oracle_solve(prob.deck, o1, x0[b])
        for ph in range(3):
            for name in ("X", "U", "Y", "Defect", "l", "lx", "lu", "ly", "lxx", "luu", "lyy", "A", "B", "C", "D", "Phix", "Phixx"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-11, (name, ph)
            for name in ("Quu", "Qux", "Qu", "K", "dU", "G", "dX"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-10, (name, ph)
        for ph in range(2):
            assert relerr(s.debug_get("Px", ph, b), oracle_get("Px", ph)) < 1e-11, ph


def test_mhpc_perturbed_batch_counts_bit_exact(cm, opt):
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV)
    x0 = workload.mhpc_batch(40)
    s = solve_gpu(cm, prob, opt, x0)
    compare_with_oracle(cm, prob, opt, x0, s, range(40))


def test_mhpc_batch_1024_properties(cm, opt):
    """BASELINE config 3: 1024 MHPC trot problems on one B200 — determinism, convergence flags, spot checks vs the oracle."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV)
    B = 1024
    x0 = workload.mhpc_batch(B)
    s = solve_gpu(cm, prob, opt, x0)
    i1 = s.get_solver_info(); c1 = s.get_commands(8)
    s.solve(opt)
    i2 = s.get_solver_info(); c2 = s.get_commands(8)
    assert np.array_equal(c1, c2) and i1 == i2
    assert all(i["status"] == 0 for i in i1)
    assert all(i["feas"] <= opt.dynamics_feas_thresh for i in i1)
    assert all(1 <= i["iter"] <= opt.max_AL_iter * opt.max_DDP_iter for i in i1)
    compare_with_oracle(cm, prob, opt, x0, s, (0, 1, 333, 512, 1023))
    # value-function property: the feedback gain reduces the quadratic model, Quu of every knot is positive definite
    sol = cm.unpack_solution(prob.deck, s.get_solution(5, 1)[0])
    for ph in sol:
        for Quu in ph["Quu"]:
            assert np.all(np.linalg.eigvalsh(0.5 * (Quu + Quu.T)) > 0)


@pytest.mark.parametrize("k0", [0, 20])
def test_mhpc_whole_problem_single_shooting_matches_oracle(cm, opt, k0):
    """MS = false on the cascaded-fidelity deck: the chain runs through whole-body phases, the touchdown impact (k0 = 20) and the
    whole-body -> single-rigid-body reset map; expected cost change from the sweep. GPU == oracle: counters bit-exact; costs and
    trajectories at 1e-7 / 1e-6 instead of 1e-9 - a 35-knot open-loop chain compounds the 1e-13 per-knot differences of the two
    whole-body models (measured 1.7e-9 on one history entry)."""
    from cafe_mpc_b200 import workload
    o = copy.copy(opt); o.MS = 0; o.max_AL_iter = 3; o.max_DDP_iter = 6
    prob = cm.MHPCProblem(CSV, k0=k0)
    x0 = workload.mhpc_batch(3)
    s = solve_gpu(cm, prob, o, x0)
    info = s.get_solver_info(); hist = s.get_history(64); sol = s.get_solution()
    for b in range(3):
        oi, oh, ot, osol = oracle_solve(prob.deck, o, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (b, info[b], oi)
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=1e-7)
        assert info[b]["feas"] == 0.0
        gp, op = cm.unpack_solution(prob.deck, sol[b]), cm.unpack_solution(prob.deck, osol)
        for pg, po in zip(gp, op):
            for name in ("Xbar", "Ubar", "Y", "K"):
                assert relerr(pg[name], po[name]) < 1e-6, (b, name)


def test_mhpc_headline_batch_4096_against_the_oracle(cm, opt):
    """The headline batch (BASELINE metric: 4096 MHPC trot problems on one GPU; the two-stream tick and the list compaction are active at
    this size only): 64 problems spread over the batch against the oracle (counters bit-exact, cost history / solution 1e-9), every
    problem converged, and the first 1024 problems equal to the same problems solved as a 1024 batch, bit for bit (a problem's result does
    not depend on the batch it is solved in)."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV)
    B = 4096
    x0 = workload.mhpc_batch(B)
    assert len(np.unique(x0.round(12), axis=0)) == B
    s = solve_gpu(cm, prob, opt, x0)
    info = s.get_solver_info()
    assert all(i["status"] == 0 and i["feas"] <= opt.dynamics_feas_thresh for i in info)
    compare_with_oracle(cm, prob, opt, x0, s, tuple(range(17, B, 64)))
    head = s.get_commands(8)[:1024].copy()
    s.close()
    s2 = solve_gpu(cm, prob, opt, x0[:1024])
    assert np.array_equal(head, s2.get_commands(8))
    assert [tuple(i[k] for k in COUNTS) for i in info[:1024]] == [tuple(i[k] for k in COUNTS) for i in s2.get_solver_info()]


# ---- BASELINE config 4: MHPC running barrel roll (multi-phase with impact jumps)
@pytest.mark.parametrize("k0", [0, 205])
def test_barrel_roll_matches_oracle_and_golden(cm, opt, k0):
    """k0 = 0: stance -> diagonal pair -> flight -> SRB; k0 = 205: mid-roll flight (22 knots) -> 4-foot landing impact with four
    touchdown constraints (AL loop, the reference's `segment<3>(i)` impulse scatter with several landing feet) -> stance -> SRB."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=k0)
    g = np.load(os.path.join(REPO, "tests/golden/mhpc_barrel.npz"))
    key = "k0" if k0 == 0 else "k205"
    assert np.array_equal(np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td] for p in prob.phases()]), g[key + "_phases"])
    x0 = workload.barrel_batch(prob, 4)
    assert np.array_equal(x0, g[key + "_x0"])
    s = solve_gpu(cm, prob, opt, x0)
    compare_with_oracle(cm, prob, opt, x0, s, range(3))
    info = s.get_solver_info(); hist = s.get_history(256); sol = s.get_solution()
    for b in (0,):
        assert [info[b][k] for k in COUNTS] == list(g["%s_counts_%d" % (key, b)])
        np.testing.assert_allclose(hist[b, :info[b]["n_hist"], 0], g["%s_hist_%d" % (key, b)][:, 0], rtol=RTOL)
        assert relerr(sol[b], g["%s_sol_%d" % (key, b)]) < RTOL
    if k0 == 205:
        assert info[0]["outer_iter"] > 1 and info[0]["max_tconstr"] < 0.005   # the landing constraints were enforced by the AL loop


def test_barrel_roll_impact_knot_parity(cm, opt):
    """One iteration at the impact-bearing offset: flight-phase (no contact) dynamics partials, the 4-foot impact map and its
    Jacobian Px (36x36), touchdown-constraint terms and the value jump Px^T H Px, per knot against the oracle."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=workload.BARREL_K0_IMPACT)
    o1 = copy.copy(opt)
    o1.max_DDP_iter = 1; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    x0 = workload.barrel_batch(prob, 4)
    s = solve_gpu(cm, prob, o1, x0)
    for b in (0, 2):
        oracle_solve(prob.deck, o1, x0[b])
        for ph in range(3):
            for name in ("X", "U", "Y", "Defect", "l", "lx", "lu", "ly", "lxx", "luu", "lyy", "A", "B", "C", "D", "Phix", "Phixx"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-11, (name, ph)
            for name in ("Quu", "Qux", "Qu", "K", "dU", "G", "dX"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-10, (name, ph)
        for ph in range(2):
            assert relerr(s.debug_get("Px", ph, b), oracle_get("Px", ph)) < 1e-11, ph


def test_barrel_roll_batch_256_properties(cm, opt):
    """256 perturbed mid-roll problems: deterministic, every problem ends with a legal status and iteration count, problems that
    converge satisfy the landing constraints; spot checks against the oracle."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=workload.BARREL_K0_IMPACT)
    B = 256
    x0 = workload.barrel_batch(prob, B)
    s = solve_gpu(cm, prob, opt, x0)
    i1 = s.get_solver_info(); c1 = s.get_commands(8)
    s.solve(opt)
    assert np.array_equal(c1, s.get_commands(8)) and i1 == s.get_solver_info()
    cap = opt.max_AL_iter * opt.max_DDP_iter
    assert all(1 <= i["iter"] <= cap for i in i1)
    done = [i for i in i1 if i["iter"] < cap and i["status"] == 0]
    assert len(done) > B // 2
    assert all(np.isfinite([i["cost"], i["feas"], i["max_tconstr"], i["max_pconstr"]]).all() for i in i1)
    compare_with_oracle(cm, prob, opt, x0, s, (0, 2, 3, 4))   # problem 3 runs into the 10 x 20 iteration cap (1863 line-search trials)


def test_lcm_command_record_is_the_float_cast_of_the_solution(cm, opt):
    """the per-problem MHPC_Command_lcmt fields packed on the device == cast<float>() of the double solution
    (MHPCLocomotion.cpp:236-281), for steps that cross the WB0 -> WB1 phase boundary."""
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV, k0=20)   # WB0 h=16, WB1 h=9
    x0 = workload.mhpc_batch(6)
    s = solve_gpu(cm, prob, opt, x0)
    N = 20
    rec = s.get_lcm_commands(N)
    assert rec.dtype == np.float32 and rec.shape == (6, 1080 * N)
    sol = s.get_solution()
    for b in (0, 5):
        ph = cm.unpack_solution(prob.deck, sol[b])
        cat = lambda name: np.concatenate([np.asarray(ph[0][name])[:16], np.asarray(ph[1][name])[:N - 16]])
        X, U, Y = cat("Xbar"), cat("Ubar"), cat("Y")
        f = cm.unpack_lcm_command(rec[b], N)
        exp = {"torque": U, "pos": X[:, 0:3], "eul": X[:, 3:6], "qJ": X[:, 6:18], "vWorld": X[:, 18:21], "eulrate": X[:, 21:24], "qJd": X[:, 24:36], "GRF": Y,
               "Qu": cat("Qu"), "feedback": np.stack([k.flatten(order="F") for k in cat("K")]), "Quu": np.stack([k.flatten(order="F") for k in cat("Quu")]),
               "Qux": np.stack([k.flatten(order="F") for k in cat("Qux")])}
        for name, v in exp.items():
            assert np.array_equal(f[name], v.astype(np.float32)), name
    with pytest.raises(Exception):
        s.get_lcm_commands(26)   # more steps than whole-body knots


def test_per_problem_references_match_oracle(cm, opt):
    """every problem tracks its own reference (different commanded forward speed) on the shared phase schedule;
    GPU == oracle run problem by problem on a deck carrying that problem's records. Impact-bearing offset so that the AL loop runs."""
    from cafe_mpc_b200 import workload
    from oracle_bindings import deck_with_references
    prob = cm.MHPCProblem(CSV, k0=20)
    B = 6
    x0 = workload.mhpc_batch(B)
    refs = workload.speed_command_references(prob, B)
    assert np.array_equal(refs[0], prob.reference_records()) and not np.array_equal(refs[3], refs[0])
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(x0)
    s.solve(opt)
    shared_info = s.get_solver_info()
    s.set_references(refs)
    s.solve(opt)
    info = s.get_solver_info(); hist = s.get_history(256); sol = s.get_solution()
    assert info[0] == shared_info[0]                 # problem 0 kept the deck's reference
    assert any(info[b]["cost"] != shared_info[b]["cost"] for b in range(1, B))
    for b in range(B):
        dk, keep = deck_with_references(prob.deck, refs[b])
        oi, oh, ot, osol = oracle_solve(dk, opt, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], b
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=RTOL)
        gp, op = cm.unpack_solution(prob.deck, sol[b]), cm.unpack_solution(prob.deck, osol)
        for pg, po in zip(gp, op):
            for name in ("Xbar", "Ubar", "Y", "K", "dU", "Qu", "Quu", "Qux", "G"):
                assert relerr(pg[name], po[name], FLOOR.get(name, 1e-6)) < RTOL, (b, name)
    # a reference set that changes the contact schedule is refused, and NULL restores the shared records
    bad = refs.copy(); bad[2, 5, 99] = 1.0 - bad[2, 5, 99]
    with pytest.raises(Exception):
        s.set_references(bad)
    s.set_references(None)
    s.solve(opt)
    assert s.get_solver_info() == shared_info


def test_receding_horizon_warm_start_matches_oracle(cm, opt):
    """SURVEY §8(f)1 (first slice): receding-horizon re-solves. The plan is shifted by dt_mpc / dt_wb = 2 knots per step (phase removal
    at k0 = 12, a new one-knot phase opened at the tail), the previous solution becomes the warm-start guess (cafe_mpc_b200/mpc.py),
    and the re-solve runs under the run-time iteration caps (max_AL_iter_runtime x max_DDP_iter_runtime). GPU == oracle at every step.
    The decks are marked as products of MHPCProblem::update: the one-knot tail phase opened at k0 = 12 has no shooting states in that
    step (MHPCProblem.cpp:366-369) and is integrated sequentially from the state the previous phase hands over."""
    from cafe_mpc_b200 import mpc, workload
    ort = copy.copy(opt)
    ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    B, k0 = 4, 8
    prob = cm.MHPCProblem(CSV, k0=k0)
    x0 = workload.mhpc_batch(B)
    s = solve_gpu(cm, prob, opt, x0)
    sol = s.get_solution()
    al = s.get_al_params()
    for step in range(3):
        k1 = k0 + 2
        p1 = cm.MHPCProblem(CSV, k0=k1, mpc_update_nsteps=2)
        assert p1.single_shooting_phase == (1 if k1 == 12 else -1)
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol)
        al1 = mpc.shift_al(prob, k0, p1, k1, al)      # the reference never resets sigma / lambda between MPC steps (ConstraintsBase.h:367-374)
        # the "measured" state of the next step: the plan's own prediction two knots ahead, nudged
        x1 = np.stack([mpc.state_at(prob, cm.unpack_solution(prob.deck, sol[b]), 2) for b in range(B)]) + 1e-3 * (x0 - x0[0])
        s1 = cm.MultiPhaseDDP(p1, 0, B)
        s1.set_initial_condition(x1)
        s1.set_initial_guess(guess)
        s1.set_al_params(al1)
        s1.solve(ort)
        info = s1.get_solver_info(); hist = s1.get_history(256); sol1 = s1.get_solution(); al = s1.get_al_params()
        for b in range(B):
            oi, oh, ot, osol, oal = oracle_solve(p1.deck, ort, x1[b], guess=guess[b], al=al1[b])
            assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (step, b)
            np.testing.assert_allclose(al[b], oal, rtol=RTOL, atol=1e-12)
            np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=RTOL)
            gp, op = cm.unpack_solution(p1.deck, sol1[b]), cm.unpack_solution(p1.deck, osol)
            for pg, po in zip(gp, op):
                for name in ("Xbar", "Ubar", "Y", "K", "dU", "Qu", "Quu", "Qux", "G"):
                    assert relerr(pg[name], po[name], FLOOR.get(name, 1e-6)) < RTOL, (step, b, name)
        if k1 == 12:
            # no shooting states in the tail phase: its defects vanish identically, and the flag is live (an all-shooting deck at the
            # same offset ends elsewhere)
            for b in range(B):
                assert not np.any(s1.debug_get("Defect", 1, b))
            p1ms = cm.MHPCProblem(CSV, k0=k1)
            s2 = cm.MultiPhaseDDP(p1ms, 0, B)
            s2.set_initial_condition(x1); s2.set_initial_guess(guess); s2.set_al_params(al1); s2.solve(ort)
            assert all(a["cost"] != c["cost"] for a, c in zip(info, s2.get_solver_info()))
            s2.close()
        # the warm start pays: same caps from the cold start end far from feasible
        s1.set_initial_guess(None)
        s1.solve(ort)
        cold = s1.get_solver_info()
        assert all(info[b]["feas"] < 0.2 * cold[b]["feas"] for b in range(B))
        prob, k0, sol = p1, k1, sol1


@pytest.mark.gpu
def test_device_shift_equals_host_shift(cm, opt):
    """the receding-horizon shift done on the device (cafe_gpu_shift_guess: previous solver's arrays -> packed guess of the
    next deck) is the same data movement as cafe_mpc_b200/mpc.py on the host: identical warm-started solves, bit for bit, over a chain
    that removes a phase at the front and opens one at the tail; cafe_gpu_get_planned_state == mpc.state_at."""
    from cafe_mpc_b200 import mpc, workload
    ort = copy.copy(opt)
    ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    B, k0 = 6, 8
    prob = cm.MHPCProblem(CSV, k0=k0)
    x0 = workload.mhpc_batch(B)
    s = solve_gpu(cm, prob, opt, x0)
    for step in range(3):
        k1 = k0 + 2
        p1 = cm.MHPCProblem(CSV, k0=k1)
        sol = s.get_solution()
        x1_host = mpc.state_at(prob, mpc.unpack_batch(prob, sol), 2)
        x1 = s.planned_state(2)
        assert np.array_equal(x1, x1_host), step
        x1 = x1 + 1e-3 * (x0 - x0[0])
        # host path
        sh = cm.MultiPhaseDDP(p1, 0, B)
        sh.set_initial_condition(x1)
        sh.set_initial_guess(mpc.shifted_guess_batch(prob, k0, p1, k1, sol))
        sh.set_al_params(mpc.shift_al(prob, k0, p1, k1, s.get_al_params()))
        sh.solve(ort)
        # device path
        sd = cm.MultiPhaseDDP(p1, 0, B)
        sd.set_initial_condition(x1)
        sd.shift_guess_from(s, k0, k1)
        sd.solve(ort)
        ih, idv = sh.get_solver_info(), sd.get_solver_info()
        assert [[i[k] for k in COUNTS] for i in ih] == [[i[k] for k in COUNTS] for i in idv], step
        assert np.array_equal(sh.get_solution(), sd.get_solution()), step
        sh.close(); s.close()
        s, prob, k0 = sd, p1, k1


def test_update_deck_on_one_handle_equals_fresh_handles(cm, opt):
    """cafe_gpu_update_deck: the MPC update on ONE solver (deck replaced, previous solution shifted into the warm start, device arena re-used)
    gives bit for bit the solves of the two-handle path (a fresh solver per MPC step + cafe_gpu_shift_guess) over eight consecutive updates
    on marked decks - start offsets 8 .. 22: the front phase shrinks to one knot (10) and disappears (12), a single-shooting tail opens
    (12) and grows - and the float32 wire records agree as well."""
    from cafe_mpc_b200 import workload
    ort = copy.copy(opt)
    ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    B, k0 = 5, 6
    prob = cm.MHPCProblem(CSV, k0=k0)
    x0 = workload.mhpc_batch(B)
    one = solve_gpu(cm, prob, opt, x0)          # the solver that is updated in place
    two = solve_gpu(cm, prob, opt, x0)          # the chain of fresh solvers
    shapes = set()
    for step in range(8):
        k1 = k0 + 2
        p1 = cm.MHPCProblem(CSV, k0=k1, mpc_update_nsteps=2)
        shapes.add(tuple(p.horizon for p in p1.phases()))
        x1 = two.planned_state(2)
        assert np.array_equal(x1, one.planned_state(2)), step
        x1 = x1 + 1e-3 * (x0 - x0[0])
        nxt = cm.MultiPhaseDDP(p1, 0, B)
        nxt.set_initial_condition(x1)
        nxt.shift_guess_from(two, k0, k1)
        nxt.solve(ort)
        one.update_deck(p1, 2)
        one.set_initial_condition(x1)
        one.solve(ort)
        ia, ib = nxt.get_solver_info(), one.get_solver_info()
        assert [[i[k] for k in COUNTS] for i in ia] == [[i[k] for k in COUNTS] for i in ib], step
        assert np.array_equal(nxt.get_solution(), one.get_solution()), step
        assert np.array_equal(nxt.get_lcm_commands(8), one.get_lcm_commands(8)), step
        two.close()
        two, prob, k0 = nxt, p1, k1
    assert (1, 24, 10) in shapes and (24, 1, 10) in shapes and len(shapes) == 8
    # cold start on a new deck through the same entry: equals a fresh solver
    p2 = cm.MHPCProblem(CSV, k0=0)
    one.update_deck(p2, 0, B=0)
    one.set_initial_condition(x0)
    one.solve(opt)
    fresh = solve_gpu(cm, p2, opt, x0)
    assert np.array_equal(fresh.get_solution(), one.get_solution())


# ---- next tier: LocoProblem, whole-body-only locomotion TO (9 WB phases, 100 knots, three flight -> stance impacts)
def test_loco_problem_matches_oracle_and_golden(cm):
    from cafe_mpc_b200 import workload
    prob = cm.LocoProblem()
    lopt = cm.load_hsddp_setting(workload.LOCO_DDP_SETTING)
    g = np.load(os.path.join(REPO, "tests/golden/loco_flypace.npz"))
    x0 = workload.mhpc_batch(4)
    assert np.array_equal(x0, g["x0"])
    s = solve_gpu(cm, prob, lopt, x0)
    compare_with_oracle(cm, prob, lopt, x0, s, (0, 1))
    info = s.get_solver_info(); hist = s.get_history(256); sol = s.get_solution()
    for b in (0, 3):
        assert [info[b][k] for k in COUNTS] == list(g["counts_%d" % b])
        np.testing.assert_allclose(hist[b, :info[b]["n_hist"], 0], g["hist_%d" % b][:, 0], rtol=RTOL)
        assert relerr(sol[b], g["sol_%d" % b]) < RTOL
    assert all(i["outer_iter"] > 1 for i in info)


def test_mhpc_deck_without_joint_and_height_barriers_matches_oracle(cm, opt):
    """The constraint-set flags on an MHPC deck (WB + SRB): GPU == oracle with the two barriers dropped, and != the full deck."""
    import ctypes as C
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200._ctypes_defs import Deck

    class _P:
        pass
    base = cm.MHPCProblem(CSV, k0=20)
    d2 = Deck.from_buffer_copy(base.deck.contents)
    for i in range(d2.n_phases):
        d2.phase[i].no_joint_limit = 1; d2.phase[i].no_min_height = 1
    p2 = _P(); p2.deck = C.pointer(d2); p2._keep = base
    x0 = workload.mhpc_batch(4)
    s = solve_gpu(cm, p2, opt, x0)
    compare_with_oracle(cm, p2, opt, x0, s, (0, 2))
    s0 = solve_gpu(cm, base, opt, x0)
    assert s0.get_solver_info()[2]["cost"] != s.get_solver_info()[2]["cost"]


# ---- next tier: in-place barrel roll (BarrelRollTO.cpp): joint-speed barrier, per-phase weights, fixed desired
#      states, two four-foot landings, solve started from the interpolated state trajectory
def test_barrel_to_matches_oracle_and_golden(cm):
    """The solve starts from a wildly infeasible interpolated trajectory (defect norm 57, cost 3 635) and differences of 1e-11 in the
    first sweep grow by a factor 3-5 per iteration while the roll is being found (1e-6 at worst around iteration 30) before they
    collapse again at convergence: every DECISION (iteration, line-search, regularisation and AL counters) still matches the oracle bit
    for bit, the first iteration matches per knot at 1e-11, the final cost at the full 30 x 10 caps at 1e-8."""
    from cafe_mpc_b200 import workload
    prob = cm.BarrelRollProblem()
    fopt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
    g = np.load(os.path.join(REPO, "tests/golden/barrel_to.npz"))
    x0 = workload.mhpc_batch(4)
    guess = prob.initial_guess(x0)
    s = cm.MultiPhaseDDP(prob, 0, 4)
    s.set_initial_condition(x0)
    s.set_initial_guess(guess)
    # (1) first iteration, per knot: LQ data incl. the joint-speed barrier terms, the two four-foot impact maps, every sweep product
    o1 = copy.copy(fopt)
    o1.max_DDP_iter = 1; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    s.solve(o1)
    for b in (0, 2):
        oracle_solve(prob.deck, o1, x0[b], guess=guess[b])
        for ph in range(6):
            for name in ("X", "U", "Y", "Defect", "l", "lx", "lu", "ly", "lxx", "luu", "lyy", "A", "B", "C", "D", "Phix", "Phixx"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-11, (name, ph)
            for name in ("Quu", "Qux", "Qu", "K", "dU", "G", "dX"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-10, (name, ph)
        for ph in range(5):
            assert relerr(s.debug_get("Px", ph, b), oracle_get("Px", ph)) < 1e-11, ph
    # (2) 5 x 10 caps against the committed golden and the oracle: counters bit-exact, costs within the amplification described above
    bopt = copy.copy(fopt)
    bopt.max_AL_iter = 5
    s.solve(bopt)
    info = s.get_solver_info(); hist = s.get_history(320); sol = s.get_solution()
    for b in (0, 3):
        assert [info[b][k] for k in COUNTS] == list(g["counts_%d" % b]), (b, info[b])
        np.testing.assert_allclose(hist[b, :3, 0], g["hist_%d" % b][:3, 0], rtol=2e-9)
        np.testing.assert_allclose(hist[b, :info[b]["n_hist"], 0], g["hist_%d" % b][:, 0], rtol=1e-5)
        gp, op = cm.unpack_solution(prob.deck, sol[b]), cm.unpack_solution(prob.deck, g["sol_%d" % b])
        for pg, po in zip(gp, op):
            for name in ("Xbar", "Ubar", "Y"):
                assert relerr(pg[name], po[name]) < 1e-4, (b, name)
    oi, oh, ot, osol = oracle_solve(prob.deck, bopt, x0[1], cap=320, guess=guess[1])
    assert [info[1][k] for k in COUNTS] == [oi[k] for k in COUNTS]
    # (3) full caps, problem 0: 300 iterations / ~2 100 line-search trials decided identically, same final cost
    s.solve(fopt)
    i0 = s.get_solver_info()[0]
    oi, oh, ot, osol = oracle_solve(prob.deck, fopt, x0[0], cap=320, guess=guess[0])
    assert [i0[k] for k in COUNTS] == [oi[k] for k in COUNTS]
    assert abs(i0["cost"] - oi["cost"]) < 1e-8 * abs(oi["cost"]) and abs(i0["max_tconstr"] - oi["max_tconstr"]) < 1e-6 * abs(oi["max_tconstr"])
    assert oi["cost"] < 0.01 * oh[0, 0] and oi["max_tconstr"] < fopt.tconstr_thresh   # the roll was found, both landings enforced


def test_barrel_to_32_problems_at_full_caps_against_four_roundings_of_the_oracle(cm):
    """All 32 problems of the in-place barrel roll at the full 30 x 10 caps (~270 iterations, ~2 000 line-search trials each) against the
    committed counters of the CPU oracle in FOUR roundings of one algorithm (tests/golden/barrel_to_four_roundings.json, made by
    tools/oracle_sensitivity.py): the oracle's sources with / without FMA contraction x the reference's CasADi kinematic-partial file
    compiled -O1 / -O3. The four builds agree with each other on 28 problems and part ways on four (11, 15, 27, 30: a last-bit difference in
    the first sweep, amplified 3-5 x per iteration while the roll is being found, flips one Armijo test): "decisions bit-exact" is well posed
    on the 28, and there the GPU must reproduce every counter, and the final cost to 1e-6
    (the four oracle builds themselves spread by up to 4e-8 there); on the four it must land on one of the outcomes
    the oracle builds produce."""
    import json
    from cafe_mpc_b200 import workload
    prob = cm.BarrelRollProblem()
    fopt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
    g = json.load(open(os.path.join(REPO, "tests/golden/barrel_to_four_roundings.json")))["builds"]
    x0 = workload.mhpc_batch(32)
    s = cm.MultiPhaseDDP(prob, 0, 32)
    s.set_initial_condition(x0)
    s.set_initial_guess(prob.initial_guess(x0))
    s.solve(fopt)
    info = s.get_solver_info()
    split = [b for b in range(32) if len(set(tuple(col[b][0]) for col in g.values())) > 1]
    assert split == [11, 15, 27, 30]
    bad = []
    for b in range(32):
        got = [info[b][k] for k in COUNTS]
        ref = g["fma_kinO3"][b]
        if b in split:   # one of the outcomes the oracle builds produce, final cost within their spread
            assert got in [col[b][0] for col in g.values()], b
            assert min(abs(info[b]["cost"] - col[b][1]) for col in g.values()) < 1e-5 * abs(ref[1]), b
        elif got != ref[0] or abs(info[b]["cost"] - ref[1]) > 1e-6 * abs(ref[1]):
            bad.append((b, got, ref[0], info[b]["cost"], ref[1]))
    assert not bad, bad


def test_single_shooting_first_phase_matches_oracle(cm, opt):
    """CafePhase::single_shooting on the FIRST phase (integrated from the solver's x0; every knot thread repeats the chain of up to 11
    whole-body steps) and the create-time checks on the flag."""
    import ctypes as C
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200._ctypes_defs import Deck
    from cafe_mpc_b200.lib import CafeError

    class _P:
        pass
    base = cm.MHPCProblem(CSV)
    d2 = Deck.from_buffer_copy(base.deck.contents)
    d2.phase[0].single_shooting = 1
    p2 = _P(); p2.deck = C.pointer(d2); p2._keep = base
    x0 = workload.mhpc_batch(3)
    s = solve_gpu(cm, p2, opt, x0)
    compare_with_oracle(cm, p2, opt, x0, s, (0, 2))
    for b in range(3):
        assert not np.any(s.debug_get("Defect", 0, b))   # no shooting states: the defects of that phase vanish identically
    assert s.get_solver_info()[1]["cost"] != solve_gpu(cm, base, opt, x0).get_solver_info()[1]["cost"]
    d3 = Deck.from_buffer_copy(d2)
    d3.phase[1].single_shooting = 1          # two phases in a row without shooting states: not supported
    p3 = _P(); p3.deck = C.pointer(d3)
    with pytest.raises(CafeError):
        cm.MultiPhaseDDP(p3, 0, 1)
    d4 = Deck.from_buffer_copy(base.deck.contents)
    d4.phase[2].single_shooting = 1          # SRB after WB: the hand-over needs the same model on both sides
    p4 = _P(); p4.deck = C.pointer(d4)
    with pytest.raises(CafeError):
        cm.MultiPhaseDDP(p4, 0, 1)
