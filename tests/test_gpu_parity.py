"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs, against the committed golden fixtures, and through size-independent properties at full batch size.

Tolerances (BASELINE.json north_star): contact/phase schedules and iteration counts bit-exact; per-iteration
cost, gains and final trajectories within 1e-9 relative (relative to the largest entry of each array)."""
import os

import numpy as np
import pytest

from oracle_bindings import oracle_get, oracle_solve

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")
RTOL = 1e-9


def relerr(g, o):
    o = np.asarray(o); g = np.asarray(g)
    if o.size == 0:
        return 0.0
    return float(np.max(np.abs(g - o)) / max(np.max(np.abs(o)), 1e-300))


def solve_gpu(cm, prob, opt, x0):
    s = cm.MultiPhaseDDP(prob, 0, len(x0))
    s.set_initial_condition(x0)
    s.solve(opt)
    return s


def test_single_hkd_solve_matches_oracle_and_golden(cm, hkd_problem, hkd_options):
    """BASELINE config 1: the single HKD trot solve of HKDMPCSolver::initialize."""
    from cafe_mpc_b200 import workload
    x0 = workload.hkd_batch(hkd_problem, 1)
    s = solve_gpu(cm, hkd_problem, hkd_options, x0)
    gi = s.get_solver_info()[0]
    oi, oh, ot, osol = oracle_solve(hkd_problem.deck, hkd_options, x0[0])
    assert [gi[k] for k in COUNTS] == [oi[k] for k in COUNTS]
    gh = s.get_history(256)[0, :gi["n_hist"]]
    np.testing.assert_allclose(gh[:, 0], oh[:, 0], rtol=RTOL)          # per-iteration cost
    np.testing.assert_allclose(gh[:, 1:], oh[:, 1:], rtol=RTOL, atol=1e-13)  # feasibility / violations (values ~1e-5)
    gt = s.get_trace(256)[0, :gi["iter"]]
    assert np.array_equal(gt[:, 6:10], ot[:, 6:10])                     # reg iters, ls iters, success, accepted eps: exact
    np.testing.assert_allclose(gt[:, 2:4], ot[:, 2:4], rtol=RTOL, atol=1e-12)       # expected cost change
    gsol = s.get_solution()[0]
    gp, op = cm.unpack_solution(hkd_problem.deck, gsol), cm.unpack_solution(hkd_problem.deck, osol)
    for pg, po in zip(gp, op):
        for name in ("Xbar", "Ubar", "K", "dU", "Quu", "Qux", "G"):
            assert relerr(pg[name], po[name]) < RTOL, name
    g = np.load(os.path.join(REPO, "tests/golden/hkd_trot_nominal.npz"))
    assert [gi[k] for k in COUNTS] == list(g["counts"])
    np.testing.assert_allclose(gh[:, 0], g["hist"][:, 0], rtol=RTOL)
    assert abs(gi["cost"] - g["final"][0]) <= RTOL * abs(g["final"][0])


def test_one_iteration_per_knot_parity(cm, hkd_problem, hkd_options):
    """LQ data and every backward-sweep / linear-rollout product of the first iteration, per knot."""
    import copy
    from cafe_mpc_b200 import workload
    opt = copy.copy(hkd_options)
    opt.max_DDP_iter = 1; opt.max_AL_iter = 1; opt.cost_thresh = 1e30; opt.dynamics_feas_thresh = 1e30  # stop before the line search
    x0 = workload.hkd_batch(hkd_problem, 4)
    s = solve_gpu(cm, hkd_problem, opt, x0)
    for b in (0, 3):
        oracle_solve(hkd_problem.deck, opt, x0[b])
        for ph in range(3):
            for name in ("X", "U", "Defect", "l", "lx", "lu", "lxx", "luu", "A", "B", "Phix", "Phixx"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-13, (name, ph)
            for name in ("Quu", "Qux", "Qu", "K", "dU", "G", "dX"):
                assert relerr(s.debug_get(name, ph, b), oracle_get(name, ph)) < 1e-11, (name, ph)
        for ph in range(2):
            assert relerr(s.debug_get("Px", ph, b), oracle_get("Px", ph)) < 1e-14


def test_perturbed_batch_counts_bit_exact(cm, hkd_problem, hkd_options):
    from cafe_mpc_b200 import workload
    B = 48
    x0 = workload.hkd_batch(hkd_problem, B)
    s = solve_gpu(cm, hkd_problem, hkd_options, x0)
    info = s.get_solver_info()
    hist = s.get_history(256)
    sol = s.get_solution()
    for b in range(B):
        oi, oh, ot, osol = oracle_solve(hkd_problem.deck, hkd_options, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], b
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=RTOL)
        gp, op = cm.unpack_solution(hkd_problem.deck, sol[b]), cm.unpack_solution(hkd_problem.deck, osol)
        for pg, po in zip(gp, op):
            assert relerr(pg["Xbar"], po["Xbar"]) < RTOL and relerr(pg["Ubar"], po["Ubar"]) < RTOL
    g = np.load(os.path.join(REPO, "tests/golden/hkd_trot_batch8.npz"))
    for b in range(8):
        assert [info[b][k] for k in COUNTS] == [int(v) for v in g["rows"][b, :6]]
        assert abs(info[b]["cost"] - g["rows"][b, 6]) <= RTOL * abs(g["rows"][b, 6])


@pytest.mark.parametrize("B", [1, 33])
def test_ragged_batch_sizes_and_batch_position_invariance(cm, hkd_problem, hkd_options, B):
    """A problem's result must not depend on the batch size or on its slot in the batch."""
    from cafe_mpc_b200 import workload
    x0 = workload.hkd_batch(hkd_problem, 40)
    ref = solve_gpu(cm, hkd_problem, hkd_options, x0[:40])
    rs = ref.get_solution()
    perm = np.arange(40)[::-1][:B].copy()
    s = solve_gpu(cm, hkd_problem, hkd_options, x0[perm])
    ss = s.get_solution()
    ri, si = ref.get_solver_info(), s.get_solver_info()
    for j, b in enumerate(perm):
        assert [si[j][k] for k in COUNTS] == [ri[b][k] for k in COUNTS]
        assert np.array_equal(ss[j], rs[b])  # bitwise: same instruction stream per problem


def test_outer_iteration_cap_and_history_edges(cm, hkd_problem, hkd_options):
    import copy
    from cafe_mpc_b200 import workload
    x0 = workload.hkd_batch(hkd_problem, 4)
    for al, ddp in ((1, 1), (1, 3), (2, 2)):
        opt = copy.copy(hkd_options); opt.max_AL_iter = al; opt.max_DDP_iter = ddp
        s = solve_gpu(cm, hkd_problem, opt, x0)
        info = s.get_solver_info()
        for b in range(4):
            oi, oh, _, _ = oracle_solve(hkd_problem.deck, opt, x0[b])
            assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS]
            assert info[b]["iter"] <= al * ddp


def test_full_size_batch_properties(cm, hkd_problem, hkd_options):
    """BASELINE-size batch (1024): determinism, convergence flags, and the value-function identity
    dV = -Qu^T dU >= 0 that every accepted backward sweep must satisfy."""
    from cafe_mpc_b200 import workload
    B = 1024
    x0 = workload.hkd_batch(hkd_problem, B)
    s = solve_gpu(cm, hkd_problem, hkd_options, x0)
    i1 = s.get_solver_info()
    c1 = s.get_commands(8)
    s.solve(hkd_options)
    i2 = s.get_solver_info()
    c2 = s.get_commands(8)
    assert np.array_equal(c1, c2) and i1 == i2  # run-to-run bitwise determinism
    st = np.array([i["status"] for i in i1])
    assert np.all(st == 0)
    feas = np.array([i["feas"] for i in i1])
    it = np.array([i["iter"] for i in i1])
    assert np.all(feas <= hkd_options.dynamics_feas_thresh) or np.all(it == hkd_options.max_AL_iter * hkd_options.max_DDP_iter)
    assert np.all(it >= 1) and np.all(it <= 50)
    # spot-check 6 problems of the big batch against the oracle
    for b in (0, 1, 511, 512, 1000, 1023):
        oi, _, _, _ = oracle_solve(hkd_problem.deck, hkd_options, x0[b])
        assert [i1[b][k] for k in COUNTS] == [oi[k] for k in COUNTS]
        assert abs(i1[b]["cost"] - oi["cost"]) <= RTOL * abs(oi["cost"])


def test_unsupported_options_fail_loudly(cm, hkd_problem, hkd_options):
    import copy
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200.lib import CafeError
    x0 = workload.hkd_batch(hkd_problem, 2)
    s = cm.MultiPhaseDDP(hkd_problem, 0, 2)
    s.set_initial_condition(x0)
    opt = copy.copy(hkd_options); opt.max_AL_iter = 400
    with pytest.raises(CafeError):   # more outer iterations than the history / barrier-update counters hold
        s.solve(opt)
    s.set_initial_condition(workload.hkd_batch(hkd_problem, 3))
    with pytest.raises(CafeError):  # batch larger than the handle's capacity
        s.solve(hkd_options)


def test_whole_problem_single_shooting_matches_oracle(cm, hkd_problem, hkd_options):
    """HSDDP_OPTION::MS = false (MultiPhaseDDP.cpp:65-68, :330-333; SinglePhase.cpp:187-221): no shooting states anywhere - every phase is
    integrated from the state the previous one hands over, defects vanish, there is no linear rollout and the expected cost change comes
    from the backward sweep (SinglePhase.cpp:383-387). GPU == oracle: counters bit-exact, history / solution 1e-9; feasibility exactly 0."""
    import copy
    from cafe_mpc_b200 import workload
    opt = copy.copy(hkd_options); opt.MS = 0
    x0 = workload.hkd_batch(hkd_problem, 4)
    s = solve_gpu(cm, hkd_problem, opt, x0)
    info = s.get_solver_info(); hist = s.get_history(64); sol = s.get_solution()
    for b in range(4):
        oi, oh, ot, osol = oracle_solve(hkd_problem.deck, opt, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (b, info[b], oi)
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=1e-9)
        assert np.all(hist[b, :oi["n_hist"], 1] == 0.0) and info[b]["feas"] == 0.0
        assert np.abs(sol[b] - osol).max() <= 1e-9 * np.abs(osol).max(), b


def test_hkd_receding_horizon_chain_matches_oracle(cm, hkd_options):
    """SURVEY §8(f)1 for the HKD application: HKDProblem::update (HKDProblem.cpp:117-222) shifts the plan by nsteps_between_mpc = 2 knots,
    HKDMPCSolver::update re-solves it under the caps 2 x 1 (HKDMPC.cpp:102-103) from the shifted previous solution. Six steps from the
    start of the trot reference: a one-knot tail phase opens at offset 2 (no shooting states in that step, HKDProblem.cpp:213-217), the
    front phase disappears at offset 12; Ubar[0] of the front phase is zeroed every step (:220). GPU == oracle at every step."""
    import copy
    from cafe_mpc_b200 import mpc, workload
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    B, k0 = 3, 0
    prob = cm.HKDProblem(csv, k0=k0)
    x0 = workload.hkd_batch(prob, B)
    s = solve_gpu(cm, prob, hkd_options, x0)
    sol = s.get_solution()
    al = s.get_al_params()
    seen_ss = seen_removal = False
    for step in range(6):
        k1 = k0 + 2
        p1 = cm.HKDProblem(csv, k0=k1, mpc_update=True)
        seen_ss = seen_ss or p1.single_shooting_phase >= 0
        seen_removal = seen_removal or len(p1.phases()) < len(prob.phases())
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol)
        al1 = mpc.shift_al(prob, k0, p1, k1, al)      # the reference never resets sigma / lambda between MPC steps (ConstraintsBase.h:367-374)
        assert not mpc.unpack_batch(p1, guess)[0]["Ubar"][:, 0].any()
        x1 = np.stack([mpc.state_at(prob, cm.unpack_solution(prob.deck, sol[b]), 2) for b in range(B)])
        x1[:, 3:6] += 1e-3 * (x0[:, 3:6] - x0[0, 3:6])   # the "measured" state: the plan's own prediction, position nudged
        s1 = cm.MultiPhaseDDP(p1, 0, B)
        s1.set_initial_condition(x1)
        s1.set_initial_guess(guess)
        s1.set_al_params(al1)
        s1.solve(ort)
        info = s1.get_solver_info(); hist = s1.get_history(64); sol1 = s1.get_solution(); al = s1.get_al_params()
        for b in range(B):
            oi, oh, ot, osol, oal = oracle_solve(p1.deck, ort, x1[b], guess=guess[b], al=al1[b])
            assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (step, b)
            np.testing.assert_allclose(al[b], oal, rtol=RTOL, atol=1e-12)
            np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=RTOL, atol=1e-9)
            gp, op = cm.unpack_solution(p1.deck, sol1[b]), cm.unpack_solution(p1.deck, osol)
            for pg, po in zip(gp, op):
                for name in ("Xbar", "Ubar", "K", "Quu", "Qux"):
                    assert relerr(pg[name], po[name]) < RTOL, (step, b, name)
        if p1.single_shooting_phase >= 0:
            for b in range(B):
                assert not np.any(s1.debug_get("Defect", p1.single_shooting_phase, b))
        # the warm start pays: the same caps from the cold start end far from feasible
        s1.set_initial_guess(None)
        s1.solve(ort)
        cold = s1.get_solver_info()
        assert all(info[b]["feas"] < 0.05 * cold[b]["feas"] for b in range(B))
        s1.close()
        prob, k0, sol = p1, k1, sol1
    assert seen_ss and seen_removal


def test_hkd_device_shift_equals_host_shift(cm, hkd_options):
    """cafe_gpu_shift_guess on HKD decks (HKDProblem::update incl. the Ubar[0] = 0 quirk) == cafe_mpc_b200/mpc.py on the host, bit for
    bit, over a chain that opens a tail phase (offset 2) — and cafe_gpu_get_planned_state == mpc.state_at."""
    import copy
    from cafe_mpc_b200 import mpc, workload
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    B, k0 = 5, 0
    prob = cm.HKDProblem(csv, k0=k0)
    x0 = workload.hkd_batch(prob, B)
    s = solve_gpu(cm, prob, hkd_options, x0)
    for step in range(3):
        k1 = k0 + 2
        p1 = cm.HKDProblem(csv, k0=k1, mpc_update=True)
        sol = s.get_solution()
        x1 = s.planned_state(2)
        assert np.array_equal(x1, mpc.state_at(prob, mpc.unpack_batch(prob, sol), 2)), step
        sh = cm.MultiPhaseDDP(p1, 0, B)
        sh.set_initial_condition(x1)
        sh.set_initial_guess(mpc.shifted_guess_batch(prob, k0, p1, k1, sol))
        sh.set_al_params(mpc.shift_al(prob, k0, p1, k1, s.get_al_params()))
        sh.solve(ort)
        sd = cm.MultiPhaseDDP(p1, 0, B)
        sd.set_initial_condition(x1)
        sd.shift_guess_from(s, k0, k1)
        sd.solve(ort)
        ih, idv = sh.get_solver_info(), sd.get_solver_info()
        assert [[i[k] for k in COUNTS] for i in ih] == [[i[k] for k in COUNTS] for i in idv], step
        assert np.array_equal(sh.get_solution(), sd.get_solution()), step
        sh.close(); s.close()
        s, prob, k0 = sd, p1, k1


def test_hkd_update_deck_on_one_handle_equals_fresh_handles(cm, hkd_options):
    """cafe_gpu_update_deck on HKD decks: seven consecutive MPC updates on one solver (tail phase opened at offset 2, front phase removed at 12)
    equal the chain of fresh solvers + cafe_gpu_shift_guess bit for bit, wire records included."""
    import copy
    from cafe_mpc_b200 import workload
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    B, k0 = 5, 0
    prob = cm.HKDProblem(csv, k0=k0)
    x0 = workload.hkd_batch(prob, B)
    one = solve_gpu(cm, prob, hkd_options, x0)
    two = solve_gpu(cm, prob, hkd_options, x0)
    for step in range(7):
        k1 = k0 + 2
        p1 = cm.HKDProblem(csv, k0=k1, mpc_update=True)
        x1 = two.planned_state(2)
        assert np.array_equal(x1, one.planned_state(2)), step
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
        assert np.array_equal(nxt.get_hkd_lcm_commands(9), one.get_hkd_lcm_commands(9)), step
        two.close()
        two, prob, k0 = nxt, p1, k1


def test_hkd_lcm_command_packing(cm, hkd_options):
    """cafe_gpu_get_hkd_lcm_commands == the loops of HKDMPCSolver::publish_mpc_cmd (HKDMPC.cpp:243-290) applied to the packed solution:
    float32 casts of Ubar, Xbar[:12] and K(m, n) for m, n < 12, nine steps that cross the first phase boundary (start offset 8: h = 3)."""
    from cafe_mpc_b200 import workload
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    prob = cm.HKDProblem(csv, k0=8)
    assert prob.phases()[0].horizon == 3
    B, N = 5, 9
    x0 = workload.hkd_batch(prob, B)
    s = solve_gpu(cm, prob, hkd_options, x0)
    wire = s.get_hkd_lcm_commands(N)
    assert wire.dtype == np.float32 and wire.shape == (B, 180 * N)
    sol = s.get_solution()
    for b in range(B):
        phases = cm.unpack_solution(prob.deck, sol[b])
        U = np.concatenate([p["Ubar"] for p in phases])[:N]
        X = np.concatenate([p["Xbar"][:-1] for p in phases])[:N]
        K = np.concatenate([p["K"] for p in phases])[:N]            # [k, m, n]
        assert np.array_equal(wire[b, :24 * N].reshape(N, 24), U.astype(np.float32))
        assert np.array_equal(wire[b, 24 * N:36 * N].reshape(N, 12), X[:, :12].astype(np.float32))
        assert np.array_equal(wire[b, 36 * N:].reshape(N, 12, 12), K[:, :12, :12].astype(np.float32))
        assert np.abs(K[:, :12, :12]).max() > 0
    from cafe_mpc_b200.lib import CafeError
    with pytest.raises(CafeError):
        s.get_hkd_lcm_commands(1000)


def test_asynchronous_collection_equals_the_blocking_one(cm, hkd_problem, hkd_options):
    """cafe_gpu_get_commands_async / cafe_gpu_commands_wait (pack on the solver's stream, D2H on a copy stream, two slots) deliver the bytes of
    cafe_gpu_get_commands - also when the next solve, on other initial states, is issued before the records of the previous one are awaited."""
    import torch
    from cafe_mpc_b200 import workload
    B = 48
    x0 = workload.hkd_batch(hkd_problem, 2 * B)
    s = cm.MultiPhaseDDP(hkd_problem, 0, B)
    want = []
    for part in (x0[:B], x0[B:]):
        s.set_initial_condition(part); s.solve(hkd_options)
        want.append(s.get_commands(8).copy())
    assert not np.array_equal(want[0], want[1])
    pin = [torch.zeros(want[0].shape, dtype=torch.float64).pin_memory() for _ in range(2)]
    s.set_initial_condition(x0[:B]); s.solve(hkd_options)
    s.get_commands_async(8, pin[0].numpy(), 0)
    s.set_initial_condition(x0[B:]); s.solve(hkd_options)      # overwrites the solver's arrays: the pack of slot 0 is ordered before it
    s.get_commands_async(8, pin[1].numpy(), 1)
    s.commands_wait(0); s.commands_wait(1)
    assert np.array_equal(pin[0].numpy(), want[0]) and np.array_equal(pin[1].numpy(), want[1])
    # a slot can be re-used: the new pack waits for the slot's previous transfer
    s.set_initial_condition(x0[:B]); s.solve(hkd_options)
    s.get_commands_async(8, pin[1].numpy(), 1)
    s.commands_wait(1)
    assert np.array_equal(pin[1].numpy(), want[0])
