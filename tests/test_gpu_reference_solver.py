"""The CUDA path against the REFERENCE's own solver (GPU suite): tests/golden/ref_hkd_trot.npz is the record of the reference's
MultiPhaseDDP<T>::solve / HKDProblem<T>::initialization / update, compiled unchanged from its sources (oracle/refbuild, tools/make_ref_golden.py;
see tests/test_cpu_reference_solver.py for what the fixture holds and why 1e-9 rather than the last bit). Decisions bit-exact, per-iteration
cost, gains and trajectories at 1e-9 relative (BASELINE.json north_star) - no oracle in between."""
import copy
import os

import numpy as np
import pytest

from test_cpu_reference_solver import (CSV, LONG_RTOL, REB_RTOL, RTOL, SS_RTOL, _Prefixed, barrel_problem, check_deck_layout, check_program, check_solve,  # noqa: F401
                                       mhpc_options, program_problem, ref, ref_barrel, ref_mhpc, ref_programs, reb_case, ref_reb, ref_ss, relerr, single_shooting_case)

pytestmark = pytest.mark.gpu


def test_gpu_reproduces_the_reference_solver_on_the_initial_solves(cm, hkd_options, ref):
    """BASELINE config 1 against the reference itself: the HKD trot solve of HKDMPCSolver::initialize (problem 0) and three perturbed starts,
    one batch; every decision of every DDP iteration (30 - 38 iterations, up to 73 line-search trials), the per-iteration record and the
    solution; the nominal problem's gains, Quu, Qux, G, Qu in full."""
    prob = cm.HKDProblem(CSV)
    B = len(ref["body"])
    x0 = np.stack([ref["p%d_s0_x0" % b] for b in range(B)])
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(x0)
    s.solve(hkd_options)
    info, trace, sol = s.get_solver_info(), s.get_trace(256), s.get_solution()
    for b in range(B):
        pre = "p%d_s0_" % b
        check_deck_layout(prob, ref, pre)
        check_solve(cm, prob, ref, pre, info[b], trace[b, :info[b]["iter"]], sol[b], full=(b == 0))


def test_gpu_update_deck_chain_reproduces_the_reference_update_chain(cm, hkd_options, ref):
    """SURVEY §8(f)1 against the reference itself: HKDProblem<T>::update x 6 + the re-solves of HKDMPCSolver<T>::update (caps 2 x 1) on ONE
    solver through cafe_gpu_update_deck - re-cut deck, on-device shift of the previous solution (Ubar[0] = 0), on-device carry-over of the
    touchdown constraints' sigma / lambda (the reference never resets them) - all four problems in one batch. x0 of every step = the plan's
    own prediction two knots ahead (cafe_gpu_get_planned_state) + the recorded nudge, which must equal the state the reference started from."""
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    B = len(ref["body"])
    n_upd = ref["nudge"].shape[1]
    prob = cm.HKDProblem(CSV)
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(np.stack([ref["p%d_s0_x0" % b] for b in range(B)]))
    s.solve(hkd_options)
    for step in range(1, n_upd + 1):
        p1 = cm.HKDProblem(CSV, k0=2 * step, mpc_update=True)
        x1 = s.planned_state(2)
        x1[:, 3:6] += ref["nudge"][:, step - 1]
        for b in range(B):
            np.testing.assert_allclose(x1[b], ref["p%d_s%d_x0" % (b, step)], rtol=RTOL, atol=1e-12)
        s.update_deck(p1, 2)
        s.set_initial_condition(x1)
        s.solve(ort)
        info, trace, sol = s.get_solver_info(), s.get_trace(64), s.get_solution()
        for b in range(B):
            pre = "p%d_s%d_" % (b, step)
            check_deck_layout(p1, ref, pre)
            check_solve(cm, p1, ref, pre, info[b], trace[b, :info[b]["iter"]], sol[b])


def test_gpu_fresh_solver_chain_with_carried_al_parameters(cm, hkd_options, ref):
    """The same chain through the other entries: a fresh solver per MPC step, host-side shift (cafe_mpc_b200/mpc.py) handed over with
    cafe_gpu_set_initial_guess, sigma / lambda read with cafe_gpu_get_al_params, shifted with mpc.shift_al and handed over with
    cafe_gpu_set_al_params - and what happens without the carry-over: the re-solves part ways with the reference's."""
    from cafe_mpc_b200 import mpc
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    B = 2
    prob, k0 = cm.HKDProblem(CSV), 0
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(np.stack([ref["p%d_s0_x0" % b] for b in range(B)]))
    s.solve(hkd_options)
    assert np.any(s.get_al_params() != mpc.initial_al(prob, B))       # the initial solve did move sigma / lambda
    differs = False
    for step in range(1, 4):
        k1 = k0 + 2
        p1 = cm.HKDProblem(CSV, k0=k1, mpc_update=True)
        sol = s.get_solution()
        x1 = np.stack([ref["p%d_s%d_x0" % (b, step)] for b in range(B)])
        s1 = cm.MultiPhaseDDP(p1, 0, B)
        s1.set_initial_condition(x1)
        s1.set_initial_guess(mpc.shifted_guess_batch(prob, k0, p1, k1, sol))
        s1.solve(ort)                                                  # sigma / lambda back at the deck's values
        reset_cost = [i["cost"] for i in s1.get_solver_info()]
        s1.set_al_params(mpc.shift_al(prob, k0, p1, k1, s.get_al_params()))
        s1.solve(ort)
        info, trace, sol1 = s1.get_solver_info(), s1.get_trace(64), s1.get_solution()
        for b in range(B):
            check_solve(cm, p1, ref, "p%d_s%d_" % (b, step), info[b], trace[b, :info[b]["iter"]], sol1[b])
            differs = differs or abs(reset_cost[b] - info[b]["cost"]) > 1e-6 * abs(info[b]["cost"])
        s.close()
        s, prob, k0 = s1, p1, k1
    assert differs


def test_gpu_reproduces_the_reference_mhpc_problem_on_the_initial_solves(cm, mhpc_options, ref_mhpc):
    """The headline workload (MHPC trot: whole-body 11 + 14 knots, SRB 10 knots) against the reference's own problem code and solver
    (tests/golden/ref_mhpc_trot.npz; what stands behind Pinocchio's names in that build: see the fixture in test_cpu_reference_solver.py)."""
    ref = ref_mhpc
    prob = cm.MHPCProblem(CSV)
    B = len(ref["x0"])
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(ref["x0"])
    s.solve(mhpc_options)
    info, trace, sol = s.get_solver_info(), s.get_trace(64), s.get_solution()
    for b in range(B):
        pre = "p%d_s0_" % b
        check_deck_layout(prob, ref, pre)
        check_solve(cm, prob, ref, pre, info[b], trace[b, :info[b]["iter"]], sol[b], full=(b == 0))


def test_gpu_update_deck_chain_reproduces_the_reference_mhpc_update_chain(cm, mhpc_options, ref_mhpc):
    """MHPCProblem<T>::update x 8 + the re-solves of MHPCLocomotion<T>::update (run-time caps) on ONE solver through cafe_gpu_update_deck: the
    front phase shrinks from 11 knots to 1 and disappears, a one-knot whole-body phase opens at the tail and grows, touchdown constraints
    appear and keep their sigma / lambda, the SRB plan keeps its data. Three problems in one batch, every step against the reference's record."""
    ref = ref_mhpc
    ort = copy.copy(mhpc_options)
    ort.max_AL_iter = mhpc_options.max_AL_iter_runtime; ort.max_DDP_iter = mhpc_options.max_DDP_iter_runtime
    B = len(ref["x0"])
    n_upd = ref["nudge"].shape[1]
    prob = cm.MHPCProblem(CSV)
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(ref["x0"])
    s.solve(mhpc_options)
    seen_td = False
    for step in range(1, n_upd + 1):
        p1 = cm.MHPCProblem(CSV, k0=2 * step, mpc_update_nsteps=2)
        seen_td = seen_td or any(p.n_td > 0 for p in p1.phases())
        x1 = s.planned_state(2) + ref["nudge"][:, step - 1]
        for b in range(B):
            np.testing.assert_allclose(x1[b], ref["p%d_s%d_x0" % (b, step)], rtol=RTOL, atol=1e-12)
        s.update_deck(p1, 2)
        s.set_initial_condition(x1)
        s.solve(ort)
        info, trace, sol = s.get_solver_info(), s.get_trace(64), s.get_solution()
        for b in range(B):
            pre = "p%d_s%d_" % (b, step)
            check_deck_layout(p1, ref, pre)
            check_solve(cm, p1, ref, pre, info[b], trace[b, :info[b]["iter"]], sol[b])
    assert seen_td


@pytest.mark.parametrize("k0", [0, 205])
def test_gpu_reproduces_the_reference_running_barrel_roll(cm, mhpc_options, ref_barrel, k0):
    """BASELINE config 4 against the reference's own code (see the CPU counterpart for what the two start offsets exercise): at offset 205 the
    landing impact, its partials, the reset map and four touchdown constraints; the perturbed problem's 200 iterations / 1 829 line-search
    trials are the reference's, decision for decision."""
    ref = _Prefixed(ref_barrel, "k%d_" % k0)
    prob = barrel_problem(cm, k0)
    s = cm.MultiPhaseDDP(prob, 0, 2)
    s.set_initial_condition(ref["x0"])
    s.solve(mhpc_options)
    info, trace, sol = s.get_solver_info(), s.get_trace(256), s.get_solution()
    for b in range(2):
        long_run = info[b]["iter"] >= 100
        check_solve(cm, prob, ref, "p%d_s0_" % b, info[b], trace[b, :info[b]["iter"]], sol[b], full=(b == 0 and k0 == 0), rtol=RTOL if not long_run else LONG_RTOL)


@pytest.mark.parametrize("name", ["loco", "barrel_to"])
def test_gpu_reproduces_the_reference_programs(cm, ref_programs, name):
    """SURVEY section 8(f)2 against the reference's stand-alone programs run unchanged (Loco_TO.cpp: 14 iterations; BarrelRollTO.cpp: 300
    iterations, 2 089 line-search trials): every decision of every iteration is the program's; tolerances as in check_program."""
    ref = _Prefixed(ref_programs, name + "_")
    prob, opt, x0, guess = program_problem(cm, name)
    s = cm.MultiPhaseDDP(prob, 0, 1)
    s.set_initial_condition(x0[None])
    if guess is not None:
        s.set_initial_guess(guess[None])
    s.solve(opt)
    info = s.get_solver_info()[0]
    check_program(cm, prob, ref, info, s.get_trace(320)[0, :info["iter"]], s.get_solution()[0], long_run=(name == "barrel_to"))


@pytest.mark.parametrize("kind", ["hkd", "mhpc"])
def test_gpu_reproduces_the_reference_in_whole_problem_single_shooting(cm, hkd_options, mhpc_options, ref_ss, kind):
    """HSDDP_OPTION::MS = false on the GPU against the reference's own run with that setting (HKD and MHPC trot, two problems each)."""
    ref, prob, opt, x0 = single_shooting_case(cm, ref_ss, kind, hkd_options, mhpc_options)
    s = cm.MultiPhaseDDP(prob, 0, 2)
    s.set_initial_condition(x0)
    s.solve(opt)
    info, trace, sol = s.get_solver_info(), s.get_trace(256), s.get_solution()
    for b in range(2):
        assert info[b]["feas"] == 0.0
        check_solve(cm, prob, ref, "p%d_s0_" % b, info[b], trace[b, :info[b]["iter"]], sol[b], rtol=SS_RTOL[kind])


def test_gpu_reproduces_the_reference_relaxed_barrier_updates(cm, mhpc_options, ref_reb):
    """PathConstraintBase::update_params with update_relax = 0.5, update_ReB = 2 (k_reb_update, RebCtx) against the reference's own run: the
    landing problem of the running barrel roll, 104 iterations / 301 trials and 200 / 1 864."""
    prob, opt = reb_case(cm, ref_reb, mhpc_options)
    s = cm.MultiPhaseDDP(prob, 0, 2)
    s.set_initial_condition(ref_reb["x0"])
    s.solve(opt)
    info, trace, sol = s.get_solver_info(), s.get_trace(256), s.get_solution()
    for b in range(2):
        check_solve(cm, prob, ref_reb, "p%d_s0_" % b, info[b], trace[b, :info[b]["iter"]], sol[b], rtol=REB_RTOL)


def test_gpu_update_deck_carries_changing_barrier_parameters_like_the_reference(cm, mhpc_options, ref_reb):
    """cafe_gpu_update_deck with update_relax = 0.5, update_ReB = 2: the per-(knot, element) update counts travel with the knots (an appended
    knot copies the last knot's) - four MPC updates of the barrel roll's landing problem on one solver against the reference's own chain."""
    from cafe_mpc_b200 import workload
    ref = _Prefixed(ref_reb, "chain_")
    prob, opt = reb_case(cm, ref_reb, mhpc_options)
    ort = copy.copy(opt)
    ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    s = cm.MultiPhaseDDP(prob, 0, 2)
    s.set_initial_condition(np.stack([ref_reb["x0"][0], ref_reb["x0"][0]]))      # two copies: the carry-over is per problem
    s.solve(opt)
    for step in range(1, 5):
        p1 = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=205 + 2 * step, mpc_update_nsteps=2)
        x1 = s.planned_state(2)
        s.update_deck(p1, 2)
        s.set_initial_condition(x1)
        s.solve(ort)
        info, trace, sol = s.get_solver_info(), s.get_trace(64), s.get_solution()
        for b in range(2):
            check_solve(cm, p1, ref, "p0_s%d_" % step, info[b], trace[b, :info[b]["iter"]], sol[b], rtol=REB_RTOL)
