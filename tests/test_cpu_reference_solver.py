"""The oracle against the REFERENCE's own solver (CPU suite).

tests/golden/ref_hkd_trot.npz holds what the reference's MultiPhaseDDP<T>::solve + HKDProblem<T>::initialization / update decided and
produced on four HKD trot problems through the initial solve and six consecutive MPC updates: it was written by
tools/make_ref_golden.py from oracle/_ref/ref_hkd, i.e. from the reference's HSDDPSolver / HKD-TrajOpt / QuadReference sources compiled
unchanged (oracle/refbuild). These tests pin the oracle's restatement of the solver layer, the repo's deck builders, compute_hkd_state and
the MPC shift (cafe_mpc_b200/mpc.py) to it: decisions (iteration, line-search and regularisation counts per DDP iteration, accepted step
sizes, AL updates, phase layouts) bit-exact, everything else at 1e-9 relative (BASELINE.json north_star). The reference binary was built
against a plain-loop stand-in for Eigen (not in this image), so last-bit agreement is not expected - 1e-9 is met with orders to spare.
The GPU counterpart is tests/test_gpu_reference_solver.py."""
import copy
import os

import numpy as np
import pytest

from oracle_bindings import oracle_solve

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
RTOL = 1e-9
LONG_RTOL = 1e-9     # runs of 100+ DDP iterations: no allowance needed so far


def relerr(g, o):
    o = np.asarray(o); g = np.asarray(g)
    if o.size == 0:
        return 0.0
    return float(np.max(np.abs(g - o)) / max(np.max(np.abs(o)), 1e-300))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(REPO, "tests/golden/ref_hkd_trot.npz"))


@pytest.fixture(scope="module")
def ref_mhpc():
    """The same record for the headline workload: the reference's MHPCProblem / WBM / MHPCCost / MHPCConstraint / MHPCReset / MHPCReference /
    SRBM code and solver on the MHPC trot (oracle/_ref/ref_mhpc). Pinocchio is not in this image: behind its names stand the oracle's own
    rigid-body recursions (oracle/refbuild/shim/pinocchio), so this file pins every layer of the whole-body problem EXCEPT the rigid-body
    algorithms themselves (mass matrix, bias forces, RNEA derivatives, foot kinematics), which are pinned by the reference's known answers
    (tests/test_cpu_mhpc.py::test_wb_contact_dynamics_known_answers) and its CasADi kinematic partials."""
    return np.load(os.path.join(REPO, "tests/golden/ref_mhpc_trot.npz"))


class _Prefixed:
    """view of an npz with a key prefix (the barrel-roll file holds two start offsets)"""
    def __init__(self, npz, prefix):
        self.npz, self.prefix = npz, prefix

    def __getitem__(self, k):
        return self.npz[k] if k == "kv" else self.npz[self.prefix + k]


@pytest.fixture(scope="module")
def ref_barrel():
    """BASELINE config 4 (running barrel roll) at the start offsets 0 and 205, oracle/_ref/ref_mhpc with mhpc_config_barrel.info."""
    return np.load(os.path.join(REPO, "tests/golden/ref_mhpc_barrel.npz"))


@pytest.fixture(scope="module")
def ref_programs():
    """The reference's stand-alone programs Loco_TO.cpp and BarrelRollTO.cpp run unchanged, main() included (oracle/_ref/ref_loco,
    ref_barrel_to; tools/make_ref_golden.py::main_programs)."""
    return np.load(os.path.join(REPO, "tests/golden/ref_programs.npz"))


def program_problem(cm, name):
    """(problem, options, x0 [36], packed initial guess or None) of the reference program `name`, as this repo builds them."""
    from cafe_mpc_b200 import workload
    x0 = workload.mhpc_batch(1)[0]          # both mains start from (0, 0, 0.2183), qJ = (0, -1, 2) x 4 (Loco_TO.cpp:49-55, BarrelRollTO.cpp:97-110)
    if name == "loco":
        return cm.LocoProblem(), cm.load_hsddp_setting(workload.LOCO_DDP_SETTING), x0, None
    prob = cm.BarrelRollProblem()
    return prob, cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING), x0, prob.initial_guess(x0[None])[0]


def check_program(cm, prob, ref, info, trace, sol, long_run):
    """Against a reference program's record. Decisions of every DDP iteration bit-exact in both programs. Loco_TO (14 iterations): everything
    else at 1e-9. BarrelRollTO (300 iterations, 2 089 line-search trials from a start with defect norm 58): rounding differences of 1e-13 in the
    first sweep grow to 1e-6 in the per-iteration cost around iteration 30 and shrink again - measured between four roundings of the oracle
    itself - so its history is held at 1e-5, its final cost at 1e-8 and its trajectories at 1e-5."""
    pre = "p0_s0_"
    assert [info["iter"], info["ls_iter_total"], info["reg_iter_total"]] == list(ref[pre + "counters"])
    rt = ref[pre + "trace"]
    assert np.array_equal(trace[:, 6:10], rt[:, 6:10])
    if not long_run:
        check_solve(cm, prob, ref, pre, info, trace, sol)
        return
    np.testing.assert_allclose(trace[:, [0, 10]], rt[:, [0, 10]], rtol=1e-5)
    np.testing.assert_allclose(trace[:, [1, 11]], rt[:, [1, 11]], rtol=2e-3, atol=1e-8)       # infeasibility: down to 1e-3 at the end, where it is the more sensitive figure
    assert abs(info["cost"] - ref[pre + "final"][0]) <= 1e-8 * abs(ref[pre + "final"][0])
    for i, p in enumerate(cm.unpack_solution(prob.deck, sol)):
        for name in ("Xbar", "Ubar"):
            assert relerr(p[name], ref[pre + "ph%d_" % i + name]) < 1e-5, (i, name)


@pytest.mark.parametrize("name", ["loco", "barrel_to"])
def test_oracle_reproduces_the_reference_programs(cm, ref_programs, name):
    """LocoProblem (Loco_TO.cpp) and the in-place barrel roll (BarrelRollTO.cpp, whose problem - switching times, six
    desired states, per-phase weights, BarrelRoll:: barriers incl. the joint-speed limit, the interpolated initial trajectory - is built inside
    its main()) against the programs themselves."""
    ref = _Prefixed(ref_programs, name + "_")
    prob, opt, x0, guess = program_problem(cm, name)
    np.testing.assert_array_equal(x0, ref["p0_s0_x0"])
    assert [p.horizon for p in prob.phases()] == list(ref["p0_s0_horizons"])
    info, hist, trace, sol = oracle_solve(prob.deck, opt, x0, cap=320, guess=guess)
    check_program(cm, prob, ref, info, trace, sol, long_run=(name == "barrel_to"))


@pytest.fixture(scope="module")
def mhpc_options(cm):
    return cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))


def check_deck_layout(prob, ref, pre):
    ph = prob.phases()
    assert [p.horizon for p in ph] == list(ref[pre + "horizons"])
    assert [list(p.contact) for p in ph] == [list(c) for c in ref[pre + "contacts"]]


def check_solve(cm, prob, ref, pre, info, trace, sol, full=False, rtol=RTOL):
    """One solve (oracle or GPU: info dict, trace [iter,12], packed solution) against the reference's record `pre`."""
    assert [info["iter"], info["ls_iter_total"], info["reg_iter_total"]] == list(ref[pre + "counters"]), pre
    rt = ref[pre + "trace"]
    assert trace.shape == rt.shape
    assert np.array_equal(trace[:, 6:10], rt[:, 6:10]), pre            # sweeps, line-search trials, success, accepted step size: exact
    assert np.array_equal(trace[:, 5] == 0, rt[:, 5] == 0)
    np.testing.assert_allclose(trace[:, 0], rt[:, 0], rtol=rtol, err_msg=pre)              # cost at the start of every DDP iteration
    np.testing.assert_allclose(trace[:, 10], rt[:, 10], rtol=rtol, err_msg=pre)            # ... and after its line search
    np.testing.assert_allclose(trace[:, [1, 11]], rt[:, [1, 11]], rtol=rtol, atol=1e-13, err_msg=pre)    # dynamics infeasibility (down to 1e-5)
    np.testing.assert_allclose(trace[:, 2:4], rt[:, 2:4], rtol=rtol, atol=1e-12, err_msg=pre)            # expected cost change
    np.testing.assert_allclose(trace[:, 4:6], rt[:, 4:6], rtol=1e-8, atol=1e-12, err_msg=pre)            # merit parameter (a ratio of the above), regularisation
    assert abs(info["cost"] - ref[pre + "final"][0]) <= rtol * abs(ref[pre + "final"][0])
    assert abs(info["feas"] - ref[pre + "final"][1]) <= rtol * abs(ref[pre + "final"][1]) + 1e-13
    for i, p in enumerate(cm.unpack_solution(prob.deck, sol)):
        q = pre + "ph%d_" % i
        kv = ref["kv"][:p["Xbar"].shape[1]]
        for name in ("Xbar", "Ubar", "dU"):
            assert relerr(p[name], ref[q + name]) < rtol, (pre, i, name)
        assert relerr(p["K"] @ kv, ref[q + "Kv"]) < rtol, (pre, i, "K v")
        if i == 0:
            assert relerr(p["K"][:4], ref[q + "K4"]) < rtol, (pre, "K4")
        if full:
            for name in ("K", "Quu", "Qux", "G", "Qu"):
                assert relerr(p[name], ref[q + name]) < rtol, (pre, i, name)


def test_initial_state_is_the_references(cm, ref):
    """compute_hkd_state (HKDModel.h:66-96) as called by HKDMPCSolver::initialize, run by the reference itself."""
    prob = cm.HKDProblem(CSV)
    for b in range(len(ref["body"])):
        x0 = prob.initial_state(ref["body"][b], ref["qJ"][b])
        np.testing.assert_allclose(x0, ref["p%d_s0_x0" % b], rtol=0, atol=1e-15)


def test_oracle_reproduces_the_reference_solver_on_the_initial_solves(cm, hkd_options, ref):
    """MultiPhaseDDP<T>::solve on the problem HKDProblem<T>::initialization builds, caps 10 x 5 (30 - 38 DDP iterations, up to 73
    line-search trials, four AL updates): every decision of every iteration and the whole solution."""
    prob = cm.HKDProblem(CSV)
    for b in range(len(ref["body"])):
        pre = "p%d_s0_" % b
        check_deck_layout(prob, ref, pre)
        info, hist, trace, sol = oracle_solve(prob.deck, hkd_options, ref[pre + "x0"])
        assert info["outer_iter"] - 1 == int(ref[pre + "n_al"]) or info["outer_iter"] == int(ref[pre + "n_al"])
        check_solve(cm, prob, ref, pre, info, trace, sol, full=(b == 0))


def test_oracle_and_mpc_shift_reproduce_the_reference_update_chain(cm, hkd_options, ref):
    """HKDProblem<T>::update x 6 (HKDProblem.cpp:117-222: tail phase opened at offset 2, front phase removed at 12, Ubar[0] = 0) with the
    re-solves of HKDMPCSolver<T>::update (caps 2 x 1): the re-cut decks, the warm start mpc.py shifts out of OUR previous solution and the
    re-solve all equal the reference's at every step. x0 of step s = the plan's own prediction + the recorded nudge."""
    from cafe_mpc_b200 import mpc
    ort = copy.copy(hkd_options)
    ort.max_AL_iter = 2; ort.max_DDP_iter = 1
    kv = ref["kv"]
    n_upd = ref["nudge"].shape[1]
    for b in range(2):
        prob, k0 = cm.HKDProblem(CSV), 0
        info, hist, trace, sol, al = oracle_solve(prob.deck, hkd_options, ref["p%d_s0_x0" % b], al=mpc.initial_al(prob))
        layouts = set()
        for s in range(1, n_upd + 1):
            pre = "p%d_s%d_" % (b, s)
            k1 = k0 + 2
            p1 = cm.HKDProblem(CSV, k0=k1, mpc_update=True)
            check_deck_layout(p1, ref, pre)
            layouts.add(tuple(ref[pre + "horizons"]))
            guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol[None, :])[0]
            for i, g in enumerate(cm.unpack_solution(p1.deck, guess)):
                q = pre + "ph%d_" % i
                if p1.phases()[i].single_shooting:
                    # a tail phase this update opened: the reference leaves its fresh Trajectory zeroed (HKDProblem.cpp:166-167), mpc.py
                    # writes reference states; neither is ever read (no shooting states, K = 0, the first rollout overwrites Xbar)
                    assert not ref[q + "gXbar"].any() and not ref[q + "gUbar"].any() and not g["Ubar"].any() and not g["K"].any()
                    continue
                assert relerr(g["Xbar"], ref[q + "gXbar"]) < RTOL and relerr(g["Ubar"], ref[q + "gUbar"]) < RTOL, (pre, i)
                assert relerr(g["K"] @ kv, ref[q + "gKv"]) < RTOL, (pre, i)
            x1 = mpc.state_at(prob, cm.unpack_solution(prob.deck, sol), 2)
            x1[3:6] += ref["nudge"][b, s - 1]
            np.testing.assert_allclose(x1, ref[pre + "x0"], rtol=RTOL, atol=1e-12)
            al = mpc.shift_al(prob, k0, p1, k1, al)
            info, hist, trace, sol, al = oracle_solve(p1.deck, ort, x1, guess=guess, al=al)
            check_solve(cm, p1, ref, pre, info, trace, sol)
            prob, k0 = p1, k1
        assert len(layouts) >= 3      # the chain went through a tail-phase opening and a front-phase removal


def test_oracle_reproduces_the_reference_mhpc_problem_on_the_initial_solves(cm, mhpc_options, ref_mhpc):
    """BASELINE config 2 / the headline workload against the reference's own whole-body + SRB problem code and solver: MHPCProblem<T>::
    initialization + MultiPhaseDDP<T>::solve (MHPCLocomotion.cpp:20-66) on the nominal and two perturbed trot starts."""
    ref = ref_mhpc
    prob = cm.MHPCProblem(CSV)
    for b in range(len(ref["x0"])):
        pre = "p%d_s0_" % b
        check_deck_layout(prob, ref, pre)
        np.testing.assert_array_equal(ref["x0"][b], ref[pre + "x0"])
        info, hist, trace, sol = oracle_solve(prob.deck, mhpc_options, ref["x0"][b])
        check_solve(cm, prob, ref, pre, info, trace, sol, full=(b == 0))


def test_oracle_and_mpc_shift_reproduce_the_reference_mhpc_update_chain(cm, mhpc_options, ref_mhpc):
    """MHPCProblem<T>::update x 8 (MHPCProblem.cpp:252-397: the front phase shrinks 11 -> 1 and disappears, a one-knot tail phase opens and
    grows, the SRB plan keeps its data) with the re-solves of MHPCLocomotion<T>::update under the run-time caps: re-cut decks, shifted warm
    start, carried sigma / lambda and the re-solves equal the reference's at every step."""
    from cafe_mpc_b200 import mpc
    ref = ref_mhpc
    ort = copy.copy(mhpc_options)
    ort.max_AL_iter = mhpc_options.max_AL_iter_runtime; ort.max_DDP_iter = mhpc_options.max_DDP_iter_runtime
    n_upd = ref["nudge"].shape[1]
    for b in (0, 2):
        prob, k0 = cm.MHPCProblem(CSV), 0
        info, hist, trace, sol, al = oracle_solve(prob.deck, mhpc_options, ref["x0"][b], al=mpc.initial_al(prob))
        layouts = set()
        for s in range(1, n_upd + 1):
            pre = "p%d_s%d_" % (b, s)
            k1 = k0 + 2
            p1 = cm.MHPCProblem(CSV, k0=k1, mpc_update_nsteps=2)
            check_deck_layout(p1, ref, pre)
            layouts.add(tuple(ref[pre + "horizons"]))
            guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol[None, :])[0]
            for i, g in enumerate(cm.unpack_solution(p1.deck, guess)):
                q = pre + "ph%d_" % i
                if p1.phases()[i].single_shooting:
                    assert not ref[q + "gUbar"].any() and not g["Ubar"].any() and not g["K"].any()     # a fresh phase: nothing of it is read
                    continue
                kv = ref["kv"][:g["Xbar"].shape[1]]
                assert relerr(g["Xbar"], ref[q + "gXbar"]) < RTOL and relerr(g["Ubar"], ref[q + "gUbar"]) < RTOL, (pre, i)
                assert relerr(g["K"] @ kv, ref[q + "gKv"]) < RTOL, (pre, i)
            x1 = mpc.state_at(prob, cm.unpack_solution(prob.deck, sol), 2) + ref["nudge"][b, s - 1]
            np.testing.assert_allclose(x1, ref[pre + "x0"], rtol=RTOL, atol=1e-12)
            al = mpc.shift_al(prob, k0, p1, k1, al)
            info, hist, trace, sol, al = oracle_solve(p1.deck, ort, x1, guess=guess, al=al)
            check_solve(cm, p1, ref, pre, info, trace, sol)
            prob, k0 = p1, k1
        assert (1, 24, 10) in layouts and (24, 1, 10) in layouts


def barrel_problem(cm, k0):
    from cafe_mpc_b200 import workload
    return cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=k0)


@pytest.mark.parametrize("k0", [0, 205])
def test_oracle_reproduces_the_reference_running_barrel_roll(cm, mhpc_options, ref_barrel, k0):
    """BASELINE config 4 against the reference's own code. Offset 0: stance -> diagonal pair -> flight. Offset 205: 22 flight knots, then the
    four-foot landing - WBM::impact / impact_partial (incl. the impulse indexing of WBM.cpp:451-454), MHPCReset, four touchdown constraints
    under the augmented Lagrangian, seven AL updates; the perturbed problem runs into the 10 x 20 iteration cap with 1 829 line-search trials
    and still takes every decision like the reference; its 200 per-iteration records hold at 1e-9 like everything else."""
    ref = _Prefixed(ref_barrel, "k%d_" % k0)
    prob = barrel_problem(cm, k0)
    for b in range(2):
        pre = "p%d_s0_" % b
        check_deck_layout(prob, ref, pre)
        info, hist, trace, sol = oracle_solve(prob.deck, mhpc_options, ref["x0"][b])
        long_run = info["iter"] >= 100
        check_solve(cm, prob, ref, pre, info, trace, sol, full=(b == 0 and k0 == 0), rtol=RTOL if not long_run else LONG_RTOL)


@pytest.fixture(scope="module")
def ref_ss():
    """HSDDP_OPTION::MS = false run by the reference itself (ref_hkd / ref_mhpc with `MS false` in ddp_setting.info; tools/make_ref_golden.py)."""
    return np.load(os.path.join(REPO, "tests/golden/ref_single_shooting.npz"))


# whole-problem single shooting integrates the entire horizon open loop around every trial: on the whole-body problem (35 knots from a cost of 978)
# rounding differences between two correct implementations reach 1e-8 in the early iterations' cost (decisions stay equal); HKD holds 1e-9
SS_RTOL = {"hkd": RTOL, "mhpc": 1e-6}


def single_shooting_case(cm, ref_ss, kind, hkd_options, mhpc_options):
    ref = _Prefixed(ref_ss, kind + "_")
    prob = cm.HKDProblem(CSV) if kind == "hkd" else cm.MHPCProblem(CSV)
    opt = copy.copy(hkd_options if kind == "hkd" else mhpc_options)
    opt.MS = 0
    x0 = np.stack([ref["p%d_s0_x0" % b] for b in range(2)])
    return ref, prob, opt, x0


@pytest.mark.parametrize("kind", ["hkd", "mhpc"])
def test_oracle_reproduces_the_reference_in_whole_problem_single_shooting(cm, hkd_options, mhpc_options, ref_ss, kind):
    """MS = false (MultiPhaseDDP.cpp:65-68: every phase's shooting set cleared; :330-333: no linear rollout; SinglePhase.cpp:383-387: the
    expected cost change comes from the sweep) against the reference's own run with that setting, HKD and MHPC trot."""
    ref, prob, opt, x0 = single_shooting_case(cm, ref_ss, kind, hkd_options, mhpc_options)
    for b in range(2):
        info, hist, trace, sol = oracle_solve(prob.deck, opt, x0[b])
        assert info["feas"] == 0.0
        check_solve(cm, prob, ref, "p%d_s0_" % b, info, trace, sol, rtol=SS_RTOL[kind])


@pytest.fixture(scope="module")
def ref_reb():
    """PathConstraintBase::update_params with update_relax = 0.5, update_ReB = 2 run by the reference itself (ref_mhpc on the running barrel roll at
    start offset 205, ddp_setting.info with those two factors; tools/make_ref_golden.py::main_reb)."""
    return np.load(os.path.join(REPO, "tests/golden/ref_reb_update.npz"))


REB_RTOL = 1e-8    # 104 / 200 iterations with barriers that tighten on the way: the feed-forward term dU (a residual that vanishes at convergence) differs by 1.3e-9


def reb_case(cm, ref_reb, mhpc_options):
    opt = copy.copy(mhpc_options)
    opt.update_relax = float(ref_reb["update_relax"]); opt.update_ReB = float(ref_reb["update_ReB"])
    return barrel_problem(cm, 205), opt


def test_oracle_reproduces_the_reference_relaxed_barrier_updates(cm, mhpc_options, ref_reb):
    """SURVEY section 8 row a10 against the reference's own code: relaxed-barrier parameters that change during the solve (delta halved down to
    delta_min, eps doubled, per violated (knot, element), 11 and 20 update rounds) on the landing problem of the running barrel roll - 104
    iterations / 301 line-search trials and 200 / 1 864, every decision the reference's."""
    prob, opt = reb_case(cm, ref_reb, mhpc_options)
    for b in range(2):
        info, hist, trace, sol = oracle_solve(prob.deck, opt, ref_reb["x0"][b])
        check_solve(cm, prob, ref_reb, "p%d_s0_" % b, info, trace, sol, rtol=REB_RTOL)


def test_oracle_and_shift_reb_reproduce_the_reference_chain_with_changing_barriers(cm, mhpc_options, ref_reb):
    """The relaxed-barrier parameters travel with the knots through MHPCProblem::update (PathConstraintBase::pop_front / push_back,
    ConstraintsBase.h:296-306: an appended knot copies the LAST knot's delta / eps; reset_params empty): four MPC updates of the barrel roll's
    landing problem with update_relax = 0.5, update_ReB = 2 against the reference's own chain. Without the carry-over the second re-solve is
    off by 50 % in cost; with it (oracle entry cafe_oracle_solve_carry + mpc.shift_reb) every step equals the reference's."""
    from cafe_mpc_b200 import mpc, workload
    from oracle_bindings import oracle_reb_init
    ref = _Prefixed(ref_reb, "chain_")
    prob, opt = reb_case(cm, ref_reb, mhpc_options)
    ort = copy.copy(opt)
    ort.max_AL_iter = opt.max_AL_iter_runtime; ort.max_DDP_iter = opt.max_DDP_iter_runtime
    k0 = 205
    info, hist, trace, sol, al, reb = oracle_solve(prob.deck, opt, ref_reb["x0"][0], al=mpc.initial_al(prob), reb=oracle_reb_init(prob.deck))
    check_solve(cm, prob, ref, "p0_s0_", info, trace, sol, rtol=REB_RTOL)
    assert any(np.any(r != r0) for r, r0 in zip(reb, oracle_reb_init(prob.deck)))
    for s in range(1, 5):
        k1 = k0 + 2
        p1 = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=k1, mpc_update_nsteps=2)
        check_deck_layout(p1, ref, "p0_s%d_" % s)
        guess = mpc.shifted_guess_batch(prob, k0, p1, k1, sol[None, :])[0]
        x1 = mpc.state_at(prob, cm.unpack_solution(prob.deck, sol), 2)
        np.testing.assert_allclose(x1, ref["p0_s%d_x0" % s], rtol=REB_RTOL, atol=1e-10)
        al = mpc.shift_al(prob, k0, p1, k1, al)
        reb = mpc.shift_reb(prob, k0, p1, k1, reb, oracle_reb_init(p1.deck))
        info, hist, trace, sol, al, reb = oracle_solve(p1.deck, ort, x1, guess=guess, al=al, reb=reb)
        check_solve(cm, p1, ref, "p0_s%d_" % s, info, trace, sol, rtol=REB_RTOL)
        prob, k0 = p1, k1


def test_reference_record_of_the_headline_batch_matches_the_oracle_on_a_spread_sample(cm):
    """tests/golden/ref_mhpc_headline.npz = the reference's own solver build on all 4096 problems of the headline batch (tools/ref_headline.py; that tool
    compares the oracle with every one of them). Here: the record belongs to the table `workload.mhpc_batch` produces today, and 32 problems spread over
    the batch through the oracle take every decision like the reference (final cost 1e-9)."""
    import hashlib
    from cafe_mpc_b200 import workload
    from oracle_bindings import oracle_solve
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = np.load(os.path.join(repo, "tests/golden/ref_mhpc_headline.npz"))
    x0 = workload.mhpc_batch(4096)
    assert hashlib.sha1(np.ascontiguousarray(x0).tobytes()).hexdigest() == str(ref["x0_sha1"])
    assert ref["counters"].shape == (4096, 3) and ref["counters"][:, 0].min() >= 1
    prob = cm.MHPCProblem(os.path.join(repo, "data/Reference/Data/trot/heuristic/quad_reference.csv"))
    opt = cm.load_hsddp_setting(os.path.join(repo, "data/MHPC/settings/ddp_setting.info"))
    for b in range(5, 4096, 128):
        info, _, _, _ = oracle_solve(prob.deck, opt, x0[b], cap=320)
        assert [info["iter"], info["ls_iter_total"], info["reg_iter_total"]] == ref["counters"][b].tolist(), b
        assert abs(info["cost"] - ref["final_cost"][b]) <= 1e-9 * abs(ref["final_cost"][b]), b
