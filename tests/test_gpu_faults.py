"""Failure branches of the solve ("these branches are part of iteration counts bit-exact"), GPU against the oracle:
  * a non-positive-definite Quu climbs the regularisation ladder (MultiPhaseDDP::backward_sweep_regularized, MultiPhaseDDP.cpp:136-165):
    recovered solves and the give-up beyond 1e2 (status CAFE_STATUS_REG_FAIL, "bad_solve" :317-320, :436-442);
  * line-search trials whose rollout leaves the 1e6 ball are rejected whatever their merit (SinglePhase.cpp:205-208,
    MultiPhaseDDP.cpp:80-84, :123); a later, smaller step is accepted and the solve carries on.
Decks are the shipped MHPC trot deck with the control weights / the initial state pushed out of the region the problem was tuned for."""
import copy
import ctypes as C
import os

import numpy as np
import pytest

from oracle_bindings import oracle_solve

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
COUNTS = ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")


class _Prob:   # what MultiPhaseDDP needs from a problem object: a deck pointer
    def __init__(self, base, deck_ptr, keep):
        self.deck = deck_ptr; self._keep = keep; self._base = base

    def phases(self):
        return self._base.phases()


def _with_control_weight(cm, prob, r):
    from cafe_mpc_b200._ctypes_defs import Deck
    d2 = Deck.from_buffer_copy(prob.deck.contents)
    for i in range(d2.n_phases):
        for j in range(24):
            d2.phase[i].r[j] = r
    return _Prob(prob, C.pointer(d2), d2)


@pytest.fixture(scope="module")
def setup(cm):
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    opt = copy.copy(opt); opt.max_AL_iter = 2; opt.max_DDP_iter = 3
    return prob, opt


def _solve(cm, prob, opt, x0):
    s = cm.MultiPhaseDDP(prob, 0, len(x0))
    s.set_initial_condition(x0)
    s.solve(opt)
    return s


@pytest.mark.parametrize("r", [-0.5, -20.0, -1000.0])
def test_regularisation_ladder_recovers_like_the_oracle(cm, setup, r):
    """negative control weights make Quu indefinite: every sweep is repeated with reg = max(2 reg, 1e-3) until it factorises"""
    from cafe_mpc_b200 import workload
    prob, opt = setup
    p2 = _with_control_weight(cm, prob, r)
    x0 = workload.mhpc_batch(3)
    s = _solve(cm, p2, opt, x0)
    info = s.get_solver_info(); hist = s.get_history(64); trace = s.get_trace(64)
    for b in range(3):
        oi, oh, ot, _ = oracle_solve(p2.deck, opt, x0[b])
        assert oi["reg_iter_total"] > oi["iter"]                      # the ladder was climbed
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (r, b)
        assert np.array_equal(trace[b, :oi["iter"], 6:10], ot[:, 6:10])   # sweeps, trials, success, accepted step per iteration
        np.testing.assert_allclose(trace[b, :oi["iter"], 5], ot[:, 5], rtol=0, atol=0)   # regularisation left behind, exact
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=1e-9)


def test_regularisation_beyond_1e2_gives_up_like_the_oracle(cm, setup):
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200._ctypes_defs import CAFE_STATUS_REG_FAIL
    prob, opt = setup
    p2 = _with_control_weight(cm, prob, -1e5)
    x0 = workload.mhpc_batch(3)
    s = _solve(cm, p2, opt, x0)
    info = s.get_solver_info()
    for b in range(3):
        oi, oh, ot, _ = oracle_solve(p2.deck, opt, x0[b])
        assert oi["status"] == CAFE_STATUS_REG_FAIL and oi["reg_iter_total"] == 18   # 1e-3 * 2^17 > 1e2
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], b
        assert abs(info[b]["cost"] - oi["cost"]) <= 1e-9 * abs(oi["cost"])


def test_far_initial_state_matches_the_oracle(cm, setup):
    """an initial state 1 km from the plan: large search directions, full steps are rejected by the merit test, small ones accepted
    (6, 5 and 4 trials in the first three iterations). Absolute positions of 1e3 leave 1e-13 for the leg kinematics and differences
    of that size grow by an order of magnitude per iteration (1e-9 in the third, 1e-4 in the sixth, where they start to flip
    regularisation decisions - further out, at 1e4, within three iterations): the comparison stops after three iterations."""
    from cafe_mpc_b200 import workload
    prob, opt = setup
    opt = copy.copy(opt); opt.max_AL_iter = 1; opt.max_DDP_iter = 3
    x0 = workload.mhpc_batch(3)
    x0[:, 0] += 1e3
    s = _solve(cm, prob, opt, x0)
    info = s.get_solver_info(); hist = s.get_history(64); trace = s.get_trace(64)
    for b in range(3):
        oi, oh, ot, _ = oracle_solve(prob.deck, opt, x0[b])
        assert oi["ls_iter_total"] >= 12                                        # full steps were rejected
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], b
        n = oi["iter"]
        assert np.array_equal(trace[b, :n, 6:10], ot[:n, 6:10])
        np.testing.assert_allclose(hist[b, :oi["n_hist"], 0], oh[:, 0], rtol=1e-7)


@pytest.mark.parametrize("off,n_ls", [(4e7, 7), (1e8, 8)])
def test_trials_leaving_the_1e6_ball_are_rejected_like_the_oracle(cm, setup, off, n_ls):
    """SinglePhase.cpp:205-208 / MultiPhaseDDP.cpp:80-84, :123: a trial whose rollout leaves |Xsim| <= 1e6 is rejected whatever its merit.
    The first trial state is Xbar[0] + eps * (x0 - Xbar[0]): with x0 4e7 (1e8) away, the steps down to 1/32 (1/64) start outside the ball.
    The merit test alone accepts 1/32 (6 trials, as for every smaller offset: the problem is translation invariant); the ball makes it
    7 (8) trials. One DDP iteration, so that only these decisions are compared."""
    from cafe_mpc_b200 import workload
    prob, opt = setup
    opt = copy.copy(opt); opt.max_AL_iter = 1; opt.max_DDP_iter = 1
    x0 = workload.mhpc_batch(3)
    x0[:, 0] += off
    s = _solve(cm, prob, opt, x0)
    info = s.get_solver_info(); trace = s.get_trace(64)
    for b in range(3):
        oi, oh, ot, _ = oracle_solve(prob.deck, opt, x0[b])
        assert oi["ls_iter_total"] == n_ls and oi["status"] == 0
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (off, b)
        assert np.array_equal(trace[b, :1, 6:10], ot[:1, 6:10])                      # sweeps, trials, success, accepted step
        assert abs(info[b]["cost"] - oi["cost"]) <= 1e-6 * abs(oi["cost"])         # positions of 1e6: 1e-10 absolute
