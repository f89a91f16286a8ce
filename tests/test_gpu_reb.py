"""PathConstraintBase::update_params for update_relax / update_ReB != 1 (HSDDPSolver/header/ConstraintsBase.h:79-85, :194-209, called from
MultiPhaseDDP::solve at MultiPhaseDDP.cpp:417-420): at the end of every outer iteration each path-constraint element that is violated
(g <= -pconstr_thresh on the last rollout) gets eps *= update_ReB and delta = max(delta * update_relax, delta_min) - per knot and per
element, so the barrier parameters of a batch diverge problem by problem. GPU against the oracle on decks whose limits are tightened until
constraints are violated at the end of outer iterations (shipped settings never get there: their factors are 1)."""
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


class _Prob:
    def __init__(self, base, deck_ptr, keep):
        self.deck = deck_ptr; self._keep = keep; self._base = base

    def phases(self):
        return self._base.phases()


def _tightened(prob, mu=None, torque=None, delta_scale=None):
    from cafe_mpc_b200._ctypes_defs import Deck
    d2 = Deck.from_buffer_copy(prob.deck.contents)
    for i in range(d2.n_phases):
        p = d2.phase[i]
        if mu is not None: p.mu = mu
        if torque is not None: p.torque_limit = torque
        if delta_scale is not None:
            p.reb_grf.delta = p.reb_grf.delta_min * delta_scale      # below its floor: the first update lifts it to delta_min
    return _Prob(prob, C.pointer(d2), d2)


def _compare(cm, prob, opt, x0, hist_rtol=1e-9):
    s = cm.MultiPhaseDDP(prob, 0, len(x0))
    s.set_initial_condition(x0)
    s.solve(opt)
    info = s.get_solver_info(); hist = s.get_history(320); sol = s.get_solution()
    out = []
    for b in range(len(x0)):
        oi, oh, ot, osol = oracle_solve(prob.deck, opt, x0[b])
        assert [info[b][k] for k in COUNTS] == [oi[k] for k in COUNTS], (b, info[b], oi)
        np.testing.assert_allclose(hist[b, :oi["n_hist"]], oh, rtol=hist_rtol, atol=1e-12)
        scale = np.abs(osol).max()
        assert np.abs(sol[b] - osol).max() <= 1e-9 * scale, b
        out.append((oi, oh))
    return out


@pytest.mark.parametrize("relax,weight", [(0.5, 0.2), (0.1, 5.0), (1.0, 0.5)])
def test_mhpc_reb_update_matches_oracle(cm, relax, weight):
    from cafe_mpc_b200 import workload
    base = cm.MHPCProblem(CSV, k0=20)
    prob = _tightened(base, mu=0.25, torque=9.0)
    opt = copy.copy(cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info")))
    opt.max_AL_iter = 4; opt.max_DDP_iter = 4
    x0 = workload.mhpc_batch(4)
    o1 = copy.copy(opt)
    ref = [oracle_solve(prob.deck, o1, x0[b])[1] for b in range(2)]            # factors 1
    opt.update_relax = relax; opt.update_ReB = weight
    got = _compare(cm, prob, opt, x0)
    # the update took effect: the cost history parts ways with the factor-1 run after the first outer iteration
    assert any(len(h) != len(r) or not np.allclose(h, r, rtol=1e-6) for (_, h), r in zip(got, ref))
    assert all(oi["outer_iter"] >= 2 for oi, _ in got[:2])


def test_mhpc_delta_below_its_floor_is_lifted_by_the_first_update(cm):
    """factors 1 but delta < delta_min: update_relax still applies the floor to violated elements (:79-82) - refused in round 1"""
    from cafe_mpc_b200 import workload
    prob = _tightened(cm.MHPCProblem(CSV), mu=0.25, torque=9.0, delta_scale=0.25)
    opt = copy.copy(cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info")))
    opt.max_AL_iter = 3; opt.max_DDP_iter = 4
    _compare(cm, prob, opt, workload.mhpc_batch(3))


@pytest.mark.parametrize("relax,weight", [(0.5, 0.2), (0.2, 4.0)])
def test_hkd_reb_update_matches_oracle(cm, relax, weight):
    from cafe_mpc_b200 import workload
    base = cm.HKDProblem(CSV)
    prob = _tightened(base, mu=0.15)
    opt = copy.copy(cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info")))
    opt.max_AL_iter = 4; opt.max_DDP_iter = 5
    opt.update_relax = relax; opt.update_ReB = weight
    x0 = workload.hkd_batch(base, 4)
    got = _compare(cm, prob, opt, x0)
    assert all(oi["outer_iter"] >= 2 for oi, _ in got)
