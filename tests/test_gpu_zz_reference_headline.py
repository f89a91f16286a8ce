"""The whole headline batch (BASELINE metric: 4096 MHPC trot problems) on the GPU against the record of the REFERENCE's OWN solver build:
tests/golden/ref_mhpc_headline.npz holds, for every problem of the SplitMix64 table, the iterations / line-search trials / regularisation steps and the
final cost that MHPCProblem + WBM + HSDDPSolver, compiled unchanged from the reference's sources (oracle/refbuild, tools/ref_headline.py), produced.
Runs last (file name): it needs nothing but the fixture at run time."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")


def test_headline_batch_takes_every_decision_like_the_reference_solver(cm):
    from cafe_mpc_b200 import workload
    ref = np.load(os.path.join(REPO, "tests/golden/ref_mhpc_headline.npz"))
    B = len(ref["final_cost"])
    assert B == 4096
    x0 = workload.mhpc_batch(B)
    assert hashlib.sha1(np.ascontiguousarray(x0).tobytes()).hexdigest() == str(ref["x0_sha1"])   # the table the reference was run on
    prob = cm.MHPCProblem(CSV)
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    s = cm.MultiPhaseDDP(prob, 0, B)
    s.set_initial_condition(x0)
    s.solve(opt)
    info = s.get_solver_info()
    s.close()
    got = np.array([[i["iter"], i["ls_iter_total"], i["reg_iter_total"]] for i in info], dtype=np.int64)
    bad = np.nonzero(np.any(got != ref["counters"].astype(np.int64), axis=1))[0]
    assert bad.size == 0, (bad[:8], got[bad[:8]], ref["counters"][bad[:8]])              # bit-exact decisions, all 4096 problems
    cost = np.array([i["cost"] for i in info])
    np.testing.assert_allclose(cost, ref["final_cost"], rtol=1e-9)                          # north star: 1e-9 relative
