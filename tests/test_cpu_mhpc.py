"""CPU suite for the MHPC (whole-body + single-rigid-body) path: phase-deck goldens, the Pinocchio-free rigid-body
model against the reference's only known-answer vectors and against its CasADi kinematic partials, the finite-difference
recipe of the reference's own test, and committed oracle goldens."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle_bindings import casadi_eval, oracle_dynamics, oracle_resetmap, oracle_solve, oracle_wb_dynamics

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSV = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")


@pytest.fixture(scope="module")
def mhpc(cm):
    return cm.MHPCProblem(CSV)


@pytest.fixture(scope="module")
def mhpc_impact(cm):
    return cm.MHPCProblem(CSV, k0=20)


@pytest.fixture(scope="module")
def mhpc_options(cm):
    return cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))


def test_mhpc_phase_schedule_golden(mhpc):
    """MHPC trot deck at t0 = 0 — WB0 h=11 (1,1,1,1); WB1 h=14 (0,1,1,0); SRB h=10."""
    ph = mhpc.phases()
    assert [(p.model, p.horizon) for p in ph] == [(1, 11), (1, 14), (2, 10)]
    assert [tuple(p.contact) for p in ph[:2]] == [(1, 1, 1, 1), (0, 1, 1, 0)]
    assert [p.next_model for p in ph] == [1, 2, -1]
    assert [p.n_td for p in ph] == [0, 0, 0]
    assert ph[0].dt == 0.01 and ph[2].dt == 0.05          # doubles in MHPCConfig (MHPCProblem.h:25-35), unlike the HKD float time step
    assert ph[2].t_offset == 0.25
    d = mhpc.deck.contents
    assert d.BG_alpha == 10.0 and d.hip_yaw == 3.1415 and d.n_records == 12 + 15 + 11
    # weights from cost_weights_regular.JSON and constraint_params_regular.info
    assert list(ph[0].q)[:6] == [0.0, 0.0, 10.0, 1.0, 2.0, 2.0] and list(ph[0].q)[6:9] == [1.0, 1.0, 1.0] and ph[0].q[35] == 0.01
    assert list(ph[0].w_footreg) == [20.0, 20.0, 1.0] and list(ph[0].w_swingvel) == [2.0, 2.0, 2.0]
    assert ph[0].reb_torque.delta == 1.0 and ph[0].reb_torque.eps == 0.01 and ph[0].reb_grf.eps == 0.05 and ph[0].mu == 0.6
    assert ph[2].r[0] == 0.01 and ph[2].h_min == 0.18 and ph[0].h_min == 0.20


def test_mhpc_impact_deck_golden(mhpc_impact):
    ph = mhpc_impact.phases()
    assert [(p.model, p.horizon) for p in ph] == [(1, 16), (1, 9), (2, 10)]
    assert tuple(ph[0].contact) == (0, 1, 1, 0) and tuple(ph[0].next_contact) == (1, 0, 0, 1)
    assert ph[0].n_td == 2 and tuple(ph[0].td_foot)[:2] == (0, 3)


# ---- the only numeric pin on the Pinocchio boundary in the reference: test/testKKTDynamics.cpp:95-121
QDD_REF = np.array([-6.3095, -4.2604, -14.1384, 22.9058, 17.9408, 39.8478, 90.4579, -65.9947, 66.0292, 138.4558, 22.0186, 7.8347,
                    -434.0716, 5.2086, 99.9593, -386.0737, -60.6831, -18.4982])
GRF_REF = np.array([5.6718, 3.3482, 4.9412, 9.5903, 5.4040, 6.7458, -40.7861, -21.8598, -24.2717, -37.9467, -20.9369, -24.4426])
FF_REF = np.array([0.0167, 0.0347, -9.8007, 3.0514, -0.7017, 4.1830, 0.8268, 0.6603, -5.2010, -0.2201, 1.3537, -4.9655, 1.0438, -1.0924,
                   -0.5153, 0.1417, -0.2949, -0.2164])


def test_wb_contact_dynamics_known_answers():
    """q = qd = 1, u = 0 (default BG_alpha = 10). The printed vectors are reproduced to their 4 decimals when the hip yaw
    is pi; with the 3.1415 of the shipped URDF the difference is 0.07 (the vectors predate that URDF edit)."""
    q = v = np.ones(18)
    u = np.zeros(12)
    qdd, grf = oracle_wb_dynamics(q, v, u, (1, 1, 1, 1), math.pi)
    assert np.max(np.abs(qdd - QDD_REF)) < 6e-5 and np.max(np.abs(grf - GRF_REF)) < 6e-5  # rounding of 4 printed decimals
    qdd, grf = oracle_wb_dynamics(q, v, u, (0, 0, 0, 0), math.pi)
    assert np.max(np.abs(qdd - FF_REF)) < 6e-5 and np.all(grf == 0)
    qdd, grf = oracle_wb_dynamics(q, v, u, (1, 1, 1, 1), 3.1415)
    assert 0.05 < np.linalg.norm(qdd - QDD_REF) < 0.1


@pytest.fixture(scope="module")
def gen_lib():
    out = os.path.join(REPO, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libgen_host.so")
    src = os.path.join(REPO, "tests", "gen_host.cpp")
    deps = [src] + [os.path.join(REPO, "cafe_mpc_b200/csrc", f) for f in ("gen/wb_gen.h", "gen/wb_leg_gen.h", "wb_leg_tables.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    return C.CDLL(so)


def gen_wb(lib, name, ins, out_shapes):
    ins = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in ins]
    outs = [np.zeros(int(np.prod(s))) for s in out_shapes]
    pin = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
    assert lib.gen_eval_wb(name.encode(), pin, pout) == 0
    return [o.reshape(s, order="F") for o, s in zip(outs, out_shapes)]


def gen_wbl(lib, name, ins, out_shapes):
    ins = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in ins]
    outs = [np.zeros(int(np.prod(s))) for s in out_shapes]
    pin = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
    assert lib.gen_eval_wbl(name.encode(), pin, pout) == 0
    return [o.reshape(s, order="F") for o, s in zip(outs, out_shapes)]


def test_leg_generic_routines_match_per_leg_routines_and_reference_casadi(gen_lib):
    """The leg-generic routines of the running whole-body knots (gen/wb_leg_gen.h: one routine for all four legs, the mirrored leg
    constants as data; derivatives in forward mode, a few directions per routine), assembled through their compact-slot tables like
    the device kernels do, equal the per-leg routines of gen/wb_gen.h and the reference's CasADi kinematic partials."""
    rng = np.random.default_rng(23)
    for _ in range(4):
        q, v, a, F = rng.normal(size=18) * .6, rng.normal(size=18), rng.normal(size=18) * 3, rng.normal(size=12) * 20
        shapes_t = [(18,), (18, 18), (12, 18), (12,), (12,), (12,)]
        new = gen_wbl(gen_lib, "wbl_terms", [q, v], shapes_t)
        old = gen_wb(gen_lib, "wb_terms_pieces", [q, v], shapes_t)
        for x, y in zip(new, old):
            np.testing.assert_allclose(x, y, rtol=0, atol=1e-14 * max(1.0, np.max(np.abs(y))))
        shapes_d = [(18, 18), (18, 18), (12, 18), (12, 18), (12, 18), (18, 18)]
        dq, dv, dvq, daq, dav, djtf = gen_wbl(gen_lib, "wbl_derivs", [q, v, a, F], shapes_d)
        odq, odv = gen_wb(gen_lib, "wb_rnea_derivs", [q, v, a], [(18, 18)] * 2)
        odvq, odaq, odav, odjtf = gen_wb(gen_lib, "wb_kin_partials", [q, v, a, F], [(12, 18)] * 3 + [(18, 18)])
        for x, y in ((dq, odq), (dv, odv), (dvq, odvq), (daq, odaq), (dav, odav), (djtf, odjtf)):
            np.testing.assert_allclose(x, y, rtol=0, atol=2e-13 * max(1.0, np.max(np.abs(y))))
        rv = casadi_eval("footVelPartialDq", [q, v], [(3, 18)] * 4)
        raq = casadi_eval("footAccPartialDq", [q, v, a], [(3, 18)] * 4)
        rav = casadi_eval("footAccPartialDv", [q, v, a], [(3, 18)] * 4)
        rf = casadi_eval("footForcePartialDq", [q, F], [(18, 18)] * 4)
        for f in range(4):
            np.testing.assert_allclose(dvq[3 * f:3 * f + 3], rv[f], rtol=0, atol=1e-13)
            np.testing.assert_allclose(daq[3 * f:3 * f + 3], raq[f], rtol=0, atol=2e-12)
            np.testing.assert_allclose(dav[3 * f:3 * f + 3], rav[f], rtol=0, atol=1e-12)
        np.testing.assert_allclose(djtf, sum(rf), rtol=0, atol=1e-12)


def test_generated_kinematic_partials_match_reference_casadi(gen_lib):
    """The symbolic whole-body tree (hip yaw = pi) reproduces footVelPartialDq / footAccPartialDq / footAccPartialDv /
    footForcePartialDq of the reference (106k generated ops) at machine precision with 11.6k ops."""
    rng = np.random.default_rng(11)
    for _ in range(4):
        q, v, a, F = rng.normal(size=18) * .6, rng.normal(size=18), rng.normal(size=18) * 3, rng.normal(size=12) * 20
        dvq, daq, dav, djtf = gen_wb(gen_lib, "wb_kin_partials", [q, v, a, F], [(12, 18)] * 3 + [(18, 18)])
        rv = casadi_eval("footVelPartialDq", [q, v], [(3, 18)] * 4)
        raq = casadi_eval("footAccPartialDq", [q, v, a], [(3, 18)] * 4)
        rav = casadi_eval("footAccPartialDv", [q, v, a], [(3, 18)] * 4)
        rf = casadi_eval("footForcePartialDq", [q, F], [(18, 18)] * 4)
        for f in range(4):
            np.testing.assert_allclose(dvq[3 * f:3 * f + 3], rv[f], rtol=0, atol=1e-13)
            np.testing.assert_allclose(daq[3 * f:3 * f + 3], raq[f], rtol=0, atol=2e-12)
            np.testing.assert_allclose(dav[3 * f:3 * f + 3], rav[f], rtol=0, atol=1e-12)
        np.testing.assert_allclose(djtf, sum(rf), rtol=0, atol=1e-12)
        (dvq2,) = gen_wb(gen_lib, "wb_footvel_partial", [q, v], [(12, 18)])
        np.testing.assert_allclose(dvq2, dvq, rtol=0, atol=1e-14)


def test_generated_dynamics_terms_match_oracle_model(gen_lib):
    """M, nle, J, Jdot v from the generated (symbolic) routine give the same KKT solution as the oracle's numeric
    Newton-Euler implementation (independent code, same hip yaw 3.1415)."""
    rng = np.random.default_rng(5)
    for contact in ((1, 1, 1, 1), (0, 1, 1, 0), (0, 0, 0, 0)):
        q, v, u = rng.normal(size=18) * .5, rng.normal(size=18), rng.normal(size=12) * 5
        nle, Ml, J, gam, pf, vf = gen_wb(gen_lib, "wb_terms", [q, v], [(18,), (18, 18), (12, 18), (12,), (12,), (12,)])
        M = np.tril(Ml) + np.tril(Ml, -1).T
        rows = [3 * f + r for f in range(4) if contact[f] for r in range(3)]
        Jc = J[rows]
        tau = np.concatenate([np.zeros(6), u])
        g = gam[rows] + 2 * 10.0 * vf[rows]
        if rows:
            K = np.block([[M, -Jc.T], [Jc, np.zeros((len(rows), len(rows)))]])
            sol = np.linalg.solve(K, np.concatenate([tau - nle, -g]))
        else:
            sol = np.linalg.solve(M, tau - nle)
        qdd, grf = oracle_wb_dynamics(q, v, u, contact, 3.1415)
        np.testing.assert_allclose(sol[:18], qdd, rtol=1e-9, atol=1e-9)
        if rows:
            np.testing.assert_allclose(sol[18:], grf[rows], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(J @ v, vf, atol=1e-13)  # v_foot = J v


def test_generated_pieces_sum_to_the_whole(gen_lib):
    """The device path evaluates M, nle, the foot kinematics and the RNEA derivatives as trunk + one piece per leg (smaller live
    sets): the composed pieces equal the monolithic generated routine, and the RNEA-derivative pieces equal central differences of
    tau = M qdd + nle built from it."""
    rng = np.random.default_rng(23)
    shapes = [(18,), (18, 18), (12, 18), (12,), (12,), (12,)]
    for _ in range(3):
        q, v, a = rng.normal(size=18) * .5, rng.normal(size=18), rng.normal(size=18) * 3
        whole = gen_wb(gen_lib, "wb_terms", [q, v], shapes)
        parts = gen_wb(gen_lib, "wb_terms_pieces", [q, v], shapes)
        for w, p_ in zip(whole, parts):
            np.testing.assert_allclose(p_, w, rtol=0, atol=1e-13 * max(1.0, np.abs(w).max()))

        def tau(q_, v_):
            nle, Ml = gen_wb(gen_lib, "wb_terms", [q_, v_], shapes)[:2]
            return (np.tril(Ml) + np.tril(Ml, -1).T) @ a + nle
        dq, dv = gen_wb(gen_lib, "wb_rnea_derivs", [q, v, a], [(18, 18), (18, 18)])
        e = 1e-6
        for i in range(18):
            d = np.zeros(18); d[i] = e
            np.testing.assert_allclose(dq[:, i], (tau(q + d, v) - tau(q - d, v)) / (2 * e), atol=2e-6 * max(1.0, np.abs(dq).max()))
            np.testing.assert_allclose(dv[:, i], (tau(q, v + d) - tau(q, v - d)) / (2 * e), atol=2e-6 * max(1.0, np.abs(dv).max()))


def _fd_check(prob, phase, x, u, n, tol):
    xn, y, A, B, Cm, D = oracle_dynamics(prob.deck, phase, 3, x, u, partials=True)
    e = 1e-6
    Afd = np.zeros_like(A); Cfd = np.zeros_like(Cm); Bfd = np.zeros_like(B); Dfd = np.zeros_like(D)
    for i in range(n):
        d = np.zeros(n); d[i] = e
        xp, yp = oracle_dynamics(prob.deck, phase, 3, x + d, u); xm, ym = oracle_dynamics(prob.deck, phase, 3, x - d, u)
        Afd[:, i] = (xp - xm) / (2 * e)
        if Cm.size: Cfd[:, i] = (yp - ym) / (2 * e)
    for i in range(12):
        d = np.zeros(12); d[i] = e
        xp, yp = oracle_dynamics(prob.deck, phase, 3, x, u + d); xm, ym = oracle_dynamics(prob.deck, phase, 3, x, u - d)
        Bfd[:, i] = (xp - xm) / (2 * e)
        if D.size: Dfd[:, i] = (yp - ym) / (2 * e)
    for an, fd in ((A, Afd), (B, Bfd), (Cm, Cfd), (D, Dfd)):
        if an.size:
            assert np.max(np.abs(an - fd)) < tol * max(1.0, np.max(np.abs(an)))


def test_wb_dynamics_partials_finite_differences(cm):
    """The reference's own recipe (test/testKKTDynamics.cpp:39-93): A, B, C against differences with eps = 1e-6
    (isApprox 1e-4 there), impact Jacobian against differences. With the hip yaw of the shipped URDF (3.1415) the analytic
    partials carry the reference's pi-vs-3.1415 inconsistency and agree to ~1e-4; with a consistent
    yaw (pi everywhere) the same formulas agree with central differences to 1e-6."""
    from cafe_mpc_b200 import workload
    rng = np.random.default_rng(2)
    for yaw, tol in ((3.1415, 5e-4), (math.pi, 2e-6)):
        prob = cm.MHPCProblem(CSV)
        prob.deck.contents.hip_yaw = yaw
        for phase in (0, 1, 2):
            n = 36 if phase < 2 else 12
            x = (workload.MHPC_NOMINAL + rng.normal(size=36) * 0.1) if n == 36 else np.array([0, 0, .25, 0, 0, 0, .3, 0, 0, 0, 0, 0]) + rng.normal(size=12) * .05
            u = rng.normal(size=12) * 2
            _fd_check(prob, phase, x, u, n, tol)
    # impact + WB->WB reset (two feet land at the end of phase 0 of the k0 = 20 deck), consistent yaw
    prob = cm.MHPCProblem(CSV, k0=20)
    prob.deck.contents.hip_yaw = math.pi
    x = workload.MHPC_NOMINAL + rng.normal(size=36) * 0.1
    xn, Px = oracle_resetmap(prob.deck, 0, x, jac=True)
    assert xn.shape == (36,) and np.array_equal(xn[:18], x[:18])
    fd = np.zeros((36, 36))
    for i in range(36):
        d = np.zeros(36); d[i] = 1e-6
        fd[:, i] = (oracle_resetmap(prob.deck, 0, x + d) - oracle_resetmap(prob.deck, 0, x - d)) / 2e-6
    # the reference's dv+/dq uses its segment<3>(i) impulse scatter (WBM.cpp:454), so only dv+/dv is exactly the derivative
    assert np.max(np.abs(Px[:, 18:] - fd[:, 18:])) < 1e-6
    assert np.max(np.abs(Px[:18] - fd[:18])) < 1e-9
    assert np.max(np.abs(Px[18:, :18] - fd[18:, :18])) < 0.5 * max(1.0, np.max(np.abs(fd[18:, :18])))  # same structure, quirk-limited accuracy
    # WB -> SRB projection keeps (pos, eul) and (v, eulrate)
    p0 = cm.MHPCProblem(CSV)
    xs = oracle_resetmap(p0.deck, 1, x)
    assert np.array_equal(xs, np.concatenate([x[:6], x[18:24]]))


def test_oracle_mhpc_matches_committed_golden(mhpc, mhpc_impact, mhpc_options):
    from cafe_mpc_b200 import workload
    g = np.load(os.path.join(REPO, "tests/golden/mhpc_trot.npz"))
    x0 = workload.mhpc_batch(4)
    for key, prob in (("k0", mhpc), ("k20", mhpc_impact)):
        for b in (0, 3):
            info, hist, trace, sol = oracle_solve(prob.deck, mhpc_options, x0[b])
            assert [info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] == list(g["%s_counts_%d" % (key, b)])
            np.testing.assert_allclose(hist[:, 0], g["%s_hist_%d" % (key, b)][:, 0], rtol=1e-9)
            np.testing.assert_allclose(sol, g["%s_sol_%d" % (key, b)], rtol=0, atol=1e-8 * np.abs(sol).max())


# ---- BASELINE config 4: MHPC running barrel roll (Reference/Data/running_br, barrel cost weights / constraint parameters)
def test_barrel_roll_phase_schedule_golden(cm):
    """SURVEY.md §8 phase table: t0 = 0 -> WB0 h=6 (1,1,1,1); WB1 h=15 (0,1,0,1); WB2 h=4 flight; SRB h=10. Start offset 205 = inside
    the roll's flight: 22 flight knots, 4-foot landing (four touchdown constraints, impact), 3 stance knots, SRB h=10."""
    from cafe_mpc_b200 import workload
    prob0 = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=0)   # owns the deck the phase views point into
    p0 = prob0.phases()
    assert [(p.model, p.horizon, tuple(p.contact)) for p in p0] == [(1, 6, (1, 1, 1, 1)), (1, 15, (0, 1, 0, 1)), (1, 4, (0, 0, 0, 0)), (2, 10, (0, 0, 0, 0))]
    assert [p.n_td for p in p0] == [0, 0, 0, 0]
    prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=workload.BARREL_K0_IMPACT)
    p1 = prob.phases()
    assert [(p.model, p.horizon, tuple(p.contact)) for p in p1] == [(1, 22, (0, 0, 0, 0)), (1, 3, (1, 1, 1, 1)), (2, 10, (0, 0, 0, 0))]
    assert p1[0].n_td == 4 and sorted(tuple(p1[0].td_foot)) == [0, 1, 2, 3] and tuple(p1[0].next_contact) == (1, 1, 1, 1)
    # barrel weights / parameters were picked up
    assert list(p1[0].q)[:6] == [0.0, 0.0, 15.0, 5.0, 5.0, 5.0] and list(p1[0].w_footreg) == [0.0, 0.0, 0.0]
    assert p1[0].reb_torque.delta == 0.1 and p1[0].reb_torque.eps == 0.1 and p1[0].reb_grf.delta == 0.1 and p1[0].al_td.sigma == 10
    x0 = workload.reference_state(prob)
    assert abs(x0[5]) > 3.0   # mid-roll: the roll angle of the tracked motion is beyond pi


def test_oracle_barrel_roll_matches_committed_golden(cm, mhpc_options):
    from cafe_mpc_b200 import workload
    g = np.load(os.path.join(REPO, "tests/golden/mhpc_barrel.npz"))
    for key, k0 in (("k0", 0), ("k205", workload.BARREL_K0_IMPACT)):
        prob = cm.MHPCProblem(workload.BARREL_CSV, mhpc_config=workload.BARREL_CONFIG, k0=k0)
        x0 = workload.barrel_batch(prob, 4)
        assert np.array_equal(x0, g[key + "_x0"])
        info, hist, trace, sol = oracle_solve(prob.deck, mhpc_options, x0[0])
        assert [info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] == list(g["%s_counts_0" % key])
        np.testing.assert_allclose(hist[:, 0], g["%s_hist_0" % key][:, 0], rtol=1e-9)
        np.testing.assert_allclose(sol, g["%s_sol_0" % key], rtol=0, atol=1e-8 * np.abs(sol).max())


def test_per_problem_reference_set_keeps_the_schedule(cm, mhpc_impact, mhpc_options):
    """workload.speed_command_references: same contact flags in every record, shifted targets; the oracle run on a deck that
    carries another problem's records gives another solution (what the GPU per-problem path is checked against)."""
    from cafe_mpc_b200 import workload
    from oracle_bindings import deck_with_references
    refs = workload.speed_command_references(mhpc_impact, 4)
    base = mhpc_impact.reference_records()
    assert refs.shape == (4,) + base.shape and np.array_equal(refs[0], base)
    assert np.array_equal(refs[:, :, 99:103], np.tile(base[:, 99:103], (4, 1, 1)))       # CAFE_REF_CONTACT
    assert np.abs(refs[2, :, 18] - base[:, 18])[:26].min() > 0                            # WB forward-velocity target moved
    x0 = workload.mhpc_batch(1)[0]
    i0, _, _, s0 = oracle_solve(mhpc_impact.deck, mhpc_options, x0)
    dk, keep = deck_with_references(mhpc_impact.deck, refs[2])
    i2, _, _, s2 = oracle_solve(dk, mhpc_options, x0)
    assert i0["status"] == 0 and i2["status"] == 0 and abs(i0["cost"] - i2["cost"]) > 1e-6 and not np.allclose(s0, s2)


def test_receding_horizon_shift_and_oracle_warm_start(cm, mhpc_options):
    """cafe_mpc_b200/mpc.py: shifting a solution by two knots keeps the overlapping knots, pads the tail like push_back_state, drops an
    exhausted first phase, and the oracle started from that guess under the run-time caps ends far closer to feasible than a cold start."""
    import copy
    from cafe_mpc_b200 import mpc, workload
    ort = copy.copy(mhpc_options)
    ort.max_AL_iter = mhpc_options.max_AL_iter_runtime; ort.max_DDP_iter = mhpc_options.max_DDP_iter_runtime
    x0 = workload.mhpc_batch(2)[1]
    p0 = cm.MHPCProblem(CSV, k0=10)      # WB0 h=1: the next shift removes it, and a new phase opens at the tail
    p1 = cm.MHPCProblem(CSV, k0=12)
    assert [p.horizon for p in p0.phases()] == [1, 24, 10] and [p.horizon for p in p1.phases()] == [24, 1, 10]
    i0, _, _, sol = oracle_solve(p0.deck, mhpc_options, x0)
    old = cm.unpack_solution(p0.deck, sol)
    g = mpc.shift_guess(p0, 10, p1, 12, old)
    np.testing.assert_array_equal(g[0]["Xbar"][:24], old[1]["Xbar"][1:25])         # absolute knots 12..35 of the old second phase
    np.testing.assert_array_equal(g[0]["Ubar"][:23], old[1]["Ubar"][1:24])
    np.testing.assert_array_equal(g[0]["K"][:23], old[1]["K"][1:24])
    np.testing.assert_array_equal(g[0]["Xbar"][24], old[1]["Xbar"][24])            # push_back_state(X.back())
    assert not g[0]["Ubar"][23].any() and not g[0]["K"][23].any()
    np.testing.assert_array_equal(g[1]["Xbar"], p1.reference_records()[25:27, :36])  # the phase the old plan did not have
    np.testing.assert_array_equal(g[2]["Xbar"], old[2]["Xbar"])                    # SRB phase is not shifted (dt_mpc < dt_srb)
    packed = mpc.pack_solution(p1, g)
    back = cm.unpack_solution(p1.deck, packed)
    for a, b_ in zip(g, back):
        for k in ("Xbar", "Ubar", "K"):
            np.testing.assert_array_equal(a[k], b_[k])
    x1 = mpc.state_at(p0, old, 2)
    iw, _, _, _ = oracle_solve(p1.deck, ort, x1, guess=packed)
    ic, _, _, _ = oracle_solve(p1.deck, ort, x1)
    assert iw["iter"] <= 4 and ic["iter"] <= 4 and iw["feas"] < 0.2 * ic["feas"]


def _mask_bits(words, rows, cols):
    """[rows, cols] boolean matrix from a column-major bit mask (bit i + rows * j)."""
    m = np.zeros((rows, cols), dtype=bool)
    for e in range(rows * cols):
        if (int(words[e >> 6]) >> (e & 63)) & 1:
            m[e % rows, e // rows] = True
    return m


def _pattern(cm, prob, phase, knot, which, rows, cols):
    from cafe_mpc_b200.lib import lib
    out = (C.c_ulonglong * 21)()
    n = lib.cafe_deck_lq_pattern(prob.deck, phase, knot, which, out)
    assert n > 0
    return _mask_bits(list(out), rows, cols)


@pytest.mark.parametrize("k0", [0, 20])
def test_structural_lxx_patterns_cover_the_oracle(cm, mhpc_options, k0):
    """The backward sweep fetches the whole-body lxx only at the structural non-zeros of cafe_deck_lq_pattern (diagonal, base block,
    per-foot blocks from the knot's contact flags). Safety net: every non-zero the oracle produces (tracking, foot placement, swing-foot
    position / velocity costs, ReB joint-limit and height terms) lies inside the mask of its knot - in the stance / swing mixes of both
    whole-body phases and after an impact - and luu is diagonal, lyy block diagonal per foot, as the sweep assumes."""
    import copy
    from cafe_mpc_b200 import workload
    prob = cm.MHPCProblem(CSV, k0=k0)
    o1 = copy.copy(mhpc_options)
    o1.max_DDP_iter = 2; o1.max_AL_iter = 1; o1.cost_thresh = 1e30; o1.dynamics_feas_thresh = 1e30
    x0 = workload.mhpc_batch(3)[2]
    oracle_solve(prob.deck, o1, x0)
    from oracle_bindings import oracle_get
    for ph, p in enumerate(prob.phases()):
        if p.model != 1:
            continue
        lxx = oracle_get("lxx", ph).reshape(p.horizon, 36, 36).transpose(0, 2, 1)   # column-major per knot -> [k][i][j]
        luu = oracle_get("luu", ph).reshape(p.horizon, 12, 12)
        lyy = oracle_get("lyy", ph).reshape(p.horizon, 12, 12)
        nnz = []
        for k in range(p.horizon):
            mask = _pattern(cm, prob, ph, k, 2, 36, 36)
            assert np.array_equal(mask, mask.T)
            assert not np.any((lxx[k] != 0) & ~mask), (ph, k, np.argwhere((lxx[k] != 0) & ~mask)[:4])
            nnz.append(int(mask.sum()))
            assert np.count_nonzero(luu[k] - np.diag(np.diag(luu[k]))) == 0
            blk = np.kron(np.eye(4), np.ones((3, 3))) > 0
            assert not np.any((lyy[k] != 0) & ~blk)
        assert 100 < min(nnz) and max(nnz) < 750   # a third to a half of the 1296 entries


# ---- next tier: LocoProblem, the whole-body-only locomotion trajectory optimisation (Locomotion/Loco_TO.cpp)
def test_loco_problem_deck_golden(cm):
    """loco_config.info: 1.0 s whole-body plan on the flypace reference, no SRB tail. LocoProblem::create_problem_one_phase
    (LocoProblem.cpp:29-84) attaches the torque and GRF barriers only; add_tconstr_one_phase still adds the touchdown constraints."""
    from cafe_mpc_b200 import workload
    prob = cm.LocoProblem()
    ph = prob.phases()
    g = np.load(os.path.join(REPO, "tests/golden/loco_flypace.npz"))
    assert np.array_equal(np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td, p.no_joint_limit, p.no_min_height] for p in ph]), g["phases"])
    assert all(p.model == 1 and p.no_joint_limit == 1 and p.no_min_height == 1 for p in ph) and sum(p.horizon for p in ph) == 100
    assert [p.next_model for p in ph] == [1] * (len(ph) - 1) + [-1]
    assert [(p.horizon, tuple(p.contact)) for p in ph[:3]] == [(6, (1, 1, 1, 1)), (15, (0, 1, 0, 1)), (10, (0, 0, 0, 0))]
    assert ph[2].n_td == 2 and tuple(ph[2].next_contact) == (1, 0, 1, 0)
    # loco_constraint_params.info / loco_cost_weights.JSON were picked up
    assert ph[0].reb_grf.delta == 0.2 and ph[0].reb_torque.delta == 0.1 and ph[0].reb_torque.eps == 0.01 and ph[0].al_td.sigma == 20
    opt = cm.load_hsddp_setting(workload.LOCO_DDP_SETTING)
    assert (opt.max_AL_iter, opt.max_DDP_iter) == (30, 10)
    # an MHPC deck keeps all four barriers
    assert all(p.no_joint_limit == 0 and p.no_min_height == 0 for p in cm.MHPCProblem(CSV).phases())


def test_oracle_loco_matches_committed_golden(cm):
    from cafe_mpc_b200 import workload
    g = np.load(os.path.join(REPO, "tests/golden/loco_flypace.npz"))
    prob = cm.LocoProblem()
    opt = cm.load_hsddp_setting(workload.LOCO_DDP_SETTING)
    x0 = workload.mhpc_batch(4)
    assert np.array_equal(x0, g["x0"])
    info, hist, trace, sol = oracle_solve(prob.deck, opt, x0[0], cap=320)
    assert [info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] == list(g["counts_0"])
    np.testing.assert_allclose(hist[:, 0], g["hist_0"][:, 0], rtol=1e-9)
    np.testing.assert_allclose(sol, g["sol_0"], rtol=0, atol=1e-8 * np.abs(sol).max())
    assert info["outer_iter"] > 1   # the touchdown constraints of the flight phases drove the AL loop


def test_oracle_joint_and_height_barriers_matter(cm, mhpc, mhpc_options):
    """Dropping the joint-limit / min-height barriers from an MHPC deck changes the cost the oracle reports (the flags are live)."""
    from cafe_mpc_b200 import workload
    from cafe_mpc_b200._ctypes_defs import Deck
    x0 = workload.mhpc_batch(2)[1]
    i0, _, _, _ = oracle_solve(mhpc.deck, mhpc_options, x0)
    d2 = Deck.from_buffer_copy(mhpc.deck.contents)
    for i in range(d2.n_phases):
        d2.phase[i].no_joint_limit = 1; d2.phase[i].no_min_height = 1
    i1, _, _, _ = oracle_solve(C.pointer(d2), mhpc_options, x0)
    assert i0["cost"] != i1["cost"] and i0["status"] == i1["status"] == 0   # (-log barriers of well-satisfied constraints are negative: the cost goes up)


# ---- next tier: the in-place barrel roll of BarrelRoll/BarrelRollTO.cpp
def test_barrel_to_deck_and_guess(cm):
    """Six hand-scheduled whole-body phases (switching times 0, .12, .33, .75, .90, 1.10, 1.25 s), fixed desired states, per-phase
    weights, joint-speed barrier on every phase, four-foot touchdown constraints after both flight phases, interpolated initial states."""
    from cafe_mpc_b200 import mpc, workload
    prob = cm.BarrelRollProblem()
    ph = prob.phases()
    g = np.load(os.path.join(REPO, "tests/golden/barrel_to.npz"))
    assert np.array_equal(np.array([[p.model, p.horizon] + list(p.contact) + [p.n_td, p.joint_speed_limit] for p in ph]), g["phases"])
    assert [p.horizon for p in ph] == [12, 21, 42, 15, 20, 15]
    assert [tuple(p.contact) for p in ph] == [(1, 1, 1, 1), (0, 1, 0, 1), (0, 0, 0, 0), (1, 1, 1, 1), (0, 0, 0, 0), (1, 1, 1, 1)]
    assert [p.n_td for p in ph] == [0, 0, 4, 0, 4, 0] and all(p.joint_speed_limit == 1 and p.no_joint_limit == 0 and p.h_min == 0.13 for p in ph)
    assert ph[0].dt == 0.01 and (ph[0].jointvel_lb, ph[0].jointvel_ub) == (-20.0, 20.0) and ph[0].reb_jointvel.delta == 0.1
    assert list(ph[0].w_footreg) == [0, 0, 0] and list(ph[0].w_swingpos) == [0, 0, 0] and list(ph[0].w_tdvel) == [0, 0, 0]
    # weights of cost_phase_1 / cost_phase_2 (br_cost_weights.JSON), desired states (BarrelRollTO.cpp:277-339)
    assert list(ph[0].q)[:6] == [0.0, 5.0, 10.0, 2.0, 2.0, 2.0] and ph[0].r[0] == 0.2 and ph[1].r[0] == 0.05 and list(ph[1].qf)[18:24] == [1.0, 1.0, 5.0, 1.0, 1.0, 5.0]
    rec = prob.reference_records()
    x1 = rec[ph[1].knot_offset, :36]
    assert np.allclose(x1[:6], [0, -0.25, 0.33, 0, 0, 0.5 * math.pi]) and np.allclose(x1[6:12], [math.pi / 6, -1.0, 2.0, -math.pi / 5, -0.5, 1.0]) and x1[23] == 3.0 * math.pi
    assert np.array_equal(rec[ph[5].knot_offset, :36], rec[ph[4].knot_offset, :36]) and rec[ph[3].knot_offset, 5] == 2 * math.pi and rec[ph[3].knot_offset, 23] == 0
    # interpolated initial states: phase 0 starts at x0 and ends (float time arithmetic) next to xf_des[0]; controls and gains are zero
    x0 = workload.mhpc_batch(4)
    gs = prob.initial_guess(x0)
    parts = mpc.unpack_batch(prob, gs)
    assert np.array_equal(parts[0]["Xbar"][:, 0], x0)
    assert np.allclose(parts[0]["Xbar"][2, -1], rec[0, :36], atol=1e-6) and np.allclose(parts[3]["Xbar"][1, 0], rec[ph[2].knot_offset, :36], atol=1e-12)
    assert all(not r["Ubar"].any() and not r["K"].any() for r in parts)
    assert np.array_equal(np.concatenate([r["Xbar"][0] for r in parts]), g["guess_xbar_0"])
    # hand check of one interpolated knot with the reference's float arithmetic: s = float(t) / float(dur), t accumulated in float
    t = np.float32(0.0)
    for _ in range(5):
        t = np.float32(np.float64(t) + 0.01)
    s = float(t / np.float32(0.33 - 0.12))
    xa, xb = rec[0, :36], rec[ph[1].knot_offset, :36]
    assert np.array_equal(parts[1]["Xbar"][0, 5], xa + (xb - xa) * s)


def test_oracle_barrel_to_matches_committed_golden(cm):
    from cafe_mpc_b200 import workload
    g = np.load(os.path.join(REPO, "tests/golden/barrel_to.npz"))
    prob = cm.BarrelRollProblem()
    opt = cm.load_hsddp_setting(workload.BARREL_TO_DDP_SETTING)
    assert (opt.max_AL_iter, opt.max_DDP_iter) == (30, 10)
    opt.max_AL_iter = 5   # the golden's caps (the full caps take ~30 s per problem on the CPU)
    x0 = workload.mhpc_batch(4)
    info, hist, trace, sol = oracle_solve(prob.deck, opt, x0[0], cap=320, guess=prob.initial_guess(x0)[0])
    assert [info[k] for k in ("status", "iter", "ls_iter_total", "reg_iter_total", "outer_iter", "n_hist")] == list(g["counts_0"])
    np.testing.assert_allclose(hist[:, 0], g["hist_0"][:, 0], rtol=1e-9)
    np.testing.assert_allclose(sol, g["sol_0"], rtol=0, atol=1e-8 * np.abs(sol).max())
    assert hist[-1, 0] < 0.05 * hist[0, 0]   # 3 635 -> 125: the roll is being found


# ------------------------------------------------------------------ the deck against an independent reader of the input files
@pytest.mark.parametrize("k0", [0, 8, 12, 20, 36])
def test_mhpc_deck_equals_an_independent_reading_of_the_input_files(cm, k0):
    """Every field of every CafePhase and every per-knot reference record of the MHPC trot deck equals what tests/deck_check.py derives
    from quad_reference.csv / mhpc_config.info / cost_weights / constraint_params by following the reference's set-up code with the
    reference's float / double types (no code shared with csrc/host or oracle/): a wrong weight, reference index, ReB / AL parameter,
    touchdown foot or time step in the deck builders fails here, where the GPU-vs-oracle tests (same deck on both sides) cannot see it."""
    from deck_check import expected_mhpc_deck
    root = os.path.join(REPO, "data")
    exp = expected_mhpc_deck(CSV, os.path.join(root, "MHPC/settings/mhpc_config.info"), root, k0)
    prob = cm.MHPCProblem(CSV, k0=k0)
    d = prob.deck.contents
    assert d.n_phases == len(exp)
    ref = prob.reference_records()
    off = 0
    for i, e in enumerate(exp):
        p = d.phase[i]
        n, m = (36, 12) if e["model"] == 1 else (12, 12)
        assert (p.model, p.horizon, p.knot_offset, p.next_model) == (e["model"], e["horizon"], off, e["next_model"]), i
        assert p.dt == e["dt"] and p.t_offset == float(e["t_offset"])
        assert tuple(p.contact) == e["contact"] and tuple(p.next_contact) == tuple(e["next_contact"])
        assert p.n_td == len(e["td_foot"]) and list(p.td_foot)[:p.n_td] == e["td_foot"]
        assert list(p.q)[:n] == [float(v) for v in e["q"]] and list(p.qf)[:n] == [float(v) for v in e["qf"]] and list(p.r)[:m] == e["r"]
        for name, val in e["reb"].items():
            r = getattr(p, name)
            assert (r.delta, r.delta_min, r.eps) == val, (i, name)
        if e["model"] == 1:
            assert list(p.w_footreg) == e["w_footreg"] and list(p.w_swingpos) == e["w_swingpos"] and list(p.w_swingvel) == e["w_swingvel"]
            assert (p.al_td.lambda_, p.al_td.sigma, p.al_td.sigma_max) == e["al_td"]
            assert p.has_reset == 1 and (p.mu, p.h_min, p.torque_limit) == (0.6, 0.20, 17.0)     # MHPCConstraint.cpp:11,77; MHPCConstraint.h:148
        else:
            assert p.has_reset == 0 and p.h_min == 0.18                                            # MHPCConstraint.h:199
        for k, r in enumerate(e["records"]):
            g = ref[off + k]
            assert np.array_equal(g[0:n], r["xr"]), (i, k)
            assert np.array_equal(g[36:36 + 12], r["ur"]) and np.array_equal(g[60:60 + len(r["yr"])], r["yr"])
            assert np.array_equal(g[72:84], r["pf"]) and np.array_equal(g[84:87], r["pcom"]) and np.array_equal(g[87:99], r["vf"])
            assert np.array_equal(g[99:103], r["contact"]) and np.array_equal(g[103:115], r["qJ"])
        off += e["horizon"] + 1
    assert d.n_records == off


# ------------------------------------------------------------------ the MPC update against an independent restatement of the reference's deque operations
def test_mpc_update_chain_equals_an_independent_restatement_of_the_reference_update(cm):
    """Forty consecutive MPC updates (start offsets 0, 2, ... 80: phases vanish at the front, grow and open at the tail) walked with
    tests/update_check.py - the reference's own sequence of deque operations (MHPCProblem::update / update_WB_plan, QuadReference::step,
    Trajectory::pop_front / push_back_state, quirk 14), no code shared with the product - against the three product pieces that stand for it:
    the deck re-cut at the new start offset (cafe_deck_build_mhpc + cafe_deck_mark_mpc_update: horizons, contacts, touchdown feet, which tail
    phase is single shooting) and the shifted guess (cafe_mpc_b200/mpc.py, which the device shift is held bit-identical to by the GPU tests)."""
    from cafe_mpc_b200 import mpc
    from update_check import MHPCPlan
    root = os.path.join(REPO, "data")
    plan = MHPCPlan(CSV, os.path.join(root, "MHPC/settings/mhpc_config.info"), root, 0)
    rng = np.random.default_rng(7)
    prev, k_prev = cm.MHPCProblem(CSV, k0=0), 0
    opened = removed = 0
    for step in range(40):
        # a tagged "solution" of the previous problem (random numbers: every knot recognisable)
        old = []
        for ph in prev.phases():
            n, m = (36, 12) if ph.model == 1 else (12, 12)
            old.append({"Xbar": rng.standard_normal((ph.horizon + 1, n)), "Ubar": rng.standard_normal((ph.horizon, m)), "K": rng.standard_normal((ph.horizon, m, n))})
        plan.load_solution(old)
        n_before = len(plan.horizon)
        nsteps = plan.update()
        assert nsteps == 2
        k_new = k_prev + nsteps
        new = cm.MHPCProblem(CSV, k0=k_new, mpc_update_nsteps=nsteps)
        wb = [p for p in new.phases() if p.model == 1]
        assert [p.horizon for p in wb] == plan.horizon, (k_new, [p.horizon for p in wb], plan.horizon)
        assert [tuple(p.contact) for p in wb] == plan.contact
        for i, p in enumerate(wb):
            assert tuple(list(p.td_foot)[:p.n_td]) == plan.touchdown_feet(i), (k_new, i)
            assert bool(p.single_shooting) == (not plan.has_ss[i]), (k_new, i)
        removed += len(plan.horizon) < n_before or (len(plan.horizon) == n_before and not plan.has_ss[-1])
        opened += not plan.has_ss[-1]
        g = mpc.shift_guess(prev, k_prev, new, k_new, old)
        for i in range(len(wb)):
            if not plan.has_ss[i]:
                continue                         # freshly opened phase: no shooting states, its Xbar is never read (mpc.py header)
            np.testing.assert_array_equal(g[i]["Xbar"], np.array(plan.X[i]), err_msg="X %d %d" % (k_new, i))
            np.testing.assert_array_equal(g[i]["Ubar"], np.array(plan.U[i]))
            np.testing.assert_array_equal(g[i]["K"], np.array(plan.K[i][:-1]))
        np.testing.assert_array_equal(g[len(wb)]["Xbar"], old[-1]["Xbar"])       # update_SRB_plan: nsteps = 0, the SRB arrays stay
        prev, k_prev = new, k_new
    assert opened >= 3 and removed >= 3
