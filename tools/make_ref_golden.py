#!/usr/bin/env python
"""Golden vectors from the REFERENCE's own HS-DDP solver (test infrastructure; needs /root/reference, so it runs in the build
container only - the fixture it writes travels).

oracle/_ref/ref_hkd is the reference's MultiPhaseDDP / SinglePhase / constraints / trajectory management / HKDProblem / HKD costs
and constraints / QuadReference / CasADi interface compiled unchanged from /root/reference (oracle/refbuild/Makefile) against the
Eigen / Boost / LCM stand-ins of oracle/refbuild/shim. This script runs it on the first N_PROB problems of the HKD trot workload
(problem 0 = the nominal problem of HKDMPCSolver::initialize, the others SplitMix64-perturbed, cafe_mpc_b200/workload.py) through the
initial solve (caps of HKDMPC/settings/ddp_setting.info) and N_UPD consecutive MPC updates (HKDProblem::update, caps 2 x 1), and
stores what the reference decided and produced in tests/golden/ref_hkd_trot.npz:

  per problem b and solve s (0 = initial, s >= 1 = after the s-th update), prefix p{b}_s{s}_:
    x0 [24], counters [3] = iter, ls_iter_total, reg_iter_total, final [4] = cost, feas, max_tconstr, max_pconstr,
    trace [iter, 12] in the oracle's CAFE_TRACE_W layout (oracle/oracle_api.h; taken in full precision by a decorator around every
      phase; column 4 (merit_rho) and 5 (reg_after) are derived from the observed columns with the solver's formulas),
    n_al = number of AL parameter updates, horizons [n_phases], contacts [n_phases, 4],
    per phase i: ph{i}_Xbar [(h+1), 24], ph{i}_Ubar [h, 24], ph{i}_dU [h, 24], ph{i}_Kv [h, 24] = K v for the fixed vector
      v = cos(1 + 0..23), ph0_K4 [min(h,4), 24, 24] (the first gains of the plan); for s >= 1 also the warm start the reference's
      update left behind: ph{i}_gXbar, ph{i}_gUbar, ph{i}_gKv.
  p0_s0_ph{i}_K, _Quu, _Qux, _G, _Qu in full for the nominal initial solve.
  inputs: body [N_PROB, 12], qJ [N_PROB, 12], nudge [N_PROB, N_UPD, 3].

Usage: python tools/make_ref_golden.py   (builds oracle/_ref/ref_hkd first)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
N_PROB, N_UPD = 4, 6
os.makedirs(os.path.join(REPO, "data/_run"), exist_ok=True)     # the working directory the reference resolves "../HKDMPC/settings/..." from
KV = np.cos(1.0 + np.arange(24))
ARRAYS = ("Xbar", "Ubar", "K", "dU", "G", "Qu", "Quu", "Qux", "Defect")


def parse(path):
    """-> list over problems of list over solves of dict(guess=[phase dicts], x0, counters, final, events, solution=[phase dicts])"""
    L = open(path).read().split("\n")
    pos = 0

    def phases(tag):
        nonlocal pos
        head = L[pos].split(); pos += 1
        assert head[0] == tag and head[1] == "n_phases", head
        out = []
        for i in range(int(head[2])):
            h = L[pos].split(); pos += 1
            assert h[0] == "phase" and int(h[1]) == i
            ph = dict(horizon=int(h[3]), contact=[int(v) for v in h[5:9]], start=float(h[10]), end=float(h[12]), n=int(h[14]), m=int(h[16]))
            for name in ARRAYS:
                t = L[pos].split(); pos += 1
                assert t[0] == name, (t[0], name)
                ph[name] = np.array(t[1:], dtype=np.float64)
            hz, n, m = ph["horizon"], ph["n"], ph["m"]
            ph["Xbar"] = ph["Xbar"].reshape(hz + 1, n); ph["G"] = ph["G"].reshape(hz + 1, n); ph["Defect"] = ph["Defect"].reshape(hz + 1, n)
            for name in ("Ubar", "dU", "Qu"):
                ph[name] = ph[name].reshape(hz, m)
            ph["K"] = ph["K"].reshape(hz, n, m).transpose(0, 2, 1)      # written column-major: [knot][column][row] -> [knot][m][n]
            ph["Qux"] = ph["Qux"].reshape(hz, n, m).transpose(0, 2, 1)
            ph["Quu"] = ph["Quu"].reshape(hz, m, m).transpose(0, 2, 1)
            out.append(ph)
        return out

    head = L[pos].split(); pos += 1
    n_prob, n_upd = int(head[1]), int(head[3])
    probs = []
    for b in range(n_prob):
        assert L[pos].split() == ["problem", str(b)]; pos += 1
        solves = []
        for s in range(n_upd + 1):
            if s > 0:
                assert L[pos].split() == ["update", str(s - 1)]; pos += 1
            rec = dict(guess=phases("guess"))
            t = L[pos].split(); pos += 1; assert t[0] == "x0"; rec["x0"] = np.array(t[1:], dtype=np.float64)
            t = L[pos].split(); pos += 1; assert t[0] == "counters"; rec["counters"] = np.array([int(t[2]), int(t[4]), int(t[6])])
            t = L[pos].split(); pos += 1; assert t[0] == "final"; rec["final"] = np.array([float(t[2]), float(t[4]), float(t[6]), float(t[8])])
            t = L[pos].split(); pos += 1; assert t[0] == "float_cost_buffer"
            t = L[pos].split(); pos += 1; assert t[0] == "events"
            ne = int(t[1])
            rec["events"] = [(int(a), int(p), float(x), float(y)) for a, p, x, y in (ln.split() for ln in L[pos:pos + ne])]
            pos += ne
            rec["solution"] = phases("solution")
            solves.append(rec)
        probs.append(solves)
    return probs


EV_ROLLOUT, EV_COST, EV_FEAS, EV_LQ, EV_BWD, EV_LIN, EV_ACCEPT, EV_AL, EV_BWD_DV = 1, 2, 3, 4, 5, 6, 7, 8, 12


def trace_of(events, n_phases, opt):
    """The 12-column per-iteration record (oracle/oracle_api.h CAFE_TRACE_W) from the decorator's event list."""
    # split at the LQ events of phase 0: an iteration = [cost + feas before it] LQ, sweeps, linear rollout, line-search trials
    lq = [i for i, e in enumerate(events) if e[0] == EV_LQ and e[1] == 0]
    rows = []
    for n, i0 in enumerate(lq):
        i1 = lq[n + 1] if n + 1 < len(lq) else len(events)
        # cost / feas evaluated just before this LQ: the last n_phases COST / FEAS events before i0
        pre = events[:i0]
        cost = sum(e[2] for e in [e for e in pre if e[0] == EV_COST][-n_phases:])           # phase order, like MultiPhaseDDP::compute_cost
        feas = np.sqrt(sum(e[2] for e in [e for e in pre if e[0] == EV_FEAS][-n_phases:]))
        body = events[i0:i1]
        # the next iteration's leading compute_cost / feasibility belong to it, not to this one: cut them off (they follow the last
        # ACCEPT or the last line-search trial)
        sweeps = [e for e in body if e[0] == EV_BWD and e[1] == n_phases - 1]
        ok_sweep = [e for e in body if e[0] == EV_BWD and e[1] == 0 and e[3] == 1.0]
        lin = [e for e in body if e[0] == EV_LIN]
        dV1 = 0.0; dV2 = 0.0
        for e in lin:            # phase order 0..n-1 (MultiPhaseDDP::linear_rollout)
            dV1 += e[2]; dV2 += e[3]
        if not lin:
            # MS = false: no linear rollout, the expected cost change is the successful sweep's, summed last phase first (MultiPhaseDDP.cpp:174-213)
            for e in [e for e in body if e[0] == EV_BWD_DV][-n_phases:]:
                dV1 += e[2]; dV2 += e[3]
        row = np.zeros(12)
        row[0], row[1], row[2], row[3] = cost, feas, dV1, dV2
        dV_abs = abs(dV1 + 0.5 * dV2)
        row[4] = dV_abs / ((1 - opt["merit_scale"]) * feas) + opt["merit_offset"] if feas > opt["dynamics_feas_thresh"] else 0.0
        reg = ok_sweep[-1][2] if ok_sweep else sweeps[-1][2]
        reg_after = reg / 20
        row[5] = 0.0 if reg_after < 1e-06 else reg_after
        row[6] = len(sweeps)
        trials = [j for j, e in enumerate(body) if e[0] == EV_ROLLOUT and e[1] == 0]
        row[7] = len(trials)
        accepted = any(e[0] == EV_ACCEPT for e in body)
        row[8] = 1.0 if accepted else 0.0
        if trials:
            j = trials[-1]
            nxt = body[j:]
            row[9] = body[j][2] if accepted else 0.0
            c_after = sum(e[2] for e in [e for e in nxt if e[0] == EV_COST][:n_phases])
            f_after = np.sqrt(sum(e[2] for e in [e for e in nxt if e[0] == EV_FEAS][:n_phases]))
            row[10] = c_after if accepted else cost
            row[11] = f_after
        else:
            row[10], row[11] = cost, feas
        rows.append(row)
    return np.array(rows).reshape(-1, 12)


def store(out, probs, optd, kv_of):
    for b, solves in enumerate(probs):
        for s, rec in enumerate(solves):
            pre = "p%d_s%d_" % (b, s)
            sol = rec["solution"]
            out[pre + "x0"] = rec["x0"]; out[pre + "counters"] = rec["counters"]; out[pre + "final"] = rec["final"]
            out[pre + "trace"] = trace_of(rec["events"], len(sol), optd)
            assert len(out[pre + "trace"]) == rec["counters"][0]
            assert int(out[pre + "trace"][:, 7].sum()) == rec["counters"][1] and int(out[pre + "trace"][:, 6].sum()) == rec["counters"][2]
            out[pre + "n_al"] = np.array(sum(1 for e in rec["events"] if e[0] == EV_AL and e[1] == 0))
            out[pre + "horizons"] = np.array([p["horizon"] for p in sol]); out[pre + "contacts"] = np.array([p["contact"] for p in sol])
            for i, p in enumerate(sol):
                q = pre + "ph%d_" % i
                kv = kv_of(p["n"])
                out[q + "Xbar"], out[q + "Ubar"], out[q + "dU"] = p["Xbar"], p["Ubar"], p["dU"]
                out[q + "Kv"] = p["K"] @ kv
                if i == 0:
                    out[q + "K4"] = p["K"][:4]
                if b == 0 and s == 0:
                    out[q + "K"], out[q + "Quu"], out[q + "Qux"], out[q + "G"], out[q + "Qu"] = p["K"], p["Quu"], p["Qux"], p["G"], p["Qu"]
                if s > 0:
                    g = rec["guess"][i]
                    out[q + "gXbar"], out[q + "gUbar"], out[q + "gKv"] = g["Xbar"], g["Ubar"], g["K"] @ kv


def main_mhpc():
    """tests/golden/ref_mhpc_trot.npz: the reference's MHPCProblem (whole-body 11 + 14 knots, SRB 10 knots at the start of the trot reference)
    through MHPCLocomotion's initial solve and N_UPD_MHPC MPC updates (run-time caps), oracle/_ref/ref_mhpc. Same keys as the HKD file; the
    inputs are x0 [N, 36] and nudge [N, N_UPD_MHPC, 36]; kv is the 36-vector, SRB phases use its first 12 entries."""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    N, U = 3, 8
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    prob = cm.MHPCProblem(csv)
    x0 = w.mhpc_batch(N)
    nudge = np.zeros((N, U, 36))
    for b in range(N):
        nudge[b, :] = 1e-3 * (x0[b] - x0[0])
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        with open(fin, "w") as f:
            f.write("%d %d\n" % (N, U))
            for b in range(N):
                f.write(" ".join(repr(float(v)) for v in x0[b]) + "\n")
                for u in range(U):
                    f.write(" ".join(repr(float(v)) for v in nudge[b, u]) + "\n")
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_mhpc"), csv, repr(float(prob.deck.contents.hip_yaw)), fin, fout],
                              cwd=os.path.join(REPO, "data/_run"), stdout=subprocess.DEVNULL)
        probs = parse(fout)
    kv36 = np.cos(1.0 + np.arange(36))
    out = dict(x0=x0, nudge=nudge, kv=kv36)
    store(out, probs, optd, lambda n: kv36[:n])
    dst = os.path.join(REPO, "tests/golden/ref_mhpc_trot.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB;", "iterations of the initial solves:", [int(out["p%d_s0_counters" % b][0]) for b in range(N)],
          "layouts:", sorted({tuple(out["p0_s%d_horizons" % s]) for s in range(U + 1)}))


def main():
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle/refbuild"), "-j8"], stdout=subprocess.DEVNULL)
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    prob = cm.HKDProblem(csv)
    body = np.tile(w.HKD_NOMINAL_BODY, (N_PROB, 1)); qJ = np.tile(w.HKD_NOMINAL_QJ, (N_PROB, 1))
    for b in range(1, N_PROB):
        for j in range(12):
            body[b, j] += w.HKD_BODY_SCALE[j] * (2 * w.uniform(b, j) - 1)
            qJ[b, j] += w.HKD_QJ_SCALE[j] * (2 * w.uniform(b, 12 + j) - 1)
    x0 = np.stack([prob.initial_state(body[b], qJ[b]) for b in range(N_PROB)])
    assert np.array_equal(x0, w.hkd_batch(prob, N_PROB))
    nudge = np.zeros((N_PROB, N_UPD, 3))
    for b in range(N_PROB):
        nudge[b, :] = 1e-3 * (x0[b, 3:6] - x0[0, 3:6])
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
    optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        with open(fin, "w") as f:
            f.write("%d %d\n" % (N_PROB, N_UPD))
            for b in range(N_PROB):
                f.write(" ".join(repr(float(v)) for v in np.concatenate([body[b], qJ[b]])) + "\n")
                for u in range(N_UPD):
                    f.write(" ".join(repr(float(v)) for v in nudge[b, u]) + "\n")
        # the reference reads "../HKDMPC/settings/*.info" relative to its working directory (HKDProblem.cpp:71, HKDMPC.cpp:23)
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_hkd"), csv, fin, fout], cwd=os.path.join(REPO, "data/_run"), stdout=subprocess.DEVNULL)
        probs = parse(fout)
    out = dict(body=body, qJ=qJ, nudge=nudge, kv=KV)
    store(out, probs, optd, lambda n: KV)
    dst = os.path.join(REPO, "tests/golden/ref_hkd_trot.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB;", "iterations of the initial solves:", [int(out["p%d_s0_counters" % b][0]) for b in range(N_PROB)])


def main_barrel():
    """tests/golden/ref_mhpc_barrel.npz: BASELINE config 4, the running barrel roll (Reference/Data/running_br, cost_weights_barrel.JSON,
    constraint_params_barrel.info through data/MHPC/settings/mhpc_config_barrel.info) at the start offsets 0 (stance -> diagonal pair -> flight)
    and 205 (mid-roll flight of 22 knots -> four-foot landing impact with four touchdown constraints -> stance): the reference starts a plan
    at the first record of its reference file, so the offset is produced by handing it the file without its first k0 records (18 lines each).
    Two problems per offset (nominal + perturbed, workload.barrel_batch), initial solve only; keys prefixed k{k0}_."""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    N = 2
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
    kv36 = np.cos(1.0 + np.arange(36))
    out = dict(kv=kv36)
    src = open(w.BARREL_CSV).read().split("\n")
    for k0 in (0, w.BARREL_K0_IMPACT):
        prob = cm.MHPCProblem(w.BARREL_CSV, mhpc_config=w.BARREL_CONFIG, k0=k0)
        x0 = w.barrel_batch(prob, N)
        with tempfile.TemporaryDirectory() as td:
            fin, fout, fcsv = os.path.join(td, "in.txt"), os.path.join(td, "out.txt"), os.path.join(td, "quad_reference.csv")
            open(fcsv, "w").write("\n".join(src[:2] + src[2 + 18 * k0:]))
            with open(fin, "w") as f:
                f.write("%d 0\n" % N)
                for b in range(N):
                    f.write(" ".join(repr(float(v)) for v in x0[b]) + "\n")
            subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_mhpc"), fcsv, repr(float(prob.deck.contents.hip_yaw)), fin, fout,
                                   "../MHPC/settings/mhpc_config_barrel.info"], cwd=os.path.join(REPO, "data/_run"), stdout=subprocess.DEVNULL)
            probs = parse(fout)
        sub = {}
        store(sub, probs, optd, lambda n: kv36[:n])
        if k0 != 0:      # the full gain / Q arrays once are enough
            for k in [k for k in sub if "_ph" in k and k.rsplit("_", 1)[1] in ("K", "Quu", "Qux", "G", "Qu")]:
                del sub[k]
        out["k%d_x0" % k0] = x0
        for k, v in sub.items():
            out["k%d_%s" % (k0, k)] = v
        print("k0", k0, "counters", [list(sub["p%d_s0_counters" % b]) for b in range(N)])
    dst = os.path.join(REPO, "tests/golden/ref_mhpc_barrel.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB")


def settings_tree(td, ms_false, overrides=None):
    """A copy of the settings directories the reference reads relative to its working directory, with the shooting switch (and other ddp_setting
    keys: overrides = {key: text}) overridden; returns the cwd."""
    import re
    import shutil
    for sub in ("HKDMPC/settings", "MHPC/settings"):
        shutil.copytree(os.path.join(REPO, "data", sub), os.path.join(td, sub))
        f = os.path.join(td, sub, "ddp_setting.info")
        txt = open(f).read()
        if ms_false:
            txt, n = re.subn(r"(\bMS\s+)true", r"\1false", txt)
            assert n == 1
        for key, val in (overrides or {}).items():
            txt, n = re.subn(r"(\b%s\s+)\S+" % key, r"\g<1>%s" % val, txt)
            assert n == 1, key
        open(f, "w").write(txt)
    run = os.path.join(td, "_run")
    os.makedirs(run)
    return run


def main_single_shooting():
    """tests/golden/ref_single_shooting.npz: HSDDP_OPTION::MS = false (whole-problem single shooting: MultiPhaseDDP.cpp:65-68, :330-333,
    SinglePhase.cpp:187-221, :383-387) run by the reference itself - ddp_setting.info with `MS false`, everything else as shipped - on two HKD
    and two MHPC trot problems (ref_hkd / ref_mhpc, initial solve). Keys prefixed hkd_ / mhpc_."""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    csv = os.path.join(REPO, "data/Reference/Data/trot/heuristic/quad_reference.csv")
    kv36 = np.cos(1.0 + np.arange(36))
    out = dict(kv=kv36)
    N = 2
    with tempfile.TemporaryDirectory() as td:
        run = settings_tree(td, True)
        # HKD
        opt = cm.load_hsddp_setting(os.path.join(REPO, "data/HKDMPC/settings/ddp_setting.info"))
        optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
        fin, fout = os.path.join(td, "in.txt"), os.path.join(td, "out.txt")
        body = np.tile(w.HKD_NOMINAL_BODY, (N, 1)); qJ = np.tile(w.HKD_NOMINAL_QJ, (N, 1))
        for b in range(1, N):
            for j in range(12):
                body[b, j] += w.HKD_BODY_SCALE[j] * (2 * w.uniform(b, j) - 1)
                qJ[b, j] += w.HKD_QJ_SCALE[j] * (2 * w.uniform(b, 12 + j) - 1)
        with open(fin, "w") as f:
            f.write("%d 0\n" % N)
            for b in range(N):
                f.write(" ".join(repr(float(v)) for v in np.concatenate([body[b], qJ[b]])) + "\n")
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_hkd"), csv, fin, fout], cwd=run, stdout=subprocess.DEVNULL)
        sub = {}
        store(sub, parse(fout), optd, lambda n: kv36[:n])
        for k, v in sub.items():
            if not ("_ph" in k and k.rsplit("_", 1)[1] in ("K", "Quu", "Qux", "G", "Qu")):
                out["hkd_" + k] = v
        print("hkd MS=false counters", [list(sub["p%d_s0_counters" % b]) for b in range(N)])
        # MHPC
        prob = cm.MHPCProblem(csv)
        x0 = w.mhpc_batch(N)
        with open(fin, "w") as f:
            f.write("%d 0\n" % N)
            for b in range(N):
                f.write(" ".join(repr(float(v)) for v in x0[b]) + "\n")
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_mhpc"), csv, repr(float(prob.deck.contents.hip_yaw)), fin, fout], cwd=run, stdout=subprocess.DEVNULL)
        sub = {}
        store(sub, parse(fout), optd, lambda n: kv36[:n])
        for k, v in sub.items():
            if not ("_ph" in k and k.rsplit("_", 1)[1] in ("K", "Quu", "Qux", "G", "Qu")):
                out["mhpc_" + k] = v
        out["mhpc_x0"] = x0
        print("mhpc MS=false counters", [list(sub["p%d_s0_counters" % b]) for b in range(N)])
    dst = os.path.join(REPO, "tests/golden/ref_single_shooting.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB")


def main_reb():
    """tests/golden/ref_reb_update.npz: PathConstraintBase::update_params (ConstraintsBase.h:79-85, :194-209) with update factors other than the
    shipped 1 / 1, run by the reference itself: ddp_setting.info with update_relax = 0.5, update_ReB = 2 on the running barrel roll at start
    offset 205 (the landing leaves friction / torque barriers violated at the end of outer iterations, so the relaxation and the weight of
    those elements are updated seven times), two problems, ref_mhpc."""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    N, k0 = 2, w.BARREL_K0_IMPACT
    opt = cm.load_hsddp_setting(os.path.join(REPO, "data/MHPC/settings/ddp_setting.info"))
    optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
    kv36 = np.cos(1.0 + np.arange(36))
    out = dict(kv=kv36, update_relax=np.array(0.5), update_ReB=np.array(2.0))
    src = open(w.BARREL_CSV).read().split("\n")
    prob = cm.MHPCProblem(w.BARREL_CSV, mhpc_config=w.BARREL_CONFIG, k0=k0)
    x0 = w.barrel_batch(prob, N)
    with tempfile.TemporaryDirectory() as td:
        run = settings_tree(td, False, {"update_relax": "0.5", "update_ReB": "2"})
        fin, fout, fcsv = os.path.join(td, "in.txt"), os.path.join(td, "out.txt"), os.path.join(td, "quad_reference.csv")
        open(fcsv, "w").write("\n".join(src[:2] + src[2 + 18 * k0:]))
        with open(fin, "w") as f:
            f.write("%d 0\n" % N)
            for b in range(N):
                f.write(" ".join(repr(float(v)) for v in x0[b]) + "\n")
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_mhpc"), fcsv, repr(float(prob.deck.contents.hip_yaw)), fin, fout,
                               "../MHPC/settings/mhpc_config_barrel.info"], cwd=run, stdout=subprocess.DEVNULL)
        probs = parse(fout)
        # the same settings through four MPC updates (problem 0): the relaxed-barrier parameters travel with the knots - PathConstraintBase::pop_front /
        # push_back (ConstraintsBase.h:296-306: a knot appended at the tail copies the LAST knot's parameters), reset_params is empty
        with open(fin, "w") as f:
            f.write("1 4\n" + " ".join(repr(float(v)) for v in x0[0]) + "\n")
            for u in range(4):
                f.write(" ".join(["0"] * 36) + "\n")
        fout2 = os.path.join(td, "out2.txt")
        subprocess.check_call([os.path.join(REPO, "oracle/_ref/ref_mhpc"), fcsv, repr(float(prob.deck.contents.hip_yaw)), fin, fout2,
                               "../MHPC/settings/mhpc_config_barrel.info"], cwd=run, stdout=subprocess.DEVNULL)
        chain = parse(fout2)
    sub = {}
    store(sub, chain, optd, lambda n: kv36[:n])
    for k, v in sub.items():
        if not ("_ph" in k and k.rsplit("_", 1)[1] in ("K", "Quu", "Qux", "G", "Qu")):
            out["chain_" + k] = v
    print("reb chain counters", [list(sub["p0_s%d_counters" % s_]) for s_ in range(5)], [list(sub["p0_s%d_horizons" % s_]) for s_ in range(5)])
    sub = {}
    store(sub, probs, optd, lambda n: kv36[:n])
    for k, v in sub.items():
        if not ("_ph" in k and k.rsplit("_", 1)[1] in ("K", "Quu", "Qux", "G", "Qu")):
            out[k] = v
    out["x0"] = x0
    print("reb counters", [list(sub["p%d_s0_counters" % b]) for b in range(N)], "n_reb_updates", [sum(1 for e in probs[b][0]["events"] if e[0] == 9 and e[1] == 0) for b in range(N)])
    dst = os.path.join(REPO, "tests/golden/ref_reb_update.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB")


def main_programs():
    """tests/golden/ref_programs.npz: the reference's stand-alone programs run UNCHANGED, main() included (oracle/_ref/ref_loco = Loco_TO.cpp,
    oracle/_ref/ref_barrel_to = BarrelRollTO.cpp; the solve is wrapped at link time, oracle/refbuild/ref_program_driver.cpp). Their initial
    states are hard-coded in main(): one problem each. Keys prefixed loco_ / barrel_to_, then as in the other files (p0_s0_...)."""
    import cafe_mpc_b200 as cm
    from cafe_mpc_b200 import workload as w
    kv36 = np.cos(1.0 + np.arange(36))
    out = dict(kv=kv36)
    for name, exe, setting in (("loco", "ref_loco", w.LOCO_DDP_SETTING), ("barrel_to", "ref_barrel_to", w.BARREL_TO_DDP_SETTING)):
        opt = cm.load_hsddp_setting(setting)
        optd = dict(merit_scale=opt.merit_scale, merit_offset=opt.merit_offset, dynamics_feas_thresh=opt.dynamics_feas_thresh)
        prob = cm.LocoProblem() if name == "loco" else cm.BarrelRollProblem()
        with tempfile.TemporaryDirectory() as td:
            fout = os.path.join(td, "out.txt")
            env = dict(os.environ, REF_OUT=fout, REF_HIP_YAW=repr(float(prob.deck.contents.hip_yaw)))
            subprocess.check_call([os.path.join(REPO, "oracle/_ref", exe)], cwd=os.path.join(REPO, "data/_run"), stdout=subprocess.DEVNULL, env=env)
            text = open(fout).read()
            cut = text.index("published ")
            open(fout, "w").write(text[:cut])
            probs = parse(fout)
            pub = np.array([[float(v) for v in ln.split()[1:]] for ln in text[cut:].split("\n") if ln.startswith("wbtraj")])
        sub = {}
        store(sub, probs, optd, lambda n: kv36[:n])
        for k in [k for k in sub if "_ph" in k and k.rsplit("_", 1)[1] in ("Quu", "Qux", "G", "Qu") or k.endswith("_K")]:
            del sub[k]
        for k, v in sub.items():
            out["%s_%s" % (name, k)] = v
        # what the program itself published for the visualiser (wbTraj_lcmt, doubles): Xbar[k], Ubar[k] of every knot but the terminal ones
        out[name + "_published"] = pub
        print(name, "counters", list(sub["p0_s0_counters"]), "final", sub["p0_s0_final"], "published", pub.shape)
    dst = os.path.join(REPO, "tests/golden/ref_programs.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst) // 1024, "KB")


if __name__ == "__main__":
    main()
    main_mhpc()
    main_barrel()
    main_programs()
    main_single_shooting()
    main_reb()
