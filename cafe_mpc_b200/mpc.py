"""Receding-horizon re-solves: build the warm-start guess for the deck at a later start offset from the previous solution.

Restates what MHPCProblem::update / update_WB_plan / update_SRB_plan do to the trajectories (MHPC/MHPC-Trajopt/MHPCProblem.cpp:252-397,
HSDDPSolver/source/TrajectoryManagement.cpp:130-228, SinglePhase.cpp:513-528) as a pure function of the previous solution:

  * the first `shift` whole-body knots are popped (a phase that shrinks to a point disappears with its terminal state),
  * the last whole-body phase grows by push_back_state(X.back()): new states are copies of the last state, new controls and gains zero,
  * the SRB phase is NOT shifted while dt_mpc < dt_srb (update_SRB_plan: nsteps = floor(dt_mpc / dt_srb) = 0) - its arrays are kept,
  * everything else (dU, value function, Q terms) does not matter for the re-solve: MultiPhaseDDP::solve starts with
    hybrid_rollout(eps = 0), U = Ubar + K (X - Xbar).
The same function serves the HKD application (HKDProblem::update, HKDMPC/HKD-TrajOpt/HKDProblem.cpp:117-222: identical front / back
bookkeeping on its 24-state phases, no reduced-order tail, plus the quirk Ubar[0] = 0 of the front trajectory, :220).
Differences to the reference, by construction: the last state copied is Xbar (the reference copies X, equal unless the final line search
ended on a rejected trial); a phase opened at the tail by a contact change is given the new deck's reference states where the reference
constructs a fresh zero trajectory - immaterial once the deck is marked as an MPC update (MHPCProblem(..., mpc_update_nsteps=2),
HKDProblem(..., mpc_update=True)): that phase then has no shooting states (MHPCProblem.cpp:366-369, HKDProblem.cpp:213-217), is integrated
from the previous phase's hand-over state with zero gains, and its Xbar is never read."""
import numpy as np

from ._ctypes_defs import MODEL_DIMS

WB = 1
HKD = 0


def _wb_ranges(problem, k0):
    """[(phase index, absolute first knot, absolute last state index)] of the leading full-order phases (whole-body for MHPC decks,
    hybrid kinodynamic for HKD decks)."""
    out, s = [], k0
    lead = problem.phases()[0].model
    for i, ph in enumerate(problem.phases()):
        if ph.model != lead:
            break
        out.append((i, s, s + ph.horizon, tuple(ph.contact)))
        s += ph.horizon
    return out


def unpack_batch(problem, sols):
    """[B, solution_size] -> per phase {Xbar [B,h+1,n], Ubar [B,h,m], K [B,h,m,n]} (views where possible)."""
    sols = np.asarray(sols)
    B = sols.shape[0]
    out, off = [], 0
    for ph in problem.phases():
        n, m, p = MODEL_DIMS[ph.model]
        h = ph.horizon
        r = {}
        r["Xbar"] = sols[:, off:off + (h + 1) * n].reshape(B, h + 1, n); off += (h + 1) * n
        r["Ubar"] = sols[:, off:off + h * m].reshape(B, h, m); off += h * m
        off += h * p + h * m
        r["K"] = sols[:, off:off + h * m * n].reshape(B, h, n, m).transpose(0, 1, 3, 2); off += h * m * n   # stored column-major per knot
        off += h * m + h * m * m + h * m * n + (h + 1) * n
        out.append(r)
    return out


def pack_solution(problem, phases):
    """Inverse of api.unpack_solution / unpack_batch for the arrays a guess needs (Xbar, Ubar, K); the rest of the record is zero.
    Arrays may carry a leading batch dimension."""
    batched = np.asarray(phases[0]["Xbar"]).ndim == 3
    chunks = []
    for ph, r in zip(problem.phases(), phases):
        n, m, p = MODEL_DIMS[ph.model]
        h = ph.horizon
        X, U, K = (np.asarray(r[k]) if batched else np.asarray(r[k])[None] for k in ("Xbar", "Ubar", "K"))
        B = X.shape[0]
        z = lambda c: np.zeros((B, c))
        chunks += [X.reshape(B, -1), U.reshape(B, -1), z(h * p), z(h * m), K.transpose(0, 1, 3, 2).reshape(B, -1), z(h * m), z(h * m * m), z(h * m * n),
                   z((h + 1) * n)]
    out = np.concatenate(chunks, axis=1)
    return out if batched else out[0]


def shift_guess(old_problem, old_k0, new_problem, new_k0, old_phases):
    """old_phases: api.unpack_solution(...) of one problem or unpack_batch(...) of a batch (deck old_problem, start offset old_k0).
    Returns the per-phase dicts {Xbar, Ubar, K} of the guess for new_problem (start offset new_k0 >= old_k0, same reference file)."""
    old_r, new_r = _wb_ranges(old_problem, old_k0), _wb_ranges(new_problem, new_k0)
    ref = new_problem.reference_records()
    d = new_problem.deck.contents
    lead = np.asarray(old_phases[0]["Xbar"]).shape[:-2]          # () or (B,)
    last_state = np.asarray(old_phases[old_r[-1][0]]["Xbar"])[..., -1, :]
    old_end = old_r[-1][2]
    out = []
    lead_model = new_problem.phases()[0].model
    for i, ph in enumerate(new_problem.phases()):
        n, m, p = MODEL_DIMS[ph.model]
        h = ph.horizon
        if ph.model != lead_model:
            # trailing reduced-order phase: kept as it is when the horizons agree, else cold start from the reference
            j = len(old_r) + (i - len(new_r))
            if j < len(old_phases) and np.asarray(old_phases[j]["Ubar"]).shape[-2] == h:
                out.append({k: np.array(old_phases[j][k]) for k in ("Xbar", "Ubar", "K")})
            else:
                X = np.broadcast_to(ref[ph.knot_offset:ph.knot_offset + h + 1, :n], lead + (h + 1, n)).copy()
                out.append({"Xbar": X, "Ubar": np.zeros(lead + (h, m)), "K": np.zeros(lead + (h, m, n))})
            continue
        _, s, e, contact = new_r[i]
        X, U, K = np.zeros(lead + (h + 1, n)), np.zeros(lead + (h, m)), np.zeros(lead + (h, m, n))
        src = [r for r in old_r if r[3] == contact and r[1] <= e and r[2] >= s]   # the old phase this one continues (same stance, overlapping)
        continues_last = bool(src) and src[0][0] == old_r[-1][0]
        for k in range(h + 1):
            a = s + k
            if src and src[0][1] <= a <= src[0][2]:
                X[..., k, :] = np.asarray(old_phases[src[0][0]]["Xbar"])[..., a - src[0][1], :]
            elif continues_last and a > old_end:
                X[..., k, :] = last_state                              # push_back_state(X.back())
            else:
                X[..., k, :] = ref[ph.knot_offset + k, :n]              # a phase the old plan did not have yet
            if k < h and src and src[0][1] <= a < src[0][2]:
                U[..., k, :] = np.asarray(old_phases[src[0][0]]["Ubar"])[..., a - src[0][1], :]
                K[..., k, :, :] = np.asarray(old_phases[src[0][0]]["K"])[..., a - src[0][1], :, :]
        out.append({"Xbar": X, "Ubar": U, "K": K})
    if lead_model == HKD and out[0]["Ubar"].shape[-2] > 0:
        out[0]["Ubar"][..., 0, :] = 0.0                                 # HKDProblem.cpp:220
    assert d.n_phases == len(out)
    return out


def state_at(problem, phases, knots_ahead):
    """Planned whole-body state `knots_ahead` knots after the start of the plan (crossing phase boundaries: the post-reset state)."""
    k = knots_ahead
    for ph, r in zip(problem.phases(), phases):
        if k < ph.horizon or ph.model not in (WB, HKD):
            return np.array(np.asarray(r["Xbar"])[..., k, :])
        k -= ph.horizon
    raise ValueError("beyond the plan")


def shifted_guess_batch(old_problem, old_k0, new_problem, new_k0, old_solutions):
    """Packed guesses [B, solution_size(new deck)] from packed previous solutions [B, solution_size(old deck)]."""
    return pack_solution(new_problem, shift_guess(old_problem, old_k0, new_problem, new_k0, unpack_batch(old_problem, old_solutions)))


def initial_al(problem, B=None):
    """[n_phases, 4, 2] (or [B, n_phases, 4, 2]) = (sigma, lambda) every touchdown-constraint element starts from (the deck's TD_AL values)."""
    ph = problem.phases()
    al = np.zeros((len(ph), 4, 2))
    for i, p in enumerate(ph):
        al[i, :p.n_td, 0] = p.al_td.sigma
        al[i, :p.n_td, 1] = p.al_td.lambda_
    return al if B is None else np.broadcast_to(al, (B,) + al.shape).copy()


def shift_al(old_problem, old_k0, new_problem, new_k0, old_al):
    """The augmented-Lagrangian parameters the re-solve after an MPC update starts from: old_al [..., n_old_phases, 4, 2] as the previous
    solve left them -> [..., n_new_phases, 4, 2]. The reference keeps every phase's TouchDownConstraint object (and its sigma / lambda) for
    as long as the phase lives - reset_params(), called by every update, is an empty function (ConstraintsBase.h:367-374, HKDProblem.cpp:208,
    MHPCProblem.cpp:363) - so a phase that continues an old phase with the same touchdown feet inherits its values; a constraint that did
    not exist before (a phase the old plan did not have, or a tail phase that only now got its touchdown constraint) starts from the deck's."""
    old_al = np.asarray(old_al)
    lead = old_al.shape[:-3]
    new = np.broadcast_to(initial_al(new_problem), lead + (len(new_problem.phases()), 4, 2)).copy()
    old_r, new_r = _wb_ranges(old_problem, old_k0), _wb_ranges(new_problem, new_k0)
    oph, nph = old_problem.phases(), new_problem.phases()
    for i, s, e, contact in new_r:
        src = [r for r in old_r if r[3] == contact and r[1] <= e and r[2] >= s]
        if not src:
            continue
        j = src[0][0]
        if oph[j].n_td > 0 and oph[j].n_td == nph[i].n_td and list(oph[j].td_foot)[:oph[j].n_td] == list(nph[i].td_foot)[:nph[i].n_td]:
            new[..., i, :, :] = old_al[..., j, :, :]
    return new


def shift_reb(old_problem, old_k0, new_problem, new_k0, old_reb, init_reb):
    """The relaxed-barrier parameters the re-solve after an MPC update starts from. old_reb / init_reb: lists per phase of [h, ne, 2] = (delta, eps)
    per running knot and element (tests/oracle_bindings.py) - what the previous solve left behind, and what the new deck would start from.
    The reference keeps them with the knots of a phase (PathConstraintBase::pop_front / push_back, ConstraintsBase.h:296-306; reset_params, called
    by every update, is empty, :191-193): popped knots take theirs along, a knot appended at the tail COPIES THE LAST KNOT's values, a phase the
    old plan did not have starts from the deck's. Under the shipped settings (update_relax = update_ReB = 1) nothing ever changes them."""
    old_r, new_r = _wb_ranges(old_problem, old_k0), _wb_ranges(new_problem, new_k0)
    out = [np.array(r) for r in init_reb]
    lead = new_problem.phases()[0].model
    n_lead_old, n_lead_new = len(old_r), len(new_r)
    for i, ph in enumerate(new_problem.phases()):
        if ph.model != lead:
            j = n_lead_old + (i - n_lead_new)       # the trailing reduced-order phase keeps its data while its horizon is unchanged
            if 0 <= j < len(old_reb) and old_reb[j].shape == out[i].shape:
                out[i] = np.array(old_reb[j])
            continue
        _, s, e, contact = new_r[i]
        src = [r for r in old_r if r[3] == contact and r[1] <= e and r[2] >= s]
        if not src or old_reb[src[0][0]].shape[1:] != out[i].shape[1:]:
            continue
        j, js, je = src[0][0], src[0][1], src[0][2]
        for k in range(ph.horizon):
            a = s + k
            if js <= a < je:
                out[i][k] = old_reb[j][a - js]
            elif a >= je and je - js > 0:
                out[i][k] = old_reb[j][je - js - 1]
    return out
