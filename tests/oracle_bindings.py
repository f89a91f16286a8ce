"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/). Only tests/, smoke() and
bench.py's CPU-baseline legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from cafe_mpc_b200._ctypes_defs import CAFE_TRACE_W, Deck, Info, Options

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "build", "libcafe_oracle.so")


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-j8", "all"], check=True, capture_output=True)
    return ORACLE_LIB


_lib = None


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB) or os.path.exists("/root/reference"):
            build_oracle()
        _lib = C.CDLL(ORACLE_LIB)
        _lib.cafe_oracle_solution_size.restype = C.c_long
        _lib.cafe_oracle_solution_size.argtypes = [C.POINTER(Deck)]
        _lib.cafe_oracle_solve.argtypes = [C.POINTER(Deck), C.POINTER(Options), C.c_void_p, C.POINTER(Info), C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def oracle_reb_layout(deck):
    """[(h, ne)] per phase: knots with running constraints and relaxed-barrier elements per knot"""
    lib = oracle()
    n_ph = deck.contents.n_phases
    ne = (C.c_int * n_ph)()
    assert lib.cafe_oracle_reb_ne(deck, ne) == 0
    return [(deck.contents.phase[i].horizon, ne[i]) for i in range(n_ph)]


def oracle_reb_init(deck):
    """list per phase of [h, ne, 2] = (delta, eps) a fresh deck starts from"""
    lib = oracle()
    lay = oracle_reb_layout(deck)
    flat = np.zeros(sum(h * ne * 2 for h, ne in lay))
    assert lib.cafe_oracle_reb_init(deck, flat.ctypes.data_as(C.c_void_p)) == 0
    return reb_split(flat, lay)


def reb_split(flat, lay):
    out, o = [], 0
    for h, ne in lay:
        out.append(flat[o:o + h * ne * 2].reshape(h, ne, 2).copy()); o += h * ne * 2
    return out


def oracle_solve(deck, opt, x0, cap=256, guess=None, al=None, reb=None):
    """Returns (info dict, hist [n_hist,4], trace [iter,12], packed solution). guess: packed solution whose Xbar/Ubar/K start the solve.
    al: [n_phases, 4, 2] (sigma, lambda) the touchdown constraints start from (the MPC loop's carry-over); then a fifth value is returned:
    the parameters the solve left behind. reb: list per phase of [h, ne, 2] (delta, eps) the relaxed barriers start from; then a sixth value is
    returned: what the solve left behind (same layout)."""
    lib = oracle()
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    info = Info()
    hist = np.zeros((cap, 4))
    trace = np.zeros((cap, CAFE_TRACE_W))
    sol = np.zeros(lib.cafe_oracle_solution_size(deck))
    g = None
    if guess is not None:
        g = np.ascontiguousarray(guess, dtype=np.float64)
        assert g.size == sol.size
    n_ph = deck.contents.n_phases
    a_in = None
    if al is not None:
        a_in = np.ascontiguousarray(al, dtype=np.float64)
        assert a_in.size == n_ph * 8
    a_out = np.zeros((n_ph, 4, 2))
    r_in = r_out = None
    if reb is not None:
        lay = oracle_reb_layout(deck)
        assert [r.shape for r in reb] == [(h, ne, 2) for h, ne in lay], ([r.shape for r in reb], lay)
        r_in = np.ascontiguousarray(np.concatenate([r.ravel() for r in reb]) if reb else np.zeros(0))
        r_out = np.zeros_like(r_in)
    rc = lib.cafe_oracle_solve_carry(deck, C.byref(opt), x0.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p) if g is not None else None,
                                     a_in.ctypes.data_as(C.c_void_p) if a_in is not None else None, a_out.ctypes.data_as(C.c_void_p),
                                     r_in.ctypes.data_as(C.c_void_p) if r_in is not None else None, r_out.ctypes.data_as(C.c_void_p) if r_out is not None else None,
                                     C.byref(info), hist.ctypes.data_as(C.c_void_p), cap, trace.ctypes.data_as(C.c_void_p), cap,
                                     sol.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("oracle solve failed")
    d = info.as_dict()
    if reb is not None:
        return d, hist[:d["n_hist"]], trace[:d["iter"]], sol, a_out, reb_split(r_out, lay)
    if al is not None:
        return d, hist[:d["n_hist"]], trace[:d["iter"]], sol, a_out
    return d, hist[:d["n_hist"]], trace[:d["iter"]], sol


def deck_with_references(deck, records):
    """A copy of the deck whose reference records are `records` ([n_records, 120]); keep the returned tuple alive while in use."""
    from cafe_mpc_b200._ctypes_defs import Deck
    d2 = Deck.from_buffer_copy(deck.contents)
    rec = np.ascontiguousarray(records, dtype=np.float64)
    d2.ref = rec.ctypes.data_as(C.POINTER(C.c_double))
    return C.pointer(d2), (d2, rec)


def oracle_get(name, phase):
    lib = oracle()
    lib.cafe_oracle_get.restype = C.c_long
    buf = np.zeros(64 * 36 * 36 + 64)
    n = lib.cafe_oracle_get(name.encode(), phase, buf.ctypes.data_as(C.c_void_p))
    if n < 0:
        raise KeyError(name)
    return buf[:n].copy()


def oracle_dynamics(deck, phase, k, x, u, partials=False):
    from cafe_mpc_b200._ctypes_defs import MODEL_DIMS
    lib = oracle()
    n, m, p = MODEL_DIMS[deck.contents.phase[phase].model]
    x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
    xn, y = np.zeros(n), np.zeros(max(p, 1))
    A, B, Cm, D = np.zeros(n * n), np.zeros(n * m), np.zeros(max(p * n, 1)), np.zeros(max(p * m, 1))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    null = C.c_void_p()
    rc = lib.cafe_oracle_dynamics(deck, phase, k, vp(x), vp(u), vp(xn), vp(y), vp(A) if partials else null, vp(B) if partials else null,
                                  vp(Cm) if partials else null, vp(D) if partials else null)
    assert rc == 0
    if not partials:
        return xn, y[:p]
    return xn, y[:p], A.reshape(n, n, order="F"), B.reshape(n, m, order="F"), Cm[:p * n].reshape(p, n, order="F"), D[:p * m].reshape(p, m, order="F")


def oracle_resetmap(deck, phase, x, jac=False):
    from cafe_mpc_b200._ctypes_defs import MODEL_DIMS
    lib = oracle()
    n = MODEL_DIMS[deck.contents.phase[phase].model][0]
    x = np.ascontiguousarray(x, dtype=np.float64)
    xn = np.zeros(36); Px = np.zeros(36 * 36)
    nn = lib.cafe_oracle_resetmap(deck, phase, x.ctypes.data_as(C.c_void_p), xn.ctypes.data_as(C.c_void_p), Px.ctypes.data_as(C.c_void_p) if jac else C.c_void_p())
    assert nn > 0
    return (xn[:nn], Px[:nn * n].reshape(nn, n, order="F")) if jac else xn[:nn]


def oracle_wb_dynamics(q, v, u, contact, hip_yaw, BG_alpha=10.0):
    lib = oracle()
    lib.cafe_oracle_wb_dynamics.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    q = np.ascontiguousarray(q, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
    qdd, grf = np.zeros(18), np.zeros(12)
    c = (C.c_int * 4)(*contact)
    assert lib.cafe_oracle_wb_dynamics(hip_yaw, BG_alpha, q.ctypes.data, v.ctypes.data, u.ctypes.data, c, qdd.ctypes.data, grf.ctypes.data) == 0
    return qdd, grf


def casadi_eval(name, ins, out_shapes):
    """Evaluate a reference CasADi function (oracle/_ref) on dense inputs; returns dense column-major outputs."""
    lib = oracle()
    ins = [np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=np.float64))) for a in ins]
    outs = [np.zeros(int(np.prod(s))) for s in out_shapes]
    pin = (C.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * len(outs))(*[a.ctypes.data for a in outs])
    rc = lib.cafe_oracle_casadi_eval(name.encode(), pin, pout)
    if rc != 0:
        raise KeyError(name)
    return [o.reshape(s, order="F") for o, s in zip(outs, out_shapes)]
