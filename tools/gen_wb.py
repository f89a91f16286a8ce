#!/usr/bin/env python3
"""Generates cafe_mpc_b200/csrc/gen/wb_gen.h: straight-line device functions of the whole-body model, emitted
from the symbolic model in tools/wb_model.py (no Pinocchio, no CasADi at run time).

  wb_terms        (q, v)        -> nle[18], M[18x18 lower incl. diag], J[12x18], Jdv[12], pf[12], vf[12]    hip yaw 3.1415
                   replaces crba / nonLinearEffects / computeJointJacobians / getFrameJacobian / getFrameVelocity /
                   getFrameAcceleration at MHPC/MHPC-Trajopt/WBM.cpp:375-411
  wb_feet         (q, v)        -> pf[12], vf[12], J[12x18]                                                   hip yaw 3.1415
                   replaces the kinematics getters WBM.cpp:260-364
  wb_rnea_derivs_{trunk,leg0..3} (q, v, a) -> that inertia group's share of dtau_dq[18x18], dtau_dv[18x18]   hip yaw 3.1415
                   together they replace pinocchio::computeRNEADerivatives (WBM.cpp:474, :514)
  wb_grav_derivs  (q)           -> dg_dq[18x18]                                                               hip yaw 3.1415
                   replaces computeGeneralizedGravityDerivatives (WBM.cpp:520)
  wb_kin_partials_foot{0..3} (q, v, a, F) -> that foot's rows of dv_dq[12x18], da_dq[12x18], da_dv[12x18] and its share of
                   dJTF_dq[18x18]                                                                              hip yaw pi
                   replaces the CasADi functions footVelPartialDq / footAccPartialDq / footAccPartialDv /
                   footForcePartialDq (MCKinematicsDerivativs.cpp), which were generated with exactly pi
                   (SURVEY.md §9 Q16); verified against the compiled reference code to 1e-14 in tests/.
Foot rows are stacked FL,FR,HL,HR (3 rows each)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from casadi2cuda import HEADER  # noqa: E402

# the banner of casadi2cuda.py describes re-emitted reference code; these files come from the repository's own symbolic model
_CASADI_NOTE = '// Re-emission (CSE + constant folding + dead-code elimination, SSA form, functor\n// outputs) of CasADi expression graphs shipped with the reference; see the tool\n// for the exact source function of every routine below.\n'
_OWN_NOTE = "// Straight-line code (CSE + constant folding + dead-code elimination, SSA form, functor\n// outputs) emitted from this repository's OWN symbolic rigid-body model of the Mini Cheetah\n// (tools/wb_model.py, tools/symbolic.py); not derived from the reference's CasADi-generated code.\n"
from symbolic import Ctx, emit_function  # noqa: E402
from wb_model import WBModel, hardcoded_params, make_vars, params_from_urdf  # noqa: E402

URDF = "/root/reference/urdf/mini_cheetah_simple_correctedInertia.urdf"
HIP_YAW_URDF = 3.1415
SYNC_EVERY = 1024   # operations between two CAFE_GEN_SYNC markers in the routines that only k_lq calls


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "cafe_mpc_b200/csrc/gen/wb_gen.h"
    P = params_from_urdf(URDF) if os.path.exists(URDF) else hardcoded_params()
    pieces = []
    # ---- Pinocchio-side quantities (hip yaw as in the URDF)
    ctx = Ctx()
    m = WBModel(ctx, P, HIP_YAW_URDF)
    q, v, a = make_vars(ctx, 0, 18), make_vars(ctx, 1, 18), make_vars(ctx, 2, 18)
    zero = [ctx.const(0.0)] * 18
    nle = m.rnea(q, v, zero)
    M = m.mass_matrix(q)
    feet = m.feet(q, v, None)
    o_nle = [(i, nle[i]) for i in range(18)]
    o_M = [(r + 18 * c, M[r][c]) for c in range(18) for r in range(c, 18)]
    o_J = [(3 * f + r + 12 * c, feet[f]["J"][r][c]) for f in range(4) for c in range(18) for r in range(3)]
    o_g = [(3 * f + r, feet[f]["a"][r]) for f in range(4) for r in range(3)]
    o_p = [(3 * f + r, feet[f]["p"][r]) for f in range(4) for r in range(3)]
    o_v = [(3 * f + r, feet[f]["v"][r]) for f in range(4) for r in range(3)]
    pieces.append(emit_function(ctx, "wb_terms", 2, [o_nle, o_M, o_J, o_g, o_p, o_v]))   # monolithic version: host tests only
    # the same quantities split by inertia group (M and nle are linear in the inertias; foot f's kinematics involve the base and
    # leg f only): trunk -> its share of nle, M; leg f -> its share of nle, M and foot f's J, Jdot v, position, velocity. The
    # shares overlap in rows/columns 0..5 only (accumulated by the caller, wb_pieces.h).
    for name, only, f in [("trunk", {5}, None)] + [("leg%d" % f, {6 + 3 * f, 7 + 3 * f, 8 + 3 * f}, f) for f in range(4)]:
        nle_p = m.rnea(q, v, zero, only=only)
        cols = {}
        def mcol(c):
            if c not in cols:
                e = [ctx.const(1.0 if i == c else 0.0) for i in range(18)]
                cols[c] = m.rnea(q, zero, e, gravity=False, only=only)
            return cols[c]
        outs = [[(i, nle_p[i]) for i in range(18)], [(r + 18 * c, mcol(c)[r]) for c in range(18) for r in range(c, 18)]]
        if f is not None:
            outs += [[(3 * f + r + 12 * c, feet[f]["J"][r][c]) for c in range(18) for r in range(3)], [(3 * f + r, feet[f]["a"][r]) for r in range(3)],
                     [(3 * f + r, feet[f]["p"][r]) for r in range(3)], [(3 * f + r, feet[f]["v"][r]) for r in range(3)]]
        pieces.append(emit_function(ctx, "wb_terms_" + name, 2, outs))
    pieces.append(emit_function(ctx, "wb_feet", 2, [o_p, o_v, o_J]))
    # RNEA derivatives, split by inertia: tau = tau[trunk] + sum_f tau[leg f] (RNEA is linear in the inertias); tau[leg f] depends
    # on the base and on leg f only, so every piece is a short function with a small live set (the monolithic 34.6 k-op version
    # spilled ~260 KB per call). The pieces overlap only in rows 0..5 x columns 3..5 (summed by the caller, wb_pieces.h).
    for name, only in [("trunk", {5})] + [("leg%d" % f, {6 + 3 * f, 7 + 3 * f, 8 + 3 * f}) for f in range(4)]:
        tau = m.rnea(q, v, a, only=only)
        o_dq = [(r + 18 * c, tau[r].d(q[c])) for c in range(18) for r in range(18)]
        o_dv = [(r + 18 * c, tau[r].d(v[c])) for c in range(18) for r in range(18)]
        pieces.append(emit_function(ctx, "wb_rnea_derivs_" + name, 3, [o_dq, o_dv], sync_every=SYNC_EVERY))   # k_lq only: lock-step markers
    grav = m.rnea(q, zero, zero)
    o_gq = [(r + 18 * c, grav[r].d(q[c])) for c in range(18) for r in range(18)]
    pieces.append(emit_function(ctx, "wb_grav_derivs", 1, [o_gq]))
    # ---- CasADi-side kinematic partials (hip yaw exactly pi)
    ctx2 = Ctx()
    m2 = WBModel(ctx2, P, math.pi)
    q, v, a, F = make_vars(ctx2, 0, 18), make_vars(ctx2, 1, 18), make_vars(ctx2, 2, 18), make_vars(ctx2, 3, 12)
    feet = m2.feet(q, v, a)
    # one function per foot: its rows of dv/dq, da/dq, da/dv and its share d(J_f^T F_f)/dq of the contact-force term; the shares
    # overlap only in rows 3..5 x columns 3..5 (summed by the caller, wb_pieces.h)
    for f in range(4):
        fv, fa, J = feet[f]["v"], feet[f]["a"], feet[f]["J"]
        o_dv = [(3 * f + r + 12 * c, fv[r].d(q[c])) for c in range(18) for r in range(3)]
        o_daq = [(3 * f + r + 12 * c, fa[r].d(q[c])) for c in range(18) for r in range(3)]
        o_dav = [(3 * f + r + 12 * c, fa[r].d(v[c])) for c in range(18) for r in range(3)]
        jtf = [J[0][i] * F[3 * f] + J[1][i] * F[3 * f + 1] + J[2][i] * F[3 * f + 2] for i in range(18)]
        o_jtf = [(r + 18 * c, jtf[r].d(q[c])) for c in range(18) for r in range(18)]
        pieces.append(emit_function(ctx2, "wb_kin_partials_foot%d" % f, 4, [o_dv, o_daq, o_dav, o_jtf], sync_every=SYNC_EVERY))
        nz = [(i % 18, i // 18) for i, s_ in o_jtf if not s_.is_zero()]
        print("jtf foot", f, "rows", sorted(set(r for r, c in nz)), "cols", sorted(set(c for r, c in nz)))
    # dv_dq alone (swing-foot velocity costs, touchdown velocity penalty, impact derivatives)
    feet_v = m2.feet(q, v, None)
    o_dv2 = []
    for f in range(4):
        o_dv2 += [(3 * f + r + 12 * c, feet_v[f]["v"][r].d(q[c])) for c in range(18) for r in range(3)]
    pieces.append(emit_function(ctx2, "wb_footvel_partial", 2, [o_dv2]))
    with open(out, "w") as fh:
        fh.write(HEADER.replace("tools/casadi2cuda.py", "tools/gen_wb.py (symbolic whole-body model, tools/wb_model.py)").replace(_CASADI_NOTE, _OWN_NOTE))
        fh.write("#ifndef CAFE_GEN_SYNC\n#define CAFE_GEN_SYNC\n#endif\n")
        fh.write("namespace cafe_gen_wb {\n\n")
        for code, meta in pieces:
            fh.write("// %s: %d ops, output non-zeros %s\n" % (meta["name"], meta["ops"], meta["nnz"]))
            fh.write(code)
            fh.write("\n\n")
        fh.write("}  // namespace cafe_gen_wb\n")
    for code, meta in pieces:
        print(meta)


if __name__ == "__main__":
    main()
