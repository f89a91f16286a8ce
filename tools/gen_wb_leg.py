#!/usr/bin/env python3
"""Generates cafe_mpc_b200/csrc/gen/wb_leg_gen.h: LEG-GENERIC straight-line device functions of the whole-body model.

The four legs of the Mini Cheetah differ only in nine mirrored constants (hip offsets, y of two centres of mass, the xy / yz
products of inertia). The per-leg pieces of gen/wb_gen.h are therefore ONE routine each here, with those constants as an input
array: a kernel runs one thread per (problem, knot, leg), all lanes of a warp in the same routine with the warp's leg constants
coming from constant memory, a quarter of the code (instruction-cache footprint) and four times the threads of the per-knot form.

  wbl_terms_leg   (ql[9], vl[9], P[9])         -> this leg's share of nle, M (local 9 x 9 lower), its foot's J (3 x 9), Jdot v,
                                                  foot position, foot velocity                                hip yaw 3.1415
  wbl_rnea_{q3,q4,q5,q678,v345,v678} (ql, vl, al, P) -> columns (= directions) of this leg's share of dtau/dq, dtau/dv   hip yaw 3.1415
  wbl_kin_{q345,q678,v345678} (ql, vl, al, F[3], P) -> columns of this foot's dv/dq, da/dq, d(J^T F)/dq | da/dv          hip yaw pi
The derivative routines are generated in FORWARD MODE (symbolic.Dual): value and tangents interleaved, a few directions per routine, so
that a routine's live set fits the register file (differentiating the finished expression graph column by column keeps every
intermediate alive across all columns: 13 KB / 26 KB of spill stores / loads per call of the 8.7 k-op version).
Local coordinates: 0..5 = floating base (x y z yaw pitch roll), 6..8 = the leg's abduction, hip, knee.
Outputs are COMPACT: output k of a routine is its k-th structural non-zero; the tables at the end of the header map compact slots
to (array, local row, local column). The trunk shares come from wb_terms_trunk / wb_rnea_derivs_trunk of gen/wb_gen.h.
Reference call sites replaced: see tools/gen_wb.py (WBM.cpp:375-411, :474, :514; MCKinematicsDerivativs.cpp)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from casadi2cuda import HEADER  # noqa: E402

# the banner of casadi2cuda.py describes re-emitted reference code; these files come from the repository's own symbolic model
_CASADI_NOTE = '// Re-emission (CSE + constant folding + dead-code elimination, SSA form, functor\n// outputs) of CasADi expression graphs shipped with the reference; see the tool\n// for the exact source function of every routine below.\n'
_OWN_NOTE = "// Straight-line code (CSE + constant folding + dead-code elimination, SSA form, functor\n// outputs) emitted from this repository's OWN symbolic rigid-body model of the Mini Cheetah\n// (tools/wb_model.py, tools/symbolic.py); not derived from the reference's CasADi-generated code.\n"
from symbolic import Ctx, Dual, emit_function  # noqa: E402
from wb_model import WBModel, hardcoded_params, make_vars, params_from_urdf  # noqa: E402

URDF = "/root/reference/urdf/mini_cheetah_simple_correctedInertia.urdf"
HIP_YAW_URDF = 3.1415
SYNC_EVERY = 1024
NP = 9   # leg constants: abd x, abd y, hip y, abd com y, thigh com y, abd Ixy, abd Iyz, thigh Ixy, thigh Iyz


def leg_constants(leg):
    return [leg["abd_xyz"][0], leg["abd_xyz"][1], leg["hip_xyz"][1], leg["abd"]["com"][1], leg["thigh"]["com"][1],
            leg["abd"]["I"][0][1], leg["abd"]["I"][1][2], leg["thigh"]["I"][0][1], leg["thigh"]["I"][1][2]]


def symbolic_leg(leg, pv):
    """copy of a leg's parameter record with the nine mirrored constants replaced by the variables pv"""
    import copy
    g = copy.deepcopy(leg)
    g["abd_xyz"] = [pv[0], pv[1], leg["abd_xyz"][2]]
    g["hip_xyz"] = [leg["hip_xyz"][0], pv[2], leg["hip_xyz"][2]]
    g["abd"]["com"] = [leg["abd"]["com"][0], pv[3], leg["abd"]["com"][2]]
    g["thigh"]["com"] = [leg["thigh"]["com"][0], pv[4], leg["thigh"]["com"][2]]
    for body, (ixy, iyz) in (("abd", (pv[5], pv[6])), ("thigh", (pv[7], pv[8]))):
        I = [list(r) for r in leg[body]["I"]]
        I[0][1] = I[1][0] = ixy
        I[1][2] = I[2][1] = iyz
        g[body]["I"] = I
    return g


def check_generic(P):
    """everything that is NOT one of the nine constants must be the same for all four legs"""
    import copy
    ref = None
    for leg in P["legs"]:
        g = copy.deepcopy(leg)
        g["abd_xyz"][0] = g["abd_xyz"][1] = g["hip_xyz"][1] = 0
        g["abd"]["com"][1] = g["thigh"]["com"][1] = 0
        for b in ("abd", "thigh"):
            g[b]["I"][0][1] = g[b]["I"][1][0] = g[b]["I"][1][2] = g[b]["I"][2][1] = 0
        if ref is None:
            ref = g
        assert g == ref, "legs differ in more than the nine mirrored constants"


def compact(pairs):
    """[(local_dense_index, Sym)] -> ([(compact_index, Sym)], [local_dense_index of every compact slot])"""
    out, table = [], []
    for dense, s in pairs:
        if s.is_zero():
            continue
        out.append((len(table), s))
        table.append(dense)
    return out, table


def build(ctx, P, hip_yaw, n_arg_params):
    pv = [ctx.var(n_arg_params, i) for i in range(NP)]
    Pg = {"body": P["body"], "legs": [symbolic_leg(P["legs"][0], pv)] + P["legs"][1:]}
    return WBModel(ctx, Pg, hip_yaw)


def local_vars(ctx, arg):
    """18 coordinates of which only the base and leg 0 are variables (local 0..8); the other legs are never touched"""
    loc = [ctx.var(arg, i) for i in range(9)]
    return loc + [ctx.const(0.0)] * 9, loc


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "cafe_mpc_b200/csrc/gen/wb_leg_gen.h"
    P = params_from_urdf(URDF) if os.path.exists(URDF) else hardcoded_params()
    check_generic(P)
    pieces, tables = [], {}
    only = {6, 7, 8}
    # ---------------- Pinocchio-side quantities (hip yaw as in the URDF); inputs ql, vl[, al], P
    ctx = Ctx()
    m = build(ctx, P, HIP_YAW_URDF, 2)
    q, ql = local_vars(ctx, 0)
    v, vl = local_vars(ctx, 1)
    zero = [ctx.const(0.0)] * 18
    nle = m.rnea(q, v, zero, only=only)
    cols = [m.rnea(q, zero, [ctx.const(1.0 if i == c else 0.0) for i in range(18)], gravity=False, only=only) for c in range(9)]
    foot = m.feet(q, v, None)[0]
    o_nle, t_nle = compact([(i, nle[i]) for i in range(9)])
    o_M, t_M = compact([(r + 9 * c, cols[c][r]) for c in range(9) for r in range(c, 9)])
    o_J, t_J = compact([(r + 3 * c, foot["J"][r][c]) for c in range(9) for r in range(3)])
    o_g, t_g = compact([(r, foot["a"][r]) for r in range(3)])
    o_p, t_p = compact([(r, foot["p"][r]) for r in range(3)])
    o_v, t_v = compact([(r, foot["v"][r]) for r in range(3)])
    assert len(t_g) == 3 and len(t_p) == 3 and len(t_v) == 3
    pieces.append(emit_function(ctx, "wbl_terms_leg", 3, [o_nle, o_M, o_J, o_g, o_p, o_v]))
    tables["TM"] = [("NLE", 9, t_nle), ("M", 9, t_M), ("J", 3, t_J), ("GAM", 3, t_g), ("PF", 3, t_p), ("VF", 3, t_v)]
    # ---------------- RNEA derivatives of the leg's inertia group, forward mode (symbolic.Dual), a few directions per routine:
    # inputs ql, vl, al, P; ONE output functor whose index is the compact slot inside the leg's DP block.
    # DP block of a leg: RQ 9 x 6 (columns = local coordinates 3..8) | RV 9 x 6 | DVQ 3 x 6 | DAQ 3 x 6 | DAV 3 x 6 | JTF 6 x 6 (rows 3..8)
    RQ, RV, DVQ, DAQ, DAV, JTF, DPW = 0, 54, 108, 126, 144, 162, 198
    dp_tab = [None] * DPW
    for c in range(3, 9):
        for r in range(9):
            dp_tab[RQ + r + 9 * (c - 3)] = (0, r, c)
            dp_tab[RV + r + 9 * (c - 3)] = (1, r, c)
        for r in range(3):
            dp_tab[DVQ + r + 3 * (c - 3)] = (2, r, c)
            dp_tab[DAQ + r + 3 * (c - 3)] = (3, r, c)
            dp_tab[DAV + r + 3 * (c - 3)] = (4, r, c)
        for r in range(3, 9):
            dp_tab[JTF + (r - 3) + 6 * (c - 3)] = (5, r, c)
    assert all(t is not None for t in dp_tab)

    def dual_vars(ctx, arg, dirs, active):
        z, one = ctx.const(0.0), ctx.const(1.0)
        loc = [ctx.var(arg, i) for i in range(9)]
        return [Dual(loc[i], [(one if (active and i == d) else z) for d in dirs]) for i in range(9)] + [z] * 9

    for kind, dirs in (("q", [3]), ("q", [4]), ("q", [5]), ("q", [6, 7, 8]), ("v", [3, 4, 5]), ("v", [6, 7, 8])):
        ctx = Ctx()
        m = build(ctx, P, HIP_YAW_URDF, 3)
        q = dual_vars(ctx, 0, dirs, kind == "q")
        v = dual_vars(ctx, 1, dirs, kind == "v")
        a = dual_vars(ctx, 2, dirs, False)
        tau = m.rnea(q, v, a, only=only)
        base = RQ if kind == "q" else RV
        outs = [(base + r + 9 * (c - 3), tau[r].t[j]) for j, c in enumerate(dirs) for r in range(9)]
        pieces.append(emit_function(ctx, "wbl_rnea_%s%s" % (kind, "".join(map(str, dirs))), 4, [outs], sync_every=SYNC_EVERY))
    # translation directions: the group's torque does not depend on the base position or on the base linear velocity (checked)
    # ---------------- CasADi-side kinematic partials (hip yaw exactly pi), forward mode; inputs ql, vl, al, F[3], P
    for kind, dirs in (("q", [3, 4, 5]), ("q", [6, 7, 8]), ("v", [3, 4, 5, 6, 7, 8])):
        ctx = Ctx()
        m = build(ctx, P, math.pi, 4)
        q = dual_vars(ctx, 0, dirs, kind == "q")
        v = dual_vars(ctx, 1, dirs, kind == "v")
        a = dual_vars(ctx, 2, dirs, False)
        F = [ctx.var(3, i) for i in range(3)]
        foot = m.feet(q, v, a)[0]
        outs = []
        if kind == "q":
            J = foot["J"]
            jtf = [J[0][i] * F[0] + J[1][i] * F[1] + J[2][i] * F[2] for i in range(9)]
            for j, c in enumerate(dirs):
                outs += [(DVQ + r + 3 * (c - 3), foot["v"][r].t[j]) for r in range(3)]
                outs += [(DAQ + r + 3 * (c - 3), foot["a"][r].t[j]) for r in range(3)]
                for r in range(9):
                    t = jtf[r].t[j] if isinstance(jtf[r], Dual) else ctx.const(0.0)
                    if r < 3:
                        assert t.is_zero(), "d(J^T F)/dq has a non-zero translation row"
                    else:
                        outs.append((JTF + (r - 3) + 6 * (c - 3), t))
        else:
            for j, c in enumerate(dirs):
                outs += [(DAV + r + 3 * (c - 3), foot["a"][r].t[j]) for r in range(3)]
        pieces.append(emit_function(ctx, "wbl_kin_%s%s" % (kind, "".join(map(str, dirs))), 5, [outs], sync_every=SYNC_EVERY))
    tables["DP"] = dp_tab

    tab_path = os.path.join(os.path.dirname(out), "..", "wb_leg_tables.h")
    with open(tab_path, "w") as fh:
        fh.write("// GENERATED by tools/gen_wb_leg.py -- do not edit. Compact output layouts of the leg-generic whole-body routines\n")
        fh.write("// (gen/wb_leg_gen.h) and the mirrored constants of the four legs (urdf/mini_cheetah_simple_correctedInertia.urdf).\n")
        fh.write("#pragma once\n")
        fh.write("// the nine mirrored constants of leg f (FL, FR, HL, HR): abd x, abd y, hip y, abd com y, thigh com y, abd Ixy, abd Iyz, thigh Ixy, thigh Iyz\n")
        fh.write("#define CAFE_WBL_NP %d\n" % NP)
        fh.write("#define CAFE_WBL_LEG_CONSTANTS { \\\n")
        for leg in P["legs"]:
            fh.write("  {" + ", ".join(repr(float(x)) for x in leg_constants(leg)) + "}, \\\n")
        fh.write("}\n\n")
        # compact layouts: per leg  TM = [NLE | M | J | GAM | PF | VF],  DP = [RQ | RV | DVQ | DAQ | DAV | JTF]
        def write_table(name, kinds, rows, colsv):
            fh.write("#define CAFE_WBL_%s_W %d\n" % (name, len(kinds)))
            fh.write("// compact slot -> (array, local row, local column); arrays in the order of the offsets above\n")
            fh.write("#define CAFE_WBL_%s_KIND {%s}\n" % (name, ", ".join(map(str, kinds))))
            fh.write("#define CAFE_WBL_%s_ROW {%s}\n" % (name, ", ".join(map(str, rows))))
            fh.write("#define CAFE_WBL_%s_COL {%s}\n\n" % (name, ", ".join(map(str, colsv))))
        off = 0
        kinds, rows, colsv = [], [], []
        for ki, (pn, nrow, tab) in enumerate(tables["TM"]):
            fh.write("#define CAFE_WBL_TM_%s %d   // %d entries\n" % (pn, off, len(tab)))
            for dense in tab:
                kinds.append(ki); rows.append(dense % nrow); colsv.append(dense // nrow)
            off += len(tab)
        write_table("TM", kinds, rows, colsv)
        for pn, o in (("RQ", 0), ("RV", 54), ("DVQ", 108), ("DAQ", 126), ("DAV", 144), ("JTF", 162)):
            fh.write("#define CAFE_WBL_DP_%s %d\n" % (pn, o))
        write_table("DP", [t[0] for t in tables["DP"]], [t[1] for t in tables["DP"]], [t[2] for t in tables["DP"]])
    with open(out, "w") as fh:
        fh.write(HEADER.replace("tools/casadi2cuda.py", "tools/gen_wb_leg.py (symbolic whole-body model, tools/wb_model.py; leg-generic pieces)").replace(_CASADI_NOTE, _OWN_NOTE))
        fh.write("#include \"../wb_leg_tables.h\"\n")
        fh.write("#ifndef CAFE_GEN_SYNC\n#define CAFE_GEN_SYNC\n#endif\n")
        fh.write("namespace cafe_gen_wbl {\n\n")
        for code, meta in pieces:
            fh.write("// %s: %d ops, output non-zeros %s\n" % (meta["name"], meta["ops"], meta["nnz"]))
            fh.write(code)
            fh.write("\n\n")
        fh.write("}  // namespace cafe_gen_wbl\n")
    for code, meta in pieces:
        print(meta)


if __name__ == "__main__":
    main()
