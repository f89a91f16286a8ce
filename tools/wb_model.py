"""Mini-Cheetah whole-body rigid-body model, Pinocchio-free, written over symbolic scalars.

Replaces, for the WB phases, what the reference gets from Pinocchio 2.6.10 (not vendored, not installed):
  kinematic tree      MHPC/MHPC-Trajopt/PinocchioInteface.cpp:17-56 (PX,PY,PZ,RZ,RY + URDF with an RX root joint)
  inertial data       urdf/mini_cheetah_simple_correctedInertia.urdf
  FK / Jacobians / frame velocity / classical acceleration, CRBA, non-linear effects, RNEA
                      call sites MHPC/MHPC-Trajopt/WBM.cpp:266-362, :375-421, :430-454, :474, :514-531
Generalised coordinates q = [x y z yaw pitch roll | FL(abd hip knee) FR HL HR], v = qdot (WBM.h:13-21).

The hip-pitch joint placement yaw is a parameter: 3.1415 for everything that replaces Pinocchio, pi for everything
that replaces the CasADi-generated kinematic partials."""
import math
import xml.etree.ElementTree as ET

from symbolic import Ctx, Dual, Sym

GRAV = 9.81
LEGS = ["fl", "fr", "hl", "hr"]


def parse_urdf(path):
    root = ET.parse(path).getroot()
    links = {}
    for l in root.findall("link"):
        ine = l.find("inertial")
        if ine is None:
            links[l.get("name")] = None
            continue
        m = float(ine.find("mass").get("value"))
        o = ine.find("origin")
        com = [float(v) for v in o.get("xyz").split()]
        assert o.get("rpy") in (None, "0 0 0", "0.0 0.0 0.0")
        i = ine.find("inertia")
        I = {k: float(i.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")}
        links[l.get("name")] = {"m": m, "com": com, "I": [[I["ixx"], I["ixy"], I["ixz"]], [I["ixy"], I["iyy"], I["iyz"]], [I["ixz"], I["iyz"], I["izz"]]]}
    joints = {}
    for j in root.findall("joint"):
        o = j.find("origin")
        ax = j.find("axis")
        joints[j.find("child").get("link")] = {
            "parent": j.find("parent").get("link"), "type": j.get("type"),
            "xyz": [float(v) for v in o.get("xyz").split()], "rpy": [float(v) for v in o.get("rpy").split()],
            "axis": [float(v) for v in ax.get("xyz").split()] if ax is not None else None}
    return links, joints


def hardcoded_params():
    """The URDF numbers, for use where /root/reference is absent."""
    legs = []
    for name, sx, sy in (("fl", 1, 1), ("fr", 1, -1), ("hl", -1, 1), ("hr", -1, -1)):
        legs.append({
            "abd_xyz": [sx * 0.19, sy * 0.049, 0.0], "hip_xyz": [0.0, sy * 0.062, 0.0], "knee_xyz": [0.0, 0.0, -0.209], "foot_xyz": [0.0, 0.0, -0.195],
            "abd": {"m": 0.54, "com": [0.0, sy * 0.036, 0.0], "I": [[0.000381, sy * 0.000058, 0.00000045], [sy * 0.000058, 0.000560, sy * 0.00000095], [0.00000045, sy * 0.00000095, 0.000444]]},
            "thigh": {"m": 0.634, "com": [0.0, sy * 0.016, -0.02], "I": [[0.001983, sy * 0.000245, 0.000013], [sy * 0.000245, 0.002103, sy * 0.0000015], [0.000013, sy * 0.0000015, 0.000408]]},
            "shank": {"m": 0.064, "com": [0.0, 0.0, -0.061], "I": [[0.000245, 0, 0], [0, 0.000248, 0], [0, 0, 0.000006]]}})
    body = {"m": 3.3, "com": [0.0, 0.0, 0.0], "I": [[0.011253, 0, 0], [0, 0.036203, 0], [0, 0, 0.042673]]}
    return {"body": body, "legs": legs}


def params_from_urdf(path):
    links, joints = parse_urdf(path)
    legs = []
    for n in LEGS:
        a, t, s, f = joints["abduct_" + n], joints["thigh_" + n], joints["shank_" + n], joints["foot_" + n]
        assert a["axis"] == [1, 0, 0] and t["axis"] == [0, 1, 0] and s["axis"] == [0, 1, 0]
        assert t["rpy"][:2] == [0, 0] and abs(t["rpy"][2] - 3.1415) < 1e-12 and a["rpy"] == [0, 0, 0] and s["rpy"] == [0, 0, 0]
        legs.append({"abd_xyz": a["xyz"], "hip_xyz": t["xyz"], "knee_xyz": s["xyz"], "foot_xyz": f["xyz"],
                     "abd": links["abduct_" + n], "thigh": links["thigh_" + n], "shank": links["shank_" + n]})
    return {"body": links["body"], "legs": legs}


# ------------------------------------------------------------------ small linear algebra over Sym / float
def cross(a, b): return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
def add(a, b): return [x + y for x, y in zip(a, b)]
def sub(a, b): return [x - y for x, y in zip(a, b)]
def scale(a, s): return [x * s for x in a]
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def matvec(M, v): return [M[i][0] * v[0] + M[i][1] * v[1] + M[i][2] * v[2] for i in range(3)]
def matmul(A, B): return [[A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] for j in range(3)] for i in range(3)]
def transpose(M): return [[M[j][i] for j in range(3)] for i in range(3)]


class WBModel:
    """18 one-dof joints; joint i: parent joint, type 'P'/'R', axis (unit vector in the joint frame), placement
    (R_off, p_off) in the parent joint frame, attached body inertia (or None)."""

    def __init__(self, ctx, params, hip_yaw):
        self.c = ctx
        # numbers become constants of the DAG; symbolic scalars pass through (leg-generic pieces: tools/gen_wb_leg.py
        # hands the mirrored leg constants in as input variables)
        K = lambda x: x if isinstance(x, (Sym, Dual)) else ctx.const(x)
        self.K = K
        I3 = [[K(1.0 if i == j else 0.0) for j in range(3)] for i in range(3)]
        z3 = [K(0.0)] * 3
        cy, sy = math.cos(hip_yaw), math.sin(hip_yaw)
        Rz = [[K(cy), K(-sy), K(0.0)], [K(sy), K(cy), K(0.0)], [K(0.0), K(0.0), K(1.0)]]
        J = []
        ex, ey, ez = [K(1.0), K(0.0), K(0.0)], [K(0.0), K(1.0), K(0.0)], [K(0.0), K(0.0), K(1.0)]
        for i, (t, ax) in enumerate((("P", ex), ("P", ey), ("P", ez), ("R", ez), ("R", ey), ("R", ex))):
            J.append({"parent": i - 1, "type": t, "axis": ax, "R": I3, "p": z3, "body": params["body"] if i == 5 else None})
        self.foot_parent, self.foot_off = [], []
        for l, leg in enumerate(params["legs"]):
            b = 6 + 3 * l
            J.append({"parent": 5, "type": "R", "axis": ex, "R": I3, "p": [K(v) for v in leg["abd_xyz"]], "body": leg["abd"]})
            J.append({"parent": b, "type": "R", "axis": ey, "R": Rz, "p": [K(v) for v in leg["hip_xyz"]], "body": leg["thigh"]})
            J.append({"parent": b + 1, "type": "R", "axis": ey, "R": I3, "p": [K(v) for v in leg["knee_xyz"]], "body": leg["shank"]})
            self.foot_parent.append(b + 2)
            self.foot_off.append([K(v) for v in leg["foot_xyz"]])
        self.joints = J

    def rot(self, axis, q):
        c, s, K = q.cos(), q.sin(), self.K
        ax = [self.c.d.cval(a.i) for a in axis]
        if ax == [1.0, 0.0, 0.0]: return [[K(1.0), K(0.0), K(0.0)], [K(0.0), c, -s], [K(0.0), s, c]]
        if ax == [0.0, 1.0, 0.0]: return [[c, K(0.0), s], [K(0.0), K(1.0), K(0.0)], [-s, K(0.0), c]]
        return [[c, -s, K(0.0)], [s, c, K(0.0)], [K(0.0), K(0.0), K(1.0)]]

    def kinematics(self, q, v=None, a=None):
        """World-frame forward pass. Returns per joint: R, p, axis_w, omega, vlin, alpha, alin (alin = acceleration of
        the joint-frame origin, gravity NOT included)."""
        K = self.K
        z3 = [K(0.0)] * 3
        out = []
        for i, j in enumerate(self.joints):
            if j["parent"] < 0:
                Rp, pp, wp, vp, alp, ap = [[K(1.0 if r == c else 0.0) for c in range(3)] for r in range(3)], z3, z3, z3, z3, z3
            else:
                P = out[j["parent"]]
                Rp, pp, wp, vp, alp, ap = P["R"], P["p"], P["w"], P["v"], P["al"], P["a"]
            Ro = matmul(Rp, j["R"])
            axw = matvec(Ro, j["axis"])
            r = matvec(Rp, j["p"])
            qd = v[i] if v is not None else K(0.0)
            qdd = a[i] if a is not None else K(0.0)
            if j["type"] == "R":
                R = matmul(Ro, self.rot(j["axis"], q[i]))
                p = add(pp, r)
                w = add(wp, scale(axw, qd))
                vl = add(vp, cross(wp, r))
                al = add(add(alp, scale(axw, qdd)), cross(wp, scale(axw, qd)))
                ac = add(add(ap, cross(alp, r)), cross(wp, cross(wp, r)))
            else:
                R = Ro
                r = add(r, scale(axw, q[i]))
                p = add(pp, r)
                w = wp
                vl = add(add(vp, cross(wp, r)), scale(axw, qd))
                al = alp
                ac = add(add(add(ap, cross(alp, r)), cross(wp, cross(wp, r))), add(scale(cross(wp, axw), qd * 2.0), scale(axw, qdd)))
            out.append({"R": R, "p": p, "axw": axw, "w": w, "v": vl, "al": al, "a": ac})
        return out

    def rnea(self, q, v, a, gravity=True, only=None):
        """tau = M(q) a + nle(q, v)  (gravity (0,0,-9.81)); world-frame Newton-Euler. `only`: set of joint indices whose bodies
        are kept (RNEA is linear in the inertias: tau = tau[trunk] + sum over legs tau[leg f], and tau[leg f] depends on the base
        and on leg f only)."""
        K = self.K
        kin = self.kinematics(q, v, a)
        n = len(self.joints)
        F = [[K(0.0)] * 3 for _ in range(n)]
        N = [[K(0.0)] * 3 for _ in range(n)]
        for i, j in enumerate(self.joints):
            b = j["body"]
            if b is None or (only is not None and i not in only):
                continue
            k = kin[i]
            cw = matvec(k["R"], [K(x) for x in b["com"]])
            acom = add(add(k["a"], cross(k["al"], cw)), cross(k["w"], cross(k["w"], cw)))
            if gravity:
                acom = add(acom, [K(0.0), K(0.0), K(GRAV)])
            f = scale(acom, b["m"])
            Ib = [[K(x) for x in row] for row in b["I"]]
            Iw = matmul(matmul(k["R"], Ib), transpose(k["R"]))
            nc = add(matvec(Iw, k["al"]), cross(k["w"], matvec(Iw, k["w"])))
            F[i] = f
            N[i] = add(nc, cross(cw, f))
        tau = [None] * n
        for i in range(n - 1, -1, -1):
            j = self.joints[i]
            k = kin[i]
            tau[i] = dot(k["axw"], N[i]) if j["type"] == "R" else dot(k["axw"], F[i])
            pa = j["parent"]
            if pa >= 0:
                F[pa] = add(F[pa], F[i])
                N[pa] = add(add(N[pa], N[i]), cross(sub(k["p"], kin[pa]["p"]), F[i]))
        return tau

    def mass_matrix(self, q):
        K = self.K
        n = len(self.joints)
        zero = [K(0.0)] * n
        M = [[None] * n for _ in range(n)]
        for c in range(n):
            e = [K(1.0 if i == c else 0.0) for i in range(n)]
            col = self.rnea(q, zero, e, gravity=False)
            for r in range(n):
                M[r][c] = col[r]
        return M

    def feet(self, q, v=None, a=None):
        """Per foot: position, velocity, acceleration (d^2 p/dt^2 for the given v and a), Jacobian (3 x 18)."""
        K = self.K
        kin = self.kinematics(q, v, a)
        res = []
        for f in range(4):
            k = kin[self.foot_parent[f]]
            r = matvec(k["R"], self.foot_off[f])
            p = add(k["p"], r)
            vel = add(k["v"], cross(k["w"], r))
            acc = add(add(k["a"], cross(k["al"], r)), cross(k["w"], cross(k["w"], r)))
            Jc = [[K(0.0)] * 18 for _ in range(3)]
            i = self.foot_parent[f]
            while i >= 0:
                ki = kin[i]
                col = cross(ki["axw"], sub(p, ki["p"])) if self.joints[i]["type"] == "R" else ki["axw"]
                for rr in range(3):
                    Jc[rr][i] = col[rr]
                i = self.joints[i]["parent"]
            res.append({"p": p, "v": vel, "a": acc, "J": Jc})
        return res


def make_vars(ctx, arg, n):
    return [ctx.var(arg, i) for i in range(n)]
