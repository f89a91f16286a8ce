"""Tiny symbolic-scalar layer on top of the hash-consed expression DAG of casadi2cuda.py.

Used to write the Mini-Cheetah whole-body rigid-body model (tools/wb_model.py) ONCE with operator
overloading, then (a) evaluate it numerically for validation against the reference's known-answer vectors,
(b) differentiate it (forward mode on the DAG, memoised) and (c) emit straight-line CUDA/C functions with the
same emitter that re-emits the CasADi graphs."""
import math

from casadi2cuda import Dag, fmt_const


class Ctx:
    def __init__(self):
        self.d = Dag()
        self._dmemo = {}

    def const(self, v):
        return Sym(self, self.d.const(v))

    def var(self, arg, idx):
        return Sym(self, self.d.mk("arg", arg, idx))

    # ---- numeric evaluation
    def evaluate(self, outs, inputs):
        """outs: list of Sym; inputs: {arg_index: sequence}. Returns list of floats."""
        d = self.d
        need = set()
        stack = [o.i for o in outs]
        while stack:
            i = stack.pop()
            if i in need:
                continue
            need.add(i)
            nd = d.nodes[i]
            if nd[0] not in ("const", "arg"):
                stack.extend(nd[1:])
        val = {}
        for i in sorted(need):
            nd = d.nodes[i]
            op = nd[0]
            if op == "const": val[i] = nd[1]
            elif op == "arg": val[i] = float(inputs[nd[1]][nd[2]])
            elif op == "+": val[i] = val[nd[1]] + val[nd[2]]
            elif op == "-": val[i] = val[nd[1]] - val[nd[2]]
            elif op == "*": val[i] = val[nd[1]] * val[nd[2]]
            elif op == "/": val[i] = val[nd[1]] / val[nd[2]]
            elif op == "neg": val[i] = -val[nd[1]]
            elif op == "sin": val[i] = math.sin(val[nd[1]])
            elif op == "cos": val[i] = math.cos(val[nd[1]])
            elif op == "sqrt": val[i] = math.sqrt(val[nd[1]])
            else: raise ValueError(op)
        return [val[o.i] for o in outs]

    # ---- differentiation (forward mode, one variable at a time, memoised per (node, var))
    def diff(self, node, var):
        d = self.d
        key = (node, var)
        m = self._dmemo.get(key)
        if m is not None:
            return m
        # iterative post-order to avoid recursion limits
        stack = [node]
        while stack:
            i = stack[-1]
            if (i, var) in self._dmemo:
                stack.pop()
                continue
            nd = d.nodes[i]
            op = nd[0]
            if op == "const":
                self._dmemo[(i, var)] = d.const(0.0); stack.pop(); continue
            if op == "arg":
                self._dmemo[(i, var)] = d.const(1.0 if i == var else 0.0); stack.pop(); continue
            pend = [a for a in nd[1:] if (a, var) not in self._dmemo]
            if pend:
                stack.extend(pend)
                continue
            da = self._dmemo[(nd[1], var)]
            if op in ("+", "-"):
                db = self._dmemo[(nd[2], var)]
                r = d.bin(op, da, db)
            elif op == "*":
                db = self._dmemo[(nd[2], var)]
                r = d.bin("+", d.bin("*", da, nd[2]), d.bin("*", nd[1], db))
            elif op == "/":
                db = self._dmemo[(nd[2], var)]
                # (a/b)' = (a' - (a/b) b') / b
                r = d.bin("/", d.bin("-", da, d.bin("*", i, db)), nd[2])
            elif op == "neg":
                r = d.un("neg", da)
            elif op == "sin":
                r = d.bin("*", d.un("cos", nd[1]), da)
            elif op == "cos":
                r = d.un("neg", d.bin("*", d.un("sin", nd[1]), da))
            elif op == "sqrt":
                r = d.bin("/", da, d.bin("*", d.const(2.0), i))
            else:
                raise ValueError(op)
            self._dmemo[(i, var)] = r
            stack.pop()
        return self._dmemo[key]


class Sym:
    __slots__ = ("c", "i")

    def __init__(self, c, i):
        self.c = c
        self.i = i

    def _w(self, o):
        return o if isinstance(o, Sym) else self.c.const(o)

    def __add__(self, o):
        if isinstance(o, Dual): return NotImplemented
        return Sym(self.c, self.c.d.bin("+", self.i, self._w(o).i))
    __radd__ = __add__
    def __sub__(self, o):
        if isinstance(o, Dual): return NotImplemented
        return Sym(self.c, self.c.d.bin("-", self.i, self._w(o).i))
    def __rsub__(self, o): return Sym(self.c, self.c.d.bin("-", self._w(o).i, self.i))
    def __mul__(self, o):
        if isinstance(o, Dual): return NotImplemented
        return Sym(self.c, self.c.d.bin("*", self.i, self._w(o).i))
    __rmul__ = __mul__
    def __truediv__(self, o): return Sym(self.c, self.c.d.bin("/", self.i, self._w(o).i))
    def __neg__(self): return Sym(self.c, self.c.d.un("neg", self.i))
    def sin(self): return Sym(self.c, self.c.d.un("sin", self.i))
    def cos(self): return Sym(self.c, self.c.d.un("cos", self.i))
    def is_zero(self): return self.c.d.is_const(self.i) and self.c.d.cval(self.i) == 0.0
    def d(self, var): return Sym(self.c, self.c.diff(self.i, var.i))


class Dual:
    """value + tangents along a few directions (forward mode by operator overloading). The nodes of the tangents are created right
    after the nodes of their value, so the emitter (which keeps creation order) interleaves them: the live set of the generated code
    stays a small multiple of that of the plain evaluation. Differentiating the finished DAG output by output (Sym.d) instead keeps
    every intermediate of the evaluation alive for every derivative column: the 8.7 k-op RNEA-derivative routine of a leg spilled
    13 KB / 26 KB (stores / loads) per call at 128 registers that way (tools/gen_wb_leg.py)."""
    __slots__ = ("v", "t")

    def __init__(self, v, t):
        self.v = v
        self.t = list(t)

    def _l(self, o):
        if isinstance(o, Dual): return o
        c = self.v.c
        return Dual(o if isinstance(o, Sym) else c.const(o), [c.const(0.0)] * len(self.t))

    def __add__(self, o): o = self._l(o); return Dual(self.v + o.v, [a + b for a, b in zip(self.t, o.t)])
    __radd__ = __add__
    def __sub__(self, o): o = self._l(o); return Dual(self.v - o.v, [a - b for a, b in zip(self.t, o.t)])
    def __rsub__(self, o): o = self._l(o); return Dual(o.v - self.v, [b - a for a, b in zip(self.t, o.t)])
    def __mul__(self, o):
        o = self._l(o)
        v = self.v * o.v
        return Dual(v, [a * o.v + self.v * b for a, b in zip(self.t, o.t)])
    __rmul__ = __mul__
    def __neg__(self): return Dual(-self.v, [-a for a in self.t])
    def sin(self): s = self.v.sin(); c = self.v.cos(); return Dual(s, [c * a for a in self.t])
    def cos(self): c = self.v.cos(); s = self.v.sin(); return Dual(c, [-(s * a) for a in self.t])
    def is_zero(self): return self.v.is_zero() and all(a.is_zero() for a in self.t)


def emit_function(ctx, name, n_in, outputs, out_names=None, decl="CAFE_HD", sync_every=0):
    """outputs: list (one per output array) of lists of (dense_index, Sym). Zero entries are skipped.
    sync_every > 0: a CAFE_GEN_SYNC marker after every sync_every operations (a CTA barrier where the includer defines it so: the
    warps of a CTA then stay inside the same stretch of straight-line code and share instruction-cache lines; empty by default).
    Emits `template<class O0,...> CAFE_HD void name(const double* i0, ..., O0 o0, ...)`."""
    d = ctx.d
    live = set()
    stack = []
    by_node = {}
    for oi, lst in enumerate(outputs):
        for dense, s in lst:
            if s.is_zero():
                continue
            by_node.setdefault(s.i, []).append((oi, dense))
            stack.append(s.i)
    while stack:
        i = stack.pop()
        if i in live:
            continue
        live.add(i)
        nd = d.nodes[i]
        if nd[0] not in ("const", "arg"):
            stack.extend(nd[1:])
    n_out = len(outputs)
    tparams = ", ".join("class O%d" % i for i in range(n_out))
    args = ", ".join(["const double* __restrict__ i%d" % i for i in range(n_in)] + ["O%d o%d" % (i, i) for i in range(n_out)])
    lines = ["template <%s>" % tparams, "%s void %s(%s) {" % (decl, name, args)]
    n_ops = 0

    def ref(i):
        nd = d.nodes[i]
        return fmt_const(nd[1]) if nd[0] == "const" else "t%d" % i

    def store(i):
        for (oi, dense) in by_node.get(i, []):
            lines.append("  o%d(%d, %s);" % (oi, dense, ref(i)))

    for i in by_node:
        if d.is_const(i):
            store(i)
    for i in sorted(live):
        nd = d.nodes[i]
        op = nd[0]
        if op == "const":
            continue
        if op == "arg":
            lines.append("  const double t%d = i%d[%d];" % (i, nd[1], nd[2]))
        elif op in ("+", "-", "*", "/"):
            lines.append("  const double t%d = %s %s %s;" % (i, ref(nd[1]), op, ref(nd[2]))); n_ops += 1
        elif op == "neg":
            lines.append("  const double t%d = -%s;" % (i, ref(nd[1]))); n_ops += 1
        else:
            lines.append("  const double t%d = %s(%s);" % (i, op, ref(nd[1]))); n_ops += 1
        store(i)
        if sync_every > 0 and op != "arg" and n_ops % sync_every == 0 and lines[-1] != "  CAFE_GEN_SYNC":
            lines.append("  CAFE_GEN_SYNC")
    lines.append("}")
    nnz = [sum(1 for _, s in lst if not s.is_zero()) for lst in outputs]
    return "\n".join(lines), {"name": name, "ops": n_ops, "nnz": nnz}
