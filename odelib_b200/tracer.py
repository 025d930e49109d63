"""Trace a user's ODE right-hand side once and emit it as CUDA device code.

The reference calls the user's Python ``f(y, t, ps)`` 100-250 times per solve from inside
scipy's LSODA (Framework.py:656; the demo models are Demo_InfectionStates.ipynb:60-128).
Here the callable is run ONCE with recording proxies.  Every arithmetic operation becomes a
node of a small SSA graph in the order Python evaluated it (so the floating-point evaluation
order of the original is preserved), the graph is differentiated symbolically for the
analytic Jacobian the Rosenbrock path needs, and both are printed as ``__device__`` functions
that NVRTC compiles for sm_100a together with the integrator kernels.

Supported inside the RHS: + - * / ** unary -, abs, numpy ufuncs exp/log/sqrt/sin/cos/tan/tanh/
log10/log2/exp2/log1p/expm1/square/minimum/maximum/power/abs on proxies, indexing / slicing /
iteration of ``y`` and ``ps``, ``np.array([...])``, ``np.dot``/``sum`` over object arrays.
Data-dependent Python control flow (``if y[0] > 0``) cannot be traced and raises TraceError --
there is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np


class TraceError(RuntimeError):
    pass


_UNARY = {"neg", "abs", "exp", "log", "sqrt", "sin", "cos", "tan", "tanh", "log10", "log2", "exp2",
          "log1p", "expm1"}
_BINARY = {"add", "sub", "mul", "div", "pow", "min", "max"}


class Graph:
    """Hash-consed SSA graph.  node = (op, a, b); leaves: ('y',i), ('p',i), ('t',), ('c',value)."""

    def __init__(self):
        self.nodes = []
        self._memo = {}

    def _mk(self, key):
        idx = self._memo.get(key)
        if idx is None:
            idx = len(self.nodes)
            self.nodes.append(key)
            self._memo[key] = idx
        return idx

    def const(self, v):
        v = float(v)
        return self._mk(("c", v.hex()))  # hex: distinguishes -0.0 / keeps NaN hashable

    def leaf(self, kind, i=None):
        return self._mk((kind,) if i is None else (kind, int(i)))

    def is_const(self, i, value=None):
        n = self.nodes[i]
        if n[0] != "c":
            return False
        return value is None or float.fromhex(n[1]) == value

    def cval(self, i):
        return float.fromhex(self.nodes[i][1])

    # -- constructors with light algebraic simplification (only exact identities) --------
    def op(self, name, a, b=None, simplify=False):
        if simplify:
            s = self._simplify(name, a, b)
            if s is not None:
                return s
        return self._mk((name, a) if b is None else (name, a, b))

    def _simplify(self, name, a, b):
        c = self.is_const
        if name == "add":
            if c(a, 0.0):
                return b
            if c(b, 0.0):
                return a
        elif name == "sub":
            if c(b, 0.0):
                return a
            if c(a, 0.0):
                return self.op("neg", b, simplify=True)
        elif name == "mul":
            if c(a, 0.0) or c(b, 0.0):
                return self.const(0.0)
            if c(a, 1.0):
                return b
            if c(b, 1.0):
                return a
            if c(a, -1.0):
                return self.op("neg", b, simplify=True)
            if c(b, -1.0):
                return self.op("neg", a, simplify=True)
        elif name == "div":
            if c(a, 0.0):
                return self.const(0.0)
            if c(b, 1.0):
                return a
        elif name == "neg":
            if c(a):
                return self.const(-self.cval(a))
            if self.nodes[a][0] == "neg":
                return self.nodes[a][1]
        if b is not None and c(a) and c(b) and name in ("add", "sub", "mul"):
            x, y = self.cval(a), self.cval(b)
            return self.const({"add": x + y, "sub": x - y, "mul": x * y}[name])
        return None


class Sym:
    """Recording proxy for one scalar."""
    __slots__ = ("g", "i")
    __array_priority__ = 1000

    def __init__(self, g, i):
        self.g, self.i = g, i

    # -- helpers ---------------------------------------------------------------------------
    def _w(self, other):
        if isinstance(other, Sym):
            if other.g is not self.g:
                raise TraceError("mixing symbols of two traces")
            return other.i
        if isinstance(other, (bool, np.bool_)):
            other = float(other)
        if isinstance(other, (int, float, np.integer, np.floating)):
            return self.g.const(other)
        if isinstance(other, np.ndarray) and other.ndim == 0:
            return self._w(other.item())
        return None

    def _bin(self, name, other, swap=False):
        j = self._w(other)
        if j is None:
            return NotImplemented
        a, b = (j, self.i) if swap else (self.i, j)
        return Sym(self.g, self.g.op(name, a, b))

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __pow__(self, o): return self._bin("pow", o)
    def __rpow__(self, o): return self._bin("pow", o, True)
    def __neg__(self): return Sym(self.g, self.g.op("neg", self.i))
    def __pos__(self): return self
    def __abs__(self): return Sym(self.g, self.g.op("abs", self.i))

    def _unary(self, name):
        return Sym(self.g, self.g.op(name, self.i))

    # numpy calls these method names on object arrays: np.exp(obj_array) -> x.exp()
    def exp(self): return self._unary("exp")
    def log(self): return self._unary("log")
    def sqrt(self): return self._unary("sqrt")
    def sin(self): return self._unary("sin")
    def cos(self): return self._unary("cos")
    def tan(self): return self._unary("tan")
    def tanh(self): return self._unary("tanh")
    def log10(self): return self._unary("log10")
    def log2(self): return self._unary("log2")
    def exp2(self): return self._unary("exp2")
    def log1p(self): return self._unary("log1p")
    def expm1(self): return self._unary("expm1")

    _UFUNC = {"add": "add", "subtract": "sub", "multiply": "mul", "true_divide": "div", "divide": "div",
              "power": "pow", "minimum": "min", "maximum": "max", "fmin": "min", "fmax": "max"}
    _UFUNC1 = {"negative": "neg", "absolute": "abs", "fabs": "abs", "exp": "exp", "log": "log", "sqrt": "sqrt",
               "sin": "sin", "cos": "cos", "tan": "tan", "tanh": "tanh", "log10": "log10", "log2": "log2",
               "exp2": "exp2", "log1p": "log1p", "expm1": "expm1"}

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            raise TraceError(f"numpy ufunc {ufunc.__name__}.{method} is not traceable")
        name = ufunc.__name__
        if any(isinstance(x, np.ndarray) and x.ndim > 0 for x in inputs):
            # array (op) scalar-symbol: broadcast ourselves through an object array
            arrs = [np.asarray(x, dtype=object) if isinstance(x, np.ndarray) else x for x in inputs]
            f = np.frompyfunc(lambda *xs: ufunc(*xs), len(inputs), 1)
            return f(*arrs)
        if name == "square":
            return inputs[0] * inputs[0]
        if name == "reciprocal":
            return 1.0 / inputs[0]
        if name in self._UFUNC1:
            x = inputs[0]
            return Sym(x.g, x.g.op(self._UFUNC1[name], x.i))
        if name in self._UFUNC:
            a, b = inputs
            if isinstance(a, Sym):
                return a._bin(self._UFUNC[name], b)
            return b._bin(self._UFUNC[name], a, True)
        raise TraceError(f"numpy ufunc '{name}' is not supported inside a traced ODE")

    def _nope(self, *_):
        raise TraceError("data-dependent control flow / comparison on a state, parameter or time "
                         "value cannot be traced into a CUDA device function")

    # == / != as well: falling back to object identity would trace `if y[0] == 0:` down ONE branch, silently
    __bool__ = __lt__ = __le__ = __gt__ = __ge__ = __eq__ = __ne__ = _nope
    __float__ = __int__ = __index__ = _nope
    __hash__ = object.__hash__

    def __repr__(self):
        return f"<Sym v{self.i}>"


def _sym_array(g, kind, n):
    a = np.empty(n, dtype=object)
    for i in range(n):
        a[i] = Sym(g, g.leaf(kind, i))
    return a


class TracedModel:
    """Result of tracing: graph + output node ids, with code generation."""

    def __init__(self, graph, outputs, n_state, n_param):
        self.g = graph
        self.outputs = list(outputs)
        self.n_state = n_state
        self.n_param = n_param
        self._jac = None
        self._dfdt = None

    # -- differentiation ---------------------------------------------------------------------
    def _diff(self, wrt):
        """Forward-mode derivative of every node w.r.t. leaf ``wrt``; returns dict node->node."""
        g = self.g
        zero, one = g.const(0.0), g.const(1.0)
        d = {}
        S = lambda n, a, b=None: g.op(n, a, b, simplify=True)
        order = self._reachable(self.outputs)
        for i in order:
            node = g.nodes[i]
            op = node[0]
            if op in ("y", "p", "t", "c"):
                d[i] = one if node == wrt else zero
                continue
            a = node[1]
            da = d[a]
            if len(node) == 2:
                if g.is_const(da, 0.0):
                    d[i] = zero
                elif op == "neg":
                    d[i] = S("neg", da)
                elif op == "exp":
                    d[i] = S("mul", i, da)
                elif op == "log":
                    d[i] = S("div", da, a)
                elif op == "sqrt":
                    d[i] = S("div", da, S("mul", g.const(2.0), i))
                elif op == "sin":
                    d[i] = S("mul", g.op("cos", a), da)
                elif op == "cos":
                    d[i] = S("neg", S("mul", g.op("sin", a), da))
                elif op == "tan":
                    d[i] = S("mul", S("add", one, S("mul", i, i)), da)
                elif op == "tanh":
                    d[i] = S("mul", S("sub", one, S("mul", i, i)), da)
                elif op == "log10":
                    d[i] = S("div", da, S("mul", a, g.const(math.log(10.0))))
                elif op == "log2":
                    d[i] = S("div", da, S("mul", a, g.const(math.log(2.0))))
                elif op == "exp2":
                    d[i] = S("mul", S("mul", i, g.const(math.log(2.0))), da)
                elif op == "log1p":
                    d[i] = S("div", da, S("add", one, a))
                elif op == "expm1":
                    d[i] = S("mul", S("add", i, one), da)
                elif op == "abs":
                    d[i] = S("mul", g.op("sign", a), da)
                else:
                    raise TraceError(f"no derivative rule for {op}")
                continue
            b = node[2]
            db = d[b]
            za, zb = g.is_const(da, 0.0), g.is_const(db, 0.0)
            if za and zb:
                d[i] = zero
            elif op == "add":
                d[i] = S("add", da, db)
            elif op == "sub":
                d[i] = S("sub", da, db)
            elif op == "mul":
                d[i] = S("add", S("mul", da, b), S("mul", a, db))
            elif op == "div":
                # d(a/b) = (da - (a/b) db) / b
                d[i] = S("div", S("sub", da, S("mul", i, db)), b)
            elif op == "pow":
                if zb and g.is_const(b):
                    c = g.cval(b)
                    am1 = a if c == 2.0 else (one if c == 1.0 else g.op("pow", a, g.const(c - 1.0)))
                    d[i] = S("mul", S("mul", g.const(c), am1), da)
                else:
                    # a**b (da*b/a + db*log(a))
                    d[i] = S("mul", i, S("add", S("div", S("mul", da, b), a), S("mul", db, g.op("log", a))))
            elif op in ("min", "max"):
                sel = g.op("sel_" + op, a, b)  # 1.0 where a is selected, else 0.0
                d[i] = S("add", S("mul", sel, da), S("mul", S("sub", one, sel), db))
            else:
                raise TraceError(f"no derivative rule for {op}")
        return d

    def jacobian(self):
        """n x n list of node ids (row = equation, column = state)."""
        if self._jac is None:
            cols = [self._diff(("y", j)) for j in range(self.n_state)]
            self._jac = [[cols[j][o] for j in range(self.n_state)] for o in self.outputs]
            dt = self._diff(("t",))
            self._dfdt = [dt[o] for o in self.outputs]
        return self._jac

    def dfdt(self):
        self.jacobian()
        return self._dfdt

    def jacobian_sparsity(self):
        J = self.jacobian()
        return [[not self.g.is_const(J[i][j], 0.0) for j in range(self.n_state)] for i in range(self.n_state)]

    @property
    def autonomous(self):
        return all(self.g.is_const(x, 0.0) for x in self.dfdt())

    # -- analysis ----------------------------------------------------------------------------
    def _reachable(self, roots):
        g = self.g
        seen = set()
        stack = list(roots)
        while stack:
            i = stack.pop()
            if i in seen:
                continue
            seen.add(i)
            node = g.nodes[i]
            if node[0] not in ("y", "p", "t", "c"):
                stack.extend(node[1:])
        return sorted(seen)   # creation order == a valid topological order

    def flops(self, roots=None):
        """Algorithmic flop count of the traced RHS: every arithmetic node and every function = 1."""
        n = 0
        for i in self._reachable(self.outputs if roots is None else roots):
            if self.g.nodes[i][0] not in ("y", "p", "t", "c"):
                n += 1
        return n

    def evaluate(self, roots, y, t, p):
        """Host evaluation of graph nodes in plain Python floats (tests / diagnostics only)."""
        g, val = self.g, {}
        f1 = {"neg": lambda a: -a, "abs": abs, "exp": math.exp, "log": math.log, "sqrt": math.sqrt,
              "sin": math.sin, "cos": math.cos, "tan": math.tan, "tanh": math.tanh, "log10": math.log10,
              "log2": math.log2, "exp2": lambda a: 2.0 ** a, "log1p": math.log1p, "expm1": math.expm1,
              "sign": lambda a: (a > 0) - (a < 0)}
        f2 = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
              "div": lambda a, b: a / b, "pow": lambda a, b: a ** b, "min": min, "max": max,
              "sel_min": lambda a, b: float(a <= b), "sel_max": lambda a, b: float(a >= b)}
        for i in self._reachable(roots):
            node = g.nodes[i]
            op = node[0]
            if op == "y":
                val[i] = float(y[node[1]])
            elif op == "p":
                val[i] = float(p[node[1]])
            elif op == "t":
                val[i] = float(t)
            elif op == "c":
                val[i] = float.fromhex(node[1])
            elif len(node) == 2:
                val[i] = f1[op](val[node[1]])
            else:
                val[i] = f2[op](val[node[1]], val[node[2]])
        return [val[r] for r in roots]

    # -- cooperative kernels: one lane, a few outputs ---------------------------------------------
    def _canonical(self, o):
        """Shape of output ``o``: its expression DAG with the leaf INDICES abstracted away.  Returns (signature, leaves):
        signature = (ops in evaluation order, root reference, kinds of the leaf slots); leaves = [(kind, index)] per slot.
        Two outputs with the same signature are the same computation on different states / parameters."""
        g = self.g
        local, ops, leaves, slot = {}, [], [], {}

        def rec(i):
            r = local.get(i)
            if r is not None:
                return r
            node = g.nodes[i]
            op = node[0]
            if op in ("y", "p"):
                key = (op, node[1])
                if key not in slot:
                    slot[key] = len(leaves)
                    leaves.append(key)
                r = ("L", op, slot[key])
            elif op == "t":
                r = ("T",)
            elif op == "c":
                r = ("C", node[1])
            else:
                args = tuple(rec(a) for a in node[1:])
                ops.append((op,) + args)
                r = ("N", len(ops) - 1)
            local[i] = r
            return r

        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 20000))
        try:
            root = rec(o)
        finally:
            sys.setrecursionlimit(old)
        return (tuple(ops), root, tuple(k for k, _ in leaves)), leaves

    def slice_plan(self, lanes):
        """How ``lanes`` lanes share the right-hand side of one system without every lane evaluating all of it.

        Outputs are grouped into CLASSES of identical shape (``_canonical``); components are laid out class by class,
        ``lanes`` per round, so that the lanes of a round run the same code on different leaves -- looked up in an index
        table -- instead of each lane evaluating the whole traced RHS and keeping 1/lanes of it.  A round that holds the
        end of one class and the start of the next runs both codes under lane predicates.
        -> dict: rounds, perm [rounds*lanes] (output index or -1), classes [(signature, n_slots)],
                 segments [per round: (class id, lane_lo, lane_hi, table offset)], table (flat leaf indices)."""
        sigs, members, leaves_of = {}, [], {}
        for k, o in enumerate(self.outputs):
            sig, leaves = self._canonical(o)
            cid = sigs.setdefault(sig, len(sigs))
            if cid == len(members):
                members.append([])
            members[cid].append(k)
            leaves_of[k] = leaves
        classes = [None] * len(sigs)
        for sig, cid in sigs.items():
            classes[cid] = (sig, len(sig[2]))
        order = [k for cid in range(len(members)) for k in members[cid]]
        cls_of = {k: cid for cid in range(len(members)) for k in members[cid]}
        rounds = (len(order) + lanes - 1) // lanes
        perm = order + [-1] * (rounds * lanes - len(order))
        table, segments = [], []
        for c in range(rounds):
            segs, lane = [], 0
            while lane < lanes and perm[c * lanes + lane] >= 0:
                cid = cls_of[perm[c * lanes + lane]]
                hi = lane
                while hi < lanes and perm[c * lanes + hi] >= 0 and cls_of[perm[c * lanes + hi]] == cid:
                    hi += 1
                segs.append((cid, lane, hi, len(table)))
                for l in range(lane, hi):
                    table += [idx for _, idx in leaves_of[perm[c * lanes + l]]]
                lane = hi
            segments.append(segs)
        return {"rounds": rounds, "perm": perm, "classes": classes, "segments": segments, "table": table, "lanes": lanes}

    def evaluate_sliced(self, plan, y, t, p):
        """Host evaluation THROUGH a slice plan (tests): every output from its class code and its table row."""
        f1 = {"neg": lambda a: -a, "abs": abs, "exp": math.exp, "log": math.log, "sqrt": math.sqrt, "sin": math.sin,
              "cos": math.cos, "tan": math.tan, "tanh": math.tanh, "log10": math.log10, "log2": math.log2,
              "exp2": lambda a: 2.0 ** a, "log1p": math.log1p, "expm1": math.expm1, "sign": lambda a: (a > 0) - (a < 0)}
        f2 = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b, "div": lambda a, b: a / b,
              "pow": lambda a, b: a ** b, "min": min, "max": max, "sel_min": lambda a, b: float(a <= b),
              "sel_max": lambda a, b: float(a >= b)}
        out = [None] * len(self.outputs)
        G = plan["lanes"]
        for c, segs in enumerate(plan["segments"]):
            for cid, lo, hi, off in segs:
                (ops, root, kinds), ns = plan["classes"][cid]
                for lane in range(lo, hi):
                    ix = plan["table"][off + (lane - lo) * ns: off + (lane - lo + 1) * ns]
                    vals = []

                    def ref(r):
                        if r[0] == "L":
                            return float(y[ix[r[2]]]) if r[1] == "y" else float(p[ix[r[2]]])
                        if r[0] == "T":
                            return float(t)
                        if r[0] == "C":
                            return float.fromhex(r[1])
                        return vals[r[1]]
                    for op in ops:
                        vals.append(f1[op[0]](ref(op[1])) if len(op) == 2 else f2[op[0]](ref(op[1]), ref(op[2])))
                    out[plan["perm"][c * G + lane]] = ref(root)
        return out

    def _emit_sliced(self, L, lanes, fmad):
        """odl_rhs_slice<...>: this lane's outputs of every round (see slice_plan)."""
        plan = self.slice_plan(lanes)
        g = self.g
        L.append(f"#define ODL_COOP_SLICED 1")
        L.append(f"#define ODL_SLICE_G {lanes}")
        L.append(f"#define ODL_CS {plan['rounds']}")
        L.append("__device__ const short ODL_SLICE_PERM[ODL_CS * ODL_SLICE_G] = {" + ", ".join(str(v) for v in plan["perm"]) + "};")
        tab = plan["table"] or [0]
        L.append(f"__device__ const short ODL_SLICE_IX[{len(tab)}] = {{" + ", ".join(str(v) for v in tab) + "};")
        L.append("template <class YV, class PV>")
        L.append("__device__ __forceinline__ void odl_rhs_slice(const YV& y, const double t, const PV& p, const int sub, "
                 "double (&f)[ODL_CS]) {")
        uid = 0
        for c, segs in enumerate(plan["segments"]):
            L.append(f"  f[{c}] = 0.0;")
            for cid, lo, hi, off in segs:
                (ops, root, kinds), ns = plan["classes"][cid]
                full = lo == 0 and hi == lanes
                L.append("  {" if full else f"  if (sub >= {lo} && sub < {hi}) {{")
                L.append(f"    const short* ix = ODL_SLICE_IX + {off} + (sub - {lo}) * {ns};")
                for s_, kind in enumerate(kinds):
                    L.append(f"    const double l{s_} = {kind}[ix[{s_}]];")
                names = []

                def ref(r):
                    if r[0] == "L":
                        return f"l{r[2]}"
                    if r[0] == "T":
                        return "t"
                    if r[0] == "C":
                        return _c_double(float.fromhex(r[1]))
                    return names[r[1]]
                for op in ops:
                    a = ref(op[1])
                    b = ref(op[2]) if len(op) == 3 else None
                    node = (op[0], None, None)
                    if op[0] == "pow" and op[2][0] == "C":
                        expr = _c_pow_const(a, float.fromhex(op[2][1]), fmad)
                    else:
                        expr = _c_expr(op[0], a, b, fmad, None, None) if op[0] != "pow" else f"pow({a}, {b})"
                    names.append(f"w{uid}")
                    L.append(f"    const double w{uid} = {expr};")
                    uid += 1
                L.append(f"    f[{c}] = {ref(root)};")
                L.append("  }")
        L.append("}")

    # -- code generation -----------------------------------------------------------------------
    def _emit(self, roots, lines, fmad, names=None):
        g = self.g
        names = {} if names is None else names
        for i in self._reachable(roots):
            if i in names:
                continue
            node = g.nodes[i]
            op = node[0]
            if op == "y":
                names[i] = f"y[{node[1]}]"
            elif op == "p":
                names[i] = f"p[{node[1]}]"
            elif op == "t":
                names[i] = "t"
            elif op == "c":
                names[i] = _c_double(float.fromhex(node[1]))
            else:
                a = names[node[1]]
                b = names[node[2]] if len(node) == 3 else None
                names[i] = f"v{i}"
                lines.append(f"  const double v{i} = {_c_expr(op, a, b, fmad, g, node)};")
        return names

    def cuda_source(self, fmad=True, observe_groups=None, coop_lanes=0):
        """CUDA source of odl_rhs / odl_jac / odl_dfdt / odl_observe for this model.

        fmad=False prints every add/sub/mul as an ``__d*_rn`` intrinsic, which the compiler never
        contracts into FMAs: the RHS then rounds exactly like the CPython evaluation of the model.
        observe_groups: list (one per output column) of tuples of state indices that are summed
        (Framework.py:659-664); default identity.
        """
        n, P = self.n_state, self.n_param
        L = []
        L.append(f"#define ODL_N {n}")
        L.append(f"#define ODL_P {P}")
        groups = observe_groups if observe_groups is not None else [(i,) for i in range(n)]
        L.append(f"#define ODL_NOUT {len(groups)}")
        L.append(f"#define ODL_RHS_FLOPS {self.flops()}")
        L.append(f"#define ODL_AUTONOMOUS {1 if self.autonomous else 0}")
        # vector arguments are template parameters: plain register arrays for thread-per-system kernels, shared-memory
        # views / slice-keeping output proxies for the cooperative (several lanes per system) kernels
        L.append("template <class YV, class PV, class DV>")
        L.append("__device__ __forceinline__ void odl_rhs(const YV& y, const double t, const PV& p, DV& dy) {")
        names = self._emit(self.outputs, L, fmad)
        for k, o in enumerate(self.outputs):
            L.append(f"  dy[{k}] = {names[o]};")
        L.append("}")
        # Jacobian (dense n x n, row-major, structural zeros written as literal 0.0)
        J = self.jacobian()
        L.append("template <class YV, class PV>")
        L.append("__device__ __forceinline__ void odl_jac(const YV& y, const double t, const PV& p, "
                 "double (&J)[ODL_N][ODL_N]) {")
        roots = [J[i][j] for i in range(n) for j in range(n)]
        names = self._emit(roots, L, True)
        for i in range(n):
            for j in range(n):
                L.append(f"  J[{i}][{j}] = {names[J[i][j]]};")
        L.append("}")
        L.append("template <class YV, class PV, class DV>")
        L.append("__device__ __forceinline__ void odl_dfdt(const YV& y, const double t, const PV& p, DV& ft) {")
        names = self._emit(self.dfdt(), L, True)
        for k, o in enumerate(self.dfdt()):
            L.append(f"  ft[{k}] = {names[o]};")
        L.append("}")
        if coop_lanes:
            self._emit_sliced(L, int(coop_lanes), fmad)
        L.append("template <class YV>")
        L.append("__device__ __forceinline__ void odl_observe(const YV& y, double (&out)[ODL_NOUT]) {")
        for c, grp in enumerate(groups):
            # numpy's .sum(axis=1) over <8 columns adds left to right
            expr = f"y[{grp[0]}]"
            for s in grp[1:]:
                expr = f"__dadd_rn({expr}, y[{s}])"
            L.append(f"  out[{c}] = {expr};")
        L.append("}")
        return "\n".join(L) + "\n"


def _c_double(v):
    if math.isnan(v):
        return "__longlong_as_double(0x7ff8000000000000LL)"
    if math.isinf(v):
        return ("-" if v < 0 else "") + "__longlong_as_double(0x7ff0000000000000LL)"
    return f"({v!r})" if v < 0 or (v == 0 and math.copysign(1, v) < 0) else repr(v)


def _c_pow_const(a, c, fmad):
    if c == 2.0:
        return f"{a} * {a}" if fmad else f"__dmul_rn({a}, {a})"
    if c == 0.5:
        return f"sqrt({a})"
    return f"pow({a}, {_c_double(c)})"


def _c_expr(op, a, b, fmad, g, node):
    if op in ("add", "sub", "mul"):
        if fmad:
            return f"{a} {'+' if op == 'add' else '-' if op == 'sub' else '*'} {b}"
        if op == "sub":
            return f"__dadd_rn({a}, -({b}))"
        return f"__d{op}_rn({a}, {b})"
    if op == "div":
        return f"{a} / {b}"
    if op == "neg":
        return f"-({a})"
    if op == "abs":
        return f"fabs({a})"
    if op == "sign":
        return f"(({a}) > 0.0 ? 1.0 : (({a}) < 0.0 ? -1.0 : 0.0))"
    if op == "pow":
        if g.is_const(node[2]):
            c = g.cval(node[2])
            if c == 2.0:
                return f"{a} * {a}" if fmad else f"__dmul_rn({a}, {a})"
            if c == 0.5:
                return f"sqrt({a})"
        return f"pow({a}, {b})"
    if op == "min":
        return f"fmin({a}, {b})"
    if op == "max":
        return f"fmax({a}, {b})"
    if op == "sel_min":
        return f"(({a}) <= ({b}) ? 1.0 : 0.0)"
    if op == "sel_max":
        return f"(({a}) >= ({b}) ? 1.0 : 0.0)"
    if op in _UNARY:
        return f"{op}({a})"
    raise TraceError(f"cannot print op {op}")


def trace(ode, n_state, n_param):
    """Run ``ode(y, t, ps)`` once on recording proxies (same call signature as Framework.py:656 uses)."""
    g = Graph()
    y = _sym_array(g, "y", n_state)
    p = _sym_array(g, "p", n_param)
    t = Sym(g, g.leaf("t"))
    try:
        res = ode(y, t, p)
    except TraceError:
        raise
    except Exception as exc:  # noqa: BLE001 - report anything the user's function raised on proxies
        raise TraceError(f"the ODE function could not be traced symbolically: {exc!r}") from exc
    res = np.asarray(res, dtype=object).ravel()
    if res.size != n_state:
        raise TraceError(f"ODE returned {res.size} derivatives for {n_state} state variables")
    outs = []
    for r in res:
        if isinstance(r, Sym):
            outs.append(r.i)
        elif isinstance(r, (int, float, np.integer, np.floating)):
            outs.append(g.const(float(r)))
        else:
            raise TraceError(f"ODE returned an untraceable value of type {type(r).__name__}")
    return TracedModel(g, outs, n_state, n_param)
