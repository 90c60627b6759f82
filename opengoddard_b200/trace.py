"""Symbolic tracing of the user callbacks.

The reference evaluates `dynamics / equality / inequality / cost / running_cost`
eagerly in numpy, once per perturbed decision vector
(/root/reference/OpenGoddard/optimize.py:674,685,703,706,727 called n+1 times per
Jacobian by scipy/optimize/_numdiff.py:683-712).  Here each callback is called ONCE
with a `TraceView` of the problem whose accessors return `Sym` objects; the unmodified
user code (python operators, numpy ufuncs, `[0]` / `[-1]` / `[a:b]` indexing,
boolean-mask assignment) then records a hash-consed expression DAG instead of
numbers.  `opengoddard_b200.tape` lowers the DAG to the register programs the CUDA
kernels interpret, keeping the reference's operation order so that results agree with
numpy to the last bit for + - * / sqrt and to the libm ulp for transcendental calls.

Structure exploited (and checked): expressions over whole node arrays are *pointwise
in the node* -- row k only reads node k of its phase -- which is what makes the
Jacobian sparse and lets the sweep kernel recompute only the rows a perturbed variable
can touch.  Anything that is not node-local (a vector mixed with `x[0]`, shifted
slices, ...) is expanded element by element into scalar rows, which is always correct.
"""
import numbers

import numpy as np


class TraceError(RuntimeError):
    """The callback did something the tracer cannot record (data-dependent python
    control flow, conversion of a traced value to float, unsupported numpy call)."""


# --------------------------------------------------------------------------- DAG
class Node:
    __slots__ = ("op", "args", "value", "uid")

    def __init__(self, op, args, value, uid):
        self.op, self.args, self.value, self.uid = op, args, value, uid

    def __repr__(self):
        if self.op == "const":
            return "c(%r)" % self.value
        if self.op in ("blk", "var"):
            return "%s%r" % (self.op, self.args)
        return "%s(%s)" % (self.op, ",".join("n%d" % a.uid for a in self.args))


UNARY = ("neg", "sqrt", "exp", "log", "sin", "cos", "tan", "abs", "square", "recip", "asin",
         "acos", "atan", "sinh", "cosh", "tanh", "log10", "sign", "floor", "ceil", "not")
BINARY = ("add", "sub", "mul", "div", "pow", "min", "max", "atan2", "lt", "le", "gt", "ge",
          "eq", "ne", "and", "or")


class Graph:
    """Hash-consing node store (common subexpressions are shared)."""

    def __init__(self):
        self._table = {}
        self.nodes = []

    def _intern(self, op, args, value=None):
        key = (op, tuple(a.uid if isinstance(a, Node) else a for a in args),
               None if value is None else np.float64(value).tobytes())
        n = self._table.get(key)
        if n is None:
            n = Node(op, tuple(args), value, len(self.nodes))
            self._table[key] = n
            self.nodes.append(n)
        return n

    def const(self, v):
        return self._intern("const", (), float(v))

    def blk(self, section, blk):
        return self._intern("blk", (section, blk))

    def var(self, index):
        return self._intern("var", (index,))

    def nodec(self, section, cid):
        """per-node constant vector `cid` of phase `section` (TraceContext.node_const) at the node"""
        return self._intern("nodec", (section, cid))

    def op(self, name, *args):
        # x*1.0 and x/1.0 are exact identities in IEEE-754 (the reference multiplies and
        # divides by unit 1.0 all the time, optimize.py:284,1018,1125)
        if name == "mul":
            a, b = args
            if b.op == "const" and b.value == 1.0:
                return a
            if a.op == "const" and a.value == 1.0:
                return b
        elif name == "div":
            a, b = args
            if b.op == "const" and b.value == 1.0:
                return a
        return self._intern(name, args)

    def interp(self, table_id, x):
        """linear lookup table `table_id` (see TraceContext.table_of) evaluated at node x"""
        return self._intern("interp", (x,), float(table_id))


def substitute_blocks(graph, node, fn, memo, fn_nodec=None):
    """Rebuild `node` with every ('blk', s, b) leaf replaced by fn(s, b) (and every ('nodec', s, cid)
    leaf by fn_nodec(s, cid) if given)."""
    got = memo.get(node.uid)
    if got is not None:
        return got
    if node.op == "blk":
        out = fn(*node.args)
    elif node.op == "nodec":
        out = fn_nodec(*node.args) if fn_nodec is not None else node
    elif node.op in ("const", "var"):
        out = node
    elif node.op == "interp":
        out = graph.interp(int(node.value), substitute_blocks(graph, node.args[0], fn, memo, fn_nodec))
    else:
        out = graph.op(node.op, *[substitute_blocks(graph, a, fn, memo, fn_nodec) for a in node.args])
    memo[node.uid] = out
    return out


def leaves(node, seen=None, out=None):
    """Set of ('blk', s, b) / ('var', i) leaves reachable from node."""
    if seen is None:
        seen, out = set(), set()
    stack = [node]
    while stack:
        n = stack.pop()
        if n.uid in seen:
            continue
        seen.add(n.uid)
        if n.op == "blk":
            out.add(("blk",) + n.args)
        elif n.op == "var":
            out.add(("var",) + n.args)
        elif n.op == "nodec":
            out.add(("nodec",) + n.args)
        else:
            stack.extend(n.args)
    return out


# --------------------------------------------------------------------------- context
class TraceContext:
    """Layout of the problem being traced (reference optimize.py:237-245, :781)."""

    def __init__(self, prob):
        self.graph = Graph()
        self.nsec = int(prob.number_of_section)
        self.nodes = [int(v) for v in prob.nodes]
        self.nstates = [int(v) for v in prob.number_of_states]
        self.ncontrols = [int(v) for v in prob.number_of_controls]
        self.off, self.g0 = [], []
        o = g = 0
        for s in range(self.nsec):
            self.off.append(o)
            self.g0.append(g)
            o += (self.nstates[s] + self.ncontrols[s]) * self.nodes[s]
            g += self.nodes[s]
        self.gtot = g
        self.nvars = o + self.nsec
        self.tables = []              # lookup tables met while tracing: dicts x, y, variant, ...
        self._table_ids = {}
        self.node_consts = [[] for _ in range(self.nsec)]    # per phase: constant vectors, one value per node
        self._node_const_ids = {}

    def node_const(self, s, values):
        """Register a per-node constant vector of phase s (prob.time[s], prob.tau[s], a user table of one
        value per node ...); returns its id within the phase."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.nodes[s],)
        key = (s, v.tobytes())
        if key not in self._node_const_ids:
            self.node_consts[s].append(v.copy())
            self._node_const_ids[key] = len(self.node_consts[s]) - 1
        return self._node_const_ids[key]

    def const_vector(self, rng, values):
        """A numpy vector aligned with the node range rng as a traced vector of per-node constants, or None
        if rng does not consist of whole phases (the caller then expands element by element)."""
        glo, ghi = rng
        parts, at = {}, glo
        for s in range(self.nsec):
            lo, hi = self.g0[s], self.g0[s] + self.nodes[s]
            if hi <= glo or lo >= ghi:
                continue
            if lo < glo or hi > ghi:
                return None
            parts[s] = self.graph.nodec(s, self.node_const(s, values[lo - glo:hi - glo]))
            at = hi
        return Sym(self, rng, parts) if parts else None

    def is_time_var(self, index):
        return index >= self.nvars - self.nsec

    def table_of(self, f):
        """Register a scipy.interpolate.interp1d object; returns its table id."""
        key = id(f)
        if key in self._table_ids:
            return self._table_ids[key]
        kind = getattr(f, "_kind", None)
        call = getattr(getattr(f, "_call", None), "__name__", "")
        y = np.asarray(f.y, dtype=float)
        if kind != "linear" or y.ndim != 1 or call not in ("_call_linear_np", "_call_linear"):
            raise TraceError("only 1-D linear scipy.interpolate.interp1d tables can be compiled for the "
                             "device (got kind=%r, y.ndim=%d)" % (kind, y.ndim))
        extrap = bool(getattr(f, "_extrapolate", False))
        nan = float("nan")
        below = above = nan
        if not extrap and not f.bounds_error:
            below = float(np.asarray(f._fill_value_below, dtype=float).ravel()[0])
            above = float(np.asarray(f._fill_value_above, dtype=float).ravel()[0])
        self.tables.append(dict(x=np.asarray(f.x, dtype=float).copy(), y=y.copy(),
                                variant=0 if call == "_call_linear_np" else 1, extrapolate=extrap,
                                fill_below=below, fill_above=above, keep=f))
        self._table_ids[key] = len(self.tables) - 1
        return self._table_ids[key]

    def table_of_arrays(self, xp, fp, left, right):
        """Register the table of a numpy.interp(x, xp, fp, left, right) call; returns its table id.
        numpy.interp is the formula ogb_interp calls variant 0 (compiled_base.c arr_interp), with
        `left` / `right` (default fp[0] / fp[-1]) outside the table."""
        xp = np.asarray(xp, dtype=float)
        fp = np.asarray(fp, dtype=float)
        if xp.ndim != 1 or fp.shape != xp.shape or len(xp) < 2 or not (np.diff(xp) >= 0).all():
            raise TraceError("numpy.interp on traced values needs 1-D ascending xp and fp of equal length >= 2")
        below = float(fp[0]) if left is None else float(left)
        above = float(fp[-1]) if right is None else float(right)
        key = ("np.interp", xp.tobytes(), fp.tobytes(), below, above)
        if key not in self._table_ids:
            self.tables.append(dict(x=xp.copy(), y=fp.copy(), variant=0, extrapolate=False,
                                    fill_below=below, fill_above=above, keep=None))
            self._table_ids[key] = len(self.tables) - 1
        return self._table_ids[key]

    def section_of(self, g):
        for s in range(self.nsec - 1, -1, -1):
            if g >= self.g0[s]:
                return s
        raise IndexError(g)

    def var_index(self, s, blk, k):
        return self.off[s] + blk * self.nodes[s] + k


# --------------------------------------------------------------------------- Sym
def is_sym(x):
    return isinstance(x, (Sym, SymList))


def _is_number(x):
    return isinstance(x, (numbers.Real, np.floating, np.integer, np.bool_)) or \
        (isinstance(x, np.ndarray) and x.ndim == 0)


class Sym:
    """A traced scalar (`rng is None`, `parts` is a Node) or a traced vector over the
    global node range rng = (glo, ghi) whose element for a node of phase s is the
    pointwise expression parts[s]."""

    __array_priority__ = 1000.0
    __hash__ = None

    def __init__(self, ctx, rng, parts):
        self.ctx, self.rng, self.parts = ctx, rng, parts

    # ---- structure
    @property
    def is_scalar(self):
        return self.rng is None

    def __len__(self):
        if self.rng is None:
            raise TypeError("len() of a traced scalar")
        return self.rng[1] - self.rng[0]

    @property
    def shape(self):
        return () if self.rng is None else (len(self),)

    @property
    def size(self):
        return 1 if self.rng is None else len(self)

    @property
    def ndim(self):
        return 0 if self.rng is None else 1

    def copy(self):
        return Sym(self.ctx, self.rng, self.parts if self.rng is None else dict(self.parts))

    def __array__(self, *a, **k):
        raise TraceError("a traced value was converted to a numpy array; this numpy call is "
                         "not supported by the OpenGoddard-B200 tracer")

    def __float__(self):
        raise TraceError("a traced value was converted to float (data-dependent python code "
                         "cannot be compiled for the device)")

    __int__ = __index__ = __float__

    def __bool__(self):
        raise TraceError("a traced value was used in an `if` / `while` (data-dependent python "
                         "control flow cannot be compiled for the device); use boolean-mask "
                         "assignment or numpy.where instead")

    def __iter__(self):
        if self.rng is None:
            raise TypeError("iteration over a traced scalar")
        return (self[i] for i in range(len(self)))

    # ---- element access
    def _element(self, k):
        n = len(self)
        if k < -n or k >= n:
            raise IndexError("index %d out of range for traced vector of length %d" % (k, n))
        g = self.rng[0] + (k + n if k < 0 else k)
        ctx = self.ctx
        s = ctx.section_of(g)
        kl = g - ctx.g0[s]
        node = substitute_blocks(ctx.graph, self.parts[s],
                                 lambda sec, b: ctx.graph.var(ctx.var_index(sec, b, kl)), {},
                                 lambda sec, cid: ctx.graph.const(ctx.node_consts[sec][cid][kl]))
        return Sym(ctx, None, node)

    def __getitem__(self, key):
        if self.rng is None:
            raise TraceError("indexing a traced scalar")
        if isinstance(key, (int, np.integer)):
            return self._element(int(key))
        if isinstance(key, slice):
            lo, hi, step = key.indices(len(self))
            if step == 1:
                hi = max(hi, lo)
                glo, ghi = self.rng[0] + lo, self.rng[0] + hi
                parts = {s: p for s, p in self.parts.items()
                         if self.ctx.g0[s] < ghi and self.ctx.g0[s] + self.ctx.nodes[s] > glo}
                return Sym(self.ctx, (glo, ghi), parts)
            return SymList([self._element(i) for i in range(lo, hi, step)])
        if isinstance(key, (list, np.ndarray)) and np.asarray(key).dtype.kind in "iu":
            return SymList([self._element(int(i)) for i in np.asarray(key).ravel()])
        raise TraceError("unsupported index %r on a traced vector" % (key,))

    def __setitem__(self, key, value):
        # boolean-mask assignment, e.g. h[h < -100.0] = -100.0
        if isinstance(key, Sym) and key.rng == self.rng and self.rng is not None:
            val = _coerce(self.ctx, value)
            for s in list(self.parts):
                vnode = val.parts if val.rng is None else val.parts[s]
                self.parts[s] = self.ctx.graph.op("sel", key.parts[s], vnode, self.parts[s])
            return
        raise TraceError("only boolean-mask assignment `x[mask] = value` is supported on traced vectors")

    # ---- arithmetic
    def _bin(self, name, other, swap=False):
        if isinstance(other, SymList):
            return NotImplemented
        if isinstance(other, (list, tuple)) and len(other) and all(_is_number(v) for v in other):
            other = np.asarray(other, dtype=np.float64)
        if isinstance(other, np.ndarray) and other.ndim > 0:
            # one value per node of the same node range: a per-node constant vector (prob.time[s], prob.tau[s],
            # a user table) -- stays node-local; anything else is expanded element by element
            if other.ndim == 1 and self.rng is not None and len(other) == len(self) and other.dtype.kind in "fiu":
                vec = self.ctx.const_vector(self.rng, other.astype(np.float64))
                if vec is not None:
                    return self._bin(name, vec, swap)
            if other.ndim == 1 and self.rng is None and other.dtype.kind in "fiu":
                # a traced scalar (a final time) with a table of one value per node: the table's node range is the
                # one phase with that many nodes, or all phases
                ctx = self.ctx
                cands = [(ctx.g0[s], ctx.g0[s] + ctx.nodes[s]) for s in range(ctx.nsec) if ctx.nodes[s] == len(other)]
                if len(other) == ctx.gtot and ctx.nsec > 1:
                    cands.append((0, ctx.gtot))
                if len(cands) == 1:
                    return self._bin(name, ctx.const_vector(cands[0], other.astype(np.float64)), swap)
            return SymList.from_any(self)._bin(name, other, swap)
        try:
            o = _coerce(self.ctx, other)
        except TypeError:
            return NotImplemented
        a, b = (o, self) if swap else (self, o)
        g = self.ctx.graph
        if a.rng is None and b.rng is None:
            return Sym(self.ctx, None, g.op(name, a.parts, b.parts))
        if a.rng is not None and b.rng is not None:
            if a.rng != b.rng:
                if len(a) != len(b):
                    raise ValueError("operands could not be broadcast together with shapes "
                                     "(%d,) (%d,)" % (len(a), len(b)))
                return SymList.from_any(a)._bin(name, SymList.from_any(b))
            return Sym(self.ctx, a.rng, {s: g.op(name, a.parts[s], b.parts[s]) for s in a.parts})
        vec = a if a.rng is not None else b
        if a.rng is None:
            return Sym(self.ctx, vec.rng, {s: g.op(name, a.parts, p) for s, p in vec.parts.items()})
        return Sym(self.ctx, vec.rng, {s: g.op(name, p, b.parts) for s, p in vec.parts.items()})

    def _interp_table(self, tid):
        g = self.ctx.graph
        if self.rng is None:
            return Sym(self.ctx, None, g.interp(tid, self.parts))
        return Sym(self.ctx, self.rng, {s: g.interp(tid, p) for s, p in self.parts.items()})

    def _interp(self, f):
        """this value pushed through the scipy interp1d object `f`"""
        g, tid = self.ctx.graph, self.ctx.table_of(f)
        if self.rng is None:
            return Sym(self.ctx, None, g.interp(tid, self.parts))
        return Sym(self.ctx, self.rng, {s: g.interp(tid, p) for s, p in self.parts.items()})

    def _un(self, name):
        g = self.ctx.graph
        if self.rng is None:
            return Sym(self.ctx, None, g.op(name, self.parts))
        return Sym(self.ctx, self.rng, {s: g.op(name, p) for s, p in self.parts.items()})

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __neg__(self): return self._un("neg")
    def __pos__(self): return self
    def __abs__(self): return self._un("abs")
    def __lt__(self, o): return self._bin("lt", o)
    def __le__(self, o): return self._bin("le", o)
    def __gt__(self, o): return self._bin("gt", o)
    def __ge__(self, o): return self._bin("ge", o)
    def __eq__(self, o): return self._bin("eq", o)
    def __ne__(self, o): return self._bin("ne", o)
    def __and__(self, o): return self._bin("and", o)
    def __or__(self, o): return self._bin("or", o)
    def __invert__(self): return self._un("not")

    def __pow__(self, o):
        # numpy's scalar-exponent fast paths (x**2 -> square, x**0.5 -> sqrt, x**-1 ->
        # reciprocal, x**1 -> x); everything else is pow()
        if _is_number(o):
            e = float(o)
            if e == 2.0:
                return self._un("square")
            if e == 1.0:
                return self
            if e == 0.5:
                return self._un("sqrt")
            if e == -1.0:
                return self._un("recip")
        return self._bin("pow", o)

    def __rpow__(self, o): return self._bin("pow", o, True)

    # ---- numpy protocol
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs.get("out") is not None:
            raise TraceError("numpy.%s.%s is not supported on traced values" % (ufunc.__name__, method))
        return _apply_ufunc(ufunc.__name__, inputs)

    def __array_function__(self, func, types, args, kwargs):
        return _apply_function(func, args, kwargs)


_UFUNC_UNARY = {"negative": "neg", "sqrt": "sqrt", "exp": "exp", "log": "log", "sin": "sin",
                "cos": "cos", "tan": "tan", "absolute": "abs", "fabs": "abs", "square": "square",
                "reciprocal": "recip", "arcsin": "asin", "arccos": "acos", "arctan": "atan",
                "sinh": "sinh", "cosh": "cosh", "tanh": "tanh", "log10": "log10", "sign": "sign",
                "floor": "floor", "ceil": "ceil", "logical_not": "not"}
_UFUNC_BINARY = {"add": "add", "subtract": "sub", "multiply": "mul", "divide": "div",
                 "true_divide": "div", "minimum": "min", "maximum": "max", "fmin": "min",
                 "fmax": "max", "arctan2": "atan2", "less": "lt", "less_equal": "le",
                 "greater": "gt", "greater_equal": "ge", "equal": "eq", "not_equal": "ne",
                 "logical_and": "and", "logical_or": "or"}


def _first_sym(items):
    for x in items:
        if isinstance(x, Sym):
            return x
        if isinstance(x, SymList):
            return x
    return None


def _apply_ufunc(name, inputs):
    anchor = _first_sym(inputs)
    if isinstance(anchor, SymList) or any(isinstance(x, SymList) for x in inputs):
        lst = SymList.from_any(anchor)
        if name in _UFUNC_UNARY:
            return SymList([_apply_ufunc(name, (e,)) for e in lst.items])
        n = len(lst.items)
        cols = [SymList.from_any(x, n, lst.ctx).items for x in inputs]
        return SymList([_apply_ufunc(name, tuple(c[i] for c in cols)) for i in range(n)])
    ctx = anchor.ctx
    if name in _UFUNC_UNARY:
        return inputs[0]._un(_UFUNC_UNARY[name])
    if name in _UFUNC_BINARY:
        a, b = inputs
        if isinstance(a, Sym):
            return a._bin(_UFUNC_BINARY[name], b)
        return b._bin(_UFUNC_BINARY[name], a, True)
    if name == "power":
        a, b = inputs
        return a.__pow__(b) if isinstance(a, Sym) else b.__rpow__(a)
    if name == "expm1":                       # exp(x) - 1 (absolute error ~1e-16: enough for an FD Jacobian)
        return _apply_ufunc("exp", (inputs[0],)) - 1.0
    if name == "log1p":
        return _apply_ufunc("log", (inputs[0] + 1.0,))
    if name == "isnan":
        return _apply_ufunc("not_equal", (inputs[0], inputs[0]))
    if name in ("remainder", "mod"):          # sign of the divisor, like numpy: x - floor(x / y) * y
        a, b = inputs
        return a - _apply_ufunc("floor", (a / b,)) * b
    if name == "fmod":                        # sign of the dividend, like C: x - trunc(x / y) * y
        a, b = inputs
        q = a / b
        tr = _apply_function(np.where, (q < 0.0, _apply_ufunc("ceil", (q,)), _apply_ufunc("floor", (q,))), {})
        return a - tr * b
    if name == "heaviside":
        x, h0 = inputs
        return _apply_function(np.where, (x < 0.0, 0.0, _apply_function(np.where, (x > 0.0, 1.0, h0), {})), {})
    if name == "float_power":
        a, b = inputs
        return a.__pow__(b) if isinstance(a, Sym) else b.__rpow__(a)
    if name == "hypot":                       # sqrt(a^2 + b^2): within an ulp of libm's hypot
        a, b = inputs
        return _apply_ufunc("sqrt", (a * a + b * b,))
    if name == "exp2":
        return inputs[0].__rpow__(2.0)
    if name == "log2":
        return _apply_ufunc("log", (inputs[0],)) / float(np.log(2.0))
    if name == "deg2rad" or name == "radians":
        return inputs[0] * (np.pi / 180.0)
    if name == "rad2deg" or name == "degrees":
        return inputs[0] * (180.0 / np.pi)
    if name == "positive":
        return inputs[0]
    raise TraceError("numpy.%s is not supported on traced values" % name)


def _apply_function(func, args, kwargs):
    name = getattr(func, "__name__", str(func))
    if name == "where" and len(args) == 3:
        cond, a, b = args
        anchor = _first_sym(args)
        ctx = anchor.ctx
        cond, a, b = _coerce(ctx, cond), _coerce(ctx, a), _coerce(ctx, b)
        rng = next((x.rng for x in (cond, a, b) if x.rng is not None), None)
        g = ctx.graph
        if rng is None:
            return Sym(ctx, None, g.op("sel", cond.parts, a.parts, b.parts))
        for x in (cond, a, b):
            if x.rng is not None and x.rng != rng:
                raise TraceError("numpy.where over traced vectors of different node ranges")
        secs = next(x.parts for x in (cond, a, b) if x.rng is not None).keys()
        pick = lambda x, s: x.parts if x.rng is None else x.parts[s]
        return Sym(ctx, rng, {s: g.op("sel", pick(cond, s), pick(a, s), pick(b, s)) for s in secs})
    if name in ("zeros_like", "ones_like", "full_like") and len(args) >= 1 and isinstance(args[0], Sym):
        value = 0.0 if name == "zeros_like" else 1.0 if name == "ones_like" else \
            (args[1] if len(args) > 1 else kwargs["fill_value"])
        x = args[0]                            # a constant with the node range of x (a select that always
        return _apply_function(np.where, (x == x, value, value), {})     # takes `value`, also for NaN x)
    if name in ("mean", "max", "min", "amax", "amin", "dot", "norm", "prod") and len(args) >= 1 and is_sym(args[0]) \
            and not kwargs:
        # reductions over the nodes are not node-local: expanded element by element (scalar rows / cost)
        items = SymList.from_any(args[0]).items
        if name == "mean":
            return _apply_function(np.sum, (args[0],), {}) / float(len(items))
        if name in ("max", "amax", "min", "amin"):
            uf = "maximum" if name in ("max", "amax") else "minimum"
            acc = items[0]
            for e in items[1:]:
                acc = _apply_ufunc(uf, (acc, e))
            return acc
        if name == "prod":
            acc = items[0]
            for e in items[1:]:
                acc = acc * e
            return acc
        if name == "dot" and len(args) == 2:
            other = SymList.from_any(args[1], len(items), args[0].ctx).items
            acc = items[0] * other[0]
            for a_, b_ in zip(items[1:], other[1:]):
                acc = acc + a_ * b_
            return acc
        if name == "norm" and len(args) == 1:
            acc = items[0] * items[0]
            for e in items[1:]:
                acc = acc + e * e
            return _apply_ufunc("sqrt", (acc,))
    if name == "clip" and len(args) + len(kwargs) >= 2 and is_sym(args[0]):
        lo = args[1] if len(args) > 1 else kwargs.get("a_min", kwargs.get("min"))
        hi = args[2] if len(args) > 2 else kwargs.get("a_max", kwargs.get("max"))
        out = args[0]
        if lo is not None:
            out = np.maximum(out, lo)
        if hi is not None:
            out = np.minimum(out, hi)
        return out
    if name == "interp" and len(args) >= 3 and is_sym(args[0]) and not is_sym(args[1]) and not is_sym(args[2]):
        if kwargs.get("period") is not None:
            raise TraceError("numpy.interp(period=...) is not supported on traced values")
        x = args[0]
        ctx = x.ctx
        tid = ctx.table_of_arrays(args[1], args[2], kwargs.get("left", args[3] if len(args) > 3 else None),
                                  kwargs.get("right", args[4] if len(args) > 4 else None))
        if isinstance(x, SymList):
            return SymList([e._interp_table(tid) if isinstance(e, Sym) else
                            float(np.interp(e, args[1], args[2], left=ctx.tables[tid]["fill_below"],
                                            right=ctx.tables[tid]["fill_above"])) for e in x.items])
        return x._interp_table(tid)
    if name in ("hstack", "concatenate") and len(args) >= 1:
        seq = list(args[0])
        out = []
        for x in seq:
            if isinstance(x, Sym) and x.rng is None:
                out.append(x)
            elif is_sym(x):
                out.extend(SymList.from_any(x).items)
            else:
                out.extend(np.atleast_1d(np.asarray(x, dtype=float)).tolist())
        return SymList(out)
    if name in ("atleast_1d",) and len(args) == 1:
        return args[0]
    if name == "sum" and len(args) == 1 and is_sym(args[0]):
        items = SymList.from_any(args[0]).items
        acc = items[0]
        for e in items[1:]:
            acc = acc + e
        return acc
    raise TraceError("numpy.%s is not supported on traced values" % name)


def _coerce(ctx, x):
    if isinstance(x, Sym):
        return x
    if _is_number(x):
        return Sym(ctx, None, ctx.graph.const(float(x)))
    raise TypeError("cannot combine %r with a traced value" % type(x))


class SymList:
    """Fallback: an explicit list of traced scalars / numbers (not node-local)."""

    __array_priority__ = 1001.0
    __hash__ = None

    def __init__(self, items):
        self.items = list(items)
        self.ctx = next((e.ctx for e in self.items if isinstance(e, Sym)), None)

    @staticmethod
    def from_any(x, n=None, ctx=None):
        if isinstance(x, SymList):
            return x
        if isinstance(x, Sym):
            if x.rng is None:
                return SymList([x] * (n or 1))
            return SymList([x._element(i) for i in range(len(x))])
        arr = np.asarray(x, dtype=float)
        if arr.ndim == 0:
            return SymList([float(arr)] * (n or 1))
        return SymList([float(v) for v in arr.ravel()])

    def __len__(self):
        return len(self.items)

    @property
    def shape(self):
        return (len(self.items),)

    def __iter__(self):
        return iter(self.items)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return SymList(self.items[key])
        return self.items[key]

    def _bin(self, name, other, swap=False):
        n = len(self.items)
        o = SymList.from_any(other, n, self.ctx)
        if len(o.items) != n:
            raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)"
                             % (n, len(o.items)))
        out = []
        for a, b in zip(self.items, o.items):
            if swap:
                a, b = b, a
            if isinstance(a, Sym):
                out.append(a._bin(name, b))
            elif isinstance(b, Sym):
                out.append(b._bin(name, a, True))
            else:
                out.append(_NUMERIC[name](a, b))
        return SymList(out)

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __neg__(self): return SymList([-e for e in self.items])

    def __pow__(self, o):
        return SymList([e ** o for e in self.items])

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            raise TraceError("numpy.%s.%s is not supported on traced values" % (ufunc.__name__, method))
        return _apply_ufunc(ufunc.__name__, inputs)

    def __array_function__(self, func, types, args, kwargs):
        return _apply_function(func, args, kwargs)


_NUMERIC = {"add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
            "div": lambda a, b: a / b, "pow": lambda a, b: a ** b, "min": min, "max": max}


# --------------------------------------------------------------------------- the traced view of a Problem
class TraceView:
    """What the callbacks receive as `prob` while being traced: attribute reads fall
    through to the real Problem (nodes, unit_states, index_states, ...), the value
    accessors return Sym objects.  Mirrors reference optimize.py:271-375."""

    def __init__(self, prob, ctx):
        object.__setattr__(self, "_prob", prob)
        object.__setattr__(self, "_ctx", ctx)

    # Attributes of the real Problem a traced callback may read: layout, LGL data, units and the
    # index helpers -- none of them depends on the decision vector.  Everything else the class
    # defines (time_update, time_knots, to_csv, plot, solve, the setters ...) reads or writes prob.p
    # and would bake the CURRENT decision vector into the tape as constants, so it is refused
    # unless overridden below with a traced version.
    _PASS = frozenset((
        "nodes", "number_of_states", "number_of_controls", "number_of_section", "number_of_variables",
        "number_of_param", "div", "D", "time_init", "t0",
        "unit_states", "unit_controls", "unit_time", "maxIterator", "iterator", "bounds",
        "knot_states_smooth", "dynamics", "cost", "running_cost", "cost_derivative", "equality",
        "inequality", "index_states", "index_controls", "index_time_final", "time_to_tau",
        "_division_states", "_division_controls", "_span_state", "_span_control", "bounds_arrays",
        "backend", "device"))

    def _node_vectors(self, arrays):
        ctx = self._ctx
        return [ctx.const_vector((ctx.g0[s], ctx.g0[s] + ctx.nodes[s]), np.asarray(arrays[s], dtype=float))
                for s in range(ctx.nsec)]

    def __getattr__(self, name):
        if name == "p":
            raise TraceError("direct access to prob.p inside a callback cannot be traced; "
                             "use prob.states()/controls()/time_final()")
        prob = self._prob
        # per-node data of the problem: traced vectors of per-node constants, so that they combine with traced
        # scalars (final times) as well as with state / control vectors and stay node-local
        if name in ("tau", "w", "time"):
            return self._node_vectors(getattr(prob, name))
        if name == "time_all_section":
            return self._ctx.const_vector((0, self._ctx.gtot), np.asarray(prob.time_all_section, dtype=float))
        if name in self._PASS or (name in vars(prob) and not hasattr(type(prob), name)):
            return getattr(prob, name)          # whitelisted, or an attribute the user attached
        if hasattr(prob, name):
            raise TraceError("prob.%s reads or writes the decision vector and cannot be used inside a "
                             "callback that is compiled for the device" % name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        raise TraceError("callbacks must not modify the Problem while being evaluated")

    def _block(self, lo, hi, section, unit):
        ctx = self._ctx
        s = range(ctx.nsec)[section]
        N = ctx.nodes[s]
        rel = lo - ctx.off[s]
        nb = ctx.nstates[s] + ctx.ncontrols[s]
        if hi - lo != N or rel < 0 or rel % N or rel // N >= nb:
            raise TraceError("accessor resolves to p[%d:%d], which is not one block of phase %d"
                             % (lo, hi, s))
        node = ctx.graph.op("mul", ctx.graph.blk(s, rel // N), ctx.graph.const(unit))
        return Sym(ctx, (ctx.g0[s], ctx.g0[s] + N), {s: node})

    def states(self, state, section):
        hi, lo = self._prob._division_states(state, section)
        return self._block(lo, hi, section, self._prob.unit_states[section][state])

    def controls(self, control, section):
        hi, lo = self._prob._division_controls(control, section)
        return self._block(lo, hi, section, self._prob.unit_controls[section][control])

    def _all(self, getter, index):
        ctx = self._ctx
        parts = {}
        for s in range(ctx.nsec):
            parts.update(getter(index, s).parts)
        return Sym(ctx, (0, ctx.gtot), parts)

    def states_all_section(self, state):
        return self._all(self.states, state)

    def controls_all_section(self, control):
        return self._all(self.controls, control)

    def _time(self, index):
        ctx = self._ctx
        v = ctx.nvars + index if index < 0 else index
        node = ctx.graph.op("mul", ctx.graph.var(v), ctx.graph.const(self._prob.unit_time))
        return Sym(ctx, None, node)

    def time_start(self, section):
        if section == 0:
            return self._prob.t0
        return self._time(range(-self._ctx.nsec - 1, 0)[section])

    def time_final(self, section):
        return self._time(range(-self._ctx.nsec, 0)[section])

    def time_final_all_section(self):
        return [self.time_final(s) for s in range(self._ctx.nsec)]

    def time_knots(self):
        # reference optimize.py:533-540 ([0] + final times; the reference assumes t0 = 0 here)
        return [0] + self.time_final_all_section()

    def time_update(self):
        """reference optimize.py:518-531 as a traced vector over all nodes: per phase (t_{s+1} - t_s) / 2 * tau_s
        + (t_{s+1} + t_s) / 2 with t = [0] + final times -- final-time scalars times the per-node constants tau.
        (The reference also stores the result in prob.time; a traced callback has no side effects.)"""
        ctx = self._ctx
        t = self.time_knots()
        parts = {}
        for s in range(ctx.nsec):
            seg = (t[s + 1] - t[s]) / 2.0 * self.tau[s] + (t[s + 1] + t[s]) / 2.0
            parts[s] = seg.parts[s]
        return Sym(ctx, (0, ctx.gtot), parts)


class interp1d_tracing:
    """While callbacks are being traced, calling a `scipy.interpolate.interp1d` object with a
    traced value records a table lookup instead of evaluating (reference
    examples/11_Polar_TSTO_Taiki.py:94-98 does `obj.airDensity(R - obj.Re)`)."""

    def __enter__(self):
        from scipy.interpolate import interp1d
        self.cls = interp1d
        self.orig = interp1d.__call__
        orig = self.orig

        def call(f, x):
            if isinstance(x, Sym):
                return x._interp(f)
            if isinstance(x, SymList):
                return SymList([e._interp(f) if isinstance(e, Sym) else orig(f, e) for e in x.items])
            return orig(f, x)

        interp1d.__call__ = call
        return self

    def __exit__(self, *exc):
        self.cls.__call__ = self.orig
        return False


class SymRows:
    """Ordered pieces appended to a Condition while tracing (Sym / SymList / floats)."""

    def __init__(self, pieces):
        self.pieces = pieces


class SymDynamics:
    """Per-state right-hand sides returned by Dynamics.__call__ while tracing."""

    def __init__(self, section, rhs):
        self.section, self.rhs = section, rhs
