"""Lower traced expression DAGs (trace.py) to the register tapes of include/ogb200.h
and assemble the problem IR the C ABI consumes.

A tape is a straight-line program: one 64-bit word per instruction
(op[63:56] dst[55:42] a[41:28] b[27:14] c[13:0]).  Registers are allocated by last
use so a tape needs only as many registers as it has simultaneously live values.
Operation order inside an expression is the user's (python evaluates left to right),
so the device reproduces numpy's rounding.
"""
from dataclasses import dataclass, field

import numpy as np

from . import trace as T

OPCODES = {
    "nop": 0, "ldp": 1, "ldc": 2, "out": 3, "add": 4, "sub": 5, "mul": 6, "div": 7, "pow": 8,
    "min": 9, "max": 10, "atan2": 11, "lt": 12, "le": 13, "gt": 14, "ge": 15, "eq": 16, "ne": 17,
    "sel": 18, "neg": 19, "sqrt": 20, "exp": 21, "log": 22, "sin": 23, "cos": 24, "tan": 25,
    "abs": 26, "square": 27, "recip": 28, "asin": 29, "acos": 30, "atan": 31, "sinh": 32,
    "cosh": 33, "tanh": 34, "log10": 35, "sign": 36, "floor": 37, "ceil": 38, "and": 39,
    "or": 40, "not": 41, "interp": 42,
}
LEAF_OPS = ("blk", "var", "const", "nodec")
MAX_REG = 96
MAX_FIELD = 16383

OUT_DYN, OUT_EQ_POINT, OUT_INEQ_POINT, OUT_RUNNING, OUT_EQ_SCALAR, OUT_INEQ_SCALAR, OUT_COST = range(7)


def _word(op, d=0, a=0, b=0, c=0):
    for f in (d, a, b, c):
        if f < 0 or f > MAX_FIELD:
            raise ValueError("tape operand %d does not fit the 14-bit field" % f)
    return (op << 56) | (d << 42) | (a << 28) | (b << 14) | c


@dataclass
class Tape:
    code: np.ndarray            # uint64
    consts: np.ndarray          # float64
    outs: list                  # [(kind, row, glo, ghi)] ; slot i <-> outs[i]
    nreg: int
    # node programs only: LDP operand a addresses block a (a < nb), then per-node constant vector a - nb,
    # then global variable a - nb - len(nodec) (a final time read at every node)
    nodec: list = field(default_factory=list)       # arrays of one value per node of the phase
    globals: list = field(default_factory=list)     # decision-variable indices


def compile_tape(out_nodes, outs, leaf_arg):
    """out_nodes[i] feeds output slot i (described by outs[i]).
    leaf_arg(node) -> LDP operand for a 'blk' / 'var' leaf."""
    # ---- topological order (iterative post-order DFS), shared nodes once
    order, seen = [], set()
    for root in out_nodes:
        stack = [(root, False)]
        while stack:
            n, done = stack.pop()
            if done:
                order.append(n)
                continue
            if n.uid in seen:
                continue
            seen.add(n.uid)
            stack.append((n, True))
            for a in reversed(n.args if n.op not in LEAF_OPS else ()):
                if a.uid not in seen:
                    stack.append((a, False))
    pos = {n.uid: i for i, n in enumerate(order)}
    # ---- last use (outputs are emitted right after their node is available: see below)
    last = {n.uid: pos[n.uid] for n in order}
    for n in order:
        if n.op not in LEAF_OPS:
            for a in n.args:
                last[a.uid] = max(last[a.uid], pos[n.uid])
    out_slots = {}
    for slot, n in enumerate(out_nodes):
        out_slots.setdefault(n.uid, []).append(slot)
    # ---- emit
    code, consts, cindex = [], [], {}
    reg, free, nreg = {}, [], 0
    for i, n in enumerate(order):
        # operands whose last use is this instruction can donate their register
        srcs = [reg[a.uid] for a in n.args] if n.op not in LEAF_OPS else []
        for a in (n.args if n.op not in LEAF_OPS else ()):
            if last[a.uid] == i and a.uid in reg:
                r = reg.pop(a.uid)
                free.append(r)
        if free:
            d = free.pop()
        else:
            d = nreg
            nreg += 1
        reg[n.uid] = d
        if n.op == "const":
            key = np.float64(n.value).tobytes()
            if key not in cindex:
                cindex[key] = len(consts)
                consts.append(n.value)
            code.append(_word(OPCODES["ldc"], d, cindex[key]))
        elif n.op in ("blk", "var", "nodec"):
            code.append(_word(OPCODES["ldp"], d, leaf_arg(n)))
        elif n.op == "interp":
            code.append(_word(OPCODES["interp"], d, srcs[0], int(n.value)))
        else:
            s = srcs + [0] * (3 - len(srcs))
            code.append(_word(OPCODES[n.op], d, s[0], s[1], s[2]))
        for slot in out_slots.get(n.uid, ()):
            code.append(_word(OPCODES["out"], slot, d))
        if last[n.uid] == i:                # never read again (pure output / dead)
            free.append(reg.pop(n.uid))
    if nreg > MAX_REG:
        raise T.TraceError("expression needs %d live registers (limit %d)" % (nreg, MAX_REG))
    return Tape(np.array(code, dtype=np.uint64), np.array(consts, dtype=np.float64),
                list(outs), max(nreg, 1))


@dataclass
class ProblemIR:
    nodes: list
    nstates: list
    ncontrols: list
    unit_states: list           # flattened over phases
    unit_time: float
    t0: float
    knot_smooth: list
    meq_user: int
    mineq_user: int
    has_running_cost: bool
    node_tapes: list            # one Tape per phase
    scalar_tape: Tape
    nvars: int = 0
    tables: list = field(default_factory=list)   # dicts: x, y, variant, extrapolate, fill_below, fill_above
    meta: dict = field(default_factory=dict)
    unit_controls: list = field(default_factory=list)   # flattened over phases (guess / trajectory kernels only)


def _node_local(ctx, leaf, s, any_var=False):
    """May a node program of phase s read this leaf?  The phase's own blocks and per-node constants at the
    node, and global variables every node reads alike (their Jacobian columns are dense in the phase): the
    final times -- and, for dynamics / running cost (any_var), picked states / controls such as `m[0]`.  User
    rows that mix a vector with a picked element keep being expanded into scalar rows instead (sparser)."""
    if leaf[0] in ("blk", "nodec"):
        return leaf[1] == s
    return leaf[0] == "var" and (any_var or ctx.is_time_var(leaf[1]))


class _RowBuilder:
    """Turns the ordered pieces of a traced Condition into output slots."""

    def __init__(self, ctx, point_kind, scalar_kind):
        self.ctx, self.point_kind, self.scalar_kind = ctx, point_kind, scalar_kind
        self.nrows = 0
        self.point = {s: [] for s in range(ctx.nsec)}   # per phase: (node, (kind,row,glo,ghi))
        self.scalar = []                                # (node, (kind,row,0,0))

    def _scalar(self, node):
        self.scalar.append((node, (self.scalar_kind, self.nrows, 0, 0)))
        self.nrows += 1

    def add(self, piece):
        g = self.ctx.graph
        if isinstance(piece, T.Sym):
            if piece.rng is None:
                self._scalar(piece.parts)
                return
            local = all(all(_node_local(self.ctx, l, s) for l in T.leaves(p))
                        for s, p in piece.parts.items())
            if local:
                glo, ghi = piece.rng
                if ghi > glo:
                    for s, p in piece.parts.items():
                        self.point[s].append((p, (self.point_kind, self.nrows, glo, ghi)))
                    self.nrows += ghi - glo
                return
            piece = T.SymList.from_any(piece)
        if isinstance(piece, T.SymList):
            for e in piece.items:
                self.add(e if isinstance(e, T.Sym) else float(e))
            return
        arr = np.atleast_1d(np.asarray(piece, dtype=float)).ravel()
        for v in arr:
            self._scalar(g.const(float(v)))


def build_ir(prob, obj):
    """Trace every callback of `prob` once and return the ProblemIR (see _build_ir)."""
    with T.interp1d_tracing():
        return _build_ir(prob, obj)


def _build_ir(prob, obj):
    """Trace every callback of `prob` once and return the ProblemIR.
    Row order = the reference's: user equality rows, defects per phase/state/node, knot
    rows, user inequality rows, cost (optimize.py:670-698, :723-728)."""
    from .optimize import Condition  # noqa: F401  (facade classes cooperate with the tracer)
    ctx = T.TraceContext(prob)
    view = T.TraceView(prob, ctx)
    g = ctx.graph
    nsec = ctx.nsec

    # ---- dynamics: one call per phase (optimize.py:685)
    dyn_nodes = []
    for s in range(nsec):
        fn = prob.dynamics[s]
        if fn is None:
            raise AssertionError("It must be set dynamics")
        res = fn(view, obj, s)
        rhs = []
        if isinstance(res, T.SymDynamics):
            items = res.rhs
        else:
            arr = np.asarray(res, dtype=float).reshape(ctx.nstates[s], ctx.nodes[s])
            items = [row for row in arr]
        if len(items) != ctx.nstates[s]:
            raise T.TraceError("dynamics of phase %d returned %d states, expected %d"
                               % (s, len(items), ctx.nstates[s]))
        for a, it in enumerate(items):
            if isinstance(it, T.SymList):
                raise T.TraceError("dynamics[%d] of phase %d mixes nodes (a reversed / shifted vector, or vectors of "
                                   "different node ranges): on the device a dynamics function is evaluated node by node" % (a, s))
            if isinstance(it, T.Sym):
                if it.rng is None:
                    node = it.parts
                elif it.rng == (ctx.g0[s], ctx.g0[s] + ctx.nodes[s]) and s in it.parts:
                    node = it.parts[s]
                else:
                    raise T.TraceError("dynamics[%d] of phase %d is not a vector over that phase's nodes" % (a, s))
                for lf in T.leaves(node):
                    if not _node_local(ctx, lf, s, any_var=True):
                        raise T.TraceError("dynamics of phase %d reads %r: on the device a dynamics function may read "
                                           "that phase's states / controls at the same node, per-node constants "
                                           "(prob.time[s], prob.tau[s], tables of one value per node), the final "
                                           "times and picked elements (x[0], m[-1]) -- not whole vectors of other "
                                           "nodes or phases" % (s, lf))
            else:
                arr = np.atleast_1d(np.asarray(it, dtype=float))
                if np.all(arr == arr.flat[0]):
                    node = g.const(float(arr.flat[0]))
                elif arr.shape == (ctx.nodes[s],):          # a per-node data table as the right-hand side
                    node = g.nodec(s, ctx.node_const(s, arr))
                else:
                    raise T.TraceError("dynamics[%d] of phase %d is a constant array that is not one value per node" % (a, s))
            rhs.append(node)
        dyn_nodes.append(rhs)

    # ---- user rows
    eq = _RowBuilder(ctx, OUT_EQ_POINT, OUT_EQ_SCALAR)
    res = prob.equality(view, obj)
    for piece in (res.pieces if isinstance(res, T.SymRows) else [res]):
        eq.add(piece)
    ineq = _RowBuilder(ctx, OUT_INEQ_POINT, OUT_INEQ_SCALAR)
    res = prob.inequality(view, obj)
    for piece in (res.pieces if isinstance(res, T.SymRows) else [res]):
        ineq.add(piece)

    # ---- cost (optimize.py:700-709)
    cres = prob.cost(view, obj)
    if isinstance(cres, T.Sym):
        if cres.rng is not None:
            raise T.TraceError("cost() must return a scalar")
        cost_node = cres.parts
    else:
        cost_node = g.const(float(cres))
    run_parts = None
    if prob.running_cost is not None:
        rres = prob.running_cost(view, obj)
        if not (isinstance(rres, T.Sym) and rres.rng == (0, ctx.gtot)):
            raise T.TraceError("running_cost() must return one value per node of every phase")
        for s, p in rres.parts.items():
            for lf in T.leaves(p):
                if not _node_local(ctx, lf, s, any_var=True):
                    raise T.TraceError("running_cost() must be pointwise in the node")
        run_parts = rres.parts

    # ---- tapes
    node_tapes = []
    for s in range(nsec):
        nodes_out = list(dyn_nodes[s])
        outs = [(OUT_DYN, a, 0, 0) for a in range(ctx.nstates[s])]
        for rb in (eq, ineq):
            for node, desc in rb.point[s]:
                nodes_out.append(node)
                outs.append(desc)
        if run_parts is not None:
            nodes_out.append(run_parts[s])
            outs.append((OUT_RUNNING, 0, 0, 0))
        nb, nnc = ctx.nstates[s] + ctx.ncontrols[s], len(ctx.node_consts[s])
        gvars = sorted({lf[1] for nd in nodes_out for lf in T.leaves(nd) if lf[0] == "var"})

        def leaf_arg(n, nb=nb, nnc=nnc, gvars=gvars):
            if n.op == "blk":
                return n.args[1]
            if n.op == "nodec":
                return nb + n.args[1]
            return nb + nnc + gvars.index(n.args[0])
        tp = compile_tape(nodes_out, outs, leaf_arg)
        tp.nodec = [np.asarray(v, dtype=np.float64) for v in ctx.node_consts[s]]
        tp.globals = list(gvars)
        node_tapes.append(tp)
    sc_nodes = [n for n, _ in eq.scalar] + [n for n, _ in ineq.scalar] + [cost_node]
    sc_outs = [d for _, d in eq.scalar] + [d for _, d in ineq.scalar] + [(OUT_COST, 0, 0, 0)]
    for n in sc_nodes:
        for lf in T.leaves(n):
            if lf[0] != "var":
                raise T.TraceError("internal: vector leaf in a scalar row")
    scalar_tape = compile_tape(sc_nodes, sc_outs, lambda n: n.args[0])

    units = [float(u) for s in range(nsec) for u in prob.unit_states[s][:ctx.nstates[s]]]
    smooth = list(prob.knot_states_smooth) + [True] * nsec
    return ProblemIR(
        nodes=list(ctx.nodes), nstates=list(ctx.nstates), ncontrols=list(ctx.ncontrols),
        unit_states=units, unit_time=float(prob.unit_time), t0=float(prob.t0),
        knot_smooth=[bool(smooth[k]) for k in range(nsec - 1)],
        meq_user=eq.nrows, mineq_user=ineq.nrows, has_running_cost=run_parts is not None,
        node_tapes=node_tapes, scalar_tape=scalar_tape, nvars=ctx.nvars,
        tables=[{k: v for k, v in t.items() if k != "keep"} for t in ctx.tables],
        meta={"graph_nodes": len(g.nodes)},
        unit_controls=[float(u) for s in range(nsec) for u in prob.unit_controls[s][:ctx.ncontrols[s]]])


# ------------------------------------------------------------------ (de)serialisation
def ir_to_arrays(ir):
    """ProblemIR -> dict of numpy arrays (np.savez-able): a traced problem can be shipped without the
    Python callbacks it came from (tests/golden/example_XX_ir.npz hold the traces of the reference's
    shipped example scripts; engine.DeviceProblem(ir_from_arrays(...), bounds) rebuilds the engine)."""
    out = {"layout": np.array([ir.nodes, ir.nstates, ir.ncontrols], dtype=np.int64),
           "unit_states": np.array(ir.unit_states, dtype=np.float64),
           "scalars": np.array([ir.unit_time, ir.t0], dtype=np.float64),
           "flags": np.array([ir.meq_user, ir.mineq_user, int(ir.has_running_cost), ir.nvars, len(ir.tables)],
                             dtype=np.int64),
           "knot_smooth": np.array([1 if k else 0 for k in ir.knot_smooth], dtype=np.uint8),
           "unit_controls": np.array(ir.unit_controls, dtype=np.float64)}
    for name, tp in [("node%d" % s, t) for s, t in enumerate(ir.node_tapes)] + [("scalar", ir.scalar_tape)]:
        out[name + "_code"] = np.asarray(tp.code, dtype=np.uint64)
        out[name + "_consts"] = np.asarray(tp.consts, dtype=np.float64)
        out[name + "_outs"] = np.array(tp.outs, dtype=np.int64).reshape(-1, 4)
        out[name + "_nreg"] = np.array([tp.nreg], dtype=np.int64)
        out[name + "_nodec"] = (np.array(tp.nodec, dtype=np.float64).reshape(len(tp.nodec), -1) if len(tp.nodec)
                                else np.zeros((0, 0)))
        out[name + "_globals"] = np.array(tp.globals, dtype=np.int64)
    for i, t in enumerate(ir.tables):
        out["table%d_x" % i] = np.asarray(t["x"], dtype=np.float64)
        out["table%d_y" % i] = np.asarray(t["y"], dtype=np.float64)
        out["table%d_meta" % i] = np.array([t["variant"], 1.0 if t["extrapolate"] else 0.0, t["fill_below"],
                                            t["fill_above"]], dtype=np.float64)
    return out


def ir_from_arrays(d):
    """Inverse of ir_to_arrays (d: a dict or an open .npz)."""
    def tp(name):
        t = Tape(np.asarray(d[name + "_code"], dtype=np.uint64), np.asarray(d[name + "_consts"], dtype=np.float64),
                 [tuple(int(v) for v in row) for row in np.asarray(d[name + "_outs"]).reshape(-1, 4)],
                 int(d[name + "_nreg"][0]))
        if name + "_nodec" in d:
            t.nodec = [np.asarray(v, dtype=np.float64) for v in np.asarray(d[name + "_nodec"])]
            t.globals = [int(v) for v in d[name + "_globals"]]
        return t
    layout = np.asarray(d["layout"])
    flags = [int(v) for v in d["flags"]]
    tables = []
    for i in range(flags[4]):
        meta = np.asarray(d["table%d_meta" % i])
        tables.append(dict(x=np.asarray(d["table%d_x" % i]), y=np.asarray(d["table%d_y" % i]), variant=int(meta[0]),
                           extrapolate=bool(meta[1]), fill_below=float(meta[2]), fill_above=float(meta[3])))
    nsec = layout.shape[1]
    return ProblemIR(nodes=[int(v) for v in layout[0]], nstates=[int(v) for v in layout[1]],
                     ncontrols=[int(v) for v in layout[2]], unit_states=[float(v) for v in d["unit_states"]],
                     unit_time=float(d["scalars"][0]), t0=float(d["scalars"][1]),
                     knot_smooth=[bool(v) for v in d["knot_smooth"]], meq_user=flags[0], mineq_user=flags[1],
                     has_running_cost=bool(flags[2]), node_tapes=[tp("node%d" % s) for s in range(nsec)],
                     scalar_tape=tp("scalar"), nvars=flags[3], tables=tables, meta={"loaded": True},
                     unit_controls=[float(v) for v in d["unit_controls"]] if "unit_controls" in d else [])
