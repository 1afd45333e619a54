"""Parser for CasADi SX straight-line C (the format of the reference's
CasadiGen/source/*.c, described in SURVEY.md Appendix A.4) into an SSA DAG.

The emitted C has one function body per file with exactly these statement forms
(reference example lines: CasadiGen/source/Dyn_BS.c:55-80, FBDynamics_par.c:57-116):

    aN=arg[i]? arg[i][j] : 0;          load of input i, element j
    aN=<literal>;                      constant (MSVC style exponents, "1.", "0.")
    aN=(aP+aQ); (aP-aQ) (aP*aQ) (aP/aQ)
    aN=(-aP);  aN=sin(aP);  cos  sqrt  aN=casadi_sq(aP);
    if (res[k]!=0) res[k][j]=aN;       store of non-zero j of output k

Work variables are re-assigned (CasADi register allocation), so the parser
SSA-renames on the fly.  Output sparsity comes from the `casadi_sK` CCS arrays
({nrow, ncol, colptr[ncol+1], row[nnz]}) selected by `<name>_sparsity_out`.

Nothing here is copied into the repository: the parser runs at build time on the
reference tree (see codegen/gen_all.py) and the emitted CUDA is a build artefact.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# op kinds
IN, CONST, ADD, SUB, MUL, DIV, NEG, SQ, SIN, COS, SQRT = (
    "in", "const", "add", "sub", "mul", "div", "neg", "sq", "sin", "cos", "sqrt")
BINOPS = {"+": ADD, "-": SUB, "*": MUL, "/": DIV}
UNOPS = {"sin": SIN, "cos": COS, "sqrt": SQRT, "casadi_sq": SQ}

_re_load = re.compile(r"^a(\d+)=arg\[(\d+)\]\? arg\[(\d+)\]\[(\d+)\] : 0;$")
_re_const = re.compile(r"^a(\d+)=(-?[0-9][0-9.eE+\-]*);$")
_re_bin = re.compile(r"^a(\d+)=\(a(\d+)([+\-*/])a(\d+)\);$")
_re_neg = re.compile(r"^a(\d+)=\(-a(\d+)\);$")
_re_un = re.compile(r"^a(\d+)=(sin|cos|sqrt|casadi_sq)\(a(\d+)\);$")
_re_store = re.compile(r"^if \(res\[(\d+)\]!=0\) res\[(\d+)\]\[(\d+)\]=a(\d+);$")
_re_sp = re.compile(r"^static const casadi_int casadi_s(\d+)\[(\d+)\] = \{([^}]*)\};$")
_re_sig = re.compile(r"^/\* (\w+):\((.*)\)->\((.*)\) \*/$")


@dataclass
class Node:
    op: str
    a: int = -1          # first operand (node id)
    b: int = -1          # second operand (node id)
    val: float = 0.0     # CONST value
    lit: str = ""        # CONST literal as written (normalised)
    arg: int = -1        # IN: input index
    idx: int = -1        # IN: element index


@dataclass
class Store:
    out: int             # output index k
    nz: int              # non-zero index j within output k
    node: int            # node id of the stored value
    pos: int             # number of nodes defined before this store (program position)
    row: int = -1        # dense row (filled from sparsity)
    col: int = -1        # dense col


@dataclass
class Sparsity:
    nrow: int
    ncol: int
    colptr: List[int]
    rows: List[int]

    @property
    def nnz(self) -> int:
        return self.colptr[-1]

    def coords(self) -> List[Tuple[int, int]]:
        out = []
        for c in range(self.ncol):
            for k in range(self.colptr[c], self.colptr[c + 1]):
                out.append((self.rows[k], c))
        return out


@dataclass
class SXFunction:
    name: str
    n_in: int
    n_out: int
    in_sizes: List[int]
    nodes: List[Node] = field(default_factory=list)
    stores: List[Store] = field(default_factory=list)
    out_sp: List[Sparsity] = field(default_factory=list)

    # ---- statistics -------------------------------------------------------
    def op_census(self) -> Dict[str, int]:
        c: Dict[str, int] = {}
        for n in self.nodes:
            c[n.op] = c.get(n.op, 0) + 1
        return c

    def flops(self) -> int:
        """`+ - * / neg sq` count (SURVEY.md section 2.4 convention)."""
        c = self.op_census()
        return sum(c.get(k, 0) for k in (ADD, SUB, MUL, DIV, NEG, SQ))


def _norm_literal(s: str) -> Tuple[float, str]:
    # MSVC style 3-digit exponents: 2.5e-002 ; trailing-dot forms "1." "0."
    v = float(s)
    return v, repr(v)


def _parse_sparsity(vals: List[int]) -> Sparsity:
    nrow, ncol = vals[0], vals[1]
    colptr = vals[2:2 + ncol + 1]
    nnz = colptr[-1]
    rows = vals[2 + ncol + 1:2 + ncol + 1 + nnz]
    if len(vals) == 2 + ncol + 1 + nnz:
        pass
    elif len(vals) == 3 and nrow * ncol == 1:      # dense scalar short form
        colptr, rows = [0, 1], [0]
    else:
        raise ValueError("unexpected sparsity array length")
    return Sparsity(nrow, ncol, list(colptr), list(rows))


def parse_casadi_c(path: str, cse: bool = True) -> SXFunction:
    """Parse one CasADi-generated C file into an SXFunction (SSA, optional CSE)."""
    with open(path, "r") as fh:
        lines = [ln.rstrip("\n") for ln in fh]

    sp_arrays: Dict[int, Sparsity] = {}
    name = None
    sig_in: List[int] = []
    body_start = None
    for i, ln in enumerate(lines):
        m = _re_sp.match(ln.strip())
        if m:
            vals = [int(t) for t in m.group(3).split(",")]
            assert len(vals) == int(m.group(2))
            sp_arrays[int(m.group(1))] = _parse_sparsity(vals)
            continue
        m = _re_sig.match(ln.strip())
        if m and name is None:
            name = m.group(1)
            for tok in m.group(2).split(","):
                mm = re.match(r"i\d+(?:\[(\d+)\])?$", tok.strip())
                sig_in.append(int(mm.group(1)) if mm and mm.group(1) else 1)
            continue
        if ln.startswith("static int casadi_f0("):
            body_start = i + 1
            break
    assert name is not None and body_start is not None, path

    # map outputs to sparsity arrays through <name>_sparsity_out
    out_sp_idx: List[int] = []
    in_fn = False
    for ln in lines:
        if f"{name}_sparsity_out(casadi_int i)" in ln:
            in_fn = True
            continue
        if in_fn:
            m = re.match(r"\s*case (\d+): return casadi_s(\d+);", ln)
            if m:
                assert int(m.group(1)) == len(out_sp_idx)
                out_sp_idx.append(int(m.group(2)))
            elif "default" in ln:
                break
    out_sp = [sp_arrays[k] for k in out_sp_idx]

    fn = SXFunction(name=name, n_in=len(sig_in), n_out=len(out_sp), in_sizes=sig_in, out_sp=out_sp)
    nodes = fn.nodes
    cur: Dict[int, int] = {}                 # work variable -> node id
    memo: Dict[tuple, int] = {}

    def mk(key: tuple, node: Node) -> int:
        if cse and key in memo:
            return memo[key]
        nodes.append(node)
        nid = len(nodes) - 1
        if cse:
            memo[key] = nid
        return nid

    i = body_start
    while i < len(lines):
        s = lines[i].strip()
        i += 1
        if not s:
            continue
        if s == "return 0;":
            break
        if s.startswith("casadi_real "):
            continue
        m = _re_bin.match(s)
        if m:
            d, p, o, q = int(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4))
            op = BINOPS[o]
            a, b = cur[p], cur[q]
            cur[d] = mk((op, a, b), Node(op, a, b))
            continue
        m = _re_load.match(s)
        if m:
            d, ai, ai2, j = (int(m.group(k)) for k in (1, 2, 3, 4))
            assert ai == ai2
            cur[d] = mk((IN, ai, j), Node(IN, arg=ai, idx=j))
            continue
        m = _re_store.match(s)
        if m:
            k, k2, j, p = (int(m.group(t)) for t in (1, 2, 3, 4))
            assert k == k2
            fn.stores.append(Store(k, j, cur[p], len(nodes)))
            continue
        m = _re_un.match(s)
        if m:
            d, f, p = int(m.group(1)), m.group(2), int(m.group(3))
            op = UNOPS[f]
            cur[d] = mk((op, cur[p]), Node(op, cur[p]))
            continue
        m = _re_neg.match(s)
        if m:
            d, p = int(m.group(1)), int(m.group(2))
            cur[d] = mk((NEG, cur[p]), Node(NEG, cur[p]))
            continue
        m = _re_const.match(s)
        if m:
            d = int(m.group(1))
            v, lit = _norm_literal(m.group(2))
            cur[d] = mk((CONST, lit), Node(CONST, val=v, lit=lit))
            continue
        raise ValueError(f"{path}:{i}: statement outside the closed SX form set: {s!r}")

    # dense coordinates of every store
    coords = [sp.coords() for sp in out_sp]
    seen = [set() for _ in out_sp]
    for st in fn.stores:
        st.row, st.col = coords[st.out][st.nz]
        seen[st.out].add(st.nz)
    for k, sp in enumerate(out_sp):
        assert len(seen[k]) == sp.nnz, f"{name}: output {k} has unwritten non-zeros"
    return fn


def evaluate(fn: SXFunction, args: List[List[float]]) -> List[List[float]]:
    """Pure-Python evaluation in definition order (tests, small cases only)."""
    import math
    v = [0.0] * len(fn.nodes)
    for i, n in enumerate(fn.nodes):
        if n.op == IN:
            v[i] = args[n.arg][n.idx]
        elif n.op == CONST:
            v[i] = n.val
        elif n.op == ADD:
            v[i] = v[n.a] + v[n.b]
        elif n.op == SUB:
            v[i] = v[n.a] - v[n.b]
        elif n.op == MUL:
            v[i] = v[n.a] * v[n.b]
        elif n.op == DIV:
            v[i] = v[n.a] / v[n.b]
        elif n.op == NEG:
            v[i] = -v[n.a]
        elif n.op == SQ:
            v[i] = v[n.a] * v[n.a]
        elif n.op == SIN:
            v[i] = math.sin(v[n.a])
        elif n.op == COS:
            v[i] = math.cos(v[n.a])
        elif n.op == SQRT:
            v[i] = math.sqrt(v[n.a])
    res = [[0.0] * sp.nnz for sp in fn.out_sp]
    for st in fn.stores:
        res[st.out][st.nz] = v[st.node]
    return res


def share_reciprocals(fn: SXFunction, min_uses: int = 2) -> SXFunction:
    """Return a copy of `fn` in which every quotient a/d whose denominator d is shared by at least `min_uses` divisions
    becomes a * (1/d), with ONE reciprocal per distinct denominator placed right after d.  The generated Jacobians divide
    by the same handful of values hundreds of times (Dyn_BS_par: 853 divisions, 18 distinct denominators - the pivots of
    the mass-matrix factorisation), and an FP64 division costs ~14 instructions plus a branch on the GPU.
    Each rewritten quotient differs from the IEEE quotient by at most 1 ulp; the exact form is kept for the CPU
    bit-exactness check of the generator."""
    import collections
    cnt = collections.Counter(nd.b for nd in fn.nodes if nd.op == DIV)
    shared = {b for b, c in cnt.items() if c >= min_uses}
    new_nodes: List[Node] = []
    remap: Dict[int, int] = {}
    rcp_of: Dict[int, int] = {}
    one_id: Optional[int] = None
    for i, nd in enumerate(fn.nodes):
        a = remap[nd.a] if nd.a >= 0 else -1
        b = remap[nd.b] if nd.b >= 0 else -1
        if nd.op == DIV and nd.b in shared:
            new_nodes.append(Node(MUL, a, rcp_of[nd.b]))
        else:
            new_nodes.append(Node(nd.op, a, b, nd.val, nd.lit, nd.arg, nd.idx))
        remap[i] = len(new_nodes) - 1
        if i in shared:
            if one_id is None:
                new_nodes.append(Node(CONST, val=1.0, lit="1.0"))
                one_id = len(new_nodes) - 1
            new_nodes.append(Node(DIV, one_id, remap[i]))
            rcp_of[i] = len(new_nodes) - 1
    out = SXFunction(name=fn.name, n_in=fn.n_in, n_out=fn.n_out, in_sizes=fn.in_sizes, nodes=new_nodes, out_sp=fn.out_sp)
    # store positions are only used for ordering hints; keep them monotone
    out.stores = [Store(st.out, st.nz, remap[st.node], remap[st.node] + 1, st.row, st.col) for st in fn.stores]
    return out
