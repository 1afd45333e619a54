"""Emitters: SSA DAG (sxdag.SXFunction) -> straight-line bodies.

Two back ends share one statement generator:

* ``emit_inc``  - a macro-parameterised body (``*.inc``) that the CUDA kernels
  ``#include`` inside a ``__device__`` function after defining

      SX_IN(i, j)                     value of element j of input i
      SX_OUT(k, nz, row, col, v)      non-zero `nz` of output k == dense (row, col)

  (row, col) is the CCS -> dense scatter of the reference adaptor
  (CasadiGen/source/CasadiGen.cpp:46-68) resolved at generation time: all four
  integers are literals, so the includer can route every store at compile time.

* ``emit_c``    - a plain C function with the CasADi calling convention
  ``(const double** arg, double** res)`` writing non-zeros only; used by the CPU
  tests to check the parser/emitter bit-for-bit against the reference objects.

``schedule`` optionally re-orders nodes (any topological order is valid).
"""
from __future__ import annotations

from typing import Iterable, List, Optional

from .sxdag import (ADD, CONST, COS, DIV, IN, MUL, NEG, SIN, SQ, SQRT, SUB,
                    SXFunction)

_BIN = {ADD: "+", SUB: "-", MUL: "*", DIV: "/"}


def _statements(fn: SXFunction, order: Optional[List[int]] = None,
                pair_sincos: bool = True) -> List[str]:
    nodes = fn.nodes
    n = len(nodes)
    if order is None:
        order = list(range(n))
    pos_of = {nid: p for p, nid in enumerate(order)}

    # stores are emitted right after the node they read is available, but never
    # earlier than in the original program order relative to other stores of the
    # same output (keeps the listing readable; order of stores is irrelevant).
    stores_after = {}
    for st in fn.stores:
        stores_after.setdefault(st.node, []).append(st)

    # sin/cos pairs on the same argument -> one sincos()
    partner = {}
    if pair_sincos:
        by_arg = {}
        for nid, nd in enumerate(nodes):
            if nd.op in (SIN, COS):
                by_arg.setdefault(nd.a, {})[nd.op] = nid
        for a, d in by_arg.items():
            if SIN in d and COS in d:
                partner[d[SIN]] = d[COS]
                partner[d[COS]] = d[SIN]

    out: List[str] = []
    done = set()
    for nid in order:
        nd = nodes[nid]
        if nid in done:
            continue
        v = f"v{nid}"
        if nd.op == IN:
            out.append(f"const double {v} = SX_IN({nd.arg}, {nd.idx});")
        elif nd.op == CONST:
            out.append(f"const double {v} = {nd.lit};")
        elif nd.op == DIV and nodes[nd.a].op == CONST and nodes[nd.a].val == 1.0 and nodes[nd.b].op != CONST:
            out.append(f"const double {v} = SX_RCP(v{nd.b});")      # reciprocal of a pivot: the includer picks the sequence
        elif nd.op in _BIN:
            out.append(f"const double {v} = v{nd.a} {_BIN[nd.op]} v{nd.b};")
        elif nd.op == NEG:
            out.append(f"const double {v} = -v{nd.a};")
        elif nd.op == SQ:
            out.append(f"const double {v} = v{nd.a} * v{nd.a};")
        elif nd.op == SQRT:
            out.append(f"const double {v} = sqrt(v{nd.a});")
        elif nd.op in (SIN, COS):
            if nid in partner:
                other = partner[nid]
                s_id, c_id = (nid, other) if nd.op == SIN else (other, nid)
                out.append(f"double v{s_id}, v{c_id}; SX_SINCOS(v{nd.a}, v{s_id}, v{c_id});")
                done.add(other)
            else:
                f = "sin" if nd.op == SIN else "cos"
                out.append(f"const double {v} = {f}(v{nd.a});")
        else:
            raise AssertionError(nd.op)
        done.add(nid)
        for x in ([nid, partner[nid]] if nid in partner else [nid]):
            for st in stores_after.get(x, []):
                out.append(f"SX_OUT({st.out}, {st.nz}, {st.row}, {st.col}, v{x});")
    return out


def critical_path_order(fn: SXFunction, issue: int = 2, lat_fp: int = 8, lat_rcp: int = 60, lat_trig: int = 150) -> List[int]:
    """Topological order by list scheduling with longest-path-to-sink priority (cycles: a warp-wide FP64 instruction
    issues every `issue` cycles per scheduler, dependent results come `lat_*` cycles later; scripts/micro/fp64_lat.cu).
    The DAG and every operation stay the same - only the order of the straight-line statements changes, so that the
    serial spine of the function (factorisation pivots, substitutions) is started as early as its operands exist
    instead of where the generator happened to create it.  A rollout thread runs at ~2 warps per scheduler, where the
    length of that spine is the step time."""
    import heapq
    nodes = fn.nodes
    n = len(nodes)

    def lat(i):
        op = nodes[i].op
        if op in (IN, CONST):
            return 0
        if op in (SIN, COS, SQRT):
            return lat_trig
        if op == DIV:
            return lat_rcp
        return lat_fp
    users: List[List[int]] = [[] for _ in range(n)]
    npred = [0] * n
    for i, nd in enumerate(nodes):
        for o in {nd.a, nd.b}:
            if o >= 0:
                users[o].append(i)
                npred[i] += 1
    prio = [0] * n
    for i in reversed(range(n)):
        best = 0
        for u in users[i]:
            if prio[u] > best:
                best = prio[u]
        prio[i] = best + lat(i)
    ready_at = [0] * n
    heap = [(-prio[i], i) for i in range(n) if npred[i] == 0]
    heapq.heapify(heap)
    waiting: List = []          # (ready cycle, -prio, node): operands issued, result not yet available
    order: List[int] = []
    clock = 0
    while heap or waiting:
        while waiting and waiting[0][0] <= clock:
            _, pr, i = heapq.heappop(waiting)
            heapq.heappush(heap, (pr, i))
        if not heap:
            clock = waiting[0][0]
            continue
        _, i = heapq.heappop(heap)
        order.append(i)
        done = clock + lat(i)
        if nodes[i].op not in (IN, CONST):
            clock += issue
        for u in users[i]:
            npred[u] -= 1
            if done > ready_at[u]:
                ready_at[u] = done
            if npred[u] == 0:
                heapq.heappush(waiting, (ready_at[u], -prio[u], u))
    assert len(order) == n
    return order


def reordered(fn: SXFunction, order: List[int]) -> SXFunction:
    """The same function with its nodes renumbered in `order` (a topological order)."""
    from .sxdag import Node, Store
    new_id = {old: k for k, old in enumerate(order)}
    nodes = []
    for old in order:
        nd = fn.nodes[old]
        nodes.append(Node(nd.op, new_id[nd.a] if nd.a >= 0 else -1, new_id[nd.b] if nd.b >= 0 else -1, nd.val, nd.lit, nd.arg, nd.idx))
    out = SXFunction(name=fn.name, n_in=fn.n_in, n_out=fn.n_out, in_sizes=fn.in_sizes, nodes=nodes, out_sp=fn.out_sp)
    out.stores = [Store(s.out, s.nz, new_id[s.node], new_id[s.node] + 1, s.row, s.col) for s in fn.stores]
    return out


def emit_inc(fn: SXFunction, order: Optional[List[int]] = None) -> str:
    hdr = [
        f"// GENERATED at build time by mhpc_minimal_env_b200/codegen from the CasADi SX",
        f"// function `{fn.name}` (SSA form, common sub-expressions merged). Do not edit.",
        f"// inputs: {fn.in_sizes}; outputs: " +
        ", ".join(f"o{k}[{sp.nrow}x{sp.ncol},{sp.nnz}nz]" for k, sp in enumerate(fn.out_sp)),
        f"// nodes: {len(fn.nodes)}  flops(+-*/ neg sq): {fn.flops()}",
    ]
    return "\n".join(hdr + _statements(fn, order)) + "\n"


def emit_c(fn: SXFunction, symbol: str, order: Optional[List[int]] = None) -> str:
    body = _statements(fn, order, pair_sincos=False)
    return "\n".join([
        "#include <math.h>",
        "#define SX_IN(i, j) (arg[i][j])",
        "#define SX_OUT(k, nz, row, col, v) res[k][nz] = (v)",
        "#define SX_RCP(x) (1.0 / (x))",
        f"void {symbol}(const double** arg, double** res) {{",
        *("  " + s for s in body),
        "}",
        "#undef SX_IN",
        "#undef SX_OUT",
        "#undef SX_RCP",
        "",
    ])
