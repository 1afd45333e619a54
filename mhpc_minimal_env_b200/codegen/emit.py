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
