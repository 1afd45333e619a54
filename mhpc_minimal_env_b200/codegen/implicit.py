"""Structured re-derivation of the whole-body dynamics (`Dyn_{BS,FS,FL}`) and their Jacobians (`Dyn_{BS,FS,FL}_par`).

The reference's `Dyn_*_par.c` are CasADi's derivative of `Dyn_*.c` (o0 = xdot, o1 = contact forces), and `Dyn_*.c`
solves the 9x9 (stance) / 7x7 (flight) constrained-dynamics system  K(q) z = b(q, qd, u)  with a *symbolic*
Gram-Schmidt QR (one sqrt per column, Dyn_BS.c has 9, Dyn_FL.c has 7).  CasADi differentiates straight through that
factorisation, which is why the Jacobian DAGs have 12-26 k nodes and ~900 simultaneously live values - far more than a
GPU thread can hold.  The same Jacobian follows from the implicit-function theorem with a fraction of the work and of
the live set:

        K z = b      =>      K dz = db - dK z       (one factorisation of K, one pair of triangular solves per column)

This module rebuilds `Dyn_*_par` that way, mechanically, from the primal DAG alone:

 1. nodes of the primal without a sqrt ancestor are "pre" nodes (kinematics, mass matrix, bias forces), the rest is the
    QR solve;
 2. the right-hand-side entries `beta` are the pre nodes the outputs depend on affinely, and  K' = (d out / d beta)^-1
    is recovered numerically at random inputs and matched, entry by entry, to (signed) pre nodes - so K' is known as
    a matrix of DAG nodes without any knowledge of how the model was written;
 3. a row permutation / sign change that makes K' the symmetric quasi-definite matrix [M -J^T; -J 0] is found
    numerically, so K' can be factorised as L D L^T without pivoting (positive pivots from the mass matrix, then the
    negative-definite Schur complement) - 45 stored values and 81 multiply-adds per solve instead of 126 / 126 for QR;
 4. forward-mode tangents of the pre nodes (with zero pruning and hash-consing) and one solve per input direction give
    d out / d in for exactly the non-zeros of the reference's output sparsity; the same factorisation gives the primal
    outputs for the rollouts (`structured_primal`).

The result is a different (mathematically equal) floating-point evaluation of the same Jacobian: it is checked against
the reference's `Dyn_*_par` objects to 1e-10 of the matrix norm at random inputs (tests/test_cpu.py), within the
1e-9 norm-relative Jacobian tolerance of SURVEY.md section 7.  The impact maps (`Imp_{F,B}`, `Imp_{F,B}_par`) have
structural zeros in their right-hand side; their system matrix is taken from the stance dynamics of the same foot
(identify_system(donor=...)).
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Tuple

from .sxdag import (ADD, CONST, COS, DIV, IN, MUL, NEG, SIN, SQ, SQRT, SUB, Node, Store, SXFunction)


class Builder:
    """Hash-consing DAG builder with constant folding; `None` stands for a structural zero."""

    def __init__(self) -> None:
        self.nodes: List[Node] = []
        self.memo: Dict[tuple, int] = {}

    def _mk(self, key: tuple, node: Node) -> int:
        i = self.memo.get(key)
        if i is None:
            self.nodes.append(node)
            i = len(self.nodes) - 1
            self.memo[key] = i
        return i

    def const(self, v: float) -> Optional[int]:
        if v == 0.0:
            return None
        return self._mk(("c", repr(float(v))), Node(CONST, val=float(v), lit=repr(float(v))))

    def cval(self, a: Optional[int]) -> Optional[float]:
        if a is None:
            return 0.0
        n = self.nodes[a]
        return n.val if n.op == CONST else None

    def inp(self, arg: int, idx: int) -> int:
        return self._mk(("i", arg, idx), Node(IN, arg=arg, idx=idx))

    def neg(self, a: Optional[int]) -> Optional[int]:
        if a is None:
            return None
        n = self.nodes[a]
        if n.op == NEG:
            return n.a
        if n.op == CONST:
            return self.const(-n.val)
        return self._mk(("n", a), Node(NEG, a))

    def add(self, a: Optional[int], b: Optional[int]) -> Optional[int]:
        if a is None:
            return b
        if b is None:
            return a
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca + cb)
        if self.nodes[b].op == NEG:
            return self.sub(a, self.nodes[b].a)
        if self.nodes[a].op == NEG:
            return self.sub(b, self.nodes[a].a)
        lo, hi = (a, b) if a <= b else (b, a)
        return self._mk(("+", lo, hi), Node(ADD, lo, hi))

    def sub(self, a: Optional[int], b: Optional[int]) -> Optional[int]:
        if b is None:
            return a
        if a is None:
            return self.neg(b)
        if a == b:
            return None
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca - cb)
        if self.nodes[b].op == NEG:
            return self.add(a, self.nodes[b].a)
        return self._mk(("-", a, b), Node(SUB, a, b))

    def mul(self, a: Optional[int], b: Optional[int]) -> Optional[int]:
        if a is None or b is None:
            return None
        ca, cb = self.cval(a), self.cval(b)
        if ca is not None and cb is not None:
            return self.const(ca * cb)
        if ca is not None:
            a, b, ca, cb = b, a, cb, ca
        if cb == 1.0:
            return a
        if cb == -1.0:
            return self.neg(a)
        # pull negations out so that x*y and (-x)*y share the product
        if self.nodes[a].op == NEG:
            return self.neg(self.mul(self.nodes[a].a, b))
        if self.nodes[b].op == NEG:
            return self.neg(self.mul(a, self.nodes[b].a))
        lo, hi = (a, b) if a <= b else (b, a)
        if lo == hi:
            return self._mk(("q", lo), Node(SQ, lo))
        return self._mk(("*", lo, hi), Node(MUL, lo, hi))

    def div(self, a: Optional[int], b: int) -> Optional[int]:
        if a is None:
            return None
        return self._mk(("/", a, b), Node(DIV, a, b))

    def sq(self, a: Optional[int]) -> Optional[int]:
        return self.mul(a, a)

    def un(self, op: str, a: int) -> int:
        return self._mk((op, a), Node(op, a))

    def dot(self, xs, ys) -> Optional[int]:
        acc = None
        for x, y in zip(xs, ys):
            acc = self.add(acc, self.mul(x, y))
        return acc


def _eval_nodes(nodes: List[Node], args, override: Optional[Dict[int, float]] = None) -> List[float]:
    v = [0.0] * len(nodes)
    for i, nd in enumerate(nodes):
        if override is not None and i in override:
            v[i] = override[i]
            continue
        op = nd.op
        a = v[nd.a] if nd.a >= 0 else 0.0
        b = v[nd.b] if nd.b >= 0 else 0.0
        if op == IN:
            v[i] = args[nd.arg][nd.idx]
        elif op == CONST:
            v[i] = nd.val
        elif op == ADD:
            v[i] = a + b
        elif op == SUB:
            v[i] = a - b
        elif op == MUL:
            v[i] = a * b
        elif op == DIV:
            v[i] = a / b
        elif op == NEG:
            v[i] = -a
        elif op == SQ:
            v[i] = a * a
        elif op == SIN:
            v[i] = math.sin(a)
        elif op == COS:
            v[i] = math.cos(a)
        elif op == SQRT:
            v[i] = math.sqrt(a)
    return v


def identify_system(primal: SXFunction, n_points: int = 3, seed: int = 7, donor: Optional[SXFunction] = None):
    """Returns (post, outs, beta, Kp, nodes) where `outs` are the output nodes that come out of the solve, `beta` the pre
    nodes forming the right-hand side (None = structurally zero row) and Kp[i][j] = (node, sign) | None with
    sum_j Kp[i][j] * value(outs[j]) = value(beta[i]); node ids refer to `nodes` (the primal's list, extended by donor
    nodes when the system matrix is taken from `donor`).

    `donor`: a function with the SAME system matrix and a full right-hand side.  The impact maps solve
    [M -J^T; J 0] [qd+; Lambda] = [M qd-; 0]: two rows of their right-hand side are structurally zero, so the probing
    below only sees a 9x7 response and cannot recover the matrix; the stance dynamics of the same foot have all nine
    right-hand-side entries, and their identified matrix (a function of q alone) is re-used after checking numerically
    that it maps the impact outputs onto the impact right-hand side."""
    import numpy as np
    nodes = primal.nodes
    n = len(nodes)
    post = [False] * n
    for i, nd in enumerate(nodes):
        if nd.op == SQRT:
            post[i] = True
        for o in (nd.a, nd.b):
            if o >= 0 and post[o]:
                post[i] = True
    frontier = sorted({o for i, nd in enumerate(nodes) if post[i] for o in (nd.a, nd.b) if o >= 0 and not post[o]})
    outs: List[int] = []
    for s in primal.stores:
        if post[s.node] and s.node not in outs:
            outs.append(s.node)
    rnd = random.Random(seed)
    beta: Optional[List[int]] = None
    Ks, vals = [], []
    for _ in range(n_points):
        args = [[rnd.uniform(-1.0, 1.0) for _ in range(sz)] for sz in primal.in_sizes]
        v0 = _eval_nodes(nodes, args)
        z0 = np.array([v0[o] for o in outs])

        def S(ov):
            v = _eval_nodes(nodes, args, ov)
            return np.array([v[o] for o in outs])
        b_here = []
        for f in frontier:
            if nodes[f].op == CONST:
                continue
            t = 0.37
            z1, z2 = S({f: v0[f] + t}), S({f: v0[f] + 2 * t})
            if np.abs(z2 - 2 * z1 + z0).max() < 1e-9 * max(1.0, np.abs(z0).max()) and np.abs(z1 - z0).max() > 0:
                b_here.append(f)
        if beta is None:
            beta = b_here
        if beta != b_here:
            raise RuntimeError(f"{primal.name}: right-hand-side nodes differ between sample points")
        if len(beta) != len(outs):
            if donor is None:
                raise RuntimeError(f"{primal.name}: {len(beta)} right-hand-side nodes for {len(outs)} solved outputs")
            return _identify_with_donor(primal, donor, post, outs, beta, n_points, seed)
        zero = {f: 0.0 for f in beta}
        zz = S(zero)
        if np.abs(zz).max() > 1e-12:
            raise RuntimeError(f"{primal.name}: solved outputs are not homogeneous in the right-hand side")
        M = np.zeros((len(outs), len(beta)))
        for k, f in enumerate(beta):
            ov = dict(zero)
            ov[f] = 1.0
            M[:, k] = S(ov)
        Ks.append(np.linalg.inv(M))
        vals.append(v0)
    nb = len(beta)
    Kp: List[List[Optional[Tuple[int, int]]]] = [[None] * nb for _ in range(nb)]
    pre_ids = [p for p in range(n) if not post[p]]
    for i in range(nb):
        for j in range(nb):
            ks = [K[i, j] for K in Ks]
            if all(abs(k) < 1e-10 for k in ks):
                continue
            hit = None
            for p in pre_ids:
                for sg in (1, -1):
                    if all(abs(sg * v[p] - k) <= 1e-9 * max(1.0, abs(k)) for v, k in zip(vals, ks)):
                        hit = (p, sg)
                        break
                if hit:
                    break
            if hit is None:
                raise RuntimeError(f"{primal.name}: system-matrix entry ({i},{j}) is not a node of the primal DAG")
            Kp[i][j] = hit
    return post, outs, beta, Kp, nodes


def _identify_with_donor(primal: SXFunction, donor: SXFunction, post, outs, beta_nodes, n_points: int, seed: int):
    import itertools
    import numpy as np
    dpost, douts, dbeta, dKp, dnodes = identify_system(donor)
    if len(douts) != len(outs):
        raise RuntimeError(f"{primal.name}: donor {donor.name} solves a system of another size")
    # donor pre nodes the matrix depends on, appended to the primal's node list
    need, stack = set(), [e[0] for row in dKp for e in row if e]
    while stack:
        q = stack.pop()
        if q in need:
            continue
        need.add(q)
        nd = dnodes[q]
        if nd.op == IN and nd.arg != 0:
            raise RuntimeError(f"{primal.name}: the donor matrix depends on an input the primal does not have")
        stack.extend(o for o in (nd.a, nd.b) if o >= 0)
    merged = list(primal.nodes)
    remap: Dict[int, int] = {}
    for q in sorted(need):
        nd = dnodes[q]
        merged.append(Node(nd.op, remap[nd.a] if nd.a >= 0 else -1, remap[nd.b] if nd.b >= 0 else -1, nd.val, nd.lit, nd.arg, nd.idx))
        remap[q] = len(merged) - 1
    post_m = list(post) + [False] * len(need)
    nb = len(outs)
    mult_cols = [j for j, o in enumerate(outs) if any(st.node == o and st.out >= 1 for st in primal.stores)]   # multiplier outputs
    rnd = random.Random(seed)
    samples = []
    for _ in range(n_points):
        args = [[rnd.uniform(-1.0, 1.0) for _ in range(sz)] for sz in primal.in_sizes]
        samples.append(_eval_nodes(merged, args))
    for flips in itertools.product((1, -1), repeat=len(mult_cols)):   # sign convention of the multipliers
        colsg = [1] * nb
        for j, f in zip(mult_cols, flips):
            colsg[j] = f
        Kp = [[(remap[e[0]], e[1] * colsg[j]) if e else None for j, e in enumerate(row)] for row in dKp]
        rows = [[sum((Kp[i][j][1] * v[Kp[i][j][0]] * v[outs[j]]) for j in range(nb) if Kp[i][j]) for v in samples] for i in range(nb)]
        beta: List[Optional[int]] = []
        sign: List[int] = []
        ok = True
        for i in range(nb):
            if all(abs(r) < 1e-9 for r in rows[i]):
                beta.append(None)
                sign.append(1)
                continue
            hit = None
            for bnode in beta_nodes:
                for sg in (1, -1):
                    if all(abs(sg * v[bnode] - r) <= 1e-8 * max(1.0, abs(r)) for v, r in zip(samples, rows[i])):
                        hit = (bnode, sg)
            if hit is None:
                ok = False
                break
            beta.append(hit[0])
            sign.append(hit[1])
        if ok and sorted(b for b in beta if b is not None) == sorted(beta_nodes):
            # fold the row signs into the matrix so that  K' out = beta  holds with +beta
            Kp = [[(e[0], e[1] * sign[i]) if e else None for e in Kp[i]] for i in range(nb)]
            return post_m, outs, beta, Kp, merged
    raise RuntimeError(f"{primal.name}: the system matrix of {donor.name} does not reproduce its right-hand side")


def symmetric_form(nodes: List[Node], in_sizes: List[int], Kp, n_points: int = 2, seed: int = 11):
    """Row permutation `perm` and row signs `sg` such that  diag(sg) * Kp[perm, :]  is symmetric with its non-zero diagonal
    entries (the mass-matrix block) first - or None.  Decided numerically at random inputs."""
    import numpy as np
    nb = len(Kp)
    rnd = random.Random(seed)
    mats = []
    for _ in range(n_points):
        args = [[rnd.uniform(-1.0, 1.0) for _ in range(sz)] for sz in in_sizes]
        v = _eval_nodes(nodes, args)
        mats.append(np.array([[(Kp[i][j][1] * v[Kp[i][j][0]] if Kp[i][j] else 0.0) for j in range(nb)] for i in range(nb)]))

    def close(a, b):
        return abs(a - b) <= 1e-9 * max(1.0, abs(a), abs(b))
    perm: List[int] = []
    sg: List[int] = []

    def place(i: int) -> bool:
        if i == nb:
            return True
        for r in range(nb):
            if r in perm:
                continue
            for s_i in (1, -1):
                ok = True
                for M in mats:
                    for j in range(i):
                        if not close(s_i * M[r][j], sg[j] * M[perm[j]][i]):
                            ok = False
                            break
                    if not ok:
                        break
                    # mass-matrix diagonal must be positive; constraint rows have a structurally zero diagonal
                    if s_i * M[r][i] < 0:
                        ok = False
                        break
                if ok:
                    perm.append(r)
                    sg.append(s_i)
                    if place(i + 1):
                        return True
                    perm.pop()
                    sg.pop()
        return False
    if not place(0):
        return None
    # the no-pivot LDL^T below needs every zero-diagonal (constraint) row after all mass-matrix rows
    zero_diag = [all(abs(M[perm[i]][i]) < 1e-14 for M in mats) for i in range(nb)]
    if any(zero_diag[i] and not zero_diag[j] for i in range(nb) for j in range(i + 1, nb)):
        return None
    return perm, sg


def make_solver(B: "Builder", K, sym):
    """Factorise the symbolic matrix K once and return solve(rhs) -> x with K x = rhs.
    `sym` = (perm, signs) from symmetric_form: LDL^T of the symmetric quasi-definite form [M -J^T; -J 0] without pivoting
    (positive pivots from the mass matrix, then the negative-definite Schur complement); otherwise a modified Gram-Schmidt QR
    (what the reference's generated code does)."""
    nb = len(K)
    if sym is not None:
        perm, sg = sym
        A = [[(K[perm[i]][j] if sg[i] > 0 else B.neg(K[perm[i]][j])) for j in range(nb)] for i in range(nb)]
        L: List[List[Optional[int]]] = [[None] * nb for _ in range(nb)]
        dval: List[Optional[int]] = []
        dinv: List[int] = []
        for j in range(nb):
            w = [B.mul(L[j][k], dval[k]) for k in range(j)]      # w_k = L_jk d_k
            dj = A[j][j]
            for k in range(j):
                dj = B.sub(dj, B.mul(L[j][k], w[k]))
            dval.append(dj)
            inv = B.div(B.const(1.0), dj)
            dinv.append(inv)
            for i in range(j + 1, nb):
                acc = A[i][j]
                for k in range(j):
                    acc = B.sub(acc, B.mul(L[i][k], w[k]))
                L[i][j] = B.mul(acc, inv)

        def solve_ldlt(rhs):
            b = [(rhs[perm[i]] if sg[i] > 0 else B.neg(rhs[perm[i]])) for i in range(nb)]
            y: List[Optional[int]] = []
            for i in range(nb):
                acc = b[i]
                for k in range(i):
                    acc = B.sub(acc, B.mul(L[i][k], y[k]))
                y.append(acc)
            x: List[Optional[int]] = [None] * nb
            for i in reversed(range(nb)):
                acc = B.mul(y[i], dinv[i])
                for k in range(i + 1, nb):
                    acc = B.sub(acc, B.mul(L[k][i], x[k]))
                x[i] = acc
            return x
        return solve_ldlt

    Q: List[List[Optional[int]]] = []          # Q[j] = column j
    R: List[List[Optional[int]]] = [[None] * nb for _ in range(nb)]
    rinv: List[int] = []
    for j in range(nb):
        v = [K[i][j] for i in range(nb)]
        for i in range(j):
            r = B.dot(Q[i], v)
            R[i][j] = r
            if r is not None:
                v = [B.sub(v[k], B.mul(r, Q[i][k])) for k in range(nb)]
        nrm = B.un(SQRT, B.dot(v, v))
        inv = B.div(B.const(1.0), nrm)
        rinv.append(inv)
        Q.append([B.mul(x, inv) for x in v])

    def solve_qr(rhs):
        y = [B.dot(Q[i], rhs) for i in range(nb)]
        x: List[Optional[int]] = [None] * nb
        for i in reversed(range(nb)):
            acc = y[i]
            for j in range(i + 1, nb):
                acc = B.sub(acc, B.mul(R[i][j], x[j]))
            x[i] = B.mul(acc, rinv[i])
        return x
    return solve_qr


def _finish(B: "Builder", like: SXFunction, stores: List[Store]) -> SXFunction:
    """Drop nodes nothing depends on, keep creation order."""
    used = [False] * len(B.nodes)
    for s in stores:
        used[s.node] = True
    for i in reversed(range(len(B.nodes))):
        if used[i]:
            nd = B.nodes[i]
            for o in (nd.a, nd.b):
                if o >= 0:
                    used[o] = True
    remap, new_nodes = {}, []
    for i, nd in enumerate(B.nodes):
        if used[i]:
            remap[i] = len(new_nodes)
            new_nodes.append(Node(nd.op, remap[nd.a] if nd.a >= 0 else -1, remap[nd.b] if nd.b >= 0 else -1,
                                  nd.val, nd.lit, nd.arg, nd.idx))
    out = SXFunction(name=like.name, n_in=like.n_in, n_out=like.n_out, in_sizes=like.in_sizes, nodes=new_nodes, out_sp=like.out_sp)
    out.stores = [Store(s.out, s.nz, remap[s.node], remap[s.node] + 1, s.row, s.col) for s in stores]
    return out


def _copier(B: "Builder", pn: List[Node]):
    cp: Dict[int, Optional[int]] = {}

    def copy(p: int) -> Optional[int]:
        stack = [p]
        while stack:
            q = stack[-1]
            if q in cp:
                stack.pop()
                continue
            nd = pn[q]
            need = [o for o in (nd.a, nd.b) if o >= 0 and o not in cp]
            if need:
                stack.extend(need)
                continue
            stack.pop()
            a = cp[nd.a] if nd.a >= 0 else None
            b = cp[nd.b] if nd.b >= 0 else None
            if nd.op == IN:
                cp[q] = B.inp(nd.arg, nd.idx)
            elif nd.op == CONST:
                cp[q] = B.const(nd.val)
            elif nd.op == ADD:
                cp[q] = B.add(a, b)
            elif nd.op == SUB:
                cp[q] = B.sub(a, b)
            elif nd.op == MUL:
                cp[q] = B.mul(a, b)
            elif nd.op == DIV:
                cp[q] = B.div(a, b)
            elif nd.op == NEG:
                cp[q] = B.neg(a)
            elif nd.op == SQ:
                cp[q] = B.sq(a)
            elif nd.op in (SIN, COS):
                cp[q] = B.un(nd.op, a) if a is not None else B.const(0.0 if nd.op == SIN else 1.0)
            else:
                raise RuntimeError("sqrt in a pre node")
        return cp[p]
    return copy


def structured_primal(primal: SXFunction, use_ldlt: bool = True, donor: Optional[SXFunction] = None) -> SXFunction:
    """The primal function with its symbolic QR solve replaced by an LDL^T solve of the identified system  K z = beta
    (same value up to rounding; ~2.2x fewer operations).  Used for the rollouts on the device."""
    post, outs, beta, Kp, nodes = identify_system(primal, donor=donor)
    nb = len(beta)
    B = Builder()
    copy = _copier(B, nodes)

    def kentry(e):
        if e is None:
            return None
        c = copy(e[0])
        return c if e[1] > 0 else B.neg(c)
    K = [[kentry(Kp[i][j]) for j in range(nb)] for i in range(nb)]
    solve = make_solver(B, K, symmetric_form(nodes, primal.in_sizes, Kp) if use_ldlt else None)
    z = solve([copy(b) if b is not None else None for b in beta])
    out_index = {o: k for k, o in enumerate(outs)}
    stores: List[Store] = []
    zero_id: Optional[int] = None
    for st in primal.stores:
        val = z[out_index[st.node]] if post[st.node] else copy(st.node)
        if val is None:
            if zero_id is None:
                B.nodes.append(Node(CONST, val=0.0, lit="0.0"))
                zero_id = len(B.nodes) - 1
            val = zero_id
        stores.append(Store(st.out, st.nz, val, val + 1, st.row, st.col))
    return _finish(B, primal, stores)


def implicit_jacobian(primal: SXFunction, par: SXFunction, use_ldlt: bool = True, donor: Optional[SXFunction] = None) -> SXFunction:
    """A drop-in replacement for `par` (same inputs, outputs, sparsity, store list) derived from `primal` by implicit
    differentiation of its linear solve.  par.out[k] must be d primal.out[k // n_in] / d in[k % n_in]."""
    post, outs, beta, Kp, pn = identify_system(primal, donor=donor)
    nb = len(beta)
    B = Builder()

    copy = _copier(B, pn)

    # ---- forward-mode tangents of pre nodes, one memo per input direction ------------------
    tan: Dict[Tuple[int, int], Dict[int, Optional[int]]] = {}

    def tangent(p: int, d: Tuple[int, int]) -> Optional[int]:
        memo = tan.setdefault(d, {})
        stack = [p]
        while stack:
            q = stack[-1]
            if q in memo:
                stack.pop()
                continue
            nd = pn[q]
            need = [o for o in (nd.a, nd.b) if o >= 0 and o not in memo]
            if need:
                stack.extend(need)
                continue
            stack.pop()
            ta = memo[nd.a] if nd.a >= 0 else None
            tb = memo[nd.b] if nd.b >= 0 else None
            op = nd.op
            if op == IN:
                memo[q] = B.const(1.0) if (nd.arg, nd.idx) == d else None
            elif op == CONST:
                memo[q] = None
            elif op == ADD:
                memo[q] = B.add(ta, tb)
            elif op == SUB:
                memo[q] = B.sub(ta, tb)
            elif op == NEG:
                memo[q] = B.neg(ta)
            elif op == MUL:
                t1 = B.mul(ta, copy(nd.b)) if ta is not None else None
                t2 = B.mul(copy(nd.a), tb) if tb is not None else None
                memo[q] = B.add(t1, t2)
            elif op == SQ:
                memo[q] = B.mul(B.mul(B.const(2.0), copy(nd.a)), ta) if ta is not None else None
            elif op == DIV:
                if ta is None and tb is None:
                    memo[q] = None
                else:
                    t1 = B.div(ta, copy(nd.b)) if ta is not None else None
                    t2 = B.mul(B.div(copy(q), copy(nd.b)), tb) if tb is not None else None
                    memo[q] = B.sub(t1, t2)
            elif op == SIN:
                memo[q] = B.mul(B.un(COS, copy(nd.a)), ta) if ta is not None else None
            elif op == COS:
                memo[q] = B.neg(B.mul(B.un(SIN, copy(nd.a)), ta)) if ta is not None else None
            else:
                raise RuntimeError("sqrt in a pre node")
        return memo[p]

    # ---- symbolic system matrix, factorised once ----------------------------------------------
    def kentry(e):
        if e is None:
            return None
        c = copy(e[0])
        return c if e[1] > 0 else B.neg(c)
    K = [[kentry(Kp[i][j]) for j in range(nb)] for i in range(nb)]
    solve = make_solver(B, K, symmetric_form(pn, primal.in_sizes, Kp) if use_ldlt else None)

    z = solve([copy(b) if b is not None else None for b in beta])
    out_index = {o: k for k, o in enumerate(outs)}

    # ---- one solve per input direction that is asked for -----------------------------------
    n_in = primal.n_in
    prim_store = {(s.out, s.nz): s.node for s in primal.stores}     # column vectors: nz == row for dense outputs
    for k, sp in enumerate(primal.out_sp):
        if sp.nnz != sp.nrow:
            raise RuntimeError("primal outputs are expected to be dense column vectors")
    dz: Dict[Tuple[int, int], List[Optional[int]]] = {}

    def dsol(d: Tuple[int, int]) -> List[Optional[int]]:
        if d not in dz:
            rhs = []
            for i in range(nb):
                acc = tangent(beta[i], d) if beta[i] is not None else None
                for j in range(nb):
                    if Kp[i][j] is None:
                        continue
                    t = tangent(Kp[i][j][0], d)
                    if t is None:
                        continue
                    term = B.mul(t, z[j])
                    acc = B.sub(acc, term) if Kp[i][j][1] > 0 else B.add(acc, term)
                rhs.append(acc)
            dz[d] = solve(rhs) if any(r is not None for r in rhs) else [None] * nb
        return dz[d]

    stores: List[Store] = []
    zero_id: Optional[int] = None
    for st in par.stores:
        po, pi = divmod(st.out, n_in)
        p = prim_store[(po, st.row)]
        d = (pi, st.col)
        if post[p]:
            val = dsol(d)[out_index[p]]
        else:
            val = tangent(p, d)
        if val is None:
            if zero_id is None:
                B.nodes.append(Node(CONST, val=0.0, lit="0.0"))
                zero_id = len(B.nodes) - 1
            val = zero_id
        stores.append(Store(st.out, st.nz, val, val + 1, st.row, st.col))

    return _finish(B, par, stores)
