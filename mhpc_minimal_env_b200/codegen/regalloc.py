"""Generator-side register allocation for the large Jacobian bodies.

The CasADi Jacobian DAGs keep ~900 FP64 values live at once; a thread has ~120 FP64 registers.  Left to ptxas, the
emitted straight-line code spills ~2x more than necessary (measured: 106 KB of spill loads per Dyn_BS_par evaluation at
128 registers against a Belady optimum of 57 KB for the same instruction order).  This emitter does the allocation itself:

* a fixed file of `cap` mutable FP64 variables (r0..r{cap-1}) that ptxas keeps in registers,
* an explicit per-thread spill array (`SX_SPILL(slot)`, a volatile local-memory array in the kernel) with slot reuse,
* Belady (furthest-next-use) eviction over the emitted instruction order,
* software prefetch: the operands of instruction s+lookahead are made resident at step s, so the reload latency
  (L1/L2) overlaps with `lookahead` independent instructions instead of stalling the consumer.

The arithmetic (operation set, operand order, evaluation order) is exactly that of the SSA form; only the storage of
intermediate values changes, so results are bit-identical to `emit_inc` (checked on the CPU by emitting the same
allocation as plain C: tests/test_cpu.py::test_codegen_bit_exact_against_reference_objects).

Macros expected from the includer:  SX_IN(i, j), SX_OUT(k, nz, row, col, v), SX_SINCOS(a, s, c).  The body declares its own
register variables r0..r{cap-1} and the volatile local spill array `sx_spill`.
"""
from __future__ import annotations

import bisect
import os
from typing import Dict, List, Tuple

from .sxdag import ADD, CONST, COS, DIV, IN, MUL, NEG, SIN, SQ, SQRT, SUB, SXFunction

_BIN = {ADD: "+", SUB: "-", MUL: "*", DIV: "/"}
INF = 1 << 60


def ilp_schedule(instrs, operands_of, window: int = 48, lat_fp: int = 6, lat_slow: int = 20, slow=None):
    """Window-limited list scheduling: re-order the straight-line program so that dependent instructions are at least
    `lat` issue slots apart when independent work is available among the next `window` instructions of the original
    order (bounded look-ahead keeps the live ranges, hence the spill traffic, close to the original order's)."""
    T = len(instrs)
    if window <= 1:
        return list(range(T))
    def_of = {}
    for t, ins in enumerate(instrs):
        for v in ins:
            def_of[v] = t
    preds = []
    for t, ins in enumerate(instrs):
        preds.append([def_of[p] for p in operands_of(ins) if p in def_of])
    done_at = [None] * T          # issue slot at which the result becomes available
    scheduled = [False] * T
    order = []
    head = 0                      # first unscheduled instruction in the original order
    clock = 0
    while head < T:
        best, best_ready = -1, None
        seen = 0
        t = head
        while t < T and seen < window:
            if not scheduled[t]:
                seen += 1
                ok = True
                ready = 0
                for q in preds[t]:
                    if not scheduled[q]:
                        ok = False
                        break
                    if done_at[q] > ready:
                        ready = done_at[q]
                if ok:
                    if ready <= clock:
                        best, best_ready = t, ready
                        break               # oldest ready instruction wins
                    if best < 0 or ready < best_ready:
                        best, best_ready = t, ready
            t += 1
        if best_ready > clock:
            clock = best_ready
        scheduled[best] = True
        order.append(best)
        done_at[best] = clock + (lat_slow if (slow and slow(instrs[best])) else lat_fp)
        clock += 1
        while head < T and scheduled[head]:
            head += 1
    return order


def pressure_order(fn: SXFunction, tie: str = "index", noise: float = 0.0, seed: int = 0) -> List[int]:
    """A topological order that keeps few values alive: list scheduling where the next node is the ready one with the smallest
    net change of the live set (+1 if its value is used later, -1 per operand it is the last user of).  Ties: `index` - the
    node that comes first in the given order; `unlock` - the node that makes ready a user which will itself end live ranges.
    `noise` > 0 perturbs the last tie-break (position in the given order) by seeded random priorities: the given order is a
    poor tie-break, and the best of a few dozen seeds moves a third fewer spill wavefronts (gen_all.py keeps the winners).
    Same DAG, same operations: only the evaluation order (hence the spill traffic) changes."""
    import random
    rnd = random.Random(seed)
    pri = [rnd.random() for _ in range(len(fn.nodes))]
    nodes = fn.nodes
    n = len(nodes)
    users: List[List[int]] = [[] for _ in range(n)]
    npred = [0] * n
    for i, nd in enumerate(nodes):
        ps = set(p for p in (nd.a, nd.b) if p >= 0)
        npred[i] = len(ps)
        for p in ps:
            users[p].append(i)
    users = [sorted(set(u)) for u in users]
    remaining = [len(u) for u in users]

    def ops_of(i: int):
        nd = nodes[i]
        return set(q for q in (nd.a, nd.b) if q >= 0 and nodes[q].op != CONST)

    def score(i: int):
        nd = nodes[i]
        if nd.op == CONST:
            return (-10, 0, 0)
        kills = sum(1 for p in ops_of(i) if remaining[p] == 1)
        base = (1 if users[i] else 0) - kills
        bonus = 0
        if tie == "unlock":
            for u in users[i]:
                if npred[u] == 1:
                    bonus = max(bonus, sum(1 for p in ops_of(u) if remaining[p] == 1))
        return (base, -bonus, i + noise * n * pri[i])

    ready = set(i for i in range(n) if npred[i] == 0)
    order: List[int] = []
    while ready:
        best = min(ready, key=score)
        ready.discard(best)
        order.append(best)
        for p in set(q for q in (nodes[best].a, nodes[best].b) if q >= 0):
            remaining[p] -= 1
        for u in users[best]:
            npred[u] -= 1
            if npred[u] == 0:
                ready.add(u)
    assert len(order) == n
    return order


def spill_cost(stats: Dict[str, int]) -> int:
    """data-pipe wavefronts of the spill traffic of one warp: a far-level (local / global) access of 8 bytes per lane is 8
    sectors, a shared-memory access 2 wavefronts"""
    return 8 * (stats["loads"] + stats["stores"]) + 2 * (stats["smem_loads"] + stats["smem_stores"])


def allocate(fn: SXFunction, cap: int = 104, lookahead: int = 48, window: int = 0, smem_slots: int = 0,
             far_lookahead: int = 0, quad_of=None, stage_slots: int = 0) -> Tuple[List[str], Dict[str, int]]:
    """`lookahead`: prefetch distance (instructions) for operands held in shared memory / produced recently;
    `far_lookahead` (> lookahead, 0 = off): longer distance used for operands that sit in LOCAL memory (HBM latency);
    `quad_of(store) -> element index | None`: outputs that land in a record are held back until the four elements of
    their aligned quad exist and leave together through SX_OUT4 (one full 32-byte sector per thread instead of four
    partial ones).
    `stage_slots` > 0: the far level is an explicit scratch array in global memory (SX_FAR_ST / SX_FAR_LD) and its reloads
    are ASYNCHRONOUS: `far_lookahead` instructions ahead of the consumer the value is copied scratch -> a per-thread staging
    slot in shared memory (SX_FAR_ISSUE(stage, slot): cp.async + commit), and `lookahead` instructions ahead it moves
    staging slot -> register after SX_FAR_WAIT(n) (n = copies issued since: a compile-time count).  The copy in flight holds
    no register, so the far distance can cover a DRAM round trip; with the synchronous form (stage_slots = 0) the reload
    itself pins a register for the whole distance, which limits it to ~128 instructions - not enough: ncu attributed 42 % of
    the stall samples of the Jacobian kernels to those loads."""
    nodes = fn.nodes
    n = len(nodes)
    is_const = [nd.op == CONST for nd in nodes]

    # sin/cos pairs on the same argument -> one sincos producing two values
    partner: Dict[int, int] = {}
    by_arg: Dict[int, Dict[str, int]] = {}
    for i, nd in enumerate(nodes):
        if nd.op in (SIN, COS):
            by_arg.setdefault(nd.a, {})[nd.op] = i
    for d in by_arg.values():
        if SIN in d and COS in d:
            partner[d[SIN]] = d[COS]
            partner[d[COS]] = d[SIN]

    # instruction list: every non-constant node in emitted order; a sincos pair is one instruction at the first member
    instrs: List[Tuple[int, ...]] = []      # (primary node, [secondary node])
    skip = set()
    # Inputs are not instructions: an input value lives in global memory for the whole body and can be (re)loaded at any time,
    # so it is treated like a value in the far spill level - loaded `far_lookahead` instructions ahead of its consumer and
    # dropped without a store when its register is needed.  (Loaded at the point of use, each of the 16 inputs of a
    # Jacobian body cost a full global-memory round trip: ncu showed the long-scoreboard stalls on exactly those lines.)
    in_of: Dict[int, Tuple[int, int]] = {}
    for i in range(n):
        if is_const[i] or i in skip:
            continue
        if nodes[i].op == IN:
            in_of[i] = (nodes[i].arg, nodes[i].idx)
            continue
        if i in partner:
            instrs.append((i, partner[i]))
            skip.add(partner[i])
        else:
            instrs.append((i,))
    # output quads: a flush pseudo-instruction (negative id) right after the instruction that completes the quad
    quads: Dict[int, Dict[int, object]] = {}
    quad_stores = set()
    if quad_of is not None:
        elem = {}
        for st in fn.stores:
            e = quad_of(st)
            if e is not None:
                elem[id(st)] = e
                quads.setdefault(e // 4, {})[e % 4] = st
        quads = {g: m for g, m in quads.items() if len(m) == 4}
        if quads:
            pos_of = {}
            for t, ins in enumerate(instrs):
                for v in ins:
                    pos_of[v] = t
            after: Dict[int, List[int]] = {}
            for g, m in quads.items():
                last = max((pos_of[st.node] for st in m.values() if not is_const[st.node]), default=-1)
                after.setdefault(last, []).append(g)
                for st in m.values():
                    quad_stores.add(id(st))
            new_instrs: List[Tuple[int, ...]] = [(-1 - g,) for g in sorted(after.get(-1, []))]
            for t, ins in enumerate(instrs):
                new_instrs.append(ins)
                for g in sorted(after.get(t, [])):
                    new_instrs.append((-1 - g,))
            instrs = new_instrs
    T = len(instrs)

    def operands(ins) -> List[int]:
        if ins[0] < 0:
            ops = []
            for k in range(4):
                p = quads[-1 - ins[0]][k].node
                if not is_const[p] and p not in ops:
                    ops.append(p)
            return ops
        nd = nodes[ins[0]]
        ops = []
        for p in (nd.a, nd.b):
            if p >= 0 and not is_const[p] and p not in ops:
                ops.append(p)
        return ops

    if window > 1:
        perm = ilp_schedule(instrs, operands, window, slow=lambda ins: ins[0] >= 0 and nodes[ins[0]].op in (DIV, SQRT, SIN, COS))
        instrs = [instrs[k] for k in perm]
    def_pos: Dict[int, int] = {}
    for t, ins in enumerate(instrs):
        if ins[0] < 0:
            continue
        for v in ins:
            def_pos[v] = t

    uses: Dict[int, List[int]] = {}
    for t, ins in enumerate(instrs):
        for p in operands(ins):
            uses.setdefault(p, []).append(t)
    stores_of: Dict[int, list] = {}
    for st in fn.stores:
        stores_of.setdefault(st.node, []).append(st)
        assert st.node not in in_of, "an input that is stored unchanged has no instruction to hang the store on"

    def next_use(v: int, t: int) -> int:
        u = uses.get(v)
        if not u:
            return INF
        k = bisect.bisect_right(u, t)
        return u[k] if k < len(u) else INF

    out: List[str] = []
    # constant outputs
    for st in fn.stores:
        if is_const[st.node] and id(st) not in quad_stores:
            out.append(f"SX_OUT({st.out}, {st.nz}, {st.row}, {st.col}, {nodes[st.node].lit});")

    reg_of: Dict[int, int] = {}          # value -> register
    free_regs = list(range(cap - 1, -1, -1))
    slot_of: Dict[int, int] = {}         # value -> spill slot (valid while the value is live)
    free_slots: List[int] = []
    n_slots = 0
    pins: Dict[int, int] = {}            # value -> last instruction index that pinned it
    stats = {"loads": 0, "stores": 0, "instrs": T, "smem_loads": 0, "smem_stores": 0, "smem_to_local": 0}
    # second level: `smem_slots` per-thread shared-memory slots (SX_SMEM(slot)), also managed furthest-next-use
    sm_of: Dict[int, int] = {}           # value -> shared-memory slot
    free_sm = list(range(smem_slots - 1, -1, -1))
    # asynchronous far level: value -> (staging slot, sequence number of its copy group)
    stage_of: Dict[int, Tuple[int, int]] = {}
    free_stage = list(range(stage_slots - 1, -1, -1))
    n_groups = 0
    far_st = (lambda sl, r: f"SX_FAR_ST({sl}, {r});") if stage_slots > 0 else (lambda sl, r: f"sx_spill[{sl}] = {r};")
    # The inputs start out in shared-memory slots: the body opens with one burst of loads (all in flight together: one global
    # round trip) whose results go straight into slots, and from there an input is an ordinary second-level value.  Loading
    # an input "early" into a register does not work: ptxas sinks the load to its first use to save the register, and the
    # consumer then waits out the whole round trip - ncu showed the long-scoreboard stalls on exactly those consumers.
    if smem_slots >= len(in_of) + 8 and os.environ.get("MHPC_RA_IN_SMEM", "1") == "1":
        ins_sorted = sorted(in_of, key=lambda q: in_of[q])
        assert len(ins_sorted) <= cap
        # all loads first (into the still empty register file), then all stores: the includer's loads and the shared-memory
        # slots are volatile, so the compiler keeps them in program order - load, store, load, store would serialise the round trips
        for k, v in enumerate(ins_sorted):
            out.append(f"r{k} = SX_IN({in_of[v][0]}, {in_of[v][1]});")
            stats["in_loads"] = stats.get("in_loads", 0) + 1
        for k, v in enumerate(ins_sorted):
            sm_of[v] = free_sm.pop()
            out.append(f"SX_SMEM({sm_of[v]}) = r{k};")
        in_of.clear()
    far_window = far_lookahead if far_lookahead > lookahead else 0

    def opnd(v: int) -> str:
        return nodes[v].lit if is_const[v] else f"r{reg_of[v]}"

    def evict(t: int, protect, before: bool = False) -> int:
        # `before`: called ahead of executing instruction t (prefetch stage), so a use AT t still counts
        best, best_nu = -1, -1
        for v in reg_of:
            if v in protect or pins.get(v, -1) >= t:
                continue
            nu = next_use(v, t - 1 if before else t)
            if nu > best_nu:
                best, best_nu = v, nu
        if best < 0:
            raise RuntimeError("register file too small for the pinned working set; raise cap or lower lookahead")
        r = reg_of.pop(best)
        if best_nu < INF and best not in slot_of and best not in sm_of and best not in in_of:
            spill_value(best, best_nu, r, t - 1 if before else t)
        return r

    def local_slot() -> int:
        nonlocal n_slots
        if free_slots:
            return free_slots.pop()
        n_slots += 1
        return n_slots - 1

    def stage_now(v: int, v_nu: int, now: int) -> None:
        """a value that has just gone to the far level although its next use lies inside the far window (its far pass is
        over): start its way back at once"""
        nonlocal n_groups
        if stage_slots > 0 and far_window > 0 and v_nu - now <= far_window and v not in stage_of and not free_stage:
            stats["stage_full_spill"] = stats.get("stage_full_spill", 0) + 1
        if stage_slots > 0 and far_window > 0 and v_nu - now <= far_window and v not in stage_of and free_stage:
            stats["staged_at_spill"] = stats.get("staged_at_spill", 0) + 1
            st_slot = free_stage.pop()
            out.append(f"SX_FAR_ISSUE({st_slot}, {slot_of[v]});")
            stage_of[v] = (st_slot, n_groups)
            n_groups += 1

    def spill_value(v: int, v_nu: int, r: int, now: int) -> None:
        """value v (in register r, next use v_nu) leaves the register file: shared memory if it is needed sooner than
        something already there (which then moves on to local memory), else local memory."""
        if smem_slots > 0:
            if free_sm:
                sm_of[v] = free_sm.pop()
                out.append(f"SX_SMEM({sm_of[v]}) = r{r};")
                stats["smem_stores"] += 1
                return
            far, far_nu = -1, -1
            for w in sm_of:
                nu = next_use(w, now)
                if nu > far_nu:
                    far, far_nu = w, nu
            if far_nu > v_nu:
                sl = sm_of.pop(far)
                if far_nu < INF and far not in slot_of:   # still needed: hand it down to local memory through the scratch register
                    slot_of[far] = local_slot()
                    out.append(f"rt = SX_SMEM({sl}); " + far_st(slot_of[far], "rt"))
                    stats["smem_to_local"] += 1
                    stats["stores"] += 1
                    stage_now(far, far_nu, now)
                sm_of[v] = sl
                out.append(f"SX_SMEM({sl}) = r{r};")
                stats["smem_stores"] += 1
                return
        slot_of[v] = local_slot()
        out.append(far_st(slot_of[v], f"r{r}"))
        stats["stores"] += 1
        stage_now(v, v_nu, now)

    def reload(p: int, r: int) -> None:
        if p in sm_of:
            out.append(f"r{r} = SX_SMEM({sm_of[p]});")
            stats["smem_loads"] += 1
        elif p in stage_of:
            st_slot, seq = stage_of.pop(p)
            out.append(f"SX_FAR_WAIT({n_groups - 1 - seq}); r{r} = SX_STAGE({st_slot});")
            free_stage.append(st_slot)
            stats["loads"] += 1
        elif p in slot_of:
            out.append(f"r{r} = SX_FAR_LD({slot_of[p]});" if stage_slots > 0 else f"r{r} = sx_spill[{slot_of[p]}];")
            stats["loads"] += 1
            if stage_slots > 0:
                stats["sync_loads"] = stats.get("sync_loads", 0) + 1
        else:
            out.append(f"r{r} = SX_IN({in_of[p][0]}, {in_of[p][1]});")
            stats["in_loads"] = stats.get("in_loads", 0) + 1

    def get_reg(t: int, protect, before: bool = False) -> int:
        if free_regs:
            return free_regs.pop()
        return evict(t, protect, before)

    def release(v: int) -> None:
        if v in reg_of:
            free_regs.append(reg_of.pop(v))
        if v in slot_of:
            free_slots.append(slot_of.pop(v))
        if v in sm_of:
            free_sm.append(sm_of.pop(v))
        if v in stage_of:                   # staged for a consumer that never came to fetch it (cannot happen by construction; kept safe)
            free_stage.append(stage_of.pop(v)[0])

    def make_resident(q: int, t: int, local_only: bool = False) -> None:
        """operands of instruction q become resident and pinned (issued at step t <= q)"""
        nonlocal n_groups
        ops = operands(instrs[q])
        for p in ops:
            if p not in reg_of and p not in slot_of and p not in sm_of and p not in in_of:
                continue   # not computed yet: it is produced (and pinned) before q
            if local_only and (p in reg_of or p in sm_of):
                continue   # the far pass only starts the long-latency (local memory) reloads early
            if local_only and stage_slots > 0 and p in slot_of:
                # asynchronous reload: scratch -> staging slot; no register, no pin.  Without a free staging slot the value is
                # fetched synchronously when its turn comes in the near pass.
                if p not in stage_of and not free_stage:
                    stats["stage_full"] = stats.get("stage_full", 0) + 1
                    pending_far.append((q, p))      # retried as soon as a staging slot is free (see the main loop)
                if p not in stage_of and free_stage:
                    st_slot = free_stage.pop()
                    out.append(f"SX_FAR_ISSUE({st_slot}, {slot_of[p]});")
                    stage_of[p] = (st_slot, n_groups)
                    n_groups += 1
                continue
            if p not in reg_of:
                r = get_reg(t, set(ops), before=True)
                reg_of[p] = r
                reload(p, r)
            pins[p] = max(pins.get(p, -1), q)

    far = far_lookahead if far_lookahead > lookahead else 0
    pending_far: List[Tuple[int, int]] = []     # (consumer, value): far-level values whose asynchronous reload found no staging slot
    for q in range(min(lookahead, T)):
        make_resident(q, 0)
    for t, ins in enumerate(instrs):
        while pending_far and free_stage:
            q, p = pending_far.pop(0)
            if q <= t + lookahead or p in reg_of or p in sm_of or p in stage_of or p not in slot_of:
                continue                        # too late to be worth it, or no longer where it was
            st_slot = free_stage.pop()
            out.append(f"SX_FAR_ISSUE({st_slot}, {slot_of[p]});")
            stage_of[p] = (st_slot, n_groups)
            n_groups += 1
            stats["staged_late"] = stats.get("staged_late", 0) + 1
        if far and t + far < T:
            make_resident(t + far, t, local_only=True)
        if t + lookahead < T:
            make_resident(t + lookahead, t)
        v = ins[0]
        ops = operands(ins)
        for p in ops:   # safety net: an operand defined inside the lookahead window may have been evicted since
            if p not in reg_of:
                r = get_reg(t, set(ops), before=True)
                reg_of[p] = r
                reload(p, r)
                stats["late_loads"] = stats.get("late_loads", 0) + 1
        if v < 0:       # flush one complete output quad
            g = -1 - v
            args = []
            for k in range(4):
                st = quads[g][k]
                args.append(f"{st.out}, {st.nz}, {st.row}, {st.col}, {opnd(st.node)}")
            out.append(f"SX_OUT4({4 * g}, " + ", ".join(args) + ");")
            stats["quads"] = stats.get("quads", 0) + 1
            for p in ops:
                if next_use(p, t) == INF and pins.get(p, -1) <= t:
                    release(p)
            continue
        nd = nodes[v]
        a = opnd(nd.a) if nd.a >= 0 else None
        b = opnd(nd.b) if nd.b >= 0 else None
        # operands that die here free their registers before the destination is chosen
        for p in ops:
            if next_use(p, t) == INF and pins.get(p, -1) <= t:
                release(p)
        protect = set(p for p in ops if p in reg_of)
        dests = []
        for w in ins:
            r = get_reg(t, protect | set(x for x, _ in dests))
            reg_of[w] = r
            dests.append((w, r))
            nu = next_use(w, t)
            if nu - t <= lookahead:   # its first consumer has already had its prefetch step: pin until then
                pins[w] = max(pins.get(w, -1), nu)
        if len(ins) == 2:
            s_id, c_id = (ins[0], ins[1]) if nodes[ins[0]].op == SIN else (ins[1], ins[0])
            out.append(f"SX_SINCOS({a}, r{reg_of[s_id]}, r{reg_of[c_id]});")
        elif nd.op == DIV and is_const[nd.a] and nodes[nd.a].val == 1.0 and not is_const[nd.b]:
            out.append(f"r{dests[0][1]} = SX_RCP({b});")
        elif nd.op in _BIN:
            out.append(f"r{dests[0][1]} = {a} {_BIN[nd.op]} {b};")
        elif nd.op == NEG:
            out.append(f"r{dests[0][1]} = -{a};")
        elif nd.op == SQ:
            out.append(f"r{dests[0][1]} = {a} * {a};")
        elif nd.op == SQRT:
            out.append(f"r{dests[0][1]} = sqrt({a});")
        elif nd.op == SIN:
            out.append(f"r{dests[0][1]} = sin({a});")
        elif nd.op == COS:
            out.append(f"r{dests[0][1]} = cos({a});")
        else:
            raise AssertionError(nd.op)
        for w, r in dests:
            for st in stores_of.get(w, []):
                if id(st) not in quad_stores:
                    out.append(f"SX_OUT({st.out}, {st.nz}, {st.row}, {st.col}, r{r});")
            if not uses.get(w):
                release(w)
    stats["slots"] = n_slots
    stats["cap"] = cap
    stats["stage_slots"] = stage_slots
    decl = ["double rt, " + ", ".join(f"r{k}" for k in range(cap)) + ";"]
    decl.append(f"SX_FAR_DECL({max(n_slots, 1)});" if stage_slots > 0 else f"volatile double sx_spill[{max(n_slots, 1)}];")
    if stage_slots > 0 and n_groups:
        out.append("SX_FAR_WAIT(0);")       # nothing of this body is in flight when the next one starts
    return decl + out, stats


def emit_inc_regalloc(fn: SXFunction, cap: int = 104, lookahead: int = 48, window: int = 0, smem_slots: int = 0,
                      far_lookahead: int = 0, quad_of=None, stage_slots: int = 0) -> Tuple[str, Dict[str, int]]:
    body, stats = allocate(fn, cap, lookahead, window, smem_slots, far_lookahead, quad_of, stage_slots)
    hdr = [
        f"// GENERATED at build time by mhpc_minimal_env_b200/codegen (regalloc) from the CasADi SX function `{fn.name}`.",
        f"// register file {cap} FP64, lookahead {lookahead}, ILP window {window}: {stats['instrs']} instructions, {stats['loads']} reloads, "
        f"{stats['stores']} spill stores, {stats['slots']} local slots; shared-memory level: {smem_slots} slots, {stats['smem_loads']} loads, {stats['smem_stores']} stores. Do not edit.",
    ]
    return "\n".join(hdr + body) + "\n", stats


def emit_c_regalloc(fn: SXFunction, symbol: str, cap: int = 104, lookahead: int = 48, window: int = 0, smem_slots: int = 0,
                    far_lookahead: int = 0, quad_of=None, stage_slots: int = 0) -> str:
    body, stats = allocate(fn, cap, lookahead, window, smem_slots, far_lookahead, quad_of, stage_slots)
    return "\n".join([
        "#include <math.h>",
        f"void {symbol}(const double** arg, double** res) {{",
        f"  volatile double sx_smem_c[{max(smem_slots, 1)}];",
        "#define SX_SMEM(s) sx_smem_c[s]",
        "#define SX_IN(i, j) (arg[i][j])",
        "#define SX_OUT(k, nz, row, col, v) res[k][nz] = (v)",
        "#define SX_OUT4(e0, k0, n0, r0, c0, v0, k1, n1, r1, c1, v1, k2, n2, r2, c2, v2, k3, n3, r3, c3, v3) "
        "do { res[k0][n0] = (v0); res[k1][n1] = (v1); res[k2][n2] = (v2); res[k3][n3] = (v3); } while (0)",
        "#define SX_SINCOS(a, s, c) do { const double t_ = (a); (s) = sin(t_); (c) = cos(t_); } while (0)",
        "#define SX_RCP(x) (1.0 / (x))",
        # the asynchronous far level, executed synchronously: a copy in flight is modelled by a NaN in the staging slot until
        # SX_FAR_WAIT has been told that at most n later copies may still be pending (so a wait count that is too large shows)
        f"  volatile double sx_stage_c[{max(stage_slots, 1)}]; double sx_pend_v[{max(stage_slots, 1)}]; long sx_pend_g[{max(stage_slots, 1)}]; long sx_ngrp = 0;",
        f"  for (int q_ = 0; q_ < {max(stage_slots, 1)}; ++q_) sx_pend_g[q_] = -1;",
        "#define SX_FAR_DECL(n) volatile double sx_far_c[n]",
        "#define SX_FAR_ST(s, r) sx_far_c[s] = (r)",
        "#define SX_FAR_LD(s) sx_far_c[s]",
        "#define SX_FAR_ISSUE(st, s) do { sx_stage_c[st] = NAN; sx_pend_v[st] = sx_far_c[s]; sx_pend_g[st] = sx_ngrp++; } while (0)",
        f"#define SX_FAR_WAIT(n) do {{ for (int q_ = 0; q_ < {max(stage_slots, 1)}; ++q_) if (sx_pend_g[q_] >= 0 && sx_pend_g[q_] < sx_ngrp - (n)) {{ sx_stage_c[q_] = sx_pend_v[q_]; sx_pend_g[q_] = -1; }} }} while (0)",
        "#define SX_STAGE(st) sx_stage_c[st]",
        *("  " + s for s in body),
        "#undef SX_IN", "#undef SX_OUT", "#undef SX_OUT4", "#undef SX_SINCOS", "#undef SX_RCP", "#undef SX_SMEM",
        "#undef SX_FAR_DECL", "#undef SX_FAR_ST", "#undef SX_FAR_LD", "#undef SX_FAR_ISSUE", "#undef SX_FAR_WAIT", "#undef SX_STAGE",
        "}", "",
    ])
