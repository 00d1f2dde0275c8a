"""Exhaustive interleaving check of the peer-memory exchange protocol (csrc/peer.cu) on a small model: CPU only.

The CUDA kernels cannot be run here, but the protocol they implement -- step numbers, per-rank flag and ack words, one
region per rank whose slots are overwritten every step -- is small enough to explore completely.  The model has one
process per rank executing the same launch sequence as the library; every store into a peer's region, every flag /
ack store and every wait is a separate atomic action, and a breadth-first search visits every reachable interleaving
for 2 and 3 ranks over several steps.  Checked in every state: no rank ever reads a slot that holds another step's
message (the overwrite hazard the acks exist for), and every non-final state has an enabled action (no deadlock).

Two launch sequences are modelled:
  * the catalog exchange       publish -> wait -> consume -> ack                         (rs_peer_publish / wait / ack)
  * the embedding lookup       ids publish (region A) -> rows kernel: wait A, wait acks B, write rows into the
                               requesters' B regions, flag B, ack A -> restore: wait B, consume, ack B
                               (rs_peer_a2a_ids / rs_peer_a2a_rows / rs_peer_a2a_restore)
The model also shows the acks are NECESSARY: without the ack wait in publish, the search finds the overwrite.
"""

from collections import deque


def _explore(n_ranks, n_steps, program, check_acks=True):
    """program(rank) -> list of micro-ops for ONE step.  Micro-ops (region X is "A" or "B"):
       ("begin", X)            seq[X] of this rank += 1 is decided; blocks until acks[X][rank][g] >= seq - 1 for all g
       ("write", X, g)         data[X][g][rank] = seq            (this rank's message lands in rank g's region)
       ("flag", X, g)          flags[X][g][rank] = seq
       ("wait", X)             blocks until flags[X][rank][g] >= seq for all g
       ("read", X)             asserts data[X][rank][g] == seq for all g   (what the consumer kernels see)
       ("ack", X, g)           acks[X][g][rank] = seq
    Stores of one rank are issued in program order (kernel boundaries / release semantics in the real code)."""
    progs = [program(r) for r in range(n_ranks)]
    regions = sorted({op[1] for p in progs for op in p})
    zero = tuple(tuple(0 for _ in range(n_ranks)) for _ in range(n_ranks))
    init = (tuple((0, 0) for _ in range(n_ranks)),                       # per rank: (step index, pc)
            tuple((x, zero, zero, zero) for x in regions))               # per region: data, flags, acks  [owner][writer]
    seen, todo, states = {init}, deque([init]), 0

    def put(mat, owner, writer, v):
        row = list(mat[owner])
        row[writer] = v
        out = list(mat)
        out[owner] = tuple(row)
        return tuple(out)

    while todo:
        state = todo.popleft()
        states += 1
        pcs, mem = state
        mem_d = {x: (d, f, a) for x, d, f, a in mem}
        progressed, finished = False, True
        for r in range(n_ranks):
            step, pc = pcs[r]
            if step == n_steps:
                continue
            finished = False
            op = progs[r][pc]
            x = op[1]
            d, f, a = mem_d[x]
            seq = step + 1
            if op[0] == "begin":
                if check_acks and any(a[r][g] < seq - 1 for g in range(n_ranks)):
                    continue                                              # blocked: a peer has not consumed the last step
            elif op[0] == "wait":
                if any(f[r][g] < seq for g in range(n_ranks)):
                    continue                                              # blocked: a message has not arrived
            elif op[0] == "read":
                for g in range(n_ranks):
                    assert d[r][g] == seq, (f"rank {r} step {seq}: slot {g} of region {x} holds step {d[r][g]}", state)
            elif op[0] == "write":
                d = put(d, op[2], r, seq)
            elif op[0] == "flag":
                f = put(f, op[2], r, seq)
            elif op[0] == "ack":
                a = put(a, op[2], r, seq)
            progressed = True
            npc = pc + 1
            nstep = step
            if npc == len(progs[r]):
                npc, nstep = 0, step + 1
            new_pcs = list(pcs)
            new_pcs[r] = (nstep, npc)
            new_mem = tuple((y, *((d, f, a) if y == x else mem_d[y])) for y in regions)
            nxt = (tuple(new_pcs), new_mem)
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
        assert finished or progressed, ("deadlock", state)
    return states


def _exchange(n_ranks):
    def program(rank):
        ops = [("begin", "A")]
        ops += [("write", "A", g) for g in range(n_ranks)]
        ops += [("flag", "A", g) for g in range(n_ranks)]               # last CTA, after the fence: flags follow all writes
        ops += [("wait", "A"), ("read", "A")]
        ops += [("ack", "A", g) for g in range(n_ranks)]
        return ops
    return program


def _lookup(n_ranks):
    def program(rank):
        ops = [("begin", "A")]                                           # rs_peer_a2a_ids
        ops += [("write", "A", g) for g in range(n_ranks)]
        ops += [("flag", "A", g) for g in range(n_ranks)]
        ops += [("wait", "A"), ("begin", "B"), ("read", "A")]           # rs_peer_a2a_rows: lists in, rows out
        ops += [("write", "B", g) for g in range(n_ranks)]
        ops += [("flag", "B", g) for g in range(n_ranks)]
        ops += [("ack", "A", g) for g in range(n_ranks)]
        ops += [("wait", "B"), ("read", "B")]                            # rs_peer_a2a_restore
        ops += [("ack", "B", g) for g in range(n_ranks)]
        return ops
    return program


def test_catalog_exchange_protocol_has_no_overwrite_and_no_deadlock():
    assert _explore(2, 4, _exchange(2)) > 200          # states visited: the search really branched
    assert _explore(3, 2, _exchange(3)) > 1000


def test_lookup_protocol_has_no_overwrite_and_no_deadlock():
    assert _explore(2, 3, _lookup(2)) > 500
    assert _explore(3, 2, _lookup(3)) > 5000


def test_without_the_ack_wait_a_slot_is_overwritten_early():
    import pytest
    with pytest.raises(AssertionError, match="holds step"):
        _explore(2, 3, _exchange(2), check_acks=False)
