"""CPU model check of the hand-over protocol of the dual-tile attention kernel (powerpaint_b200/csrc/attention2.cuh).

The kernel's correctness rests on a handful of mbarriers shared by one TMA producer, one tcgen05 issuer and two groups
of softmax warps, all running at their own pace. Round 2 found a real race there (a softmax warp a key block ahead of a
slower warp of its group released PV over raw scores: NaN rows, but only at 16384 tokens) that no amount of testing at
small sizes had shown. This file restates the protocol as a small transition system — every `mbar_wait` with the parity
the source uses, every arrival / commit, the in-order tcgen05 pipe, the K / V^T ring — and explores EVERY interleaving
for small key counts, checking at each tensor-core operation and each softmax access that the data it touches is the
data it expects:

  S(k) executes   -> its key block is loaded in the ring stage, and its score buffer is free: PV(k - NBUF) has executed
  softmax reads   -> the buffer holds S(k), complete
  O rescale       -> every PV of the tile's earlier blocks has executed
  PV(k) executes  -> the buffer holds S(k) and EVERY warp of the tile has written its rows of P(k); the V^T stage holds
                     key block j with the ones row (row sums) written
  TMA load        -> the stage it overwrites is no longer needed by any pending S / PV
  epilogue        -> all PV of the tile have executed
  no deadlock     -> every run ends with all agents finished and the pipe empty

The model FOUND a race this way (end of round 2, after the GPU budget was spent): with three score buffers a warp can
finish the last key block while a slower warp of its group still holds back PV(nkv - 2); `bar_pv_done` is then two
phases short of the parity the epilogue waits for, and a parity wait cannot tell "two behind" from "done" — the wait
passes at once and O is read without the last two key blocks (a mild, silent error in 32 rows; no test tolerance would
see it). The kernel now takes the rescale path's wait on the last block (`final_guard`), where the barrier is at most
one phase short. `test_the_unguarded_epilogue_aliases` keeps the finding reproducible.
Three mutants show that the model can see the bugs it is meant to exclude: that unguarded epilogue, the
start-of-round-2 design (ONE P hand-over barrier per tile instead of one per key-block parity) and a ring slot released
after tile 0's PV instead of tile 1's.
The model is tied to the source by `test_model_matches_the_source`: the waits / commits it restates must be present in
attention2.cuh verbatim, so a protocol edit there fails here until the model follows.

Model conventions: an mbarrier is (completed phases, pending arrivals); `wait(parity)` passes iff the phase in progress
has the other parity (PTX `mbarrier.try_wait.parity`). tcgen05 operations of the single issuing thread execute in issue
order; a commit arrives on its barrier when everything issued before it has executed.
"""
import os
from collections import deque

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.path.dirname(HERE), "powerpaint_b200", "csrc", "attention2.cuh")


class Violation(Exception):
    pass


def _issuer_program(nkv, NBUF, S, kv_release_tile=1, single_p_barrier=False):
    """attention2.cuh `warp == 1` branch, linearised. Instructions: ("wait", bar, parity) | ("mma", op) | ("commit", bar)"""
    prog = [("wait", ("q",), 0)]
    nprod = 2 * nkv

    def issue_s(k):
        j, t, buf = k >> 1, k & 1, k % NBUF
        out = []
        if t == 0:
            out.append(("wait", ("kv_full", j % S), (j // S) & 1))
        out += [("mma", ("S", k)), ("commit", ("s_full", buf))]
        return out

    for k in range(min(NBUF, nprod)):
        prog += issue_s(k)
    for k in range(nprod):
        j, t = k >> 1, k & 1
        if single_p_barrier:  # start-of-round-2 design: one hand-over barrier per tile, phase = block parity
            prog.append(("wait", ("p_full", t, 0), j & 1))
        else:
            prog.append(("wait", ("p_full", t, j & 1), (j >> 1) & 1))
        prog += [("mma", ("PV", k)), ("commit", ("pv_done", t))]
        if t == kv_release_tile:
            prog.append(("commit", ("kv_empty", j % S)))
        if k + NBUF < nprod:
            prog.append(("wait", ("pv_done", t), j & 1))
            prog += issue_s(k + NBUF)
    return prog


def explore(nkv, NBUF, S, W, single_p_barrier=False, kv_release_tile=1, allow_rescale=True, final_guard=True,
            max_states=3_000_000):
    """exhaustive search over all interleavings; returns the number of distinct states; raises Violation"""
    prog = _issuer_program(nkv, NBUF, S, kv_release_tile, single_p_barrier)
    counts = {("q",): 1}
    for s in range(S):
        counts[("kv_full", s)] = 1
        counts[("kv_empty", s)] = 1
    for b in range(NBUF):
        counts[("s_full", b)] = 1
    for t in range(2):
        counts[("pv_done", t)] = 1
        for par in range(2):
            counts[("p_full", t, par)] = W
    bars = sorted(counts)
    bidx = {b: i for i, b in enumerate(bars)}

    def p_bar(t, j):
        return ("p_full", t, 0 if single_p_barrier else (j & 1))

    # ---- state: (pc, fifo, barriers, warps, producer, bufs, pwritten, stages, ones, pv_exec, s_exec)
    #  barriers: tuple of (completed, pending); warps[t][w] = (j, step); producer = (j, loading?)
    #  bufs[b] = k of the score product held (or -1); pwritten[b][w] = k written by warp w of the buffer's tile (or -1)
    #  stages[s] = (block, loaded) ; ones[s] = block whose ones row is written ; pv_exec[t], s_exec = executed counts
    def arrive(bs, bar):
        i = bidx[bar]
        c, p = bs[i]
        p += 1
        if p == counts[bar]:
            c, p = c + 1, 0
        return bs[:i] + ((c, p),) + bs[i + 1:]

    def passed(bs, bar, parity):
        return (bs[bidx[bar]][0] & 1) != parity

    init = (0, (), tuple((0, 0) for _ in bars), tuple(tuple((0, 0) for _ in range(W)) for _ in range(2)), (-1, 0),
            tuple(-1 for _ in range(NBUF)), tuple(tuple(-1 for _ in range(W)) for _ in range(NBUF)),
            tuple((-1, 0) for _ in range(S)), tuple(-1 for _ in range(S)), (0, 0), frozenset())
    seen = {init}
    todo = deque([init])
    DONE_STEP = 9

    def successors(st):
        pc, fifo, bs, warps, prod, bufs, pw, stages, ones, pv_exec, s_exec = st
        out = []
        # ---------------- TMA producer (warp 0): Q first, then the K / V^T ring
        pj, loading = prod
        if pj == -1:
            out.append((pc, fifo, arrive(bs, ("q",)), warps, (0, 0), bufs, pw, stages, ones, pv_exec, s_exec))
        elif pj < nkv:
            s = pj % S
            if not loading:
                if passed(bs, ("kv_empty", s), ((pj // S) & 1) ^ 1):
                    old = stages[s][0]
                    if old >= 0:  # the block being overwritten must not be needed any more
                        if not ({2 * old, 2 * old + 1} <= s_exec) or pv_exec[0] <= old or pv_exec[1] <= old:
                            raise Violation(f"TMA overwrites ring stage {s} (block {old}) while S / PV of it are pending")
                    ns = stages[:s] + ((pj, 0),) + stages[s + 1:]
                    out.append((pc, fifo, bs, warps, (pj, 1), bufs, pw, ns, ones, pv_exec, s_exec))
            else:
                ns = stages[:s] + ((pj, 1),) + stages[s + 1:]
                out.append((pc, fifo, arrive(bs, ("kv_full", s)), warps, (pj + 1, 0), bufs, pw, ns, ones, pv_exec,
                            s_exec))
        # ---------------- tcgen05 issuer (warp 1)
        if pc < len(prog):
            ins = prog[pc]
            if ins[0] == "wait":
                if passed(bs, ins[1], ins[2]):
                    out.append((pc + 1, fifo, bs, warps, prod, bufs, pw, stages, ones, pv_exec, s_exec))
            else:
                out.append((pc + 1, fifo + (ins,), bs, warps, prod, bufs, pw, stages, ones, pv_exec, s_exec))
        # ---------------- the tensor-core pipe executes its head
        if fifo:
            ins, rest = fifo[0], fifo[1:]
            if ins[0] == "commit":
                out.append((pc, rest, arrive(bs, ins[1]), warps, prod, bufs, pw, stages, ones, pv_exec, s_exec))
            else:
                kind, k = ins[1]
                j, t, buf = k >> 1, k & 1, k % NBUF
                stg = stages[j % S]
                if kind == "S":
                    if stg != (j, 1):
                        raise Violation(f"S({k}) reads ring stage {j % S} holding {stg}, not the loaded block {j}")
                    prev = bufs[buf]
                    if prev >= 0 and pv_exec[prev & 1] <= (prev >> 1):
                        raise Violation(f"S({k}) overwrites buffer {buf} before PV({prev}) has read P({prev})")
                    nb = bufs[:buf] + (k,) + bufs[buf + 1:]
                    npw = pw[:buf] + (tuple(-1 for _ in range(W)),) + pw[buf + 1:]
                    out.append((pc, rest, bs, warps, prod, nb, npw, stages, ones, pv_exec, s_exec | {k}))
                else:
                    if bufs[buf] != k:
                        raise Violation(f"PV({k}) reads buffer {buf} holding S({bufs[buf]})")
                    if any(x != k for x in pw[buf]):
                        raise Violation(f"PV({k}) released while P({k}) is incomplete: rows written {pw[buf]} "
                                        "(raw fp32 scores read as fp16 probabilities)")
                    if stg != (j, 1) or ones[j % S] != j:
                        raise Violation(f"PV({k}) reads V^T stage {j % S}: holds {stg}, ones row of block {ones[j % S]}")
                    if pv_exec[t] != j:
                        raise Violation(f"PV({k}) out of order for tile {t}")
                    npv = (pv_exec[0] + (t == 0), pv_exec[1] + (t == 1))
                    out.append((pc, rest, bs, warps, prod, bufs, pw, stages, ones, npv, s_exec))
        # ---------------- softmax warps (warps 2..9): W per tile in the model
        for t in range(2):
            for w in range(W):
                j, step = warps[t][w]
                if step == DONE_STEP:
                    continue

                def upd(nj, nstep, nbs=bs, npw=pw, nones=ones):
                    nw = warps[t][:w] + ((nj, nstep),) + warps[t][w + 1:]
                    nwarps = (nw, warps[1]) if t == 0 else (warps[0], nw)
                    return (pc, fifo, nbs, nwarps, prod, bufs, npw, stages, nones, pv_exec, s_exec)

                if j == nkv:  # epilogue: wait for the tile's last PV
                    if passed(bs, ("pv_done", t), (nkv - 1) & 1):
                        if pv_exec[t] != nkv:
                            raise Violation(f"epilogue of tile {t} reads O after {pv_exec[t]} of {nkv} PV products")
                        out.append(upd(j, DONE_STEP))
                    continue
                k = 2 * j + t
                buf = k % NBUF
                if step == 0:  # mbar_wait(bar_s_full(buf), (k / NBUF) & 1), then the score reads
                    if passed(bs, ("s_full", buf), (k // NBUF) & 1):
                        if bufs[buf] != k or k not in s_exec:
                            raise Violation(f"softmax warp ({t},{w}) reads buffer {buf} for S({k}); it holds S({bufs[buf]})")
                        if final_guard and NBUF == 3 and j == nkv - 1 and j > 0:
                            out.append(upd(j, 1))       # last block: wait for PV(j - 1) like a rescale (see the header)
                        else:
                            out.append(upd(j, 2))       # steady state: no rescale
                            if allow_rescale and j > 0:
                                out.append(upd(j, 1))   # new row maximum: O must be rescaled first
                elif step == 1:  # mbar_wait(bar_pv_done(t), (j - 1) & 1), then O *= alpha
                    if passed(bs, ("pv_done", t), (j - 1) & 1):
                        if pv_exec[t] != j:
                            raise Violation(f"warp ({t},{w}) rescales O at block {j} with {pv_exec[t]} PV products executed")
                        out.append(upd(j, 2))
                else:  # P store over the scores, (warp 2: ones row), arrive on the hand-over barrier
                    if bufs[buf] != k:
                        raise Violation(f"warp ({t},{w}) stores P({k}) into buffer {buf} holding S({bufs[buf]})")
                    npw = pw[:buf] + (pw[buf][:w] + (k,) + pw[buf][w + 1:],) + pw[buf + 1:]
                    nones = ones
                    if t == 0 and w == 0:
                        if stages[j % S] != (j, 1):
                            raise Violation(f"ones row written into stage {j % S} holding {stages[j % S]}")
                        nones = ones[:j % S] + (j,) + ones[j % S + 1:]
                    out.append(upd(j + 1, 0, nbs=arrive(bs, p_bar(t, j)), npw=npw, nones=nones))
        return out

    while todo:
        st = todo.popleft()
        nxt = successors(st)
        if not nxt:
            pc, fifo, _, warps, prod, *_ = st
            finished = (pc == len(prog) and not fifo and prod[0] == nkv
                        and all(x[1] == DONE_STEP for tw in warps for x in tw))
            if not finished:
                raise Violation(f"deadlock: issuer at {pc}/{len(prog)}, pipe {fifo[:2]}, warps {warps}, producer {prod}")
        for n in nxt:
            if n not in seen:
                seen.add(n)
                todo.append(n)
                if len(seen) > max_states:
                    raise RuntimeError("state space larger than expected")
    return len(seen)


@pytest.mark.parametrize("NBUF,S,W,nkv", [
    (3, 2, 2, 1), (3, 2, 2, 2), (3, 2, 2, 3), (3, 2, 2, 4),   # d <= 48 (SD-1.5 d = 40): three score buffers
    (2, 2, 2, 1), (2, 2, 2, 2), (2, 2, 2, 3), (2, 2, 2, 4),   # d = 56 .. 112: one score buffer per tile
    (3, 3, 2, 4), (3, 4, 2, 6), (3, 4, 2, 9),                  # ring depths incl. the kernel's four stages; longer runs
    (3, 2, 4, 2), (2, 2, 4, 2), (3, 2, 3, 4),                  # the kernel's four warps per tile; three with rescales
])
def test_every_interleaving_hands_over_complete_data(NBUF, S, W, nkv):
    n = explore(nkv, NBUF, S, W, allow_rescale=(W < 4 and nkv < 9))
    assert n > 100


@pytest.mark.skipif(os.environ.get("PP_SLOW_TESTS") != "1", reason="~2 min: set PP_SLOW_TESTS=1")
def test_the_kernels_own_configuration():
    """attn2_kernel<3,1,4> as it is launched: three score buffers, four ring stages, four warps per tile (1.4 M states),
    and four key blocks on a two-stage ring (1.1 M states)"""
    assert explore(3, 3, 4, 4, allow_rescale=False, max_states=20_000_000) > 1_000_000
    assert explore(4, 3, 2, 4, allow_rescale=False, max_states=20_000_000) > 1_000_000


def test_the_unguarded_epilogue_aliases():
    """the protocol as measured in round 2 (no wait on the last block): the epilogue's parity wait passes two phases
    early when a warp runs a block ahead at the end — three score buffers only"""
    for nkv in (2, 3, 4):
        with pytest.raises(Violation, match=f"epilogue of tile \\d reads O after {nkv - 2} of {nkv} PV products"):
            explore(nkv, 3, 2, 2, final_guard=False, allow_rescale=False)
        explore(nkv, 2, 2, 2, final_guard=False)  # two score buffers: S(k + 2) needs PV(k), nobody runs ahead
    explore(1, 3, 2, 2, final_guard=False)        # a single key block (cross-attention over 77 tokens) has no alias


def test_the_round_2_race_is_visible_to_the_model():
    """one P hand-over barrier per tile (the design at the start of round 2) with three score buffers: a warp that runs a
    key block ahead completes the barrier's phase for a slower warp -> PV over raw scores. Two score buffers cannot run
    ahead (S(k + 2) needs PV(k)), which is why only the d <= 48 instantiation was affected."""
    with pytest.raises(Violation, match="P\\(\\d+\\) is incomplete"):
        explore(3, 3, 2, 2, single_p_barrier=True, allow_rescale=False)
    explore(3, 2, 2, 2, single_p_barrier=True, allow_rescale=False)  # NBUF = 2: safe even with one barrier


def test_an_early_ring_release_is_visible_to_the_model():
    """kv_empty committed after tile 0's PV(j) instead of tile 1's: the producer may refill the stage under PV(2j + 1)"""
    with pytest.raises(Violation, match="stage"):
        explore(4, 3, 2, 2, kv_release_tile=0, allow_rescale=False)


def test_model_matches_the_source():
    """the waits / commits the model restates, verbatim in attention2.cuh"""
    with open(SRC) as f:
        src = f.read()
    for line in [
        "auto bar_p_full = [&](int t, int j) { return bars + 8u * (7 + 2 * MAXS + 2 * t + (j & 1)); };",
        "mbar_init(bar_p_full(t, 0), 4);",
        "mbar_init(bar_p_full(t, 1), 4);",
        "mbar_wait(bar_kv_empty(s), ph ^ 1u);",
        "mbar_wait(bar_kv_full(j % S), (j / S) & 1);",
        "umma_commit(bar_s_full(buf));",
        "umma_commit(bar_pv_done(t));",
        "for (int k = 0; k < NBUF && k < nprod; ++k) issue_s(k);",
        "mbar_wait(bar_p_full(t, j), (j >> 1) & 1);",
        "if (t == 1) umma_commit(bar_kv_empty(j % S));",
        "if (k + NBUF < nprod) {",
        "mbar_wait(bar_pv_done(t), j & 1);",
        "issue_s(k + NBUF);",
        "mbar_wait(bar_s_full(buf), (k / NBUF) & 1);",
        "mbar_wait(bar_pv_done(t), (j - 1) & 1);",
        "if constexpr (NBUF == 3 && ATT2_FINAL_GUARD != 0) {\n                if (j == nkv - 1 && j > 0) {\n                    mbar_wait("
        "bar_pv_done(t), "
        "(j - 1) & 1);",
        "#define ATT2_FINAL_GUARD 1",
        "if (lane_id() == 0) mbar_arrive(bar_p_full(t, j));",
        "mbar_wait(bar_pv_done(t), (nkv - 1) & 1);",
    ]:
        assert line in src, f"attention2.cuh no longer contains `{line}`: update the protocol model"
