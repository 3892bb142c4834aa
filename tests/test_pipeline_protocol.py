"""Flow control of the pipelined sharded search (qb_comm.cu, qb_comm_pipelined_step), model-checked on the CPU.

Per rank two in-order streams: the SCAN stream runs scan(i) — which writes this shard's lists into LOCAL ring slot i mod R — once the
rank's own merge(i - W) has completed; the EXCHANGE stream runs, per step, push(i) — after scan(i): copies local slot i mod R into REMOTE
slot i mod R of every peer, then the flag — and merge(i), which waits for every peer's flag of step i and reads slot i mod R of its own
buffer.  No other synchronisation exists between ranks.  Two hazards: a scan overwriting a local slot that has not been pushed yet (the
window W bounds how far the scan stream runs ahead: R >= W), and a push overwriting a remote slot its owner has not merged yet (the
in-order exchange streams keep ranks within one step of each other: R >= 2).  The library uses W = 2, R = 4.  This test runs the protocol
under random (adversarial) schedules, checks safety and liveness, and shows that the checker does find the overwrite when the ring is
smaller than the window."""
import random

import pytest


def run(world, steps, W, R, seed, bias=None):
    rng = random.Random(seed)
    scan_done = [0] * world            # steps whose scan has completed, per rank
    pushed = [0] * world               # steps pushed, per rank
    merged = [0] * world               # steps merged, per rank
    # slot[y][x][k] = step number currently stored in rank y's buffer, written by x, ring position k
    slot = [[[-1] * R for _ in range(world)] for _ in range(world)]
    local = [[-1] * R for _ in range(world)]        # local[r][k] = step whose lists sit in rank r's own ring slot k
    while min(merged) < steps:
        actions = []
        for r in range(world):
            i = scan_done[r]
            if i < steps and (i < W or merged[r] >= i - W + 1):       # scan(i) waits for this rank's merge(i - W)
                actions.append(("scan", r))
            if pushed[r] < scan_done[r] and pushed[r] == merged[r]:    # exchange stream in order: push(i) follows merge(i - 1)
                actions.append(("push", r))
            j = merged[r]
            if j < pushed[r] and all(slot[r][x][j % R] >= j for x in range(world)):   # every peer's lists of step j (or later!) have landed
                actions.append(("merge", r))
        assert actions, f"deadlock: scan_done={scan_done} pushed={pushed} merged={merged}"
        if bias is not None and rng.random() < 0.7:
            fav = [a for a in actions if a[1] == bias]                 # let one rank race ahead as far as the protocol allows
            actions = fav or actions
        kind, r = rng.choice(actions)
        if kind == "scan":
            local[r][scan_done[r] % R] = scan_done[r]
            scan_done[r] += 1
        elif kind == "push":
            i = pushed[r]
            if local[r][i % R] != i:
                return f"rank {r} pushes step {i} but its local slot holds step {local[r][i % R]}"
            for y in range(world):
                slot[y][r][i % R] = i
            pushed[r] += 1
        else:
            j = merged[r]
            for x in range(world):
                if slot[r][x][j % R] != j:
                    return f"rank {r} merges step {j} but rank {x}'s slot holds step {slot[r][x][j % R]}"
            merged[r] += 1
    return None


@pytest.mark.parametrize("world", [2, 3, 8])
def test_window_2_ring_4_is_safe_and_live(world):
    for seed in range(60):
        assert run(world, steps=40, W=2, R=4, seed=seed) is None
        assert run(world, steps=40, W=2, R=4, seed=seed, bias=seed % world) is None


def test_the_checker_finds_the_overwrite_when_the_ring_is_too_small():
    found = [run(3, steps=40, W=3, R=2, seed=s, bias=s % 3) for s in range(200)]
    assert any(f is not None for f in found), "a window of 3 steps over a ring of 2 should be caught"
    assert all(run(3, steps=40, W=2, R=2, seed=s, bias=s % 3) is None for s in range(100))     # R = W is enough
    # a barrier per step (W = 1) with two slots: the non-pipelined exchange
    assert all(run(3, steps=30, W=1, R=2, seed=s, bias=s % 3) is None for s in range(60))
