"""Model check of the fused peer-memory allreduce (logreg_b200/csrc/sampler.cuh, `finish_eval`):
the two-slot mailbox protocol restated step by step in Python and run under random interleavings
of 2..8 ranks, some of them much slower than others.

Protocol: exchange k uses slot k & 1.  Rank r stores its sums into mailbox[peer][slot][r] of every
peer, then raises flag[peer][slot][r] to k + 1 (release), then waits until flag[r][slot][q] >= k + 1
for every peer q (acquire), then adds mailbox[r][slot][q] over q in rank order.  The claim in the
kernel's comment -- "two slots suffice: a peer cannot run two evaluations ahead because it needs my
sums of evaluation k to finish k" -- is what is checked: every value a rank adds is the one its peer
produced for THAT exchange, for every interleaving, so all ranks compute the same total."""
import random


def rank_proc(r, world, n_exchanges, mailbox, flags, totals):
    for k in range(n_exchanges):
        slot = k & 1
        mine = (r + 1) * 1000 + k                      # this rank's "sums" of exchange k
        for q in range(world):
            if q != r:
                mailbox[q][slot][r] = (k, mine)        # remote stores ...
                yield
        for q in range(world):
            if q != r:
                flags[q][slot][r] = k + 1              # ... then the flag (st.release.sys)
                yield
        for q in range(world):                         # one lane per peer spins (ld.acquire.sys)
            if q != r:
                while flags[r][slot][q] < k + 1:
                    yield
        tot = 0
        for q in range(world):                         # add in rank order
            if q == r:
                tot += mine
            else:
                kk, val = mailbox[r][slot][q]
                assert kk == k, f"rank {r} read rank {q}'s sums of exchange {kk} during exchange {k}"
                tot += val
            yield
        totals[r].append(tot)
        for _ in range(random.randrange(0, 3)):        # "streaming" of the next evaluation
            yield


def run(rng, world, n_exchanges, slow):
    mailbox = [[[None] * world for _ in range(2)] for _ in range(world)]
    flags = [[[0] * world for _ in range(2)] for _ in range(world)]
    totals = [[] for _ in range(world)]
    procs = [rank_proc(r, world, n_exchanges, mailbox, flags, totals) for r in range(world)]
    weights = [0.05 if r in slow else 1.0 for r in range(world)]
    alive = list(range(world))
    steps = 0
    while alive:
        r = rng.choices(alive, weights=[weights[a] for a in alive])[0]
        try:
            next(procs[r])
        except StopIteration:
            alive.remove(r)
        steps += 1
        assert steps < 10_000_000
    return totals


def test_two_slots_suffice_for_any_interleaving():
    rng = random.Random(4242)
    for trial in range(120):
        world = rng.choice([2, 3, 4, 8])
        slow = set(rng.sample(range(world), rng.randrange(0, world)))
        totals = run(rng, world, 12, slow)
        expect = [sum((r + 1) * 1000 + k for r in range(world)) for k in range(12)]
        for r in range(world):
            assert totals[r] == expect, (trial, world, r)
