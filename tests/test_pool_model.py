"""Model check of the drive-mode batch pool (logreg_b200/csrc/eval_persist_kernel.cuh, `next_dynamic`
/ `producer_visit`): a step-by-step Python restatement of the protocol, run under thousands of
random interleavings of the warps of several CTAs.

Protocol being modelled (per CTA, shared memory): `claims` counts draws; draw c is slot c % 8 of chunk
c // 8; chunk k's first batch (relative to the dynamic region) is published in base[k % 16] with tag k+1
by ONE producer (lane 0 of warp 0), which issues every global claim of the CTA (so the bases it
publishes are monotone), keeps the pool up to 4 chunks ahead of the draws, publishes a claim one
visit after issuing it and marks the first exhausted chunk in `last`.  A consumer spins until its
chunk is tagged or `last` says the pool ended; warp 0 keeps producing while it spins.

Checked: every batch of the dynamic region is handed out exactly once across all CTAs, every warp
terminates, and no consumer ever reads a slot that has been recycled for a later chunk (with 8 slots
an adversarial schedule could do that; with 16 the producer's 4-chunk look-ahead leaves 12 chunks of
slack and 6000 adversarial interleavings never got there).

History: this model was written after the GPU budget of round 2 was spent, to understand why drive
mode + the peer-memory exchange had stopped at N = 4 / 8.  It found a deadlock in the protocol as it
then was (test_the_model_finds_the_deadlock_of_a_producer_that_leaves_early); the kernel now keeps
the producer in the loop until the end of the pool is published."""
import random

CHUNK, SLOTS, NONE = 8, 16, 0xFFFFFFFF


class Cta:
    def __init__(self):
        self.claims, self.last = 0, NONE
        self.base, self.tag = [0] * SLOTS, [0] * SLOTS
        self.issued, self.pend = 0, []            # producer-private (registers of warp 0, lane 0)


def producer_visit(cta, glob, span):
    for chunk, base in cta.pend:                  # publish what the previous visit issued
        assert cta.tag[chunk % SLOTS] <= chunk + 1
        cta.base[chunk % SLOTS] = base
        cta.tag[chunk % SLOTS] = chunk + 1
        if base >= span and cta.last == NONE:
            cta.last = chunk
    cta.pend = []
    target = 0 if cta.last != NONE else cta.claims // CHUNK + 4
    while cta.issued < target and len(cta.pend) < 2:
        cta.pend.append((cta.issued, glob[0]))    # atomicAdd(work, 8)
        glob[0] += CHUNK
        cta.issued += 1


def warp(cta, w, glob, span, ks_static, out, producer_stays=True):
    """One warp of the streaming loop as a generator: each `yield` is a point where another warp may
    run.  Static batches are irrelevant to the pool; only their count (how late the warp starts
    drawing) matters.  producer_stays: the producer does not leave the loop before the end of the
    pool is published (the rule whose absence this model found: see the test below)."""
    if w == 0:
        producer_visit(cta, glob, span)
    yield
    for _ in range(ks_static):                    # static prefix: warp 0 visits the producer every iteration
        if w == 0:
            producer_visit(cta, glob, span)
        yield
    while True:
        if w == 0:
            producer_visit(cta, glob, span)
        yield
        c = cta.claims                            # atomicAdd(&s_claims, 1)
        cta.claims += 1
        chunk, slot = c // CHUNK, c % CHUNK
        yield
        base = NONE
        spins = 0
        while True:
            if cta.tag[chunk % SLOTS] == chunk + 1:
                base = cta.base[chunk % SLOTS]
                break
            assert cta.tag[chunk % SLOTS] < chunk + 1, "slot recycled before its chunk was read"
            if chunk >= cta.last:
                break
            if w == 0:
                producer_visit(cta, glob, span)
            spins += 1
            assert spins < 100_000, "consumer starved"
            yield
        if base >= span or base + slot >= span:
            # `cur >= nbatch`: the warp leaves the streaming loop (bases are monotone per CTA, so
            # everything this CTA draws later is exhausted as well).  The producer may only leave once
            # the first exhausted chunk is published -- it can get here through an out-of-range slot
            # of the chunk that straddles the end of X, with the end of the pool still in its registers.
            if w == 0 and producer_stays:
                visits = 0
                while cta.last == NONE:
                    producer_visit(cta, glob, span)
                    visits += 1
                    assert visits <= 8, "the end of the pool is at most a few visits away"
                    yield
            return
        out.append(base + slot)                   # "process" the batch (the claim came early: nothing else to model)
        yield


def run_once(rng, nctas, nwarps, span, ks_static, bias, producer_stays=True):
    glob = [0]
    ctas = [Cta() for _ in range(nctas)]
    out = []
    gens = [(ci, w, warp(ctas[ci], w, glob, span, ks_static, out, producer_stays))
            for ci in range(nctas) for w in range(nwarps)]
    weights = [bias if w == 0 else 1.0 for _, w, _ in gens]      # bias < 1: a slow producer warp
    steps = 0
    while gens:
        i = rng.choices(range(len(gens)), weights=weights)[0]
        try:
            next(gens[i][2])
        except StopIteration:
            gens.pop(i)
            weights.pop(i)
        steps += 1
        assert steps < 5_000_000, "no termination"
    return sorted(out)


def test_every_dynamic_batch_is_handed_out_exactly_once():
    rng = random.Random(12345)
    for trial in range(300):
        nctas = rng.choice([1, 2, 3, 5])
        nwarps = rng.choice([2, 4, 8])
        span = rng.choice([0, 1, 7, 8, 9, 63, 64, 200, 333])
        ks = rng.choice([0, 0, 1, 3])
        bias = rng.choice([1.0, 0.2, 5.0])
        out = run_once(rng, nctas, nwarps, span, ks, bias)
        assert out == list(range(span)), (trial, nctas, nwarps, span, ks, bias, out[:20])


def test_an_extremely_slow_producer_only_delays_the_cta():
    rng = random.Random(7)
    for _ in range(20):
        out = run_once(rng, 2, 8, 500, 0, 0.02)
        assert out == list(range(500))


def test_the_model_finds_the_deadlock_of_a_producer_that_leaves_early():
    """Round-2 finding: when the dynamic region is not a multiple of 8 batches, one chunk straddles the
    end of X; if the producer warp draws one of its out-of-range slots it used to leave the loop with
    the (exhausted) next chunk unpublished, and the other warps of that CTA waited for it forever.
    Without the `producer_stays` rule the model starves a consumer within a few hundred interleavings."""
    import pytest
    rng = random.Random(99)
    with pytest.raises(AssertionError, match="starved"):
        for _ in range(400):
            run_once(rng, rng.choice([1, 2, 3]), rng.choice([4, 8]), rng.choice([1, 9, 65, 201, 333]), 0,
                     rng.choice([1.0, 0.2]), producer_stays=False)
