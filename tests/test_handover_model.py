"""Model check of the hand-over between consecutive evaluations inside ONE drive-mode launch
(logreg_b200/csrc/eval_persist_kernel.cuh): running sums accumulated with atomics, an atomic ticket
electing the last CTA, read-and-zero of the sums, reset of ticket and batch counter, epoch flag.
Sequentially consistent Python restatement under random interleavings of the CTAs; checks that the
last CTA of evaluation e sees exactly the contributions of evaluation e (none missing, none of e+1),
that exactly one CTA is elected per evaluation and that the counters are at rest at the end."""
import random


def cta(b, G, n_evals, sh, log):
    for e in range(n_evals):
        while sh["epoch"] < e:                       # wait_epoch (ld.acquire.gpu)
            yield
        beta = sh["beta"]                            # the point published by the previous last CTA
        assert beta == e, (b, e, beta)
        sh["work"][e & 1] += 1                       # (claims of this evaluation)
        yield
        sh["acc"] += (b + 1) * 10 ** 3 + beta        # red.global.add.f64 of this CTA's sums
        yield
        t = sh["ticket"]                             # atomicAdd(ticket, 1) after fence + barrier
        sh["ticket"] = t + 1
        yield
        if t != G - 1:
            continue                                 # on to the next evaluation (prefetch, then wait)
        tot, sh["acc"] = sh["acc"], 0                # atomicExch(acc, 0)
        sh["ticket"] = 0
        sh["work"][e & 1] = 0
        yield
        assert tot == sum((c + 1) * 10 ** 3 + e for c in range(G)), (e, tot)
        log.append((e, b))
        sh["beta"] = e + 1                           # finish_eval: sampler update -> next evaluation point
        yield
        sh["epoch"] = e + 1                          # st.release.gpu


def test_handover_between_evaluations():
    rng = random.Random(31337)
    for trial in range(200):
        G = rng.choice([1, 2, 3, 7, 16])
        n_evals = rng.choice([1, 2, 5, 9])
        sh = {"epoch": 0, "beta": 0, "acc": 0, "ticket": 0, "work": [0, 0]}
        log = []
        procs = {b: cta(b, G, n_evals, sh, log) for b in range(G)}
        weights = {b: rng.choice([1.0, 0.1, 5.0]) for b in range(G)}
        steps = 0
        while procs:
            live = list(procs)
            b = rng.choices(live, weights=[weights[x] for x in live])[0]
            try:
                next(procs[b])
            except StopIteration:
                del procs[b]
            steps += 1
            assert steps < 2_000_000
        assert [e for e, _ in log] == list(range(n_evals))          # exactly one last CTA per evaluation, in order
        assert sh["epoch"] == n_evals and sh["acc"] == 0 and sh["ticket"] == 0 and sh["work"] == [0, 0]
