"""CPU model of the filter path's ALGORITHM (tensorrec_b200/csrc/score_filter_tc.cu + rescore_topk.cu): block-bound
admission, 32-entry buffers, compaction to at most 16 entries >= (k-th best - 2.25 m), overflow tracking, exact
re-scoring and the certificate.  It checks the claim DESIGN.md section 4 makes about the kernel pair, independently of
any GPU: whenever the certificate accepts a row, the reported top-k equals the exact top-k in tf.nn.top_k order --
for adversarial inputs too (massive ties, approximation errors right at the bound, any processing order)."""
import numpy as np
import pytest

BUF, KEEP, MARGINS = 32, 16, 2.25


def run_row(exact, approx, bias_order_pos, block_max_of, m, k, block=128, step=16, theta_init=-np.inf, raw=False):
    """One user row.  exact/approx: per-item scores (approx within m of exact); processing order = bias_order_pos.
    block_max_of(pos) -> admission slack of the block (>= 0): the bound test admits a superset of approx > tau.
    theta_init: a threshold the row starts from (section "shared thresholds" below); raw=True returns the survivors
    and the row's final threshold instead of the certified top-k."""
    n = len(exact)
    theta, tau, drop_max = theta_init, theta_init, -np.inf
    buf = []                                           # [(approx score, item id)]

    def compact():
        nonlocal buf, theta, tau, drop_max
        buf.sort(key=lambda e: (-e[0], e[1]))
        if len(buf) >= k:
            floor = buf[k - 1][0] - MARGINS * m
            kept = [e for e in buf if e[0] >= floor]
            if len(kept) > KEEP:
                drop_max = max(drop_max, kept[KEEP][0])
                kept = kept[:KEEP]
            floor = max(floor, theta)                  # (a carried threshold is never lowered)
            buf, theta, tau = kept, floor, floor       # (the kernel lowers tau by a few ulps: more admissions only)
        # fewer than k entries: nothing can be dropped yet

    for p0 in range(0, n, step):
        positions = range(p0, min(p0 + step, n))
        slack = block_max_of(p0 // block)
        passing = [p for p in positions if approx[bias_order_pos[p]] + slack > tau]
        if len(buf) + len(passing) > BUF:
            compact()
        buf.extend((approx[bias_order_pos[p]], bias_order_pos[p]) for p in passing)
        assert len(buf) <= BUF
    compact()
    row_theta = max(theta, drop_max)
    if raw:
        return [i for _, i in buf], row_theta, theta
    # rescore_topk_kernel: exact scores of the survivors, tf.nn.top_k order, certificate
    surv = sorted(((exact[i], i) for _, i in buf), key=lambda e: (-e[0], e[1]))
    top = surv[:k]
    certified = len(surv) >= k and (row_theta == -np.inf or row_theta + m < top[k - 1][0])
    return [i for _, i in top], certified


def exact_topk(exact, k):
    order = np.lexsort((np.arange(len(exact)), -np.asarray(exact, dtype=np.float64)))
    return list(order[:k])


@pytest.mark.parametrize('seed', range(12))
def test_certified_rows_equal_the_exact_topk(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(40, 900))
    k = int(rng.integers(1, 13))
    kind = seed % 4
    if kind == 0:                                       # continuous scores
        exact = rng.standard_normal(n)
    elif kind == 1:                                     # massive ties
        exact = rng.integers(-2, 3, n).astype(np.float64)
    elif kind == 2:                                     # a plateau exactly at the k-th place
        exact = np.concatenate([rng.standard_normal(n - 30) - 5.0, np.full(30, 1.0)])
    else:                                               # near-ties inside the error bound
        exact = 1.0 + 1e-4 * rng.standard_normal(n)
    m = 1e-3 if kind != 1 else 0.0
    noise = rng.uniform(-m, m, n) if m > 0 else np.zeros(n)
    if kind == 3:
        noise = np.where(rng.random(n) < 0.5, m, -m)   # errors right at the bound
    approx = exact + noise
    order = rng.permutation(n)                          # ANY processing order must be safe
    slack_per_block = rng.uniform(0.0, 0.01, n // 128 + 1)
    top, certified = run_row(exact, approx, order, lambda b: slack_per_block[b], m, k)
    if certified:
        assert top == exact_topk(exact, k)
    # rows that are not certified go to the exact kernel: nothing to check here except that the model says so for ties
    if kind == 1 and n > 5 * k:
        assert not certified or top == exact_topk(exact, k)


def test_clear_cut_rows_are_certified():
    """Well separated scores, honest error bound: the certificate must accept (the filter path has to be the common case)."""
    rng = np.random.default_rng(99)
    n, k, m = 2000, 10, 1e-4
    exact = rng.standard_normal(n)
    approx = exact + rng.uniform(-m, m, n)
    accepted = 0
    for trial in range(20):
        order = rng.permutation(n)
        top, certified = run_row(exact, approx, order, lambda b: 0.0, m, k)
        accepted += int(certified)
        if certified:
            assert top == exact_topk(exact, k)
    assert accepted >= 18


def test_overflow_of_near_ties_is_never_silently_wrong():
    """More than 16 items within the margin of the k-th best: entries are dropped, drop_max remembers the best of them
    and the certificate refuses the row."""
    n, k, m = 400, 10, 1e-3
    exact = np.full(n, 2.0)
    exact[:5] = 3.0                                      # 5 clear winners, then 395 exact ties for places 6..10
    approx = exact + np.linspace(-m, m, n)
    top, certified = run_row(exact, approx, np.arange(n), lambda b: 0.0, m, k)
    assert not certified


# ---- shared thresholds across item shards (the design the admission model argues for) ---------------
# A shard only has to keep the items that can reach the GLOBAL top-k.  A row may therefore start a shard's sweep from
# any theta for which k items with approximate score >= theta + 2.25 m exist SOMEWHERE (here: the threshold the row
# ended the previous shards with, carried round a ring).  The per-shard list is then no longer the shard's own top-k,
# so the certificate moves after the merge: exact k-th best of the union > max over shards of (theta_r, dropped_r) + m.
def run_sharded_row(exact, approx, shards, m, k, rng):
    theta, union, row_thetas = -np.inf, [], []
    for ids in shards:
        order = rng.permutation(len(ids))
        surv, row_theta, theta_out = run_row(exact[ids], approx[ids], order, lambda b: 0.0, m, k, theta_init=theta,
                                             raw=True)
        union += [int(ids[i]) for i in surv]
        row_thetas.append(row_theta)
        theta = max(theta, theta_out)
    merged = sorted(((exact[i], i) for i in union), key=lambda e: (-e[0], e[1]))[:k]
    bound = max(row_thetas)
    certified = len(merged) >= k and (bound == -np.inf or bound + m < merged[k - 1][0])
    return [i for _, i in merged], certified, len(union)


@pytest.mark.parametrize('seed', range(16))
def test_carried_thresholds_keep_the_certificate_sound(seed):
    rng = np.random.default_rng(1000 + seed)
    n, n_shards = int(rng.integers(600, 4000)), int(rng.integers(2, 9))
    k = int(rng.integers(1, 13))
    kind = seed % 4
    if kind == 0:
        exact = rng.standard_normal(n)
    elif kind == 1:
        exact = rng.integers(-2, 3, n).astype(np.float64)                     # massive ties
    elif kind == 2:
        exact = np.concatenate([rng.standard_normal(n - 30) - 5.0, np.full(30, 1.0)])
        exact = exact[rng.permutation(n)]                                     # a plateau at the k-th place, spread over shards
    else:
        exact = 1.0 + 1e-4 * rng.standard_normal(n)                           # near-ties inside the error bound
    m = 1e-3 if kind != 1 else 0.0
    approx = exact + (np.where(rng.random(n) < 0.5, m, -m) if kind == 3 else rng.uniform(-m, m, n) if m > 0 else 0.0)
    cuts = np.sort(rng.choice(np.arange(1, n), n_shards - 1, replace=False))
    shards = np.split(np.arange(n), cuts)
    top, certified, n_union = run_sharded_row(exact, approx, shards, m, k, rng)
    if certified:
        assert top == exact_topk(exact, k)


def test_carried_thresholds_shrink_the_lists_and_still_certify_clear_rows():
    rng = np.random.default_rng(5)
    n, k, m, n_shards = 16000, 10, 1e-4, 8
    exact = rng.standard_normal(n)
    approx = exact + rng.uniform(-m, m, n)
    shards = np.split(np.arange(n), n_shards)
    top, certified, n_union = run_sharded_row(exact, approx, shards, m, k, rng)
    assert certified and top == exact_topk(exact, k)
    # every shard on its own keeps between k and 16 survivors (8 k ... 128 in total); with a carried threshold a later
    # shard keeps only what beats the k-th best seen so far (theta alone is carried, not the lists: it rises only when a
    # shard finds k better items by itself, so this is an upper bound on what the design sends through the exchange)
    assert n_union < 8 * k
