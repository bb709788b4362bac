"""The (clip, step) item schedule of a multi-step launch (csrc/ls_fused.cu): item q = step * B + clip, CTA i takes
q = i, i + grid, ... in order, item (k, b) may start only when (k - 1, b) has been published.  A host model of
that protocol: with all CTAs co-resident (grid <= #SMs, which launch_fused guarantees) it always terminates, every
item runs exactly once and after its predecessor - for independent CTAs (the default build) for every B >= 1, for
cluster pairs sharing a weight ring (LS_MULTICAST=1) for B >= 2 (why lsf_steps runs B = 1 step by step there)."""
import itertools

import pytest


def simulate(B, K, grid, paired):
    """Round-robin 'hardware': a CTA makes progress when its current item's dependency is published; with paired
    CTAs both CTAs of a pair must be able to progress (they consume one weight stream in lock step)."""
    n_items = B * K
    n_rounds = -(-n_items // grid)
    cur = [0] * grid                      # round index per CTA
    done = set()
    order = []

    def item(c):
        q = c + cur[c] * grid
        return q if q < n_items else None

    def ready(c):
        if cur[c] >= n_rounds:
            return False
        q = item(c)
        return q is None or q < B or (q - B) in done      # invalid items recompute item 0: no dependency

    for _ in range(4 * n_rounds * grid + 8):
        progressed = False
        units = [(c, c + 1) for c in range(0, grid, 2)] if paired else [(c,) for c in range(grid)]
        for unit in units:
            if all(cur[c] < n_rounds for c in unit) and all(ready(c) for c in unit):
                for c in unit:
                    q = item(c)
                    if q is not None:
                        assert q not in done
                        done.add(q)
                        order.append(q)
                    cur[c] += 1
                progressed = True
        if all(r >= n_rounds for r in cur):
            break
        if not progressed:
            return None                   # deadlock
    assert len(done) == n_items
    pos = {q: i for i, q in enumerate(order)}
    assert all(pos[q - B] < pos[q] for q in order if q >= B)
    return order


@pytest.mark.parametrize("B,K", list(itertools.product([1, 2, 3, 5, 64, 147, 148, 149, 300, 512], [1, 2, 7, 16])))
def test_independent_ctas_always_finish(B, K):
    grid = min(B * K, 148)
    assert simulate(B, K, grid, paired=False) is not None


@pytest.mark.parametrize("B,K", list(itertools.product([2, 3, 5, 64, 147, 149, 300, 512], [1, 2, 7, 16])))
def test_cluster_pairs_finish_for_two_or_more_clips(B, K):
    grid = min(B * K + 1, 148) & ~1
    assert simulate(B, K, grid, paired=True) is not None


def test_cluster_pairs_deadlock_with_one_clip():
    """B = 1, K > 1: the two CTAs of a pair hold consecutive steps of the same clip; the second waits for the first,
    the first for its peer to drain the shared ring."""
    assert simulate(1, 4, 4, paired=True) is None
