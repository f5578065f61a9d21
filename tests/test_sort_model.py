"""CPU model of the step's counting sort by cell (csrc/sph_sort.cu + k_reorder<true>): whatever order
the atomics hand out inside a cell, ranking a cell's members by index gives exactly the stable sort by
key -- the order the radix passes produce and the oracle's definition of "sorted order"."""
import numpy as np
import pytest


def counting_sort_by_cell(keys, table_size, rng):
    n = len(keys)
    # count: the atomic's return value is the provisional rank; the order in which the particles of a
    # cell reach the atomic is arbitrary -- modelled by visiting the particles in a random order
    count = np.zeros(table_size + 1, np.int64)
    rank = np.zeros(n, np.int64)
    for i in rng.permutation(n):
        rank[i] = count[keys[i]]
        count[keys[i]] += 1
    # scan: cell_start = exclusive prefix
    cell_start = np.concatenate([[0], np.cumsum(count)[:-1]])
    # scatter: (key, index) -> cell_start[key] + provisional rank
    pairs = np.full(n, -1, np.int64)
    pairs[cell_start[keys] + rank] = np.arange(n)
    assert (pairs >= 0).all()
    # reorder: the thread of provisional slot s ranks its index among the cell's members
    order = np.full(n, -1, np.int64)
    for s in range(n):
        src = pairs[s]
        c0, c1 = cell_start[keys[src]], cell_start[keys[src] + 1]
        assert c0 <= s < c1
        order[c0 + np.count_nonzero(pairs[c0:c1] < src)] = src
    return order, cell_start


@pytest.mark.parametrize("n,table_size", [(1, 8), (500, 64), (3000, 27), (2000, 4096)])
def test_rank_by_index_inside_the_cell_is_a_stable_sort(n, table_size):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, table_size, n)
    order, cell_start = counting_sort_by_cell(keys, table_size, rng)
    np.testing.assert_array_equal(order, np.argsort(keys, kind="stable"))
    np.testing.assert_array_equal(cell_start, np.searchsorted(np.sort(keys), np.arange(table_size + 1)))


def test_warp_aggregated_count_matches_one_atomic_per_particle():
    """cell_rank(): neighbouring lanes with equal keys share one atomic of the run length; the ranks
    inside the run follow the lane order.  Same counts, a valid set of ranks per cell."""
    rng = np.random.default_rng(3)
    keys = np.sort(rng.integers(0, 40, 256))          # nearly sorted input: long runs
    keys[rng.integers(0, 256, 20)] = rng.integers(0, 40, 20)
    count = np.zeros(41, np.int64)
    rank = np.zeros(256, np.int64)
    for w in rng.permutation(8):                       # warps reach their atomics in any order
        lanes = np.arange(32 * w, 32 * w + 32)
        heads = [l for l in lanes if l == lanes[0] or keys[l] != keys[l - 1]]
        for h in heads:
            end = h + 1
            while end < lanes[-1] + 1 and keys[end] == keys[h]:
                end += 1
            base = count[keys[h]]
            count[keys[h]] += end - h
            rank[h:end] = base + np.arange(end - h)
    np.testing.assert_array_equal(count[:40], np.bincount(keys, minlength=40))
    for k in range(40):
        assert sorted(rank[keys == k]) == list(range(int(count[k])))
