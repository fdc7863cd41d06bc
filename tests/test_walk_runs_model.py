"""The table walk by runs (motes_b200/csrc/walktab.cuh: walk_runs_kernel), restated in numpy and checked against the
literal column-by-column walk of numpy's masked-rejection draws (sky.py:34-38 -> np.random.choice -> one generator
stream per frame).  What the kernel relies on: consecutive columns that draw with the same number n of good sky
pixels share their acceptance test, so the k-th of them ends at the (k + 1) * 100 * n-th accepted word behind the run's
start - positions from cumulative accept counts, no walk inside the run - and the draws the running column has had at
every 2048-word piece boundary follow from the same counts.  The CUDA kernel itself is checked bit for bit against the
literal replay on the GPU (tests/test_gpu_fullsize.py); this test pins the algorithm.
"""
import numpy as np
import pytest

PIECE = 2048


def literal_walk(words, mask, ns, p0):
    """Column after column: position of the first draw, position behind the last accepted draw."""
    pos, starts, ends = p0, [], []
    for n in ns:
        starts.append(pos)
        if n > 1:
            need, have = 100 * n, 0
            while have < need:
                have += (words[pos] & mask) < n
                pos += 1
        ends.append(pos)
    return np.array(starts), np.array(ends)


def runs_walk(words, mask, ns, p0):
    """Runs of equal n (columns with n <= 1 draw nothing and belong to whatever run surrounds them)."""
    starts, ends = np.zeros(len(ns), dtype=np.int64), np.zeros(len(ns), dtype=np.int64)
    subs = {}
    pos, c = p0, 0
    while c < len(ns):
        window = ns[c:c + 32]
        drawing = [i for i, n in enumerate(window) if n > 1]
        if not drawing:
            starts[c:c + len(window)] = ends[c:c + len(window)] = pos
            c += len(window)
            continue
        n = window[drawing[0]]
        other = [i for i in drawing if window[i] != n]
        length = other[0] if other else len(window)
        members = [i for i in drawing if i < length]
        need, k_total = 100 * n, len(members)
        piece0 = pos // PIECE
        accept = (words[piece0 * PIECE:] & mask) < n
        cum = np.cumsum(accept)                                     # accepted in [piece0 * PIECE, position]
        a = int(cum[pos - piece0 * PIECE - 1]) if pos > piece0 * PIECE else 0
        targets = a + need * np.arange(1, k_total + 1)
        t_pos = piece0 * PIECE + np.searchsorted(cum, targets, side="left") + 1      # behind the target-th accepted word
        # draws of the running column at the piece boundaries inside the run
        last_piece = (int(t_pos[-1]) - 1) // PIECE
        for piece in range(piece0 + 1, last_piece + 1):
            s_lo = int(cum[(piece - piece0) * PIECE - 1])
            done = min(k_total, (s_lo - a) // need)
            if done < k_total:
                subs[piece] = s_lo - a - done * need
        before = 0
        for i in range(length):
            start = pos if before == 0 else int(t_pos[before - 1])
            starts[c + i] = start
            if i in members:
                ends[c + i] = int(t_pos[before])
                before += 1
            else:
                ends[c + i] = start
        pos = int(t_pos[-1])
        c += length
    return starts, ends, subs


@pytest.mark.parametrize("seed,nspat", [(1, 80), (2, 200), (3, 400)])
def test_runs_walk_equals_literal_walk(seed, nspat):
    rng = np.random.default_rng(seed)
    mask = (1 << int(np.ceil(np.log2(nspat)))) - 1
    base = int(0.8 * nspat)
    ns = []
    while len(ns) < 150:                                            # runs of random length, some columns without a sky model
        n = max(mask // 2 + 2, base + int(rng.integers(-3, 4)))     # one rejection mask for the frame
        ns += [n] * int(rng.integers(1, 45))
    ns = np.array(ns[:150])
    ns[rng.random(150) < 0.05] = 0
    words = rng.integers(0, 1 << 16, size=150 * 100 * (mask + 1) * 2 + 100000, dtype=np.uint16).astype(np.int64)
    p0 = int(rng.integers(0, 5000))
    want_s, want_e = literal_walk(words, mask, ns, p0)
    got_s, got_e, subs = runs_walk(words, mask, ns, p0)
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_e, want_e)
    # the running count at a piece boundary is read by the column that holds words on both sides of it (a column that
    # ends exactly at the boundary never asks, the next one starts there and counts from zero)
    checked = 0
    for piece, count in subs.items():
        sb = piece * PIECE
        inside = np.nonzero((want_s < sb) & (want_e > sb))[0]
        if len(inside) == 0:
            continue
        col = int(inside[0])
        assert count == int(np.sum((words[want_s[col]:sb] & mask) < ns[col])), piece
        checked += 1
    assert checked > 100
