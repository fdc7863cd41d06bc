"""The bootstrap column kernel's medians (motes_b200/csrc/skyseg.cuh: sky_column_hist32_kernel), restated in numpy.

A column's 100 n draws are indices into its good sky pixels; draw k belongs to sample k % 100 (np.random.choice fills
its (n, 100) result row by row, sky.py:34-38).  Per sample the kernel only keeps a histogram of the drawn pixels' RANKS
inside a window of 128 ranks around n / 2 plus the number of draws below and above it; the median is read off a running
sum: the bin holding the sample's t-th draw inside the window is the number of bins whose running count stays <= t.
A lane's accepted draws go to consecutive samples without a wrap test: rows 100 .. 106 take what passes sample 99 and
are added to rows 0 .. 6 afterwards.  This pins that arithmetic against np.nanmedian of the sampled values.
"""
import numpy as np
import pytest

BINS, ROWS = 128, 107


@pytest.mark.parametrize("n,seed", [(2, 0), (3, 1), (97, 2), (256, 3), (323, 4), (400, 5), (512, 6)])
def test_window_histogram_medians_equal_numpy(n, seed):
    rng = np.random.default_rng(seed)
    pixels = rng.normal(100.0, 5.0, size=n)
    pixels[rng.random(n) < 0.2] = np.round(pixels.mean())                            # some equal values
    order = np.argsort(pixels, kind="stable")
    srt = pixels[order]
    rank_of = np.empty(n, dtype=np.int64)
    rank_of[order] = np.arange(n)
    draws = rng.integers(0, n, size=100 * n)
    want = np.nanmedian(pixels[draws].reshape(n, 100), axis=0)                        # sky.py:36-39

    lo = max(0, n // 2 - BINS // 2)
    hist = np.zeros((ROWS, BINS + 2), dtype=np.int64)
    # groups of up to 8 consecutive draws start at sample (ordinal % 100) and run on without wrapping
    k = 0
    while k < len(draws):
        run = int(rng.integers(1, 9))
        smp = k % 100
        for d in draws[k:k + run]:
            r = rank_of[d] - lo
            hist[smp, r if 0 <= r < BINS else (BINS if r < 0 else BINS + 1)] += 1
            smp += 1
        k += run
    hist[:7] += hist[100:107]
    k1, k2 = (n - 1) // 2, n // 2
    got = np.empty(100)
    for s in range(100):
        nb, na = hist[s, BINS], hist[s, BINS + 1]
        cum = np.cumsum(hist[s, :BINS])
        assert nb + cum[-1] + na == n
        if not (k1 >= nb and k2 < nb + cum[-1]):
            pytest.skip("median outside the window: the kernel hands such a column to the exact one")
        x1 = int(np.sum(cum <= k1 - nb))
        x2 = int(np.sum(cum <= k2 - nb))
        got[s] = srt[lo + x1] if k1 == k2 else (srt[lo + x1] + srt[lo + x2]) * 0.5
    assert np.array_equal(got, want)
