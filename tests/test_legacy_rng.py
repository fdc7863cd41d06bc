"""oracle/legacy_rng.py must reproduce numpy's legacy RandomState stream as sky.py consumes it."""
import numpy as np
import pytest

from oracle.legacy_rng import MT19937, bounded_indices, bootstrap_median_replay, LegacyGauss, rejection_mask


@pytest.mark.parametrize("seed", [0, 1, 1234, 2**32 - 1])
def test_raw_stream(seed):
    ref = np.random.RandomState(seed)._bit_generator.random_raw(3000).astype(np.uint32)
    assert np.array_equal(MT19937(seed).raw(3000), ref)


@pytest.mark.parametrize("n", [1, 2, 3, 64, 65, 170, 372, 512, 513])
def test_choice_indices(n):
    seed = 77 + n
    np.random.seed(seed)
    ref = np.random.choice(np.arange(n), (n, 100), replace=True)
    got = bounded_indices(MT19937(seed), n, n * 100).reshape(n, 100)
    assert np.array_equal(ref, got)


def test_masks():
    assert [rejection_mask(n) for n in (2, 3, 4, 5, 256, 257, 512)] == [1, 3, 3, 7, 255, 511, 511]


def test_bootstrap_median_value_and_stream_position():
    rng = np.random.default_rng(5)
    good = rng.normal(100.0, 7.0, size=173)
    np.random.seed(99)
    draws = np.random.choice(good, (len(good), 100), replace=True)
    meds = np.nanmedian(draws, axis=0)
    ref = (np.nanmean(meds), np.std(meds) / (99 ** 0.5))
    after = np.random.randint(0, 2**31 - 1)
    gen = MT19937(99)
    got = bootstrap_median_replay(gen, good)
    assert ref == got
    # both generators sit at the same place in the stream afterwards
    np.random.seed(99)
    np.random.choice(good, (len(good), 100), replace=True)
    assert np.random.get_state()[2] % 624 == gen.pos % 624


def test_legacy_gauss():
    np.random.seed(4242)
    ref = np.concatenate([np.random.normal(loc=3.0, scale=2.5, size=100), np.random.normal(loc=-1.0, scale=0.5, size=100)])
    g = LegacyGauss(MT19937(4242))
    got = np.concatenate([g.normal(3.0, 2.5, 100), g.normal(-1.0, 0.5, 100)])
    assert np.array_equal(ref, got)


def test_jump_polynomials_reach_the_same_generator_window():
    """The jump-ahead table the segmented bootstrap uploads (csrc/mtjump.hpp: characteristic polynomial by
    Berlekamp-Massey, x^(2^s) by repeated squaring): jumping 2^s words equals running the recurrence 2^s words."""
    import ctypes, os, subprocess
    from conftest import ROOT
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    lib = ctypes.CDLL(os.path.join(d, "libmotes_hostsim.so"))
    lib.hostsim_jump_mismatches.argtypes = [ctypes.c_uint32, ctypes.c_int]
    for seed, s in ((1234, 0), (1, 3), (5489, 10), (2024, 16), (7, 18), (4242, 20)):
        assert lib.hostsim_jump_mismatches(seed, s) == 0, (seed, s)
