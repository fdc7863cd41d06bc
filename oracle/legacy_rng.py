"""Restatement of numpy's legacy global RNG as the reference's sky bootstrap consumes it.  TEST INFRASTRUCTURE ONLY.

The reference draws its sky bootstrap from ``np.random.choice`` and
``np.random.normal`` on the process-global legacy ``RandomState``
(``/root/reference/src/motes/sky.py:34-38,219-223``) and never seeds it; the
parity harness calls ``np.random.seed(k)`` before each frame.  numpy is a
third-party dependency of the reference (``/root/reference/poetry.lock:417-418``
pins 2.2.0; this image carries 2.3.5).  What is restated here is the published
MT19937 generator (Matsumoto & Nishimura 1998, ``init_genrand`` seeding) and the
three consumption rules of numpy's legacy distributions:

* ``randint``-style bounded integers by masked rejection on 32-bit outputs
  (used by ``choice``);
* 53-bit doubles ``(a >> 5) * 2**26 + (b >> 6)) / 2**53`` from two outputs;
* the polar Box-Muller ``legacy_gauss`` that returns ``f*x2`` and caches ``f*x1``.

``tests/test_legacy_rng.py`` pins every function against ``np.random.RandomState``.
The CUDA sky kernel (``motes_b200/csrc/sky.cuh``, ``mt19937.cuh``) implements the
same stream on the device.
"""

from __future__ import annotations

import math

import numpy as np

N, M = 624, 397
UPPER, LOWER, MATRIX_A = 0x80000000, 0x7FFFFFFF, 0x9908B0DF


class MT19937:
    """Plain MT19937 with numpy's ``seed(int)`` initialisation; ``raw(k)`` returns k tempered 32-bit outputs."""

    def __init__(self, seed: int):
        mt = np.zeros(N, dtype=np.uint64)
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, N):
            prev = int(mt[i - 1])
            mt[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xFFFFFFFF
        self.mt = mt.astype(np.uint32)
        self.pos = N

    def _twist(self):
        mt = self.mt
        # three dependency-free sweeps: new[i] needs old[i], old[i+1] and (old or new)[i+397 mod 624]
        for first, last in ((0, N - M), (N - M, 2 * (N - M)), (2 * (N - M), N - 1)):
            i = np.arange(first, last)
            y = (mt[i] & np.uint32(UPPER)) | (mt[i + 1] & np.uint32(LOWER))
            mag = np.where(y & np.uint32(1), np.uint32(MATRIX_A), np.uint32(0))
            mt[i] = mt[(i + M) % N] ^ (y >> np.uint32(1)) ^ mag
        y = (mt[N - 1] & np.uint32(UPPER)) | (mt[0] & np.uint32(LOWER))
        mag = np.uint32(MATRIX_A) if (y & np.uint32(1)) else np.uint32(0)
        mt[N - 1] = mt[M - 1] ^ (y >> np.uint32(1)) ^ mag
        self.pos = 0

    @staticmethod
    def _temper(y):
        y = y ^ (y >> np.uint32(11))
        y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
        y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
        return y ^ (y >> np.uint32(18))

    def raw(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=np.uint32)
        filled = 0
        while filled < count:
            if self.pos >= N:
                self._twist()
            take = min(N - self.pos, count - filled)
            out[filled:filled + take] = self._temper(self.mt[self.pos:self.pos + take])
            self.pos += take
            filled += take
        return out


def rejection_mask(n: int) -> int:
    """Smallest 2**k - 1 >= n - 1 (numpy ``_bounded_uint32`` mask for range n-1)."""
    rng = n - 1
    mask = rng
    for s in (1, 2, 4, 8, 16):
        mask |= mask >> s
    return mask


def bounded_indices(gen: MT19937, n: int, count: int) -> np.ndarray:
    """``count`` integers in [0, n) as ``np.random.choice``/``randint`` draws them: one 32-bit output per attempt, keep ``(u & mask) <= n-1``.  n == 1 consumes nothing."""
    if n == 1:
        return np.zeros(count, dtype=np.int64)
    mask = np.uint32(rejection_mask(n))
    out = np.empty(count, dtype=np.int64)
    filled = 0
    while filled < count:
        u = gen.raw(1)[0] & mask
        if u <= n - 1:
            out[filled] = u
            filled += 1
    return out


def bootstrap_median_replay(gen: MT19937, good: np.ndarray):
    """The value ``sky.sky_median`` returns when the global RNG is at ``gen``'s position (sky.py:17-43)."""
    n = len(good)
    idx = bounded_indices(gen, n, n * 100).reshape(n, 100)     # row-major fill: pixel major, sample minor
    meds = np.nanmedian(good[idx], axis=0)
    return np.nanmean(meds), np.std(meds) / (99 ** 0.5)


def uniform53(gen: MT19937) -> float:
    a, b = (int(v) for v in gen.raw(2))
    return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0


class LegacyGauss:
    """numpy ``legacy_gauss``: polar method, second deviate cached across calls."""

    def __init__(self, gen: MT19937):
        self.gen = gen
        self.cached = None

    def draw(self) -> float:
        if self.cached is not None:
            g, self.cached = self.cached, None
            return g
        while True:
            x1 = 2.0 * uniform53(self.gen) - 1.0
            x2 = 2.0 * uniform53(self.gen) - 1.0
            r2 = x1 * x1 + x2 * x2
            if 0.0 < r2 < 1.0:
                break
        f = math.sqrt(-2.0 * math.log(r2) / r2)        # C libm, as numpy's legacy_gauss
        self.cached = f * x1
        return f * x2

    def normal(self, loc, scale, size):
        return np.array([loc + scale * self.draw() for _ in range(size)])
