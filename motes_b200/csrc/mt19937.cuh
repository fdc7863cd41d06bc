// MT19937 exactly as numpy's legacy RandomState drives it (np.random.seed / choice / normal,
// consumed by /root/reference/src/motes/sky.py:34-38,219-223).  The 624-word state lives in
// memory shared by a thread scope; the twist runs as three dependency-free sweeps.
#pragma once

#include "common.cuh"

namespace motes {

constexpr int kMtN = 624;
constexpr int kMtM = 397;

// np.random.seed(int) -> init_genrand (serial by construction; run by one thread)
MOTES_HD void mt_seed(uint32_t* mt, uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < kMtN; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
}

MOTES_HD uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

MOTES_HD uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t far) {
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// Regenerate all 624 words.  new[i] needs old[i], old[i+1] and word i+397 (old for i < 227, new after),
// so [0,227), [227,454), [454,623) are each embarrassingly parallel; word 623 needs the new word 0.
template <class Scope>
MOTES_HD void mt_twist(const Scope& sc, uint32_t* mt) {
    constexpr int d = kMtN - kMtM;   // 227
    if (sc.size() < 32) {               // canonical in-place order (host simulation / tiny scopes)
        if (sc.tid == 0) {
            for (int i = 0; i < kMtN - 1; ++i) mt[i] = mt_mix(mt[i], mt[i + 1], mt[(i + kMtM) % kMtN]);
            mt[kMtN - 1] = mt_mix(mt[kMtN - 1], mt[0], mt[kMtM - 1]);
        }
        sc.sync();
        return;
    }
    // sweep 1 and 2 read old[i+1] which the same sweep may overwrite: stage through registers
    for (int base = 0; base < 2; ++base) {
        int first = base * d, last = first + d;
        uint32_t nv[8];
        int cnt = 0;
        for (int i = first + sc.tid; i < last; i += sc.size()) {
            uint32_t far = (base == 0) ? mt[i + kMtM] : mt[i - d];
            nv[cnt++] = mt_mix(mt[i], mt[i + 1], far);
        }
        sc.sync();
        cnt = 0;
        for (int i = first + sc.tid; i < last; i += sc.size()) mt[i] = nv[cnt++];
        sc.sync();
    }
    {
        int first = 2 * d, last = kMtN - 1;
        uint32_t nv[8];
        int cnt = 0;
        for (int i = first + sc.tid; i < last; i += sc.size()) nv[cnt++] = mt_mix(mt[i], mt[i + 1], mt[i - d]);
        sc.sync();
        cnt = 0;
        for (int i = first + sc.tid; i < last; i += sc.size()) mt[i] = nv[cnt++];
        sc.sync();
    }
    if (sc.tid == 0) mt[kMtN - 1] = mt_mix(mt[kMtN - 1], mt[0], mt[kMtM - 1]);
    sc.sync();
}

// smallest 2^k - 1 >= n - 1 (numpy _bounded_uint32 mask)
MOTES_HD uint32_t rejection_mask(uint32_t n) {
    uint32_t m = n - 1;
    m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16;
    return m;
}

}  // namespace motes
