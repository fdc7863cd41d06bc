// MT19937 jump-ahead polynomials (host side, plain C++).
//
// The reference draws every bootstrap index of a frame from ONE numpy legacy MT19937 stream
// (np.random.choice, /root/reference/src/motes/sky.py:34-38), so a bit-exact replay has to consume
// a frame's stream in order.  To spread one stream over many CTAs the device code needs the
// generator state at positions k * 2^s of the stream.  With T the linear map "advance the
// 624-word window by one word" and phi its characteristic polynomial (degree 19937),
//   window(J) = T^J window(0) = g_J(T) window(0),   g_J(x) = x^J mod phi(x),
// and g_J(T) w = XOR over the set coefficients c_i of the window starting at stream word i.
// This header computes phi once (Berlekamp-Massey over GF(2) on 2 x 19937 bits of the stream) and
// the table g_{2^s}, s = 0 .. kJumpMaxLog2, by repeated squaring.  Algorithm: Haramoto, Matsumoto,
// Nishimura, Panneton, L'Ecuyer, "Efficient jump ahead for F2-linear random number generators"
// (2008), polynomial method without the sliding window.
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

namespace motes {

constexpr int kJumpDegree = 19937;
constexpr int kJumpWords64 = 313;              // 64-bit words holding bits 0 .. 19937 (phi has bit 19937)
constexpr int kJumpWords32 = 624;              // 32-bit words handed to the device: bits 0 .. 19936 (+ padding)
constexpr int kJumpMaxLog2 = 40;

namespace jump_detail {

struct HostMt {
    uint32_t mt[624];
    int pos = 624;
    explicit HostMt(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    }
    uint32_t next_raw() {                       // untempered state word, in stream order
        if (pos == 624) {
            for (int i = 0; i < 624; ++i) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            pos = 0;
        }
        return mt[pos++];
    }
};

inline int parity64(uint64_t v) { return __builtin_parityll(v); }

// dst ^= src << shift (bit shift), arrays of n 64-bit words
inline void xor_shifted(uint64_t* dst, const uint64_t* src, int n, int shift) {
    const int ws = shift >> 6, bs = shift & 63;
    if (bs == 0) {
        for (int i = n - 1; i >= ws; --i) dst[i] ^= src[i - ws];
    } else {
        for (int i = n - 1; i > ws; --i) dst[i] ^= (src[i - ws] << bs) | (src[i - ws - 1] >> (64 - bs));
        if (ws < n) dst[ws] ^= src[0] << bs;
    }
}

// Characteristic polynomial of the one-word step: bit i of the result = coefficient of x^i.
inline std::vector<uint64_t> characteristic_polynomial() {
    const int N = 2 * kJumpDegree;
    const int W = kJumpWords64 + 1;
    HostMt gen(19650218u);
    std::vector<uint8_t> s(N);
    for (int n = 0; n < N; ++n) s[n] = (uint8_t)(gen.next_raw() >> 31);
    std::vector<uint64_t> C(W, 0), B(W, 0), T(W, 0), R(W, 0);
    C[0] = 1;
    B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < N; ++n) {
        // R: bit j = s[n - j]
        for (int i = W - 1; i > 0; --i) R[i] = (R[i] << 1) | (R[i - 1] >> 63);
        R[0] = (R[0] << 1) | s[n];
        uint64_t acc = 0;
        const int used = (L >> 6) + 1;
        for (int i = 0; i < used && i < W; ++i) acc ^= C[i] & R[i];
        if (!parity64(acc)) { ++m; continue; }
        if (2 * L <= n) {
            T = C;
            xor_shifted(C.data(), B.data(), W, m);
            L = n + 1 - L;
            B = T;
            m = 1;
        } else {
            xor_shifted(C.data(), B.data(), W, m);
            ++m;
        }
    }
    // connection polynomial C (1 + c_1 x + .. + c_L x^L)  ->  phi(x) = x^L C(1/x)
    std::vector<uint64_t> phi(kJumpWords64, 0);
    if (L != kJumpDegree) return phi;           // all-zero signals failure (never happens for MT19937)
    for (int i = 0; i <= L; ++i)
        if ((C[i >> 6] >> (i & 63)) & 1) { int j = L - i; phi[j >> 6] |= 1ull << (j & 63); }
    return phi;
}

// r = a * a mod phi
inline void square_mod(const uint64_t* a, const uint64_t* phi, uint64_t* r) {
    const int W2 = 2 * kJumpWords64;
    uint64_t wide[2 * kJumpWords64 + 1];
    memset(wide, 0, sizeof(wide));
    for (int i = 0; i < kJumpWords64; ++i) {
        uint64_t v = a[i], lo = 0, hi = 0;
        for (int b = 0; b < 32; ++b) lo |= ((v >> b) & 1ull) << (2 * b);
        for (int b = 0; b < 32; ++b) hi |= ((v >> (32 + b)) & 1ull) << (2 * b);
        wide[2 * i] = lo;
        wide[2 * i + 1] = hi;
    }
    for (int bit = 2 * (kJumpDegree - 1); bit >= kJumpDegree; --bit) {
        if (!((wide[bit >> 6] >> (bit & 63)) & 1)) continue;
        const int shift = bit - kJumpDegree;
        const int ws = shift >> 6, bs = shift & 63;
        if (bs == 0) {
            for (int i = 0; i < kJumpWords64; ++i) wide[i + ws] ^= phi[i];
        } else {
            uint64_t carry = 0;
            for (int i = 0; i < kJumpWords64; ++i) {
                wide[i + ws] ^= (phi[i] << bs) | carry;
                carry = phi[i] >> (64 - bs);
            }
            wide[kJumpWords64 + ws] ^= carry;
        }
    }
    (void)W2;
    for (int i = 0; i < kJumpWords64; ++i) r[i] = wide[i];
}

}  // namespace jump_detail

// Table of x^(2^s) mod phi as 32-bit coefficient words (device layout), s = 0 .. kJumpMaxLog2.
struct JumpTable {
    std::vector<uint32_t> words;        // [(kJumpMaxLog2 + 1) * kJumpWords32]
    bool ok = false;
    const uint32_t* poly(int log2_stride) const { return words.data() + (size_t)log2_stride * kJumpWords32; }
};

inline const JumpTable& jump_table() {
    static const JumpTable table = [] {
        JumpTable t;
        std::vector<uint64_t> phi = jump_detail::characteristic_polynomial();
        bool nonzero = false;
        for (uint64_t w : phi) nonzero = nonzero || w;
        if (!nonzero) return t;
        t.words.assign((size_t)(kJumpMaxLog2 + 1) * kJumpWords32, 0u);
        std::vector<uint64_t> cur(kJumpWords64, 0), nxt(kJumpWords64, 0);
        cur[0] = 2;                              // the polynomial x
        for (int s = 0; s <= kJumpMaxLog2; ++s) {
            uint32_t* out = t.words.data() + (size_t)s * kJumpWords32;
            for (int i = 0; i < kJumpWords32; ++i) {
                uint64_t w = (i >> 1) < kJumpWords64 ? cur[i >> 1] : 0;
                out[i] = (uint32_t)((i & 1) ? (w >> 32) : w);
            }
            out[kJumpWords32 - 1] &= 0x1u;       // bits 19936 is the last coefficient (word 623 bit 0)
            if (s == kJumpMaxLog2) break;
            jump_detail::square_mod(cur.data(), phi.data(), nxt.data());
            cur.swap(nxt);
        }
        t.ok = true;
        return t;
    }();
    return table;
}

// Reference application of a jump on the host (tests): window[0..624) -> window at +2^log2_stride.
// Word 0 of the result carries only its top bit (the low 31 bits of a window's first word are not
// part of the generator state).
inline void jump_window_host(const uint32_t* poly, const uint32_t* window, uint32_t* out) {
    const int total = kJumpDegree + 624;
    std::vector<uint32_t> w(total);
    for (int i = 0; i < 624; ++i) w[i] = window[i];
    for (int n = 624; n < total; ++n) {
        uint32_t y = (w[n - 624] & 0x80000000u) | (w[n - 623] & 0x7fffffffu);
        w[n] = w[n - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (int j = 0; j < 624; ++j) out[j] = 0;
    for (int i = 0; i < kJumpDegree; ++i) {
        if (!((poly[i >> 5] >> (i & 31)) & 1u)) continue;
        for (int j = 0; j < 624; ++j) out[j] ^= w[i + j];
    }
}

}  // namespace motes
