// Segmented replay of a frame's numpy MT19937 stream for the order-0 bootstrap sky
// (sky.sky_median, /root/reference/src/motes/sky.py:17-43, called per column from
// sky.subtract_sky :257-399).  Device only.
//
// The reference consumes ONE generator stream per frame, column after column, and the number of
// words a column consumes depends on its own rejections - so a literal replay is a serial walk.
// Here the walk is reduced to its irreducible part and everything else runs on the whole GPU:
//
//   plan     per frame: how much of the stream can be needed (sum over columns of 100 * (mask + 1)
//            plus 2^20 words of slack, > 50 sigma of the rejection count);
//   jump     generator windows at stream positions k * 2^s by polynomial jump-ahead (mtjump.hpp),
//            doubling rounds: window(k) = g_{2^(s+r)}(T) window(k - 2^r);
//   stream   one CTA per (frame, segment) runs the recurrence from its window and writes the low
//            16 bits of every tempered output (all a masked-rejection draw of <= 65536 values
//            looks at) to HBM, plus a raw 624-word snapshot every 2^16 words for the state export;
//   walk     one CTA per frame scans the 16-bit stream once: for every column the position of its
//            first draw and of its last accepted draw, and the accept count of the running column
//            at every 2048-word boundary.  This is the only serial step (~2 bytes per word read);
//   column   one CTA per (frame, column): every warp takes 2048-word pieces of the column's range,
//            numbers the accepted draws with ballots (the piece's first ordinal comes from the walk),
//            writes rank slots in numpy's row-major (pixel, sample) order into shared memory and
//            reads the 100 sample medians off them; mean / std in numpy's summation order;
//   export   numpy's RandomState after the frame: the 624-word block holding the last consumed
//            word, regenerated from the nearest snapshot.
#pragma once

#include "sky.cuh"

namespace motes {

constexpr int kSnapLog2 = 16;                    // snapshot spacing (words)
constexpr uint32_t kSnapStride = 1u << kSnapLog2;
constexpr int kWalkChunk = 8192;                 // words per walk iteration (256 threads x 4 x 8)
constexpr int kSubLog2 = 11;
constexpr int kSubChunk = 1 << kSubLog2;         // granularity of the walk's running counts
constexpr uint32_t kSegSlack = 1u << 20;
constexpr int kJumpThreads = 640;
constexpr int kJumpStreamWords = 19937 + 624 + 15;
constexpr int kColThreads = 512;

// ---- plan -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
seg_plan_kernel(const uint32_t* __restrict__ rng, int ndisp, const int32_t* __restrict__ n_good,
                const int32_t* __restrict__ flag, SegLayout lay, int32_t* __restrict__ seg_need) {
    __shared__ unsigned long long part[8];
    const int f = blockIdx.x;
    unsigned long long acc = 0;
    for (int c = threadIdx.x; c < ndisp; c += blockDim.x) {
        const size_t col = (size_t)f * ndisp + c;
        const int n = n_good[col];
        if (flag[col] == kSkyModelled && n > 1) acc += 100ull * ((unsigned long long)rejection_mask((uint32_t)n) + 1ull);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long total = rng[(size_t)f * 625 + 624];
        for (int w = 0; w < 8; ++w) total += part[w];
        total += kSegSlack;
        unsigned long long need = (total + lay.stride - 1) >> lay.log2_stride;
        if (need > (unsigned long long)lay.segs) need = lay.segs;
        if (need < 1) need = 1;
        seg_need[f] = (int32_t)need;
    }
}

// Every segment of every frame (the generator is started before the sky pixels are known).
__global__ void seg_plan_all_kernel(int nframes, SegLayout lay, int32_t* __restrict__ seg_need) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < nframes) seg_need[f] = lay.segs;
}

// ---- jump -------------------------------------------------------------------------------------
// window(dst) = XOR over set coefficients c_i of the words [i, i + 624) of the stream that starts
// at window(src).  Word 0 of a jumped window carries only its top bit (all the recurrence reads).
constexpr int kJumpParts = 4;        // CTAs sharing one jump: each XORs the windows of a quarter of the coefficient words

__global__ void __launch_bounds__(kJumpThreads)
mt_jump_kernel(const uint32_t* __restrict__ rng, uint32_t* __restrict__ windows, int segs, int round_log2,
               const uint32_t* __restrict__ poly, const int32_t* __restrict__ seg_need) {
    extern __shared__ __align__(16) uint32_t jump_smem[];
    uint32_t* w = jump_smem;                          // kJumpStreamWords
    uint32_t* coef = jump_smem + kJumpStreamWords;    // 624
    const int f = blockIdx.y;
    const int k = (1 << round_log2) + blockIdx.x;
    if (k >= seg_need[f]) return;
    const int src = k - (1 << round_log2);
    const uint32_t* from = (src == 0) ? rng + (size_t)f * 625 : windows + ((size_t)f * segs + src) * 624;
    for (int i = threadIdx.x; i < 624; i += blockDim.x) { w[i] = from[i]; coef[i] = poly[i]; }
    __syncthreads();
    // this CTA's share of the coefficient words; the stream is only needed up to its last window
    const int part = blockIdx.z, per = 624 / kJumpParts;
    const int iw0 = part * per, iw1 = iw0 + per;
    const int last_word = min(19937 + 624, 32 * iw1 + 624);
    for (int base = 624; base < last_word; base += kSweep) {
        const int n = base + (int)threadIdx.x;
        if ((int)threadIdx.x < kSweep && n < last_word) w[n] = mt_mix(w[n - 624], w[n - 623], w[n - kSweep]);
        __syncthreads();
    }
    const int j = threadIdx.x;
    if (j < 624) {
        uint32_t acc = 0;
        for (int iw = iw0; iw < iw1; ++iw) {
            uint32_t cw = coef[iw];
            const uint32_t* row = w + 32 * iw + j;
            while (cw) {                               // uniform across the CTA
                const int b = __ffs(cw) - 1;
                cw &= cw - 1;
                acc ^= row[b];
            }
        }
        atomicXor(&windows[((size_t)f * segs + k) * 624 + j], acc);       // the window was zeroed before the round
    }
}

// ---- stream -----------------------------------------------------------------------------------
// Segment k of frame f: words (k L, (k+1) L] of the stream (segment 0 also the 624 state words),
// tempered and truncated to 16 bits; raw snapshots of the generator window at multiples of 2^16.
//
// The recurrence w[n] = mix(w[n-624], w[n-623]) ^ w[n-227] is run 454 words per barrier: thread
// t < 227 produces w[n0] (n0 = gen + t) and w[n0 + 227]; the far operand of the first is the
// thread's own second word of the previous round and that of the second is the word just made, so
// both stay in registers.  The ring is 18 rows of 454 words and the loop is unrolled 18 times, which
// turns every shared-memory and global address into "per-thread base + compile-time offset".
constexpr int kRowWords = 2 * kSweep;                  // 454
constexpr int kRingRows = 18;
constexpr int kRingWords = kRingRows * kRowWords;       // 8172

// ring address of stream word p (gen0 = first generated word; round j fills row (j + 2) % 18)
__device__ __forceinline__ int stream_ring_slot(uint32_t p, uint32_t gen0) {
    const int rel = (int)(p - gen0) + 2 * kRowWords;            // >= 284 for every word of the window
    const int row = rel / kRowWords, col = rel - row * kRowWords;
    return (row % kRingRows) * kRowWords + col;
}

// The recurrence and the tempering written with three-input logic ops (lop3) so that each step is one shift and
// one logic instruction: the generator kernel is integer-issue bound (ncu: 72 % of the issue slots).
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
// Pipe balance (ncu, mt_stream_kernel: ALU pipe 79 % busy, FMA pipe 12 %): shifts and logic ops all issue to the ALU
// pipe at one warp instruction per two cycles, so constant right shifts can be moved to the FMA pipe as the high half
// of a multiplication by 2^(32 - s) (MOTES_MT_HI: bit 0 = the twist's y >> 1, bit 1 = y >> 11, bit 2 = y >> 18) and
// the twist's conditional constant as a product (MOTES_MT_MAG).  Measured on B200, 128 frames of 4096 x 400 (29.7 ms as
// it was): the product 28.6 ms; IMAD.HI is slow - with the product, y >> 18 through it 29.8 ms, y >> 11 and y >> 18
// 31.2 ms, all three 32.8 ms.  So: the product, no high-half shifts.
#ifndef MOTES_MT_HI
#define MOTES_MT_HI 0
#endif
#ifndef MOTES_MT_MAG
#define MOTES_MT_MAG 1
#endif
template <int S, bool HI>
__device__ __forceinline__ uint32_t shr_const(uint32_t y) {
    return HI ? __umulhi(y, 1u << (32 - S)) : (y >> S);
}
__device__ __forceinline__ uint32_t mt_mix_fast(uint32_t a, uint32_t b, uint32_t far) {
    const uint32_t y = lop3<0xE4>(b, a, 0x7fffffffu);                 // (b & 0x7fffffff) | (a & 0x80000000)
#if MOTES_MT_MAG
    uint32_t mag;                                                      // y & 1 == b & 1; (a plain product is turned into a select)
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(mag) : "r"(b & 1u), "r"(0x9908b0dfu));
#else
    const uint32_t mag = (uint32_t)(-(int32_t)(b & 1u)) & 0x9908b0dfu;  // y & 1 == b & 1
#endif
    return lop3<0x96>(far, shr_const<1, (MOTES_MT_HI & 1) != 0>(y), mag);       // far ^ (y >> 1) ^ mag
}
__device__ __forceinline__ uint32_t mt_temper_fast(uint32_t y) {
    y ^= shr_const<11, (MOTES_MT_HI & 2) != 0>(y);
    y = lop3<0x78>(y, y << 7, 0x9d2c5680u);                            // y ^ ((y << 7) & B)
    y = lop3<0x78>(y, y << 15, 0xefc60000u);                           // y ^ ((y << 15) & C)
    return y ^ shr_const<18, (MOTES_MT_HI & 4) != 0>(y);
}

template <int U, class OutT>
__device__ __forceinline__ void stream_round(uint32_t& far, const uint32_t* pa0, const uint32_t* pa0w, const uint32_t* pa1,
                                             const uint32_t* pa1w, const uint32_t* pb, const uint32_t* pbw, uint32_t* pw,
                                             uint32_t* pww, OutT* out_t) {
    // operands of word n0: rows U / U + 1 (per thread), of word n1: row U + 1; results: row U + 2
    const uint32_t a0 = (U == kRingRows - 1) ? pa0w[U * kRowWords] : pa0[U * kRowWords];
    const uint32_t a1 = (U == kRingRows - 1) ? pa1w[U * kRowWords] : pa1[U * kRowWords];
    const uint32_t* b = (U == kRingRows - 1) ? pbw : pb;
    const uint32_t b0 = b[U * kRowWords], b1 = b[U * kRowWords + 1];
    const uint32_t v0 = mt_mix_fast(a0, a1, far);
    const uint32_t v1 = mt_mix_fast(b0, b1, v0);
    uint32_t* w = (U >= kRingRows - 2) ? pww : pw;
    w[U * kRowWords] = v0;
    w[U * kRowWords + kSweep] = v1;
    out_t[U * kRowWords] = (OutT)mt_temper_fast(v0);
    out_t[U * kRowWords + kSweep] = (OutT)mt_temper_fast(v1);
    far = v1;
}

// OutT = uint16_t: the low half of every output (order-0 bootstrap); uint32_t: whole outputs (polynomial sky).
template <class OutT>
__global__ void __launch_bounds__(256)
mt_stream_kernel(const uint32_t* __restrict__ rng, const uint32_t* __restrict__ windows, SegLayout lay,
                 const int32_t* __restrict__ seg_need, OutT* __restrict__ stream, uint32_t* __restrict__ snaps) {
    __shared__ uint32_t ring[kRingWords];
    const int f = blockIdx.y, k = blockIdx.x;
    if (k >= seg_need[f]) return;
    const uint32_t base = (uint32_t)k << lay.log2_stride;
    const uint32_t gen0 = base + 624;
    const uint32_t* from = (k == 0) ? rng + (size_t)f * 625 : windows + ((size_t)f * lay.segs + k) * 624;
    OutT* out = stream + (size_t)f * lay.stream_len;
    uint32_t* snap = snaps + (size_t)f * lay.snaps * 624;
    for (int i = threadIdx.x; i < 624; i += blockDim.x) {
        const uint32_t v = from[i];
        ring[stream_ring_slot(base + i, gen0)] = v;
        if (k == 0 || i >= 1) out[base + i] = (OutT)mt_temper(v);     // word 0 of a jumped window is not a stream word
        if (k == 0) snap[i] = v;
    }
    __syncthreads();
    const uint32_t last_out = base + lay.stride;            // last stream position this segment writes
    const uint32_t stop = last_out + 624;                   // words [gen0, stop) are needed (the tail feeds the last snapshot)
    uint32_t next_snap = ((base >> kSnapLog2) + 1) << kSnapLog2;
    const int t = threadIdx.x;
    const bool worker = t < kSweep;
    auto take_snapshot = [&](uint32_t have) {               // after a barrier; words below `have` exist
        if (next_snap <= last_out && next_snap + 624 <= have) {
            uint32_t* dst = snap + (size_t)(next_snap >> kSnapLog2) * 624;
            for (int i = threadIdx.x; i < 624; i += blockDim.x) dst[i] = ring[stream_ring_slot(next_snap + i, gen0)];
            next_snap += kSnapStride;
        }
    };
    // per-thread operand columns (see the header comment); "w" pointers serve the rounds that wrap
    const int tt = worker ? t : 0;
    const int oa0 = (tt < 170) ? tt + 284 : kRowWords + tt - 170;
    const int oa1 = (tt < 169) ? tt + 285 : kRowWords + tt - 169;
    const uint32_t* pa0 = ring + oa0;
    const uint32_t* pa1 = ring + oa1;
    const uint32_t* pa0w = ring + (oa0 + (kRingRows - 1) * kRowWords) % kRingWords - (kRingRows - 1) * kRowWords;
    const uint32_t* pa1w = ring + (oa1 + (kRingRows - 1) * kRowWords) % kRingWords - (kRingRows - 1) * kRowWords;
    const uint32_t* pb = ring + kRowWords + tt + 57;
    const uint32_t* pbw = pb - kRingWords;
    uint32_t* pw = ring + 2 * kRowWords + tt;
    uint32_t* pww = pw - kRingWords;
    uint32_t far = ring[kRowWords + kSweep + tt];            // word gen0 + t - 227
    uint32_t gen = gen0;
    OutT* out_t = out + gen0 + tt;
    // ---- full blocks of 18 rounds, every output inside the segment
    while (gen + (uint32_t)kRingWords <= last_out + 1u) {
        if (worker) {
            stream_round<0, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        }
        __syncthreads();
        if (worker) stream_round<1, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<2, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<3, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<4, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<5, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<6, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<7, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<8, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        take_snapshot(gen + 9u * kRowWords);
        if (worker) stream_round<9, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<10, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<11, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<12, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<13, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<14, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<15, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<16, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        if (worker) stream_round<17, OutT>(far, pa0, pa0w, pa1, pa1w, pb, pbw, pw, pww, out_t);
        __syncthreads();
        gen += (uint32_t)kRingWords;
        out_t += kRingWords;
        take_snapshot(gen);
    }
    // ---- tail: generic rounds (the rest of the segment and the 624 words behind it)
    // gen - gen0 is a multiple of 8172 here, so round j of the tail again fills row (j + 2) % 18
    for (int j = 0; gen < stop; ++j, gen += (uint32_t)kRowWords) {
        if (worker) {
            const uint32_t n0 = gen + (uint32_t)t, n1 = n0 + (uint32_t)kSweep;
            const uint32_t v0 = mt_mix(ring[stream_ring_slot(n0 - 624, gen0)], ring[stream_ring_slot(n0 - 623, gen0)], far);
            const uint32_t v1 = mt_mix(ring[stream_ring_slot(n1 - 624, gen0)], ring[stream_ring_slot(n1 - 623, gen0)], v0);
            ring[stream_ring_slot(n0, gen0)] = v0;
            ring[stream_ring_slot(n1, gen0)] = v1;
            if (n0 <= last_out) out[n0] = (OutT)mt_temper(v0);
            if (n1 <= last_out) out[n1] = (OutT)mt_temper(v1);
            far = v1;
        }
        __syncthreads();
        if ((j % 9) == 8 || gen + (uint32_t)kRowWords >= stop) take_snapshot(gen + (uint32_t)kRowWords);
    }
}

// ---- walk -------------------------------------------------------------------------------------
// Position bookkeeping of one frame.  The 16-bit stream is staged through a shared-memory ring by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx): one CTA has to pull ~420 MB through one
// SM, so ~200 KB are kept in flight to cover the HBM latency.  Thread t of sub-round q owns the 8
// words at chunk + q * 2048 + 8 t .. + 7, so (q, t, element) order is stream order.
constexpr int kWalkStages = 12;                                    // x 16 KB = 192 KB in flight
constexpr int kWalkChunkBytes = kWalkChunk * 2;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// accepted draws among the 8 words of d (positions pos0 .. pos0 + 7) that lie in [from, limit)
__device__ __forceinline__ int walk_count8(const uint4& d, uint32_t pos0, uint32_t from, uint32_t limit, uint32_t mask,
                                           uint32_t top) {
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t v = ((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu) & mask;
        const uint32_t p = pos0 + (uint32_t)e;
        cnt += (v <= top) && (p >= from) && (p < limit);
    }
    return cnt;
}
// the same without the position test, two words per operation: for v <= mask < 2 (top + 1) <= 0x10000 the
// 16-bit sum v + 0x8000 - (top + 1) has bit 15 set exactly when v > top and never carries
__device__ __forceinline__ int walk_count8_packed(const uint4& d, uint32_t mask2, uint32_t kk2) {
    const uint32_t t0 = (d.x & mask2) + kk2, t1 = (d.y & mask2) + kk2, t2 = (d.z & mask2) + kk2, t3 = (d.w & mask2) + kk2;
    // the flag is the top bit of the high byte of each half: gather the four high bytes of two words with one PRMT
    const uint32_t u0 = __byte_perm(t0, t1, 0x7531), u1 = __byte_perm(t2, t3, 0x7531);
    return 8 - __popc((u0 & 0x80808080u) | ((u1 >> 1) & 0x40404040u));
}

// the 8 words at pos0 .. pos0 + 7 of a piece that ends inside the stream: packed count, with the words in
// front of the column start `from` left out (only the one group that straddles `from` takes the slow path)
__device__ __forceinline__ int walk_count8_from(const uint4& d, uint32_t pos0, uint32_t from, uint32_t limit,
                                                uint32_t mask, uint32_t top, uint32_t mask2, uint32_t kk2) {
    if (pos0 >= from) return walk_count8_packed(d, mask2, kk2);
    if (pos0 + 8u <= from) return 0;
    return walk_count8(d, pos0, from, limit, mask, top);
}

constexpr int kTabRange = 64;                       // thresholds per frame the walk table covers (row = 64 x u16 = 128 B)
struct WalkTabInfo {                                // walktab.cuh
    int n_lo, range;                                // thresholds n_lo .. n_lo + range - 1
    uint32_t mask;                                  // their common rejection mask
    int ok;                                         // 1: walk_serial_kernel serves this frame; 0: sky_walk_kernel does
};

constexpr int kWalkThreads = 512;
constexpr int kWalkWarps = kWalkThreads / 32;
constexpr int kWalkPieces = 32;                            // pieces per round (one per lane of the prefix)
constexpr int kSubPerChunk = kWalkChunk / kSubChunk;
constexpr int kPieceRows = kSubChunk / 256;                // 256 words per warp-wide 16-byte load

// One CTA (16 warps) per frame.  Per column: the warps count the accepted draws of the 2048-word pieces of the
// range the column is expected to span (100 (mask + 1) words, +1 piece: > 10 sigma; up to 32 pieces a round),
// one barrier, a prefix over the pieces gives the running counts and the piece in which the column's last
// draw falls; that piece's warp pins the exact word (one packed lane-prefix of its eight row counts).
// The per-column cost is serial latency, not data (measured 1.9 us per column at nspat = 100, 2.2 us at 400):
// the next column's n / flag are fetched a column ahead, the copy engine is fed by a warp off the critical
// path, and 16 warps rather than 32 keep the replicated bookkeeping short.
__global__ void __launch_bounds__(kWalkThreads, 1)
sky_walk_kernel(const uint16_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng,
                const int32_t* __restrict__ seg_need, int ndisp, const int32_t* __restrict__ n_good,
                const int32_t* __restrict__ flag, uint32_t* __restrict__ col_start, uint32_t* __restrict__ col_end,
                uint32_t* __restrict__ sub_have, uint32_t* __restrict__ end_pos, FrameState* __restrict__ st,
                const WalkTabInfo* __restrict__ served /* optional: frames walk_serial_kernel has done */) {
    extern __shared__ __align__(128) unsigned char walk_smem[];      // kWalkStages x 16 KB
    __shared__ __align__(8) uint64_t full[kWalkStages];
    __shared__ uint32_t piece_cnt[2][kWalkPieces];
    __shared__ uint32_t cross;
    const int f = blockIdx.x;
    if (served && served[f].ok) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint16_t* s = stream + (size_t)f * lay.stream_len;
    uint32_t* subs = sub_have + (size_t)f * lay.sub_len;
    uint32_t* cs = col_start + (size_t)f * ndisp;
    uint32_t* ce = col_end + (size_t)f * ndisp;
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;     // positions [0, limit) exist
    uint32_t P = rng[(size_t)f * 625 + 624];
    if (st[f].status != 0) {
        if (tid == 0) end_pos[f] = P;
        return;
    }
    // chunk i of this walk = stream words [first + i * 8192, +8192); it lives in stage i % kWalkStages
    const uint32_t first = P & ~(uint32_t)(kWalkChunk - 1);
    if (tid == 0) {
        for (int i = 0; i < kWalkStages; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t issued = 0;                       // thread 0: chunks handed to the copy engine so far
    auto top_up = [&](uint32_t head_chunk) {   // thread 0 only; stages of chunks < head_chunk are free
        while (issued < head_chunk + (uint32_t)kWalkStages) {
            const uint32_t pos = first + issued * kWalkChunk;
            if (pos >= limit) break;
            const int stage = (int)(issued % kWalkStages);
            mbar_expect_tx(&full[stage], kWalkChunkBytes);
            tma_bulk_load(walk_smem + (size_t)stage * kWalkChunkBytes, s + pos, kWalkChunkBytes, &full[stage]);
            ++issued;
        }
    };
    // the copy engine is fed by a warp that rarely has a second piece to count
    const bool producer = tid == 32 * 4;
    if (producer) top_up(0);
    int parity = 0;
    bool failed = false;
    // per-column inputs are fetched one column ahead: their latency would otherwise sit on the serial path
    int n_next = n_good[(size_t)f * ndisp], flag_next = flag[(size_t)f * ndisp];
    for (int c = 0; c < ndisp; ++c) {
        const size_t col = (size_t)f * ndisp + c;
        const int n = n_next, col_flag = flag_next;
        if (c + 1 < ndisp) { n_next = n_good[col + 1]; flag_next = flag[col + 1]; }
        if (failed || col_flag != kSkyModelled || n <= 1) {
            if (tid == 0) { cs[c] = P; ce[c] = P; }
            continue;
        }
        const uint32_t need = (uint32_t)n * kBootstrap;
        const uint32_t mask = rejection_mask((uint32_t)n), top = (uint32_t)(n - 1);
        const bool packed_ok = top < 0x8000u;
        const uint32_t mask2 = mask * 0x00010001u, kk2 = (0x8000u - (top + 1u)) * 0x00010001u;
        if (tid == 0) cs[c] = P;
        uint32_t have = 0;
        uint32_t piece0 = P >> kSubLog2;            // first piece of the next span (absolute piece index)
        while (true) {
            if ((piece0 << kSubLog2) >= limit) { failed = true; break; }
            // span: the pieces the rest of the column is expected to need, +1, at most one per warp and
            // never more than the ring holds beyond the chunk the span starts in
            const float expect = (float)(need - have) * __fdividef((float)(mask + 1u), (float)n);     // words, a heuristic
            unsigned long long want = (unsigned long long)(expect * (1.0f / (float)kSubChunk)) + 2ull;
            const uint32_t room = (uint32_t)(kWalkStages - 1) * kSubPerChunk;
            int span = (int)(want < (unsigned long long)kWalkPieces ? want : (unsigned long long)kWalkPieces);
            if ((uint32_t)span > room) span = (int)room;
            const uint32_t head_chunk = ((piece0 << kSubLog2) - first) / kWalkChunk;
            if (producer) top_up(head_chunk);
            // this warp's pieces: piece w in the first round, piece 31 - w in the second, so that warp 0 - whose first
            // piece starts inside the previous column and takes the slower count - is the last to get a second one
            for (int round = 0; round < (kWalkPieces + kWalkWarps - 1) / kWalkWarps; ++round) {
            const int pc = (round & 1) ? (round + 1) * kWalkWarps - 1 - warp : round * kWalkWarps + warp;
            if (pc >= span) continue;
            int cnt = 0;
            const uint32_t sb = (piece0 + (uint32_t)pc) << kSubLog2;
            if (sb < limit) {
                const uint32_t chunk = (sb - first) / kWalkChunk;
                const int stage = (int)(chunk % kWalkStages);
                mbar_wait(&full[stage], (chunk / kWalkStages) & 1u);
                const uint4* src = reinterpret_cast<const uint4*>(walk_smem + (size_t)stage * kWalkChunkBytes +
                                                                  (size_t)((sb - first) % kWalkChunk) * 2);
                if (packed_ok && sb >= P && sb + kSubChunk <= limit) {
#pragma unroll
                    for (int k = 0; k < kPieceRows; ++k) cnt += walk_count8_packed(src[k * 32 + lane], mask2, kk2);
                } else if (packed_ok && sb + kSubChunk <= limit) {
#pragma unroll
                    for (int k = 0; k < kPieceRows; ++k)
                        cnt += walk_count8_from(src[k * 32 + lane], sb + (uint32_t)(k * 256 + lane * 8), P, limit, mask, top,
                                                mask2, kk2);
                } else {
#pragma unroll
                    for (int k = 0; k < kPieceRows; ++k)
                        cnt += walk_count8(src[k * 32 + lane], sb + (uint32_t)(k * 256 + lane * 8), P, limit, mask, top);
                }
                cnt = __reduce_add_sync(0xffffffffu, cnt);       // one REDUX instead of five dependent shuffles
            }
            if (lane == 0) piece_cnt[parity][pc] = (uint32_t)cnt;
            }
            __syncthreads();
            // inclusive prefix over the span's pieces, replicated in every warp
            uint32_t incl = (lane < span) ? piece_cnt[parity][lane] : 0u;
            const uint32_t own = incl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            parity ^= 1;
            const unsigned reach = __ballot_sync(0xffffffffu, lane < span && have + incl >= need);
            const uint32_t span_total = __shfl_sync(0xffffffffu, incl, 31);
            // running count of this column at the piece boundaries it covers
            if (warp == 0 && lane < span) {
                const uint32_t bpos = (piece0 + (uint32_t)lane) << kSubLog2;
                if (bpos >= P && bpos < limit) subs[piece0 + (uint32_t)lane] = have + incl - own;
            }
            if (reach) {
                const int j = __ffs(reach) - 1;
                if (warp == (((j / kWalkWarps) & 1) ? kWalkWarps - 1 - j % kWalkWarps : j % kWalkWarps)) {      // the warp that counted piece j
                    const uint32_t sb = (piece0 + (uint32_t)j) << kSubLog2;      // the piece the last draw falls in
                    const uint32_t before = have + __shfl_sync(0xffffffffu, incl - own, j);
                    uint32_t target = need - before;                    // 1-based ordinal inside this piece
                    const uint32_t chunk = (sb - first) / kWalkChunk;
                    const int stage = (int)(chunk % kWalkStages);
                    const uint4* src = reinterpret_cast<const uint4*>(walk_smem + (size_t)stage * kWalkChunkBytes +
                                                                      (size_t)((sb - first) % kWalkChunk) * 2);
                    if (packed_ok && sb >= P && sb + kSubChunk <= limit) {
                        // a whole piece of this column: reject flags of the lane's eight words per row (bit 15 - i of the
                        // low half = word 2 i, of the high half = word 2 i + 1), all rows loaded up front; per row the
                        // lanes' counts (0..8) go round as four ballots, so there is no shuffle chain
                        uint32_t rej[kPieceRows];
#pragma unroll
                        for (int k = 0; k < kPieceRows; ++k) {
                            const uint4 d = src[k * 32 + lane];
                            const uint32_t r0 = ((d.x & mask2) + kk2) & 0x80008000u, r1 = ((d.y & mask2) + kk2) & 0x80008000u;
                            const uint32_t r2 = ((d.z & mask2) + kk2) & 0x80008000u, r3 = ((d.w & mask2) + kk2) & 0x80008000u;
                            rej[k] = r0 | (r1 >> 1) | (r2 >> 2) | (r3 >> 3);
                        }
                        // lane-inclusive prefix of the eight row counts at once: 10-bit fields (a row holds <= 256 draws),
                        // three rows to a register, five shuffle steps; lane 31 then holds the row totals
                        static_assert(kPieceRows == 8, "the packed prefix below is laid out for eight rows");
                        uint32_t mine[kPieceRows];
#pragma unroll
                        for (int k = 0; k < kPieceRows; ++k) mine[k] = 8u - (uint32_t)__popc(rej[k]);
                        uint32_t pre[3] = {mine[0] | (mine[1] << 10) | (mine[2] << 20), mine[3] | (mine[4] << 10) | (mine[5] << 20),
                                           mine[6] | (mine[7] << 10)};
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const uint32_t t = __shfl_up_sync(0xffffffffu, pre[q], o);
                                if (lane >= o) pre[q] += t;
                            }
                        }
                        uint32_t tot[3];
#pragma unroll
                        for (int q = 0; q < 3; ++q) tot[q] = __shfl_sync(0xffffffffu, pre[q], 31);
                        uint32_t acc = 0, want = 0, incl = 0, own = 0, flags = 0;
                        int row = -1;
#pragma unroll
                        for (int k = 0; k < kPieceRows; ++k) {
                            const uint32_t row_total = (tot[k / 3] >> (10 * (k % 3))) & 0x3ffu;
                            if (row < 0 && target <= acc + row_total) {
                                row = k;
                                want = target - acc;
                                incl = (pre[k / 3] >> (10 * (k % 3))) & 0x3ffu;
                                own = mine[k];
                                flags = rej[k];
                            }
                            acc += row_total;
                        }
                        if (row >= 0 && incl - own < want && want <= incl) {      // this lane holds the last draw
                            uint32_t seen = incl - own;
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const uint32_t bit = 1u << ((e & 1) * 16 + 15 - (e >> 1));
                                if (!(flags & bit) && ++seen == want) cross = sb + (uint32_t)(row * 256 + lane * 8 + e) + 1u;
                            }
                        }
                    } else
                    for (int k = 0; k < kPieceRows; ++k) {
                        const uint4 d = src[k * 32 + lane];
                        const uint32_t pos0 = sb + (uint32_t)(k * 256 + lane * 8);
                        const uint32_t mine = (packed_ok && sb + kSubChunk <= limit)
                                                  ? (uint32_t)walk_count8_from(d, pos0, P, limit, mask, top, mask2, kk2)
                                                  : (uint32_t)walk_count8(d, pos0, P, limit, mask, top);
                        uint32_t li = mine;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t t = __shfl_up_sync(0xffffffffu, li, o);
                            if (lane >= o) li += t;
                        }
                        const uint32_t row_total = __shfl_sync(0xffffffffu, li, 31);
                        if (target <= row_total) {
                            if (li - mine < target && target <= li) {
                                const uint32_t w4[4] = {d.x, d.y, d.z, d.w};
                                uint32_t seen = li - mine;
                                for (int e = 0; e < 8; ++e) {
                                    const uint32_t v = ((w4[e >> 1] >> ((e & 1) * 16)) & 0xffffu) & mask;
                                    const uint32_t p = pos0 + (uint32_t)e;
                                    if ((v <= top) && (p >= P) && (p < limit)) {
                                        if (++seen == target) { cross = p + 1; break; }
                                    }
                                }
                            }
                            break;
                        }
                        target -= row_total;
                    }
                }
                __syncthreads();
                P = cross;
                if (tid == 0) ce[c] = P;
                break;
            }
            have += span_total;
            piece0 += (uint32_t)span;
        }
        if (failed) {
            if (tid == 0) { ce[c] = P; atomicMin(&st[f].status, (int)MOTES_E_INTERNAL); }
        }
    }
    if (tid == 0) end_pos[f] = P;
}

// ---- column -----------------------------------------------------------------------------------
struct ColShared {
    uint32_t* slots;     // 100 rows x row_words words (uint16 rank per drawn pixel)
    uint16_t* ranks;     // nspat
    double* meds;        // 100
    double* sq;          // 100
};

// median of every bootstrap sample from its n rank slots: bitwise binary search, one warp per sample.
// Packed path (n <= 64 NW): each lane keeps NW words = 2 NW ranks in registers; "rank >= cand" for
// both halves of a word is bit 15 of (half + 0x8000 - cand).
template <int NW>
__device__ __forceinline__ void column_medians_packed(const ColShared& sh, int n, int rw, const double* __restrict__ srt,
                                                      int lane, int warp, int nwarps) {
    const int k1 = (n - 1) / 2, k2 = n / 2;
    int nbits = 1;
    while ((1 << nbits) < n) ++nbits;
    const int nwords = (n + 1) >> 1;
    constexpr int slots_total = 64 * NW;
    for (int smp = warp; smp < kBootstrap; smp += nwarps) {
        const uint32_t* row = sh.slots + smp * rw;
        uint32_t reg[NW];
#pragma unroll
        for (int t = 0; t < NW; ++t) {
            const int wd = lane + 32 * t;
            uint32_t wv = 0x7fff7fffu;                      // padding: larger than any rank
            if (wd < nwords) {
                wv = row[wd];
                if (2 * wd + 1 >= n) wv = (wv & 0xffffu) | 0x7fff0000u;
            }
            reg[t] = wv;
        }
        int x1 = 0;
        for (int bit = nbits - 1; bit >= 0; --bit) {
            const int cand = x1 | (1 << bit);
            const uint32_t kk = (uint32_t)(0x8000 - cand) * 0x00010001u;
            uint32_t ge = 0;
#pragma unroll
            for (int t = 0; t < NW; ++t) ge += ((reg[t] + kk) >> 15) & 0x00010001u;
            const int ge_all = __reduce_add_sync(0xffffffffu, (int)((ge & 0xffffu) + (ge >> 16)));
            if (slots_total - ge_all <= k1) x1 = cand;
        }
        int x2 = x1;
        if (k2 != k1) {
            const uint32_t kk = (uint32_t)(0x8000 - (x1 + 1)) * 0x00010001u;
            uint32_t ge = 0;
            int nxt = 0x7fff;
#pragma unroll
            for (int t = 0; t < NW; ++t) {
                ge += ((reg[t] + kk) >> 15) & 0x00010001u;
                const int h0 = (int)(reg[t] & 0xffffu), h1 = (int)(reg[t] >> 16);
                if (h0 > x1 && h0 < nxt) nxt = h0;
                if (h1 > x1 && h1 < nxt) nxt = h1;
            }
            const int le = slots_total - __reduce_add_sync(0xffffffffu, (int)((ge & 0xffffu) + (ge >> 16)));
            nxt = __reduce_min_sync(0xffffffffu, nxt);
            if (le <= k2) x2 = nxt;
        }
        if (lane == 0) sh.meds[smp] = (k1 == k2) ? srt[x1] : (srt[x1] + srt[x2]) * 0.5;
    }
}

__device__ __forceinline__ void column_medians(const ColShared& sh, int n, int rw, const double* __restrict__ srt,
                                               int lane, int warp, int nwarps) {
    switch ((n + 63) >> 6) {
        case 1: column_medians_packed<1>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 2: column_medians_packed<2>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 3: column_medians_packed<3>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 4: column_medians_packed<4>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 5: column_medians_packed<5>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 6: column_medians_packed<6>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 7: column_medians_packed<7>(sh, n, rw, srt, lane, warp, nwarps); return;
        case 8: column_medians_packed<8>(sh, n, rw, srt, lane, warp, nwarps); return;
        default: break;
    }
    const int k1 = (n - 1) / 2, k2 = n / 2;
    int nbits = 1;
    while ((1 << nbits) < n) ++nbits;
    const uint16_t* slots16 = reinterpret_cast<const uint16_t*>(sh.slots);
    for (int smp = warp; smp < kBootstrap; smp += nwarps) {
        const uint16_t* row = slots16 + smp * (2 * rw);
        int x1 = 0;
        for (int bit = nbits - 1; bit >= 0; --bit) {
            const int cand = x1 | (1 << bit);
            int cnt = 0;
            for (int j = lane; j < n; j += 32) cnt += ((int)row[j] < cand);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt <= k1) x1 = cand;
        }
        int x2 = x1;
        if (k2 != k1) {
            int le = 0, nxt = 0x7fffffff;
            for (int j = lane; j < n; j += 32) {
                const int r = (int)row[j];
                le += (r <= x1);
                if (r > x1 && r < nxt) nxt = r;
            }
            le = __reduce_add_sync(0xffffffffu, le);
            nxt = __reduce_min_sync(0xffffffffu, nxt);
            if (le <= k2) x2 = nxt;
        }
        if (lane == 0) sh.meds[smp] = (k1 == k2) ? srt[x1] : (srt[x1] + srt[x2]) * 0.5;
    }
}

// np.add.reduce over 100 doubles (pairwise_sum_DOUBLE: 8 accumulators, then the tail), by one warp
__device__ __forceinline__ double warp_numpy_sum100(const double* a, int lane) {
    double v = 0.0;
    if (lane < 8) {
        v = a[lane];
        for (int i = 1; i < 12; ++i) v += a[8 * i + lane];
    }
    const double t = v + __shfl_down_sync(0xffffffffu, v, 1);
    const double u = t + __shfl_down_sync(0xffffffffu, t, 2);
    const double w = u + __shfl_down_sync(0xffffffffu, u, 4);
    double res = ((w + a[96]) + a[97]) + a[98];
    res = res + a[99];
    return __shfl_sync(0xffffffffu, res, 0);
}

// The accepted draws of one column, piece by piece: every warp takes 2048-word pieces of [P, Pend); the
// ordinal of a piece's first accept comes from the walk.  sink(ord, m4, v) receives, per lane and row of
// 128 words, the ordinal of the lane's first accept, the 4-bit accept mask of its four words (stream order)
// and the four masked words.
template <class Sink>
__device__ __forceinline__ void column_draws(const uint16_t* __restrict__ s, const uint32_t* __restrict__ subs, uint32_t P,
                                             uint32_t Pend, int n, int lane, int warp, int nwarps, Sink sink) {
    if (Pend <= P) return;
    const int need = n * kBootstrap;
    const uint32_t mask = rejection_mask((uint32_t)n), top = (uint32_t)(n - 1);
    const unsigned lanes_below = (1u << lane) - 1u;
    const uint32_t first = P >> kSubLog2, last = (Pend - 1) >> kSubLog2;
    for (uint32_t sub = first + warp; sub <= last; sub += nwarps) {
        const uint32_t sb = sub << kSubLog2;
        const uint32_t start = max(P, sb), end = min(Pend, sb + (uint32_t)kSubChunk);
        int ord0 = (sb > P) ? (int)subs[sub] : 0;
        // four rows (512 words) of loads in flight per warp: the stream comes from HBM / L2
        for (uint32_t row0 = start & ~127u; row0 < end; row0 += 512) {
            uint2 raw4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) raw4[r] = __ldg(reinterpret_cast<const uint2*>(s + row0 + 128u * r + 4u * lane));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t row = row0 + 128u * r;
                if (row >= end) break;
                const uint32_t pos = row + 4u * lane;
                const uint2 raw = raw4[r];
                const uint32_t v[4] = {raw.x & mask, (raw.x >> 16) & mask, raw.y & mask, (raw.y >> 16) & mask};
                unsigned m4 = (unsigned)(v[0] <= top) | ((unsigned)(v[1] <= top) << 1) | ((unsigned)(v[2] <= top) << 2) |
                              ((unsigned)(v[3] <= top) << 3);
                if (row < start || row + 128u > end) {           // first / last row of the piece
                    unsigned inside = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) inside |= (unsigned)(pos + e >= start && pos + e < end) << e;
                    m4 &= inside;
                }
                const int cnt = __popc(m4);
                // ordinal of the lane's first accept: accepts of the lower lanes (count <= 4: three ballots)
                const unsigned c0 = __ballot_sync(0xffffffffu, cnt & 1), c1 = __ballot_sync(0xffffffffu, cnt & 2),
                               c2 = __ballot_sync(0xffffffffu, cnt & 4);
                const int before = __popc(c0 & lanes_below) + 2 * __popc(c1 & lanes_below) + 4 * __popc(c2 & lanes_below);
                const int total = __popc(c0) + 2 * __popc(c1) + 4 * __popc(c2);
                if (ord0 + total <= need) sink(ord0 + before, m4, v);      // always, unless walk and column disagree
                ord0 += total;
            }
        }
    }
}

// np.nanmean / np.std of the 100 medians in numpy's summation order (sky.py:40-41); warp 0 of the CTA
__device__ __forceinline__ void column_mean_std(const double* meds, double* sq, int lane, double* sky, double* sky_err) {
    const double mean = warp_numpy_sum100(meds, lane) / (double)kBootstrap;
    for (int smp = lane; smp < kBootstrap; smp += 32) { const double d = meds[smp] - mean; sq[smp] = d * d; }
    __syncwarp();
    const double var = warp_numpy_sum100(sq, lane) / (double)kBootstrap;
    if (lane == 0) {
        *sky = mean;
        *sky_err = sqrt(var) / kSqrt99;
    }
}

// Exact for any n: every draw's rank is stored in numpy's (pixel, sample) slot, medians by bitwise search.
// With `list` the CTAs walk a list of columns (the ones the histogram kernel below could not settle).
__global__ void __launch_bounds__(kColThreads, 2)
sky_column_kernel(const uint16_t* __restrict__ stream, SegLayout lay, int nspat, int ndisp,
                  const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                  const double* __restrict__ sorted, const uint16_t* __restrict__ rank_of,
                  const uint32_t* __restrict__ col_start, const uint32_t* __restrict__ col_end,
                  const uint32_t* __restrict__ sub_have, double* __restrict__ sky, double* __restrict__ sky_err,
                  const FrameState* __restrict__ st, const uint32_t* __restrict__ list,
                  const uint32_t* __restrict__ list_count) {
    extern __shared__ __align__(16) unsigned char col_smem[];
    ColShared sh;
    sh.meds = reinterpret_cast<double*>(col_smem);
    sh.sq = sh.meds + kBootstrap;
    sh.slots = reinterpret_cast<uint32_t*>(sh.sq + kBootstrap);
    sh.ranks = reinterpret_cast<uint16_t*>(sh.slots + (size_t)kBootstrap * boot_row_words(nspat));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kColThreads / 32;
    const uint32_t items = list ? *list_count : 1u;
    for (uint32_t item = list ? blockIdx.x : 0u; item < items; item += gridDim.x) {
        const int c = list ? (int)(list[item] % (uint32_t)ndisp) : (int)blockIdx.x;
        const int f = list ? (int)(list[item] / (uint32_t)ndisp) : (int)blockIdx.y;
        if (st[f].status != 0) continue;
        const size_t col = (size_t)f * ndisp + c;
        if (flag[col] != kSkyModelled) {
            if (tid == 0) { sky[col] = 0.0; sky_err[col] = 0.0; }
            continue;
        }
        const int n = n_good[col];
        if (n == 0) {
            if (tid == 0) { sky[col] = NAN; sky_err[col] = NAN; }
            continue;
        }
        const double* srt = sorted + col * nspat;
        __syncthreads();                 // list mode: the previous column's medians have been consumed
        if (n > 1) {
            const int rw = boot_row_words(n);
            uint16_t* slots16 = reinterpret_cast<uint16_t*>(sh.slots);
            for (int i = tid; i < n; i += kColThreads) sh.ranks[i] = rank_of[col * nspat + i];
            __syncthreads();
            const uint16_t* ranks = sh.ranks;
            column_draws(stream + (size_t)f * lay.stream_len, sub_have + (size_t)f * lay.sub_len, col_start[col], col_end[col],
                         n, lane, warp, nwarps, [&](int ord, unsigned m4, const uint32_t (&v)[4]) {
                             const int pix = ord / kBootstrap;
                             int smp = ord - pix * kBootstrap;
                             int at = smp * (2 * rw) + pix;          // slot of ordinal ord; the next one is a row down
#pragma unroll
                             for (int e = 0; e < 4; ++e) {
                                 if (m4 & (1u << e)) {
                                     slots16[at] = ranks[v[e]];
                                     at += 2 * rw;
                                     if (++smp == kBootstrap) { smp = 0; at -= kBootstrap * 2 * rw - 1; }
                                 }
                             }
                         });
            __syncthreads();
            column_medians(sh, n, rw, srt, lane, warp, nwarps);
        } else {
            for (int smp = tid; smp < kBootstrap; smp += kColThreads) sh.meds[smp] = srt[0];
        }
        __syncthreads();
        if (warp == 0) column_mean_std(sh.meds, sh.sq, lane, sky + col, sky_err + col);
    }
}

// The same result from counts instead of slots: the median rank of a bootstrap sample of n ranks sits
// within a few sqrt(n) / 2 of n / 2, so each sample only keeps an 8-bit histogram of the ranks inside a
// window around n / 2 plus the number of draws below and above it; its median is read off a prefix sum.
// A sample whose median falls outside the window (or a count that does not add up to n: a histogram byte
// overflowed) sends the column to the exact kernel above through `list`.
//
// Two layouts of a sample's row: PACKED - 64 words of four 8-bit bins (a window of 256 ranks, for n up to ~1500);
// plain - 128 words of one 32-bit bin each (a window of 128 ranks, enough for n <= 400: 6.4 sigma), where a draw
// is an atomic add of the constant 1 at a table-supplied word and costs a third fewer instructions.
template <bool PACKED>
struct HistLayout {
    static constexpr int kWords = PACKED ? 64 : 128;
    static constexpr int kBins = PACKED ? 256 : 128;
    static constexpr int kStride = kWords + 3;      // + draws below / above the window, + 1 word of bank skew
};

template <bool PACKED>
__global__ void __launch_bounds__(kColThreads, 3)
sky_column_hist_kernel(const uint16_t* __restrict__ stream, SegLayout lay, int nspat, int ndisp, int window,
                       const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                       const double* __restrict__ sorted, const uint16_t* __restrict__ rank_of,
                       const uint32_t* __restrict__ col_start, const uint32_t* __restrict__ col_end,
                       const uint32_t* __restrict__ sub_have, double* __restrict__ sky, double* __restrict__ sky_err,
                       const FrameState* __restrict__ st, uint32_t frame0, uint32_t* __restrict__ list,
                       uint32_t* __restrict__ list_count) {
    constexpr int kHistWords = HistLayout<PACKED>::kWords, kHistStride = HistLayout<PACKED>::kStride;
    extern __shared__ __align__(16) unsigned char colh_smem[];
    double* meds = reinterpret_cast<double*>(colh_smem);
    double* sq = meds + kBootstrap;
    // per sample: 64 words of 8-bit bins, then the counts of draws below and above the window
    uint32_t* hist = reinterpret_cast<uint32_t*>(sq + kBootstrap);                 // 100 x kHistStride
    // per good pixel (in the order the draws index them): which counter a draw of it bumps, and by how much:
    // low byte = word inside the sample's row, high byte = shift of the increment
    uint16_t* entry = reinterpret_cast<uint16_t*>(hist + kBootstrap * kHistStride);   // nspat
    __shared__ int failed;
    const int c = blockIdx.x, f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kColThreads / 32;
    if (st[f].status != 0) return;
    const size_t col = (size_t)f * ndisp + c;
    if (flag[col] != kSkyModelled) {
        if (tid == 0) { sky[col] = 0.0; sky_err[col] = 0.0; }
        return;
    }
    const int n = n_good[col];
    if (n == 0) {
        if (tid == 0) { sky[col] = NAN; sky_err[col] = NAN; }
        return;
    }
    const double* srt = sorted + col * nspat;
    const int lo = max(0, n / 2 - window / 2);     // ranks [lo, lo + window) are histogrammed
    if (n > 1) {
        for (int i = tid; i < kBootstrap * kHistStride; i += kColThreads) hist[i] = 0u;
        for (int i = tid; i < n; i += kColThreads) {
            const int r = (int)rank_of[col * nspat + i] - lo;
            const bool in = (unsigned)r < (unsigned)window;
            if (PACKED) entry[i] = in ? (uint16_t)((r >> 2) | ((8 * (r & 3)) << 8)) : (uint16_t)(r < 0 ? kHistWords : kHistWords + 1);
            else entry[i] = in ? (uint16_t)r : (uint16_t)(r < 0 ? kHistWords : kHistWords + 1);
        }
        if (tid == 0) failed = 0;
        __syncthreads();
        column_draws(stream + (size_t)f * lay.stream_len, sub_have + (size_t)f * lay.sub_len, col_start[col], col_end[col], n,
                     lane, warp, nwarps, [&](int ord, unsigned m4, const uint32_t (&v)[4]) {
                         int smp = ord % kBootstrap;
                         uint32_t* row = hist + smp * kHistStride;
#pragma unroll
                         for (int e = 0; e < 4; ++e) {
                             if (m4 & (1u << e)) {
                                 const uint32_t en = entry[v[e]];
                                 if (PACKED) atomicAdd(row + (en & 0xffu), 1u << (en >> 8));      // one shared-memory atomic per draw
                                 else atomicAdd(row + en, 1u);
                                 row += kHistStride;
                                 if (++smp == kBootstrap) { smp = 0; row = hist; }
                             }
                         }
                     });
        __syncthreads();
        const int k1 = (n - 1) / 2, k2 = n / 2;
        for (int smp = warp; smp < kBootstrap; smp += nwarps) {
            const uint32_t* row = hist + smp * kHistStride;
            // this lane's share of the window: 8 bins (two packed words) or 4 bins (four plain words)
            constexpr int kLaneBins = HistLayout<PACKED>::kBins / 32;
            uint32_t cntb[kLaneBins];
            if (PACKED) {
                const uint32_t w0 = row[2 * lane], w1 = row[2 * lane + 1];
#pragma unroll
                for (int bin = 0; bin < kLaneBins; ++bin) cntb[bin] = ((bin < 4 ? w0 : w1) >> (8 * (bin & 3))) & 0xffu;
            } else {
#pragma unroll
                for (int bin = 0; bin < kLaneBins; ++bin) cntb[bin] = row[kLaneBins * lane + bin];
            }
            int mine = 0;
#pragma unroll
            for (int bin = 0; bin < kLaneBins; ++bin) mine += (int)cntb[bin];
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int inside = __shfl_sync(0xffffffffu, incl, 31);
            const int nb = (int)row[kHistWords], na = (int)row[kHistWords + 1];
            const bool ok = (nb + inside + na == n) && (k1 >= nb) && (k2 < nb + inside);
            if (!ok) {
                if (lane == 0) failed = 1;
                continue;
            }
            // order statistics k1 and k2 of the sample: bin of the (k - nb)-th draw inside the window
            const int t1 = k1 - nb, t2 = k2 - nb, excl = incl - mine;
            int x1 = -1, x2 = -1;
            if ((t1 >= excl && t1 < incl) || (t2 >= excl && t2 < incl)) {
                int seen = excl;
#pragma unroll
                for (int bin = 0; bin < kLaneBins; ++bin) {
                    const int cnt = (int)cntb[bin];
                    if (t1 >= seen && t1 < seen + cnt) x1 = lo + kLaneBins * lane + bin;
                    if (t2 >= seen && t2 < seen + cnt) x2 = lo + kLaneBins * lane + bin;
                    seen += cnt;
                }
            }
            x1 = __reduce_max_sync(0xffffffffu, x1);
            x2 = __reduce_max_sync(0xffffffffu, x2);
            if (lane == 0) meds[smp] = (k1 == k2) ? srt[x1] : (srt[x1] + srt[x2]) * 0.5;
        }
    } else {
        if (tid == 0) failed = 0;
        for (int smp = tid; smp < kBootstrap; smp += kColThreads) meds[smp] = srt[0];
    }
    __syncthreads();
    if (failed) {
        if (tid == 0) list[atomicAdd(list_count, 1u)] = (frame0 + (uint32_t)f) * (uint32_t)ndisp + (uint32_t)c;
        return;
    }
    if (warp == 0) column_mean_std(meds, sq, lane, sky + col, sky_err + col);
}

// ---- plain 32-bit histogram, second form (n <= 512) ------------------------------------------------------
// The same counts as the plain layout of sky_column_hist_kernel above (which now only runs its packed layout, for
// n > 512), with the row loop rebuilt around the pipe that limited it (ncu: ALU pipe 67 %, 122 warp instructions per
// 128 words, a fifth of the kernel in the 100 warp-wide median scans):
//   * 8 words per lane and row (one 16-byte load); acceptance of two draws per operation - bit 15 of
//     (half + 0x8000 - n) - gathered into one flag word per lane by two PRMT (the walk's packed form);
//   * four ballots number the accepts of 256 words; the row total comes from one REDUX;
//   * a lane's accepted draws go to consecutive samples: the row pointer just advances, and a run that passes
//     sample 99 lands in rows 100 .. 106, which are added to rows 0 .. 6 before the medians are read - no wrap test
//     per draw: table look-up, address, increment, advance (behind a branch: ptxas does not predicate shared atomics);
//   * medians: one THREAD per sample walks its 128 bins (rows are 16-byte aligned, stride 132 words: the four
//     quarter-warp phases of a 16-byte shared load hit distinct banks) instead of one warp per sample.
constexpr int kH32Bins = 128;
#ifndef MOTES_H32_STRIDE
#define MOTES_H32_STRIDE 132
#endif
// + draws below / above the window.  132: 16-byte aligned rows (the medians read them with 16-byte loads); 133: an odd
// stride - the below / above counters of consecutive samples, 60 % of all draws, then lie in distinct banks
constexpr int kH32Stride = MOTES_H32_STRIDE;
constexpr int kH32HistWords = (kH32Stride * (kBootstrap + 7) + 3) & ~3;
constexpr int kH32Rows = kBootstrap + 7;                 // a lane's run of 8 draws may pass sample 99
constexpr int kH32RowBytes = kH32Stride * 4;
#ifndef MOTES_H32_THREADS
#define MOTES_H32_THREADS 416
#endif
#ifndef MOTES_H32_FLIGHT
#define MOTES_H32_FLIGHT 4       // B200, 128 frames of 4096 x 400: 1 row in flight 29.3 ms, 2: 26.0, 3: 25.3, 4: 24.8 (48 registers)
#endif
// 13 warps: a column of 257 .. 512 good sky pixels spans 100 x 512 words = 25 pieces of 2048, i.e. 26 with the two
// partial ones at its ends - two pieces per warp; three CTAs per SM leave 48 registers per thread.  Narrower slits
// (shorter columns) are launched with one warp per expected piece (h32_threads)
constexpr int kH32Threads = MOTES_H32_THREADS;
constexpr int kH32Flight = MOTES_H32_FLIGHT;             // 256-word rows a warp has in flight
constexpr int kH32Entries = 512;                         // table entries (mask + 1 for n <= 512: any masked word stays inside)

inline int h32_threads(int nspat) {
    unsigned p2 = 1;
    while (p2 < (unsigned)nspat) p2 <<= 1;
    int warps = (int)((100u * p2 + kSubChunk - 1) / kSubChunk) + 1;          // pieces a column of nspat draws can touch
    if (warps > kH32Threads / 32) warps = kH32Threads / 32;
    if (warps < 4) warps = 4;                                                 // (the medians: 100 threads)
    return warps * 32;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t at) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(at) : "memory");
    return v;
}
__device__ __forceinline__ void reds_inc(uint32_t at) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(at) : "memory");
}
// one draw: bump the counter its pixel's table entry names in the sample row at `rowaddr`, move on to the next sample
__device__ __forceinline__ void h32_draw(uint32_t& rowaddr, uint32_t flag, uint32_t entry_at) {
    if (flag) {
        reds_inc(rowaddr + 4u * lds_u16(entry_at));
        rowaddr += (uint32_t)kH32RowBytes;
    }
}
// lanes whose x has bit BIT set
template <uint32_t BIT>
__device__ __forceinline__ unsigned ballot_bit(uint32_t x) {
    unsigned r;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b32 t;\n"
        "and.b32 t, %1, %2;\n"
        "setp.ne.b32 p, t, 0;\n"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n"
        "}\n"
        : "=r"(r)
        : "r"(x), "n"(BIT));
    return r;
}

__global__ void __launch_bounds__(kH32Threads, 3)
sky_column_hist32_kernel(const uint16_t* __restrict__ stream, SegLayout lay, int nspat, int ndisp, int window,
                         const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                         const double* __restrict__ sorted, const uint16_t* __restrict__ rank_of,
                         const uint32_t* __restrict__ col_start, const uint32_t* __restrict__ col_end,
                         const uint32_t* __restrict__ sub_have, double* __restrict__ sky, double* __restrict__ sky_err,
                         const FrameState* __restrict__ st, uint32_t frame0, uint32_t* __restrict__ list,
                         uint32_t* __restrict__ list_count) {
    extern __shared__ __align__(16) unsigned char colh_smem[];
    double* meds = reinterpret_cast<double*>(colh_smem);
    double* sq = meds + kBootstrap;
    uint32_t* hist = reinterpret_cast<uint32_t*>(sq + kBootstrap);                   // kH32Rows x kH32Stride
    uint16_t* entry = reinterpret_cast<uint16_t*>(hist + kH32HistWords);              // kH32Entries: word a draw of this pixel bumps
    __shared__ int failed;
    const int c = blockIdx.x, f = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = (int)blockDim.x, nwarps = nthreads >> 5;     // <= kH32Threads: fewer warps for short columns
    if (st[f].status != 0) return;
    const size_t col = (size_t)f * ndisp + c;
    if (flag[col] != kSkyModelled) {
        if (tid == 0) { sky[col] = 0.0; sky_err[col] = 0.0; }
        return;
    }
    const int n = n_good[col];
    if (n == 0) {
        if (tid == 0) { sky[col] = NAN; sky_err[col] = NAN; }
        return;
    }
    const double* srt = sorted + col * nspat;
    const int lo = max(0, n / 2 - window / 2);     // ranks [lo, lo + window) are histogrammed
    if (n > 1) {
        {
            uint4* h4 = reinterpret_cast<uint4*>(hist);
            for (int i = tid; i < kH32HistWords / 4; i += nthreads) h4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int i = tid; i < kH32Entries; i += nthreads) {
            uint16_t en = (uint16_t)(kH32Bins + 2);                                   // (never a draw: a spare word of the row)
            if (i < n) {
                const int r = (int)rank_of[col * nspat + i] - lo;
                en = ((unsigned)r < (unsigned)window) ? (uint16_t)r : (uint16_t)(r < 0 ? kH32Bins : kH32Bins + 1);
            }
            entry[i] = en;
        }
        if (tid == 0) failed = 0;
        __syncthreads();
        const uint32_t P = col_start[col], Pend = col_end[col];
        if (Pend > P) {
            const uint16_t* s = stream + (size_t)f * lay.stream_len;
            const uint32_t* subs = sub_have + (size_t)f * lay.sub_len;
            // (volatile: otherwise the address conversion is rematerialised - four instructions - in every row)
            uint32_t hist_at, entry_at;
            asm volatile("{\n.reg .u64 t;\ncvta.to.shared.u64 t, %2;\ncvt.u32.u64 %0, t;\nadd.u32 %1, %0, %3;\n}\n"
                         : "=r"(hist_at), "=r"(entry_at)
                         : "l"(hist), "n"(kH32HistWords * 4));
            const int need = n * kBootstrap;
            const uint32_t mask = rejection_mask((uint32_t)n);
            const uint32_t mask2 = mask * 0x00010001u, kk2 = (uint32_t)(0x8000 - n) * 0x00010001u, mask_hi = mask << 1;
            const unsigned lanes_below = (1u << lane) - 1u;
            const uint32_t first = P >> kSubLog2, last = (Pend - 1) >> kSubLog2;
            for (uint32_t sub = first + warp; sub <= last; sub += nwarps) {
                const uint32_t sb = sub << kSubLog2;
                const uint32_t start = max(P, sb), end = min(Pend, sb + (uint32_t)kSubChunk);
                int ord0 = (sb > P) ? (int)subs[sub] : 0;
                for (uint32_t row0 = start & ~255u; row0 < end; row0 += 256u * kH32Flight) {
                    uint4 raw4[kH32Flight];
#pragma unroll
                    for (int r = 0; r < kH32Flight; ++r) raw4[r] = __ldg(reinterpret_cast<const uint4*>(s + row0 + 256u * r + 8u * lane));
#pragma unroll
                    for (int r = 0; r < kH32Flight; ++r) {
                        const uint32_t row = row0 + 256u * r;
                        if (row >= end) break;
                        const uint4 d = raw4[r];
                        const uint32_t t0 = (d.x & mask2) + kk2, t1 = (d.y & mask2) + kk2, t2 = (d.z & mask2) + kk2,
                                       t3 = (d.w & mask2) + kk2;
                        // accept flags: word e < 4 at bit 8 e + 7, word e >= 4 at bit 8 (e - 4) + 6
                        const uint32_t u0 = __byte_perm(t0, t1, 0x7531), u1 = __byte_perm(t2, t3, 0x7531);
                        uint32_t acc = (~u0 & 0x80808080u) | ((~u1 >> 1) & 0x40404040u);
                        if (row < start || row + 256u > end) {           // first / last row of the column
                            const uint32_t pos = row + 8u * lane;
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (pos + e < start || pos + e >= end) acc &= ~(e < 4 ? 0x80u << (8 * e) : 0x40u << (8 * (e - 4)));
                        }
                        const uint32_t cnt = (uint32_t)__popc(acc);
                        const unsigned c0 = ballot_bit<1>(cnt), c1 = ballot_bit<2>(cnt), c2 = ballot_bit<4>(cnt), c3 = ballot_bit<8>(cnt);
                        const int total = __reduce_add_sync(0xffffffffu, (int)cnt);
                        if (ord0 + total <= need) {                      // always, unless walk and column disagree
                            const uint32_t ord = (uint32_t)ord0 + __popc(c0 & lanes_below) + 2 * __popc(c1 & lanes_below) +
                                                 4 * __popc(c2 & lanes_below) + 8 * __popc(c3 & lanes_below);
                            const uint32_t smp = ord - __umulhi(ord, 42949673u) * (uint32_t)kBootstrap;       // ord % 100 (ord < 2^22)
                            uint32_t rowaddr = hist_at + smp * (uint32_t)kH32RowBytes;
                            const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const uint32_t bit = e < 4 ? 0x80u << (8 * e) : 0x40u << (8 * (e - 4));
                                // byte offset of the draw's table entry: 2 v
                                const uint32_t v2 = (e & 1) ? ((w[e >> 1] >> 15) & mask_hi) : ((w[e >> 1] & mask) << 1);
                                h32_draw(rowaddr, acc & bit, entry_at + v2);
                            }
                        }
                        ord0 += total;
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < 7 * kH32Stride; i += nthreads) hist[i] += hist[kBootstrap * kH32Stride + i];
        __syncthreads();
        if (tid < kBootstrap) {
            const int k1 = (n - 1) / 2, k2 = n / 2;
            // bins whose running count stays <= t: the bin holding the sample's t-th draw inside the window
            int cum = 0, x1 = 0, x2 = 0;
#if (MOTES_H32_STRIDE % 4) == 0
            const uint4* row4 = reinterpret_cast<const uint4*>(hist + tid * kH32Stride);
            const uint4 tail = row4[kH32Bins / 4];
            const int nb = (int)tail.x, na = (int)tail.y;
            const int t1 = k1 - nb, t2 = k2 - nb;
#pragma unroll 4
            for (int q = 0; q < kH32Bins / 4; ++q) {
                const uint4 b = row4[q];
                cum += (int)b.x; x1 += (cum <= t1); x2 += (cum <= t2);
                cum += (int)b.y; x1 += (cum <= t1); x2 += (cum <= t2);
                cum += (int)b.z; x1 += (cum <= t1); x2 += (cum <= t2);
                cum += (int)b.w; x1 += (cum <= t1); x2 += (cum <= t2);
            }
#else
            const uint32_t* row = hist + tid * kH32Stride;        // (odd stride: consecutive samples, distinct banks)
            const int nb = (int)row[kH32Bins], na = (int)row[kH32Bins + 1];
            const int t1 = k1 - nb, t2 = k2 - nb;
#pragma unroll 8
            for (int q = 0; q < kH32Bins; ++q) { cum += (int)row[q]; x1 += (cum <= t1); x2 += (cum <= t2); }
#endif
            const bool ok = (nb + cum + na == n) && (k1 >= nb) && (k2 < nb + cum);
            if (!ok) failed = 1;
            else meds[tid] = (k1 == k2) ? srt[lo + x1] : (srt[lo + x1] + srt[lo + x2]) * 0.5;
        }
    } else {
        if (tid == 0) failed = 0;
        for (int smp = tid; smp < kBootstrap; smp += nthreads) meds[smp] = srt[0];
    }
    __syncthreads();
    if (failed) {
        if (tid == 0) list[atomicAdd(list_count, 1u)] = (frame0 + (uint32_t)f) * (uint32_t)ndisp + (uint32_t)c;
        return;
    }
    if (warp == 0) column_mean_std(meds, sq, lane, sky + col, sky_err + col);
}

// ---- polynomial sky (sky.sky_polynomial, sky.py:177-254) over the segmented generator ----------------------
// np.random.normal (legacy polar method) consumes the stream four words per attempt and accepts an attempt
// when 0 < r2 < 1 - a property of the words alone, not of the column that draws.  So, with the whole 32-bit
// stream in HBM (mt_stream_kernel<uint32_t>): the acceptance bit of every attempt is computed in parallel
// (poly_attempts_kernel), a prefix sum over them tells where the accept with a given ordinal sits, and a column
// that needs 50 n accepted attempts after the 50 (n_0 + .. + n_{c-1}) of its predecessors gets its attempt range
// from two look-ups (poly_bounds_kernel).  The columns then run independently (sky_poly_column_kernel).
constexpr int kAttemptBlock = 1024;              // attempts per prefix-sum block (32 bitmap words)

__global__ void __launch_bounds__(256)
seg_plan_poly_kernel(const uint32_t* __restrict__ rng, int ndisp, int order, const int32_t* __restrict__ n_good,
                     const int32_t* __restrict__ flag, SegLayout lay, int32_t* __restrict__ seg_need) {
    __shared__ unsigned long long part[8];
    const int f = blockIdx.x;
    unsigned long long acc = 0;
    for (int c = threadIdx.x; c < ndisp; c += blockDim.x) {
        const size_t col = (size_t)f * ndisp + c;
        const int n = n_good[col];
        if (flag[col] == kSkyModelled && n >= order + 1) acc += (unsigned long long)n * (kBootstrap / 2);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long accepts = 0;
        for (int w = 0; w < 8; ++w) accepts += part[w];
        // attempts per accept: 4 / pi = 1.2732; the sum over 1e7+ accepts is sharp, 2^20 words of slack cover it
        unsigned long long total = rng[(size_t)f * 625 + 624] + 4ull * (unsigned long long)((double)accepts * 1.2733) + kSegSlack;
        unsigned long long need = (total + lay.stride - 1) >> lay.log2_stride;
        if (need > (unsigned long long)lay.segs) need = lay.segs;
        if (need < 1) need = 1;
        seg_need[f] = (int32_t)need;
    }
}

// x1, x2, r2 of the attempt whose four tempered words are w[0..3]; the arithmetic of numpy's legacy_gauss
__device__ __forceinline__ bool gauss_attempt(const uint32_t* __restrict__ w, double* x1, double* x2, double* r2) {
    const double u1 = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
    const double u2 = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) / 9007199254740992.0;
    *x1 = 2.0 * u1 - 1.0;
    *x2 = 2.0 * u2 - 1.0;
    *r2 = *x1 * *x1 + *x2 * *x2;
    return *r2 < 1.0 && *r2 != 0.0;
}

// attempt a of a frame = stream words pos0 + 4 a .. + 3.  One thread per attempt: bitmap word per warp,
// accepted-attempt count per block of 1024 attempts.
__global__ void __launch_bounds__(256)
poly_attempts_kernel(const uint32_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng,
                     const int32_t* __restrict__ seg_need, uint32_t* __restrict__ bitmap, size_t bitmap_len,
                     uint32_t* __restrict__ block_cnt, size_t block_len) {
    const int f = blockIdx.y;
    const uint32_t pos0 = rng[(size_t)f * 625 + 624];
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;        // stream words [0, limit) exist
    const unsigned long long a = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long p = (unsigned long long)pos0 + 4ull * a;
    if (((unsigned long long)blockIdx.x * blockDim.x) * 4ull + pos0 >= limit) return;            // whole CTA past the end
    bool acc = false;
    if (p + 3 < limit) {
        const uint32_t* w = stream + (size_t)f * lay.stream_len + p;
        double x1, x2, r2;
        acc = gauss_attempt(w, &x1, &x2, &r2);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0) {
        bitmap[(size_t)f * bitmap_len + (a >> 5)] = bal;
        if (bal) atomicAdd(&block_cnt[(size_t)f * block_len + (a / kAttemptBlock)], (uint32_t)__popc(bal));
    }
}

// index of the attempt that is the frame's accept number `ordinal` (0-based), from the exclusive block prefix
__device__ __forceinline__ unsigned long long attempt_of_accept(unsigned long long ordinal, const uint32_t* __restrict__ prefix,
                                                                int nblocks, const uint32_t* __restrict__ bitmap) {
    int lo = 0, hi = nblocks - 1;                 // last block whose prefix <= ordinal
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((unsigned long long)prefix[mid] <= ordinal) lo = mid; else hi = mid - 1;
    }
    unsigned long long left = ordinal - prefix[lo];
    const uint32_t* w = bitmap + (size_t)lo * (kAttemptBlock / 32);
    for (int i = 0; i < kAttemptBlock / 32; ++i) {
        uint32_t bits = w[i];
        const unsigned c = (unsigned)__popc(bits);
        if (left < c) {
            for (unsigned k = 0; k < (unsigned)left; ++k) bits &= bits - 1;
            return (unsigned long long)lo * kAttemptBlock + 32ull * i + (unsigned)(__ffs(bits) - 1);
        }
        left -= c;
    }
    return ~0ull;                                  // fewer accepts than planned: reported by the caller
}

// One CTA per frame: exclusive prefix of the block counts (in place), then per column the attempt range
// [first, last) it consumes, and the generator position after the frame.
__global__ void __launch_bounds__(1024)
poly_bounds_kernel(SegLayout lay, const uint32_t* __restrict__ rng, const int32_t* __restrict__ seg_need, int ndisp,
                   int order, const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                   const uint32_t* __restrict__ bitmap, size_t bitmap_len, uint32_t* __restrict__ block_cnt,
                   size_t block_len, unsigned long long* __restrict__ col_first, unsigned long long* __restrict__ col_last,
                   uint32_t* __restrict__ end_pos, FrameState* __restrict__ st) {
    __shared__ unsigned long long carry;
    __shared__ unsigned long long wsum[32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t pos0 = rng[(size_t)f * 625 + 624];
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;
    const int nblocks = (int)((((unsigned long long)(limit > pos0 ? limit - pos0 : 0) / 4ull) + kAttemptBlock - 1) / kAttemptBlock);
    uint32_t* cnt = block_cnt + (size_t)f * block_len;
    const uint32_t* bits = bitmap + (size_t)f * bitmap_len;
    // ---- exclusive prefix over the blocks, 1024 at a time
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + tid;
        const unsigned long long v = (i < nblocks) ? cnt[i] : 0u;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long below = carry;
        for (int w = 0; w < warp; ++w) below += wsum[w];
        if (i < nblocks) cnt[i] = (uint32_t)(below + incl - v);
        __syncthreads();
        if (tid == 1023) carry = below + incl;
        __syncthreads();
    }
    const unsigned long long total_accepts = carry;
    // ---- columns: ordinal of each column's first accept = 50 x (good pixels of the modelled columns before it)
    unsigned long long run = 0;
    __syncthreads();
    for (int base = 0; base < ndisp; base += 1024) {
        const int c = base + tid;
        const size_t col = (size_t)f * ndisp + c;
        unsigned long long mine = 0;
        if (c < ndisp && flag[col] == kSkyModelled && n_good[col] >= order + 1) mine = (unsigned long long)n_good[col] * (kBootstrap / 2);
        unsigned long long incl = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long below = run;
        for (int w = 0; w < warp; ++w) below += wsum[w];
        unsigned long long all = run;
        for (int w = 0; w < 32; ++w) all += wsum[w];
        if (c < ndisp) {
            const unsigned long long first_ord = below + incl - mine;        // accepts consumed before this column
            unsigned long long a0 = 0, a1 = 0;
            if (mine) {
                if (first_ord + mine > total_accepts) {
                    atomicMin(&st[f].status, (int)MOTES_E_INTERNAL);
                } else {
                    a0 = first_ord ? attempt_of_accept(first_ord - 1, cnt, nblocks, bits) + 1 : 0;
                    a1 = attempt_of_accept(first_ord + mine - 1, cnt, nblocks, bits) + 1;
                }
            }
            col_first[col] = a0;
            col_last[col] = a1;
        }
        run = all;
        __syncthreads();
    }
    // generator position after the frame: just behind the last attempt of the last drawing column
    if (tid == 0) {
        unsigned long long last = 0;
        if (run > 0 && run <= total_accepts) last = attempt_of_accept(run - 1, cnt, nblocks, bits) + 1;
        end_pos[f] = pos0 + (uint32_t)(4ull * last);
    }
}

// Persistent CTAs over (frame, column) items; each owns one slot of y scratch (100 x nspat doubles).
__global__ void __launch_bounds__(kBootThreads)
sky_poly_column_kernel(const uint32_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng, int nframes,
                       int nspat, int ndisp, int order, int spatial_floor, const double* __restrict__ errs,
                       size_t frame_stride, const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                       const double* __restrict__ sorted, const uint16_t* __restrict__ rank_of,
                       const uint16_t* __restrict__ good_row, const unsigned long long* __restrict__ col_first,
                       const unsigned long long* __restrict__ col_last, double* __restrict__ yscratch,
                       double* __restrict__ model, double* __restrict__ model_err, const FrameState* __restrict__ st,
                       unsigned int* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char spc_smem[];
    PolyShared sh;
    sh.q = reinterpret_cast<double*>(spc_smem);
    sh.rmat = sh.q + (size_t)nspat * kMaxPoly;
    sh.colscale = sh.rmat + kMaxPoly * kMaxPoly;
    sh.coef = sh.colscale + kMaxPoly;
    sh.loc = sh.coef + kBootstrap * kMaxPoly;
    sh.scl = sh.loc + nspat;
    sh.red = sh.scl + nspat;
    sh.samples = sh.red + 40;
    sh.ring = reinterpret_cast<uint32_t*>(sh.samples + (size_t)(kBootThreads / 32) * (kBootstrap + 4));     // unused here
    sh.counts = sh.ring;                                                                                // 16 words
    sh.ctl = reinterpret_cast<int*>(sh.counts + 16);
    sh.rows = reinterpret_cast<uint16_t*>(sh.ctl + 4);
    __shared__ unsigned int s_item;
    BlockScope sc{(int)threadIdx.x};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* counts8 = reinterpret_cast<unsigned char*>(sh.counts);
    const unsigned long long below_mask = (1ull << (8 * warp)) - 1ull;
    double* y = yscratch + (size_t)blockIdx.x * kBootstrap * nspat;
    const unsigned int total = (unsigned int)nframes * (unsigned int)ndisp;
    const int p = order + 1;
    int parity = 0;
    while (true) {
        if (tid == 0) s_item = atomicAdd(queue, 1u);
        __syncthreads();
        const unsigned int item = s_item;
        __syncthreads();
        if (item >= total) break;
        const int f = item / ndisp, c = item % ndisp;
        const size_t col = (size_t)item;
        if (st[f].status != 0 || flag[col] != kSkyModelled) continue;
        const int n = n_good[col];
        double* fmodel = model + (size_t)f * frame_stride;
        double* fmodel_err = model_err + (size_t)f * frame_stride;
        if (n < p) {
            for (int i = tid; i < nspat; i += kBootThreads) { fmodel[(size_t)i * ndisp + c] = NAN; fmodel_err[(size_t)i * ndisp + c] = NAN; }
            continue;
        }
        for (int g = tid; g < n; g += kBootThreads) {
            const int row = (int)good_row[col * nspat + g];
            sh.rows[g] = (uint16_t)row;
            sh.loc[g] = sorted[col * nspat + rank_of[col * nspat + g]];
            sh.scl[g] = errs[(size_t)f * frame_stride + (size_t)row * ndisp + c];
        }
        __syncthreads();
        // ---- the column's attempts, 256 at a time, in stream order
        const uint32_t pos0 = rng[(size_t)f * 625 + 624];
        const uint32_t* words = stream + (size_t)f * lay.stream_len + pos0;
        const int need = n * (kBootstrap / 2);
        int have = 0;
        for (unsigned long long a = col_first[col]; a < col_last[col]; a += kBootThreads) {
            const unsigned long long mine = a + (unsigned)tid;
            bool acc = false;
            double x1 = 0.0, x2 = 0.0, r2 = 0.0;
            if (mine < col_last[col]) acc = gauss_attempt(words + 4ull * mine, &x1, &x2, &r2);
            const unsigned ballot = __ballot_sync(0xffffffffu, acc);
            if (lane == 0) counts8[parity * 8 + warp] = (unsigned char)__popc(ballot);
            __syncthreads();
            const unsigned long long all = *reinterpret_cast<const unsigned long long*>(counts8 + parity * 8);
            const unsigned long long low = all & below_mask;
            const int total_acc = (int)(__vsadu4((unsigned)all, 0u) + __vsadu4((unsigned)(all >> 32), 0u));
            const int off = (int)(__vsadu4((unsigned)low, 0u) + __vsadu4((unsigned)(low >> 32), 0u)) +
                            __popc(ballot & ((1u << lane) - 1u));
            parity ^= 1;
            const int ord = have + off;
            if (acc && ord < need) {
                const double fr = sqrt(-2.0 * log(r2) / r2);
                const int j = ord / (kBootstrap / 2), smp = 2 * (ord % (kBootstrap / 2));
                y[(size_t)j * kBootstrap + smp] = sh.loc[j] + sh.scl[j] * (fr * x2);         // returned first
                y[(size_t)j * kBootstrap + smp + 1] = sh.loc[j] + sh.scl[j] * (fr * x1);     // the cached deviate
            }
            have += total_acc;
        }
        __syncthreads();
        sky_poly_fit_column(sc, sh, n, p, nspat, ndisp, c, spatial_floor, y, fmodel, fmodel_err);
    }
}

// ---- export -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mt_export_kernel(uint32_t* __restrict__ rng, const uint32_t* __restrict__ snaps, SegLayout lay,
                 const uint32_t* __restrict__ end_pos, const FrameState* __restrict__ st) {
    __shared__ uint32_t ring[kRing];
    const int f = blockIdx.x;
    uint32_t* state = rng + (size_t)f * 625;
    const uint32_t E = end_pos[f];
    if (st[f].status != 0 || E == state[624] || E == 0) return;
    const uint32_t block = ((E - 1) / 624u) * 624u;
    const uint32_t q = block >> kSnapLog2;
    const uint32_t base = q << kSnapLog2;
    const uint32_t* from = snaps + ((size_t)f * lay.snaps + q) * 624;
    for (int i = threadIdx.x; i < 624; i += blockDim.x) ring[(base + i) & (kRing - 1)] = from[i];
    __syncthreads();
    const uint32_t stop = block + 624;
    for (uint32_t gen = base + 624; gen < stop; gen += kSweep) {
        const uint32_t n = gen + threadIdx.x;
        if (threadIdx.x < kSweep && n < stop)
            ring[n & (kRing - 1)] = mt_mix(ring[(n - 624) & (kRing - 1)], ring[(n - 623) & (kRing - 1)],
                                           ring[(n - kSweep) & (kRing - 1)]);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 624; i += blockDim.x) state[i] = ring[(block + i) & (kRing - 1)];
    if (threadIdx.x == 0) state[624] = E - block;
}

}  // namespace motes
