// Position walk of the order-0 bootstrap for SMALL batches and LONG frames (sky.py:34-38: where in numpy's generator
// stream does every column's np.random.choice start?).
//
// sky_walk_kernel (skyseg.cuh) streams a frame's whole 16-bit generator stream through one SM and counts the accepted
// draws of ~27 pieces per column: with 128 frames in flight that runs at the HBM roofline, but one frame alone pays
// the serial per-column latency (2.2 us per column: 9 ms for 4096 columns, 45 ms for the 24000 columns of an X-shooter
// order).  Here the counting is taken off the serial path:
//
//   walk_range_kernel   per frame: the range [n_lo, n_hi] of good-sky-pixel counts (one rejection mask, <= 64 values)
//   walk_table_kernel   all SMs, one warp per 2048-word piece: how many words of the piece a column with n good pixels
//                       accepts, for every n of the range (a 64-entry u16 row per piece) - one pass over the stream at
//                       full bandwidth
//   walk_serial_kernel  one warp per frame: per column the partial first piece is counted from the piece already in
//                       registers (it is the piece the previous column ended in), the whole pieces come from the table
//                       (one lookup per lane, one warp prefix), and only the piece the last draw falls in is read -
//                       prefetched from its predicted position a column ahead
//
// Same outputs as sky_walk_kernel (col_start / col_end / running counts at piece boundaries / end position), bit for
// bit; frames the table cannot serve (a wider range of n, two masks) are left to sky_walk_kernel.
#pragma once

#include "skyseg.cuh"

namespace motes {

__global__ void __launch_bounds__(256)
walk_range_kernel(int ndisp, const int32_t* __restrict__ n_good, const int32_t* __restrict__ flag,
                  const FrameState* __restrict__ st, WalkTabInfo* __restrict__ info) {
    __shared__ int s_lo[8], s_hi[8];
    const int f = blockIdx.x, tid = threadIdx.x;
    int lo = 0x7fffffff, hi = 0;
    for (int c = tid; c < ndisp; c += 256) {
        const size_t col = (size_t)f * ndisp + c;
        const int n = n_good[col];
        if (flag[col] == kSkyModelled && n > 1) { lo = min(lo, n); hi = max(hi, n); }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
        WalkTabInfo t;
        if (hi < lo) {                              // no column draws anything: the serial kernel only copies positions
            t.n_lo = 2; t.range = 1; t.mask = 1u; t.ok = 1;
        } else {
            t.n_lo = lo;
            t.range = hi - lo + 1;
            t.mask = rejection_mask((uint32_t)lo);
            t.ok = (t.range <= kTabRange) && (rejection_mask((uint32_t)hi) == t.mask) && (hi - 1 < 0x8000);
        }
        if (st[f].status != 0) t.ok = 0;
        info[f] = t;
    }
}

// One warp per piece.  tab[piece][r][j] = words in rows 0 .. r of the piece (256 words per row, positions < limit)
// whose masked value is < n_lo + j, i.e. that a column with n = n_lo + j good pixels accepts.
constexpr int kTabPiece = kPieceRows * kTabRange;          // u16 entries per piece (1 KB)
__global__ void __launch_bounds__(256)
walk_table_kernel(const uint16_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng,
                  const int32_t* __restrict__ seg_need, const WalkTabInfo* __restrict__ info, size_t pieces_per_frame,
                  uint16_t* __restrict__ table) {
    __shared__ uint32_t hist[8][kPieceRows][kTabRange];
    const int f = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WalkTabInfo t = info[f];
    if (!t.ok) return;
    const uint32_t piece = blockIdx.x * 8u + (uint32_t)warp;
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;
    const uint32_t sb = piece << kSubLog2;
    const uint32_t P0 = rng[(size_t)f * 625 + 624];
    if (piece >= pieces_per_frame || sb >= limit || sb + (uint32_t)kSubChunk <= P0) return;
    const uint4* src = reinterpret_cast<const uint4*>(stream + (size_t)f * lay.stream_len + sb);
    uint4 d[kPieceRows];
#pragma unroll
    for (int k = 0; k < kPieceRows; ++k) d[k] = __ldg(src + k * 32 + lane);
#pragma unroll
    for (int k = 0; k < kPieceRows; ++k) { hist[warp][k][lane] = 0u; hist[warp][k][lane + 32] = 0u; }
    __syncwarp();
    const uint32_t n_lo = (uint32_t)t.n_lo, n_hi = (uint32_t)(t.n_lo + t.range - 1);      // thresholds n_lo .. n_hi
    const uint32_t mask2 = t.mask * 0x00010001u;
    const uint32_t kk_lo = (0x8000u - n_lo) * 0x00010001u, kk_hi = (0x8000u - n_hi) * 0x00010001u;
    const bool whole = sb + (uint32_t)kSubChunk <= limit;
    int below[kPieceRows];
#pragma unroll
    for (int k = 0; k < kPieceRows; ++k) {
        const uint32_t pos0 = sb + (uint32_t)(k * 256 + lane * 8);
        int c_lo, c_hi;
        if (whole) {
            c_lo = walk_count8_packed(d[k], mask2, kk_lo);      // values <= n_lo - 1
            c_hi = walk_count8_packed(d[k], mask2, kk_hi);      // values <= n_hi - 1
        } else {
            c_lo = walk_count8(d[k], pos0, 0u, limit, t.mask, n_lo - 1u);
            c_hi = walk_count8(d[k], pos0, 0u, limit, t.mask, n_hi - 1u);
        }
        below[k] = c_lo;
        if (c_hi != c_lo) {                                     // a word in the band [n_lo, n_hi): rare (range / (mask + 1))
            const uint32_t w[4] = {d[k].x, d[k].y, d[k].z, d[k].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t v = ((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu) & t.mask;
                if (v >= n_lo && v < n_hi && pos0 + (uint32_t)e < limit) atomicAdd(&hist[warp][k][v - n_lo], 1u);
            }
        }
    }
    __syncwarp();
    // per row: entry j = below + sum_{k < j} hist[k] (each lane takes entries 2 lane and 2 lane + 1), then cumulated over the rows
    uint32_t run0 = 0, run1 = 0;
    uint32_t* out = reinterpret_cast<uint32_t*>(table + ((size_t)f * pieces_per_frame + piece) * kTabPiece);
#pragma unroll
    for (int k = 0; k < kPieceRows; ++k) {
        const uint32_t row_below = (uint32_t)__reduce_add_sync(0xffffffffu, below[k]);
        const uint32_t h0 = hist[warp][k][2 * lane], h1 = hist[warp][k][2 * lane + 1];
        uint32_t incl = h0 + h1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        const uint32_t e0 = row_below + incl - (h0 + h1);
        run0 += e0;
        run1 += e0 + h0;
        out[k * (kTabRange / 2) + lane] = (run0 & 0xffffu) | (run1 << 16);
    }
}

// The lane's 8 words of a 256-word row (row start `rs`): reject flags (word e: bit (e & 1) * 16 + 15 - (e >> 1)), words
// outside [from, limit) rejected.
// reject bits of the first k (0..8) words of a lane: even words 0,2,4,6 -> bits 15..12, odd words 1,3,5,7 -> bits 31..28
__device__ __forceinline__ uint32_t first_words_mask(uint32_t k) {
    const uint32_t even = (k + 1u) >> 1, odd = k >> 1;              // how many even / odd words are among the first k
    const uint32_t me = (0xF0000u >> even) & 0xF000u, mo = (0xF0000u >> odd) & 0xF000u;
    return me | (mo << 16);
}
__device__ __forceinline__ uint32_t row_rejects(const uint4& d, uint32_t rs, int lane, uint32_t from, uint32_t limit, uint32_t mask2,
                                                uint32_t kk2) {
    const uint32_t r0 = ((d.x & mask2) + kk2) & 0x80008000u, r1 = ((d.y & mask2) + kk2) & 0x80008000u;
    const uint32_t r2 = ((d.z & mask2) + kk2) & 0x80008000u, r3 = ((d.w & mask2) + kk2) & 0x80008000u;
    uint32_t r = r0 | (r1 >> 1) | (r2 >> 2) | (r3 >> 3);
    const uint32_t pos0 = rs + (uint32_t)(lane * 8);
    // words in front of `from`: the first min(8, from - pos0) of this lane
    const uint32_t lead = from > pos0 ? min(from - pos0, 8u) : 0u;
    r |= first_words_mask(lead);
    if (pos0 + 8u > limit) {                                        // the end of the stream: rare
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (pos0 + (uint32_t)e >= limit) r |= 1u << ((e & 1) * 16 + 15 - (e >> 1));
    }
    return r;
}
// position after the target-th (1-based) accepted word of the row, 0 when the row holds fewer
__device__ __forceinline__ uint32_t row_locate(uint32_t rej, uint32_t rs, int lane, uint32_t target) {
    const uint32_t mine = 8u - (uint32_t)__popc(rej);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned holder = __ballot_sync(0xffffffffu, incl - mine < target && target <= incl);
    if (!holder) return 0u;
    const int hl = __ffs(holder) - 1;
    // the holder's flags and its ordinal inside the lane, broadcast: every lane finds the word (no divergence)
    const uint32_t hrej = __shfl_sync(0xffffffffu, rej, hl);
    const uint32_t want = target - __shfl_sync(0xffffffffu, incl - mine, hl);       // 1..8
    // accepted flags of the holder's words: even words 0, 2, 4, 6 = bits 15 .. 12 of ~hrej, odd words 1, 3, 5, 7 = bits 31 .. 28;
    // e = words in front of the want-th accepted one
    const uint32_t okb = ~hrej;
    uint32_t e = 0, seen = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        seen += (okb >> (15 - i)) & 1u;
        e += (seen < want) ? 1u : 0u;
        seen += (okb >> (31 - i)) & 1u;
        e += (seen < want) ? 1u : 0u;
    }
    return rs + (uint32_t)(hl * 8) + e + 1u;
}

constexpr int kCandRows = 5;                               // rows fetched around the predicted position of a column's last draw

// What one column asks for on behalf of the next one, staged in shared memory by cp.async: register loads carried across
// loop iterations share the warp's few scoreboards with the loads issued after them, so that their consumers wait for the
// NEW loads as well (ncu: two long-scoreboard stalls of an L2 latency each per column); asynchronous copies complete on
// their own and are collected with one wait at the start of the column.
struct WalkStage {
    uint32_t start_rows[8];        // table rows of the piece the column starts in (the word holding entry j)
    uint32_t last_rows[8];         // table rows of the piece its last draw is predicted in
    uint32_t pieces[32];           // whole-piece entries of the 32 pieces behind the start piece
    uint4 cand[kCandRows][32];     // rows of the stream around the predicted position of the last draw
};
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}

// One warp per frame.  Everything the next column needs from memory is requested at one point of the current column -
// as soon as the piece q its last draw falls in is known: the next column starts in q (table rows of q at the next
// threshold), its whole pieces follow q (one table entry per lane), and its own last draw is predicted from the
// expected consumption (table rows of the predicted piece, kCandRows rows of the stream around the predicted position).
// A wrong prediction costs one synchronous reload.
__global__ void __launch_bounds__(32)
walk_serial_kernel(const uint16_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng,
                   const int32_t* __restrict__ seg_need, int ndisp, const int32_t* __restrict__ n_good,
                   const int32_t* __restrict__ flag, const WalkTabInfo* __restrict__ info, size_t pieces_per_frame,
                   const uint16_t* __restrict__ table, uint32_t* __restrict__ col_start, uint32_t* __restrict__ col_end,
                   uint32_t* __restrict__ sub_have, uint32_t* __restrict__ end_pos, FrameState* __restrict__ st) {
    __shared__ __align__(16) WalkStage stage[2];
    const int f = blockIdx.x, lane = threadIdx.x;
    const WalkTabInfo t = info[f];
    if (!t.ok) return;
    const uint16_t* s = stream + (size_t)f * lay.stream_len;
    const uint16_t* tab = table + (size_t)f * pieces_per_frame * kTabPiece;
    uint32_t* subs = sub_have + (size_t)f * lay.sub_len;
    uint32_t* cs = col_start + (size_t)f * ndisp;
    uint32_t* ce = col_end + (size_t)f * ndisp;
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;
    const uint32_t last_piece = (limit - 1u) >> kSubLog2;
    uint32_t P = rng[(size_t)f * 625 + 624];
    const uint32_t mask2 = t.mask * 0x00010001u;
    // table entry j of a row / piece lives in the 32-bit word (j >> 1) of its 128-byte line
    auto rows_word = [&](uint32_t piece, int j) -> const uint32_t* {
        return reinterpret_cast<const uint32_t*>(tab + (size_t)piece * kTabPiece + (lane & 7) * kTabRange) + (j >> 1);
    };
    auto piece_word = [&](uint32_t piece, int j) -> const uint32_t* {
        return reinterpret_cast<const uint32_t*>(tab + (size_t)piece * kTabPiece + (kPieceRows - 1) * kTabRange) + (j >> 1);
    };
    auto entry = [](uint32_t word, int j) -> uint32_t { return (j & 1) ? word >> 16 : word & 0xffffu; };
    auto tab_rows = [&](uint32_t piece, int j) -> uint32_t {          // lane & 7 = row: cumulative count through that row
        return piece <= last_piece ? entry(*rows_word(piece, j), j) : 0u;
    };
    auto tab_pieces = [&](uint32_t base, int j) -> uint32_t {         // lane = piece base + lane: its whole count
        return base + (uint32_t)lane <= last_piece ? entry(*piece_word(base + (uint32_t)lane, j), j) : 0u;
    };
    auto row_data = [&](uint32_t rs) -> uint4 { return __ldg(reinterpret_cast<const uint4*>(s + rs) + lane); };
    // what the issue point of the previous column requested (into stage[parity])
    int parity = 0;
    int pre_j = -1;                       // threshold index the requests were made for
    uint32_t pre_start = 0xffffffffu;     // piece the column was expected to start in
    uint32_t pre_last = 0xffffffffu;      // predicted piece of the last draw
    uint32_t cand_rs = 0xffffffffu;       // row start of cand[0]
    uint4 rowd = make_uint4(0, 0, 0, 0);  // the row P lies in (it held the previous column's last draw)
    uint32_t rowd_rs = 0xffffffffu;
    bool failed = false;
#ifdef MOTES_WALK_DEBUG
    int dbg_start = 0, dbg_rowd = 0, dbg_last = 0, dbg_cand = 0, dbg_fromP = 0, dbg_cols = 0, dbg_jchange = 0;
#endif
    int n_next = n_good[(size_t)f * ndisp], flag_next = flag[(size_t)f * ndisp];
    for (int c = 0; c < ndisp; ++c) {
        const size_t col = (size_t)f * ndisp + c;
        const int n = n_next, col_flag = flag_next;
        if (c + 1 < ndisp) { n_next = n_good[col + 1]; flag_next = flag[col + 1]; }
        if (failed || col_flag != kSkyModelled || n <= 1) {
            if (lane == 0) { cs[c] = P; ce[c] = P; }
            continue;
        }
        const uint32_t need = (uint32_t)n * kBootstrap, top = (uint32_t)(n - 1);
        const uint32_t kk2 = (0x8000u - (top + 1u)) * 0x00010001u;
        const int j = n - t.n_lo;
        if (lane == 0) cs[c] = P;
        const uint32_t p0 = P >> kSubLog2, r0 = (P >> 8) & 7u, rs0 = P & ~255u;
        if (p0 > last_piece) { failed = true; if (lane == 0) ce[c] = P; continue; }
        // collect what the previous column requested for this one
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const WalkStage& in = stage[parity];
        uint32_t start_rows, pieces_tl;
#ifdef MOTES_WALK_DEBUG
        ++dbg_cols; dbg_jchange += (pre_j != j); dbg_start += (pre_j != j || pre_start != p0);
#endif
        if (pre_j != j || pre_start != p0) {                // nothing (or the wrong thing) was requested: ask now
            start_rows = tab_rows(p0, j);
            pieces_tl = tab_pieces(p0 + 1u, j);
            pre_last = 0xffffffffu;
            cand_rs = 0xffffffffu;
        } else {
            start_rows = entry(in.start_rows[lane & 7], j);
            pieces_tl = p0 + 1u + (uint32_t)lane <= last_piece ? entry(in.pieces[lane], j) : 0u;
        }
        // ---- accepted draws from P to the end of piece p0: the rest of P's row from the row itself, the rows behind it
        // from the table
        uint32_t rej0 = 0xffffffffu;                         // reject flags of P's row (positions >= P), if it was needed
        const uint32_t through_r0 = __shfl_sync(0xffffffffu, start_rows, (int)r0);
        const uint32_t piece_total = __shfl_sync(0xffffffffu, start_rows, kPieceRows - 1);
        uint32_t in_row;
        if ((P & 255u) == 0u) {
            in_row = through_r0 - (r0 ? __shfl_sync(0xffffffffu, start_rows, (int)r0 - 1) : 0u);
        } else {
#ifdef MOTES_WALK_DEBUG
            dbg_rowd += (rowd_rs != rs0);
#endif
            if (rowd_rs != rs0) { rowd = row_data(rs0); rowd_rs = rs0; }
            rej0 = row_rejects(rowd, rs0, lane, P, limit, mask2, kk2);
            in_row = (uint32_t)__reduce_add_sync(0xffffffffu, 8 - __popc(rej0));
        }
        uint32_t have = in_row + (piece_total - through_r0);
        uint32_t cross = 0;
        uint32_t q = p0, target = need;                      // piece of the last draw, and its ordinal counted from ...
        bool from_P = true;                                  // ... P (q == p0) or the start of q
        if (have < need) {
            from_P = false;
            uint32_t base = p0 + 1u;
            uint32_t tl = pieces_tl;
            while (true) {
                uint32_t incl = tl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const bool exists = base + (uint32_t)lane <= last_piece;
                const unsigned reach = __ballot_sync(0xffffffffu, exists && have + incl >= need);
                const int jq = reach ? __ffs(reach) - 1 : 32;
                if (exists && lane <= jq) subs[base + (uint32_t)lane] = have + incl - tl;     // running count at the piece boundaries
                if (reach) {
                    q = base + (uint32_t)jq;
                    target = need - (have + __shfl_sync(0xffffffffu, incl - tl, jq));
                    break;
                }
                have += __shfl_sync(0xffffffffu, incl, 31);
                base += 32u;
                if (base > last_piece) { failed = true; break; }
                tl = tab_pieces(base, j);
            }
        }
        if (failed) {
            if (lane == 0) { ce[c] = P; atomicMin(&st[f].status, (int)MOTES_E_INTERNAL); }
            continue;
        }
        // ---- issue point: q is known.  Requests for the next column (threshold jn): it starts in q ...
#ifdef MOTES_WALK_DEBUG
        dbg_fromP += from_P; dbg_last += (!from_P && pre_last != q);
#endif
        const int jn = (n_next > 1 && flag_next == kSkyModelled) ? n_next - t.n_lo : j;
        WalkStage& out = stage[parity ^ 1];
        if (lane < 8) cp_async4(&out.start_rows[lane], rows_word(q <= last_piece ? q : last_piece, jn));
        cp_async4(&out.pieces[lane], piece_word(q + 1u + (uint32_t)lane <= last_piece ? q + 1u + (uint32_t)lane : last_piece, jn));
        // ... and ends about 100 n (mask + 1) / n words behind this column's last draw, whose offset in q is about
        // target / acceptance
        const float wpd = __fdividef((float)(t.mask + 1u), (float)(jn + t.n_lo));
        const float here = (float)(q << kSubLog2) + (from_P ? (float)(P & (uint32_t)(kSubChunk - 1)) : 0.0f) + (float)target * __fdividef((float)(t.mask + 1u), (float)n);
        const float there = here + (float)((jn + t.n_lo) * kBootstrap) * wpd;
        const uint32_t guess_pos = there < 4.0e9f ? (uint32_t)there : 0xfffffff0u;
        const uint32_t guess_piece = guess_pos >> kSubLog2;
        if (lane < 8) cp_async4(&out.last_rows[lane], rows_word(guess_piece <= last_piece ? guess_piece : last_piece, jn));
        uint32_t next_cand_rs = (guess_pos & ~255u) - 2u * 256u;
        if ((guess_pos & ~255u) < 2u * 256u) next_cand_rs = 0u;
        const bool cand_ok = (size_t)next_cand_rs + (size_t)kCandRows * 256u <= (size_t)lay.stream_len;
        if (cand_ok) {
#pragma unroll
            for (int i = 0; i < kCandRows; ++i)
                cp_async16(&out.cand[i][lane], reinterpret_cast<const uint4*>(s + next_cand_rs + 256u * (uint32_t)i) + lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- the row of q the last draw falls in
        const uint32_t q_rows = from_P ? start_rows
                                       : (pre_last == q ? entry(in.last_rows[lane & 7], j) : tab_rows(q, j));
        uint32_t row, t_row;
        if (from_P) {
            // rows r0 .. 7 counted from P: in_row, then the table's rows
            const uint32_t cum = (lane & 7) >= (int)r0 ? in_row + ((q_rows) - through_r0) : 0u;       // accepted from P through row lane & 7
            const unsigned hit = __ballot_sync(0xffffffffu, lane < kPieceRows && (uint32_t)lane >= r0 && cum >= target) & 0xffu;
            row = hit ? (uint32_t)(__ffs(hit) - 1) : r0;
            const uint32_t before_row = row > r0 ? __shfl_sync(0xffffffffu, cum, (int)row - 1) : 0u;
            t_row = target - before_row;
        } else {
            const unsigned hit = __ballot_sync(0xffffffffu, lane < kPieceRows && q_rows >= target) & 0xffu;
            row = hit ? (uint32_t)(__ffs(hit) - 1) : 0u;
            const uint32_t before_row = row ? __shfl_sync(0xffffffffu, q_rows, (int)row - 1) : 0u;
            t_row = target - before_row;
        }
        const uint32_t rs = (q << kSubLog2) + row * 256u;
        const bool reuse_rej0 = from_P && row == r0 && rej0 != 0xffffffffu;
        if (!reuse_rej0) {
            const uint32_t idx = (rs - cand_rs) >> 8;
            if (rowd_rs != rs) {
                if (cand_rs != 0xffffffffu && rs >= cand_rs && idx < (uint32_t)kCandRows) {
                    rowd = in.cand[idx][lane];
                } else {
#ifdef MOTES_WALK_DEBUG
                    ++dbg_cand;
#endif
                    rowd = row_data(rs);
                }
                rowd_rs = rs;
            }
        }
        cand_rs = cand_ok ? next_cand_rs : 0xffffffffu;
        pre_j = jn;
        pre_start = q;
        pre_last = guess_piece;
        parity ^= 1;
        if (reuse_rej0) {
            cross = row_locate(rej0, rs, lane, t_row);
        } else {
            const uint32_t rej = row_rejects(rowd, rs, lane, from_P && row == r0 ? P : rs, limit, mask2, kk2);
            cross = row_locate(rej, rs, lane, t_row);
        }
        if (cross == 0u) {
            failed = true;
            if (lane == 0) { ce[c] = P; atomicMin(&st[f].status, (int)MOTES_E_INTERNAL); }
            continue;
        }
        P = cross;
        if (lane == 0) ce[c] = P;
        // two columns ahead, towards L2 (the requests above then find their lines there): the table rows of the piece the
        // column after next should start and end in, its per-piece entries, and the stream around its predicted end
        {
            const float step = (float)((jn + t.n_lo) * kBootstrap) * wpd;
            const float far_pos = there + step;
            if (far_pos < (float)limit) {
                const uint32_t far_u = (uint32_t)far_pos;
                const uint32_t far_piece = far_u >> kSubLog2;
                const char* line = reinterpret_cast<const char*>(s + (((size_t)far_u & ~(size_t)255) - 512 + 256 * (size_t)(lane >> 2))) + 128 * (lane & 3);
                if (lane < 4 * kCandRows && far_u >= 512u) asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                // rows of guess_piece (where the column after next starts) and of far_piece (where it should end) ...
                const uint32_t pz = lane < 8 ? guess_piece : lane < 16 ? far_piece : lane < 24 ? far_piece + 1u : guess_piece + 1u;
                const char* rows = reinterpret_cast<const char*>(tab + (size_t)pz * kTabPiece + (lane & 7) * kTabRange);
                if (pz <= last_piece) asm volatile("prefetch.global.L2 [%0];" ::"l"(rows));
                // ... and the entries of the pieces behind guess_piece (its whole pieces)
                const uint32_t pw = guess_piece + 1u + (uint32_t)lane;
                const char* ent = reinterpret_cast<const char*>(tab + (size_t)pw * kTabPiece + (kPieceRows - 1) * kTabRange);
                if (pw <= last_piece) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent));
            }
        }
    }
    if (lane == 0) end_pos[f] = P;
#ifdef MOTES_WALK_DEBUG
    if (lane == 0) printf("walk f=%d cols=%d jchange=%d start_reload=%d rowd_reload=%d last_rows_miss=%d cand_miss=%d fromP=%d range=%d\n", f, dbg_cols, dbg_jchange, dbg_start, dbg_rowd, dbg_last, dbg_cand, dbg_fromP, t.range);
#endif
}

// ---- walk by runs -----------------------------------------------------------------------------------------------------
// Consecutive columns that draw with the same number n of good sky pixels (the sky region moves slowly along the
// dispersion axis, and a column whose sky was not modelled draws nothing) share their acceptance test, so the k-th of
// them ends at the (k + 1) 100 n-th accepted word behind the run's start: the serial dependency is between RUNS, not
// columns.  One warp per frame, up to 32 columns per iteration (lane i looks at column c + i):
//   * A = accepted words of the piece the run starts in that lie in front of its start P (table rows + P's row);
//   * the pieces behind, 32 at a time (one table entry per lane, one warp prefix): every lane knows the accepted count
//     S at the start and end of its piece, hence which targets A + (k + 1) need fall into it (it posts piece and
//     remainder for each) and - at every piece boundary - how many draws the running column has had (sub_have);
//   * lane k resolves target k on its own: the piece's eight cumulative row counts, then a scan of the 256-word row.
// A run of one column costs about what walk_serial_kernel spends on it; a long run costs ~1 / 10 per column.
// Same outputs as walk_serial_kernel / sky_walk_kernel, bit for bit.
__device__ __forceinline__ uint32_t group_rejects(const uint4& d, uint32_t mask2, uint32_t kk2) {
    const uint32_t r0 = ((d.x & mask2) + kk2) & 0x80008000u, r1 = ((d.y & mask2) + kk2) & 0x80008000u;
    const uint32_t r2 = ((d.z & mask2) + kk2) & 0x80008000u, r3 = ((d.w & mask2) + kk2) & 0x80008000u;
    return r0 | (r1 >> 1) | (r2 >> 2) | (r3 >> 3);        // word e: bit (e & 1) * 16 + 15 - (e >> 1)
}
// words in front of the want-th (1..8) accepted one of a group with reject flags `rej`
__device__ __forceinline__ uint32_t nth_accept(uint32_t rej, uint32_t want) {
    const uint32_t okb = ~rej;
    uint32_t e = 0, seen = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        seen += (okb >> (15 - i)) & 1u;
        e += (seen < want) ? 1u : 0u;
        seen += (okb >> (31 - i)) & 1u;
        e += (seen < want) ? 1u : 0u;
    }
    return e;
}

constexpr int kRunWarps = 8;                                // warps per frame: 256 pieces per round of the piece scan

__global__ void __launch_bounds__(kRunWarps * 32)
walk_runs_kernel(const uint16_t* __restrict__ stream, SegLayout lay, const uint32_t* __restrict__ rng,
                 const int32_t* __restrict__ seg_need, int ndisp, const int32_t* __restrict__ n_good,
                 const int32_t* __restrict__ flag, const WalkTabInfo* __restrict__ info, size_t pieces_per_frame,
                 const uint16_t* __restrict__ table, uint32_t* __restrict__ col_start, uint32_t* __restrict__ col_end,
                 uint32_t* __restrict__ sub_have, uint32_t* __restrict__ end_pos, FrameState* __restrict__ st) {
    __shared__ uint32_t t_piece[32], t_rem[32], t_pos[32];
    __shared__ uint32_t chunk_tot[2][kRunWarps];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const WalkTabInfo t = info[f];
    if (!t.ok) return;
    const uint16_t* s = stream + (size_t)f * lay.stream_len;
    const uint16_t* tab = table + (size_t)f * pieces_per_frame * kTabPiece;
    uint32_t* subs = sub_have + (size_t)f * lay.sub_len;
    uint32_t* cs = col_start + (size_t)f * ndisp;
    uint32_t* ce = col_end + (size_t)f * ndisp;
    const int32_t* ng = n_good + (size_t)f * ndisp;
    const int32_t* fl = flag + (size_t)f * ndisp;
    const uint32_t limit = ((uint32_t)seg_need[f] << lay.log2_stride) + 1u;
    const uint32_t last_piece = (limit - 1u) >> kSubLog2;
    const uint32_t mask2 = t.mask * 0x00010001u;
    const unsigned below = (1u << lane) - 1u;
    uint32_t P = rng[(size_t)f * 625 + 624];
    bool failed = false;
    int c = 0;
    while (c < ndisp) {
        // ---- the run: columns c .. c + L - 1, K of them draw (all with the same n).  Every warp works this out for itself.
        const int ci = c + lane;
        const bool inside = ci < ndisp;
        const int n_i = inside ? ng[ci] : 0;
        const bool draws = inside && fl[ci] == kSkyModelled && n_i > 1;
        const unsigned draw_mask = __ballot_sync(0xffffffffu, draws);
        if (failed || draw_mask == 0u) {
            if (warp == 0 && inside) { cs[ci] = P; ce[ci] = P; }
            c += 32;
            continue;
        }
        const int n = __shfl_sync(0xffffffffu, n_i, __ffs(draw_mask) - 1);
        const unsigned other = __ballot_sync(0xffffffffu, draws && n_i != n);
        const int L = other ? __ffs(other) - 1 : min(32, ndisp - c);
        const unsigned members = (L >= 32) ? 0xffffffffu : ((1u << L) - 1u);
        const uint32_t K = (uint32_t)__popc(draw_mask & members);
        const uint32_t my_k = (uint32_t)__popc(draw_mask & members & below);      // drawing columns in front of this one
        const uint32_t need = (uint32_t)n * kBootstrap;
        const uint32_t kk2 = (0x8000u - (uint32_t)n) * 0x00010001u;
        const int j = n - t.n_lo;
        const uint32_t p0 = P >> kSubLog2, r0 = (P >> 8) & 7u, rs0 = P & ~255u;
        auto entry16 = [&](uint32_t piece, uint32_t row) -> uint32_t {              // accepted in rows 0 .. row of the piece
            return (uint32_t)__ldg(tab + (size_t)piece * kTabPiece + row * kTabRange + j);
        };
        bool bad = p0 > last_piece;
        if (!bad) {
            // the first round's entries, then A: accepted words of piece p0 in front of P
            uint32_t piece = p0 + 32u * (uint32_t)warp + (uint32_t)lane;
            uint32_t tl = piece <= last_piece ? entry16(piece, kPieceRows - 1) : 0u;
            const uint32_t rows_before = r0 ? entry16(p0, r0 - 1u) : 0u;
            uint32_t in_front = 0;
            if (P & 255u) {
                const uint4 d = __ldg(reinterpret_cast<const uint4*>(s + rs0) + lane);
                const uint32_t rej = group_rejects(d, mask2, kk2);
                const uint32_t pos0 = rs0 + (uint32_t)lane * 8u;
                const uint32_t lead = P > pos0 ? min(P - pos0, 8u) : 0u;            // this lane's words in front of P
                in_front = lead - (uint32_t)__popc(rej & first_words_mask(lead));
                in_front = (uint32_t)__reduce_add_sync(0xffffffffu, (int)in_front);
            }
            const uint32_t A = rows_before + in_front;
            // ---- the pieces from p0 on, 32 per warp and round: accepted counts S at the start / end of every piece, hence
            // the targets A + (k + 1) need inside it and the draws the running column has had at its boundary
            uint32_t S = 0, base = p0;
            for (int rnd = 0;; ++rnd) {
                uint32_t incl = tl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                if (lane == 31) chunk_tot[rnd & 1][warp] = incl;
                // (the next round's entries are requested before the barrier)
                const uint32_t piece_next = piece + 32u * (uint32_t)kRunWarps;
                const uint32_t tl_next = piece_next <= last_piece ? entry16(piece_next, kPieceRows - 1) : 0u;
                __syncthreads();
                uint32_t off = S, all = 0;
#pragma unroll
                for (int w = 0; w < kRunWarps; ++w) {
                    const uint32_t tw = chunk_tot[rnd & 1][w];
                    if (w < warp) off += tw;
                    all += tw;
                }
                const bool exists = piece <= last_piece;
                const uint32_t s_lo = off + incl - tl, s_hi = off + incl;          // accepted from p0's start to the piece's start / end
                // targets A + (k + 1) need <= s: (s - A) / need of them
                const uint32_t c_lo = s_lo >= A ? min(K, (s_lo - A) / need) : 0u;
                const uint32_t c_hi = s_hi >= A ? min(K, (s_hi - A) / need) : 0u;
                if (exists && piece > p0 && c_lo < K) subs[piece] = s_lo - A - c_lo * need;   // draws of the running column so far
                if (c_lo < c_hi) {
                    // what resolving the targets will read, towards L2: the piece's row counts and the predicted rows of the stream
                    const char* block = reinterpret_cast<const char*>(tab + (size_t)piece * kTabPiece);
#pragma unroll
                    for (int r = 0; r < kPieceRows; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(block + 128 * r));
                    for (uint32_t k = c_lo; k < c_hi; ++k) {
                        const uint32_t rem = A + (k + 1u) * need - s_lo;
                        t_piece[k] = piece;
                        t_rem[k] = rem;
                        const int guess = (int)((rem * (uint32_t)kPieceRows) / (tl + 1u));
#pragma unroll
                        for (int dr = -1; dr <= 1; ++dr) {
                            const int r = guess + dr;
                            if (r >= 0 && r < kPieceRows) {
                                const char* row = reinterpret_cast<const char*>(s + ((size_t)piece << kSubLog2) + (size_t)r * 256);
#pragma unroll
                                for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128 * l));
                            }
                        }
                    }
                }
                S += all;
                const bool done = S >= A && (S - A) / need >= K;                    // uniform over the CTA
                if (done) {
                    // the next run starts in one of these pieces: their entries, towards L2
                    if (exists) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(tab + (size_t)piece * kTabPiece + (kPieceRows - 1) * kTabRange)));
                    if (piece_next <= last_piece) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(tab + (size_t)piece_next * kTabPiece + (kPieceRows - 1) * kTabRange)));
                    break;
                }
                base += 32u * (uint32_t)kRunWarps;
                if (base > last_piece) { bad = true; break; }
                piece = piece_next;
                tl = tl_next;
            }
        }
        __syncthreads();
        // ---- where the targets lie: a warp per target (k = warp, warp + 8, ..), the eight row counts of its piece in lanes
        // 0 .. 7, then the 256-word row, 8 words per lane
        if (!bad) {
            uint32_t cum[4], rem4[4], q4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t k = (uint32_t)(warp + kRunWarps * i);
                q4[i] = k < K ? t_piece[k] : 0u;
                rem4[i] = k < K ? t_rem[k] : 0u;
                cum[i] = (k < K && lane < kPieceRows) ? entry16(q4[i], (uint32_t)lane) : 0xffffffffu;
            }
            uint32_t rs4[4], trow4[4];
            uint4 d4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t k = (uint32_t)(warp + kRunWarps * i);
                const unsigned reach = __ballot_sync(0xffffffffu, lane < kPieceRows && cum[i] >= rem4[i]) & 0xffu;
                const uint32_t row = reach ? (uint32_t)(__ffs(reach) - 1) : (uint32_t)(kPieceRows - 1);
                const uint32_t prev = row ? __shfl_sync(0xffffffffu, cum[i], (int)row - 1) : 0u;
                trow4[i] = rem4[i] - prev;
                rs4[i] = (q4[i] << kSubLog2) + row * 256u;
                d4[i] = k < K ? __ldg(reinterpret_cast<const uint4*>(s + rs4[i]) + lane) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t k = (uint32_t)(warp + kRunWarps * i);
                if (k < K) {                                                        // uniform over the warp
                    const uint32_t rej = row_rejects(d4[i], rs4[i], lane, rs4[i], limit, mask2, kk2);
                    const uint32_t pos = row_locate(rej, rs4[i], lane, trow4[i]);
                    if (lane == 0) t_pos[k] = pos;
                }
            }
        }
        __syncthreads();
        // ---- positions of the run's columns
        uint32_t good = 0;
        if (!bad) {
            const unsigned lost = __ballot_sync(0xffffffffu, (uint32_t)lane < K && t_pos[lane] == 0u);
            good = lost ? (uint32_t)(__ffs(lost) - 1) : K;                        // targets found
        }
        if (warp == 0 && lane < L) {
            const uint32_t kb = min(my_k, good);
            const uint32_t start = kb ? t_pos[kb - 1u] : P;
            cs[ci] = start;
            ce[ci] = (draws && my_k < good) ? t_pos[my_k] : start;
        }
        if (good) P = t_pos[good - 1u];
        if (good < K) {
            failed = true;
            if (threadIdx.x == 0) atomicMin(&st[f].status, (int)MOTES_E_INTERNAL);
        }
        c += L;
    }
    if (threadIdx.x == 0) end_pos[f] = P;
}

}  // namespace motes
