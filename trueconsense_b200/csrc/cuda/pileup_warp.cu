// pileup_warp.cu — variant 3 of the pileup kernel: barrier-free, one warp = one stream of reads.
//
// Same bit-parallel idea as variant 2 (pileup_swar.cu): BAM's base codes are one-hot (A=1, C=2, G=4,
// T=8), so once a read's bases sit on the reference grid as 4-bit codes, counting A/C/G/T per column
// is a positional popcount down the reads — a Harley-Seal carry-save tree of LOP3s into bit-sliced
// counters that live in registers.  What changed is everything around it (profiles/r1_v24_swar.md:
// 70 % of variant 2's instructions were divergent CIGAR / shift-change handling, at 8 warps per SM
// behind CTA-wide barriers):
//
//   partition   the start-sorted reads are cut into one contiguous range per warp of the grid (one
//               CTA per SM, all resident); a warp never synchronises with another warp.
//   window      a warp keeps its counters for the 8*WC reference columns [w0, w0 + 8*WC); reads are
//               taken 32 at a time while their starts stay inside [w0, w0 + slack), slack = 8*WC -
//               (longest reference span) - 8, so every column they touch is inside the window.  When a
//               read starts beyond it the counters are flushed to HBM (one red.add per non-zero cell)
//               and the window moves.  Deep data flushes once per thousands of reads.
//   stage       the packed SEQ words and CIGAR ops of the <= 32 reads are contiguous in HBM: the warp
//               copies them to its shared-memory slice as 16-byte vectors, byte-swapping SEQ words to
//               "first base in the top nibble"; the next sub-tile's lines are prefetched to L2.  Codes that
//               are not one-hot (N, IUPAC) are detected on the way (3 ops per word) and cleared afterwards,
//               one dirty vector per lane and iteration.
//   emit        one lane per read, walk and expansion fused (emit_read_common): lanes stay in step on
//               chunks of match ops — the part of an M/=/X op inside four consecutive 8-column row words —
//               consuming the I / D / S ops in front of it first (deletion columns and insertion anchors,
//               sparse, go to packed 16+16-bit shared counters with one red.shared.add), then five source
//               words, four funnel shifts under the op's shift D (query index = column + D), head / tail
//               masks, four red.shared.or into the lane's own row.  Nothing is written to be re-read.
//               Sub-tiles containing pads or zero-length ops (rare) take the general three-pass form instead
//               (walk_read_general + expand_rows: htslib's full look-ahead state).
//   column sum  lane j owns row words j, j+32, ...: 32 rows go through two 16-input carry-save trees
//               into the bit-sliced counters (4 + HI planes per word, in registers across sub-tiles).
//   span pass   when the caller bounds the longest reference span (tc_reads_t.max_ref_span) the kernel also
//               checks sort order and range and adds the two ends of every read's span to the coverage
//               difference array (combined inside the warp with match.any) — no separate pass over the CIGARs.
#include <limits.h>

#include "pileup.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MIN_SLACK = 64;
constexpr int HI_PLANES = 7;            // bit-sliced planes above "eights": 15 + 16*127 = 2047 reads per run
constexpr int RUN_CAP = 2047;
#ifndef TC_CHUNK_WORDS
#define TC_CHUNK_WORDS 6
#endif
constexpr int CHUNK_WORDS = TC_CHUNK_WORDS;     // row words per emitted chunk of a match op
// per op: bit0 = M/=/X, bit1 = consumes reference, bit2 = consumes query (MIDNSHP=X -> 0..8)
constexpr uint32_t OPFLAGS = 7u | (4u << 3) | (2u << 6) | (2u << 9) | (4u << 12) | (7u << 21) | (7u << 24);
// clamped shift: ops 11..15 (not defined by BAM) index past the table and read 0
__device__ __forceinline__ uint32_t op_flags(uint32_t op) { return __funnelshift_rc(OPFLAGS, 0u, 3u * op) & 7u; }

template <int WC> struct geom {
    static constexpr int ROWW = WC * 8;             // window / row width in reference columns
    static constexpr int RS = WC + 1;               // padded row stride (words)
    static constexpr int NW = WC / 32;              // row words owned by one lane in the column sum
    static constexpr int SEQ_PAD = 4;               // zero words in front of the staged SEQ stream
    static constexpr int SEQ_CAP = WC * 27;         // staged SEQ words per sub-tile (WC=64: 32 reads of 432 bases)
    static constexpr int CIG_CAP = WC * 14;         // staged CIGAR ops per sub-tile
    static constexpr int SEQ_WORDS = SEQ_PAD + SEQ_CAP + 8;
    static constexpr int CIG_WORDS = CIG_CAP + 8;
    static constexpr int WARP_WORDS = 32 * RS + SEQ_WORDS + CIG_WORDS;
    static constexpr int WARPS = (227 * 1024 / 4) / WARP_WORDS > 16 ? 16 : (227 * 1024 / 4) / WARP_WORDS;
};

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    l = a ^ b ^ c;
    h = (a & b) | (c & (a | b));
}

// non-zero <=> some nibble of w has two or more bits set
__device__ __forceinline__ uint32_t multibit(uint32_t w) { return w & ((w | 0x88888888u) - 0x11111111u); }

__device__ __forceinline__ uint32_t clear_multibit(uint32_t w) {
    const uint32_t z = multibit(w);
    const uint32_t m = (z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x11111111u;
    return w & ~(m * 15u);
}

struct walk_out { int nd, b_first, last_end, x_end; };

extern __shared__ __align__(16) uint32_t smem[];

// Shared memory is addressed through 32-bit shared-window byte addresses and explicit ld/st/red.shared:
// the lane-private rows, descriptor lists and counters are indexed with data-dependent offsets in every
// inner loop, and this keeps each access at one address add + one LDS/STS/ATOMS.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// X and I events go straight to a packed global counter per column (X count in the low, I count in the high 32
// bits): they are sparse (a few per read), and keeping them out of shared memory buys another resident warp.
// One add serves both kinds: +1 (a deletion column), +2^32 (an insertion anchor), +2^32 - 1 (an anchor on a
// deletion's last column: that column reads "*+n..", no longer "*").
__device__ __forceinline__ void red_xi(unsigned long long* p, unsigned long long v) { asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
constexpr unsigned long long XI_X = 1ull, XI_I = 1ull << 32, XI_I_MINUS_X = (1ull << 32) - 1ull;
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Walk, general form (pads / zero-length ops present in the sub-tile): one op per iteration with the full
// look-ahead state of htslib's resolve_cigar2.  One lane = one read.  cs: word offset of the read's staged
// CIGAR ops (overwritten by head-fragment descriptors), x: start column inside the window, lq: l_seq,
// row: shared address of the lane's row, xi: the packed X|I counters of the window's first column, xi_n: columns left.
template <int ROWW>
__device__ __forceinline__ walk_out walk_read_general(const uint32_t cs, int nops, int x, const int lq, const uint32_t row, unsigned long long* xi, const int xi_n, int* err) {
    const bool has_seq = lq != 0;           // SEQ '*': every base reads 'N' — events and coverage only
    int y = 0, nd = 0, b_first = INT_MAX, last_end = 0;
    int pend = 0, lastcol = 0;
    bool last_was_d = false;
    const int kmax = __reduce_max_sync(FULL, nops);
    for (int k = 0; k < kmax; ++k) {
        if (k < nops) {
            const uint32_t c = lds(cs + 4 * k);
            const uint32_t op = c & 15u;
            const int l = (int)(c >> 4);
            const uint32_t fl = op_flags(op);
            const int e = x + l;
            if ((fl & 1u) && has_seq && l > 0 && (y + l > lq || e > ROWW)) { atomicCAS(err, 0, TC_ERR_CAPACITY); nops = 0; }   // CIGAR longer than SEQ
            else if ((fl & 1u) && has_seq && l > 0) {
                const int D = y - x;
                const int fw8 = (x + 7) & ~7;
                if (fw8 < e) sts(row + (fw8 >> 1), (uint32_t)e | ((uint32_t)D << 11));
                if (x & 7) {
                    const int fe = min(e, (x | 7) + 1);
                    sts(cs + 4 * nd, (uint32_t)x | ((uint32_t)(fe - x) << 11) | ((uint32_t)D << 14));
                    ++nd;
                }
                b_first = min(b_first, x);
                last_end = e;
            }
            if (op == OP_D && e <= ROWW)
                for (int col = x; col < e; ++col) red_xi(xi + min(col, xi_n - 1), XI_X);
            if (op == OP_I) {
                if (pend && l > 0) {
                    if (lastcol >= 0 && lastcol < ROWW) red_xi(xi + min(lastcol, xi_n - 1), last_was_d ? XI_I_MINUS_X : XI_I);
                    pend = 0;
                }
            } else if (op == OP_P) {
                if (pend == 1) pend = 2;
            } else if (!(fl & 2u)) {
                if (pend == 1) pend = 0;
            }
            if (fl & 2u) { pend = 1; last_was_d = (op == OP_D); lastcol = e - 1; }
            x = (fl & 2u) ? e : x;
            y += (fl & 4u) ? l : 0;
            if (x > ROWW) { atomicCAS(err, 0, TC_ERR_CAPACITY); nops = 0; nd = 0; b_first = INT_MAX; }
        }
    }
    walk_out o; o.nd = nd; o.b_first = b_first; o.last_end = last_end; o.x_end = x;
    return o;
}

// Expansion passes of the general form (after walk_read_general).  A: one row word per iteration; a non-zero word
// is the (end column, shift) of a new regime; the word becomes funnelshift(source words under D) cut at the end.
// B: head fragments (an M op starting inside a row word), OR-ed into the lane's own row.
template <int WC>
__device__ __forceinline__ void expand_rows(const walk_out wo, const uint32_t row, const uint32_t sq, const uint32_t cs) {
    // ---- expand A: one row word per iteration; a non-zero word starts a new regime (end column, shift)
    {
        const bool any_m = wo.b_first != INT_MAX;
        const int o0 = any_m ? wo.b_first >> 3 : 0;
        const int words = any_m ? ((wo.last_end - 1) >> 3) - o0 + 1 : 0;
        const int itmax = __reduce_max_sync(FULL, words);
        // Every lane runs all itmax iterations, no branches: past its own last word a lane points at its
        // row's pad word (always zero: no new regime, and the regime in force has ended, so it stores zero).
        // Source addresses under a regime that has ended stay inside the warp's shared-memory slice.
        uint32_t rp = row + 4u * o0;            // the row word of this iteration
        const uint32_t rpad = row + 4u * WC;
        uint32_t sp = sq + 4u * o0;             // the SEQ word under it (+ 4 * (D >> 3) of the regime in force)
        int sh4 = 0, rem4 = 0, left = words;    // rem4: 4 * (columns the regime still covers from this word's first column)
#pragma unroll 4
        for (int it = 0; it < itmax; ++it) {
            const uint32_t rq = left > 0 ? rp : rpad;
            const uint32_t t = lds(rq);
            if (t) {
                const int D = (int)t >> 11;
                rem4 = 4 * ((int)(t & 0x7ffu) - 8 * (o0 + it));
                sp = sq + 4u * (o0 + it) + (uint32_t)((D >> 3) << 2);
                sh4 = (D & 7) << 2;
            }
            const uint32_t v = __funnelshift_l(lds(sp + 4), lds(sp), sh4) & ~__funnelshift_rc(0xffffffffu, 0u, max(rem4, 0));
            sts(rq, v);
            rem4 -= 32; rp += 4; sp += 4; --left;
        }
    }
    // ---- expand B: head fragments (an M op starting inside a row word), OR-ed into the lane's own row
    {
        const int itmax = __reduce_max_sync(FULL, wo.nd);
        for (int it = 0; it < itmax; ++it) {
            if (it < wo.nd) {
                const uint32_t d = lds(cs + 4 * it);
                const int b = (int)(d & 0x7ffu), flen = (int)((d >> 11) & 7u), D = (int)d >> 14;
                const int o = b >> 3, kb = b & 7;
                const int q0 = 8 * o + D;
                const uint32_t s = sq + (uint32_t)((q0 >> 3) << 2);
                const uint32_t v = __funnelshift_l(lds(s + 4), lds(s), (q0 & 7) << 2);
                const uint32_t m = (0xffffffffu >> (4 * kb)) & ~__funnelshift_rc(0xffffffffu, 0u, 4 * (kb + flen));
                sts(row + 4 * o, lds(row + 4 * o) | (v & m));
            }
        }
    }
}

// Common form (no pads, no zero-length ops, every read has SEQ), fused: walk and expansion in one pass, no
// descriptors.  One lane = one read; lanes are kept in
// step on CHUNKS of match ops: a chunk is the part of an M/=/X op that falls into CW consecutive row words
// (at most 8 * CW - (x & 7) columns; CW = 6 measured best on ONT-like reads, 4 and 8 within 3 %).  Per outer iteration a lane first consumes the ops in front of its next match
// op (typically one I or D: the sparse X / I events), then emits one chunk: CW + 1 source words, CW funnel
// shifts under the op's shift D (query index = column + D), head / tail masks, CW red.shared.or into its own
// row.  Straight-line, nothing is written to be re-read by a later pass.
// sq: shared address of the read's first staged SEQ word.  Returns the read's end column (x after the last op).
template <int ROWW, int CW>
__device__ __forceinline__ int emit_read_common(const uint32_t cs, const int nops, int x, const int y0, const int lq, const uint32_t row,
                                                unsigned long long* xi, const int xi_n, const uint32_t sq, int* err) {
    int y = y0, rem = 0;       // y0: query index of the first base relative to sq's first nibble (pieces of long reads: 0..7)
    uint32_t cp = cs;                       // next op
    const uint32_t cend = cs + 4u * (uint32_t)nops;
    // the next two ops are fetched ahead of their use (the addresses stay inside the slice: the staged CIGARs are
    // followed by padding words)
    uint32_t c0 = lds(cs), c1 = lds(cs + 4);
    // what the op before the current one was: bit 1 = it consumed the reference, bit 0 = it was a deletion
    uint32_t prev = 0;
    while (__any_sync(FULL, cp < cend || rem > 0)) {
        // Between two chunks a lane consumes at most one op that is not a match (typically the I or D between two
        // match ops) and then starts the match op behind it: straight-line code for the common "M I M D M" shape;
        // a second op in a row that is not a match (S I, D I, ...) simply waits for the next iteration.
        if (rem == 0 && cp < cend) {
            uint32_t c = c0;
            uint32_t fl = op_flags(c & 15u);
            if (!(fl & 1u)) {
                const uint32_t op = c & 15u;
                const int l = (int)(c >> 4);
                // one add serves both events: a deletion's first column (+1 X), or an insertion's anchor
                // column x-1 (+1 I, and -1 X when that column belongs to a deletion: it reads "*+n..", not "*").
                // Without zero-length ops the anchor exists whenever the previous op consumed the reference; a
                // column past the reference is clamped (such a read is a TC_ERR_RANGE, the counts are void).
                const bool is_d = (op == OP_D);
                if (is_d || (op == OP_I && (prev & 2u)))
                    red_xi(xi + min(is_d ? x : x - 1, xi_n - 1), is_d ? XI_X : ((prev & 1u) ? XI_I_MINUS_X : XI_I));
                if (is_d && l > 1) for (int col = x + 1; col < min(x + l, xi_n); ++col) red_xi(xi + col, XI_X);
                prev = (fl & 2u) | (is_d ? 1u : 0u);
                x += (fl & 2u) ? l : 0;
                y += (fl & 4u) ? l : 0;
                cp += 4;
                c = c1;
                fl = op_flags(c & 15u);
            }
            if ((fl & 1u) && cp < cend) {
                const int l = (int)(c >> 4);
                // beyond the window, or a CIGAR that consumes more query than SEQ holds: the scatter kernel's business
                if (x + l > ROWW || y + l > lq) { atomicCAS(err, 0, TC_ERR_CAPACITY); cp = cend; }
                else { rem = l; cp += 4; }
            }
            c0 = lds(cp); c1 = lds(cp + 4);
        }
        __syncwarp();           // lanes leave the loop above at different points: emit the chunks together
        if (rem > 0) {
            const int kb = x & 7;
            const int cl = min(rem, 8 * CW - kb);
            {
                const int q0 = (x - kb) + (y - x);                  // query index under the first column of row word x >> 3
                const uint32_t s = sq + (uint32_t)((q0 >> 3) << 2);
                const int sh4 = (q0 & 7) << 2;
                uint32_t w[CW + 1];
#pragma unroll
                for (int j = 0; j <= CW; ++j) w[j] = lds(s + 4 * j);
                const int rem4 = 4 * (kb + cl);                     // 4 * columns from the first row word's start to the chunk's end
                const uint32_t ro = row + (uint32_t)((x >> 3) << 2);
                reds_or(ro, __funnelshift_l(w[1], w[0], sh4) & (0xffffffffu >> (4 * kb)) & ~__funnelshift_rc(0xffffffffu, 0u, rem4));
#pragma unroll
                for (int j = 1; j < CW; ++j)
                    reds_or(ro + 4 * j, __funnelshift_l(w[j + 1], w[j], sh4) & ~__funnelshift_rc(0xffffffffu, 0u, max(rem4 - 32 * j, 0)));
            }
            x += cl; y += cl; rem -= cl;
            prev = 2u;
        }
    }
    if (x > ROWW) atomicCAS(err, 0, TC_ERR_CAPACITY);       // deletions / skips ran past the window (their event columns were clamped)
    return x;
}

template <int WC, bool PIECES>
__global__ void __launch_bounds__(geom<WC>::WARPS * 32, 1) warp_pileup_kernel(pileup_args a) {
    using G = geom<WC>;
    constexpr int ROWW = G::ROWW, RS = G::RS, NW = G::NW;
    const int lane = threadIdx.x & 31;
    const int L = a.L;

    // which geometry handles this batch: the narrowest whose slack is usable
    // the longest reference span: the caller's bound (tc_reads_t.max_ref_span) or what the span pass found
    // (PIECES: the "reads" are pieces of at most PIECE_COLS columns; the span pass ran over the real reads)
    const bool fold = !PIECES && a.span_hint > 0;   // no span pass ran: this kernel also does its checks and the coverage ends
    const int ms = PIECES ? PIECE_COLS : ((max(fold ? a.span_hint : a.status->max_span, 1) + 7) & ~7);
    // 256-column windows (short reads: half the shared memory per warp, more warps per SM), 512 or 1024
    const int slack32 = (256 - ms - 8) & ~7, slack64 = (512 - ms - 8) & ~7, slack128 = (1024 - ms - 8) & ~7;
    const bool mine = (WC == 32) ? (slack32 >= MIN_SLACK)
                    : (WC == 64) ? (slack32 < MIN_SLACK && slack64 >= MIN_SLACK) : (slack64 < MIN_SLACK);
    if (!mine) return;
    const int slack = (WC == 32) ? slack32 : (WC == 64) ? slack64 : slack128;
    if (slack < MIN_SLACK) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
        return;
    }

    // the warp's slice of shared memory, as shared-window byte addresses
    const uint32_t rows = (uint32_t)__cvta_generic_to_shared(smem) + 4u * (threadIdx.x >> 5) * G::WARP_WORDS;   // [32][RS]
    const uint32_t seq_s = rows + 4u * 32 * RS;                 // seq_s + 4 * SEQ_PAD <-> SEQ word sbase_al
    const uint32_t cig_s = seq_s + 4u * G::SEQ_WORDS;
    for (int i = lane; i < 32 * RS; i += 32) sts(rows + 4 * i, 0u);
    if (lane < G::SEQ_PAD) sts(seq_s + 4 * lane, 0u);
    __syncwarp();

    const int64_t n_reads = PIECES ? a.n_pieces : a.r.n;
    const int64_t n_seq_words = (int64_t)a.r.seq_off[a.r.n];
    const int64_t n_ops_total = PIECES ? 0 : (int64_t)a.r.cigar_off[a.r.n];
    const int64_t gw = (int64_t)blockIdx.x * G::WARPS + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * G::WARPS;
    int64_t r = gw * n_reads / n_warps;
    const int64_t r_end = (gw + 1) * n_reads / n_warps;

    uint32_t ones[NW], twos[NW], fours[NW], eights[NW], hi[NW][HI_PLANES];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        ones[j] = twos[j] = fours[j] = eights[j] = 0;
#pragma unroll
        for (int p = 0; p < HI_PLANES; ++p) hi[j][p] = 0;
    }
    int w0 = INT_MIN, run_reads = 0;
    const uint32_t row = rows + 4u * lane * RS;
    int prev_pos = (fold && r > 0 && r < r_end) ? a.r.pos[r - 1] : INT_MIN;     // sort-order check across sub-tiles
    int my_max_span = 0, my_zero_span = 0;

    auto flush = [&]() {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            uint32_t planes[4 + HI_PLANES];
            planes[0] = ones[j]; planes[1] = twos[j]; planes[2] = fours[j]; planes[3] = eights[j];
#pragma unroll
            for (int p = 0; p < HI_PLANES; ++p) planes[4 + p] = hi[j][p];
            uint32_t any = 0;
#pragma unroll
            for (int p = 0; p < 4 + HI_PLANES; ++p) any |= planes[p];
            if (any) {
#pragma unroll 4
                for (int bit = 0; bit < 32; ++bit) {
                    if (!((any >> bit) & 1u)) continue;
                    int v = 0;
#pragma unroll
                    for (int p = 0; p < 4 + HI_PLANES; ++p) v |= (int)((planes[p] >> bit) & 1u) << p;
                    const int colr = w0 + 8 * (lane + 32 * j) + (7 - (bit >> 2));
                    const int cls = bit & 3;     // bit 0 A, 1 C, 2 G, 3 T (BAM codes 1,2,4,8)
                    const int crow = cls == 0 ? TC_ROW_A : cls == 1 ? TC_ROW_C : cls == 2 ? TC_ROW_G : TC_ROW_T;
                    if (colr < L) atomicAdd(&a.counts[(size_t)crow * L + colr], v);
                }
            }
            ones[j] = twos[j] = fours[j] = eights[j] = 0;
#pragma unroll
            for (int p = 0; p < HI_PLANES; ++p) hi[j][p] = 0;
        }
        __syncwarp();
    };

    while (r < r_end) {
        const int nmax = (int)min((int64_t)32, r_end - r);
        const bool valid = lane < nmax;
        const int64_t ri = r + (valid ? lane : 0);
        int p, lq, n, nops_lane, y0 = 0;
        bool passes = true, cig_exotic = false;
        uint32_t cs, sq_lane;
        if constexpr (PIECES) {
            // ---- pieces of long reads: one record per lane, gathered through the start-sorted order
            const tc_piece* pc = a.pieces + a.piece_order[ri];
            const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(pc)), m1 = __ldg(reinterpret_cast<const uint4*>(pc) + 1);
            p = (int)m0.x; lq = (int)m1.y; y0 = (int)m1.z;
            passes = p < L;
            const uint32_t so = m0.y, co = m0.z;
            nops_lane = (int)m0.w;
            const int p0 = __shfl_sync(FULL, p, 0);
            const bool fresh = (w0 == INT_MIN);
            if (fresh) w0 = max(p0, 0) & ~7;
            // every piece is staged on its own, as 16-byte vectors from the aligned-down start of its words / ops
            const int sv = (int)(((so & 3u) + m1.x + 3u) >> 2), cv = (int)(((co & 3u) + m0.w + 3u) >> 2);
            int s_inc = valid ? sv : 0, c_inc = valid ? cv : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ts = __shfl_up_sync(FULL, s_inc, o), tc = __shfl_up_sync(FULL, c_inc, o);
                if (lane >= o) { s_inc += ts; c_inc += tc; }
            }
            const bool inwin = p >= w0 && p - w0 < slack;
            const bool fits = valid && inwin && 4 * s_inc <= G::SEQ_CAP && 4 * c_inc <= G::CIG_CAP;
            const unsigned fm = __ballot_sync(FULL, fits);
            n = (fm == FULL) ? 32 : __ffs(~fm) - 1;
            n = min(n, RUN_CAP - run_reads);
            if (n == 0) {
                const bool inwin0 = __shfl_sync(FULL, (int)inwin, 0) != 0;
                if ((!inwin0 && !fresh) || run_reads >= RUN_CAP) {
                    flush();
                    w0 = INT_MIN; run_reads = 0;
                    continue;
                }
                if (lane == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);     // one piece larger than the staging buffers
                r += 1;
                continue;
            }
            const uint32_t sdst = seq_s + 4u * G::SEQ_PAD + 16u * (uint32_t)(s_inc - sv);
            const uint32_t cdst = cig_s + 16u * (uint32_t)(c_inc - cv);
            if (lane < n) {
                const int64_t sb = (int64_t)(so & ~3u);
                for (int v = 0; v < sv; ++v) {
                    uint4 q;
                    if (sb + 4 * v + 4 <= n_seq_words) q = __ldg(reinterpret_cast<const uint4*>(a.r.seq4 + sb) + v);
                    else {              // the last words of the batch: do not read past the array
                        q.x = sb + 4 * v + 0 < n_seq_words ? __ldg(a.r.seq4 + sb + 4 * v + 0) : 0u;
                        q.y = sb + 4 * v + 1 < n_seq_words ? __ldg(a.r.seq4 + sb + 4 * v + 1) : 0u;
                        q.z = sb + 4 * v + 2 < n_seq_words ? __ldg(a.r.seq4 + sb + 4 * v + 2) : 0u;
                        q.w = sb + 4 * v + 3 < n_seq_words ? __ldg(a.r.seq4 + sb + 4 * v + 3) : 0u;
                    }
                    if (multibit(q.x) | multibit(q.y) | multibit(q.z) | multibit(q.w)) {
                        q.x = clear_multibit(q.x); q.y = clear_multibit(q.y); q.z = clear_multibit(q.z); q.w = clear_multibit(q.w);
                    }
                    q.x = __byte_perm(q.x, 0, 0x0123); q.y = __byte_perm(q.y, 0, 0x0123);
                    q.z = __byte_perm(q.z, 0, 0x0123); q.w = __byte_perm(q.w, 0, 0x0123);
                    sts4(sdst + 16 * v, q);
                }
                // the piece-CIGAR buffer is padded to whole vectors behind its last op
                const uint4* csrc = reinterpret_cast<const uint4*>(a.r.cigar + (co & ~3u));
                for (int v = 0; v < cv; ++v) sts4(cdst + 16 * v, __ldg(csrc + v));
            }
            __syncwarp();
            cs = cdst + 4u * (co & 3u);
            sq_lane = sdst + 4u * (so & 3u);
        } else {
            // ---- metadata of the next (up to) 32 reads, one per lane
            p = a.r.pos[ri];
            const uint32_t so = a.r.seq_off[ri], so_next = a.r.seq_off[ri + 1];
            const uint32_t co = a.r.cigar_off[ri], co_next = a.r.cigar_off[ri + 1];
            const uint32_t flg = a.r.flag[ri];
            lq = a.r.l_seq[ri];
            passes = !(flg & (a.flag_filter | 4u)) && !(a.ignore_orphans && (flg & 1u) && !(flg & 2u));
            if (a.min_mapq > 0 && a.r.mapq && (int)a.r.mapq[ri] < a.min_mapq) passes = false;
            if (p >= L) {       // starts past the reference: TC_ERR_RANGE (the span pass says so; without one, here) — never walked
                if (fold && valid && passes) atomicCAS(&a.status->err, 0, TC_ERR_RANGE);
                passes = false;
            }
            const int p0 = __shfl_sync(FULL, p, 0);
            const bool fresh = (w0 == INT_MIN);
            if (fresh) w0 = max(p0, 0) & ~7;
            const uint32_t sbase_al = __shfl_sync(FULL, so, 0) & ~3u;
            const uint32_t cbase_al = __shfl_sync(FULL, co, 0) & ~3u;
            const bool inwin = p >= w0 && p - w0 < slack;
            const bool fits = valid && inwin && (so_next - sbase_al + 1 <= (uint32_t)G::SEQ_CAP) && (co_next - cbase_al <= (uint32_t)G::CIG_CAP);
            const unsigned fm = __ballot_sync(FULL, fits);
            n = (fm == FULL) ? 32 : __ffs(~fm) - 1;
            n = min(n, RUN_CAP - run_reads);
            if (n == 0) {
                const bool inwin0 = __shfl_sync(FULL, (int)inwin, 0) != 0;
                if ((!inwin0 && !fresh) || run_reads >= RUN_CAP) {      // the window (or the counters' range) is used up
                    flush();
                    w0 = INT_MIN; run_reads = 0;
                    continue;
                }
                // a negative position (TC_ERR_RANGE), or one read larger than the staging buffers
                if (inwin0 && lane == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
                if (fold && lane == 0) {
                    if (p0 < prev_pos) atomicCAS(&a.status->err, 0, TC_ERR_UNSORTED);
                    if (p0 < 0 && passes) atomicCAS(&a.status->err, 0, TC_ERR_RANGE);
                }
                prev_pos = p0;
                r += 1;
                continue;
            }

            // ---- stage SEQ words [sbase_al, send) and CIGAR ops [cbase_al, cend)
            {
                const uint32_t send = __shfl_sync(FULL, so_next, n - 1) + 1;        // one word of look-ahead for the funnel shift
                const int nv = (int)((send - sbase_al + 3) >> 2);
                const uint32_t cend = __shfl_sync(FULL, co_next, n - 1);
                const int ncv = (int)((cend - cbase_al + 3) >> 2);
                // the next sub-tile starts where this one ends and is about as large: pull its lines (and the
                // metadata lines two sub-tiles ahead) towards L2 while this one is being processed
                {
                    const int64_t s_lo = (int64_t)send - 1, s_len = (int64_t)send - sbase_al;
                    for (int64_t o = 32 * lane; o < s_len && s_lo + o < n_seq_words; o += 1024) prefetch_l2(a.r.seq4 + s_lo + o);
                    const int64_t c_len = (int64_t)cend - cbase_al;
                    for (int64_t o = 32 * lane; o < c_len && (int64_t)cend + o < n_ops_total; o += 1024) prefetch_l2(a.r.cigar + cend + o);
                    const int64_t rm = r + n + 64;
                    if (rm < n_reads) {
                        if (lane == 0) prefetch_l2(a.r.pos + rm);
                        if (lane == 1) prefetch_l2(a.r.seq_off + rm);
                        if (lane == 2) prefetch_l2(a.r.cigar_off + rm);
                        if (lane == 3) prefetch_l2(a.r.l_seq + rm);
                        if (lane == 4) prefetch_l2(a.r.flag + rm);
                    }
                }
                uint32_t dirty = 0;          // dirty: bit t <=> the lane's t-th vector holds a code that is not one-hot
                const uint32_t sdst = seq_s + 4u * G::SEQ_PAD;
                if ((int64_t)sbase_al + 4ll * nv <= n_seq_words) {
                    const uint4* src = reinterpret_cast<const uint4*>(a.r.seq4 + sbase_al);
                    int t = 0;
                    for (int i = lane; i < nv; i += 128, t += 4) {
                        uint4 v[4];
    #pragma unroll
                        for (int u = 0; u < 4; ++u) if (i + 32 * u < nv) v[u] = __ldg(src + i + 32 * u);
    #pragma unroll
                        for (int u = 0; u < 4; ++u) if (i + 32 * u < nv) {
                            const uint32_t z = multibit(v[u].x) | multibit(v[u].y) | multibit(v[u].z) | multibit(v[u].w);
                            dirty |= (z != 0 ? 1u : 0u) << (t + u);
                            v[u].x = __byte_perm(v[u].x, 0, 0x0123); v[u].y = __byte_perm(v[u].y, 0, 0x0123);
                            v[u].z = __byte_perm(v[u].z, 0, 0x0123); v[u].w = __byte_perm(v[u].w, 0, 0x0123);
                            sts4(sdst + 16 * (i + 32 * u), v[u]);
                        }
                    }
                } else {            // the last sub-tile of the batch: do not read past the array
                    int t = 0;
                    for (int i = lane; i < nv; i += 32, ++t) {
                        uint32_t w[4];
    #pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int64_t wi = (int64_t)sbase_al + 4 * i + u;
                            w[u] = wi < n_seq_words ? __ldg(a.r.seq4 + wi) : 0u;
                        }
                        const uint32_t z = multibit(w[0]) | multibit(w[1]) | multibit(w[2]) | multibit(w[3]);
                        dirty |= (z != 0 ? 1u : 0u) << t;
    #pragma unroll
                        for (int u = 0; u < 4; ++u) sts(sdst + 16 * i + 4 * u, __byte_perm(w[u], 0, 0x0123));
                    }
                }
                // pads and zero-length ops: min over the words of min((op ^ P), len) is 0 exactly when one is present
                uint32_t exmin = 1u;
                if ((int64_t)cbase_al + 4ll * ncv <= n_ops_total) {
                    const uint4* csrc = reinterpret_cast<const uint4*>(a.r.cigar + cbase_al);
                    for (int i = lane; i < ncv; i += 64) {
                        uint4 v0 = __ldg(csrc + i), v1 = make_uint4(16u, 16u, 16u, 16u);
                        const bool b1 = i + 32 < ncv;
                        if (b1) v1 = __ldg(csrc + i + 32);
                        exmin = min(exmin, min(min(min((v0.x & 15u) ^ 6u, v0.x >> 4), min((v0.y & 15u) ^ 6u, v0.y >> 4)),
                                               min(min((v0.z & 15u) ^ 6u, v0.z >> 4), min((v0.w & 15u) ^ 6u, v0.w >> 4))));
                        exmin = min(exmin, min(min(min((v1.x & 15u) ^ 6u, v1.x >> 4), min((v1.y & 15u) ^ 6u, v1.y >> 4)),
                                               min(min((v1.z & 15u) ^ 6u, v1.z >> 4), min((v1.w & 15u) ^ 6u, v1.w >> 4))));
                        sts4(cig_s + 16 * i, v0);
                        if (b1) sts4(cig_s + 16 * (i + 32), v1);
                    }
                } else {
                    for (int i = lane; i < 4 * ncv; i += 32) {
                        const int64_t oi = (int64_t)cbase_al + i;
                        const uint32_t c = oi < n_ops_total ? __ldg(a.r.cigar + oi) : 16u;
                        exmin = min(exmin, min((c & 15u) ^ 6u, c >> 4));
                        sts(cig_s + 4 * i, c);
                    }
                }
                // ... and reads without SEQ ('*': every base reads 'N'): the general form below handles all of these
                cig_exotic = __any_sync(FULL, exmin == 0u || (lane < n && lq == 0));
                // some base is N / IUPAC (rare in real reads): clear those codes — they only count towards coverage.
                // Only the vectors that hold one are revisited, one per lane and iteration.
                while (__any_sync(FULL, dirty != 0)) {
                    if (dirty) {
                        const int t = __ffs(dirty) - 1;
                        dirty &= dirty - 1;
                        const uint32_t q = sdst + 16 * (lane + 32 * t);
                        uint4 v = lds4(q);
                        v.x = clear_multibit(v.x); v.y = clear_multibit(v.y); v.z = clear_multibit(v.z); v.w = clear_multibit(v.w);
                        sts4(q, v);
                    }
                }
                __syncwarp();
            }
            cs = cig_s + 4u * (co - cbase_al);
            nops_lane = (int)(co_next - co);
            sq_lane = seq_s + 4u * G::SEQ_PAD + 4u * (so - sbase_al);
        }
        // ---- walk
        const bool act = lane < n && passes;
        const int nops = act ? nops_lane : 0;
        const int x0 = p - w0;
        const uint32_t sq = act ? sq_lane : seq_s + 4u * G::SEQ_PAD;
        if (!act) cs = cig_s;           // lanes without a read still run the loops' first fetch: keep it inside the slice
        int x_end;
        if (!cig_exotic) {
            x_end = emit_read_common<ROWW, CHUNK_WORDS>(cs, nops, x0, y0, lq, row, a.xi + w0, L - w0, sq, &a.status->err);
        } else {
            // pads or zero-length ops somewhere in the sub-tile (rare): the general three-pass form
            const walk_out wo = walk_read_general<ROWW>(cs, nops, x0, lq, row, a.xi + w0, L - w0, &a.status->err);
            x_end = wo.x_end;
            expand_rows<WC>(wo, row, sq, cs);
        }

        // ---- without a span pass: sort order, range, span statistics and the two ends of every read's span in the
        // coverage difference array (adds to the same column are combined inside the warp first)
        if (fold) {
            int pprev = __shfl_up_sync(FULL, p, 1);
            if (lane == 0) pprev = prev_pos;
            if (lane < n && p < pprev) atomicCAS(&a.status->err, 0, TC_ERR_UNSORTED);
            prev_pos = __shfl_sync(FULL, p, n - 1);
            const int span = act ? x_end - x0 : 0;
            const bool bad = act && (p >= L || p + span > L);
            if (bad) atomicCAS(&a.status->err, 0, TC_ERR_RANGE);
            const bool has = act && span > 0 && !bad;
            my_max_span = max(my_max_span, span);
            my_zero_span += (act && span == 0) ? 1 : 0;
            const unsigned g0 = __match_any_sync(FULL, has ? p : -1 - lane);
            if (has && (__ffs(g0) - 1) == lane) atomicAdd(&a.diff[p], __popc(g0));
            const unsigned g1 = __match_any_sync(FULL, has ? p + span : -1 - lane);
            if (has && (__ffs(g1) - 1) == lane) atomicAdd(&a.diff[p + span], -__popc(g1));
        }

        __syncwarp();

        // ---- column sum: lane owns row words lane + 32 j, all 32 rows
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const uint32_t col = rows + 4u * (lane + 32 * j);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                if (blk * 16 < n) {
                    uint32_t xw[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) { xw[q] = lds(col + 4u * (blk * 16 + q) * RS); sts(col + 4u * (blk * 16 + q) * RS, 0u); }
                    uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteens;
                    csa(twosA, ones[j], ones[j], xw[0], xw[1]);
                    csa(twosB, ones[j], ones[j], xw[2], xw[3]);
                    csa(foursA, twos[j], twos[j], twosA, twosB);
                    csa(twosA, ones[j], ones[j], xw[4], xw[5]);
                    csa(twosB, ones[j], ones[j], xw[6], xw[7]);
                    csa(foursB, twos[j], twos[j], twosA, twosB);
                    csa(eightsA, fours[j], fours[j], foursA, foursB);
                    csa(twosA, ones[j], ones[j], xw[8], xw[9]);
                    csa(twosB, ones[j], ones[j], xw[10], xw[11]);
                    csa(foursA, twos[j], twos[j], twosA, twosB);
                    csa(twosA, ones[j], ones[j], xw[12], xw[13]);
                    csa(twosB, ones[j], ones[j], xw[14], xw[15]);
                    csa(foursB, twos[j], twos[j], twosA, twosB);
                    csa(eightsB, fours[j], fours[j], foursA, foursB);
                    csa(sixteens, eights[j], eights[j], eightsA, eightsB);
                    uint32_t carry = sixteens;
#pragma unroll
                    for (int pl = 0; pl < HI_PLANES; ++pl) { const uint32_t t = hi[j][pl] & carry; hi[j][pl] ^= carry; carry = t; }
                }
            }
        }
        __syncwarp();
        r += n;
        run_reads += n;
    }
    if (run_reads > 0) flush();
    if (fold) {
        my_max_span = __reduce_max_sync(FULL, my_max_span);
        my_zero_span = __reduce_add_sync(FULL, my_zero_span);
        if (lane == 0) {
            if (my_max_span > 0) atomicMax(&a.status->max_span, my_max_span);
            if (my_zero_span > 0) atomicAdd(&a.status->n_zero_span, my_zero_span);
        }
    }
}

}  // namespace

bool tc_pileup_warp_supported(const pileup_args& a) {
    return (((uintptr_t)a.r.seq4 | (uintptr_t)a.r.cigar) & 15u) == 0;
}

template <int WC, bool PIECES>
static int launch_geom(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    using G = geom<WC>;
    const size_t smem = sizeof(uint32_t) * (size_t)G::WARP_WORDS * G::WARPS;
    TC_CUDA(cudaFuncSetAttribute(warp_pileup_kernel<WC, PIECES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    warp_pileup_kernel<WC, PIECES><<<ctx->sm_count, G::WARPS * 32, smem, s>>>(a);
    TC_LAUNCH_CHECK();
    return TC_OK;
}

// All geometries are enqueued; each reads the longest reference span the span pass left in
// a.status and returns at once unless it is the one that fits (no host round trip in between).
int tc_pileup_warp_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    int rc = launch_geom<32, false>(ctx, a, s);
    if (rc) return rc;
    rc = launch_geom<64, false>(ctx, a, s);
    if (rc) return rc;
    return launch_geom<128, false>(ctx, a, s);
}

// pieces of long reads (pileup_long.cu): always the 512-column geometry
int tc_pileup_warp_launch_pieces(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) { return launch_geom<64, true>(ctx, a, s); }
