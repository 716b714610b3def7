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
//               "first base in the top nibble".  Codes that are not one-hot (N, IUPAC) are detected on
//               the way (3 ops per word) and cleared in a second pass only when a sub-tile has any.
//   walk        one lane per read, warp-uniform loop over CIGAR ops, straight-line body: every M/=/X op
//               [b, e) with shift D (query index = column + D) leaves (e, D) in the read's row at the
//               first 8-column word it covers the start of, and — when b is not a multiple of 8 — a
//               head-fragment descriptor in place of the consumed ops.  Deletion columns and insertion
//               anchors (sparse) go to packed 16+16-bit shared counters.
//   expand A    one lane per read, one row word per iteration: a non-zero word is the (e, D) of a new
//               regime; the word becomes funnelshift(source words under D) cut at e.  No searching, no
//               inner loop.
//   expand B    one head fragment per iteration, OR-ed into the lane's own row word.
//   column sum  lane j owns row words j, j+32, ...: 32 rows go through two 16-input carry-save trees
//               into the bit-sliced counters (4 + HI planes per word, in registers across sub-tiles).
#include <limits.h>

#include "pileup.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MIN_SLACK = 64;
constexpr int HI_PLANES = 7;            // bit-sliced planes above "eights": 15 + 16*127 = 2047 reads per run
constexpr int RUN_CAP = 2047;
// per op: bit0 = M/=/X, bit1 = consumes reference, bit2 = consumes query (MIDNSHP=X -> 0..8)
constexpr unsigned long long OPFLAGS = 7ull | (4ull << 3) | (2ull << 6) | (2ull << 9) | (4ull << 12) | (7ull << 21) | (7ull << 24);

template <int WC> struct geom {
    static constexpr int ROWW = WC * 8;             // window / row width in reference columns
    static constexpr int RS = WC + 1;               // padded row stride (words)
    static constexpr int NW = WC / 32;              // row words owned by one lane in the column sum
    static constexpr int SEQ_PAD = 4;               // zero words in front of the staged SEQ stream
    static constexpr int SEQ_CAP = WC * 28;         // staged SEQ words per sub-tile (WC=64: 32 reads of 448 bases)
    static constexpr int CIG_CAP = WC * 16;         // staged CIGAR ops per sub-tile
    static constexpr int SEQ_WORDS = SEQ_PAD + SEQ_CAP + 8;
    static constexpr int CIG_WORDS = CIG_CAP + 8;
    static constexpr int WARP_WORDS = 32 * RS + SEQ_WORDS + CIG_WORDS + ROWW;
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

struct walk_out { int nd, b_first, last_end; };

// One lane = one read.  cs: the read's staged CIGAR ops (overwritten by head-fragment descriptors),
// x: start column inside the window, lq: l_seq, row: the lane's row, xi: the warp's packed X|I counters.
template <int ROWW, bool EXOTIC>
__device__ __forceinline__ walk_out walk_read(uint32_t* cs, int nops, int x, const int lq, uint32_t* row, int* xi, int* err) {
    const bool has_seq = lq != 0;           // SEQ '*': every base reads 'N' — events and coverage only
    int y = 0, nd = 0, b_first = INT_MAX, last_end = 0;
    int pend = 0, lastcol = 0;
    bool last_was_d = false;
    const int kmax = __reduce_max_sync(FULL, nops);
    for (int k = 0; k < kmax; ++k) {
        if (k < nops) {
            const uint32_t c = cs[k];
            const uint32_t op = c & 15u;
            const int l = (int)(c >> 4);
            const uint32_t fl = (uint32_t)(OPFLAGS >> (3u * op)) & 7u;
            const int e = x + l;
            if ((fl & 1u) && has_seq && l > 0) {
                const int D = y - x;
                const int fw8 = (x + 7) & ~7;
                if (fw8 < e) row[fw8 >> 3] = (uint32_t)e | ((uint32_t)D << 11);
                if (x & 7) {
                    const int fe = min(e, (x | 7) + 1);
                    cs[nd++] = (uint32_t)x | ((uint32_t)(fe - x) << 11) | ((uint32_t)D << 14);
                }
                b_first = min(b_first, x);
                last_end = e;
            }
            if (op == OP_D && e <= ROWW)
                for (int col = x; col < e; ++col) atomicAdd(&xi[col], 1);
            if (EXOTIC) {
                // general look-ahead state of htslib's resolve_cigar2 (pads, zero-length ops)
                if (op == OP_I) {
                    if (pend && l > 0) {
                        if (lastcol >= 0 && lastcol < ROWW) atomicAdd(&xi[lastcol], last_was_d ? 0xffff : 0x10000);
                        pend = 0;
                    }
                } else if (op == OP_P) {
                    if (pend == 1) pend = 2;
                } else if (!(fl & 2u)) {
                    if (pend == 1) pend = 0;
                }
                if (fl & 2u) { pend = 1; last_was_d = (op == OP_D); lastcol = e - 1; }
            } else {
                // no pads, no zero-length ops: an insertion counts iff the op before it consumes the reference;
                // after a deletion that column reads "*+n.." and is no longer an X (one add does both)
                if (op == OP_I && pend && x >= 1 && x <= ROWW) atomicAdd(&xi[x - 1], last_was_d ? 0xffff : 0x10000);
                pend = (int)(fl & 2u);
                last_was_d = (op == OP_D);
            }
            x = (fl & 2u) ? e : x;
            y += (fl & 4u) ? l : 0;
            if (x > ROWW) { atomicCAS(err, 0, TC_ERR_CAPACITY); nops = 0; nd = 0; b_first = INT_MAX; }
        }
    }
    if (y > lq && b_first != INT_MAX) { atomicCAS(err, 0, TC_ERR_CAPACITY); }    // CIGAR longer than SEQ: scatter kernel
    walk_out o; o.nd = nd; o.b_first = b_first; o.last_end = last_end;
    return o;
}

template <int WC>
__global__ void __launch_bounds__(geom<WC>::WARPS * 32, 1) warp_pileup_kernel(pileup_args a) {
    using G = geom<WC>;
    constexpr int ROWW = G::ROWW, RS = G::RS, NW = G::NW;
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int L = a.L;

    // which geometry handles this batch: the narrowest whose slack is usable
    const int ms = (max(a.status->max_span, 1) + 7) & ~7;
    const int slack64 = (512 - ms - 8) & ~7, slack128 = (1024 - ms - 8) & ~7;
    const bool mine = (WC == 64) ? (slack64 >= MIN_SLACK) : (slack64 < MIN_SLACK);
    if (!mine) return;
    const int slack = (WC == 64) ? slack64 : slack128;
    if (slack < MIN_SLACK) {
        if (threadIdx.x == 0 && blockIdx.x == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
        return;
    }

    uint32_t* rows = smem + (threadIdx.x >> 5) * G::WARP_WORDS;     // [32][RS]
    uint32_t* seq_s = rows + 32 * RS;                               // index SEQ_PAD <-> word sbase_al
    uint32_t* cig_s = seq_s + G::SEQ_WORDS;
    int* xi = (int*)(cig_s + G::CIG_WORDS);                         // [ROWW]  X count | I count << 16
    for (int i = lane; i < 32 * RS; i += 32) rows[i] = 0;
    for (int i = lane; i < ROWW; i += 32) xi[i] = 0;
    if (lane < G::SEQ_PAD) seq_s[lane] = 0;
    __syncwarp();

    const int64_t n_reads = a.r.n;
    const int64_t n_seq_words = (int64_t)a.r.seq_off[n_reads];
    const int64_t n_ops_total = (int64_t)a.r.cigar_off[n_reads];
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
    uint32_t* row = rows + lane * RS;

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
        for (int i = lane; i < ROWW; i += 32) {
            const int v = xi[i];
            if (v) {
                const int xv = v & 0xffff, iv = (int)((uint32_t)v >> 16);
                if (w0 + i < L) {
                    if (xv) atomicAdd(&a.counts[(size_t)TC_ROW_X * L + w0 + i], xv);
                    if (iv) atomicAdd(&a.counts[(size_t)TC_ROW_I * L + w0 + i], iv);
                }
                xi[i] = 0;
            }
        }
        __syncwarp();
    };

    while (r < r_end) {
        // ---- metadata of the next (up to) 32 reads, one per lane
        const int nmax = (int)min((int64_t)32, r_end - r);
        const bool valid = lane < nmax;
        const int64_t ri = r + (valid ? lane : 0);
        const int p = a.r.pos[ri];
        const uint32_t so = a.r.seq_off[ri], so_next = a.r.seq_off[ri + 1];
        const uint32_t co = a.r.cigar_off[ri], co_next = a.r.cigar_off[ri + 1];
        const uint32_t flg = a.r.flag[ri];
        const int lq = a.r.l_seq[ri];
        bool passes = !(flg & (a.flag_filter | 4u)) && !(a.ignore_orphans && (flg & 1u) && !(flg & 2u));
        if (a.min_mapq > 0 && a.r.mapq && (int)a.r.mapq[ri] < a.min_mapq) passes = false;
        const int p0 = __shfl_sync(FULL, p, 0);
        const bool fresh = (w0 == INT_MIN);
        if (fresh) w0 = max(p0, 0) & ~7;
        const uint32_t sbase_al = __shfl_sync(FULL, so, 0) & ~3u;
        const uint32_t cbase_al = __shfl_sync(FULL, co, 0) & ~3u;
        const bool inwin = p >= w0 && p - w0 < slack;
        const bool fits = valid && inwin && (so_next - sbase_al + 1 <= (uint32_t)G::SEQ_CAP) && (co_next - cbase_al <= (uint32_t)G::CIG_CAP);
        const unsigned fm = __ballot_sync(FULL, fits);
        int n = (fm == FULL) ? 32 : __ffs(~fm) - 1;
        n = min(n, RUN_CAP - run_reads);
        if (n == 0) {
            const bool inwin0 = __shfl_sync(FULL, (int)inwin, 0) != 0;
            if ((!inwin0 && !fresh) || run_reads >= RUN_CAP) {      // the window (or the counters' range) is used up
                flush();
                w0 = INT_MIN; run_reads = 0;
                continue;
            }
            // a negative position (flagged by the span pass), or one read larger than the staging buffers
            if (inwin0 && lane == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
            r += 1;
            continue;
        }

        // ---- stage SEQ words [sbase_al, send) and CIGAR ops [cbase_al, cend)
        bool cig_exotic;
        {
            const uint32_t send = __shfl_sync(FULL, so_next, n - 1) + 1;        // one word of look-ahead for the funnel shift
            const int nv = (int)((send - sbase_al + 3) >> 2);
            const uint32_t cend = __shfl_sync(FULL, co_next, n - 1);
            const int ncv = (int)((cend - cbase_al + 3) >> 2);
            uint32_t zacc = 0, exacc = 0;
            if ((int64_t)sbase_al + 4ll * nv <= n_seq_words) {
                const uint4* src = reinterpret_cast<const uint4*>(a.r.seq4 + sbase_al);
                uint4* dst = reinterpret_cast<uint4*>(seq_s + G::SEQ_PAD);
                for (int i = lane; i < nv; i += 128) {
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i + 32 * u < nv) v[u] = __ldg(src + i + 32 * u);
#pragma unroll
                    for (int u = 0; u < 4; ++u) if (i + 32 * u < nv) {
                        zacc |= multibit(v[u].x) | multibit(v[u].y) | multibit(v[u].z) | multibit(v[u].w);
                        v[u].x = __byte_perm(v[u].x, 0, 0x0123); v[u].y = __byte_perm(v[u].y, 0, 0x0123);
                        v[u].z = __byte_perm(v[u].z, 0, 0x0123); v[u].w = __byte_perm(v[u].w, 0, 0x0123);
                        dst[i + 32 * u] = v[u];
                    }
                }
            } else {            // the last sub-tile of the batch: do not read past the array
                for (int i = lane; i < 4 * nv; i += 32) {
                    const int64_t wi = (int64_t)sbase_al + i;
                    const uint32_t w = wi < n_seq_words ? __ldg(a.r.seq4 + wi) : 0u;
                    zacc |= multibit(w);
                    seq_s[G::SEQ_PAD + i] = __byte_perm(w, 0, 0x0123);
                }
            }
            if ((int64_t)cbase_al + 4ll * ncv <= n_ops_total) {
                const uint4* csrc = reinterpret_cast<const uint4*>(a.r.cigar + cbase_al);
                uint4* cdst = reinterpret_cast<uint4*>(cig_s);
                for (int i = lane; i < ncv; i += 64) {
                    uint4 v0 = __ldg(csrc + i), v1 = make_uint4(16u, 16u, 16u, 16u);
                    const bool b1 = i + 32 < ncv;
                    if (b1) v1 = __ldg(csrc + i + 32);
                    exacc |= (uint32_t)((v0.x & 15u) == OP_P) | (uint32_t)(v0.x < 16u) | (uint32_t)((v0.y & 15u) == OP_P) | (uint32_t)(v0.y < 16u) |
                             (uint32_t)((v0.z & 15u) == OP_P) | (uint32_t)(v0.z < 16u) | (uint32_t)((v0.w & 15u) == OP_P) | (uint32_t)(v0.w < 16u) |
                             (uint32_t)((v1.x & 15u) == OP_P) | (uint32_t)(v1.x < 16u) | (uint32_t)((v1.y & 15u) == OP_P) | (uint32_t)(v1.y < 16u) |
                             (uint32_t)((v1.z & 15u) == OP_P) | (uint32_t)(v1.z < 16u) | (uint32_t)((v1.w & 15u) == OP_P) | (uint32_t)(v1.w < 16u);
                    cdst[i] = v0;
                    if (b1) cdst[i + 32] = v1;
                }
            } else {
                for (int i = lane; i < 4 * ncv; i += 32) {
                    const int64_t oi = (int64_t)cbase_al + i;
                    const uint32_t c = oi < n_ops_total ? __ldg(a.r.cigar + oi) : 16u;
                    exacc |= (uint32_t)((c & 15u) == OP_P) | (uint32_t)(c < 16u);
                    cig_s[i] = c;
                }
            }
            cig_exotic = __any_sync(FULL, exacc != 0);
            __syncwarp();
            if (__any_sync(FULL, zacc != 0)) {      // rare: some base is N / IUPAC — clear those codes (they only count towards coverage)
                for (int i = lane; i < 4 * nv; i += 32) seq_s[G::SEQ_PAD + i] = clear_multibit(seq_s[G::SEQ_PAD + i]);
                __syncwarp();
            }
        }

        // ---- walk
        const bool act = lane < n && passes;
        uint32_t* cs = cig_s + (co - cbase_al);
        const int nops = act ? (int)(co_next - co) : 0;
        const int x0 = p - w0;
        walk_out wo;
        if (cig_exotic) wo = walk_read<ROWW, true>(cs, nops, x0, lq, row, xi, &a.status->err);
        else wo = walk_read<ROWW, false>(cs, nops, x0, lq, row, xi, &a.status->err);

        // ---- expand A: one row word per iteration; a non-zero word starts a new regime (end column, shift)
        const uint32_t* sq = seq_s + G::SEQ_PAD + (so - sbase_al);
        {
            const bool any_m = wo.b_first != INT_MAX;
            const int o0 = any_m ? wo.b_first >> 3 : 0;
            const int words = any_m ? ((wo.last_end - 1) >> 3) - o0 + 1 : 0;
            const int itmax = __reduce_max_sync(FULL, words);
            uint32_t* rp = row + o0;
            const uint32_t* sp = sq + o0;
            int ue = 0, dw = 0, sh4 = 0, colbase = o0 * 8;
            for (int it = 0; it < itmax; ++it) {
                if (it < words) {
                    const uint32_t t = rp[it];
                    if (t) { ue = (int)(t & 0x7ffu); const int D = (int)t >> 11; dw = D >> 3; sh4 = (D & 7) << 2; }
                    const int nrem = ue - colbase;
                    uint32_t v = 0;
                    if (nrem > 0) {
                        const uint32_t* s = sp + it + dw;
                        v = __funnelshift_l(s[1], s[0], sh4) & ~__funnelshift_rc(0xffffffffu, 0u, min(nrem, 8) * 4);
                    }
                    rp[it] = v;
                    colbase += 8;
                }
            }
        }
        // ---- expand B: head fragments (an M op starting inside a row word), OR-ed into the lane's own row
        {
            const int itmax = __reduce_max_sync(FULL, wo.nd);
            for (int it = 0; it < itmax; ++it) {
                if (it < wo.nd) {
                    const uint32_t d = cs[it];
                    const int b = (int)(d & 0x7ffu), flen = (int)((d >> 11) & 7u), D = (int)d >> 14;
                    const int o = b >> 3, kb = b & 7;
                    const int q0 = 8 * o + D;
                    const uint32_t* s = sq + (q0 >> 3);
                    const uint32_t v = __funnelshift_l(s[1], s[0], (q0 & 7) << 2);
                    const uint32_t m = (0xffffffffu >> (4 * kb)) & ~__funnelshift_rc(0xffffffffu, 0u, 4 * (kb + flen));
                    row[o] |= v & m;
                }
            }
        }
        __syncwarp();

        // ---- column sum: lane owns row words lane + 32 j, all 32 rows
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            uint32_t* col = rows + lane + 32 * j;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                if (blk * 16 < n) {
                    uint32_t xw[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) { xw[q] = col[(blk * 16 + q) * RS]; col[(blk * 16 + q) * RS] = 0; }
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
}

}  // namespace

bool tc_pileup_warp_supported(const pileup_args& a) {
    return (((uintptr_t)a.r.seq4 | (uintptr_t)a.r.cigar) & 15u) == 0;
}

template <int WC>
static int launch_geom(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    using G = geom<WC>;
    const size_t smem = sizeof(uint32_t) * (size_t)G::WARP_WORDS * G::WARPS;
    TC_CUDA(cudaFuncSetAttribute(warp_pileup_kernel<WC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    warp_pileup_kernel<WC><<<ctx->sm_count, G::WARPS * 32, smem, s>>>(a);
    TC_LAUNCH_CHECK();
    return TC_OK;
}

// Both geometries are enqueued; each reads the longest reference span the span pass left in
// a.status and returns at once unless it is the one that fits (no host round trip in between).
int tc_pileup_warp_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    int rc = launch_geom<64>(ctx, a, s);
    if (rc) return rc;
    return launch_geom<128>(ctx, a, s);
}
