// pileup_flat.cu — variant 5 of the pileup kernel: the CIGAR walk, the expansion and the column sum are three
// separate loops of a warp over a sub-tile of <= 32 start-sorted reads, each with the lane assignment that keeps
// the warp full.
//
// Variant 3 (pileup_warp.cu) walks and expands in one loop, lane = read: 193 warp-instructions per 400-base
// ONT-like read at 21 of 32 lanes active (profiles/r1_v70_pileup_warp_tma.md) — every lane pays the trip count
// of the read with the most CIGAR ops, and every chunk pays a fixed number of row words whatever the op's
// length.  Same bit-parallel arithmetic here (BAM's base codes are one-hot, a column's A/C/G/T counts are a
// positional popcount down the reads through a carry-save tree into bit-sliced counters in registers), same
// partition / window / TMA staging, but:
//
//   walk        lane = read.  The walk only computes where things are: every maximal run of M/=/X ops becomes a
//               SEGMENT descriptor of 8 bytes — shared address of its first row word, columns from that word's
//               start to the segment's end, query nibble under the word's first column, first column inside the
//               word — appended to the list of the read's row group (ballot + popcount; lists live in shared
//               memory): segments that end within 4 row words to the SHORT list (from the front of the group's
//               buffer), longer ones to the LONG list (from its back).  D / I events go to the packed global
//               X | I counters as one 64-bit add each.  No SEQ word is touched, no row is written.  The raw
//               32-bit ops are read where the bulk copy put them (the rows, idle at that point): no 16-bit
//               repacking pass, no limit on an op's length.
//   expand      lane = descriptor, 32 at a time — all lanes busy whatever the reads' op counts.  Short list:
//               5 source words, 4 funnel shifts, a head mask and 4 tail masks (one vector load from a 33-entry
//               table), 4 red.shared.or into the segment's row; no branches.  Long list: 8 words per iteration,
//               the remainder of a still longer segment goes back as a descriptor into the slots the iteration
//               just consumed (the list is its own work queue and never grows).
//   column sum  lane = NW consecutive row words (one vector load per row), RP rows per pass through 16-input
//               carry-save trees into the counters.  A 512-column window sums its 32 reads as two groups of 16
//               rows — the rows take 4 KB of the warp's slice instead of 8.
//
// Declined (TC_ERR_CAPACITY -> tc_pileup_counts falls back to the scatter kernel, which carries htslib's full
// look-ahead state): sub-tiles with pads or zero-length ops; a single read larger than the staging buffers.
// PIECES mode (pileup_long.cu): the "reads" are pieces of long reads, gathered through their start-sorted order.
#include <limits.h>

#include "pileup_smem.cuh"

namespace {

using namespace tcsm;

constexpr int MIN_SLACK = 64;
constexpr int HI_PLANES = 7;            // bit-sliced planes above "eights": 15 + 16 * 127 = 2047 reads per run
constexpr int RUN_CAP = 2047;
#ifndef TC_FLAT_CW
#define TC_FLAT_CW 4
#endif
constexpr int CW = TC_FLAT_CW;          // row words a lane emits per descriptor and iteration
#ifndef TC_FLAT_MAX_WARPS
#define TC_FLAT_MAX_WARPS 24
#endif
#ifndef TC_FLAT_SEQ_PER_WC
#define TC_FLAT_SEQ_PER_WC 26
#endif
#ifndef TC_FLAT_CIG_PER_WC
#define TC_FLAT_CIG_PER_WC 15
#endif

template <int WC> struct fgeom {
    static constexpr int ROWW = WC * 8;                     // window width in reference columns
    static constexpr int NW = WC / 32;                      // consecutive row words owned by one lane in the column sum
    static constexpr int RP = (WC == 32) ? 32 : 16;         // rows (reads) per column-sum pass
    static constexpr int NG = 32 / RP;                      // row groups per sub-tile
    static constexpr int RS = WC + NW;                      // row stride (words): a multiple of NW, odd multiple keeps rows on different banks
    static constexpr int ROW_WORDS = RP * RS;
    static constexpr int SEQ_PAD = 4;                       // zero words in front of the staged SEQ stream
    static constexpr int SEQ_CAP = WC * TC_FLAT_SEQ_PER_WC; // staged SEQ words per sub-tile (WC=64: 32 reads of 416 bases)
    static constexpr int SEQ_WORDS = SEQ_PAD + SEQ_CAP + 8;
    static constexpr int CIG_CAP = WC * TC_FLAT_CIG_PER_WC; // staged CIGAR ops per sub-tile (raw, in the rows)
    static constexpr int GROUP_OPS = CIG_CAP / NG;          // ... and per row group
    // a segment needs a match op that is the read's first op or follows an op that is none:
    // segments of a group <= (its ops + its reads) / 2
    static constexpr int DCAP = (GROUP_OPS + RP) / 2 + 4;
    static constexpr int DESC_WORDS = 2 * NG * DCAP;
    static constexpr int WARP_WORDS = ROW_WORDS + SEQ_WORDS + DESC_WORDS;
    static constexpr int LUT_WORDS = 33 * 4 + 4;            // tail masks of 4 row words for 0..32 columns (CTA-wide, behind the warps' slices)
    static constexpr int WARPS = (227 * 1024 / 4 - LUT_WORDS) / (WARP_WORDS + 2) > TC_FLAT_MAX_WARPS ? TC_FLAT_MAX_WARPS : (227 * 1024 / 4 - LUT_WORDS) / (WARP_WORDS + 2);
    static constexpr int SMEM_WORDS = (WARP_WORDS + 2) * WARPS + LUT_WORDS;
    static_assert(CIG_CAP + 4 <= ROW_WORDS, "the raw ops of a sub-tile are staged in the rows");
    static_assert(ROW_WORDS % 4 == 0 && SEQ_WORDS % 4 == 0 && DESC_WORDS % 4 == 0, "16-byte aligned regions");
    static_assert(DCAP % 2 == 0, "8-byte descriptors, 16-byte aligned lists");
};

extern __shared__ __align__(16) uint32_t smem[];

__device__ __forceinline__ uint2 lds2(uint32_t a) {
    uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v;
}

// Segment descriptor.  Word 0: shared address of the segment's first row word (18 bits) | E << 18, E = columns from
// that word's first column to the segment's end.  Word 1: query nibble under that word's first column, relative to
// the warp's SEQ region (16 bits) | the segment's first column inside the word << 16.
constexpr int D_E_SHIFT = 18, D_KB_SHIFT = 16;
constexpr uint32_t D_ADDR_MASK = (1u << D_E_SHIFT) - 1u;

// X | I event: one 64-bit add to the column's packed counter (X in the low, I in the high 32 bits), predicated
__device__ __forceinline__ void red_xi_if(bool p, unsigned long long* xi, uint32_t col, uint32_t lo, uint32_t hi) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 v, q;\n\tsetp.ne.u32 p, %0, 0;\n\tmov.b64 v, {%3, %4};\n\tmad.wide.u32 q, %2, 8, %1;\n\t@p red.global.add.u64 [q], v;\n\t}"
                 ::"r"((uint32_t)p), "l"(xi), "r"(col), "r"(lo), "r"(hi) : "memory");
}

// red.shared.or of a word that has any bit set (predicated, no branch): words behind a segment's end, and the lanes
// of an iteration that hold no descriptor, cost no shared-memory cycles — 24 idle lanes OR-ing zero into one
// address were 24 serialised atomics (profiles/r2_v2: shared-memory pipe at 72 %, 63 M conflict cycles on atomics)
__device__ __forceinline__ void reds_or_nz(uint32_t a, uint32_t v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(a), "r"(v) : "memory");
}

// NWORDS row words of a segment (NWORDS = 4: E <= 32): NWORDS + 1 source words, funnel shifts, head / tail masks,
// red.shared.or.  Words behind the segment's end get an empty mask and are not written.
template <int NWORDS>
__device__ __forceinline__ void emit_words(const uint2 d, const uint32_t seq_s, const uint32_t lut) {
    const uint32_t ro = d.x & D_ADDR_MASK, E = d.x >> D_E_SHIFT;        // (an empty descriptor: E = 0, every mask empty)
    const uint32_t src = seq_s + ((d.y >> 1) & 0x7ffcu);        // word holding nibble d.y & 0xffff
    const uint32_t sh = d.y << 2;                               // funnel shifts take it modulo 32: 4 * (nibble & 7)
    const uint32_t head = 0xffffffffu >> (4u * (d.y >> D_KB_SHIFT));
    uint32_t w[NWORDS + 1];
#pragma unroll
    for (int j = 0; j <= NWORDS; ++j) w[j] = lds(src + 4 * j);
    uint32_t t[NWORDS];
    {
        const uint4 v = lds4(lut + 16u * min(E, 32u));
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
    }
    if constexpr (NWORDS == 8) {
        const uint4 v = lds4(lut + 16u * (min(max(E, 32u), 64u) - 32u));
        t[4] = v.x; t[5] = v.y; t[6] = v.z; t[7] = v.w;
    }
    reds_or_nz(ro, __funnelshift_l(w[1], w[0], sh) & head & t[0]);
#pragma unroll
    for (int j = 1; j < NWORDS; ++j) reds_or_nz(ro + 4 * j, __funnelshift_l(w[j + 1], w[j], sh) & t[j]);
}

template <int NW> struct rowvec;
template <> struct rowvec<1> {
    uint32_t w[1];
    __device__ __forceinline__ void load_clear(uint32_t a) { w[0] = lds(a); sts(a, 0u); }
};
template <> struct rowvec<2> {
    uint32_t w[2];
    __device__ __forceinline__ void load_clear(uint32_t a) { const uint2 v = lds2(a); w[0] = v.x; w[1] = v.y; sts2(a, 0u, 0u); }
};
template <> struct rowvec<4> {
    uint32_t w[4];
    __device__ __forceinline__ void load_clear(uint32_t a) {
        const uint4 v = lds4(a); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; sts4(a, make_uint4(0u, 0u, 0u, 0u));
    }
};

template <int WC, bool PIECES>
__global__ void __launch_bounds__(fgeom<WC>::WARPS * 32, 1) flat_pileup_kernel(pileup_args a) {
    using G = fgeom<WC>;
    constexpr int ROWW = G::ROWW, RS = G::RS, NW = G::NW, RP = G::RP, NG = G::NG;
    const int lane = threadIdx.x & 31;
    const int L = a.L;

    // which geometry handles this batch: the narrowest whose slack is usable (the longest reference span: the
    // caller's bound or what the span pass found; PIECES: pieces span at most PIECE_COLS columns)
    const bool fold = !PIECES && a.span_hint > 0;   // no span pass ran: this kernel also does its checks and the coverage ends
    const int ms = PIECES ? PIECE_COLS : ((max(fold ? a.span_hint : a.status->max_span, 1) + 7) & ~7);
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
    const uint32_t rows = (uint32_t)__cvta_generic_to_shared(smem) + 4u * (threadIdx.x >> 5) * G::WARP_WORDS;   // [RP][RS]
    const uint32_t seq_s = rows + 4u * G::ROW_WORDS;            // seq_s + 4 * SEQ_PAD <-> SEQ word sbase_al
    const uint32_t desc_s = seq_s + 4u * G::SEQ_WORDS;          // [NG][DCAP] descriptors of 8 bytes
    for (int i = lane; i < G::ROW_WORDS; i += 32) sts(rows + 4 * i, 0u);
    if (lane < G::SEQ_PAD) sts(seq_s + 4 * lane, 0u);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem) + 4u * G::WARPS * G::WARP_WORDS + 8u * (threadIdx.x >> 5);
    uint32_t tma_parity = 0;
    if (lane == 0) { mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    // tail masks: entry E, word j = the leading clamp(4 * E - 32 * j, 0, 32) bits
    const uint32_t lut = (uint32_t)__cvta_generic_to_shared(smem) + 4u * ((G::WARPS * (G::WARP_WORDS + 2) + 3) & ~3);      // 16-byte aligned
    if (threadIdx.x < 33 * 4) {
        const int bits = min(max(4 * (int)(threadIdx.x >> 2) - 32 * (int)(threadIdx.x & 3), 0), 32);
        sts(lut + 4 * threadIdx.x, bits == 0 ? 0u : 0xffffffffu << (32 - bits));
    }
    __syncthreads();            // the only CTA-wide barrier: from here on a warp never waits for another

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
    int prev_pos = (fold && r > 0 && r < r_end) ? a.r.pos[r - 1] : INT_MIN;     // sort-order check across sub-tiles
    int my_max_span = 0, my_zero_span = 0;
    // lanes of my row group, and those of them below me
    const uint32_t gmask = (NG == 1) ? FULL : (((1u << RP) - 1u) << (lane & ~(RP - 1)));
    const uint32_t glt = ((1u << lane) - 1u) & gmask;
    const uint32_t my_list = desc_s + 8u * G::DCAP * (uint32_t)(lane / RP);

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
                    const int colr = w0 + 8 * (NW * lane + j) + (7 - (bit >> 2));
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
        bool passes = true;
        uint32_t ops_lane, sq_lane;     // shared addresses of the lane's first op and of its first SEQ word
        int ops_vecs;                   // 16-byte vectors of the rows the staged ops occupy
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
            const int c_grp0 = __shfl_sync(FULL, c_inc - (valid ? cv : 0), lane & ~(RP - 1));      // op vectors in front of my group
            const bool inwin = p >= w0 && p - w0 < slack;
            const bool fits = valid && inwin && 4 * s_inc <= G::SEQ_CAP && 4 * (c_inc - c_grp0) <= G::GROUP_OPS;
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
            const uint32_t cdst = rows + 16u * (uint32_t)(c_inc - cv);
            const int s_tot = __shfl_sync(FULL, s_inc, n - 1), c_tot = __shfl_sync(FULL, c_inc, n - 1);
            const int64_t sb_p = (int64_t)(so & ~3u);
            // every lane asks for its own piece's words and ops as bulk copies that complete on the warp's mbarrier —
            // the whole gather in flight at once (the piece-CIGAR buffer is padded to whole vectors behind its last op)
            const bool tma_p = __all_sync(FULL, lane >= n || sb_p + 4ll * sv <= n_seq_words);
            if (tma_p) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_expect_tx(mbar, 16u * (uint32_t)(s_tot + c_tot));
                __syncwarp();
                if (lane < n) {
                    tma_bulk_g2s(sdst, a.r.seq4 + sb_p, 16u * (uint32_t)sv, mbar);
                    if (cv > 0) tma_bulk_g2s(cdst, a.r.cigar + (co & ~3u), 16u * (uint32_t)cv, mbar);
                }
                mbar_wait(mbar, tma_parity);
                tma_parity ^= 1u;
            } else {
                if (lane < n) {         // the last words of the batch: do not read past the array
                    for (int v = 0; v < 4 * sv; ++v) sts(sdst + 4 * v, sb_p + v < n_seq_words ? __ldg(a.r.seq4 + sb_p + v) : 0u);
                    for (int v = 0; v < 4 * cv; ++v) sts(cdst + 4 * v, __ldg(a.r.cigar + (co & ~3u) + v));
                }
                __syncwarp();
            }
            // SEQ in place: codes that are not one-hot (N, IUPAC) cleared, first base to the top nibble
            {
                const uint32_t sbase = seq_s + 4u * G::SEQ_PAD;
                for (int i = lane; i < s_tot; i += 32) {
                    uint4 v = lds4(sbase + 16 * i);
                    if (multibit(v.x) | multibit(v.y) | multibit(v.z) | multibit(v.w)) {
                        v.x = clear_multibit(v.x); v.y = clear_multibit(v.y); v.z = clear_multibit(v.z); v.w = clear_multibit(v.w);
                    }
                    v.x = __byte_perm(v.x, 0, 0x0123); v.y = __byte_perm(v.y, 0, 0x0123);
                    v.z = __byte_perm(v.z, 0, 0x0123); v.w = __byte_perm(v.w, 0, 0x0123);
                    sts4(sbase + 16 * i, v);
                }
            }
            __syncwarp();
            ops_lane = cdst + 4u * (co & 3u);
            sq_lane = sdst + 4u * (so & 3u);
            ops_vecs = c_tot;
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
            // ops in front of my row group (group 0 starts at the aligned-down base)
            const uint32_t co_g0 = __shfl_sync(FULL, co, lane & ~(RP - 1));
            const uint32_t cgrp0 = (lane < RP) ? cbase_al : co_g0;
            const bool inwin = p >= w0 && p - w0 < slack;
            const bool fits = valid && inwin && (so_next - sbase_al + 1 <= (uint32_t)G::SEQ_CAP) && (co_next - cgrp0 <= (uint32_t)G::GROUP_OPS);
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
            const uint32_t send = __shfl_sync(FULL, so_next, n - 1) + 1;        // one word of look-ahead for the funnel shift
            const int nv = (int)((send - sbase_al + 3) >> 2);
            const uint32_t cend = __shfl_sync(FULL, co_next, n - 1);
            const int ncv = (int)((cend - cbase_al + 3) >> 2);
            const uint32_t sdst = seq_s + 4u * G::SEQ_PAD;
            // the two contiguous ranges are requested as bulk copies (SEQ to its place, the raw CIGAR ops into the rows, which
            // are idle and all zero here): every byte of the sub-tile is in flight at once and no register holds any of it
            const bool use_tma = (int64_t)sbase_al + 4ll * nv <= n_seq_words && (int64_t)cbase_al + 4ll * ncv <= n_ops_total;
            if (use_tma) {
                fence_proxy_async();        // the warp's earlier accesses to these buffers come before the copies' writes
                __syncwarp();
                if (lane == 0) {
                    mbar_expect_tx(mbar, 16u * (uint32_t)(nv + ncv));
                    tma_bulk_g2s(sdst, a.r.seq4 + sbase_al, 16u * (uint32_t)nv, mbar);
                    if (ncv > 0) tma_bulk_g2s(rows, a.r.cigar + cbase_al, 16u * (uint32_t)ncv, mbar);
                }
            }
            // the next sub-tile starts where this one ends and is about as large: pull its lines (and the metadata lines
            // two sub-tiles ahead) towards L2 while this one is being processed
            {
                const int64_t s_lo = ((int64_t)send - 1) & ~3ll, s_len = (int64_t)send - sbase_al;
                const int64_t c_lo = (int64_t)cend & ~3ll, c_len = (int64_t)cend - cbase_al;
                if (lane == 0) {
                    const int64_t sw = min(s_len + 4, n_seq_words - s_lo) & ~3ll, cw = min(c_len + 4, n_ops_total - c_lo) & ~3ll;
                    if (sw > 0) prefetch_l2_bulk(a.r.seq4 + s_lo, (uint32_t)(4 * sw));
                    if (cw > 0) prefetch_l2_bulk(a.r.cigar + c_lo, (uint32_t)(4 * cw));
                }
                const int64_t rm = r + n + 64;
                if (rm < n_reads) {
                    if (lane == 0) prefetch_l2(a.r.pos + rm);
                    if (lane == 1) prefetch_l2(a.r.seq_off + rm);
                    if (lane == 2) prefetch_l2(a.r.cigar_off + rm);
                    if (lane == 3) prefetch_l2(a.r.l_seq + rm);
                    if (lane == 4) prefetch_l2(a.r.flag + rm);
                }
            }
            if (use_tma) {
                mbar_wait(mbar, tma_parity);
                tma_parity ^= 1u;
            } else {        // the last sub-tile of the batch: do not read past the arrays
                for (int i = lane; i < 4 * nv; i += 32) {
                    const int64_t wi = (int64_t)sbase_al + i;
                    sts(sdst + 4 * i, wi < n_seq_words ? __ldg(a.r.seq4 + wi) : 0u);
                }
                for (int i = lane; i < 4 * ncv; i += 32) {
                    const int64_t oi = (int64_t)cbase_al + i;
                    sts(rows + 4 * i, oi < n_ops_total ? __ldg(a.r.cigar + oi) : 0u);
                }
                __syncwarp();
            }
            // SEQ in place: codes that are not one-hot (N, IUPAC; rare) cleared, first base to the top nibble
            for (int i = lane; i < nv; i += 32) {
                uint4 v = lds4(sdst + 16 * i);
                if (multibit(v.x) | multibit(v.y) | multibit(v.z) | multibit(v.w)) {
                    v.x = clear_multibit(v.x); v.y = clear_multibit(v.y); v.z = clear_multibit(v.z); v.w = clear_multibit(v.w);
                }
                v.x = __byte_perm(v.x, 0, 0x0123); v.y = __byte_perm(v.y, 0, 0x0123);
                v.z = __byte_perm(v.z, 0, 0x0123); v.w = __byte_perm(v.w, 0, 0x0123);
                sts4(sdst + 16 * i, v);
            }
            __syncwarp();
            ops_lane = rows + 4u * (co - cbase_al);
            nops_lane = (int)(co_next - co);
            sq_lane = sdst + 4u * (so - sbase_al);
            ops_vecs = ncv;
        }

        // ---- walk: lane = read.  Segments -> descriptors in the row group's list; D / I events -> global counters.
        // One iteration takes a match op (extending the lane's run of match ops) and then the op behind it that is
        // none (which ends the run: its descriptor goes out with its final length) — the common M I M D M shape
        // costs one iteration per segment; anything else simply takes another iteration.
        const bool act = lane < n && passes;
        const int x0 = p - w0;
        int x = x0, y = y0;
        int cnt = 0;                    // descriptors in my group's list (the same in all lanes of the group)
        {
            uint32_t cp = act ? ops_lane : rows;
            const uint32_t cend = cp + 4u * (uint32_t)(act ? nops_lane : 0);
            const uint32_t qn = 2u * (sq_lane - seq_s);            // nibble address of the read's first base
            const uint32_t rowbase = rows + 4u * (uint32_t)((lane & (RP - 1)) * RS);
            const bool hasseq = lq > 0;         // reads without SEQ ('*': every base reads 'N') count coverage and events only
            uint32_t c0 = lds(cp), c1 = lds(cp + 4);
            uint32_t pr = 0, pd = 0;    // the op before consumed the reference / was a deletion
            int xo = x, yo = y, lo = 0; // the current run of match ops: first column, first query base, length (0: none)
            int ym = 0;                 // query index behind the last match op
            uint32_t exmin = 0xffffffffu;       // min over the ops of (len << 4 | op), pads counted as 0: < 16 <=> a pad or a zero-length op
#pragma unroll 1
            while (__any_sync(FULL, cp < cend)) {
                const uint32_t f0 = op_flags(c0 & 15u), f1 = op_flags(c1 & 15u);
                const bool is_m = cp < cend && (f0 & 1u);
                const int lm = is_m ? (int)(c0 >> 4) : 0;
                lo += lm; x += lm; y += lm;
                ym = is_m ? y : ym;
                exmin = is_m ? min(exmin, c0) : exmin;
                pr = is_m ? 1u : pr; pd = is_m ? 0u : pd;
                cp += is_m ? 4u : 0u;
                const uint32_t c = is_m ? c1 : c0, fl = is_m ? f1 : f0;
                const bool more_m = cp < cend && (fl & 1u);     // another match op follows: the run goes on
                const bool nm = cp < cend && !(fl & 1u);
                const bool emit = lo > 0 && !more_m && hasseq;
                const unsigned eb = __ballot_sync(FULL, emit);
                if (emit) {
                    const int kb = xo & 7;
                    sts2(my_list + 8u * (uint32_t)(cnt + __popc(eb & glt)),
                         (rowbase + (uint32_t)((xo >> 3) << 2)) | ((uint32_t)(kb + lo) << D_E_SHIFT),
                         (qn + (uint32_t)(yo - kb)) | ((uint32_t)kb << D_KB_SHIFT));
                }
                cnt += __popc(eb & gmask);
                lo = more_m ? lo : 0;
                if (nm) {
                    const uint32_t op = c & 15u;
                    const int l = (int)(c >> 4);
                    exmin = min(exmin, op == OP_P ? 0u : c);
                    // X / I events: a deletion's columns count +1 X each; an insertion counts +1 I on its anchor column x - 1 —
                    // and -1 X there when that column belongs to a deletion (it reads "*+n..", not "*").  Without zero-length
                    // ops the anchor exists whenever the previous op consumed the reference; a column past the reference is
                    // clamped (such a read is a TC_ERR_RANGE, the counts are void).  One 64-bit add per event: X is the low,
                    // I the high half of the column's counter, and 2^32 - 1 is "+1 I, -1 X" (the sums are exact modulo 2^64).
                    const bool is_d = (op == OP_D);
                    const bool ev = is_d || (op == OP_I && pr);
                    const uint32_t col = (uint32_t)min(w0 + x - (is_d ? 0 : 1), L - 1);
                    red_xi_if(ev, a.xi, col, is_d ? 1u : 0u - pd, (is_d ? 1u : pd) ^ 1u);
                    if (is_d && l > 1) for (int k = w0 + x + 1; k <= min(w0 + x + l - 1, L - 1); ++k) red_xi_if(true, a.xi, (uint32_t)k, 1u, 0u);
                    pr = (fl >> 1) & 1u; pd = is_d ? 1u : 0u;
                    x += (fl & 2u) ? l : 0;
                    y += (fl & 4u) ? l : 0;
                    xo = x; yo = y;             // a run can only start behind an op that is no match
                    cp += 4;
                }
                c0 = lds(cp); c1 = lds(cp + 4);
            }
            // beyond the window, a CIGAR that consumes more query than SEQ holds, pads, zero-length ops: the scatter kernel's business
            const bool over = act && (x > ROWW || (hasseq && ym > lq));
            if (__any_sync(FULL, over || exmin < 16u)) {
                if (lane == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
                cnt = 0;
            }
        }
        const int x_end = x;
        __syncwarp();
        // the rows held the raw ops: clear them again
        for (int i = lane; i < ops_vecs; i += 32) sts4(rows + 16 * i, make_uint4(0u, 0u, 0u, 0u));
        __syncwarp();

        // ---- per row group: expand its descriptors into its rows (lane = descriptor), then sum the rows (lane = row words)
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int ns = __shfl_sync(FULL, cnt, g * RP);
            if (ns == 0) continue;
            const uint32_t list = desc_s + 8u * G::DCAP * (uint32_t)g;
            // every segment: its first 4 row words, no branches (lanes without a descriptor run on an empty one).  What is
            // left of a longer segment goes to the LONG list — which grows, in place, over the slots already consumed
            int nl = 0;
#pragma unroll 1
            for (int h = lane; h < ns + lane; h += 32) {
                uint2 d = make_uint2(rows, 0u);
                if (h < ns) d = lds2(list + 8u * (uint32_t)h);
                emit_words<4>(d, seq_s, lut);
                const bool more = (d.x >> D_E_SHIFT) > 32u;
                const unsigned mb = __ballot_sync(FULL, more);
                if (more) sts2(list + 8u * (uint32_t)(nl + __popc(mb & ((1u << lane) - 1u))), d.x + 16u - (32u << D_E_SHIFT), (d.y & 0xffffu) + 32u);
                nl += __popc(mb);
            }
            __syncwarp();
            // long segments: 8 more row words per iteration; what is left of a still longer one goes back into the slots the
            // iteration consumed (the list is its own work queue and never grows)
#pragma unroll 1
            for (int h = 0; h < nl;) {
                const int c32 = min(32, nl - h);
                uint2 d = make_uint2(rows, 0u);
                if (lane < c32) d = lds2(list + 8u * (uint32_t)(h + lane));
                emit_words<8>(d, seq_s, lut);
                const bool more = (d.x >> D_E_SHIFT) > 64u;
                const unsigned mb = __ballot_sync(FULL, more);
                const int k = __popc(mb);
                if (more) sts2(list + 8u * (uint32_t)(h + c32 - k + __popc(mb & ((1u << lane) - 1u))), d.x + 32u - (64u << D_E_SHIFT), (d.y & 0xffffu) + 64u);
                h += c32 - k;
                __syncwarp();
            }
            // column sum: lane owns row words NW * lane .. NW * lane + NW - 1, all RP rows of the group
            const int ng = min(n - g * RP, RP);
#pragma unroll
            for (int blk = 0; blk < RP / 16; ++blk) {
                if (blk * 16 < ng) {
                    rowvec<NW> xw[16];
                    const uint32_t col = rows + 4u * (uint32_t)(NW * lane) + 4u * (uint32_t)(blk * 16 * RS);
#pragma unroll
                    for (int q = 0; q < 16; ++q) xw[q].load_clear(col + 4u * (uint32_t)(q * RS));
#pragma unroll
                    for (int j = 0; j < NW; ++j) {
                        uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteens;
                        csa(twosA, ones[j], ones[j], xw[0].w[j], xw[1].w[j]);
                        csa(twosB, ones[j], ones[j], xw[2].w[j], xw[3].w[j]);
                        csa(foursA, twos[j], twos[j], twosA, twosB);
                        csa(twosA, ones[j], ones[j], xw[4].w[j], xw[5].w[j]);
                        csa(twosB, ones[j], ones[j], xw[6].w[j], xw[7].w[j]);
                        csa(foursB, twos[j], twos[j], twosA, twosB);
                        csa(eightsA, fours[j], fours[j], foursA, foursB);
                        csa(twosA, ones[j], ones[j], xw[8].w[j], xw[9].w[j]);
                        csa(twosB, ones[j], ones[j], xw[10].w[j], xw[11].w[j]);
                        csa(foursA, twos[j], twos[j], twosA, twosB);
                        csa(twosA, ones[j], ones[j], xw[12].w[j], xw[13].w[j]);
                        csa(twosB, ones[j], ones[j], xw[14].w[j], xw[15].w[j]);
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
        }

        // ---- without a span pass: sort order, range, span statistics and the two ends of every read's span in the
        // coverage difference array (adds to the same column are combined inside the warp first)
        if (fold) {
            int pprev = __shfl_up_sync(FULL, p, 1);
            if (lane == 0) pprev = prev_pos;
            if (lane < n && p < pprev) atomicCAS(&a.status->err, 0, TC_ERR_UNSORTED);
            prev_pos = __shfl_sync(FULL, p, n - 1);
            const int span = act ? x_end - x0 : 0;
            const bool badr = act && (p >= L || p + span > L);
            if (badr) atomicCAS(&a.status->err, 0, TC_ERR_RANGE);
            const bool has = act && span > 0 && !badr;
            my_max_span = max(my_max_span, span);
            my_zero_span += (act && span == 0) ? 1 : 0;
            const unsigned g0 = __match_any_sync(FULL, has ? p : -1 - lane);
            if (has && (__ffs(g0) - 1) == lane) atomicAdd(&a.diff[p], __popc(g0));
            const unsigned g1 = __match_any_sync(FULL, has ? p + span : -1 - lane);
            if (has && (__ffs(g1) - 1) == lane) atomicAdd(&a.diff[p + span], -__popc(g1));
        }

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

template <int WC, bool PIECES>
int launch_geom(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    using G = fgeom<WC>;
    const size_t smem = sizeof(uint32_t) * (size_t)G::SMEM_WORDS;
    const uint32_t bit = 1u << (8 + (WC == 32 ? 0 : WC == 64 ? 1 : 2) + (PIECES ? 3 : 0));
    if (!(ctx->warp_attr_set & bit)) {
        TC_CUDA(cudaFuncSetAttribute(flat_pileup_kernel<WC, PIECES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ctx->warp_attr_set |= bit;
    }
    flat_pileup_kernel<WC, PIECES><<<ctx->sm_count, G::WARPS * 32, smem, s>>>(a);
    TC_LAUNCH_CHECK();
    return TC_OK;
}

}  // namespace

// Without a span bound all geometries are enqueued: each reads the longest reference span the span pass left in
// a.status and returns at once unless it is the one that fits (no host round trip in between).  With the caller's
// bound the host knows which one that is.
int tc_pileup_flat_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    if (a.span_hint > 0) {
        const int ms = (a.span_hint + 7) & ~7;
        const int slack32 = (256 - ms - 8) & ~7, slack64 = (512 - ms - 8) & ~7;
        if (slack32 >= MIN_SLACK) return launch_geom<32, false>(ctx, a, s);
        if (slack64 >= MIN_SLACK) return launch_geom<64, false>(ctx, a, s);
        return launch_geom<128, false>(ctx, a, s);
    }
    int rc = launch_geom<32, false>(ctx, a, s);
    if (rc) return rc;
    rc = launch_geom<64, false>(ctx, a, s);
    if (rc) return rc;
    return launch_geom<128, false>(ctx, a, s);
}

// pieces of long reads (pileup_long.cu): always the 512-column geometry
int tc_pileup_flat_launch_pieces(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) { return launch_geom<64, true>(ctx, a, s); }
