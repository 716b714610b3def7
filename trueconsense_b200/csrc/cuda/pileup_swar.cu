// pileup_swar.cu — variant 2 of the pileup kernel: no atomic per base.
//
// Why: the scatter kernel (pileup.cu) pays one shared-memory atomic per aligned base, and
// spread-address shared atomics run at ~0.5 lane/clk/SM on this architecture
// (B300_MICROARCH.md "Atomics") — two orders of magnitude under what HBM can feed
// (profiles/r1_v1_scatter.md: 0.9 % of the HBM roofline).  Here the per-base work is bit-parallel:
//
//   work unit   = the reads whose START lies in one window of SLACK reference columns, cut into
//                 pieces of <= UNIT_READS reads; every read belongs to exactly one unit, so it is
//                 staged and walked once.  All columns such a read can touch lie in
//                 [w0, w0 + ROWW) because SLACK + (longest reference span) <= ROWW.  Units are
//                 handed to persistent CTAs through an atomic counter.
//   stage       the packed SEQ words and CIGAR ops of a sub-tile of T consecutive reads are
//                 contiguous in HBM, so they stream into shared memory as 16-byte vectors.  On the
//                 way every SEQ word is byte-swapped (first base in the top nibble) and every code
//                 that is not exactly A/C/G/T (N, IUPAC, '=') is cleared: such bases only count
//                 towards coverage, which comes from the difference array.
//   walk        one lane per read steps through its CIGAR in a warp-uniform loop: deletion and
//                 insertion events go to small shared counters (they are sparse), every M/=/X op
//                 becomes a 32-bit segment descriptor written over the consumed ops in place.
//   expand      one lane per read, warp-uniform loop, one row word (8 columns) per iteration:
//                 funnel-shift the staged words onto the reference grid and write a
//                 reference-aligned row of one-hot 4-bit codes (row stride padded to ROWW/8+1
//                 words: the 32 lanes of a warp hit 32 different banks).
//   column sum  BAM's base codes are one-hot (A=1, C=2, G=4, T=8), so counting A/C/G/T per column is
//                 a positional popcount down the rows: thread c owns row-word column c (8 columns
//                 x 4 classes = 32 bit positions) and adds 16 rows at a time with a Harley-Seal
//                 carry-save tree (15 full adders = 30 LOP3) into bit-sliced counters that stay in
//                 registers for the whole unit; they become integers and go to HBM once per unit.
#include <limits.h>

#include "pileup.cuh"

namespace {

constexpr int ROWW = 896;           // row width in reference columns
constexpr int WC = ROWW / 8;        // row words (112)
constexpr int RS = WC + 1;          // padded row stride (words)
constexpr int T = 128;              // reads per sub-tile == threads per CTA
constexpr int SEQ_CAP = 6912;       // staged SEQ words per sub-tile (incl. pads)
constexpr int CIG_CAP = 4096;       // staged CIGAR ops per sub-tile
constexpr int SEQ_PAD = 4;          // zero words in front of the staged SEQ stream
constexpr int UNIT_READS = 2048;
constexpr int HI_PLANES = 8;        // bit-sliced planes above "eights": counts up to 15 + 16*255 = 4095 >= UNIT_READS
constexpr int SMEM_WORDS = T * RS + SEQ_CAP + 8 + CIG_CAP + 8 + 2 * ROWW;
constexpr int MIN_SLACK = 64;

struct swar_work {
    int slack;          // start-window width (multiple of 8)
    int n_windows;
    int total_units;
    int counter;
};

__global__ void swar_units_kernel(pileup_args a, swar_work* work, int* __restrict__ win_lo, int* __restrict__ unit_prefix,
                                  int max_windows) {
    __shared__ int s_slack, s_nw;
    if (threadIdx.x == 0) {
        const int ms = (max(a.status->max_span, 1) + 7) & ~7;
        int slack = (ROWW - ms - 8) & ~7;
        int nw = 0;
        if (slack < MIN_SLACK) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); slack = MIN_SLACK; }
        else nw = (a.L + slack - 1) / slack;
        if (nw > max_windows) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); nw = 0; }
        s_slack = slack; s_nw = nw;
    }
    __syncthreads();
    const int slack = s_slack, nw = s_nw;
    auto lower = [&](long long v) {
        int64_t lo = 0, hi = a.r.n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((long long)a.r.pos[mid] < v) lo = mid + 1; else hi = mid; }
        return (int)lo;
    };
    for (int w = threadIdx.x; w <= nw; w += blockDim.x) win_lo[w] = (w == nw) ? (int)a.r.n : lower((long long)w * slack);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        if (nw > 0) win_lo[0] = 0;      // reads with negative positions are flagged by the span pass; keep them in unit 0
        for (int w = 0; w < nw; ++w) {
            unit_prefix[w] = acc;
            acc += (win_lo[w + 1] - win_lo[w] + UNIT_READS - 1) / UNIT_READS;
        }
        unit_prefix[nw] = acc;
        work->slack = slack; work->n_windows = nw; work->total_units = acc; work->counter = 0;
    }
}

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    l = a ^ b ^ c;                      // one LOP3
    h = (a & b) | (c & (a | b));        // one LOP3 (majority)
}

// byte-swap to "first base in the top nibble" and clear every nibble that is not one-hot
__device__ __forceinline__ uint32_t prep_seq_word(uint32_t w) {
    w = __byte_perm(w, 0, 0x0123);
    const uint32_t z = w & ((w | 0x88888888u) - 0x11111111u);      // non-zero nibble <=> two or more bits set
    const uint32_t m = (z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x11111111u;
    return w & ~(m * 15u);
}

__global__ void __launch_bounds__(T, 2) swar_main_kernel(pileup_args a, swar_work* work, const int* __restrict__ win_lo,
                                                         const int* __restrict__ unit_prefix) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* rows = smem;                              // [T][RS]
    uint32_t* seq_s = rows + T * RS;                    // [SEQ_CAP + 8]; index SEQ_PAD <-> word sbase_al
    uint32_t* cig_s = seq_s + SEQ_CAP + 8;              // [CIG_CAP + 8]
    int* xcnt = (int*)(cig_s + CIG_CAP + 8);            // [ROWW]
    int* icnt = xcnt + ROWW;                            // [ROWW]
    __shared__ int s_unit;
    const int tid = threadIdx.x;
    const int L = a.L;
    const int64_t n_seq_words = (int64_t)a.r.seq_off[a.r.n];
    const int64_t n_ops_total = (int64_t)a.r.cigar_off[a.r.n];
    const int slack = work->slack;
    const int n_windows = work->n_windows;
    const int total_units = work->total_units;

    for (int i = tid; i < T * RS; i += T) rows[i] = 0;
    for (int i = tid; i < 2 * ROWW; i += T) xcnt[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(&work->counter, 1);
        __syncthreads();
        const int u = s_unit;
        __syncthreads();
        if (u >= total_units) break;
        int wlo = 0, whi = n_windows;
        while (whi - wlo > 1) { int mid = (wlo + whi) >> 1; if (unit_prefix[mid] <= u) wlo = mid; else whi = mid; }
        const int w = wlo;
        const int w0 = w * slack;
        const int u0 = win_lo[w] + (u - unit_prefix[w]) * UNIT_READS;
        const int u1 = min(win_lo[w + 1], u0 + UNIT_READS);

        uint32_t ones = 0, twos = 0, fours = 0, eights = 0;
        uint32_t hi[HI_PLANES];
#pragma unroll
        for (int j = 0; j < HI_PLANES; ++j) hi[j] = 0;

        int t0 = u0;
        while (t0 < u1) {
            // ---- how many reads fit the staging buffers
            const int nmax = min(T, u1 - t0);
            const uint32_t sbase_al = a.r.seq_off[t0] & ~3u;
            const uint32_t cbase_al = a.r.cigar_off[t0] & ~3u;
            bool fits = false;
            if (tid < nmax) {
                fits = (a.r.seq_off[t0 + tid + 1] - sbase_al + SEQ_PAD + 1 <= (uint32_t)SEQ_CAP) &&
                       (a.r.cigar_off[t0 + tid + 1] - cbase_al <= (uint32_t)CIG_CAP);
            }
            const int n = __syncthreads_count(fits);
            if (n == 0) {       // a single read larger than the staging buffers: not supported by this variant
                if (tid == 0) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
                t0 += 1;
                continue;
            }
            // ---- stage SEQ words [sbase_al, send) and CIGAR ops [cbase_al, cend) as 16-byte vectors
            {
                const uint32_t send = a.r.seq_off[t0 + n] + 1;           // one word of look-ahead for the funnel shift
                const int nv = (int)((send - sbase_al + 3) >> 2);
                const uint4* src = reinterpret_cast<const uint4*>(a.r.seq4 + sbase_al);
                uint4* dst = reinterpret_cast<uint4*>(seq_s + SEQ_PAD);
                const uint32_t cend = a.r.cigar_off[t0 + n];
                const int ncv = (int)((cend - cbase_al + 3) >> 2);
                const uint4* csrc = reinterpret_cast<const uint4*>(a.r.cigar + cbase_al);
                uint4* cdst = reinterpret_cast<uint4*>(cig_s);
                const bool seq_tail = (int64_t)sbase_al + 4ll * nv > n_seq_words;
                const bool cig_tail = (int64_t)cbase_al + 4ll * ncv > n_ops_total;
                if (!seq_tail && !cig_tail) {
                    // common case: whole vectors, four loads in flight per thread
                    for (int i = tid; i < nv; i += 4 * T) {
                        uint4 v0 = __ldg(src + i), v1, v2, v3;
                        const bool b1 = i + T < nv, b2 = i + 2 * T < nv, b3 = i + 3 * T < nv;
                        if (b1) v1 = __ldg(src + i + T);
                        if (b2) v2 = __ldg(src + i + 2 * T);
                        if (b3) v3 = __ldg(src + i + 3 * T);
                        v0.x = prep_seq_word(v0.x); v0.y = prep_seq_word(v0.y); v0.z = prep_seq_word(v0.z); v0.w = prep_seq_word(v0.w);
                        dst[i] = v0;
                        if (b1) { v1.x = prep_seq_word(v1.x); v1.y = prep_seq_word(v1.y); v1.z = prep_seq_word(v1.z); v1.w = prep_seq_word(v1.w); dst[i + T] = v1; }
                        if (b2) { v2.x = prep_seq_word(v2.x); v2.y = prep_seq_word(v2.y); v2.z = prep_seq_word(v2.z); v2.w = prep_seq_word(v2.w); dst[i + 2 * T] = v2; }
                        if (b3) { v3.x = prep_seq_word(v3.x); v3.y = prep_seq_word(v3.y); v3.z = prep_seq_word(v3.z); v3.w = prep_seq_word(v3.w); dst[i + 3 * T] = v3; }
                    }
                    for (int i = tid; i < ncv; i += 2 * T) {
                        uint4 v0 = __ldg(csrc + i), v1;
                        const bool b1 = i + T < ncv;
                        if (b1) v1 = __ldg(csrc + i + T);
                        cdst[i] = v0;
                        if (b1) cdst[i + T] = v1;
                    }
                } else {
                    // the last sub-tile of the batch: do not read past the arrays
                    for (int i = tid; i < nv; i += T) {
                        const int64_t wb = (int64_t)sbase_al + 4 * i;
                        uint4 v;
                        v.x = wb + 0 < n_seq_words ? prep_seq_word(__ldg(a.r.seq4 + wb + 0)) : 0u;
                        v.y = wb + 1 < n_seq_words ? prep_seq_word(__ldg(a.r.seq4 + wb + 1)) : 0u;
                        v.z = wb + 2 < n_seq_words ? prep_seq_word(__ldg(a.r.seq4 + wb + 2)) : 0u;
                        v.w = wb + 3 < n_seq_words ? prep_seq_word(__ldg(a.r.seq4 + wb + 3)) : 0u;
                        dst[i] = v;
                    }
                    for (int i = tid; i < ncv; i += T) {
                        const int64_t ob = (int64_t)cbase_al + 4 * i;
                        uint4 v;
                        v.x = ob + 0 < n_ops_total ? __ldg(a.r.cigar + ob + 0) : 0u;
                        v.y = ob + 1 < n_ops_total ? __ldg(a.r.cigar + ob + 1) : 0u;
                        v.z = ob + 2 < n_ops_total ? __ldg(a.r.cigar + ob + 2) : 0u;
                        v.w = ob + 3 < n_ops_total ? __ldg(a.r.cigar + ob + 3) : 0u;
                        cdst[i] = v;
                    }
                }
                if (tid < SEQ_PAD) seq_s[tid] = 0;
            }
            __syncthreads();

            // ---- walk: one lane per read, warp-uniform loop over CIGAR ops, forward-only state machine.
            // A read's bases land on the reference grid with a piecewise-constant shift D (source nibble =
            // row column + D); the walk records one 32-bit descriptor per shift change
            //   [0,10) b: first column of the new regime  [10,19) gap: zero columns starting at b (deletion /
            //   ref-skip)  [19,32) D (signed) for the columns from b+gap on
            // over the consumed ops in place, and counts the sparse events in shared counters:
            //   X  every column of a D op;   I  the last column of a reference-consuming op followed by an
            //   insertion (directly, or through pads) — and when that op was a D, its last column reads
            //   "*+n..", not "*", so the X is taken back (TrueConsense/indexing.py:118,130).
            const int64_t r = t0 + tid;
            const bool act = tid < n && read_passes(a, r);
            uint32_t* cs = cig_s;
            int nops = 0, x = 0;
            if (act) {
                cs = cig_s + (a.r.cigar_off[r] - cbase_al);
                nops = (int)(a.r.cigar_off[r + 1] - a.r.cigar_off[r]);
                x = a.r.pos[r] - w0;
                if (x < 0) nops = 0;        // negative position: flagged by the span pass
            }
            const int lq = act ? a.r.l_seq[r] : 0;
            const bool no_seq = lq == 0;    // SEQ '*': every base reads 'N' — events and coverage only
            int nd = 0, y = 0, seg_end = 0, d_cur = INT_MIN, pend = 0, lastcol = 0, b_first = 0;
            bool last_was_d = false;
            const int kmax = __reduce_max_sync(0xffffffffu, nops);
            for (int k = 0; k < kmax; ++k) {
                if (k < nops) {
                    const uint32_t c = cs[k];
                    const uint32_t op = c & 15u;
                    const int l = (int)(c >> 4);
                    if (op_is_match(op)) {
                        const int dn = y - x;
                        if (d_cur == INT_MIN) { seg_end = x; b_first = x; }
                        if (!no_seq && (dn != d_cur || x != seg_end)) {
                            const int gap = x - seg_end;
                            if (gap > 511 || y >= 4096) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
                            cs[nd++] = (uint32_t)seg_end | ((uint32_t)(gap & 511) << 10) | ((uint32_t)dn << 19);
                            d_cur = dn;
                        }
                        x += l; y += l; seg_end = x;
                        pend = 1; last_was_d = false; lastcol = x - 1;
                    } else if (op == OP_D) {
                        if (x + l <= ROWW)
                            for (int col = x; col < x + l; ++col) atomicAdd(&xcnt[col], 1);
                        x += l;
                        pend = 1; last_was_d = true; lastcol = x - 1;
                    } else if (op == OP_N) {
                        x += l;
                        pend = 1; last_was_d = false; lastcol = x - 1;
                    } else if (op == OP_I) {
                        y += l;
                        if (pend && l > 0) {
                            if (lastcol < ROWW) { atomicAdd(&icnt[lastcol], 1); if (last_was_d) atomicAdd(&xcnt[lastcol], -1); }
                            pend = 0;
                        }
                    } else if (op == OP_P) {
                        if (pend == 1) pend = 2;
                    } else {            // S, H: end a direct look-ahead, not one that went through a pad
                        if (op == OP_S) y += l;
                        if (pend == 1) pend = 0;
                    }
                    if (x > ROWW) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); nops = 0; nd = 0; }
                }
            }
            if (y > lq && nd > 0) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); nd = 0; }   // CIGAR longer than SEQ: scatter kernel
            __syncwarp();

            // ---- expand, pass A: one row word (8 columns) per iteration, warp-uniform trip count.  Each word is
            // written under the shift in force at its first column; shift changes inside a word are ignored here.
            const int o0 = b_first >> 3;
            const uint32_t* sq = seq_s + SEQ_PAD;
            if (act) sq += (a.r.seq_off[r] - sbase_al);
            uint32_t* row = rows + tid * RS + o0;
            {
                const int words = nd > 0 ? ((seg_end - 1) >> 3) - o0 + 1 : 0;
                const int itmax = __reduce_max_sync(0xffffffffu, words);
                int colbase = o0 * 8, zu = 0, dsh = 0, di = 0;
                int nb = INT_MAX, ngap = 0, nD = 0;
                bool started = false;
                if (nd > 0) { const uint32_t d = cs[0]; nb = (int)(d & 1023u); ngap = (int)((d >> 10) & 511u); nD = (int)d >> 19; }
                for (int it = 0; it < itmax; ++it) {
                    if (it < words) {
                        while (nb <= colbase) {                 // shift changes at or before this word's first column
                            dsh = nD; zu = nb + ngap; started = true;
                            if (++di < nd) { const uint32_t d = cs[di]; nb = (int)(d & 1023u); ngap = (int)((d >> 10) & 511u); nD = (int)d >> 19; }
                            else nb = INT_MAX;
                        }
                        uint32_t v = 0;
                        if (started) {
                            const int idx = (dsh >> 3) + o0 + it;
                            v = __funnelshift_l(sq[idx + 1], sq[idx], (dsh & 7) * 4);
                            if (zu > colbase) { const int kz = zu - colbase; v = kz >= 8 ? 0u : (v & (0xffffffffu >> (4 * kz))); }
                        }
                        row[it] = v;
                        colbase += 8;
                    }
                }
            }
            // ---- expand, pass B: one shift change per iteration.  A change at column b (not on a word boundary)
            // replaces the nibbles from b on in its row word: `gap` zero columns, then the bases under the new shift.
            {
                const int npatch = nd > 0 ? nd + 1 : 0;         // + the end of the read
                const int itmax = __reduce_max_sync(0xffffffffu, npatch);
                for (int it = 0; it < itmax; ++it) {
                    if (it < npatch) {
                        int b, gap, dn;
                        if (it < nd) { const uint32_t d = cs[it]; b = (int)(d & 1023u); gap = (int)((d >> 10) & 511u); dn = (int)d >> 19; }
                        else { b = seg_end; gap = 511; dn = 0; }
                        const int kb = b & 7;
                        if (kb) {
                            const int o = b >> 3;
                            const int g = kb + gap;
                            const int idx = (dn >> 3) + o;
                            const uint32_t v2 = __funnelshift_l(sq[idx + 1], sq[idx], (dn & 7) * 4);
                            const uint32_t m_new = g >= 8 ? 0u : (0xffffffffu >> (4 * g));
                            uint32_t* wp = rows + tid * RS + o;
                            *wp = (*wp & ~(0xffffffffu >> (4 * kb))) | (v2 & m_new);
                        }
                    }
                }
            }
            __syncthreads();

            // ---- column sum: thread c owns row-word column c, all T rows
            if (tid < WC) {
                uint32_t* col = rows + tid;
#pragma unroll 2
                for (int blk = 0; blk < T / 16; ++blk) {
                    uint32_t xw[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { xw[j] = col[(blk * 16 + j) * RS]; col[(blk * 16 + j) * RS] = 0; }
                    uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteens;
                    csa(twosA, ones, ones, xw[0], xw[1]);
                    csa(twosB, ones, ones, xw[2], xw[3]);
                    csa(foursA, twos, twos, twosA, twosB);
                    csa(twosA, ones, ones, xw[4], xw[5]);
                    csa(twosB, ones, ones, xw[6], xw[7]);
                    csa(foursB, twos, twos, twosA, twosB);
                    csa(eightsA, fours, fours, foursA, foursB);
                    csa(twosA, ones, ones, xw[8], xw[9]);
                    csa(twosB, ones, ones, xw[10], xw[11]);
                    csa(foursA, twos, twos, twosA, twosB);
                    csa(twosA, ones, ones, xw[12], xw[13]);
                    csa(twosB, ones, ones, xw[14], xw[15]);
                    csa(foursB, twos, twos, twosA, twosB);
                    csa(eightsB, fours, fours, foursA, foursB);
                    csa(sixteens, eights, eights, eightsA, eightsB);
                    uint32_t carry = sixteens;
#pragma unroll
                    for (int j = 0; j < HI_PLANES; ++j) { uint32_t t = hi[j] & carry; hi[j] ^= carry; carry = t; }
                }
            }
            __syncthreads();
            t0 += n;
        }
        // ---- unit epilogue: bit-sliced counters -> integers -> HBM
        if (tid < WC) {
            uint32_t planes[4 + HI_PLANES];
            planes[0] = ones; planes[1] = twos; planes[2] = fours; planes[3] = eights;
#pragma unroll
            for (int j = 0; j < HI_PLANES; ++j) planes[4 + j] = hi[j];
            uint32_t any = 0;
#pragma unroll
            for (int j = 0; j < 4 + HI_PLANES; ++j) any |= planes[j];
            if (any) {
#pragma unroll 4
                for (int bit = 0; bit < 32; ++bit) {
                    if (!((any >> bit) & 1u)) continue;
                    int v = 0;
#pragma unroll
                    for (int j = 0; j < 4 + HI_PLANES; ++j) v |= (int)((planes[j] >> bit) & 1u) << j;
                    const int colr = w0 + 8 * tid + (7 - (bit >> 2));
                    const int cls = bit & 3;     // bit 0 A, 1 C, 2 G, 3 T (BAM codes 1,2,4,8)
                    const int row = cls == 0 ? TC_ROW_A : cls == 1 ? TC_ROW_C : cls == 2 ? TC_ROW_G : TC_ROW_T;
                    if (colr < L) atomicAdd(&a.counts[(size_t)row * L + colr], v);
                }
            }
        }
        for (int i = tid; i < ROWW; i += T) {
            const int xv = xcnt[i], iv = icnt[i];
            if (xv) { if (w0 + i < L) atomicAdd(&a.counts[(size_t)TC_ROW_X * L + w0 + i], xv); xcnt[i] = 0; }
            if (iv) { if (w0 + i < L) atomicAdd(&a.counts[(size_t)TC_ROW_I * L + w0 + i], iv); icnt[i] = 0; }
        }
        __syncthreads();
    }
}

}  // namespace

bool tc_pileup_swar_supported(const pileup_args& a) {
    return (((uintptr_t)a.r.seq4 | (uintptr_t)a.r.cigar) & 15u) == 0;
}

int tc_pileup_swar_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    const int max_windows = (a.L + MIN_SLACK - 1) / MIN_SLACK + 1;
    int* buf = (int*)tc_dev_buf(ctx, SLOT_TILES, sizeof(swar_work) + sizeof(int) * (2 * (size_t)max_windows + 8));
    if (!buf) return TC_ERR_NOMEM;
    swar_work* work = (swar_work*)buf;
    int* win_lo = buf + sizeof(swar_work) / sizeof(int);
    int* unit_prefix = win_lo + max_windows + 2;
    swar_units_kernel<<<1, 256, 0, s>>>(a, work, win_lo, unit_prefix, max_windows);
    TC_LAUNCH_CHECK();
    const size_t smem = sizeof(uint32_t) * SMEM_WORDS;
    TC_CUDA(cudaFuncSetAttribute(swar_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    swar_main_kernel<<<ctx->sm_count * 2, T, smem, s>>>(a, work, win_lo, unit_prefix);
    TC_LAUNCH_CHECK();
    return TC_OK;
}
