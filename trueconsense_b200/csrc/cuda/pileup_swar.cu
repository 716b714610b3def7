// pileup_swar.cu — variant 2 of the pileup kernel: no atomic per base.
//
// Why: the scatter kernel (pileup.cu) pays one shared-memory atomic per aligned base, and
// spread-address shared atomics run at ~0.5 lane/clk/SM on this architecture
// (B300_MICROARCH.md "Atomics") — two orders of magnitude under what HBM can feed
// (profiles/r1_*: 0.9 % of the HBM roofline).  Here the per-base work is bit-parallel:
//
//   work unit   = (reference window of W columns) x (<= UNIT_READS start-sorted reads that can
//                 overlap it); units are handed to persistent CTAs through an atomic counter.
//   stage       the raw packed SEQ words and CIGAR ops of a sub-tile of T reads are contiguous in
//                 HBM (reads are stored back to back), so they come in as two coalesced 16-byte
//                 streams into shared memory.
//   expand      one lane per read walks its CIGAR once (deletion / insertion events go to small
//                 shared counters — they are sparse), turns every M/=/X op into a window-clipped
//                 segment descriptor, then funnel-shifts its packed bases into a reference-aligned
//                 row of 4-bit codes in shared memory (8 columns per 32-bit word, row stride padded
//                 to W/8+1 words so the 32 lanes of a warp hit 32 different banks).  Codes that are
//                 not exactly A/C/G/T are cleared on the way: they only count towards coverage,
//                 which comes from the difference array.
//   column sum  BAM's base codes are one-hot (A=1, C=2, G=4, T=8), so counting A/C/G/T in a column
//                 is a positional popcount over the rows: every thread owns one row-word column
//                 (8 reference columns x 4 classes = 32 bit positions) and adds 16 rows at a time
//                 with a Harley-Seal carry-save tree (15 full adders = 30 LOP3) into bit-sliced
//                 counters that live in registers for the whole unit; they are turned into integers
//                 and flushed to HBM once per unit.
#include "pileup.cuh"

namespace {

constexpr int W = 512;              // window columns
constexpr int WC = W / 8;           // row words
constexpr int RS = WC + 1;          // padded row stride (words)
constexpr int T = 128;              // reads per sub-tile == threads per CTA
constexpr int SEQ_CAP = 8192;       // staged SEQ words per sub-tile (incl. pads)
constexpr int CIG_CAP = 5120;       // staged CIGAR ops per sub-tile
constexpr int SEQ_PAD = 4;          // zero words in front of the staged SEQ stream
constexpr int UNIT_READS = 2048;
constexpr int HI_PLANES = 8;        // bit-sliced counter planes above "eights": up to 16 * 255 rows per thread
constexpr int SMEM_WORDS = T * RS + SEQ_CAP + 8 + CIG_CAP + 8 + 2 * W;

struct swar_work {
    int n_windows;
    int total_units;
    int counter;
    int pad;
};

__global__ void swar_units_kernel(pileup_args a, swar_work* work, int* __restrict__ win_lo, int* __restrict__ win_hi,
                                  int* __restrict__ unit_prefix, int n_windows) {
    const int ms = max(a.status->max_span, 1);
    auto lower = [&](int v) {
        int64_t lo = 0, hi = a.r.n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a.r.pos[mid] < v) lo = mid + 1; else hi = mid; }
        return (int)lo;
    };
    for (int w = threadIdx.x; w < n_windows; w += blockDim.x) {
        int w0 = w * W;
        win_lo[w] = lower(w0 - ms + 1);
        win_hi[w] = lower(w0 + W);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < n_windows; ++w) {
            unit_prefix[w] = acc;
            acc += (win_hi[w] - win_lo[w] + UNIT_READS - 1) / UNIT_READS;
        }
        unit_prefix[n_windows] = acc;
        work->n_windows = n_windows; work->total_units = acc; work->counter = 0;
    }
}

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    l = a ^ b ^ c;                      // one LOP3
    h = (a & b) | (c & (a | b));        // one LOP3 (majority)
}

__global__ void __launch_bounds__(T, 2) swar_main_kernel(pileup_args a, swar_work* work, const int* __restrict__ win_lo,
                                                         const int* __restrict__ win_hi, const int* __restrict__ unit_prefix) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* rows = smem;                              // [T][RS]
    uint32_t* seq_s = rows + T * RS;                    // [SEQ_CAP + 8]; index SEQ_PAD = word sbase_al
    uint32_t* cig_s = seq_s + SEQ_CAP + 8;              // [CIG_CAP + 8]
    int* xcnt = (int*)(cig_s + CIG_CAP + 8);            // [W]
    int* icnt = xcnt + W;                               // [W]
    __shared__ int s_unit;
    const int tid = threadIdx.x;
    const int L = a.L;
    const int64_t n_seq_words = (int64_t)a.r.seq_off[a.r.n];
    const int64_t n_ops_total = (int64_t)a.r.cigar_off[a.r.n];

    for (int i = tid; i < T * RS; i += T) rows[i] = 0;
    for (int i = tid; i < 2 * W; i += T) xcnt[i] = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) s_unit = atomicAdd(&work->counter, 1);
        __syncthreads();
        const int u = s_unit;
        __syncthreads();
        if (u >= work->total_units) break;
        // unit -> (window, read range)
        int wlo = 0, whi = work->n_windows;
        while (whi - wlo > 1) { int mid = (wlo + whi) >> 1; if (unit_prefix[mid] <= u) wlo = mid; else whi = mid; }
        const int w = wlo;
        const int w0 = w * W, w1 = w0 + W;
        const int u0 = win_lo[w] + (u - unit_prefix[w]) * UNIT_READS;
        const int u1 = min(win_hi[w], u0 + UNIT_READS);

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
            // ---- stage SEQ words [sbase_al, send) and CIGAR ops [cbase_al, cend) with 16-byte copies
            {
                const uint32_t send = a.r.seq_off[t0 + n] + 1;           // one word of look-ahead for the funnel shift
                const uint32_t nw = send - sbase_al;
                const uint4* src = reinterpret_cast<const uint4*>(a.r.seq4 + sbase_al);
                uint4* dst = reinterpret_cast<uint4*>(seq_s + SEQ_PAD);
                const int nv = (int)((nw + 3) >> 2);
                for (int i = tid; i < nv; i += T) {
                    int64_t wbase = (int64_t)sbase_al + 4 * i;
                    uint4 v;
                    if (wbase + 4 <= n_seq_words) v = __ldg(src + i);
                    else {
                        v.x = wbase + 0 < n_seq_words ? __ldg(a.r.seq4 + wbase + 0) : 0u;
                        v.y = wbase + 1 < n_seq_words ? __ldg(a.r.seq4 + wbase + 1) : 0u;
                        v.z = wbase + 2 < n_seq_words ? __ldg(a.r.seq4 + wbase + 2) : 0u;
                        v.w = wbase + 3 < n_seq_words ? __ldg(a.r.seq4 + wbase + 3) : 0u;
                    }
                    dst[i] = v;
                }
                if (tid < SEQ_PAD) seq_s[tid] = 0;
                const uint32_t cend = a.r.cigar_off[t0 + n];
                const uint32_t nc = cend - cbase_al;
                const uint4* csrc = reinterpret_cast<const uint4*>(a.r.cigar + cbase_al);
                uint4* cdst = reinterpret_cast<uint4*>(cig_s);
                const int ncv = (int)((nc + 3) >> 2);
                for (int i = tid; i < ncv; i += T) {
                    int64_t obase = (int64_t)cbase_al + 4 * i;
                    uint4 v;
                    if (obase + 4 <= n_ops_total) v = __ldg(csrc + i);
                    else {
                        v.x = obase + 0 < n_ops_total ? __ldg(a.r.cigar + obase + 0) : 0u;
                        v.y = obase + 1 < n_ops_total ? __ldg(a.r.cigar + obase + 1) : 0u;
                        v.z = obase + 2 < n_ops_total ? __ldg(a.r.cigar + obase + 2) : 0u;
                        v.w = obase + 3 < n_ops_total ? __ldg(a.r.cigar + obase + 3) : 0u;
                    }
                    cdst[i] = v;
                }
            }
            __syncthreads();
            // ---- expand: one lane per read
            if (tid < n) {
                const int64_t r = t0 + tid;
                if (read_passes(a, r)) {
                    const int p = a.r.pos[r];
                    const int lq = a.r.l_seq[r];
                    uint32_t* cs = cig_s + (a.r.cigar_off[r] - cbase_al);
                    const int nops = (int)(a.r.cigar_off[r + 1] - a.r.cigar_off[r]);
                    int nseg = 0, qfirst = 0;
                    int x = p, y = 0;
                    for (int k = 0; k < nops; ++k) {
                        const uint32_t c = cs[k];
                        const uint32_t op = c & 15u;
                        const int l = (int)(c >> 4);
                        if (!op_consumes_ref(op)) { if (op == OP_I || op == OP_S) y += l; continue; }
                        if (x >= w1) break;
                        const int xe = x + l;
                        const bool match = op_is_match(op);
                        if (xe > w0) {
                            const int last = xe - 1;
                            int indel = 0;
                            if (last >= w0 && last < w1 && k + 1 < nops) {
                                const uint32_t nx = cs[k + 1] & 15u;
                                if (nx == OP_I || nx == OP_P) indel = peek_indel(cs, nops, k);
                            }
                            if (match) {
                                const int a0 = max(x, w0), b0 = min(xe, w1);
                                const int q0 = y + (a0 - x);
                                int len = b0 - a0;
                                if (q0 + len > lq) len = lq - q0;            // SEQ '*' or short SEQ: those columns read 'N'
                                if (len > 0) {
                                    if (nseg == 0) qfirst = q0;
                                    const int qrel = q0 - qfirst;
                                    if (qrel >= (1 << 13)) atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY);
                                    cs[nseg++] = (uint32_t)(a0 - w0) | ((uint32_t)(len - 1) << 9) | ((uint32_t)(qrel & 0x1fff) << 18);
                                }
                            } else if (op == OP_D) {
                                const int c0 = max(x, w0), c1 = min(xe, w1);
                                for (int col = c0; col < c1; ++col)
                                    if (!(col == last && indel > 0)) atomicAdd(&xcnt[col - w0], 1);
                            }
                            if (indel > 0) atomicAdd(&icnt[last - w0], 1);
                        }
                        if (match) y += l;
                        x = xe;
                    }
                    if (nseg > 0) {
                        const uint32_t* sq = seq_s + SEQ_PAD + (a.r.seq_off[r] - sbase_al);
                        uint32_t* row = rows + tid * RS;
                        int i = 0;
                        uint32_t d = cs[0];
                        int c0 = (int)(d & 511u), len = (int)((d >> 9) & 511u) + 1, q0 = qfirst + (int)(d >> 18);
                        int o = c0 >> 3, oend = (c0 + len - 1) >> 3;
                        for (;;) {
                            const int s0 = q0 + 8 * o - c0;
                            const int ws = s0 >> 3;
                            const int sh = (s0 & 7) * 4;
                            const uint32_t whi = __byte_perm(sq[ws], 0, 0x0123);
                            const uint32_t wlo2 = __byte_perm(sq[ws + 1], 0, 0x0123);
                            uint32_t v = __funnelshift_l(wlo2, whi, sh);
                            const int dlo = max(c0 - 8 * o, 0), dhi = min(c0 + len - 8 * o, 8);
                            v &= (0xffffffffu >> (4 * dlo)) & (0xffffffffu << (4 * (8 - dhi)));
                            const uint32_t z = v & ((v | 0x88888888u) - 0x11111111u);
                            if (z) {        // codes with more than one bit (N, IUPAC): not A/C/G/T
                                uint32_t m = (z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x11111111u;
                                v &= ~(m * 15u);
                            }
                            row[o] |= v;
                            if (o == oend) {
                                if (++i == nseg) break;
                                d = cs[i];
                                c0 = (int)(d & 511u); len = (int)((d >> 9) & 511u) + 1; q0 = qfirst + (int)(d >> 18);
                                o = c0 >> 3; oend = (c0 + len - 1) >> 3;
                            } else ++o;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- column sum: thread -> (row word column, half of the rows)
            {
                const int c = tid & (WC - 1);
                const int g = tid / WC;
                uint32_t* col = rows + (g * (T / 2)) * RS + c;
#pragma unroll
                for (int blk = 0; blk < (T / 2) / 16; ++blk) {
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
        {
            const int c = tid & (WC - 1);
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
                    const int colr = w0 + 8 * c + (7 - (bit >> 2));
                    const int cls = bit & 3;     // bit 0 A, 1 C, 2 G, 3 T (BAM codes 1,2,4,8)
                    const int row = cls == 0 ? TC_ROW_A : cls == 1 ? TC_ROW_C : cls == 2 ? TC_ROW_G : TC_ROW_T;
                    if (colr < L) atomicAdd(&a.counts[(size_t)row * L + colr], v);
                }
            }
            for (int i = tid; i < W; i += T) {
                int xv = xcnt[i], iv = icnt[i];
                if (xv) { if (w0 + i < L) atomicAdd(&a.counts[(size_t)TC_ROW_X * L + w0 + i], xv); xcnt[i] = 0; }
                if (iv) { if (w0 + i < L) atomicAdd(&a.counts[(size_t)TC_ROW_I * L + w0 + i], iv); icnt[i] = 0; }
            }
        }
        __syncthreads();
    }
}

}  // namespace

bool tc_pileup_swar_supported(const pileup_args& a) {
    return (((uintptr_t)a.r.seq4 | (uintptr_t)a.r.cigar) & 15u) == 0;
}

int tc_pileup_swar_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    const int n_windows = (a.L + W - 1) / W;
    int* buf = (int*)tc_dev_buf(ctx, SLOT_TILES, sizeof(swar_work) + sizeof(int) * (3 * (size_t)n_windows + 4));
    if (!buf) return TC_ERR_NOMEM;
    swar_work* work = (swar_work*)buf;
    int* win_lo = buf + sizeof(swar_work) / sizeof(int);
    int* win_hi = win_lo + n_windows;
    int* unit_prefix = win_hi + n_windows;
    swar_units_kernel<<<1, 256, 0, s>>>(a, work, win_lo, win_hi, unit_prefix, n_windows);
    TC_LAUNCH_CHECK();
    const size_t smem = sizeof(uint32_t) * SMEM_WORDS;
    TC_CUDA(cudaFuncSetAttribute(swar_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    swar_main_kernel<<<ctx->sm_count * 2, T, smem, s>>>(a, work, win_lo, win_hi, unit_prefix);
    TC_LAUNCH_CHECK();
    return TC_OK;
}
