// call.cu — kernel (4): the per-position consensus / IUPAC call.
//
// One thread per reference position restates, on the seven counters of that position,
//   GetDistribution + GetNucleotide   TrueConsense/Sequences.py:119-165
//   IsAmbiguous and helpers           TrueConsense/Ambig.py:102-228
//   MinorityDel                       TrueConsense/Events.py:85-106
//   the ListInserts threshold         TrueConsense/Events.py:25-36
// and emits the candidate table (tc_call_table_t) the sequential walk of
// Sequences.BuildConsensus reads.  A second, single-CTA kernel turns the PRIMARY_X flags into
// len(WalkForward(p)) for every p (Sequences.py:44-52) with a reverse segmented scan.
//
// The percentage tests are IEEE-double (count / cov) * 100 with division and multiplication
// rounded separately, exactly like CPython's float arithmetic: __ddiv_rn / __dmul_rn / __dsub_rn
// keep ptxas from contracting anything (the file is also built with --fmad=false).  Integer
// cross-multiplication gives different answers at the thresholds (55/45 at
// coverage 100 is not ambiguous, 6/5 at coverage 10 is).
#include "tc_common.cuh"

__device__ __forceinline__ double pct(int c, int cov) { return __dmul_rn(__ddiv_rn((double)c, (double)cov), 100.0); }
__device__ __forceinline__ bool within(double a, double b, double maxdist) { return fabs(__dsub_rn(a, b)) <= maxdist; }

// letters indexed 0..4 = A,T,C,G,X in the reference's dict order; tie-break by the letter's
// ASCII value (Sequences.py:137-140 sorts (count, letter) tuples)
struct ranked5 { int cnt[5]; char let[5]; };

__device__ __forceinline__ void rank5(const int c[5], ranked5& r) {
    // rank of letter i = how many letters beat it: larger count, or equal count and larger ASCII
    const int asc[5] = {'A', 'T', 'C', 'G', 'X'};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        int beat = 0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (j != i && (c[j] > c[i] || (c[j] == c[i] && asc[j] > asc[i]))) ++beat;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (beat == k) { r.cnt[k] = c[i]; r.let[k] = (char)asc[i]; }
    }
}

// Ambig.py:18-99: the IUPAC code of a set of 2 or 3 distinct letters out of A,C,G,T
__device__ __forceinline__ char iupac_of_mask(int m) {
    // bit0 A, bit1 C, bit2 G, bit3 T
    switch (m) {
        case 0x3: return 'M'; case 0x5: return 'R'; case 0x9: return 'W'; case 0x6: return 'S';
        case 0xA: return 'Y'; case 0xC: return 'K'; case 0x7: return 'V'; case 0xB: return 'H';
        case 0xD: return 'D'; case 0xE: return 'B'; default: return 0;
    }
}
__device__ __forceinline__ int letter_bit(char l) { return l == 'A' ? 1 : l == 'C' ? 2 : l == 'G' ? 4 : l == 'T' ? 8 : 0; }

// Ambig.py:179-228
__device__ __forceinline__ char is_ambiguous(const char let[4], const int cnt[4], int cov, double maxdist) {
    if (cov == 0) return 0;
    if (let[0] == 'X' || let[1] == 'X') return 0;
    double p1 = pct(cnt[0], cov), p2 = pct(cnt[1], cov), p3 = pct(cnt[2], cov), p4 = pct(cnt[3], cov);
    if (!within(p1, p2, maxdist)) return 0;                                     // AmbiguityType: None
    if (within(p1, p3, maxdist) && within(p2, p3, maxdist)) {
        if (within(p1, p4, maxdist) && within(p2, p4, maxdist) && within(p3, p4, maxdist)) return 'N';   // type 4
        if (let[2] == 'X') return 'N';                                          // type 3 with X in the top three
        return iupac_of_mask(letter_bit(let[0]) | letter_bit(let[1]) | letter_bit(let[2]));
    }
    return iupac_of_mask(letter_bit(let[0]) | letter_bit(let[1]));              // type 2
}

struct call_args {
    const int32_t* counts; int L;
    int mincov; int include_ambig; double maxdist, mdel_pct, ins_pct;
    uint8_t* call_char; uint8_t* flags; int32_t* xrun; uint8_t* rank_letter; int32_t* rank_count; uint8_t* ambig_char;
};

__global__ void __launch_bounds__(256) call_kernel(call_args a) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.L) return;
    const size_t L = (size_t)a.L;
    const int cov = a.counts[TC_ROW_COV * L + p];
    int c[5];
    c[0] = a.counts[TC_ROW_A * L + p]; c[1] = a.counts[TC_ROW_T * L + p]; c[2] = a.counts[TC_ROW_C * L + p];
    c[3] = a.counts[TC_ROW_G * L + p]; c[4] = a.counts[TC_ROW_X * L + p];
    const int ins = a.counts[TC_ROW_I * L + p];
    ranked5 r;
    rank5(c, r);
    char amb = is_ambiguous(r.let, r.cnt, cov, a.maxdist);
    unsigned f = 0;
    if (cov < a.mincov) f |= TC_CF_LOWCOV;
    if (cov > a.mincov) f |= TC_CF_COV_GT_MINCOV;
    if (cov == 0) f |= TC_CF_ZERO_COV;
    if (r.let[0] == 'X') f |= TC_CF_PRIMARY_X;
    if (cov != 0 && pct(c[4], cov) >= a.mdel_pct) f |= TC_CF_MINORITY_DEL;                       // Events.py:102-106
    if (!(cov < a.mincov) && cov != 0 && ins != 0 && pct(ins, cov) > a.ins_pct) f |= TC_CF_INS_CANDIDATE;   // Events.py:29-36
    if (amb) f |= TC_CF_AMBIG;
    char ch;
    if (r.let[0] != 'X') {
        if (a.include_ambig && amb) ch = amb;
        else ch = (r.cnt[0] < a.mincov) ? (char)(r.let[0] | 0x20) : r.let[0];
    } else {
        ch = (r.cnt[1] < a.mincov) ? (char)(r.let[1] | 0x20) : r.let[1];
    }
    if (a.call_char) a.call_char[p] = (uint8_t)ch;
    a.flags[p] = (uint8_t)f;
    if (a.ambig_char) a.ambig_char[p] = (uint8_t)amb;
    if (a.rank_letter) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a.rank_letter[k * L + p] = (uint8_t)r.let[k];
    }
    if (a.rank_count) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a.rank_count[k * L + p] = r.cnt[k];
    }
}

// xrun[j] = number of consecutive PRIMARY_X positions starting at j+1.  One CTA; thread t owns a
// contiguous chunk, scans it backwards, thread 0 chains the chunk carries, chunks patch their tails.
__global__ void __launch_bounds__(1024) xrun_kernel(uint8_t* __restrict__ flags, int32_t* __restrict__ xrun, int L) {
    __shared__ int lead[1024];      // consecutive X from the chunk's first position
    __shared__ int full[1024];      // chunk is all X
    __shared__ int carry[1024];     // consecutive X starting right after the chunk
    const int t = threadIdx.x;
    const int per = (L + 1023) / 1024;
    const int a = min(L, t * per), b = min(L, a + per);
    int run = 0, lastnon = a - 1;   // last non-X position inside the chunk
    bool seen_non = false;
    for (int j = b - 1; j >= a; --j) {
        xrun[j] = run;
        if (flags[j] & TC_CF_PRIMARY_X) ++run; else { run = 0; if (!seen_non) { seen_non = true; lastnon = j; } }
    }
    // after the loop `run` = consecutive X from position a (the chunk's lead)
    lead[t] = run; full[t] = (b > a && !seen_non) || (b == a);
    __syncthreads();
    if (t == 0) {
        int c = 0;
        for (int k = 1023; k >= 0; --k) {
            carry[k] = c;
            int ka = min(L, k * per), kb = min(L, ka + per);
            if (kb > ka) c = full[k] ? lead[k] + c : lead[k];
        }
    }
    __syncthreads();
    const int cin = carry[t];
    // positions whose local run reaches the end of the chunk: j >= lastnon (all of j+1..b-1 are X)
    int from = seen_non ? lastnon : a;
    for (int j = max(from, a); j < b; ++j) {
        if (cin) xrun[j] += cin;
    }
    __syncthreads();
    for (int j = a; j < b; ++j) {
        if (xrun[j] == L - 1 - j) flags[j] |= TC_CF_XRUN_OFF_END;
    }
}

__global__ void is_ambiguous_kernel(const uint8_t* __restrict__ letters, const int32_t* __restrict__ cnts,
                                    const int32_t* __restrict__ cov, int64_t n, double maxdist, uint8_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    char let[4]; int cnt[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { let[k] = (char)letters[k * n + i]; cnt[k] = cnts[k * n + i]; }
    out[i] = (uint8_t)is_ambiguous(let, cnt, cov[i], maxdist);
}

// ordered compaction of the INS_CANDIDATE positions (1-based) — single CTA, L is small
__global__ void __launch_bounds__(1024) list_candidates_kernel(const uint8_t* __restrict__ flags, int L, int32_t* __restrict__ out,
                                                               int cap, int32_t* __restrict__ n_out) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < L; b0 += 1024) {
        int j = b0 + threadIdx.x;
        int is = (j < L) && (flags[j] & TC_CF_INS_CANDIDATE);
        unsigned m = __ballot_sync(0xffffffffu, is);
        int within_w = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        int total = 0;
        for (int w = 0; w < 32; ++w) total += wsum[w];
        int slot = base_s + woff + within_w;
        if (is && slot < cap) out[slot] = j + 1;
        __syncthreads();
        if (threadIdx.x == 0) base_s += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base_s;
}

// ---------------------------------------------------------------- C-ABI
TC_API int tc_call(tc_ctx_t* ctx, const int32_t* counts, int32_t ref_len, const tc_call_params_t* p,
                   const tc_call_table_t* t, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!counts || !p || !t) return tc_fail(ctx, TC_ERR_ARG, "NULL argument");
    if (ref_len <= 0) return tc_fail(ctx, TC_ERR_ARG, "ref_len must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    TC_CUDA(cudaSetDevice(ctx->device));
    const size_t L = (size_t)ref_len;
    int rc;
    call_args a;
    a.counts = (const int32_t*)tc_stage_in(ctx, SLOT_COUNTS, counts, sizeof(int32_t) * TC_NROWS * L, s, &rc);
    if (rc) return rc;
    a.L = ref_len; a.mincov = p->mincov; a.include_ambig = p->include_ambig;
    a.maxdist = p->ambig_maxdist; a.mdel_pct = p->minority_del_pct; a.ins_pct = p->insert_pct;
    struct out_t { void* user; void* dev; size_t bytes; int slot; };
    out_t outs[6] = {
        {t->call_char, nullptr, L, SLOT_CALL_A}, {t->flags, nullptr, L, SLOT_CALL_B}, {t->xrun, nullptr, 4 * L, SLOT_CALL_C},
        {t->rank_letter, nullptr, 4 * L, SLOT_CALL_D}, {t->rank_count, nullptr, 16 * L, SLOT_CALL_E}, {t->ambig_char, nullptr, L, SLOT_CALL_F}};
    for (int i = 0; i < 6; ++i) {
        bool need = outs[i].user != nullptr || i == 1;     // flags feed the xrun kernel
        if (!need) continue;
        if (outs[i].user && tc_is_device_ptr(outs[i].user)) outs[i].dev = outs[i].user;
        else { outs[i].dev = tc_dev_buf(ctx, outs[i].slot, outs[i].bytes); if (!outs[i].dev) return TC_ERR_NOMEM; }
    }
    a.call_char = (uint8_t*)outs[0].dev; a.flags = (uint8_t*)outs[1].dev; a.xrun = (int32_t*)outs[2].dev;
    a.rank_letter = (uint8_t*)outs[3].dev; a.rank_count = (int32_t*)outs[4].dev; a.ambig_char = (uint8_t*)outs[5].dev;
    call_kernel<<<(unsigned)((L + 255) / 256), 256, 0, s>>>(a);
    TC_LAUNCH_CHECK();
    if (a.xrun) {
        xrun_kernel<<<1, 1024, 0, s>>>(a.flags, a.xrun, ref_len);
        TC_LAUNCH_CHECK();
    }
    bool any_host = false;
    for (int i = 0; i < 6; ++i) {
        if (outs[i].user && outs[i].dev != outs[i].user) {
            TC_D2H(outs[i].user, outs[i].dev, outs[i].bytes, s);
            any_host = true;
        }
    }
    if (any_host) TC_CUDA(cudaStreamSynchronize(s));
    return TC_OK;
}

TC_API int tc_is_ambiguous(tc_ctx_t* ctx, const uint8_t* letters, const int32_t* cnts, const int32_t* cov, int64_t n,
                           double maxdist, uint8_t* out_char, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!letters || !cnts || !cov || !out_char || n < 0) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    if (n == 0) return TC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    TC_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const uint8_t* dl = (const uint8_t*)tc_stage_in(ctx, SLOT_TMP_A, letters, 4 * (size_t)n, s, &rc); if (rc) return rc;
    const int32_t* dc = (const int32_t*)tc_stage_in(ctx, SLOT_TMP_B, cnts, 16 * (size_t)n, s, &rc); if (rc) return rc;
    const int32_t* dv = (const int32_t*)tc_stage_in(ctx, SLOT_TMP_C, cov, 4 * (size_t)n, s, &rc); if (rc) return rc;
    bool out_dev = tc_is_device_ptr(out_char);
    uint8_t* d_out = out_dev ? out_char : (uint8_t*)tc_dev_buf(ctx, SLOT_CALL_F, (size_t)n);
    if (!d_out) return TC_ERR_NOMEM;
    is_ambiguous_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dl, dc, dv, n, maxdist, d_out);
    TC_LAUNCH_CHECK();
    if (!out_dev) {
        TC_D2H(out_char, d_out, (size_t)n, s);
        TC_CUDA(cudaStreamSynchronize(s));
    }
    return TC_OK;
}

TC_API int tc_list_insert_candidates(tc_ctx_t* ctx, const uint8_t* flags, int32_t ref_len, int32_t* cand_pos, int32_t cap,
                                     int32_t* n_out, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!flags || !cand_pos || !n_out || ref_len <= 0 || cap < 0) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    TC_CUDA(cudaSetDevice(ctx->device));
    int rc;
    const uint8_t* df = (const uint8_t*)tc_stage_in(ctx, SLOT_TMP_A, flags, (size_t)ref_len, s, &rc); if (rc) return rc;
    int32_t* d_out = (int32_t*)tc_dev_buf(ctx, SLOT_TMP_B, 4 * (size_t)(cap + 1) + 16);
    if (!d_out) return TC_ERR_NOMEM;
    list_candidates_kernel<<<1, 1024, 0, s>>>(df, ref_len, d_out + 1, cap, d_out);
    TC_LAUNCH_CHECK();
    int32_t n = 0;
    TC_D2H(ctx->host_status, d_out, 4, s);
    TC_CUDA(cudaStreamSynchronize(s));
    n = *(int32_t*)ctx->host_status;
    *n_out = n;
    if (n > cap) return tc_fail(ctx, TC_ERR_CAPACITY, "%d insertion candidates, capacity %d", n, cap);
    if (n > 0) {
        if (tc_is_device_ptr(cand_pos)) TC_CUDA(cudaMemcpyAsync(cand_pos, d_out + 1, 4 * (size_t)n, cudaMemcpyDeviceToDevice, s));
        else TC_D2H(cand_pos, d_out + 1, 4 * (size_t)n, s);
        TC_CUDA(cudaStreamSynchronize(s));
    }
    return TC_OK;
}
