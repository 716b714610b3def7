// inserts.cu — kernel (2): the insertion caller, TrueConsense/Events.py:47-82 (ExtractInserts).
//
// For every candidate position the reference piles up one column with pysam's DEFAULT arguments
// (Events.py:66): samtools stepper (flag filter 0x704, orphans skipped), max_depth 8000,
// min_base_quality 13; upper-cases the strings and takes collections.Counter's mode (ties: first
// encountered).  Here, per candidate column c (one CTA each):
//   select   reads the region fetch would return and the stepper would pass, in file order;
//            htslib's depth cap (bam_plp_push: a read is dropped when it starts on the column the
//            engine is waiting to emit and more than max_depth reads are live) reduces, for a
//            single-column fetch where every fetched read is still live, to
//            "admitted  <=>  first selected read at its start coordinate, or fewer than max_depth
//            selected reads before it";
//   emit     for every admitted read covering c: CIGAR walk to the column, base-quality test,
//            and a 64-bit key hashing exactly the characters pysam would print, upper-cased
//            (head character, sign and length of the indel, inserted bases);
//   count    keys are radix-sorted per candidate (cub::DeviceSegmentedRadixSort, stable) and
//            run-length encoded; the longest run wins, ties go to the run whose first entry came
//            first in the file.  Every member of every run is compared with its run head
//            character by character, so a hash collision is reported (TC_ERR_RANGE) and can never
//            silently change a count.
// Not emulated (deviations): htslib's mate-overlap quality rewriting.
#include <cub/device/device_segmented_radix_sort.cuh>

#include "tc_common.cuh"

struct ins_args {
    dreads r;
    const int32_t* span_end;        // [n] pos + reference span
    const int32_t* cand;            // [n_cand] 1-based positions
    int n_cand;
    uint32_t flag_filter; int min_mapq, min_bq, ignore_orphans; long long max_depth;
    int32_t* range;                 // [n_cand][2] lo, hi read indices
    const int64_t* seg_off;         // [n_cand+1] entry storage offsets
    uint64_t* ent_key; uint32_t* ent_idx;           // unsorted keys / entry ids
    uint32_t* ent_read; int32_t* ent_indel; int32_t* ent_qpos; uint8_t* ent_head;
    int32_t* seg_count;             // [n_cand] emitted entries
    tc_status* status;
};

__global__ void span_end_kernel(dreads r, int32_t* __restrict__ span_end, tc_status* status) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r.n) return;
    if (i > 0 && r.pos[i] < r.pos[i - 1]) atomicCAS(&status->err, 0, TC_ERR_UNSORTED);
    int span = 0;
    for (uint32_t k = r.cigar_off[i]; k < r.cigar_off[i + 1]; ++k) {
        uint32_t c = r.cigar[k];
        if (op_consumes_ref(c & 15u)) span += (int)(c >> 4);
    }
    span_end[i] = r.pos[i] + span;
    atomicMax(&status->max_span, span);
}

// reads that can overlap column c: pos in (c - max_span, c]
__global__ void cand_range_kernel(ins_args a) {
    int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci >= a.n_cand) return;
    const int c = a.cand[ci] - 1;
    const int ms = max(a.status->max_span, 1);
    auto lower = [&](int v) {   // first read with pos >= v
        int64_t lo = 0, hi = a.r.n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a.r.pos[mid] < v) lo = mid + 1; else hi = mid; }
        return (int)lo;
    };
    a.range[2 * ci] = lower(c - ms + 1);
    a.range[2 * ci + 1] = lower(c + 1);
}

__device__ __forceinline__ uint64_t mix_key(uint64_t h, uint64_t v) {
    h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    return h;
}

// the character pysam prints for base code `code` on a read of the given strand, upper-cased
__device__ __forceinline__ char base_char_upper(uint32_t code, bool rev) {
    const char nt16[17] = "=ACMGRSVTWYHKDBN";
    if (code == 0) return rev ? ',' : '.';      // strand_mark_char maps '=' to '.' / ','
    return nt16[code];
}

__device__ __forceinline__ char ins_char(const ins_args& a, uint32_t read, int qpos, int j, bool rev) {
    const int lq = a.r.l_seq[read];
    const int q = qpos + j;
    if (q >= lq) return 'N';
    return base_char_upper(seq_code(a.r.seq4 + a.r.seq_off[read], q), rev);
}

__global__ void __launch_bounds__(1024) ins_select_emit_kernel(ins_args a) {
    __shared__ int wsum[32];
    __shared__ int wmax[32];
    __shared__ long long sel_base_s;
    __shared__ int emit_base_s;
    __shared__ int last_sel_s;
    const int ci = blockIdx.x;
    const int c = a.cand[ci] - 1;
    const int lo = a.range[2 * ci], hi = a.range[2 * ci + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t off = a.seg_off[ci];
    if (threadIdx.x == 0) { sel_base_s = 0; emit_base_s = 0; last_sel_s = -1; }
    __syncthreads();
    for (int t0 = lo; t0 < hi; t0 += 1024) {
        const int r = t0 + threadIdx.x;
        bool sel = false; int pos = 0, end = 0;
        if (r < hi) {
            pos = a.r.pos[r]; end = a.span_end[r];
            uint32_t fl = a.r.flag[r];
            bool pass = !(fl & (a.flag_filter | 4u)) && !(a.min_mapq > 0 && a.r.mapq && (int)a.r.mapq[r] < a.min_mapq) &&
                        !(a.ignore_orphans && (fl & 1u) && !(fl & 2u));
            bool fetched = pos <= c && (end > c || (end == pos && pos == c));   // BAI query: pos < c+1 && endpos > c
            sel = pass && fetched;
        }
        // nearest earlier selected read (inclusive max-scan of selected indices)
        int idx = sel ? r : -1;
        int m = idx;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, m, o); if (lane >= o) m = max(m, t); }
        int prev_in_warp = __shfl_up_sync(0xffffffffu, m, 1);
        if (lane == 0) prev_in_warp = -1;
        if (lane == 31) wmax[warp] = m;
        unsigned bal = __ballot_sync(0xffffffffu, sel);
        int sel_before_w = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int prev_sel = last_sel_s, sel_before = 0, sel_total = 0;
        for (int w = 0; w < 32; ++w) {
            if (w < warp) { prev_sel = max(prev_sel, wmax[w]); sel_before += wsum[w]; }
            sel_total += wsum[w];
        }
        prev_sel = max(prev_sel, prev_in_warp);
        int tile_last = last_sel_s;
        for (int w = 0; w < 32; ++w) tile_last = max(tile_last, wmax[w]);
        bool first_at_start = sel && (prev_sel < 0 || a.r.pos[prev_sel] != pos);
        // a read without reference span is only linked into htslib's list when it opens its start coordinate;
        // it never yields an entry, and it only counts towards the live total when linked
        bool zero_span = sel && end == pos;
        long long rank = sel_base_s + sel_before_w + sel_before;
        bool admitted = sel && (first_at_start || rank < a.max_depth);
        // (zero-span reads that are not first at their start are not counted; the approximation only matters
        //  when such reads sit exactly on a candidate column and the cap binds at the same time)
        bool emit = false; int indel = 0, qpos = 0; char head = 0; uint64_t key = 0;
        if (admitted && !zero_span && end > c) {
            const uint32_t c0 = a.r.cigar_off[r];
            const int n = (int)(a.r.cigar_off[r + 1] - c0);
            const uint32_t* cig = a.r.cigar + c0;
            const bool rev = (a.r.flag[r] & 16u) != 0;
            const int lq = a.r.l_seq[r];
            int x = pos, y = 0;
            for (int k = 0; k < n; ++k) {
                uint32_t cc = cig[k]; uint32_t op = cc & 15u; int l = (int)(cc >> 4);
                if (!op_consumes_ref(op)) { if (op == OP_I || op == OP_S) y += l; continue; }
                if (c < x + l) {
                    bool match = op_is_match(op);
                    qpos = match ? y + (c - x) : y;
                    if (c == x + l - 1) indel = peek_indel(cig, n, k);
                    int qv = (qpos < lq) ? (int)a.r.qual[8ull * a.r.seq_off[r] + qpos] : 0;
                    if (qv >= a.min_bq) {
                        emit = true;
                        if (match) head = (qpos < lq) ? base_char_upper(seq_code(a.r.seq4 + a.r.seq_off[r], qpos), rev) : 'N';
                        else head = (op == OP_N) ? (rev ? '<' : '>') : '*';
                    }
                    break;
                }
                if (op_is_match(op)) y += l;
                x += l;
            }
            if (emit) {
                key = mix_key(0x7463696e73ull, (uint64_t)(uint8_t)head);
                key = mix_key(key, (uint64_t)(uint32_t)indel);
                for (int j = 1; j <= indel; ++j) key = mix_key(key, (uint64_t)(uint8_t)ins_char(a, r, qpos, j, rev));
            }
        }
        __syncthreads();
        unsigned ebal = __ballot_sync(0xffffffffu, emit);
        int emit_before_w = __popc(ebal & ((1u << lane) - 1));
        if (lane == 0) wsum[warp] = __popc(ebal);
        __syncthreads();
        int emit_before = 0, emit_total = 0;
        for (int w = 0; w < 32; ++w) { if (w < warp) emit_before += wsum[w]; emit_total += wsum[w]; }
        if (emit) {
            int64_t slot = off + emit_base_s + emit_before + emit_before_w;
            a.ent_key[slot] = key; a.ent_idx[slot] = (uint32_t)slot;
            a.ent_read[slot] = (uint32_t)r; a.ent_indel[slot] = indel; a.ent_qpos[slot] = qpos; a.ent_head[slot] = (uint8_t)head;
        }
        __syncthreads();
        if (threadIdx.x == 0) { sel_base_s += sel_total; emit_base_s += emit_total; last_sel_s = tile_last; }
        __syncthreads();
    }
    if (threadIdx.x == 0) a.seg_count[ci] = emit_base_s;
}

__device__ __forceinline__ bool same_entry(const ins_args& a, uint32_t e1, uint32_t e2) {
    if (a.ent_head[e1] != a.ent_head[e2] || a.ent_indel[e1] != a.ent_indel[e2]) return false;
    int indel = a.ent_indel[e1];
    uint32_t r1 = a.ent_read[e1], r2 = a.ent_read[e2];
    bool v1 = (a.r.flag[r1] & 16u) != 0, v2 = (a.r.flag[r2] & 16u) != 0;
    for (int j = 1; j <= indel; ++j)
        if (ins_char(a, r1, a.ent_qpos[e1], j, v1) != ins_char(a, r2, a.ent_qpos[e2], j, v2)) return false;
    return true;
}

// one CTA per candidate over its sorted keys
__global__ void __launch_bounds__(1024) ins_mode_kernel(ins_args a, const uint64_t* __restrict__ skey, const uint32_t* __restrict__ sidx,
                                                        tc_insert_call_t* __restrict__ calls) {
    __shared__ unsigned long long best_s;
    const int ci = blockIdx.x;
    const int64_t off = a.seg_off[ci];
    const int m = a.seg_count[ci];
    if (threadIdx.x == 0) best_s = 0ull;
    __syncthreads();
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const uint64_t k = skey[off + i];
        // run head = first element with this key
        int lo = 0, hi = i;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (skey[off + mid] < k) lo = mid + 1; else hi = mid; }
        const int head = lo;
        if (head != i) {
            if (!same_entry(a, sidx[off + head], sidx[off + i])) atomicCAS(&a.status->err, 0, TC_ERR_RANGE);
        } else {
            int l2 = i, h2 = m;
            while (l2 < h2) { int mid = (l2 + h2) >> 1; if (skey[off + mid] <= k) l2 = mid + 1; else h2 = mid; }
            unsigned long long cnt = (unsigned long long)(l2 - i);
            // stable sort: the head of a run is its earliest entry in file order
            unsigned long long first = (unsigned long long)(sidx[off + i] - (uint32_t)off);
            unsigned long long v = (cnt << 32) | (0xffffffffull - first);
            best = max(best, v);
        }
    }
    atomicMax(&best_s, best);
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_insert_call_t out;
        out.pos = a.cand[ci]; out.n_entries = m; out.mode_count = 0; out.first_read = -1; out.head = 0; out.indel = 0; out.bases_off = -1;
        if (m > 0) {
            unsigned long long v = best_s;
            uint32_t e = (uint32_t)off + (uint32_t)(0xffffffffull - (v & 0xffffffffull));
            out.mode_count = (int32_t)(v >> 32);
            out.first_read = (int32_t)a.ent_read[e];
            out.head = a.ent_head[e];
            out.indel = a.ent_indel[e];
            out.bases_off = (int64_t)e;         // entry id for now; the host turns it into a buffer offset
        }
        calls[ci] = out;
    }
}

__global__ void seg_end_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ cnt, int64_t* __restrict__ end, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) end[i] = off[i] + cnt[i];
}

__global__ void ins_bases_kernel(ins_args a, const tc_insert_call_t* __restrict__ calls, const int64_t* __restrict__ entry_of,
                                 uint8_t* __restrict__ bases) {
    const int ci = blockIdx.x;
    const tc_insert_call_t c = calls[ci];
    if (c.indel <= 0) return;
    const uint32_t e = (uint32_t)entry_of[ci];
    const uint32_t r = a.ent_read[e];
    const bool rev = (a.r.flag[r] & 16u) != 0;
    for (int j = threadIdx.x; j < c.indel; j += blockDim.x) bases[c.bases_off + j] = (uint8_t)ins_char(a, r, a.ent_qpos[e], j + 1, rev);
}

TC_API int tc_extract_inserts(tc_ctx_t* ctx, const tc_reads_t* reads, int32_t ref_len, const int32_t* cand_pos, int32_t n_cand,
                              const tc_pileup_params_t* p, tc_insert_call_t* calls, uint8_t* bases, int64_t bases_cap, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!reads || !p || (n_cand > 0 && (!cand_pos || !calls)) || n_cand < 0 || ref_len <= 0) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    if (n_cand == 0) return TC_OK;
    if (tc_is_device_ptr(cand_pos) || tc_is_device_ptr(calls) || (bases && tc_is_device_ptr(bases)))
        return tc_fail(ctx, TC_ERR_ARG, "tc_extract_inserts takes host pointers for cand_pos / calls / bases");
    for (int i = 0; i < n_cand; ++i)
        if (cand_pos[i] < 1 || cand_pos[i] > ref_len) return tc_fail(ctx, TC_ERR_ARG, "candidate position %d outside 1..%d", cand_pos[i], ref_len);
    cudaStream_t s = (cudaStream_t)stream;
    TC_CUDA(cudaSetDevice(ctx->device));
    ins_args a;
    memset(&a, 0, sizeof(a));
    int rc = tc_resolve_reads(ctx, reads, &a.r, NEED_QUAL | NEED_MATE, s);
    if (rc) return rc;
    const int64_t n = a.r.n;
    for (int i = 0; i < n_cand; ++i) { calls[i].pos = cand_pos[i]; calls[i].n_entries = 0; calls[i].mode_count = 0; calls[i].first_read = -1; calls[i].head = 0; calls[i].indel = 0; calls[i].bases_off = -1; }
    if (n == 0) return TC_OK;
    int32_t* d_end = (int32_t*)tc_dev_buf(ctx, SLOT_SPAN_END, 4 * (size_t)n);
    tc_status* d_status = (tc_status*)tc_dev_buf(ctx, SLOT_STATUS, sizeof(tc_status));
    int32_t* d_cand = (int32_t*)tc_dev_buf(ctx, SLOT_INS_A, 4 * (size_t)n_cand);
    int32_t* d_range = (int32_t*)tc_dev_buf(ctx, SLOT_INS_B, 8 * (size_t)n_cand);
    if (!d_end || !d_status || !d_cand || !d_range) return TC_ERR_NOMEM;
    TC_CUDA(cudaMemsetAsync(d_status, 0, sizeof(tc_status), s));
    TC_CUDA(cudaMemcpyAsync(d_cand, cand_pos, 4 * (size_t)n_cand, cudaMemcpyHostToDevice, s));
    span_end_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a.r, d_end, d_status);
    TC_LAUNCH_CHECK();
    a.span_end = d_end; a.cand = d_cand; a.n_cand = n_cand; a.range = d_range; a.status = d_status;
    a.flag_filter = p->flag_filter; a.min_mapq = p->min_mapq; a.min_bq = p->min_base_quality; a.ignore_orphans = p->ignore_orphans;
    a.max_depth = p->max_depth > 0 ? p->max_depth : (1ll << 62);
    cand_range_kernel<<<(n_cand + 127) / 128, 128, 0, s>>>(a);
    TC_LAUNCH_CHECK();
    int32_t* h_range = (int32_t*)malloc(8 * (size_t)n_cand);
    int64_t* h_off = (int64_t*)malloc(8 * ((size_t)n_cand + 1));
    if (!h_range || !h_off) { free(h_range); free(h_off); return tc_fail(ctx, TC_ERR_NOMEM, "out of host memory"); }
#define INS_FREE() do { free(h_range); free(h_off); } while (0)
    cudaError_t e = cudaMemcpyAsync(h_range, d_range, 8 * (size_t)n_cand, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { INS_FREE(); return tc_cuda_fail(ctx, e, "range readback"); }
    h_off[0] = 0;
    for (int i = 0; i < n_cand; ++i) h_off[i + 1] = h_off[i] + (h_range[2 * i + 1] - h_range[2 * i]);
    const int64_t total = h_off[n_cand];
    if (total >= 0x7fffffffll) { INS_FREE(); return tc_fail(ctx, TC_ERR_CAPACITY, "too many candidate column entries (%lld)", (long long)total); }
    const size_t T = (size_t)(total > 0 ? total : 1);
    int64_t* d_off = (int64_t*)tc_dev_buf(ctx, SLOT_INS_C, 8 * ((size_t)n_cand + 1));
    // one slab: keys, sorted keys, ids, sorted ids, read, indel, qpos, head, counts
    size_t bytes = T * (8 + 8 + 4 + 4 + 4 + 4 + 4 + 1) + 4 * (size_t)n_cand + 64;
    uint8_t* slab = (uint8_t*)tc_dev_buf(ctx, SLOT_INS_D, bytes);
    tc_insert_call_t* d_calls = (tc_insert_call_t*)tc_dev_buf(ctx, SLOT_INS_E, sizeof(tc_insert_call_t) * (size_t)n_cand);
    if (!d_off || !slab || !d_calls) { INS_FREE(); return TC_ERR_NOMEM; }
    uint64_t* d_key = (uint64_t*)slab; uint64_t* d_skey = d_key + T;
    uint32_t* d_idx = (uint32_t*)(d_skey + T); uint32_t* d_sidx = d_idx + T;
    a.ent_read = d_sidx + T; a.ent_indel = (int32_t*)(a.ent_read + T); a.ent_qpos = a.ent_indel + T;
    a.seg_count = a.ent_qpos + T; a.ent_head = (uint8_t*)(a.seg_count + n_cand);
    a.ent_key = d_key; a.ent_idx = d_idx; a.seg_off = d_off;
    e = cudaMemcpyAsync(d_off, h_off, 8 * ((size_t)n_cand + 1), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { INS_FREE(); return tc_cuda_fail(ctx, e, "offset upload"); }
    ins_select_emit_kernel<<<n_cand, 1024, 0, s>>>(a);
    ctx->launches++;
    // segments for the sort: [off[i], off[i] + count[i])
    int64_t* d_seg_end = (int64_t*)tc_dev_buf(ctx, SLOT_INS_F, 8 * (size_t)n_cand);
    if (!d_seg_end) { INS_FREE(); return TC_ERR_NOMEM; }
    seg_end_kernel<<<(n_cand + 127) / 128, 128, 0, s>>>(d_off, a.seg_count, d_seg_end, n_cand);
    ctx->launches++;
    size_t tmp_bytes = 0;
    e = cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_skey, d_idx, d_sidx, (int64_t)total, n_cand, d_off, d_seg_end, 0, 64, s);
    if (e != cudaSuccess) { INS_FREE(); return tc_cuda_fail(ctx, e, "cub sort sizing"); }
    void* d_tmp = tc_dev_buf(ctx, SLOT_INS_G, tmp_bytes + 16);
    if (!d_tmp) { INS_FREE(); return TC_ERR_NOMEM; }
    e = cub::DeviceSegmentedRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_skey, d_idx, d_sidx, (int64_t)total, n_cand, d_off, d_seg_end, 0, 64, s);
    if (e != cudaSuccess) { INS_FREE(); return tc_cuda_fail(ctx, e, "cub segmented radix sort"); }
    ctx->launches++;
    ins_mode_kernel<<<n_cand, 1024, 0, s>>>(a, d_skey, d_sidx, d_calls);
    ctx->launches++;
    e = cudaMemcpyAsync(calls, d_calls, sizeof(tc_insert_call_t) * (size_t)n_cand, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->host_status, d_status, sizeof(tc_status), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { INS_FREE(); return tc_cuda_fail(ctx, e, "insert calls readback"); }
    tc_status st; memcpy(&st, ctx->host_status, sizeof(st));
    if (st.err == TC_ERR_UNSORTED) { INS_FREE(); return tc_fail(ctx, TC_ERR_UNSORTED, "Unsorted input. Pileup aborts"); }
    if (st.err) { INS_FREE(); return tc_fail(ctx, st.err, "insertion key collision or device-side failure %d", st.err); }
    // lay the winners' inserted characters out in the caller's buffer
    int64_t need = 0;
    int64_t* h_entry = (int64_t*)malloc(8 * (size_t)n_cand);
    if (!h_entry) { INS_FREE(); return tc_fail(ctx, TC_ERR_NOMEM, "out of host memory"); }
    for (int i = 0; i < n_cand; ++i) {
        h_entry[i] = calls[i].bases_off;
        if (calls[i].indel > 0) { calls[i].bases_off = need; need += calls[i].indel; } else calls[i].bases_off = -1;
    }
    rc = TC_OK;
    if (need > 0) {
        if (!bases || need > bases_cap) rc = tc_fail(ctx, TC_ERR_CAPACITY, "bases buffer too small: need %lld bytes", (long long)need);
        else {
            uint8_t* d_bases = (uint8_t*)tc_dev_buf(ctx, SLOT_TMP_A, (size_t)need);
            int64_t* d_entry = (int64_t*)tc_dev_buf(ctx, SLOT_TMP_B, 8 * (size_t)n_cand);
            if (!d_bases || !d_entry) rc = TC_ERR_NOMEM;
            else {
                e = cudaMemcpyAsync(d_entry, h_entry, 8 * (size_t)n_cand, cudaMemcpyHostToDevice, s);
                if (e == cudaSuccess) e = cudaMemcpyAsync(d_calls, calls, sizeof(tc_insert_call_t) * (size_t)n_cand, cudaMemcpyHostToDevice, s);
                if (e == cudaSuccess) {
                    ins_bases_kernel<<<n_cand, 128, 0, s>>>(a, d_calls, d_entry, d_bases);
                    ctx->launches++;
                    e = cudaMemcpyAsync(bases, d_bases, (size_t)need, cudaMemcpyDeviceToHost, s);
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(s);
                if (e != cudaSuccess) rc = tc_cuda_fail(ctx, e, "inserted bases readback");
            }
        }
    }
    free(h_entry);
    INS_FREE();
#undef INS_FREE
    if (rc == TC_OK) { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) rc = tc_cuda_fail(ctx, le, "insert kernels"); }
    return rc;
}
