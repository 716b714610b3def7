// pileup_long.cu — long reads for pileup variant 3.
//
// The warp kernel (pileup_warp.cu) keeps a read's whole alignment inside one 512- or 1024-column window, so a
// read may span at most ~950 reference columns.  Longer reads (10-kb long reads, BASELINE config 5) are cut
// into PIECES here: the reference is divided into cells of PIECE_COLS columns and the part of an alignment
// inside one cell becomes a piece that looks like a short read —
//     position, the SEQ word holding its first query base (+ the base's index inside that word), the query
//     length left from there, and its own CIGAR: the read's ops over the cell, the op crossing a cell border
//     split in two.  Ops that consume no reference (I, S, H) stay with the piece of the reference-consuming op in
//     front of them, so "an insertion counts on the last column before it" never looks across a cut.
// SEQ is not copied: pieces point into the read's own packed bases.  Pieces are then ordered by start column
// (cub radix sort) and the same kernel runs over them in its PIECES mode (records gathered through the order,
// every piece staged on its own).  Coverage, span statistics and the sort / range checks come from the span
// pass over the real reads, as always.
//
//   lr_count_kernel   one thread per read: pieces and piece-CIGAR ops the read will produce (from its span)
//   exclusive sums    cub::DeviceScan -> where each read's pieces / ops go
//   lr_split_kernel   one thread per read: CIGAR walk, writes piece records and piece CIGARs
//   radix sort        piece start columns -> processing order
//   warp_pileup_kernel<64, PIECES>
// Reads with pads, zero-length ops or no SEQ are not cut (TC_ERR_CAPACITY -> the scatter kernel takes the batch).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "pileup.cuh"

namespace {

__device__ __forceinline__ int ref_span_of(const pileup_args& a, int64_t r) {
    int span = 0;
    for (uint32_t k = a.r.cigar_off[r]; k < a.r.cigar_off[r + 1]; ++k) {
        const uint32_t c = a.r.cigar[k];
        if (op_consumes_ref(c & 15u)) span += (int)(c >> 4);
    }
    return span;
}

// pieces per read = cells its span touches; ops per read <= its own ops + one extra per cut, padded to whole
// 16-byte vectors per piece (+3 ops each) so that a piece's ops can be staged as vectors
__global__ void lr_count_kernel(pileup_args a, uint32_t* __restrict__ n_pieces, uint32_t* __restrict__ n_ops) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.r.n) return;
    uint32_t np = 0, no = 0;
    const int pos = a.r.pos[r];
    if (read_passes(a, r) && pos >= 0 && pos < a.L) {
        const int span = ref_span_of(a, r);
        if (span > 0) {
            np = (uint32_t)((pos + span - 1) / PIECE_COLS - pos / PIECE_COLS + 1);
            no = (a.r.cigar_off[r + 1] - a.r.cigar_off[r]) + np * 8u;      // + cut ops, vector padding, alignment slack
        }
    }
    n_pieces[r] = np;
    n_ops[r] = (no + 3u) & ~3u;
}

// One thread per read walks its CIGAR, but the ops come through shared memory: per round the warp loads the next 32
// ops of each of its 32 reads with one coalesced 128-byte request per read (all 32 in flight together), then every lane
// consumes its own read's 32 ops.  (A thread streaming its own 2.4 KB of ops from HBM alone took 485 us on config 5.)
constexpr int LR_SPLIT_THREADS = 128;
__global__ void __launch_bounds__(LR_SPLIT_THREADS) lr_split_kernel(pileup_args a, const uint32_t* __restrict__ piece_base, const uint32_t* __restrict__ ops_base,
                                tc_piece* __restrict__ pieces, uint32_t* __restrict__ pcig, int32_t* __restrict__ piece_pos) {
    __shared__ uint32_t stage[LR_SPLIT_THREADS / 32][32][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = r < a.r.n && piece_base[r + 1] != piece_base[r];
    const int lq = valid ? a.r.l_seq[r] : 0;
    if (valid && lq == 0) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); valid = false; }        // SEQ '*': scatter kernel only
    const uint32_t sbeg = valid ? a.r.seq_off[r] : 0u;
    uint32_t pi = valid ? piece_base[r] : 0u;            // next piece record
    uint32_t w = valid ? ops_base[r] : 0u;               // next op slot (multiple of 4)
    int x = valid ? a.r.pos[r] : 0, y = 0;
    int cell_end = (x / PIECE_COLS + 1) * PIECE_COLS;
    // the open piece
    int px = x, py = 0;
    uint32_t pw = w;
    bool open = false, has_ref = false;
    auto close_piece = [&]() {
        tc_piece pc;
        const int yw = py >> 3;
        pc.pos = px; pc.seq_beg = sbeg + (uint32_t)yw; pc.cig_beg = pw; pc.cig_n = w - pw;
        pc.seq_n = (uint32_t)((y - 8 * yw + 7) / 8 + 1);
        pc.lq = lq - 8 * yw; pc.y0 = py & 7; pc.pad = 0;
        pieces[pi] = pc;
        piece_pos[pi] = px;
        ++pi;
        while (w & 3u) pcig[w++] = 0x10u | OP_H;            // pad to a whole vector with no-op hard clips (never read: cig_n stops before)
    };
    uint32_t k = valid ? a.r.cigar_off[r] : 0u;
    uint32_t k1 = valid ? a.r.cigar_off[r + 1] : 0u;
    while (__any_sync(0xffffffffu, k < k1)) {
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const uint32_t ki = __shfl_sync(0xffffffffu, k, i), k1i = __shfl_sync(0xffffffffu, k1, i);
            if (ki + lane < k1i) stage[wib][i][lane] = __ldg(a.r.cigar + ki + lane);
        }
        __syncwarp();
        const int nb = (int)min(32u, k1 - k);
        for (int t = 0; t < nb; ++t) {
            const uint32_t c = stage[wib][lane][t];
            const uint32_t op = c & 15u;
            int l = (int)(c >> 4);
            if (op == OP_P || l == 0 || op > OP_X) { atomicCAS(&a.status->err, 0, TC_ERR_CAPACITY); k1 = k; valid = false; break; }      // scatter kernel only
            if (!op_consumes_ref(op)) {
                // clips in front of the first aligned base are dropped (a soft clip only moves the query index): pieces
                // never carry an over-long leading clip into the kernel's 16-bit op staging
                if (!open && (op == OP_H || op == OP_S)) { if (op == OP_S) y += l; continue; }
                if (!open) { open = true; px = x; py = y; pw = w; }     // a leading insertion opens the first piece
                pcig[w++] = c;
                if (op == OP_I || op == OP_S) y += l;
                continue;
            }
            const bool match = op_is_match(op);
            while (l > 0) {
                if (x == cell_end) {            // the next column belongs to the next cell: cut here
                    if (open && has_ref) { close_piece(); open = false; has_ref = false; }
                    cell_end += PIECE_COLS;
                }
                if (!open) { open = true; px = x; py = y; pw = w; }
                const int take = min(l, cell_end - x);
                pcig[w++] = ((uint32_t)take << 4) | op;
                has_ref = true;
                x += take; l -= take;
                if (match) y += take;
            }
        }
        if (valid) k += (uint32_t)nb;
        __syncwarp();
    }
    if (valid && open && has_ref) close_piece();
}

__global__ void lr_iota_kernel(uint32_t* __restrict__ idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

}  // namespace

int tc_pileup_long_launch(tc_ctx* ctx, const pileup_args& a0, cudaStream_t s) {
    pileup_args a = a0;
    const int64_t n = a.r.n;
    // per read: pieces and ops, then exclusive sums (n + 1 entries so that [r + 1] - [r] is the count)
    uint32_t* cnt = (uint32_t*)tc_dev_buf(ctx, SLOT_SEGS, 4 * 4 * ((size_t)n + 1) + 64);
    if (!cnt) return TC_ERR_NOMEM;
    uint32_t* n_pieces = cnt; uint32_t* n_ops = cnt + (n + 1); uint32_t* piece_base = cnt + 2 * (n + 1); uint32_t* ops_base = cnt + 3 * (n + 1);
    TC_CUDA(cudaMemsetAsync(cnt, 0, 4 * 2 * ((size_t)n + 1), s));
    lr_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, n_pieces, n_ops);
    TC_LAUNCH_CHECK();
    size_t tmp_bytes = 0, tmp2 = 0;
    TC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, n_pieces, piece_base, (int)(n + 1), s));
    void* d_tmp = tc_dev_buf(ctx, SLOT_TMP_C, tmp_bytes + 16);
    if (!d_tmp) return TC_ERR_NOMEM;
    TC_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, n_pieces, piece_base, (int)(n + 1), s));
    TC_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, n_ops, ops_base, (int)(n + 1), s));
    ctx->launches += 2;
    uint32_t totals[2];
    TC_CUDA(cudaMemcpyAsync(&totals[0], piece_base + n, 4, cudaMemcpyDeviceToHost, s));
    TC_CUDA(cudaMemcpyAsync(&totals[1], ops_base + n, 4, cudaMemcpyDeviceToHost, s));
    TC_CUDA(cudaStreamSynchronize(s));
    ctx->d2h_bytes += 8;
    const size_t NP = totals[0], NO = totals[1];
    if (NP == 0) return TC_OK;
    if (NP >= 0x7fffffffull) return tc_fail(ctx, TC_ERR_CAPACITY, "too many pieces (%zu)", NP);
    // pieces, their start columns (sort keys), the order, the piece CIGARs
    uint8_t* slab = (uint8_t*)tc_dev_buf(ctx, SLOT_TILES, NP * (sizeof(tc_piece) + 4 * 4) + 4 * (NO + 8) + 256);
    if (!slab) return TC_ERR_NOMEM;
    tc_piece* pieces = (tc_piece*)slab;
    int32_t* key_in = (int32_t*)(pieces + NP); int32_t* key_out = key_in + NP;
    uint32_t* idx_in = (uint32_t*)(key_out + NP); uint32_t* idx_out = idx_in + NP;
    uint32_t* pcig = idx_out + NP;
    pcig = (uint32_t*)(((uintptr_t)pcig + 15) & ~(uintptr_t)15);
    lr_split_kernel<<<(unsigned)((n + LR_SPLIT_THREADS - 1) / LR_SPLIT_THREADS), LR_SPLIT_THREADS, 0, s>>>(a, piece_base, ops_base, pieces, pcig, key_in);
    TC_LAUNCH_CHECK();
    lr_iota_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, s>>>(idx_in, (int64_t)NP);
    TC_LAUNCH_CHECK();
    int bits = 1;
    while ((1ll << bits) < (long long)a.L + 1 && bits < 31) ++bits;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, key_in, key_out, idx_in, idx_out, (int)NP, 0, bits, s));
    void* d_tmp2 = tc_dev_buf(ctx, SLOT_TMP_C, (tmp2 > tmp_bytes ? tmp2 : tmp_bytes) + 16);
    if (!d_tmp2) return TC_ERR_NOMEM;
    TC_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp2, tmp2, key_in, key_out, idx_in, idx_out, (int)NP, 0, bits, s));
    ctx->launches++;
    a.pieces = pieces; a.piece_order = idx_out; a.n_pieces = (int64_t)NP;
    a.r.cigar = pcig;
    a.span_hint = 0;
    return tc_pileup_warp_launch_pieces(ctx, a, s);
}
