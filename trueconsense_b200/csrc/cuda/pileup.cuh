// pileup.cuh — arguments shared by the pileup kernel variants.
#pragma once
#include "tc_common.cuh"

struct pileup_args {
    dreads r;
    int32_t L;
    uint32_t flag_filter;
    int32_t min_mapq, min_bq, ignore_orphans;
    int32_t* counts;        // [8][L]
    int32_t* diff;          // [L+1]
    tc_status* status;
    int32_t span_hint;      // tc_reads_t.max_ref_span (0 = unknown: a span pass finds it)
};

__device__ __forceinline__ bool read_passes(const pileup_args& a, int64_t r) {
    uint32_t fl = a.r.flag[r];
    if (fl & (a.flag_filter | 4u)) return false;                 // htslib always drops UNMAP
    if (a.min_mapq > 0 && a.r.mapq && (int)a.r.mapq[r] < a.min_mapq) return false;
    if (a.ignore_orphans && (fl & 1u) && !(fl & 2u)) return false;
    return true;
}


// variant 2 (pileup_swar.cu): enqueue the SWAR column kernel(s); fills every row except coverage
int tc_pileup_swar_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s);
bool tc_pileup_swar_supported(const pileup_args& a);

// variant 3 (pileup_warp.cu): barrier-free warp-per-read-stream SWAR kernel; fills every row except coverage
int tc_pileup_warp_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s);
bool tc_pileup_warp_supported(const pileup_args& a);
