// allreduce.cu — the one collective of the path: summing per-rank count tables when an
// ultra-deep single sample is sharded by read range.  Counts are additive int32,
// so the result is bit-identical to the single-GPU table.  The table is ~1 MB (L = 29,903) to
// ~6 MB (L = 197 k): latency-bound on NVLink/NVSwitch, one ncclAllReduce.
//
// NCCL is resolved at run time from the library the host layer created the communicator with
// (TC_NCCL_LIB, default "libnccl.so.2"), so libtcb200.so itself has no link-time NCCL dependency.
#include <dlfcn.h>
#include <stdlib.h>

#include "tc_common.cuh"

typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);

static nccl_allreduce_fn g_allreduce = nullptr;
static nccl_errstr_fn g_errstr = nullptr;

static int nccl_resolve(tc_ctx* ctx) {
    if (!g_allreduce) {
        const char* path = getenv("TC_NCCL_LIB");
        void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return tc_fail(ctx, TC_ERR_ARG, "cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        g_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
        g_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
        if (!g_allreduce) return tc_fail(ctx, TC_ERR_ARG, "ncclAllReduce not found in the NCCL library");
    }
    return TC_OK;
}

TC_API int tc_allreduce_counts(tc_ctx_t* ctx, int32_t* counts_dev, int64_t n_elems, void* comm, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!counts_dev || n_elems < 0 || !comm) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    if (!tc_is_device_ptr(counts_dev)) return tc_fail(ctx, TC_ERR_ARG, "tc_allreduce_counts needs a device pointer");
    TC_CUDA(cudaSetDevice(ctx->device));
    int rc = nccl_resolve(ctx);
    if (rc) return rc;
    const int nccl_int32 = 2, nccl_sum = 0;
    int r = g_allreduce(counts_dev, counts_dev, (size_t)n_elems, nccl_int32, nccl_sum, comm, (cudaStream_t)stream);
    if (r != 0) return tc_fail(ctx, TC_ERR_CUDA, "ncclAllReduce failed: %s", g_errstr ? g_errstr(r) : "?");
    return TC_OK;
}

// ---- the shard's pileup and the sum in one enqueue -------------------------------------------------------------------
// The pad row of the table (row 7, all zero by contract) carries one word of the collective: "this rank's status block
// holds an error".  Summed with the counts, every rank learns whether ANY shard needs the slow road, without a second
// collective and without a synchronisation between the pileup and the sum.
__global__ void rr_flag_kernel(const tc_status* st, int32_t* pad0) { *pad0 = st->err != 0 ? 1 : 0; }

struct rr_host { tc_status st; int32_t any_err; };

static int rr_chain(tc_ctx* ctx, const tc_reads_t* reads, int32_t L, const tc_pileup_params_t* p, int32_t* counts, void* comm,
                    cudaStream_t s, tc_pileup_pending* pend) {
    int rc = tc_pileup_enqueue(ctx, reads, L, p, counts, s, pend);
    if (rc) return rc;
    int32_t* pad0 = counts + (size_t)(TC_NROWS - 1) * L;
    rr_flag_kernel<<<1, 1, 0, s>>>(pend->d_status, pad0);
    TC_LAUNCH_CHECK();
    const int nccl_int32 = 2, nccl_sum = 0;
    // rows 0..6 and the flag word: the rest of the pad row stays out of the collective
    const int r = g_allreduce(counts, counts, (size_t)(TC_NROWS - 1) * L + 1, nccl_int32, nccl_sum, comm, s);
    if (r != 0) return tc_fail(ctx, TC_ERR_CUDA, "ncclAllReduce failed: %s", g_errstr ? g_errstr(r) : "?");
    ctx->launches++;
    rr_host* h = (rr_host*)ctx->host_status;
    TC_D2H(&h->st, pend->d_status, sizeof(tc_status), s);
    TC_D2H(&h->any_err, pad0, sizeof(int32_t), s);
    TC_CUDA(cudaMemsetAsync(pad0, 0, sizeof(int32_t), s));
    return TC_OK;
}

TC_API int tc_pileup_counts_allreduce(tc_ctx_t* ctx, const tc_reads_t* reads, int32_t ref_len, const tc_pileup_params_t* p,
                                      int32_t* counts, void* comm, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!reads || !p || !counts || !comm || ref_len <= 0) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    if (!tc_is_device_ptr(counts)) return tc_fail(ctx, TC_ERR_ARG, "tc_pileup_counts_allreduce needs a device pointer for the table");
    static_assert(sizeof(rr_host) <= 256, "host_status block");
    TC_CUDA(cudaSetDevice(ctx->device));
    int rc = nccl_resolve(ctx);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int L = ref_len;

    unsigned char key[256];
    int kn = 0;
    auto put = [&](const void* q, size_t bytes) { memcpy(key + kn, q, bytes); kn += (int)bytes; };
    put(reads, sizeof(*reads)); put(&ref_len, 4); put(p, sizeof(*p)); put(&counts, sizeof(counts)); put(&comm, sizeof(comm));
    put(&ctx->buf_epoch, 8); put(&ctx->timing, 4);
    static_assert(sizeof(tc_reads_t) + 4 + sizeof(tc_pileup_params_t) + 16 + 12 <= 256, "shard key");
    const bool same = ctx->rr_key_len == kn && memcmp(ctx->rr_key, key, (size_t)kn) == 0;
    if (!same) {
        if (ctx->rr_exec) { cudaGraphExecDestroy(ctx->rr_exec); ctx->rr_exec = nullptr; }
        memcpy(ctx->rr_key, key, (size_t)kn); ctx->rr_key_len = kn; ctx->rr_seen = 0;
    }
    tc_pileup_pending& pend = ctx->rr_pend;        // of the run that was enqueued eagerly or captured: a replay is that run again
    if (same && ctx->rr_exec) {
        TC_CUDA(cudaGraphLaunch(ctx->rr_exec, s));
        ctx->launches += ctx->rr_launches; ctx->d2h_bytes += ctx->rr_d2h;
    } else if (same && ctx->rr_seen >= 1 && tc_reads_all_device(reads) && !getenv("TC_NO_GRAPH")) {
        if (!ctx->cap_stream) TC_CUDA(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
        const int64_t l0 = ctx->launches, d0 = ctx->d2h_bytes, epoch0 = ctx->buf_epoch;
        TC_CUDA(cudaStreamBeginCapture(ctx->cap_stream, cudaStreamCaptureModeRelaxed));
        ctx->in_capture = 1;
        rc = rr_chain(ctx, reads, L, p, counts, comm, ctx->cap_stream, &pend);
        ctx->in_capture = 0;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ctx->cap_stream, &graph);
        if (rc == TC_OK && ce == cudaSuccess && graph && ctx->buf_epoch == epoch0) {
            const cudaError_t ie = cudaGraphInstantiate(&ctx->rr_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { ctx->rr_exec = nullptr; return tc_cuda_fail(ctx, ie, "cudaGraphInstantiate"); }
            ctx->rr_launches = ctx->launches - l0; ctx->rr_d2h = ctx->d2h_bytes - d0;
            TC_CUDA(cudaGraphLaunch(ctx->rr_exec, s));
        } else {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            ctx->launches = l0; ctx->d2h_bytes = d0;
            if (rc && rc != TC_ERR_CUDA) return rc;
            ctx->rr_key_len = 0;
            rc = rr_chain(ctx, reads, L, p, counts, comm, s, &pend);
            if (rc) return rc;
        }
    } else {
        rc = rr_chain(ctx, reads, L, p, counts, comm, s, &pend);
        if (rc) return rc;
        ctx->rr_seen++;
    }
    TC_CUDA(cudaStreamSynchronize(s));              // the one synchronisation of the pass
    rr_host h;
    memcpy(&h, ctx->host_status, sizeof(h));
    if (h.any_err != 0) {
        // some rank's shard did not go through: every rank takes the separate calls (which handle the kernel variants) and
        // makes the same collective call whatever its own outcome, so nobody waits for a rank that failed
        ctx->rr_key_len = 0;
        const int prc = tc_pileup_counts(ctx, reads, ref_len, p, counts, stream);
        if (prc) TC_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * TC_NROWS * (size_t)L, s));
        char keep[sizeof(ctx->err)];
        if (prc) memcpy(keep, ctx->err, sizeof(keep));
        const int arc = tc_allreduce_counts(ctx, counts, (int64_t)TC_NROWS * L, comm, stream);
        if (prc) { memcpy(ctx->err, keep, sizeof(keep)); return prc; }
        return arc;
    }
    return tc_pileup_finish(ctx, h.st, &pend, reads, ref_len, p, counts, stream);
}
