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

TC_API int tc_allreduce_counts(tc_ctx_t* ctx, int32_t* counts_dev, int64_t n_elems, void* comm, void* stream) {
    if (!ctx) return TC_ERR_ARG;
    if (!counts_dev || n_elems < 0 || !comm) return tc_fail(ctx, TC_ERR_ARG, "bad argument");
    if (!tc_is_device_ptr(counts_dev)) return tc_fail(ctx, TC_ERR_ARG, "tc_allreduce_counts needs a device pointer");
    TC_CUDA(cudaSetDevice(ctx->device));
    if (!g_allreduce) {
        const char* path = getenv("TC_NCCL_LIB");
        void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return tc_fail(ctx, TC_ERR_ARG, "cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
        g_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
        g_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
        if (!g_allreduce) return tc_fail(ctx, TC_ERR_ARG, "ncclAllReduce not found in the NCCL library");
    }
    const int nccl_int32 = 2, nccl_sum = 0;
    int r = g_allreduce(counts_dev, counts_dev, (size_t)n_elems, nccl_int32, nccl_sum, comm, (cudaStream_t)stream);
    if (r != 0) return tc_fail(ctx, TC_ERR_CUDA, "ncclAllReduce failed: %s", g_errstr ? g_errstr(r) : "?");
    return TC_OK;
}
