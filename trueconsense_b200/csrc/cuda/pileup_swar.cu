// pileup_swar.cu — variant 2 of the pileup kernel (placeholder until the SWAR kernel lands).
#include "pileup.cuh"

int tc_pileup_swar_launch(tc_ctx* ctx, const pileup_args& a, cudaStream_t s) {
    (void)a; (void)s;
    return tc_fail(ctx, TC_ERR_ARG, "pileup kernel variant 2 is not built into this library");
}
