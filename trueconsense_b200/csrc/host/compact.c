/* compact.c — the compact transport forms of a read batch (trueconsense_b200.h: tc_reads_t.seq2 / seq_exc_* / cigar16).
 *
 * What travels over PCIe per sample is SEQ (4 bits per base in BAM and in tc_reads_t.seq4) and the CIGARs; a viral amplicon
 * sample is A / C / G / T except for the odd N, and its operations are short.  A decoder has every base in hand once: packing
 * two bits per base next to the 4-bit words costs it one more pass (here: OpenMP over reads), and halves the bytes every
 * later upload of the batch moves.  Words with anything else than A C G T over their valid bases go to an exception list as they
 * are, so the device rebuilds seq4 bit for bit. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "tc_host.h"

/* nibble (one-hot BAM code) -> 2-bit code, 0xff: not A C G T */
static const uint8_t CODE2[16] = {0xff, 0, 1, 0xff, 2, 0xff, 0xff, 0xff, 3, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff};

int tc_seq2_pack(const uint32_t* seq4, int64_t n_seq_words, const uint32_t* seq_off, const int32_t* l_seq, int64_t n_reads,
                 uint16_t* seq2, uint32_t** exc_idx_out, uint32_t** exc_val_out, int64_t* n_exc_out, int n_threads) {
    if (!seq2 || !exc_idx_out || !exc_val_out || !n_exc_out || n_reads < 0 || n_seq_words < 0) return -1;
    *exc_idx_out = NULL; *exc_val_out = NULL; *n_exc_out = 0;
    if (n_reads == 0 || n_seq_words == 0) return 0;
    if (!seq4 || !seq_off || !l_seq) return -1;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_num_procs();
#else
    n_threads = 1;
#endif
    uint8_t* flag = calloc((size_t)n_seq_words, 1);
    if (!flag) return -5;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(n_threads) reduction(|:bad)
    for (int64_t r = 0; r < n_reads; ++r) {
        const uint32_t w0 = seq_off[r], w1 = seq_off[r + 1];
        if (w1 < w0 || (int64_t)w1 > n_seq_words) { bad = 1; continue; }
        const int64_t len = l_seq[r];
        for (uint32_t w = w0; w < w1; ++w) {
            const uint32_t v = seq4[w];
            const int64_t valid = len - 8 * (int64_t)(w - w0);
            uint32_t h = 0;
            int exc = 0;
            for (int j = 0; j < 8; ++j) {
                const uint32_t nib = (v >> (8 * (j >> 1) + ((j & 1) ? 0 : 4))) & 15u;
                if (j < valid) {
                    const uint8_t c = CODE2[nib];
                    if (c == 0xff) exc = 1; else h |= (uint32_t)c << (2 * j);
                } else if (nib) exc = 1;            /* padding that is not zero: kept as it is */
            }
            seq2[w] = (uint16_t)h;
            flag[w] = (uint8_t)exc;
        }
    }
    if (bad) { free(flag); return -2; }
    int64_t n_exc = 0;
    for (int64_t w = 0; w < n_seq_words; ++w) n_exc += flag[w];
    if (n_exc > 0) {
        uint32_t* idx = malloc(sizeof(uint32_t) * (size_t)n_exc);
        uint32_t* val = malloc(sizeof(uint32_t) * (size_t)n_exc);
        if (!idx || !val) { free(idx); free(val); free(flag); return -5; }
        int64_t k = 0;
        for (int64_t w = 0; w < n_seq_words; ++w)
            if (flag[w]) { idx[k] = (uint32_t)w; val[k] = seq4[w]; ++k; }
        *exc_idx_out = idx; *exc_val_out = val;
    }
    *n_exc_out = n_exc;
    free(flag);
    return 0;
}

void tc_host_free(void* p) { free(p); }
