// CPU harness around trueconsense_b200/csrc/cuda/inflate_core.cuh (the same source the GPU kernels compile): lets the
// CPU test-suite fuzz the DEFLATE decoder and the CRC pieces against zlib.  Test infrastructure only.
#include <stdlib.h>
#include <string.h>

#include "inflate_core.cuh"

using namespace tcinf;

extern "C" int h_inflate(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out, int stride) {
    uint16_t* lut = (uint16_t*)calloc((size_t)LUT_SIZE * stride, 2);
    uint16_t* dlut = (uint16_t*)calloc((size_t)DLUT_SIZE * stride, 2);
    huff hl, hd;
    uint8_t lens[LENS_SIZE];
    (void)stride;
    const int rc = inflate_block<1>(in, n_in, out, n_out, lut, dlut, hl, hd, lens, 0);
    free(lut); free(dlut);
    return rc;
}

extern "C" uint32_t h_crc(const uint8_t* p, int64_t n) {
    uint32_t tab[256];
    for (uint32_t i = 0; i < 256; ++i) tab[i] = crc_table_entry(i);
    uint32_t c = 0xffffffffu;
    for (int64_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xffffffffu;
}

extern "C" uint32_t h_crc_combine(uint32_t a, uint32_t b, uint64_t len_b) { return crc_combine(a, b, len_b); }
