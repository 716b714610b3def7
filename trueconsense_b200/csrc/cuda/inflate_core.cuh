// inflate_core.cuh — raw DEFLATE (RFC 1951) of ONE BGZF block by ONE thread, and CRC-32 pieces.
//
// The reference reads its BAM through pysam -> htslib (TrueConsense/indexing.py:6-19 `Readbam`, :96 the pileup's reader):
// bgzf.c inflates every <= 64 KiB BGZF member with zlib and checks its CRC-32.  A BAM of one deep sample is tens of
// thousands of such members, each independent of the others — the one axis a GPU can use: a thread inflates a whole
// member sequentially (Huffman decoding is a serial bit-stream walk), thousands of members are in flight at once.
//
// Plain functions, `__host__ __device__`: the CPU test-suite compiles this header with g++ and fuzzes it against zlib
// (tests/test_inflate_core.py); the kernels in bgzf.cu instantiate it with the tables in shared memory.
//
// Decoding: canonical Huffman codes kept puff-style (count per length + symbols sorted by code), a first-level look-up
// table of LUT_BITS (literal/length) and DLUT_BITS (distance) bits in front — entry = symbol << 4 | code length, 0 =
// "longer than the table: decode bit by bit".  Length and distance bases are computed, not tabulated.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TCI_HD __host__ __device__ __forceinline__
#else
#define TCI_HD static inline
#endif

// NL lanes of a warp may run one member's decode TOGETHER (NL = 32 on the device, 1 on the host): every lane executes the same
// instructions on the same values — the decode is one serial chain either way, a warp instruction costs one issue slot whether
// one lane or all of them are active — and the lanes share what is data-parallel: a match is copied NL bytes per step, a table is
// filled NL entries per step.  Stores of single bytes are lane 0's.
#if defined(__CUDA_ARCH__)
#define TCI_WARP_SYNC() __syncwarp()
#else
#define TCI_WARP_SYNC() ((void)0)
#endif

namespace tcinf {

constexpr int LUT_BITS = 10;
constexpr int DLUT_BITS = 7;
constexpr int LUT_SIZE = 1 << LUT_BITS;
constexpr int DLUT_SIZE = 1 << DLUT_BITS;
constexpr int LENS_SIZE = 344;          // 19 code-length code lengths + up to 286 + 30 code lengths

enum {
    INF_OK = 0,
    INF_ERR_BTYPE = 1,          // reserved block type
    INF_ERR_STORED = 2,         // stored block: LEN / NLEN mismatch
    INF_ERR_HEADER = 3,         // dynamic header: bad counts, bad code lengths, no end-of-block code
    INF_ERR_CODE = 4,           // a bit pattern that is no code
    INF_ERR_DIST = 5,           // distance symbol 30 / 31, or a distance reaching in front of the output
    INF_ERR_OUTPUT = 6,         // more (or fewer) bytes than the member's ISIZE
    INF_ERR_INPUT = 7,          // the stream runs past the member's compressed bytes
};

struct huff {
    uint16_t count[16];         // codes of each length
    uint16_t symbol[288];       // symbols ordered by (length, symbol)
};

// LSB-first bit reader.  The stream is fetched in aligned 16-byte vectors, one vector AHEAD of the one being consumed: a
// member's decode is a chain of dependent steps, and a load that has to come from L2 / DRAM (~700 cycles against ~90 for a
// symbol) in the middle of it was two thirds of the time — the vector asked for 16 symbols earlier has arrived when it is
// needed.  Vectors at and behind `n_vec` read as zero: a corrupt stream cannot walk out of the buffer (which must be
// readable from the 16-byte boundary in front of the stream to the one behind it); bits_over() tells whether more bits
// were CONSUMED than the stream holds.
struct alignas(16) vec4 { uint32_t x, y, z, w; };

struct bits {
    const vec4* base;
    int64_t k, n_vec;           // next vector to fetch, vectors covering the stream
    vec4 nxt;                   // fetched ahead
    uint64_t lo, hi;            // words of the current vector not yet taken
    int words_left;
    uint32_t taken_bits, lim_bits;   // (streams of less than 2^28 bytes)
    uint64_t buf;
    int n;
};

TCI_HD uint32_t bits_pop_word(bits& b) {
    if (b.words_left == 0) {
        b.lo = (uint64_t)b.nxt.x | ((uint64_t)b.nxt.y << 32);
        b.hi = (uint64_t)b.nxt.z | ((uint64_t)b.nxt.w << 32);
        b.words_left = 4;
        if (b.k < b.n_vec) b.nxt = b.base[b.k]; else b.nxt = vec4{0, 0, 0, 0};
        b.k++;
    }
    const uint32_t w = (uint32_t)b.lo;
    b.lo = (b.lo >> 32) | (b.hi << 32);
    b.hi >>= 32;
    b.words_left--;
    return w;
}
TCI_HD void bits_init(bits& b, const uint8_t* in, int64_t n_bytes) {
    const uintptr_t a = (uintptr_t)in & 15u;
    b.base = (const vec4*)(in - a);
    b.n_vec = ((int64_t)a + n_bytes + 15) / 16;
    b.nxt = b.n_vec > 0 ? b.base[0] : vec4{0, 0, 0, 0};
    b.k = 1;
    b.words_left = 0; b.lo = b.hi = 0;
    for (uintptr_t i = 0; i < a / 4; ++i) bits_pop_word(b);
    const int skip = (int)(a & 3u);
    b.buf = (uint64_t)(bits_pop_word(b) >> (8 * skip));
    b.n = 32 - 8 * skip;
    b.taken_bits = 32;
    b.lim_bits = 8u * ((uint32_t)skip + (uint32_t)n_bytes);
}
TCI_HD void bits_refill(bits& b) {          // afterwards: at least 33 bits
    if (b.n <= 32) {
        b.buf |= (uint64_t)bits_pop_word(b) << b.n;
        b.n += 32;
        b.taken_bits += 32;
    }
}
TCI_HD uint32_t bits_take(bits& b, int n) {  // n <= 32, after a refill
    const uint32_t v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n; b.n -= n;
    return v;
}
TCI_HD bool bits_over(const bits& b) { return b.taken_bits - (uint32_t)b.n > b.lim_bits; }

// canonical code from code lengths (puff.c's construct): < 0 over-subscribed, 0 complete, > 0 incomplete
TCI_HD int huff_build(huff& h, const uint8_t* lengths, int n) {
    for (int l = 0; l < 16; ++l) h.count[l] = 0;
    for (int s = 0; s < n; ++s) h.count[lengths[s]]++;
    if (h.count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l < 16; ++l) {
        left <<= 1;
        left -= h.count[l];
        if (left < 0) return left;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
    for (int s = 0; s < n; ++s)
        if (lengths[s]) h.symbol[offs[lengths[s]]++] = (uint16_t)s;
    return left;
}

// first-level table.  LIT: entries of literals (symbols < 256) carry bit 15, so the hot loop tests one bit
constexpr uint32_t LIT_FLAG = 0x8000u;
template <int BITS, bool LIT, int NL>
TCI_HD void lut_build(const huff& h, uint16_t* lut, int lane) {
    if (NL > 1) TCI_WARP_SYNC();                                // nobody still reads the previous block's table
    for (int i = lane; i < (1 << BITS); i += NL) lut[i] = 0;
    if (NL > 1) TCI_WARP_SYNC();
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= BITS; ++l) {
        for (int j = 0; j < h.count[l]; ++j, ++idx, ++code) {
            uint32_t r = 0;                                     // the code arrives MSB first in an LSB-first stream
            for (int k = 0; k < l; ++k) r |= ((code >> k) & 1u) << (l - 1 - k);
            const uint32_t sym = h.symbol[idx];
            const uint16_t e = (uint16_t)((sym << 4) | (uint32_t)l | ((LIT && sym < 256u) ? LIT_FLAG : 0u));
            for (uint32_t v = r + ((uint32_t)lane << l); v < (1u << BITS); v += (uint32_t)NL << l) lut[v] = e;
        }
        code <<= 1;
    }
    if (NL > 1) TCI_WARP_SYNC();
}

// bit by bit (puff.c's decode): any code length
TCI_HD int huff_decode_slow(bits& b, const huff& h) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; ++l) {
        code |= (int)(b.buf & 1u);
        b.buf >>= 1; b.n -= 1;
        const int count = h.count[l];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

template <int BITS>
TCI_HD int huff_decode(bits& b, const huff& h, const uint16_t* lut) {
    const uint32_t e = lut[(uint32_t)(b.buf & ((1u << BITS) - 1u))];
    if (e) { b.buf >>= (e & 15u); b.n -= (int)(e & 15u); return (int)((e & ~LIT_FLAG) >> 4); }
    return huff_decode_slow(b, h);
}

// Inflate `in[0 .. n_in)` (a raw DEFLATE stream: what sits between a BGZF member's header and its CRC) into
// `out[0 .. n_out)`, n_out = the member's ISIZE (< 2^31).  `lut` / `dlut`: LUT_SIZE / DLUT_SIZE entries of this thread.
// `lens`, `hl`, `hd`: per-thread scratch.  Returns INF_OK or the first error; never reads outside the
// words covering the input, never writes outside `out`.
template <int NL>
TCI_HD int inflate_block(const uint8_t* in, int64_t n_in, uint8_t* out, int64_t n_out_, uint16_t* lut, uint16_t* dlut,
                         huff& hl, huff& hd, uint8_t* lens /* [LENS_SIZE] */, int lane) {
    const bool writer = NL == 1 || lane == 0;
    if (n_out_ < 0 || n_out_ > 0x7fffffff) return INF_ERR_OUTPUT;
    if (n_in < 0 || n_in >= (1 << 28)) return INF_ERR_INPUT;
    const uint32_t n_out = (uint32_t)n_out_;
    bits b;
    bits_init(b, in, n_in);
    uint32_t o = 0;
    for (;;) {
        bits_refill(b);
        if (bits_over(b)) return INF_ERR_INPUT;
        const uint32_t bfinal = bits_take(b, 1);
        const uint32_t btype = bits_take(b, 2);
        if (btype == 0) {
            bits_take(b, b.n & 7);                              // to the next byte boundary
            bits_refill(b);
            const uint32_t len = bits_take(b, 16);
            bits_refill(b);
            const uint32_t nlen = bits_take(b, 16);
            if ((len ^ 0xffffu) != nlen) return INF_ERR_STORED;
            if (len > n_out - o) return INF_ERR_OUTPUT;
            for (uint32_t i = 0; i < len; ++i) {
                bits_refill(b);
                const uint8_t v8 = (uint8_t)bits_take(b, 8);
                if (writer) out[o] = v8;
                ++o;
            }
            if (bits_over(b)) return INF_ERR_INPUT;
        } else if (btype == 1 || btype == 2) {
            if (btype == 1) {
                for (int s = 0; s < 144; ++s) lens[s] = 8;
                for (int s = 144; s < 256; ++s) lens[s] = 9;
                for (int s = 256; s < 280; ++s) lens[s] = 7;
                for (int s = 280; s < 288; ++s) lens[s] = 8;
                huff_build(hl, lens, 288);
                for (int s = 0; s < 30; ++s) lens[s] = 5;
                huff_build(hd, lens, 30);
            } else {
                const int nlen = (int)bits_take(b, 5) + 257;
                const int ndist = (int)bits_take(b, 5) + 1;
                const int ncode = (int)bits_take(b, 4) + 4;
                if (nlen > 286 || ndist > 30) return INF_ERR_HEADER;
                for (int i = 0; i < 19; ++i) lens[i] = 0;
                for (int i = 0; i < ncode; ++i) {
                    // order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
                    const int j = i - 4;
                    const int pos = i < 3 ? 16 + i : (i == 3 ? 0 : ((j & 1) ? 7 - (j >> 1) : 8 + (j >> 1)));
                    bits_refill(b);
                    lens[pos] = (uint8_t)bits_take(b, 3);
                }
                if (huff_build(hl, lens, 19) != 0) return INF_ERR_HEADER;       // the code-length code must be complete
                int idx = 0;
                while (idx < nlen + ndist) {
                    bits_refill(b);
                    int sym = huff_decode_slow(b, hl);
                    if (sym < 0) return INF_ERR_HEADER;
                    if (sym < 16) lens[19 + idx++] = (uint8_t)sym;
                    else {
                        int rep, val = 0;
                        if (sym == 16) {
                            if (idx == 0) return INF_ERR_HEADER;
                            val = lens[19 + idx - 1];
                            rep = 3 + (int)bits_take(b, 2);
                        } else if (sym == 17) rep = 3 + (int)bits_take(b, 3);
                        else rep = 11 + (int)bits_take(b, 7);
                        if (idx + rep > nlen + ndist) return INF_ERR_HEADER;
                        while (rep--) lens[19 + idx++] = (uint8_t)val;
                    }
                }
                if (bits_over(b)) return INF_ERR_INPUT;
                if (lens[19 + 256] == 0) return INF_ERR_HEADER;                 // no end-of-block code
                // (zlib: an incomplete code is only accepted when it is a single code of length 1)
                int left = huff_build(hl, lens + 19, nlen);
                if (left < 0 || (left > 0 && !(hl.count[1] == 1 && nlen - hl.count[0] == 1))) return INF_ERR_HEADER;
                left = huff_build(hd, lens + 19 + nlen, ndist);
                if (left < 0 || (left > 0 && !(hd.count[1] == 1 && ndist - hd.count[0] == 1))) return INF_ERR_HEADER;
            }
            lut_build<LUT_BITS, true, NL>(hl, lut, lane);
            lut_build<DLUT_BITS, false, NL>(hd, dlut, lane);
            for (;;) {
                bits_refill(b);
                uint32_t e = lut[(uint32_t)b.buf & (uint32_t)(LUT_SIZE - 1)];
                if (e & LIT_FLAG) {
                    // literals, up to three per refill (a code in the table has at most LUT_BITS = 10 bits, 33 are there):
                    // look-up, one flag test, store, shift — the chain a member's decode time is made of
                    if (n_out - o >= 3u) {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            if (writer) out[o] = (uint8_t)(e >> 4);
                            ++o;
                            b.buf >>= (e & 15u); b.n -= (int)(e & 15u);
                            if (r == 2) break;
                            e = lut[(uint32_t)b.buf & (uint32_t)(LUT_SIZE - 1)];
                            if (!(e & LIT_FLAG)) break;
                        }
                    } else {
                        if (o >= n_out) return INF_ERR_OUTPUT;
                        if (writer) out[o] = (uint8_t)(e >> 4);
                        ++o;
                        b.buf >>= (e & 15u); b.n -= (int)(e & 15u);
                    }
                    continue;
                }
                int sym;
                if (e) { b.buf >>= (e & 15u); b.n -= (int)(e & 15u); sym = (int)(e >> 4); }
                else {
                    sym = huff_decode_slow(b, hl);
                    if (sym < 0) return INF_ERR_CODE;
                    if (sym < 256) {
                        if (o >= n_out) return INF_ERR_OUTPUT;
                        if (writer) out[o] = (uint8_t)sym;
                        ++o;
                        continue;
                    }
                }
                if (sym == 256) break;
                if (sym > 285) return INF_ERR_CODE;
                int len;
                if (sym < 265) len = sym - 254;
                else if (sym == 285) len = 258;
                else {
                    const int x = (sym - 261) >> 2;
                    len = ((4 | ((sym - 265) & 3)) << x) + 3 + (int)bits_take(b, x);
                }
                bits_refill(b);
                const int ds = huff_decode<DLUT_BITS>(b, hd, dlut);
                if (ds < 0 || ds > 29) return INF_ERR_DIST;
                uint32_t dist;
                if (ds < 4) dist = (uint32_t)ds + 1u;
                else {
                    const int x = (ds >> 1) - 1;
                    dist = ((uint32_t)(2 | (ds & 1)) << x) + 1u + bits_take(b, x);
                }
                if (dist > o) return INF_ERR_DIST;
                if ((uint32_t)len > n_out - o) return INF_ERR_OUTPUT;
                // the copy.  Byte by byte, every byte waits for its own load to come back (the store needs the value, the
                // next load may alias the store): a round trip to L2 per byte.  So: a run (distance 1) loads once; a source
                // at least 8 bytes back is fetched 8 bytes at a time, loads first, stores after; only short overlapping
                // periods go byte by byte.  (Measured: 8 + 8 predicated by the length = 45 instructions per match; jumps into
                // unrolled runs of exactly `len` = fewer instructions and MORE time, the indirect branches stall.)
                if (NL > 1) {
                    // all lanes: NL bytes per step.  A source that overlaps its destination (distance < length) repeats with
                    // the distance as its period, so every byte is read from in front of the match
                    TCI_WARP_SYNC();                            // what other lanes wrote of this member so far is visible
                    const uint8_t* src = out + o - dist;
                    uint8_t* dst = out + o;
                    if (dist >= (uint32_t)len) {
                        for (int i = lane; i < len; i += NL) dst[i] = src[i];
                    } else {
                        for (int i = lane; i < len; i += NL) dst[i] = src[(uint32_t)i % dist];
                    }
                    o += (uint32_t)len;
                } else if (dist == 1) {
                    const uint8_t v = out[o - 1];
                    for (int i = 0; i < len; ++i) out[o + i] = v;
                    o += len;
                } else if (dist >= 8) {
                    const uint8_t* src = out + o - dist;
                    uint8_t* dst = out + o;
                    if (n_out - o >= (uint32_t)len + 7u) {
                        // whole groups of 8, no predicates: the bytes a group writes beyond the match's end are overwritten by
                        // what the stream produces next (the output only grows, and the member must end exactly at n_out)
                        o += (uint32_t)len;
                        for (int done = 0; done < len; done += 8, src += 8, dst += 8) {
                            const uint8_t t0 = src[0], t1 = src[1], t2 = src[2], t3 = src[3], t4 = src[4], t5 = src[5], t6 = src[6], t7 = src[7];
                            dst[0] = t0; dst[1] = t1; dst[2] = t2; dst[3] = t3; dst[4] = t4; dst[5] = t5; dst[6] = t6; dst[7] = t7;
                        }
                    } else {
                        // the last bytes of the member: exactly `len`
                        o += (uint32_t)len;
                        while (len > 0) {
                            const int n8 = len < 8 ? len : 8;
                            uint8_t t[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (j < n8) t[j] = src[j];
#pragma unroll
                            for (int j = 0; j < 8; ++j) if (j < n8) dst[j] = t[j];
                            src += n8; dst += n8; len -= n8;
                        }
                    }
                } else {
                    for (int i = 0; i < len; ++i, ++o) out[o] = out[o - dist];
                }
                if (bits_over(b)) return INF_ERR_INPUT;
            }
            if (bits_over(b)) return INF_ERR_INPUT;
        } else {
            return INF_ERR_BTYPE;
        }
        if (bfinal) break;
    }
    if (o != n_out) return INF_ERR_OUTPUT;
    if (bits_over(b)) return INF_ERR_INPUT;
    return INF_OK;
}

// ---------------------------------------------------------------- CRC-32 (the gzip polynomial, reflected)
constexpr uint32_t CRC_POLY = 0xedb88320u;

TCI_HD uint32_t crc_table_entry(uint32_t i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? CRC_POLY ^ (c >> 1) : c >> 1;
    return c;
}
// a(x) * b(x) mod p(x), bit-reflected operands (zlib's multmodp)
TCI_HD uint32_t crc_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod p(x)
TCI_HD uint32_t crc_x8nmodp(uint64_t n) {
    uint32_t t = 1u << 30;                       // x^1
    for (int k = 0; k < 3; ++k) t = crc_multmodp(t, t);     // x^8
    uint32_t p = 1u << 31;                       // x^0
    while (n) {
        if (n & 1u) p = crc_multmodp(t, p);
        n >>= 1;
        if (n) t = crc_multmodp(t, t);
    }
    return p;
}
// CRC of A || B from the CRCs of A and B and the length of B
TCI_HD uint32_t crc_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
    return crc_multmodp(crc_x8nmodp(len_b), crc_a) ^ crc_b;
}

}  // namespace tcinf
