// pileup_smem.cuh — device helpers shared by the warp-stream pileup kernels (pileup_warp.cu):
// CIGAR op classes, carry-save adders, one-hot checks of packed base codes, explicit shared-window
// loads / stores / reductions, and TMA bulk copies (cp.async.bulk) completing on an mbarrier.  sm_100a only.
#pragma once
#include "pileup.cuh"

namespace tcsm {

constexpr unsigned FULL = 0xffffffffu;

// per op: bit0 = M/=/X, bit1 = consumes reference, bit2 = consumes query (MIDNSHP=X -> 0..8)
constexpr uint32_t OPFLAGS = 7u | (4u << 3) | (2u << 6) | (2u << 9) | (4u << 12) | (7u << 21) | (7u << 24);
// clamped shift: ops 11..15 (not defined by BAM) index past the table and read 0
__device__ __forceinline__ uint32_t op_flags(uint32_t op) { return __funnelshift_rc(OPFLAGS, 0u, 3u * op) & 7u; }

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    l = a ^ b ^ c;
    h = (a & b) | (c & (a | b));
}

// non-zero <=> some nibble of w has two or more bits set
__device__ __forceinline__ uint32_t multibit(uint32_t w) { return w & ((w | 0x88888888u) - 0x11111111u); }

__device__ __forceinline__ uint32_t clear_multibit(uint32_t w) {
    const uint32_t z = multibit(w);
    const uint32_t m = (z | (z >> 1) | (z >> 2) | (z >> 3)) & 0x11111111u;
    return w & ~(m * 15u);
}

// Shared memory is addressed through 32-bit shared-window byte addresses and explicit ld/st/red.shared:
// the lane-private rows, descriptor lists and counters are indexed with data-dependent offsets in every
// inner loop, and this keeps each access at one address add + one LDS/STS/ATOMS.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// X and I events go straight to a packed global counter per column (X count in the low, I count in the high 32
// bits): they are sparse (a few per read), and keeping them out of shared memory buys another resident warp.
// 32-bit adds on the halves: +1 X (a deletion column), +1 I (an insertion anchor), and -1 X as well when the anchor
// is a deletion's last column (that column reads "*+n..", no longer "*").
__device__ __forceinline__ void red_u32(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// red.shared.or of a word that has any bit set (predicated, no branch): an empty word costs no shared-memory cycle
__device__ __forceinline__ void reds_or_nz(uint32_t a, uint32_t v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(a), "r"(v) : "memory");
}
// X | I event: one 64-bit add to column col's packed counter (X in the low, I in the high 32 bits), predicated.  lo = 1: +1 X;
// hi = 1: +1 I; lo = 2^32 - 1, hi = 0: +1 I and -1 X (the carry does it; the sums are exact modulo 2^64).
__device__ __forceinline__ void red_xi_if(bool p, unsigned long long* xi, uint32_t col, uint32_t lo, uint32_t hi) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 v, q;\n\tsetp.ne.u32 p, %0, 0;\n\tmov.b64 v, {%3, %4};\n\tmad.wide.u32 q, %2, 8, %1;\n\t@p red.global.add.u64 [q], v;\n\t}"
                 ::"r"((uint32_t)p), "l"(xi), "r"(col), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ uint4 lds4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts2(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
// two BAM CIGAR words -> two 16-bit ops (len << 4 | op, len < 4096) in one word
__device__ __forceinline__ uint32_t pack_ops(uint32_t c0, uint32_t c1) { return __byte_perm(c0, c1, 0x5410); }
// pads and zero-length ops: min((op ^ P), len) is 0 exactly for those
__device__ __forceinline__ uint32_t op_exotic_min(uint32_t c) { return min((c & 15u) ^ 6u, c >> 4); }
constexpr uint32_t OP_BIG = 1u << 16;           // an op of 4096+ bases does not fit 16 bits
constexpr uint32_t S_SATURATED = (4095u << 4) | OP_S;

// ---- TMA (cp.async.bulk) staging: one lane asks for the sub-tile's SEQ words and CIGAR ops as two bulk copies that
// complete on the warp's mbarrier; nothing is held in registers while they are in flight
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "TC_MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra TC_MBAR_DONE;\n\t"
        "bra TC_MBAR_WAIT;\n\t"
        "TC_MBAR_DONE:\n\t"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

}  // namespace tcsm
