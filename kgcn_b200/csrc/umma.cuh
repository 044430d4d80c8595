// tcgen05 (5th-generation tensor core) / TMEM PTX wrappers for sm_100a, kind::tf32, cta_group::1.
//
// Shared-memory operand layout used throughout: K-major, SWIZZLE_128B.  One swizzle atom is
// 8 rows x 128 bytes (32 tf32); 8-row groups are 1024 B apart (SBO), atoms along K are placed
// `atom_stride` bytes apart and are addressed by moving the descriptor start address; inside an
// atom a K-step of 8 tf32 is +32 B.  The 16-byte chunk `c` of row `r` lives at chunk position
// c ^ (r & 7) (the XOR is on absolute address bits [4,7) ^ [7,10), so atoms are 1024-B aligned).
#pragma once
#include <cstdint>

namespace kgcn {

// byte offset of element (row r, k-index kk) inside an operand whose atoms are atom_stride apart
__device__ __forceinline__ uint32_t sw128_offset(int r, int kk, uint32_t atom_stride) {
    const uint32_t atom = static_cast<uint32_t>(kk) >> 5;
    const uint32_t chunk = (static_cast<uint32_t>(kk) >> 2) & 7u;
    return atom * atom_stride + (static_cast<uint32_t>(r) >> 3) * 1024u + (static_cast<uint32_t>(r) & 7u) * 128u +
           ((chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4) + (static_cast<uint32_t>(kk) & 3u) * 4u;
}

// 64-bit shared-memory matrix descriptor (K-major, SWIZZLE_128B, SBO = 1024 B, version 1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) /* LBO (unused) */ |
           (static_cast<uint64_t>(1024 >> 4) << 32) /* SBO */ | (1ull << 46) /* version */ |
           (2ull << 61) /* SWIZZLE_128B */;
}

// instruction descriptor: D = f32, A = B = tf32, both K-major, M = m (64 or 128), N = n
// (multiple of 8 for M = 64, of 16 for M = 128)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// Round-to-nearest (ties away) tf32 with the low 13 mantissa bits zero, in 2 integer instructions.
// (cvt.rna.tf32.f32 lowers to 4+ instructions per element because it special-cases Inf/NaN; here
// Inf stays Inf, NaN stays NaN, and the largest finite values round to Inf like any rounding would.)
__device__ __forceinline__ float tf32_hi(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// warp-collective: allocate `cols` (power of two >= 32) TMEM columns, base address -> *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_slot))),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns starting at taddr
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// must be executed (by the whole warp) before the registers of preceding tmem loads are read
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ties 16 loaded values to a point AFTER tmem_ld_wait() in the volatile-asm order, so the compiler
// cannot schedule arithmetic on them above the wait
__device__ __forceinline__ void tmem_ld_fence(float (&v)[16]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                      "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
}

// TMEM -> registers, 16 lanes x (8 * REPS) fp32 columns starting at taddr (shape .16x256b).
// For repetition i (columns 8i .. 8i+7) thread t holds:
//   v[4i+0], v[4i+1] : row  t/4      , columns 8i + 2*(t%4) + {0, 1}
//   v[4i+2], v[4i+3] : row  t/4 + 8  , same columns
// (the m16n8 accumulator fragment layout), so all 32 threads carry data of a 16-row M=64 slab.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// same shape, 16 columns: v[4i + {0,1}] row t/4, v[4i + {2,3}] row t/4 + 8, columns 8i + 2*(t%4) + {0,1}, i = 0..1
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_fence8(float (&v)[8]) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}

// MN-major operands (the contraction index is the ROW index of a row-major tile).  For tf32 the only
// MN-major shared-memory layout is SWIZZLE_128B with a 32-byte base: rows of 128 bytes (32 values along
// M/N) for consecutive k, 32-value chunks `lbo` bytes apart, and inside a row the 32-byte granule
// index is XORed with (k & 3).  One K = 8 instruction spans two 4-row swizzle atoms (`sbo` = 512 B).
__device__ __forceinline__ uint32_t mn32_offset(uint32_t krow, uint32_t mn, uint32_t lbo) {   // mn % 4 == 0 for 16-byte accesses
    return (mn >> 5) * lbo + krow * 128u + ((((mn >> 3) & 3u) ^ (krow & 3u)) << 5) + ((mn & 7u) << 2);
}
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) /* version */ |
           (1ull << 61) /* SWIZZLE_128B_BASE32B */;
}
constexpr uint32_t kUmmaMajorMnA = 1u << 15, kUmmaMajorMnB = 1u << 16;   // instruction-descriptor bits

}  // namespace kgcn
