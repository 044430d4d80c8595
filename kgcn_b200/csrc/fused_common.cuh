// Device helpers shared by the fused GraphConv forward (graphconv_fused.cu) and backward
// (graphconv_fused_bwd.cu) kernels: explicit shared-memory loads/stores by 32-bit address, the
// swizzled output staging tile, warp election, the consumer-only named barrier and the per-stage
// descriptor the TMA producer hands to the consumers.
#pragma once
#include <cstdint>

#include "common.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ int2 lds_i2(uint32_t addr) {
    int2 r;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
template <int VEC>
__device__ __forceinline__ void lds_f(float (&r)[VEC], uint32_t addr);
template <>
__device__ __forceinline__ void lds_f<4>(float (&r)[4], uint32_t addr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_f<2>(float (&r)[2], uint32_t addr) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_f<1>(float (&r)[1], uint32_t addr) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[0]) : "r"(addr));
}
template <int VEC>
__device__ __forceinline__ void sts_f(uint32_t addr, const float (&r)[VEC]);
template <>
__device__ __forceinline__ void sts_f<4>(uint32_t addr, const float (&r)[4]) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
}
template <>
__device__ __forceinline__ void sts_f<2>(uint32_t addr, const float (&r)[2]) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(r[0]), "f"(r[1]) : "memory");
}
template <>
__device__ __forceinline__ void sts_f<1>(uint32_t addr, const float (&r)[1]) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(r[0]) : "memory");
}

__device__ __forceinline__ uint32_t ystage_off(uint32_t row, uint32_t col, uint32_t ypitch) {   // col % 4 == 0 or 2
    const uint32_t chunk = col >> 2;
    return row * ypitch + ((((chunk & 7u) ^ (row & 7u)) | (chunk & ~7u)) << 4) + ((col & 3u) << 2);
}


// staged tile -> global, coalesced.  y_tile points at the tile's first output row.
template <int NC>
__device__ __forceinline__ void copy_out(uint32_t ys, uint32_t ypitch, float* y_tile, int rows, int f_out, int tid, bool vec4) {
    if (vec4) {
        const int cpr = f_out >> 2;   // 16-byte chunks per row
        for (int idx = tid; idx < rows * cpr; idx += NC) {
            const uint32_t r = static_cast<uint32_t>(idx / cpr), c = static_cast<uint32_t>(idx - r * cpr);
            float t[4];
            lds_f<4>(t, ys + ystage_off(r, c << 2, ypitch));
            *reinterpret_cast<float4*>(y_tile + static_cast<size_t>(r) * f_out + (c << 2)) = make_float4(t[0], t[1], t[2], t[3]);
        }
    } else {
        for (int idx = tid; idx < rows * f_out; idx += NC) {
            const uint32_t r = static_cast<uint32_t>(idx / f_out), c = static_cast<uint32_t>(idx - r * f_out);
            y_tile[idx] = __uint_as_float(lds_u32(ys + ystage_off(r, c & ~3u, ypitch) + ((c & 3u) << 2)));
        }
    }
}

__device__ __forceinline__ bool elect_one() {   // exactly one lane of the (converged) warp gets true
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}


// 4 consumer threads per tile row (8 warps for 64-row tiles, 16 warps for 128-row tiles) + 1 producer warp
template <int NC>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

struct StageInfo {   // written by the producer before it arms full[s]
    int32_t e_first;  // first CSR entry of the tile
    int32_t n_entries;
    int32_t rp_skip;  // ints to skip in the staged rowptr slice (16-byte alignment of the copy)
    int32_t e_skip;   // entries to skip in the staged col / val slices
    int32_t staged;   // 0: the tile has more entries than the stage holds -> read col/val from global
    int32_t pad[3];
};

// One pass of the 3xTF32 contraction (pass 0: Zhi.Whi, 1: Zlo.Whi, 2: Zhi.Wlo) in K-steps of 8
// (32 bytes), fully unrolled for NA K-atoms so that every descriptor is a uniform-register add of a
// constant.  A single thread needs ~100+ cycles of scalar work per tcgen05.mma (tools/ubench), so the
// three passes are issued by three different warps into three TMEM accumulators that the epilogue adds.
template <int NA>
__device__ __forceinline__ void issue_pass(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                           uint32_t z_atom16, uint32_t w_atom16, int K) {
    uint32_t acc = 0;
#pragma unroll
    for (int at = 0; at < NA; ++at) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (at * 32 + ks * 8 < K) {  // skip k-steps that only see padding
                umma_tf32(tmem_d, da + static_cast<uint64_t>(at * z_atom16 + 2 * ks),
                          db + static_cast<uint64_t>(at * w_atom16 + 2 * ks), idesc, acc);
                acc = 1;
            }
        }
    }
}
__device__ __noinline__ void issue_pass_loop(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                             uint32_t z_atom16, uint32_t w_atom16, int K, int n_atoms) {
    uint32_t acc = 0;
    int k_left = K;
    for (int at = 0; at < n_atoms; ++at, da += z_atom16, db += w_atom16, k_left -= 32)
        for (int ks = 0; ks < 4; ++ks)
            if (ks * 8 < k_left) {
                umma_tf32(tmem_d, da + 2u * ks, db + 2u * ks, idesc, acc);
                acc = 1;
            }
}

}  // namespace
}  // namespace kgcn
