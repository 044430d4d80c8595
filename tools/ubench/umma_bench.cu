// Micro-benchmark: cycles per tcgen05.mma kind::tf32 (K = 8 per instruction) for the shapes the fused
// GraphConv kernel could use.  One CTA per SM; thread 0 issues R MMAs, commits, waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../kgcn_b200/csrc/tma.cuh"
#include "../../kgcn_b200/csrc/umma.cuh"
using namespace kgcn;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// mode: 0 = SS dependent (one accumulator), 1 = SS 3 independent accumulators, 2 = TS dependent, 3 = TS independent x3
__global__ void __launch_bounds__(128, 1) bench(int kind_f16, int M, int N, int mode_in, int reps, int ksteps, long long* out) {
    int mode = mode_in;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x; i < (96u * 1024u) / 16u; i += blockDim.x)
        asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "f"(0.001f) : "memory");
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    const bool warp_mode = mode >= 4;   // 4/5: whole warp converged, one lane elected per MMA (CUTLASS style)
    if (warp_mode) mode -= 4;
    uint32_t is_leader = 0;
    if (threadIdx.x < 32) {
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
    }
    if (warp_mode ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
        // f16 kind: a_format = b_format = 0 (F16), D = f32
        const uint32_t idesc = kind_f16 ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24)) : umma_idesc_tf32(M, N);
        const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 48 * 1024);
        const uint32_t a_atom = (uint32_t)M * 128u >> 4, b_atom = (uint32_t)N * 128u >> 4;
        uint32_t parity = 0;
        for (int warm = 0; warm < 2; ++warm) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t d = tm + ((mode & 1) ? (uint32_t)((r % 3) * 256 / 2) : 0u);   // accumulators 128 columns apart
                    const uint64_t ad = da + (uint64_t)((k >> 2) * a_atom + (k & 3) * 2);
                    const uint64_t bd = db + (uint64_t)((k >> 2) * b_atom + (k & 3) * 2);
                    if (!warp_mode || is_leader) {
                        if (mode >= 2) umma_tf32_ts(d, tm + 384 + (uint32_t)(k * 8), bd, idesc, 1);
                        else if (kind_f16) umma_f16(d, ad, bd, idesc, 1);
                        else umma_tf32(d, ad, bd, idesc, 1);
                    }
                    if (warp_mode) __syncwarp();
                }
            }
            const long long t1 = clock64();
            if (!warp_mode || is_leader) umma_commit(&bar);
            if (warp_mode) __syncwarp();
            mbar_wait(&bar, parity);
            parity ^= 1;
            const long long t2 = clock64();
            if (warm == 1 && blockIdx.x == 0 && (threadIdx.x == 0)) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int shapes[][2] = {{128, 64}, {64, 64}, {128, 128}, {128, 256}, {64, 128}, {64, 256}, {128, 32}, {128, 16}};
    const char* modes[] = {"SS dep", "SS ind3", "TS dep", "TS ind3", "SS dep W", "SS ind3 W", "TS dep W", "TS ind3 W"};
    printf("%-8s %-8s %-8s %10s %10s\n", "M", "N", "mode", "issue c/mma", "total c/mma");
    for (int kind = 0; kind < 2; ++kind)
    for (auto& sh : shapes)
        for (int mode = 0; mode < 1; ++mode) {
            const int reps = 64, ks = 8;
            bench<<<148, 128, 200 * 1024>>>(kind, sh[0], sh[1], mode, reps, ks, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2] = {0, 0};
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%s %-8d %-8d %-8s %10.1f %10.1f %s\n", kind ? "f16 " : "tf32", sh[0], sh[1], modes[mode], h[0] / double(reps * ks), h[1] / double(reps * ks),
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
