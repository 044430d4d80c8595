// Batched CSR SpMM for sm_100a: the Bspmm / Bconv / Bspmdt ops (include/kgcn_b200.h).
//
// One CTA owns one graph.  Its right-hand-side tile(s) ([n_cols, feat] fp32, contiguous in the
// [B, N, F] feature tensor) are staged into shared memory with ONE bulk-async copy (TMA 1-D,
// cp.async.bulk + mbarrier complete_tx) while the warps already fetch their CSR row extents; the
// neighbour gather then runs entirely out of shared memory.  Rows are handled by sub-warp lane
// groups (LPR lanes x VEC floats cover one feature row), entries of a row are fetched coalesced
// by the group and broadcast with shuffles ("CSR-vector" segmented sum), and results leave with
// 16-byte coalesced stores.  HBM traffic is therefore the algorithmic minimum: X once, Y once,
// CSR once.  Graphs whose tile does not fit / is not 16-byte aligned, and the B=1 large-N
// block-diagonal use (example_model/sparse.py:65-69), take the same kernel with STAGED=false
// (gathers straight from global / L2).
#include "common.cuh"

namespace kgcn {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
};
template <>
struct Vec<2> {
    using T = float2;
};
template <>
struct Vec<1> {
    using T = float;
};

template <int VEC>
__device__ __forceinline__ void vload(float (&r)[VEC], const float* p) {
    typename Vec<VEC>::T v = *reinterpret_cast<const typename Vec<VEC>::T*>(p);
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = f[k];
}
template <int VEC>
__device__ __forceinline__ void vstore(float* p, const float (&r)[VEC]) {
    typename Vec<VEC>::T v;
    float* f = reinterpret_cast<float*>(&v);
#pragma unroll
    for (int k = 0; k < VEC; ++k) f[k] = r[k];
    *reinterpret_cast<typename Vec<VEC>::T*>(p) = v;
}

struct SpmmParams {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* rhs;
    float* out;
    const float* self_scale;
    int64_t rs_g, rs_c, os_g, os_c;
    int channels, n_rows, n_cols, feat;
    int lpr_log2;  // lanes per row = 1 << lpr_log2
    int act;
};

constexpr int kSpmmThreads = 128;

template <int VEC, bool STAGED>
__global__ void __launch_bounds__(kSpmmThreads) bspmm_kernel(const SpmmParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw + 128);

    const int64_t g = blockIdx.x;
    const int C = p.channels;
    const bool sum_channels = (p.os_c == 0);
    const bool shared_rhs = (p.rs_c == 0);
    const int tile_elems = p.n_cols * p.feat;
    const float* rhs_g = p.rhs + g * p.rs_g;

    if (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const int n_tiles = shared_rhs ? 1 : C;
            const uint32_t bytes = static_cast<uint32_t>(tile_elems) * 4u;
            mbar_expect_tx(bar, bytes * n_tiles);
            for (int c = 0; c < n_tiles; ++c)
                bulk_g2s(tile + static_cast<size_t>(c) * tile_elems, rhs_g + c * p.rs_c, bytes, bar);
        }
        __syncthreads();  // barrier init visible to the waiters
    }

    const int lpr = 1 << p.lpr_log2;
    const int lane = threadIdx.x & 31;
    const int sub = lane & (lpr - 1);
    const int grp_base = lane & ~(lpr - 1);
    const unsigned gmask = (lpr == 32) ? 0xffffffffu : (((1u << lpr) - 1u) << grp_base);
    const int groups_per_cta = kSpmmThreads >> p.lpr_log2;
    const int group = threadIdx.x >> p.lpr_log2;

    const int n_out_rows = sum_channels ? p.n_rows : C * p.n_rows;
    const int chunk = lpr * VEC;
    const int n_chunks = (p.feat + chunk - 1) / chunk;
    const int32_t* rp_g = p.rowptr + g * static_cast<int64_t>(C) * p.n_rows;

    bool waited = !STAGED;
    for (int q = group; q < n_out_rows; q += groups_per_cta) {
        const int i = sum_channels ? q : q % p.n_rows;
        const int c_begin = sum_channels ? 0 : q / p.n_rows;
        const int c_end = sum_channels ? C : c_begin + 1;
        float* out_row = p.out + g * p.os_g + c_begin * p.os_c + static_cast<int64_t>(i) * p.feat;
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int f0 = ch * chunk + sub * VEC;
            const bool active = f0 < p.feat;  // feat % VEC == 0 by construction
            float acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
            for (int c = c_begin; c < c_end; ++c) {
                const int32_t s = __ldg(rp_g + c * p.n_rows + i);
                const int32_t e = __ldg(rp_g + c * p.n_rows + i + 1);
                const float* x_c;
                if (STAGED)
                    x_c = tile + (shared_rhs ? 0 : static_cast<size_t>(c) * tile_elems);
                else
                    x_c = rhs_g + c * p.rs_c;
                for (int32_t base = s; base < e; base += lpr) {
                    const int32_t mine = base + sub;
                    int my_col = 0;
                    float my_val = 0.0f;
                    if (mine < e) {
                        my_col = __ldg(p.col + mine);
                        my_val = __ldg(p.val + mine);
                    }
                    if (!waited) {  // first use of the staged tile
                        mbar_wait(bar, 0);
                        waited = true;
                    }
                    const int cnt = min(lpr, e - base);
                    for (int k = 0; k < cnt; ++k) {
                        const int j = __shfl_sync(gmask, my_col, grp_base + k);
                        const float v = __shfl_sync(gmask, my_val, grp_base + k);
                        if (active) {
                            float xv[VEC];
                            vload<VEC>(xv, x_c + static_cast<size_t>(j) * p.feat + f0);
#pragma unroll
                            for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                        }
                    }
                }
                if (p.self_scale != nullptr && active) {  // GIN: + eps_c * x[i]  (layers.py:469)
                    if (!waited) {
                        mbar_wait(bar, 0);
                        waited = true;
                    }
                    const float eps = __ldg(p.self_scale + c);
                    float xv[VEC];
                    vload<VEC>(xv, x_c + static_cast<size_t>(i) * p.feat + f0);
#pragma unroll
                    for (int t = 0; t < VEC; ++t) acc[t] = fmaf(eps, xv[t], acc[t]);
                }
            }
            if (active) {
                if (p.act != KGCN_ACT_NONE) {
#pragma unroll
                    for (int t = 0; t < VEC; ++t) acc[t] = apply_act(acc[t], p.act);
                }
                vstore<VEC>(out_row + f0, acc);
            }
        }
    }
    if (STAGED && !waited) mbar_wait(bar, 0);  // never leave with a bulk copy in flight
}

// dval[e] = < dy[row_e, :], rhs[col_e, :] >  (bspmm_call.py:49-54); one warp per CSR row.
__global__ void __launch_bounds__(128) bspmm_dvalues_kernel(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col,
                                                            const int32_t* __restrict__ perm, int64_t total_rows,
                                                            int channels, int n_rows, int feat,
                                                            const float* __restrict__ dy, int64_t ds_g, int64_t ds_c,
                                                            const float* __restrict__ rhs, int64_t rs_g, int64_t rs_c,
                                                            float* __restrict__ dval) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= total_rows) return;
    const int64_t m = row / n_rows;
    const int i = static_cast<int>(row % n_rows);
    const int64_t g = m / channels;
    const int c = static_cast<int>(m % channels);
    const float* dy_row = dy + g * ds_g + c * ds_c + static_cast<int64_t>(i) * feat;
    const float* rhs_m = rhs + g * rs_g + c * rs_c;
    const int32_t s = rowptr[row], e = rowptr[row + 1];
    for (int32_t k = s; k < e; ++k) {
        const float* r = rhs_m + static_cast<int64_t>(col[k]) * feat;
        float acc = 0.0f;
        for (int f = lane; f < feat; f += 32) acc = fmaf(dy_row[f], r[f], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) dval[perm ? perm[k] : k] = acc;
    }
}

template <int VEC>
int launch_vec(const SpmmParams& p, int64_t n_graphs, bool staged, size_t smem, cudaStream_t st) {
    if (staged) {
        if (smem > 48 * 1024) {
            static bool attr_set = false;  // idempotent; racing threads set the same value
            if (!attr_set) {
                KGCN_CUDA_OK(cudaFuncSetAttribute(bspmm_kernel<VEC, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set = true;
            }
        }
        bspmm_kernel<VEC, true><<<static_cast<unsigned>(n_graphs), kSpmmThreads, smem, st>>>(p);
    } else {
        bspmm_kernel<VEC, false><<<static_cast<unsigned>(n_graphs), kSpmmThreads, 0, st>>>(p);
    }
    KGCN_LAUNCH_OK("bspmm_kernel");
    return KGCN_OK;
}

}  // namespace

int launch_bspmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs, int channels,
                 int n_rows, int n_cols, int feat, const float* rhs, int64_t rs_g, int64_t rs_c, float* out,
                 int64_t os_g, int64_t os_c, const float* self_scale, int act, cudaStream_t st) {
    KGCN_REQUIRE(rowptr && col && val && rhs && out, KGCN_ERR_NULL, "bspmm: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_rows > 0 && n_cols > 0 && feat > 0, KGCN_ERR_BAD_SHAPE,
                 "bspmm: bad shape n_graphs=%lld channels=%d n_rows=%d n_cols=%d feat=%d", (long long)n_graphs,
                 channels, n_rows, n_cols, feat);
    KGCN_REQUIRE(n_graphs < (1ll << 31), KGCN_ERR_BAD_SHAPE, "bspmm: n_graphs too large for one launch");
    KGCN_REQUIRE(self_scale == nullptr || n_rows == n_cols, KGCN_ERR_BAD_SHAPE,
                 "bspmm: self_scale needs square matrices");
    if (n_graphs == 0) return KGCN_OK;

    // widest vector width the row pitch and all base offsets allow
    auto ok = [&](int v) {
        const uintptr_t a = static_cast<uintptr_t>(v) * 4;
        return feat % v == 0 && reinterpret_cast<uintptr_t>(rhs) % a == 0 && reinterpret_cast<uintptr_t>(out) % a == 0 &&
               rs_g % v == 0 && rs_c % v == 0 && os_g % v == 0 && os_c % v == 0;
    };
    const int vec = ok(4) ? 4 : (ok(2) ? 2 : 1);
    int lpr_log2 = 0;
    while ((1 << lpr_log2) < 32 && (1 << lpr_log2) * vec < feat) ++lpr_log2;

    const size_t tile_bytes = static_cast<size_t>(n_cols) * feat * 4;
    const size_t n_tiles = (rs_c == 0) ? 1 : channels;
    const bool staged = tile_bytes % 16 == 0 && aligned16(rhs) && (rs_g * 4) % 16 == 0 && (rs_c * 4) % 16 == 0 &&
                        tile_bytes * n_tiles <= 96 * 1024 && tile_bytes * n_tiles < (1u << 20);
    const size_t smem = staged ? 128 + tile_bytes * n_tiles : 0;

    SpmmParams p{rowptr, col, val, rhs, out, self_scale, rs_g, rs_c, os_g, os_c, channels, n_rows, n_cols, feat, lpr_log2, act};
    switch (vec) {
        case 4: return launch_vec<4>(p, n_graphs, staged, smem, st);
        case 2: return launch_vec<2>(p, n_graphs, staged, smem, st);
        default: return launch_vec<1>(p, n_graphs, staged, smem, st);
    }
}

}  // namespace kgcn

extern "C" int kgcn_bspmm_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                              int32_t channels, int32_t n_rows, int32_t n_cols, int32_t feat, const float* rhs,
                              int64_t rhs_stride_g, int64_t rhs_stride_c, float* out, int64_t out_stride_g,
                              int64_t out_stride_c, const float* self_scale, void* stream) {
    return kgcn::launch_bspmm(rowptr, col, val, n_graphs, channels, n_rows, n_cols, feat, rhs, rhs_stride_g,
                              rhs_stride_c, out, out_stride_g, out_stride_c, self_scale, KGCN_ACT_NONE,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int kgcn_bspmm_dvalues_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm, int64_t n_graphs,
                                      int32_t channels, int32_t n_rows, int32_t n_cols, int32_t feat, const float* dy,
                                      int64_t dy_stride_g, int64_t dy_stride_c, const float* rhs, int64_t rhs_stride_g,
                                      int64_t rhs_stride_c, float* dval, void* stream) {
    using namespace kgcn;
    KGCN_REQUIRE(rowptr && col && dy && rhs && dval, KGCN_ERR_NULL, "bspmm_dvalues: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_rows > 0 && n_cols > 0 && feat > 0, KGCN_ERR_BAD_SHAPE,
                 "bspmm_dvalues: bad shape");
    const int64_t total_rows = n_graphs * channels * n_rows;
    if (total_rows == 0) return KGCN_OK;
    const int64_t blocks = ceil_div<int64_t>(total_rows * 32, 128);
    KGCN_REQUIRE(blocks < (1ll << 31), KGCN_ERR_BAD_SHAPE, "bspmm_dvalues: too many rows for one launch");
    bspmm_dvalues_kernel<<<static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, col, perm, total_rows, channels, n_rows, feat, dy, dy_stride_g, dy_stride_c, rhs, rhs_stride_g,
        rhs_stride_c, dval);
    KGCN_LAUNCH_OK("bspmm_dvalues_kernel");
    return KGCN_OK;
}
