// Batched CSR SpMM for sm_100a: the Bspmm / Bconv / Bspmdt ops (include/kgcn_b200.h).
//
// Tile kernel (the hot path): one CTA owns G consecutive graphs.  Their right-hand-side tiles
// ([n_cols, feat] fp32 each, contiguous in the [B, N, F] feature tensor) are staged into shared
// memory with bulk-async copies (TMA 1-D, cp.async.bulk + mbarrier complete_tx) while all threads
// stage the tile's CSR slice (row extents, and {column byte offset, value} pairs) with coalesced
// loads; the neighbour gather then runs entirely out of shared memory.  A row is owned by a
// sub-warp lane group (LPR lanes x VEC floats cover the feature row), walks its entries with one
// broadcast LDS.64 + one LDS.128 + VEC FFMA per entry (segmented sum, entries in storage order)
// and leaves with 16-byte coalesced stores.  HBM traffic is the algorithmic minimum: X once,
// Y once, CSR once.  The previous "CSR-vector + shuffle" version of this kernel was issue-bound
// (78 % issue-slot utilisation at 48 % of HBM peak, profiles/r01_spmm_v1.txt); this layout cuts
// the instruction count per graph ~5x.
//
// Row kernel (fallback): graphs whose tile does not fit / is not 16-byte aligned, and the B=1
// large-N block-diagonal use (example_model/sparse.py:65-69), split the flat row space over CTAs
// and gather straight from global / L2.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace kgcn {
namespace {

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
__device__ __forceinline__ void vload(float (&r)[VEC], const void* p) {
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
    int64_t n_graphs;
    int channels, n_rows, n_cols, feat;
    int act;
    int graphs_per_cta;  // tile kernel
    int cv_cap;          // tile kernel: staged {col,val} capacity (entries)
    int lpr_log2;        // row kernel
};

constexpr int kTileThreads = 256;

// ---------------------------------------------------------------------------------------------
// tile kernel
// ---------------------------------------------------------------------------------------------
// Shared-memory loads by 32-bit shared-window address: keeps every address computation in a
// register that was set up once outside the row loops (ptxas otherwise rematerialises them per row).
__device__ __forceinline__ int2 lds_v2(uint32_t addr) {
    int2 r;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
__device__ __forceinline__ int lds_b32(uint32_t addr) {
    int r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
template <int VEC>
__device__ __forceinline__ void lds_vec(float (&r)[VEC], uint32_t addr);
template <>
__device__ __forceinline__ void lds_vec<4>(float (&r)[4], uint32_t addr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_vec<2>(float (&r)[2], uint32_t addr) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_vec<1>(float (&r)[1], uint32_t addr) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[0]) : "r"(addr));
}

// Accumulate CSR entries [s, e) (indices into the staged {byte offset, value} pairs at cv_addr):
// per entry one broadcast LDS.64, one LDS.(32*VEC) and VEC FFMA, in storage order.
template <int VEC>
__device__ __forceinline__ void gather_row(float (&acc)[VEC], uint32_t cv_addr, int s, int e, uint32_t xb_addr) {
    uint32_t a = cv_addr + 8u * static_cast<uint32_t>(s);
    const uint32_t a_end = cv_addr + 8u * static_cast<uint32_t>(e);
#pragma unroll 2
    for (; a < a_end; a += 8) {
        const int2 cv = lds_v2(a);
        float xv[VEC];
        lds_vec<VEC>(xv, xb_addr + static_cast<uint32_t>(cv.x));
        const float v = __int_as_float(cv.y);
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
    }
}

template <int ACT, int VEC>
__device__ __forceinline__ void act_inplace(float (&acc)[VEC]) {
    if (ACT != KGCN_ACT_NONE) {
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc[t] = apply_act(acc[t], ACT);
    }
}

// FLAT rows of one tile: output row w <- CSR row w; rows are contiguous in the output.
template <int VEC, int LPR, int ACT>
__device__ __forceinline__ void flat_rows(uint32_t rp_addr, uint32_t cv_addr, uint32_t xb_addr, int e0, int rows_total,
                                          int group, float* out_w, int feat, const float* self_scale) {
    constexpr int kGroups = 256 / LPR;
    const float eps = self_scale != nullptr ? __ldg(self_scale) : 0.0f;
    const uint32_t row_pitch = static_cast<uint32_t>(feat) * 4u;
    out_w += static_cast<int64_t>(group) * feat;
    const int64_t out_step = static_cast<int64_t>(kGroups) * feat;
    for (int w = group; w < rows_total; w += kGroups, out_w += out_step) {
        const int2 se = make_int2(lds_b32(rp_addr + 4u * w), lds_b32(rp_addr + 4u * w + 4u));
        float acc[VEC];
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc[t] = 0.0f;
        gather_row<VEC>(acc, cv_addr, se.x - e0, se.y - e0, xb_addr);
        if (self_scale != nullptr) {  // GIN: + eps * x[w]  (layers.py:469); C == 1 on this path
            float xv[VEC];
            lds_vec<VEC>(xv, xb_addr + static_cast<uint32_t>(w) * row_pitch);
#pragma unroll
            for (int t = 0; t < VEC; ++t) acc[t] = fmaf(eps, xv[t], acc[t]);
        }
        act_inplace<ACT, VEC>(acc);
        vstore<VEC>(out_w, acc);
    }
}

// FLAT: every output row is produced from exactly one CSR row and output rows of the tile are
// contiguous (C == 1, or per-matrix outputs laid out [B, C, R, F]); otherwise carry counters
// track (graph, out-channel, row) and up to C CSR rows are summed per output row.
template <int VEC, int LPR, bool FLAT>
__global__ void __launch_bounds__(kTileThreads) bspmm_tile_kernel(const SpmmParams p) {
    pdl_prologue();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.channels, n_rows = p.n_rows, feat = p.feat;
    const bool sum_channels = (p.os_c == 0);
    const bool shared_rhs = (p.rs_c == 0);
    const int n_tiles = shared_rhs ? 1 : C;
    const uint32_t tile_bytes = static_cast<uint32_t>(p.n_cols) * feat * 4u;
    const uint32_t graph_bytes = tile_bytes * n_tiles;

    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
    const int ng = static_cast<int>(min(static_cast<int64_t>(p.graphs_per_cta), p.n_graphs - g0));

    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* tile = smem_raw + 128;
    int32_t* rp_s = reinterpret_cast<int32_t*>(tile + static_cast<size_t>(p.graphs_per_cta) * graph_bytes);
    const int rows_total = ng * C * n_rows;
    // {byte offset of the source feature row inside the staged tiles, value bits}; 8-byte aligned
    int2* cv_s = reinterpret_cast<int2*>(rp_s + ((p.graphs_per_cta * C * n_rows + 1 + 1) & ~1));

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, graph_bytes * ng);
        const float* src = p.rhs + g0 * p.rs_g;
        if (shared_rhs && p.rs_g * 4 == tile_bytes) {  // the G tiles are one contiguous block
            bulk_g2s(tile, src, graph_bytes * ng, bar);
        } else {
            for (int gl = 0; gl < ng; ++gl)
                for (int c = 0; c < n_tiles; ++c)
                    bulk_g2s(tile + (static_cast<size_t>(gl) * n_tiles + c) * tile_bytes, src + gl * p.rs_g + c * p.rs_c,
                             tile_bytes, bar);
        }
    }
    const int32_t* rp_g = p.rowptr + g0 * C * n_rows;
    for (int r = threadIdx.x; r <= rows_total; r += kTileThreads) rp_s[r] = __ldg(rp_g + r);
    __syncthreads();  // rowptr slice + barrier init visible
    const int32_t e0 = rp_s[0];
    const int n_entries = rp_s[rows_total] - e0;
    const int row_pitch = feat * 4;
    const bool staged_csr = n_entries <= p.cv_cap;  // else: unusually dense tile, entries are read from global
    if (staged_csr) {
        const int n_mats = ng * n_tiles;                    // staged feature tiles
        const int mat_rows = (shared_rhs ? C : 1) * n_rows;  // CSR rows that read the same tile
        for (int k = threadIdx.x; k < n_entries; k += kTileThreads) {
            int m = 0;
            while (m + 1 < n_mats && e0 + k >= rp_s[(m + 1) * mat_rows]) ++m;
            cv_s[k] = make_int2(__ldg(p.col + e0 + k) * row_pitch + m * static_cast<int>(tile_bytes),
                                __float_as_int(__ldg(p.val + e0 + k)));
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);  // feature tiles have landed

    constexpr int kGroups = kTileThreads / LPR;
    constexpr int kChunk = LPR * VEC;
    const int sub = threadIdx.x & (LPR - 1);
    const int group = threadIdx.x / LPR;
    const uint32_t rp_addr = smem_u32(rp_s), cv_addr = smem_u32(cv_s);

    for (int f0 = sub * VEC; f0 < feat; f0 += kChunk) {  // one pass unless feat > LPR * VEC
        const uint32_t xb_addr = smem_u32(tile) + static_cast<uint32_t>(f0) * 4u;
        if (FLAT && staged_csr) {
            float* out_w = p.out + g0 * p.os_g + f0;  // rows of the tile are contiguous in the output
            switch (p.act) {
                case KGCN_ACT_RELU:
                    flat_rows<VEC, LPR, KGCN_ACT_RELU>(rp_addr, cv_addr, xb_addr, e0, rows_total, group, out_w, feat, p.self_scale);
                    break;
                case KGCN_ACT_SIGMOID:
                    flat_rows<VEC, LPR, KGCN_ACT_SIGMOID>(rp_addr, cv_addr, xb_addr, e0, rows_total, group, out_w, feat, p.self_scale);
                    break;
                case KGCN_ACT_TANH:
                    flat_rows<VEC, LPR, KGCN_ACT_TANH>(rp_addr, cv_addr, xb_addr, e0, rows_total, group, out_w, feat, p.self_scale);
                    break;
                default:
                    flat_rows<VEC, LPR, KGCN_ACT_NONE>(rp_addr, cv_addr, xb_addr, e0, rows_total, group, out_w, feat, p.self_scale);
            }
        } else {
            const unsigned char* xb = tile + f0 * 4;
            const int n_out_ch = sum_channels ? 1 : C;
            int i = group, oc = 0, gl = 0;  // carry counters instead of divisions
            while (i >= n_rows) {
                i -= n_rows;
                if (++oc == n_out_ch) { oc = 0; ++gl; }
            }
            while (gl < ng) {
                const int c_begin = sum_channels ? 0 : oc;
                const int c_end = sum_channels ? C : oc + 1;
                float acc[VEC];
#pragma unroll
                for (int t = 0; t < VEC; ++t) acc[t] = 0.0f;
                for (int c = c_begin; c < c_end; ++c) {
                    const int r = (gl * C + c) * n_rows + i;
                    const int s = rp_s[r] - e0, e = rp_s[r + 1] - e0;
                    const int tile_off = (shared_rhs ? gl : gl * C + c) * static_cast<int>(tile_bytes);
                    if (staged_csr) {
                        gather_row<VEC>(acc, cv_addr, s, e, xb_addr);
                    } else {
                        for (int k = s; k < e; ++k) {
                            const int off = __ldg(p.col + e0 + k) * row_pitch + tile_off;
                            const float v = __ldg(p.val + e0 + k);
                            float xv[VEC];
                            vload<VEC>(xv, xb + off);
#pragma unroll
                            for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                        }
                    }
                    if (p.self_scale != nullptr) {
                        const float eps = __ldg(p.self_scale + c);
                        float xv[VEC];
                        vload<VEC>(xv, xb + tile_off + i * row_pitch);
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = fmaf(eps, xv[t], acc[t]);
                    }
                }
                if (p.act != KGCN_ACT_NONE) {
#pragma unroll
                    for (int t = 0; t < VEC; ++t) acc[t] = apply_act(acc[t], p.act);
                }
                vstore<VEC>(p.out + (g0 + gl) * p.os_g + oc * p.os_c + static_cast<int64_t>(i) * feat + f0, acc);
                i += kGroups;
                while (i >= n_rows) {
                    i -= n_rows;
                    if (++oc == n_out_ch) { oc = 0; ++gl; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// row kernel (fallback): flat output rows split over CTAs, gather from global memory
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 128;

template <int VEC>
__global__ void __launch_bounds__(kRowThreads) bspmm_row_kernel(const SpmmParams p, int64_t total_out_rows) {
    pdl_prologue();
    const int C = p.channels, n_rows = p.n_rows, feat = p.feat;
    const bool sum_channels = (p.os_c == 0);
    const int lpr = 1 << p.lpr_log2;
    const int sub = threadIdx.x & (lpr - 1);
    const int64_t q = (static_cast<int64_t>(blockIdx.x) * kRowThreads + threadIdx.x) >> p.lpr_log2;
    if (q >= total_out_rows) return;
    const int n_out_ch = sum_channels ? 1 : C;
    const int i = static_cast<int>(q % n_rows);
    const int64_t t1 = q / n_rows;
    const int oc = static_cast<int>(t1 % n_out_ch);
    const int64_t g = t1 / n_out_ch;
    const int c_begin = sum_channels ? 0 : oc, c_end = sum_channels ? C : oc + 1;
    const int chunk = lpr * VEC;
    float* out_row = p.out + g * p.os_g + oc * p.os_c + static_cast<int64_t>(i) * feat;
    for (int f0 = sub * VEC; f0 < feat; f0 += chunk) {
        float acc[VEC];
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc[t] = 0.0f;
        for (int c = c_begin; c < c_end; ++c) {
            const int64_t r = (g * C + c) * n_rows + i;
            const int32_t s = __ldg(p.rowptr + r), e = __ldg(p.rowptr + r + 1);
            const float* x_c = p.rhs + g * p.rs_g + c * p.rs_c + f0;
            for (int32_t k = s; k < e; ++k) {
                const int j = __ldg(p.col + k);
                const float v = __ldg(p.val + k);
                float xv[VEC];
                vload<VEC>(xv, x_c + static_cast<int64_t>(j) * feat);
#pragma unroll
                for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
            }
            if (p.self_scale != nullptr) {
                const float eps = __ldg(p.self_scale + c);
                float xv[VEC];
                vload<VEC>(xv, x_c + static_cast<int64_t>(i) * feat);
#pragma unroll
                for (int t = 0; t < VEC; ++t) acc[t] = fmaf(eps, xv[t], acc[t]);
            }
        }
        if (p.act != KGCN_ACT_NONE) {
#pragma unroll
            for (int t = 0; t < VEC; ++t) acc[t] = apply_act(acc[t], p.act);
        }
        vstore<VEC>(out_row + f0, acc);
    }
}

// dval[e] = < dy[row_e, :], rhs[col_e, :] >  (bspmm_call.py:49-54); one warp per CSR row.
__global__ void __launch_bounds__(128) bspmm_dvalues_kernel(const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ col,
                                                            const int32_t* __restrict__ perm, int64_t total_rows,
                                                            int channels, int n_rows, int feat,
                                                            const float* __restrict__ dy, int64_t ds_g, int64_t ds_c,
                                                            const float* __restrict__ rhs, int64_t rs_g, int64_t rs_c,
                                                            float* __restrict__ dval) {
    pdl_prologue();
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

template <int VEC, int LPR, bool FLAT>
int launch_tile(const SpmmParams& p, unsigned grid, size_t smem, cudaStream_t st) {
    static bool attr_set = false;  // idempotent; racing threads set the same value
    if (!attr_set) {
        KGCN_CUDA_OK(cudaFuncSetAttribute(bspmm_tile_kernel<VEC, LPR, FLAT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    launch_pdl(bspmm_tile_kernel<VEC, LPR, FLAT>, grid, kTileThreads, smem, st, p);
    KGCN_LAUNCH_OK("bspmm_tile_kernel");
    return KGCN_OK;
}

template <int VEC, bool FLAT>
int launch_tile_vec(const SpmmParams& p, int lpr, unsigned grid, size_t smem, cudaStream_t st) {
    switch (lpr) {
        case 1: return launch_tile<VEC, 1, FLAT>(p, grid, smem, st);
        case 2: return launch_tile<VEC, 2, FLAT>(p, grid, smem, st);
        case 4: return launch_tile<VEC, 4, FLAT>(p, grid, smem, st);
        case 8: return launch_tile<VEC, 8, FLAT>(p, grid, smem, st);
        case 16: return launch_tile<VEC, 16, FLAT>(p, grid, smem, st);
        default: return launch_tile<VEC, 32, FLAT>(p, grid, smem, st);
    }
}

}  // namespace

int launch_bspmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs, int channels,
                 int n_rows, int n_cols, int feat, const float* rhs, int64_t rs_g, int64_t rs_c, float* out,
                 int64_t os_g, int64_t os_c, const float* self_scale, int act, cudaStream_t st) {
    KGCN_REQUIRE(rowptr && col && val && rhs && out, KGCN_ERR_NULL, "bspmm: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_rows > 0 && n_cols > 0 && feat > 0, KGCN_ERR_BAD_SHAPE,
                 "bspmm: bad shape n_graphs=%lld channels=%d n_rows=%d n_cols=%d feat=%d", (long long)n_graphs,
                 channels, n_rows, n_cols, feat);
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

    SpmmParams p{};
    p.rowptr = rowptr; p.col = col; p.val = val; p.rhs = rhs; p.out = out; p.self_scale = self_scale;
    p.rs_g = rs_g; p.rs_c = rs_c; p.os_g = os_g; p.os_c = os_c; p.n_graphs = n_graphs;
    p.channels = channels; p.n_rows = n_rows; p.n_cols = n_cols; p.feat = feat; p.act = act; p.lpr_log2 = lpr_log2;

    const size_t tile_bytes = static_cast<size_t>(n_cols) * feat * 4;
    const size_t graph_bytes = tile_bytes * ((rs_c == 0) ? 1 : channels);
    const int64_t rows_per_graph = static_cast<int64_t>(channels) * n_rows;
    const bool stageable = tile_bytes % 16 == 0 && aligned16(rhs) && (rs_g * 4) % 16 == 0 && (rs_c * 4) % 16 == 0 &&
                           graph_bytes <= 96 * 1024 && rows_per_graph <= 8192;
    if (stageable) {
        // graphs per CTA: ~16 KB of features (measured best at C2: 2 graphs of 32x64 -> 6.1 us at B=1024,
        // 83 % of the measured HBM peak at B=16384; 1, 3, 4 and 8 graphs per CTA are all slower)
        int64_t G = std::max<int64_t>(1, std::min<int64_t>(8, (16 * 1024) / static_cast<int64_t>(graph_bytes)));
        static const char* force_g = getenv("KGCN_SPMM_G");   // tuning knob
        if (force_g != nullptr && atoi(force_g) > 0) G = std::min<int64_t>(atoi(force_g), (96 * 1024) / static_cast<int64_t>(graph_bytes));
        const int64_t cap = std::max<int64_t>(256, 6 * G * rows_per_graph);
        const size_t smem = 128 + G * graph_bytes + ((G * rows_per_graph + 2) & ~1ll) * 4 + cap * 8;
        const int64_t grid = ceil_div<int64_t>(n_graphs, G);
        if (smem <= 200 * 1024 && grid < (1ll << 31) && G * graph_bytes < (1u << 20)) {
            p.graphs_per_cta = static_cast<int>(G);
            p.cv_cap = static_cast<int>(cap);
            const int lpr = 1 << lpr_log2;
            const unsigned ug = static_cast<unsigned>(grid);
            const int64_t mat_elems = static_cast<int64_t>(n_rows) * feat;
            // one CSR row per output row and the tile's output rows are contiguous
            const bool flat = (channels == 1 && os_g == mat_elems && (self_scale == nullptr || rs_g == mat_elems)) ||
                              (channels > 1 && os_c == mat_elems && os_g == channels * mat_elems && self_scale == nullptr);
            if (flat) {
                switch (vec) {
                    case 4: return launch_tile_vec<4, true>(p, lpr, ug, smem, st);
                    case 2: return launch_tile_vec<2, true>(p, lpr, ug, smem, st);
                    default: return launch_tile_vec<1, true>(p, lpr, ug, smem, st);
                }
            }
            switch (vec) {
                case 4: return launch_tile_vec<4, false>(p, lpr, ug, smem, st);
                case 2: return launch_tile_vec<2, false>(p, lpr, ug, smem, st);
                default: return launch_tile_vec<1, false>(p, lpr, ug, smem, st);
            }
        }
    }
    const int64_t total_out_rows = n_graphs * (os_c == 0 ? 1 : channels) * n_rows;
    const int64_t blocks = ceil_div<int64_t>(total_out_rows << lpr_log2, kRowThreads);
    KGCN_REQUIRE(blocks < (1ll << 31), KGCN_ERR_BAD_SHAPE, "bspmm: too many rows for one launch");
    switch (vec) {
        case 4: launch_pdl(bspmm_row_kernel<4>, static_cast<unsigned>(blocks), kRowThreads, 0, st, p, total_out_rows); break;
        case 2: launch_pdl(bspmm_row_kernel<2>, static_cast<unsigned>(blocks), kRowThreads, 0, st, p, total_out_rows); break;
        default: launch_pdl(bspmm_row_kernel<1>, static_cast<unsigned>(blocks), kRowThreads, 0, st, p, total_out_rows); break;
    }
    KGCN_LAUNCH_OK("bspmm_row_kernel");
    return KGCN_OK;
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
    launch_pdl(bspmm_dvalues_kernel, static_cast<unsigned>(blocks), 128, 0, static_cast<cudaStream_t>(stream), 
        rowptr, col, perm, total_rows, channels, n_rows, feat, dy, dy_stride_g, dy_stride_c, rhs, rhs_stride_g,
        rhs_stride_c, dval);
    KGCN_LAUNCH_OK("bspmm_dvalues_kernel");
    return KGCN_OK;
}
