// Exact-fp32 (FFMA) dense helpers for sm_100a.
//
// These are the general-shape building blocks behind the decomposed ("W-first", reference-order)
// GraphConv / GraphDense paths and every backward: a tiled SGEMM with fused bias / activation /
// enabled-node masking, a deterministic split-K variant for the long reductions
// dW = X^T . G (K = B*N rows) that also emits colsum(G) = dbias, and the small elementwise /
// reduction kernels (activation gradient, GraphGather forward / backward).  The fused
// tensor-core layer kernel lives in graphconv_fused.cu; this file is what it is checked against
// and what odd shapes fall back to.
#include <algorithm>

#include "common.cuh"

namespace kgcn {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, GEMM_THREADS = 256;

struct SgemmParams {
    const float* A;
    const float* B;
    float* C;
    int64_t lda, ldb, ldc;
    int64_t M;  // rows of op(A) / C
    int N;      // cols of op(B) / C
    int64_t K;  // reduction length
    int64_t k_chunk;  // split-K: reduction elements per blockIdx.z (multiple of BK)
    float* partial;   // split-K workspace [splits][M + colsum][N], or null
    int colsum;       // split-K only: also emit column sums of op(B) as row M of the partials
    const float* bias;
    const int32_t* enabled;
    int n_nodes;
    int act;
    int accumulate;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS) sgemm_kernel(const SgemmParams p) {
    pdl_prologue();
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
    const int n0 = blockIdx.x * BN;
    const int64_t k_begin = static_cast<int64_t>(blockIdx.z) * p.k_chunk;
    const int64_t k_end = min(p.K, k_begin + p.k_chunk);
    const bool do_colsum = p.colsum && blockIdx.y == 0 && ty == 0;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    float bsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};

    // Register-staged double buffering: the global loads of tile k0+BK are issued before the FFMAs of
    // tile k0, so their latency overlaps the math (these GEMMs run at ~1 CTA per SM).
    constexpr int kLdA = (BM * BK) / GEMM_THREADS, kLdB = (BN * BK) / GEMM_THREADS;
    float ra[kLdA], rb[kLdB];
    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int it = 0; it < kLdA; ++it) {
            const int idx = t + it * GEMM_THREADS;
            const int m = TA ? idx % BM : idx / BK, k = TA ? idx / BM : idx % BK;
            const int64_t gm = m0 + m, gk = k0 + k;
            ra[it] = (gm < p.M && gk < k_end) ? (TA ? p.A[gk * p.lda + gm] : p.A[gm * p.lda + gk]) : 0.0f;
        }
#pragma unroll
        for (int it = 0; it < kLdB; ++it) {
            const int idx = t + it * GEMM_THREADS;
            const int n = TB ? idx / BK : idx % BN, k = TB ? idx % BK : idx / BN;
            const int gn = n0 + n;
            const int64_t gk = k0 + k;
            rb[it] = (gn < p.N && gk < k_end) ? (TB ? p.B[static_cast<int64_t>(gn) * p.ldb + gk] : p.B[gk * p.ldb + gn]) : 0.0f;
        }
    };
    if (k_begin < k_end) fetch(k_begin);
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int it = 0; it < kLdA; ++it) {
            const int idx = t + it * GEMM_THREADS;
            As[TA ? idx / BM : idx % BK][TA ? idx % BM : idx / BK] = ra[it];
        }
#pragma unroll
        for (int it = 0; it < kLdB; ++it) {
            const int idx = t + it * GEMM_THREADS;
            Bs[TB ? idx % BK : idx / BN][TB ? idx / BK : idx % BN] = rb[it];
        }
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            if (do_colsum) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
            }
        }
        __syncthreads();
    }

    if (p.partial != nullptr) {  // split-K: raw partial sums, reduced by splitk_reduce_kernel
        const int64_t rows = p.M + (p.colsum ? 1 : 0);
        float* part = p.partial + static_cast<int64_t>(blockIdx.z) * rows * p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t gm = m0 + ty * 4 + i;
            if (gm >= p.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + tx * 4 + j;
                if (gn < p.N) part[gm * p.N + gn] = acc[i][j];
            }
        }
        if (do_colsum) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + tx * 4 + j;
                if (gn < p.N) part[p.M * p.N + gn] = bsum[j];
            }
        }
        return;
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= p.M) continue;
        bool masked = false;
        if (p.enabled != nullptr) masked = (gm % p.n_nodes) >= p.enabled[gm / p.n_nodes];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= p.N) continue;
            float v = acc[i][j];
            if (p.accumulate) {
                v += p.C[gm * p.ldc + gn];
            } else {
                if (p.bias != nullptr) v += p.bias[gn];
                v = apply_act(v, p.act);
                if (masked) v = 0.0f;
            }
            p.C[gm * p.ldc + gn] = v;
        }
    }
}

// out[r, n] = sum_z partial[z][r][n] (deterministic order); row M -> colsum output.
// Block = 32 consecutive outputs (one per lane, coalesced) x 32 warps; warp w adds the partials z = w, w + 32, ...
// in that order with up to eight independent loads in flight, then warp 0 adds the 32 warp sums in warp order.  The
// reduction is a latency chain (a partial block is a few KB), so the width matters, not the bytes.
// `n_per_ch` < N: the N columns are C = N / n_per_ch channel blocks and `out` is channel-major [C][M][n_per_ch]
// (dW of a multi-channel GraphConv: the partial blocks hold X^T.[G_0 | G_1 | ..]).
__global__ void __launch_bounds__(1024) splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t M, int N,
                                                             int has_colsum, float* __restrict__ out,
                                                             float* __restrict__ colsum_out, int n_per_ch) {
    pdl_prologue();
    __shared__ float red[32][33];
    const int64_t rows = M + (has_colsum ? 1 : 0);
    const int64_t total = rows * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    float s = 0.0f;
    if (idx < total) {
        const float* p = partial + idx;
        for (int z = warp; z < splits; z += 256) {   // predicated batches of 8: up to 256 partials in ONE round trip
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (z + 32 * j < splits) ? p[static_cast<int64_t>(z + 32 * j) * total] : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (z + 32 * j < splits) s += v[j];
        }
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && idx < total) {
        s = 0.0f;
#pragma unroll
        for (int w = 0; w < 32; ++w) s += red[w][lane];
        if (idx < M * N) {
            if (n_per_ch == N) {
                out[idx] = s;
            } else {
                const int64_t m = idx / N;
                const int cn = static_cast<int>(idx - m * N), c = cn / n_per_ch;
                out[(static_cast<int64_t>(c) * M + m) * n_per_ch + (cn - c * n_per_ch)] = s;
            }
        } else if (colsum_out != nullptr) {
            colsum_out[idx - M * N] = s;
        }
    }
}

// dy_bcast != 0: dy is [n_graphs, feat] and is broadcast over the n_nodes rows of each graph (the
// gradient of GraphGather, layers.py:164) -- fuses gather_bwd into this pass.
__global__ void act_grad_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ du,
                                int64_t n, int feat, int act, const int32_t* __restrict__ enabled, int n_nodes,
                                int dy_bcast) {
    pdl_prologue();
    const int64_t per_graph = static_cast<int64_t>(n_nodes) * feat;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float g = dy_bcast ? dy[(idx / per_graph) * feat + (idx % feat)] : dy[idx];
        float v = g * (y != nullptr ? act_grad_from_output(y[idx], act) : 1.0f);
        if (enabled != nullptr) {
            const int64_t row = idx / feat;
            if ((row % n_nodes) >= enabled[row / n_nodes]) v = 0.0f;
        }
        du[idx] = v;
    }
}

// GraphGather forward: out[g, f] = sum_i x[g, i, f], rows added in index order (layers.py:164).
__global__ void gather_fwd_kernel(const float* __restrict__ x, int64_t n_graphs, int n_nodes, int feat,
                                  float* __restrict__ out) {
    pdl_prologue();
    const int64_t total = n_graphs * feat;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t g = idx / feat;
        const int f = static_cast<int>(idx % feat);
        const float* p = x + g * n_nodes * feat + f;
        float s = 0.0f;
        for (int i = 0; i < n_nodes; ++i) s += p[static_cast<int64_t>(i) * feat];
        out[idx] = s;
    }
}

__global__ void gather_bwd_kernel(const float* __restrict__ dout, int64_t n_graphs, int n_nodes, int feat,
                                  float* __restrict__ dx) {
    pdl_prologue();
    const int64_t total = n_graphs * n_nodes * feat;
    const int64_t per_graph = static_cast<int64_t>(n_nodes) * feat;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        dx[idx] = dout[(idx / per_graph) * feat + (idx % feat)];
    }
}

inline unsigned ew_blocks(int64_t n) {
    const int64_t b = ceil_div<int64_t>(n, 256);
    return static_cast<unsigned>(b < 1 ? 1 : (b > kNumSMs * 16 ? kNumSMs * 16 : b));
}

int choose_splits(int64_t M_out, int N, int64_t K) {
    const int64_t tiles = ceil_div<int64_t>(M_out, BM) * ceil_div(N, BN);
    int64_t splits = ceil_div<int64_t>(kNumSMs, tiles);   // one wave (2 CTAs/SM measured slower: more partials to reduce)
    const int64_t max_by_k = ceil_div<int64_t>(K, 4 * BK);  // at least 4 k-steps per split
    if (splits > max_by_k) splits = max_by_k;
    if (splits < 1) splits = 1;
    return static_cast<int>(splits);
}

}  // namespace

int launch_sgemm(bool trans_a, bool trans_b, int64_t M, int N, int K, const float* A, int64_t lda, const float* B,
                 int64_t ldb, float* C, int64_t ldc, const GemmEpilogue& ep, cudaStream_t st) {
    KGCN_REQUIRE(A && B && C, KGCN_ERR_NULL, "sgemm: NULL pointer argument");
    KGCN_REQUIRE(M >= 0 && N > 0 && K > 0, KGCN_ERR_BAD_SHAPE, "sgemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    if (M == 0) return KGCN_OK;
    SgemmParams p{};
    p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.k_chunk = ceil_div<int64_t>(K, BK) * BK;
    p.partial = nullptr; p.colsum = 0;
    p.bias = ep.bias; p.enabled = ep.enabled; p.n_nodes = ep.n_nodes; p.act = ep.act; p.accumulate = ep.accumulate;
    const int64_t gy = ceil_div<int64_t>(M, BM);
    KGCN_REQUIRE(gy < 65536 * 32768ll, KGCN_ERR_BAD_SHAPE, "sgemm: M too large");
    // gridDim.y is limited to 65535: fold very tall problems by launching slabs
    const int64_t max_gy = 65535;
    for (int64_t y0 = 0; y0 < gy; y0 += max_gy) {
        const int64_t rows_here = std::min<int64_t>(max_gy, gy - y0);
        SgemmParams q = p;
        const int64_t row_off = y0 * BM;
        q.A = trans_a ? A + row_off : A + row_off * lda;
        q.C = C + row_off * ldc;
        q.M = std::min<int64_t>(M - row_off, rows_here * BM);
        if (ep.enabled != nullptr) {
            KGCN_REQUIRE(row_off % ep.n_nodes == 0 || gy <= max_gy, KGCN_ERR_UNSUPPORTED, "sgemm: masked slab split");
            q.enabled = ep.enabled + row_off / ep.n_nodes;
        }
        dim3 grid(ceil_div(N, BN), static_cast<unsigned>(rows_here), 1);
        if (trans_a && trans_b) launch_pdl(sgemm_kernel<true, true>, grid, GEMM_THREADS, 0, st, q);
        else if (trans_a) launch_pdl(sgemm_kernel<true, false>, grid, GEMM_THREADS, 0, st, q);
        else if (trans_b) launch_pdl(sgemm_kernel<false, true>, grid, GEMM_THREADS, 0, st, q);
        else launch_pdl(sgemm_kernel<false, false>, grid, GEMM_THREADS, 0, st, q);
        KGCN_LAUNCH_OK("sgemm_kernel");
    }
    return KGCN_OK;
}

int launch_splitk_reduce(const float* partial, int splits, int64_t M, int N, float* out, float* colsum_out,
                         cudaStream_t st) {
    launch_pdl(splitk_reduce_kernel, static_cast<unsigned>(ceil_div<int64_t>((M + 1) * N, 32)), 1024, 0, st, partial, splits,
               M, N, 1, out, colsum_out, N);
    KGCN_LAUNCH_OK("splitk_reduce_kernel");
    return KGCN_OK;
}

int launch_splitk_reduce_ch(const float* partial, int splits, int64_t M, int n_per_ch, int channels, float* out,
                            float* colsum_out, cudaStream_t st) {
    const int N = n_per_ch * channels;
    launch_pdl(splitk_reduce_kernel, static_cast<unsigned>(ceil_div<int64_t>((M + 1) * N, 32)), 1024, 0, st, partial, splits,
               M, N, 1, out, colsum_out, n_per_ch);
    KGCN_LAUNCH_OK("splitk_reduce_kernel");
    return KGCN_OK;
}

size_t reduce_gemm_workspace_bytes(int64_t M, int Ka, int N) {
    const int splits = choose_splits(Ka, N, M);
    return static_cast<size_t>(splits) * (static_cast<size_t>(Ka) + 1) * N * sizeof(float);
}

int launch_reduce_gemm_tn(int64_t M, int Ka, int N, const float* A, int64_t lda, const float* B, int64_t ldb,
                          float* out, float* colsum_b, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    KGCN_REQUIRE(A && B && out, KGCN_ERR_NULL, "reduce_gemm: NULL pointer argument");
    KGCN_REQUIRE(M > 0 && Ka > 0 && N > 0, KGCN_ERR_BAD_SHAPE, "reduce_gemm: bad shape");
    const int splits = choose_splits(Ka, N, M);
    const size_t need = static_cast<size_t>(splits) * (static_cast<size_t>(Ka) + 1) * N * sizeof(float);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= need, KGCN_ERR_WORKSPACE,
                 "reduce_gemm: workspace %zu < %zu bytes", workspace_bytes, need);
    SgemmParams p{};
    p.A = A; p.B = B; p.C = nullptr; p.lda = lda; p.ldb = ldb; p.ldc = 0;
    p.M = Ka; p.N = N; p.K = M;
    p.k_chunk = ceil_div<int64_t>(ceil_div<int64_t>(M, splits), BK) * BK;
    p.partial = static_cast<float*>(workspace);
    p.colsum = 1;
    dim3 grid(ceil_div(N, BN), ceil_div(Ka, BM), splits);
    launch_pdl(sgemm_kernel<true, false>, grid, GEMM_THREADS, 0, st, p);
    KGCN_LAUNCH_OK("sgemm_kernel(split-K)");
    launch_pdl(splitk_reduce_kernel, static_cast<unsigned>(ceil_div<int64_t>((static_cast<int64_t>(Ka) + 1) * N, 32)), 1024, 0, st, 
        static_cast<const float*>(workspace), splits, Ka, N, 1, out, colsum_b, N);
    KGCN_LAUNCH_OK("splitk_reduce_kernel");
    return KGCN_OK;
}

// 16-byte vectorised variant (feat % 4 == 0, aligned pointers, no row mask)
__global__ void act_grad_vec4_kernel(const float4* __restrict__ y, const float4* __restrict__ dy, float4* __restrict__ du,
                                     int64_t n4, int feat4, int act, int n_nodes, int dy_bcast) {
    pdl_prologue();
    const int64_t per_graph4 = static_cast<int64_t>(n_nodes) * feat4;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n4;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 g = dy_bcast ? dy[(idx / per_graph4) * feat4 + (idx % feat4)] : dy[idx];
        float4 o = g;
        if (y != nullptr) {
            const float4 yy = y[idx];
            o.x *= act_grad_from_output(yy.x, act);
            o.y *= act_grad_from_output(yy.y, act);
            o.z *= act_grad_from_output(yy.z, act);
            o.w *= act_grad_from_output(yy.w, act);
        }
        du[idx] = o;
    }
}

int launch_act_grad(const float* y, const float* dy, float* du, int64_t n, int feat, int act, const int32_t* enabled,
                    int n_nodes, bool dy_bcast, cudaStream_t st) {
    if (n == 0) return KGCN_OK;
    if (enabled == nullptr && feat % 4 == 0 && aligned16(dy) && aligned16(du) && (y == nullptr || aligned16(y))) {
        const int64_t n4 = n / 4;
        const int64_t blocks = std::min<int64_t>(ceil_div<int64_t>(n4, 256), kNumSMs * 8);
        launch_pdl(act_grad_vec4_kernel, static_cast<unsigned>(blocks), 256, 0, st, 
            reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(du), n4,
            feat / 4, act, n_nodes, dy_bcast ? 1 : 0);
        KGCN_LAUNCH_OK("act_grad_vec4_kernel");
        return KGCN_OK;
    }
    launch_pdl(act_grad_kernel, ew_blocks(n), 256, 0, st, y, dy, du, n, feat, act, enabled, n_nodes, dy_bcast ? 1 : 0);
    KGCN_LAUNCH_OK("act_grad_kernel");
    return KGCN_OK;
}

}  // namespace kgcn

using namespace kgcn;

extern "C" int kgcn_gather_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* out,
                                   void* stream) {
    KGCN_REQUIRE(x && out, KGCN_ERR_NULL, "gather_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "gather_fwd: bad shape");
    if (n_graphs == 0) return KGCN_OK;
    launch_pdl(gather_fwd_kernel, ew_blocks(n_graphs * feat), 256, 0, static_cast<cudaStream_t>(stream), x, n_graphs, n_nodes,
                                                                                                  feat, out);
    KGCN_LAUNCH_OK("gather_fwd_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_gather_bwd_f32(const float* dout, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* dx,
                                   void* stream) {
    KGCN_REQUIRE(dout && dx, KGCN_ERR_NULL, "gather_bwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "gather_bwd: bad shape");
    if (n_graphs == 0) return KGCN_OK;
    launch_pdl(gather_bwd_kernel, ew_blocks(n_graphs * n_nodes * feat), 256, 0, static_cast<cudaStream_t>(stream), 
        dout, n_graphs, n_nodes, feat, dx);
    KGCN_LAUNCH_OK("gather_bwd_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_graphdense_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t f_in,
                                       const float* kernel, const float* bias, int32_t f_out, int32_t act,
                                       const int32_t* enabled_node_nums, float* y, void* stream) {
    KGCN_REQUIRE(x && kernel && y, KGCN_ERR_NULL, "graphdense_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && f_in > 0 && f_out > 0, KGCN_ERR_BAD_SHAPE, "graphdense_fwd: bad shape");
    GemmEpilogue ep;
    ep.bias = bias; ep.act = act; ep.enabled = enabled_node_nums; ep.n_nodes = n_nodes;
    return launch_sgemm(false, false, n_graphs * n_nodes, f_out, f_in, x, f_in, kernel, f_out, y, f_out, ep,
                        static_cast<cudaStream_t>(stream));
}

extern "C" size_t kgcn_graphdense_workspace_bytes(int64_t n_graphs, int32_t n_nodes, int32_t f_in, int32_t f_out) {
    const int64_t rows = n_graphs * n_nodes;
    if (rows <= 0) return 0;
    return static_cast<size_t>(rows) * f_out * sizeof(float) + reduce_gemm_workspace_bytes(rows, f_in, f_out);
}

extern "C" int kgcn_graphdense_bwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t f_in,
                                       const float* kernel, int32_t f_out, int32_t act,
                                       const int32_t* enabled_node_nums, const float* y, const float* dy, float* dx,
                                       float* dkernel, float* dbias, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    KGCN_REQUIRE(x && kernel && y && dy && dkernel, KGCN_ERR_NULL, "graphdense_bwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && f_in > 0 && f_out > 0, KGCN_ERR_BAD_SHAPE, "graphdense_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t rows = n_graphs * n_nodes;
    if (rows == 0) {
        KGCN_CUDA_OK(cudaMemsetAsync(dkernel, 0, sizeof(float) * f_in * f_out, st));
        if (dbias) KGCN_CUDA_OK(cudaMemsetAsync(dbias, 0, sizeof(float) * f_out, st));
        return KGCN_OK;
    }
    KGCN_REQUIRE(workspace && workspace_bytes >= kgcn_graphdense_workspace_bytes(n_graphs, n_nodes, f_in, f_out),
                 KGCN_ERR_WORKSPACE, "graphdense_bwd: workspace too small");
    float* du = static_cast<float*>(workspace);
    void* ws2 = du + rows * f_out;
    const size_t ws2_bytes = workspace_bytes - static_cast<size_t>(rows) * f_out * sizeof(float);
    int rc = launch_act_grad(y, dy, du, rows * f_out, f_out, act, enabled_node_nums, n_nodes, false, st);
    if (rc) return rc;
    rc = launch_reduce_gemm_tn(rows, f_in, f_out, x, f_in, du, f_out, dkernel, dbias, ws2, ws2_bytes, st);
    if (rc) return rc;
    if (dx != nullptr) {
        GemmEpilogue ep;
        rc = launch_sgemm(false, true, rows, f_in, f_out, du, f_out, kernel, f_out, dx, f_in, ep, st);
    }
    return rc;
}
