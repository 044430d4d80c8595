// GraphMaxPooling, the block-diagonal (per-molecule segment) readout and GraphBatchNormalization for sm_100a.
//
// GraphMaxPooling (kgcn/layers.py:122-153).  Per molecule b, channel c, feature k the reference builds the
// sparse matrix  A[b][c] * x[b, :, k]  (column j scaled by x[b, j, k]), densifies it with
// tf.sparse_tensor_to_dense (absent entries become 0) and takes tf.reduce_max over axis 1; channels are then
// summed with tf.add_n:
//     y[b, i, k] = sum_c  max( { A_c[i, j] * x[b, j, k] : (i, j) stored }  U  { 0 if row i stores < n_cols entries } )
// The reference runs n_graphs * channels * feat separate TF op chains for this; here it is one launch over the
// BatchedCSR.  Gradient (TF: reduce_max spreads the incoming gradient EVENLY over all maximal positions of the
// densified row, implicit zeros included; the shares of absent positions are dropped by the gather of
// sparse_tensor_to_dense's gradient):
//     dx[b, j, k] = sum_c sum_{(i, j) stored} A_c[i, j] * dy[b, i, k] * [A_c[i, j] * x[b, j, k] == m_c[b, i, k]] / n_max_c[b, i, k]
// The forward stores m_c and n_max_c when asked to, the backward is a gather over the TRANSPOSED BatchedCSR:
// deterministic, no atomics.  Stored duplicates (which tf.sparse_tensor_to_dense rejects) count as separate
// candidates.
//
// Segment sum (example_model/sparse.py:79-90, the tf.scan over molecules of the block-diagonal model):
//     out[m, :] = sum of rows start[m] .. start[m] + size[m] - 1.
//
// GraphBatchNormalization (kgcn/layers.py:170-220 / kgcn/legacy/layers.py:170-218): per-feature affine
// normalisation of the first enabled_node_nums[b] rows of every molecule, the remaining rows are written as
// exact zeros (layers.py:207-213).  mode 0: moving statistics (what the Keras layer does under the reference
// trainer); mode 1: batch statistics over the enabled rows (legacy tf.layers variant,
// training=True), biased variance, also returned for the moving-average update.
#include <algorithm>

#include "common.cuh"

namespace kgcn {
namespace {

template <int VEC>
struct Vec {
    float v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> ldv(const float* p) {
    Vec<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) r.v[i] = p[i];
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const Vec<VEC>& r) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = r.v[i];
    }
}

// one thread per (graph, row, VEC features)
template <int VEC>
__global__ void maxpool_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   const float* __restrict__ val, int64_t n_graphs, int C, int N, int F,
                                   const float* __restrict__ x, float* __restrict__ y, float* __restrict__ chmax,
                                   float* __restrict__ nmax) {
    pdl_prologue();
    const int fq = F / VEC;
    const int64_t total = n_graphs * N * fq;
    const int64_t ch_stride = n_graphs * N * static_cast<int64_t>(F);
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int f0 = static_cast<int>(idx % fq) * VEC;
        const int64_t gi = idx / fq;   // graph * N + row
        const int64_t g = gi / N;
        const int i = static_cast<int>(gi - g * N);
        const float* xg = x + g * N * F + f0;
        Vec<VEC> out;
#pragma unroll
        for (int t = 0; t < VEC; ++t) out.v[t] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const int64_t r = (g * C + c) * N + i;
            const int s = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
            const bool has_zero = (e - s) < N;   // the densified row keeps at least one implicit 0
            Vec<VEC> m, cnt;
#pragma unroll
            for (int t = 0; t < VEC; ++t) {
                m.v[t] = has_zero ? 0.0f : -INFINITY;
                cnt.v[t] = has_zero ? static_cast<float>(N - (e - s)) : 0.0f;
            }
            for (int k = s; k < e; ++k) {
                const float a = __ldg(val + k);
                const Vec<VEC> xv = ldv<VEC>(xg + static_cast<int64_t>(__ldg(col + k)) * F);
#pragma unroll
                for (int t = 0; t < VEC; ++t) {
                    const float pr = __fmul_rn(a, xv.v[t]);
                    if (pr > m.v[t]) {
                        m.v[t] = pr;
                        cnt.v[t] = 1.0f;
                    } else if (pr == m.v[t]) {
                        cnt.v[t] += 1.0f;
                    }
                }
            }
            if (chmax != nullptr) {
                stv<VEC>(chmax + c * ch_stride + gi * F + f0, m);
                stv<VEC>(nmax + c * ch_stride + gi * F + f0, cnt);
            }
#pragma unroll
            for (int t = 0; t < VEC; ++t) out.v[t] += m.v[t];
        }
        stv<VEC>(y + gi * F + f0, out);
    }
}

// one thread per (graph, column j, VEC features); rowptr_t / col_t / val_t = transposed BatchedCSR
template <int VEC>
__global__ void maxpool_bwd_kernel(const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ col_t,
                                   const float* __restrict__ val_t, int64_t n_graphs, int C, int N, int F,
                                   const float* __restrict__ x, const float* __restrict__ dy,
                                   const float* __restrict__ chmax, const float* __restrict__ nmax, float* __restrict__ dx) {
    pdl_prologue();
    const int fq = F / VEC;
    const int64_t total = n_graphs * N * fq;
    const int64_t ch_stride = n_graphs * N * static_cast<int64_t>(F);
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int f0 = static_cast<int>(idx % fq) * VEC;
        const int64_t gj = idx / fq;
        const int64_t g = gj / N;
        const int j = static_cast<int>(gj - g * N);
        const Vec<VEC> xv = ldv<VEC>(x + gj * F + f0);
        Vec<VEC> acc;
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc.v[t] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const int64_t r = (g * C + c) * N + j;
            const int s = __ldg(rowptr_t + r), e = __ldg(rowptr_t + r + 1);
            for (int k = s; k < e; ++k) {
                const float a = __ldg(val_t + k);
                const int64_t off = (g * N + __ldg(col_t + k)) * F + f0;   // row i of the forward
                const Vec<VEC> m = ldv<VEC>(chmax + c * ch_stride + off);
                const Vec<VEC> n = ldv<VEC>(nmax + c * ch_stride + off);
                const Vec<VEC> d = ldv<VEC>(dy + off);
#pragma unroll
                for (int t = 0; t < VEC; ++t)
                    if (__fmul_rn(a, xv.v[t]) == m.v[t]) acc.v[t] += a * (d.v[t] / n.v[t]);
            }
        }
        stv<VEC>(dx + gj * F + f0, acc);
    }
}

// rows of segment m: start[m] .. start[m] + size[m] - 1 (sizes are exclusive-scanned on the fly by the host mirror)
__global__ void segment_sum_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ start,
                                       const int64_t* __restrict__ size, int64_t n_seg, int feat, float* __restrict__ out) {
    pdl_prologue();
    const int64_t total = n_seg * feat;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = idx / feat;
        const int f = static_cast<int>(idx - m * feat);
        const float* p = x + start[m] * feat + f;
        float s = 0.0f;
        for (int64_t i = 0; i < size[m]; ++i) s += p[i * feat];   // index order, like the reference's reduce_sum per slice
        out[idx] = s;
    }
}
__global__ void segment_sum_bwd_kernel(const float* __restrict__ dout, const int64_t* __restrict__ start,
                                       const int64_t* __restrict__ size, int64_t n_seg, int feat, float* __restrict__ dx) {
    pdl_prologue();
    // one warp per segment row block: thread per (segment, feature) walks the segment's rows
    const int64_t total = n_seg * feat;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t m = idx / feat;
        const int f = static_cast<int>(idx - m * feat);
        const float g = dout[idx];
        float* p = dx + start[m] * feat + f;
        for (int64_t i = 0; i < size[m]; ++i) p[i * feat] = g;
    }
}

// ---- GraphBatchNormalization ----
// pass 1 (mode 1 only): per-block partial sums of x and x^2 over the enabled rows, fixed-order second stage
constexpr int kBnThreads = 256;
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, const int32_t* __restrict__ enabled,
                                                              int64_t n_graphs, int N, int F, int graphs_per_block,
                                                              double* __restrict__ partial /* [blocks][2][F] */) {
    pdl_prologue();
    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * graphs_per_block;
    const int64_t g1 = min(n_graphs, g0 + graphs_per_block);
    for (int f = threadIdx.x; f < F; f += kBnThreads) {
        double s = 0.0, q = 0.0;
        for (int64_t g = g0; g < g1; ++g) {
            const int n = enabled ? min(max(enabled[g], 0), N) : N;
            const float* p = x + g * N * F + f;
            for (int i = 0; i < n; ++i) {
                const double v = p[static_cast<int64_t>(i) * F];
                s += v;
                q += v * v;
            }
        }
        partial[(static_cast<int64_t>(blockIdx.x) * 2) * F + f] = s;
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * F + f] = q;
    }
}
__global__ void bn_finish_kernel(const double* __restrict__ partial, int blocks, int F, const int32_t* __restrict__ enabled,
                                 int64_t n_graphs, int N, float* __restrict__ mean, float* __restrict__ var) {
    pdl_prologue();
    __shared__ double cnt_s;
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int64_t g = 0; g < n_graphs; ++g) c += enabled ? min(max(enabled[g], 0), N) : N;
        cnt_s = c;
    }
    __syncthreads();
    const double cnt = cnt_s > 0.0 ? cnt_s : 1.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int b = 0; b < blocks; ++b) {
            s += partial[(static_cast<int64_t>(b) * 2) * F + f];
            q += partial[(static_cast<int64_t>(b) * 2 + 1) * F + f];
        }
        const double m = s / cnt;
        mean[f] = static_cast<float>(m);
        var[f] = static_cast<float>(fmax(q / cnt - m * m, 0.0));   // biased variance (tf.nn.moments)
    }
}
__global__ void bn_apply_kernel(const float* __restrict__ x, const int32_t* __restrict__ enabled, int64_t n_graphs, int N, int F,
                                const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, float* __restrict__ y) {
    pdl_prologue();
    const int64_t total = n_graphs * N * F;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(idx % F);
        const int64_t gi = idx / F;
        const int64_t g = gi / N;
        const int i = static_cast<int>(gi - g * N);
        const int n = enabled ? enabled[g] : N;
        float out = 0.0f;
        if (i < n) {
            const float inv = rsqrtf(var[f] + eps);
            const float sc = gamma ? gamma[f] * inv : inv;
            out = (x[idx] - mean[f]) * sc + (beta ? beta[f] : 0.0f);
        }
        y[idx] = out;
    }
}

// backward: per-block partial sums of dy and dy * xhat over the enabled rows (xhat = (x - mean) / sqrt(var + eps))
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                   const int32_t* __restrict__ enabled, int64_t n_graphs, int N, int F,
                                                                   int graphs_per_block, const float* __restrict__ mean,
                                                                   const float* __restrict__ var, float eps, double* __restrict__ partial) {
    pdl_prologue();
    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * graphs_per_block;
    const int64_t g1 = min(n_graphs, g0 + graphs_per_block);
    for (int f = threadIdx.x; f < F; f += kBnThreads) {
        const float m = mean[f], inv = rsqrtf(var[f] + eps);
        double s = 0.0, q = 0.0;
        for (int64_t g = g0; g < g1; ++g) {
            const int n = enabled ? min(max(enabled[g], 0), N) : N;
            const int64_t o = g * N * F + f;
            for (int i = 0; i < n; ++i) {
                const double d = dy[o + static_cast<int64_t>(i) * F];
                s += d;
                q += d * static_cast<double>((x[o + static_cast<int64_t>(i) * F] - m) * inv);
            }
        }
        partial[(static_cast<int64_t>(blockIdx.x) * 2) * F + f] = s;
        partial[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * F + f] = q;
    }
}
// dbeta = sum dy, dgamma = sum dy * xhat; sums[0..F) / sums[F..2F) keep them (as floats) for the apply pass
__global__ void bn_bwd_finish_kernel(const double* __restrict__ partial, int blocks, int F, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ sums) {
    pdl_prologue();
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int b = 0; b < blocks; ++b) {
            s += partial[(static_cast<int64_t>(b) * 2) * F + f];
            q += partial[(static_cast<int64_t>(b) * 2 + 1) * F + f];
        }
        if (dbeta) dbeta[f] = static_cast<float>(s);
        if (dgamma) dgamma[f] = static_cast<float>(q);
        sums[f] = static_cast<float>(s);
        sums[F + f] = static_cast<float>(q);
    }
}
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const int32_t* __restrict__ enabled,
                                    int64_t n_graphs, int N, int F, const float* __restrict__ mean, const float* __restrict__ var,
                                    const float* __restrict__ gamma, float eps, int mode, const float* __restrict__ sums,
                                    float* __restrict__ dx) {
    pdl_prologue();
    __shared__ float inv_cnt_s;
    if (threadIdx.x == 0) {
        double c = 0.0;
        if (mode == 1)
            for (int64_t g = 0; g < n_graphs; ++g) c += enabled ? min(max(enabled[g], 0), N) : N;
        inv_cnt_s = c > 0.0 ? static_cast<float>(1.0 / c) : 0.0f;
    }
    __syncthreads();
    const float inv_cnt = inv_cnt_s;
    const int64_t total = n_graphs * N * F;
    for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(idx % F);
        const int64_t gi = idx / F;
        const int64_t g = gi / N;
        const int i = static_cast<int>(gi - g * N);
        const int n = enabled ? enabled[g] : N;
        float out = 0.0f;
        if (i < n) {
            const float inv = rsqrtf(var[f] + eps);
            const float sc = gamma ? gamma[f] * inv : inv;
            if (mode == 0) {
                out = dy[idx] * sc;
            } else {   // batch statistics: the mean and the variance depend on x
                const float xhat = (x[idx] - mean[f]) * inv;
                out = sc * (dy[idx] - inv_cnt * (sums[f] + xhat * sums[F + f]));
            }
        }
        dx[idx] = out;
    }
}

inline unsigned ew_blocks(int64_t n) {
    const int64_t b = ceil_div<int64_t>(n, 256);
    return static_cast<unsigned>(b < 1 ? 1 : (b > kNumSMs * 16 ? kNumSMs * 16 : b));
}

}  // namespace
}  // namespace kgcn

using namespace kgcn;

extern "C" size_t kgcn_maxpool_workspace_bytes(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t feat) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || feat <= 0) return 0;
    return 2u * static_cast<size_t>(channels) * n_graphs * n_nodes * feat * sizeof(float);
}

extern "C" int kgcn_maxpool_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                    int32_t channels, int32_t n_nodes, const float* x, int32_t feat, float* y,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    KGCN_REQUIRE(rowptr && col && val && x && y, KGCN_ERR_NULL, "maxpool_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_nodes > 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "maxpool_fwd: bad shape");
    if (n_graphs == 0) return KGCN_OK;
    float* chmax = nullptr;
    float* nmax = nullptr;
    if (workspace != nullptr) {   // training: keep the per-channel maxima and tie counts for the backward
        KGCN_REQUIRE(workspace_bytes >= kgcn_maxpool_workspace_bytes(n_graphs, channels, n_nodes, feat), KGCN_ERR_WORKSPACE,
                     "maxpool_fwd: workspace too small");
        chmax = static_cast<float*>(workspace);
        nmax = chmax + static_cast<size_t>(channels) * n_graphs * n_nodes * feat;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool v4 = feat % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(workspace);
    if (v4)
        launch_pdl(maxpool_fwd_kernel<4>, ew_blocks(n_graphs * n_nodes * (feat / 4)), 256, 0, st, rowptr, col, val, n_graphs,
                   channels, n_nodes, feat, x, y, chmax, nmax);
    else
        launch_pdl(maxpool_fwd_kernel<1>, ew_blocks(n_graphs * n_nodes * feat), 256, 0, st, rowptr, col, val, n_graphs, channels,
                   n_nodes, feat, x, y, chmax, nmax);
    KGCN_LAUNCH_OK("maxpool_fwd_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_maxpool_bwd_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                    int32_t channels, int32_t n_nodes, const float* x, int32_t feat, const float* dy,
                                    const void* workspace, size_t workspace_bytes, float* dx, void* stream) {
    KGCN_REQUIRE(rowptr_t && col_t && val_t && x && dy && dx && workspace, KGCN_ERR_NULL, "maxpool_bwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_nodes > 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "maxpool_bwd: bad shape");
    if (n_graphs == 0) return KGCN_OK;
    KGCN_REQUIRE(workspace_bytes >= kgcn_maxpool_workspace_bytes(n_graphs, channels, n_nodes, feat), KGCN_ERR_WORKSPACE,
                 "maxpool_bwd: workspace too small");
    const float* chmax = static_cast<const float*>(workspace);
    const float* nmax = chmax + static_cast<size_t>(channels) * n_graphs * n_nodes * feat;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool v4 = feat % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(workspace);
    if (v4)
        launch_pdl(maxpool_bwd_kernel<4>, ew_blocks(n_graphs * n_nodes * (feat / 4)), 256, 0, st, rowptr_t, col_t, val_t, n_graphs,
                   channels, n_nodes, feat, x, dy, chmax, nmax, dx);
    else
        launch_pdl(maxpool_bwd_kernel<1>, ew_blocks(n_graphs * n_nodes * feat), 256, 0, st, rowptr_t, col_t, val_t, n_graphs,
                   channels, n_nodes, feat, x, dy, chmax, nmax, dx);
    KGCN_LAUNCH_OK("maxpool_bwd_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_segment_sum_fwd_f32(const float* x, const int64_t* start, const int64_t* size, int64_t n_segments,
                                        int32_t feat, float* out, void* stream) {
    KGCN_REQUIRE(x && start && size && out, KGCN_ERR_NULL, "segment_sum_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_segments >= 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "segment_sum_fwd: bad shape");
    if (n_segments == 0) return KGCN_OK;
    launch_pdl(segment_sum_fwd_kernel, ew_blocks(n_segments * feat), 256, 0, static_cast<cudaStream_t>(stream), x, start, size,
               n_segments, feat, out);
    KGCN_LAUNCH_OK("segment_sum_fwd_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_segment_sum_bwd_f32(const float* dout, const int64_t* start, const int64_t* size, int64_t n_segments,
                                        int32_t feat, float* dx, void* stream) {
    KGCN_REQUIRE(dout && start && size && dx, KGCN_ERR_NULL, "segment_sum_bwd: NULL pointer argument");
    KGCN_REQUIRE(n_segments >= 0 && feat > 0, KGCN_ERR_BAD_SHAPE, "segment_sum_bwd: bad shape");
    if (n_segments == 0) return KGCN_OK;
    launch_pdl(segment_sum_bwd_kernel, ew_blocks(n_segments * feat), 256, 0, static_cast<cudaStream_t>(stream), dout, start, size,
               n_segments, feat, dx);
    KGCN_LAUNCH_OK("segment_sum_bwd_kernel");
    return KGCN_OK;
}

extern "C" size_t kgcn_graph_bn_workspace_bytes(int64_t n_graphs, int32_t feat) {
    if (n_graphs <= 0 || feat <= 0) return 0;
    const int64_t blocks = std::min<int64_t>(n_graphs, 4 * kNumSMs);
    return static_cast<size_t>(blocks) * 2 * feat * sizeof(double);
}

extern "C" int kgcn_graph_bn_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat,
                                     const int32_t* enabled_node_nums, const float* gamma, const float* beta, float* mean,
                                     float* var, float eps, int32_t mode, float* y, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    KGCN_REQUIRE(x && y && mean && var, KGCN_ERR_NULL, "graph_bn_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && feat > 0 && (mode == 0 || mode == 1), KGCN_ERR_BAD_SHAPE, "graph_bn_fwd: bad shape / mode");
    if (n_graphs == 0) return KGCN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 1) {   // batch statistics over the enabled rows -> mean / var (outputs)
        KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= kgcn_graph_bn_workspace_bytes(n_graphs, feat), KGCN_ERR_WORKSPACE,
                     "graph_bn_fwd: workspace too small");
        const int blocks = static_cast<int>(std::min<int64_t>(n_graphs, 4 * kNumSMs));
        const int gpb = static_cast<int>(ceil_div<int64_t>(n_graphs, blocks));
        const int used = static_cast<int>(ceil_div<int64_t>(n_graphs, gpb));
        launch_pdl(bn_stats_kernel, used, kBnThreads, 0, st, x, enabled_node_nums, n_graphs, n_nodes, feat, gpb,
                   static_cast<double*>(workspace));
        KGCN_LAUNCH_OK("bn_stats_kernel");
        launch_pdl(bn_finish_kernel, 1, 256, 0, st, static_cast<const double*>(workspace), used, feat, enabled_node_nums, n_graphs,
                   n_nodes, mean, var);
        KGCN_LAUNCH_OK("bn_finish_kernel");
    }
    launch_pdl(bn_apply_kernel, ew_blocks(n_graphs * n_nodes * feat), 256, 0, st, x, enabled_node_nums, n_graphs, n_nodes, feat,
               static_cast<const float*>(mean), static_cast<const float*>(var), gamma, beta, eps, y);
    KGCN_LAUNCH_OK("bn_apply_kernel");
    return KGCN_OK;
}

extern "C" int kgcn_graph_bn_bwd_f32(const float* x, const float* dy, int64_t n_graphs, int32_t n_nodes, int32_t feat,
                                     const int32_t* enabled_node_nums, const float* gamma, const float* mean, const float* var,
                                     float eps, int32_t mode, float* dx, float* dgamma, float* dbeta, void* workspace,
                                     size_t workspace_bytes, void* stream) {
    KGCN_REQUIRE(x && dy && mean && var, KGCN_ERR_NULL, "graph_bn_bwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && n_nodes > 0 && feat > 0 && (mode == 0 || mode == 1), KGCN_ERR_BAD_SHAPE, "graph_bn_bwd: bad shape / mode");
    if (n_graphs == 0) return KGCN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t part_bytes = kgcn_graph_bn_workspace_bytes(n_graphs, feat);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= part_bytes + 2u * feat * sizeof(float), KGCN_ERR_WORKSPACE,
                 "graph_bn_bwd: workspace too small (need graph_bn_workspace_bytes + 2 * feat floats)");
    double* partial = static_cast<double*>(workspace);
    float* sums = reinterpret_cast<float*>(static_cast<char*>(workspace) + part_bytes);
    const int blocks = static_cast<int>(std::min<int64_t>(n_graphs, 4 * kNumSMs));
    const int gpb = static_cast<int>(ceil_div<int64_t>(n_graphs, blocks));
    const int used = static_cast<int>(ceil_div<int64_t>(n_graphs, gpb));
    launch_pdl(bn_bwd_reduce_kernel, used, kBnThreads, 0, st, x, dy, enabled_node_nums, n_graphs, n_nodes, feat, gpb, mean, var, eps,
               partial);
    KGCN_LAUNCH_OK("bn_bwd_reduce_kernel");
    launch_pdl(bn_bwd_finish_kernel, 1, 256, 0, st, static_cast<const double*>(partial), used, feat, dgamma, dbeta, sums);
    KGCN_LAUNCH_OK("bn_bwd_finish_kernel");
    if (dx != nullptr) {
        launch_pdl(bn_bwd_apply_kernel, ew_blocks(n_graphs * n_nodes * feat), 256, 0, st, x, dy, enabled_node_nums, n_graphs, n_nodes,
                   feat, mean, var, gamma, eps, mode, static_cast<const float*>(sums), dx);
        KGCN_LAUNCH_OK("bn_bwd_apply_kernel");
    }
    return KGCN_OK;
}
