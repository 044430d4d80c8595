// GraphConv forward / backward entry points (include/kgcn_b200.h).
//
// Two implementations sit behind kgcn_graphconv_fwd_f32:
//   * the fused single-kernel layer (graphconv_fused.cu) -- aggregate-first, tensor-core X.W,
//     used whenever the shape is eligible and KGCN_FLAG_REFERENCE_ORDER is not set;
//   * the decomposed path in this file, which keeps the reference's operation order
//     (kgcn/layers.py:112-113: fw = x.W_c + b_c, then A.fw, then the channel sum of :115) with
//     exact-fp32 FFMA GEMMs.  It handles every shape (odd feature widths, B=1 / huge N) and is
//     the in-library cross-check for the fused kernel.
#include <algorithm>

#include "common.cuh"

namespace kgcn {

// implemented in graphconv_fused.cu
bool fused_fwd_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x,
                        const float* y, const int32_t* rowptr, const int32_t* col, const float* val);
int launch_graphconv_fused_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                               int channels, int n_nodes, const float* x, int f_in, const float* w, const float* bias,
                               int f_out, int act, float* y, cudaStream_t st);

// implemented in graphconv_fused_v4.cu (warp-specialised pipeline, Z in tensor memory)
bool fused_v4_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x, const float* y,
                       const int32_t* rowptr, const int32_t* col, const float* val, const float* w, const float* bias);
int launch_graphconv_fused_v4(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs, int channels,
                              int n_nodes, const float* x, int f_in, const float* w, const float* bias, int f_out, int act,
                              float* y, cudaStream_t st, bool w_transposed = false, const float* mul_src = nullptr,
                              int mul_act = KGCN_ACT_NONE, int f_out_valid = 0);

// implemented in graphconv_fused_dw.cu (weight / bias gradient of a layer, all channels, one launch + the partial reduce)
bool fused_dw_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x, const float* du,
                       const int32_t* rowptr, const int32_t* col, const float* val);
size_t fused_dw_partial_bytes(int f_in, int n_total);
int launch_graphconv_fused_dw(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs, int channels,
                              int n_nodes, const float* x, int f_in, const float* du, int f_out, float* dw, float* dbias,
                              void* workspace, size_t workspace_bytes, cudaStream_t st);
int fused_dw_splits(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out);
int launch_graphconv_fused_dw_partial(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                      int channels, int n_nodes, const float* x, int f_in, const float* du, int f_out,
                                      float* partial, size_t partial_bytes, int* splits_out, cudaStream_t st);
bool fused_v4_plannable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out);

bool fused_bwd_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int act, bool dy_bcast,
                        const float* x, const float* w, const float* y, const float* dy, const float* dx,
                        const int32_t* rowptr, const int32_t* col, const float* val);
size_t fused_bwd_partial_bytes(int f_in, int f_out);
int launch_graphconv_fused_bwd(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                               int n_nodes, const float* x, int f_in, const float* w, int f_out, int act, const float* y,
                               const float* dy, bool dy_bcast, float* dx, float* dw, float* dbias, void* workspace,
                               size_t workspace_bytes, cudaStream_t st);

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Wide layers (F = 128) have no single-CTA plan in the v4 kernel ([W ; bias] hi / lo does not fit shared memory next to the
// stages): they run on the transposed-product kernel with the weights in tensor memory.
inline bool prefer_v5(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    return !fused_v4_chainable(n_graphs, channels, n_nodes, f_in, f_out) && fused_v5_plannable(n_graphs, channels, n_nodes, f_in, f_out);
}
inline bool ptrs_aligned16(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f, const void* g) {
    return aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d) && aligned16(e) && aligned16(f) && aligned16(g);
}
}  // namespace

}  // namespace kgcn

using namespace kgcn;

extern "C" size_t kgcn_graphconv_workspace_bytes(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t f_in,
                                                 int32_t f_out) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || f_in <= 0 || f_out <= 0) return 0;
    const size_t act_elems = static_cast<size_t>(n_graphs) * n_nodes * f_out;
    // forward: H[C][B*N][f_out]; backward: du[B*N][f_out] + G[C][B*N][f_out] + split-K partials
    return align_up((1 + static_cast<size_t>(channels)) * act_elems * sizeof(float), 256) +
           align_up(std::max({reduce_gemm_workspace_bytes(n_graphs * n_nodes, f_in, f_out), fused_bwd_partial_bytes(f_in, f_out),
                              fused_dw_partial_bytes(f_in, channels * f_out)}), 256);
}

extern "C" int kgcn_graphconv_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                      int32_t channels, int32_t n_nodes, const float* x, int32_t f_in, const float* w,
                                      const float* bias, int32_t f_out, int32_t act, float* y, int32_t flags,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    KGCN_REQUIRE(rowptr && col && val && x && w && y, KGCN_ERR_NULL, "graphconv_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_nodes > 0 && f_in > 0 && f_out > 0, KGCN_ERR_BAD_SHAPE,
                 "graphconv_fwd: bad shape n_graphs=%lld channels=%d n_nodes=%d f_in=%d f_out=%d", (long long)n_graphs,
                 channels, n_nodes, f_in, f_out);
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "graphconv_fwd: unknown act %d", act);
    if (n_graphs == 0) return KGCN_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    if (!(flags & KGCN_FLAG_REFERENCE_ORDER) && prefer_v5(n_graphs, channels, n_nodes, f_in, f_out) &&
        ptrs_aligned16(x, y, rowptr, col, val, w, bias)) {
        const V4ChainJob job{rowptr, col, val, x, w, bias, y, f_in, f_out, act, 0, nullptr, KGCN_ACT_NONE, f_out};
        return launch_graphconv_fused_v5_chain(&job, 1, n_graphs, channels, n_nodes, st);
    }
    if (!(flags & KGCN_FLAG_REFERENCE_ORDER) && fused_v4_eligible(n_graphs, channels, n_nodes, f_in, f_out, x, y, rowptr, col, val, w, bias))
        return launch_graphconv_fused_v4(rowptr, col, val, n_graphs, channels, n_nodes, x, f_in, w, bias, f_out, act, y, st);
    if (!(flags & KGCN_FLAG_REFERENCE_ORDER) && fused_fwd_eligible(n_graphs, channels, n_nodes, f_in, f_out, x, y, rowptr, col, val))
        return launch_graphconv_fused_fwd(rowptr, col, val, n_graphs, channels, n_nodes, x, f_in, w, bias, f_out, act,
                                          y, st);

    const int64_t rows = n_graphs * n_nodes;
    const size_t need = static_cast<size_t>(channels) * rows * f_out * sizeof(float);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= need, KGCN_ERR_WORKSPACE,
                 "graphconv_fwd: workspace %zu < %zu bytes", workspace_bytes, need);
    float* h = static_cast<float*>(workspace);  // [C][B*N][f_out]
    for (int c = 0; c < channels; ++c) {
        GemmEpilogue ep;
        ep.bias = bias ? bias + static_cast<size_t>(c) * f_out : nullptr;
        int rc = launch_sgemm(false, false, rows, f_out, f_in, x, f_in, w + static_cast<size_t>(c) * f_in * f_out,
                              f_out, h + static_cast<size_t>(c) * rows * f_out, f_out, ep, st);
        if (rc) return rc;
    }
    return launch_bspmm(rowptr, col, val, n_graphs, channels, n_nodes, n_nodes, f_out, h,
                        static_cast<int64_t>(n_nodes) * f_out, rows * f_out, y, static_cast<int64_t>(n_nodes) * f_out,
                        0, nullptr, act, st);
}

extern "C" int kgcn_graphconv_bwd_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t,
                                      int64_t n_graphs, int32_t channels, int32_t n_nodes, const float* x,
                                      int32_t f_in, const float* w, int32_t f_out, int32_t act, const float* y,
                                      const float* dy, float* dx, float* dw, float* dbias, int32_t flags,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    const bool dy_bcast = (flags & KGCN_FLAG_DY_BROADCAST) != 0;
    KGCN_REQUIRE(rowptr_t && col_t && val_t && x && w && dy && dw, KGCN_ERR_NULL, "graphconv_bwd: NULL pointer argument");
    KGCN_REQUIRE(act == KGCN_ACT_NONE || y != nullptr, KGCN_ERR_NULL, "graphconv_bwd: y required when act != none");
    KGCN_REQUIRE(n_graphs >= 0 && channels > 0 && n_nodes > 0 && f_in > 0 && f_out > 0, KGCN_ERR_BAD_SHAPE,
                 "graphconv_bwd: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_graphs == 0) {
        KGCN_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * channels * f_in * f_out, st));
        if (dbias) KGCN_CUDA_OK(cudaMemsetAsync(dbias, 0, sizeof(float) * channels * f_out, st));
        return KGCN_OK;
    }
    const int64_t rows = n_graphs * n_nodes;
    const size_t act_elems = static_cast<size_t>(rows) * f_out;
    KGCN_REQUIRE(workspace != nullptr &&
                     workspace_bytes >= kgcn_graphconv_workspace_bytes(n_graphs, channels, n_nodes, f_in, f_out),
                 KGCN_ERR_WORKSPACE, "graphconv_bwd: workspace too small");
    float* du = static_cast<float*>(workspace);                     // [B*N][f_out]
    float* gbuf = du + act_elems;                                   // [C][B*N][f_out]
    const size_t used = align_up((1 + static_cast<size_t>(channels)) * act_elems * sizeof(float), 256);
    void* ws2 = static_cast<char*>(workspace) + used;
    const size_t ws2_bytes = workspace_bytes - used;

    // Two streaming launches with concurrent roles (feature widths that are multiples of 32, any channel count):
    //   dx = sum_c (A_c^T.dU).W_c^T  = the fused layer kernel on (A^T, dU, W^T);   dW_c, dbias_c = graphconv_fused_dw.cu
    if (!(flags & KGCN_FLAG_REFERENCE_ORDER) && ws2_bytes >= fused_dw_partial_bytes(f_in, channels * f_out)) {
        const bool need_du = act != KGCN_ACT_NONE || dy_bcast;
        const float* du_in = need_du ? du : dy;
        if (fused_dw_eligible(n_graphs, channels, n_nodes, f_in, f_out, x, du_in, rowptr_t, col_t, val_t)) {
            if (need_du) {
                int rc = launch_act_grad(act != KGCN_ACT_NONE ? y : nullptr, dy, du, static_cast<int64_t>(act_elems), f_out, act,
                                         nullptr, n_nodes, dy_bcast, st);
                if (rc) return rc;
            }
            if (dx != nullptr && prefer_v5(n_graphs, channels, n_nodes, f_out, f_in) && ptrs_aligned16(du_in, dx, rowptr_t, col_t, val_t, w, nullptr)) {
                const V4ChainJob job{rowptr_t, col_t, val_t, du_in, w, nullptr, dx, f_out, f_in, KGCN_ACT_NONE, 1, nullptr, KGCN_ACT_NONE, f_in};
                int rc = launch_graphconv_fused_v5_chain(&job, 1, n_graphs, channels, n_nodes, st);
                if (rc) return rc;
            } else if (dx != nullptr && fused_v4_eligible(n_graphs, channels, n_nodes, f_out, f_in, du_in, dx, rowptr_t, col_t, val_t, w, nullptr)) {
                int rc = launch_graphconv_fused_v4(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, du_in, f_out, w, nullptr, f_in,
                                                   KGCN_ACT_NONE, dx, st, /*w_transposed=*/true);
                if (rc) return rc;
            } else if (dx != nullptr) {   // K = C * f_out too wide for tensor memory: G through HBM, exact-fp32 GEMMs
                int rc = launch_bspmm(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, n_nodes, f_out, du_in,
                                      static_cast<int64_t>(n_nodes) * f_out, 0, gbuf, static_cast<int64_t>(n_nodes) * f_out,
                                      static_cast<int64_t>(act_elems), nullptr, KGCN_ACT_NONE, st);
                if (rc) return rc;
                for (int c = 0; c < channels; ++c) {
                    GemmEpilogue ep;
                    ep.accumulate = c > 0;
                    rc = launch_sgemm(false, true, rows, f_in, f_out, gbuf + static_cast<size_t>(c) * act_elems, f_out,
                                      w + static_cast<size_t>(c) * f_in * f_out, f_out, dx, f_in, ep, st);
                    if (rc) return rc;
                }
            }
            return launch_graphconv_fused_dw(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, x, f_in, du_in, f_out, dw, dbias,
                                             ws2, ws2_bytes, st);
        }
    }

    // fused single-kernel backward (graphconv_fused_bwd.cu) when the shape is eligible
    if (!(flags & KGCN_FLAG_REFERENCE_ORDER) && ws2_bytes >= fused_bwd_partial_bytes(f_in, f_out) &&
        fused_bwd_eligible(n_graphs, channels, n_nodes, f_in, f_out, act, dy_bcast, x, w, y, dy, dx, rowptr_t, col_t, val_t))
        return launch_graphconv_fused_bwd(rowptr_t, col_t, val_t, n_graphs, n_nodes, x, f_in, w, f_out, act, y, dy, dy_bcast,
                                          dx, dw, dbias, ws2, ws2_bytes, st);

    const float* du_ptr = dy;
    if (act != KGCN_ACT_NONE || dy_bcast) {
        int rc = launch_act_grad(act != KGCN_ACT_NONE ? y : nullptr, dy, du, static_cast<int64_t>(act_elems), f_out, act,
                                 nullptr, n_nodes, dy_bcast, st);
        if (rc) return rc;
        du_ptr = du;
    }
    // G[c] = A_c^T . dU   (bspmm_call.py:44), channel-major so each G[c] is a contiguous [B*N, f_out]
    int rc = launch_bspmm(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, n_nodes, f_out, du_ptr,
                          static_cast<int64_t>(n_nodes) * f_out, 0, gbuf, static_cast<int64_t>(n_nodes) * f_out,
                          static_cast<int64_t>(act_elems), nullptr, KGCN_ACT_NONE, st);
    if (rc) return rc;
    for (int c = 0; c < channels; ++c) {
        const float* g_c = gbuf + static_cast<size_t>(c) * act_elems;
        rc = launch_reduce_gemm_tn(rows, f_in, f_out, x, f_in, g_c, f_out, dw + static_cast<size_t>(c) * f_in * f_out,
                                   dbias ? dbias + static_cast<size_t>(c) * f_out : nullptr, ws2, ws2_bytes, st);
        if (rc) return rc;
        if (dx != nullptr) {
            GemmEpilogue ep;
            ep.accumulate = c > 0;
            rc = launch_sgemm(false, true, rows, f_in, f_out, g_c, f_out, w + static_cast<size_t>(c) * f_in * f_out,
                              f_out, dx, f_in, ep, st);
            if (rc) return rc;
        }
    }
    return KGCN_OK;
}

// ---- step-loop form of the backward: dU in, dU of the layer below out, weight-gradient partials left for the fused tail ----
extern "C" int32_t kgcn_graphconv_bwd_splits(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t f_in, int32_t f_out,
                                             int32_t need_dx) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || f_in <= 0 || f_out <= 0) return 0;
    if (need_dx && !fused_v4_plannable(n_graphs, channels, n_nodes, f_out, f_in) && !fused_v5_plannable(n_graphs, channels, n_nodes, f_out, f_in)) return 0;
    return fused_dw_splits(n_graphs, channels, n_nodes, f_in, f_out);
}

extern "C" int kgcn_graphconv_bwd_partial_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                              int32_t channels, int32_t n_nodes, const float* x, int32_t f_in, const float* w,
                                              int32_t f_out, const float* du, float* dx, int32_t act_below, float* partial,
                                              size_t partial_bytes, void* stream) {
    KGCN_REQUIRE(rowptr_t && col_t && val_t && x && w && du && partial, KGCN_ERR_NULL, "graphconv_bwd_partial: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && f_in > 0 && f_out > 0, KGCN_ERR_BAD_SHAPE, "graphconv_bwd_partial: bad shape");
    KGCN_REQUIRE(act_below >= KGCN_ACT_NONE && act_below <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "graphconv_bwd_partial: unknown act %d", act_below);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    KGCN_REQUIRE(fused_dw_eligible(n_graphs, channels, n_nodes, f_in, f_out, x, du, rowptr_t, col_t, val_t), KGCN_ERR_UNSUPPORTED,
                 "graphconv_bwd_partial: shape / alignment not supported by the fused weight-gradient kernel (see kgcn_graphconv_bwd_splits)");
    if (dx != nullptr && prefer_v5(n_graphs, channels, n_nodes, f_out, f_in) && ptrs_aligned16(du, dx, rowptr_t, col_t, val_t, w, x)) {
        const V4ChainJob job{rowptr_t, col_t, val_t, du, w, nullptr, dx, f_out, f_in, KGCN_ACT_NONE, 1,
                             act_below != KGCN_ACT_NONE ? x : nullptr, act_below, f_in};
        int rc = launch_graphconv_fused_v5_chain(&job, 1, n_graphs, channels, n_nodes, st);
        if (rc) return rc;
    } else if (dx != nullptr) {
        KGCN_REQUIRE(fused_v4_eligible(n_graphs, channels, n_nodes, f_out, f_in, du, dx, rowptr_t, col_t, val_t, w, nullptr),
                     KGCN_ERR_UNSUPPORTED, "graphconv_bwd_partial: shape / alignment not supported by the fused layer kernel");
        int rc = launch_graphconv_fused_v4(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, du, f_out, w, nullptr, f_in,
                                           KGCN_ACT_NONE, dx, st, /*w_transposed=*/true, x, act_below);
        if (rc) return rc;
    }
    return launch_graphconv_fused_dw_partial(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, x, f_in, du, f_out, partial,
                                             partial_bytes, nullptr, st);
}

// Forward of a layer whose output width is padded (feature dims rounded up to the next multiple of 32 so that the
// tcgen05 kernels take Tox21-style widths such as 75 -> 50): columns >= f_out_valid are written as exact zeros whatever
// the activation, so the zero rows / columns of the padded weights never receive a gradient.
extern "C" int kgcn_graphconv_fwd_padded_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                             int32_t channels, int32_t n_nodes, const float* x, int32_t f_in, const float* w,
                                             const float* bias, int32_t f_out, int32_t f_out_valid, int32_t act, float* y,
                                             void* stream) {
    KGCN_REQUIRE(rowptr && col && val && x && w && y, KGCN_ERR_NULL, "graphconv_fwd_padded: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && f_in > 0 && f_out > 0 && f_out_valid > 0 && f_out_valid <= f_out,
                 KGCN_ERR_BAD_SHAPE, "graphconv_fwd_padded: bad shape");
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "graphconv_fwd_padded: unknown act %d", act);
    KGCN_REQUIRE(fused_v4_eligible(n_graphs, channels, n_nodes, f_in, f_out, x, y, rowptr, col, val, w, bias), KGCN_ERR_UNSUPPORTED,
                 "graphconv_fwd_padded: shape / alignment not supported by the fused layer kernel");
    return launch_graphconv_fused_v4(rowptr, col, val, n_graphs, channels, n_nodes, x, f_in, w, bias, f_out, act, y,
                                     static_cast<cudaStream_t>(stream), false, nullptr, KGCN_ACT_NONE, f_out_valid);
}

extern "C" int32_t kgcn_graphconv_fwd_fused(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t f_in, int32_t f_out) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || f_in <= 0 || f_out <= 0) return 0;
    return (fused_v4_plannable(n_graphs, channels, n_nodes, f_in, f_out) || fused_v5_plannable(n_graphs, channels, n_nodes, f_in, f_out)) ? 1 : 0;
}

extern "C" int kgcn_reduce_partials_f32(const float* partial, int32_t splits, int32_t f_in, int32_t f_out, int32_t channels,
                                        float* dw, float* dbias, void* stream) {
    KGCN_REQUIRE(partial && dw, KGCN_ERR_NULL, "reduce_partials: NULL pointer argument");
    KGCN_REQUIRE(splits > 0 && f_in > 0 && f_out > 0 && channels > 0, KGCN_ERR_BAD_SHAPE, "reduce_partials: bad shape");
    return launch_splitk_reduce_ch(partial, splits, f_in, f_out, channels, dw, dbias, static_cast<cudaStream_t>(stream));
}

// ---- chained launches for the step loop: all forward layers / all dx layers of a network in ONE launch each ----
// forward jobs of one layer appended to jobs[k..]: one job, or several over channel groups (the later ones accumulate into y
// and the last activates).  Returns the new job count, -1 when the layer has no single-CTA plan or the table is full.
static int add_fwd_jobs(V4ChainJob* jobs, int k, int k_max, const int32_t* rowptr, const int32_t* col, const float* val, const float* x,
                        const float* w, const float* bias, float* y, int f_in, int f_out, int f_valid, int act, const V4Head* head,
                        int64_t n_graphs, int channels, int n_nodes) {
    const int cg = fused_v4_chain_group(n_graphs, channels, n_nodes, f_in, f_out, head ? head->n_labels : 0);
    if (cg == 0) return -1;
    for (int c0 = 0; c0 < channels; c0 += cg) {
        if (k >= k_max) return -1;
        const int cn = std::min(cg, channels - c0);
        const bool first = c0 == 0, last = c0 + cn == channels;
        V4ChainJob j{rowptr, col, val, x, w, bias, y, f_in, f_out, last ? act : KGCN_ACT_NONE, 0, nullptr, KGCN_ACT_NONE,
                     last ? f_valid : f_out, last ? head : nullptr};
        j.c_begin = c0;
        j.c_count = cn;
        j.acc_in = first ? 0 : 1;
        jobs[k++] = j;
    }
    return k;
}

// 1: every layer (possibly over channel groups) and every dx has a single-CTA plan in the v4 kernel, at most 6 forward jobs;
// 2: every layer and dx runs on the v5 kernel (wide layers); 0: neither
static int chain_kind(int64_t n_graphs, int channels, int n_nodes, int n_layers, const int32_t* dims) {
    bool v4 = true, v5 = true;
    int fwd_jobs = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int cg = fused_v4_chain_group(n_graphs, channels, n_nodes, dims[l], dims[l + 1], 0);
        v4 = v4 && cg > 0;
        if (cg > 0) fwd_jobs += (channels + cg - 1) / cg;
        v5 = v5 && fused_v5_plannable(n_graphs, channels, n_nodes, dims[l], dims[l + 1]);
        if (l > 0) {
            v4 = v4 && fused_v4_chainable(n_graphs, channels, n_nodes, dims[l + 1], dims[l]);
            v5 = v5 && fused_v5_plannable(n_graphs, channels, n_nodes, dims[l + 1], dims[l]);
        }
    }
    v4 = v4 && fwd_jobs <= 6;
    v5 = v5 && n_layers <= 4;
    return v4 ? 1 : (v5 ? 2 : 0);
}

extern "C" int32_t kgcn_graphconv_chain_supported(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers,
                                                  const int32_t* dims) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || n_layers < 1 || n_layers > 4 || dims == nullptr) return 0;
    return chain_kind(n_graphs, channels, n_nodes, n_layers, dims);
}

extern "C" int kgcn_graphconv_chain_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                            int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                            const int32_t* dims_valid, const float* x, const float* const* w,
                                            const float* const* bias, float* const* y, int32_t act, void* stream) {
    KGCN_REQUIRE(rowptr && col && val && dims && x && w && y, KGCN_ERR_NULL, "graphconv_chain_fwd: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && n_layers >= 1 && n_layers <= 4, KGCN_ERR_BAD_SHAPE,
                 "graphconv_chain_fwd: bad shape (1..4 layers)");
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "graphconv_chain_fwd: unknown act %d", act);
    V4ChainJob jobs[6];
    const float* in = x;
    const int kind = chain_kind(n_graphs, channels, n_nodes, n_layers, dims);
    KGCN_REQUIRE(kind != 0, KGCN_ERR_UNSUPPORTED, "graphconv_chain_fwd: network not supported (kgcn_graphconv_chain_supported)");
    int k = 0;
    for (int l = 0; l < n_layers; ++l) {
        KGCN_REQUIRE(w[l] && y[l], KGCN_ERR_NULL, "graphconv_chain_fwd: NULL weight / output of layer %d", l);
        KGCN_REQUIRE(aligned16(in) && aligned16(y[l]) && aligned16(w[l]) && (!bias || aligned16(bias[l])) && aligned16(rowptr) &&
                         aligned16(col) && aligned16(val), KGCN_ERR_MISALIGNED, "graphconv_chain_fwd: 16-byte alignment required");
        const int fv = dims_valid ? dims_valid[l + 1] : dims[l + 1];
        if (kind == 2) {
            jobs[k++] = V4ChainJob{rowptr, col, val, in, w[l], bias ? bias[l] : nullptr, y[l], dims[l], dims[l + 1], act, 0, nullptr,
                                   KGCN_ACT_NONE, fv};
        } else {
            k = add_fwd_jobs(jobs, k, 6, rowptr, col, val, in, w[l], bias ? bias[l] : nullptr, y[l], dims[l], dims[l + 1], fv, act, nullptr,
                             n_graphs, channels, n_nodes);
            KGCN_REQUIRE(k > 0, KGCN_ERR_UNSUPPORTED, "graphconv_chain_fwd: layer %d has no plan", l);
        }
        in = y[l];
    }
    if (kind == 2) return launch_graphconv_fused_v5_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream));
    return launch_graphconv_fused_v4_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream));
}

extern "C" int kgcn_graphconv_chain_dx_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                           int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                           const float* const* x, const float* const* w, float* const* du, int32_t act,
                                           void* stream) {
    KGCN_REQUIRE(rowptr_t && col_t && val_t && dims && x && w && du, KGCN_ERR_NULL, "graphconv_chain_dx: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && n_layers >= 2 && n_layers <= 5, KGCN_ERR_BAD_SHAPE,
                 "graphconv_chain_dx: bad shape (2..5 layers)");
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "graphconv_chain_dx: unknown act %d", act);
    // du[l] = dU of layer l ([B, N, dims[l + 1]]); du[n_layers - 1] is the input, du[n_layers - 2] .. du[0] are outputs:
    // du[l - 1] = (sum_c A_c^T . du[l] . W_l,c^T) (.) act'(x[l]),  x[l] = input of layer l = output of layer l - 1
    V4ChainJob jobs[4];
    int k = 0;
    for (int l = n_layers - 1; l >= 1; --l, ++k) {
        KGCN_REQUIRE(x[l] && w[l] && du[l] && du[l - 1], KGCN_ERR_NULL, "graphconv_chain_dx: NULL pointer at layer %d", l);
        KGCN_REQUIRE(aligned16(x[l]) && aligned16(w[l]) && aligned16(du[l]) && aligned16(du[l - 1]) && aligned16(rowptr_t) &&
                         aligned16(col_t) && aligned16(val_t), KGCN_ERR_MISALIGNED, "graphconv_chain_dx: 16-byte alignment required");
        jobs[k] = V4ChainJob{rowptr_t, col_t, val_t, du[l], w[l], nullptr, du[l - 1], dims[l + 1], dims[l], KGCN_ACT_NONE, 1,
                             x[l], act == KGCN_ACT_NONE ? KGCN_ACT_NONE : act, 0};
    }
    if (chain_kind(n_graphs, channels, n_nodes, n_layers, dims) == 2)
        return launch_graphconv_fused_v5_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream));
    return launch_graphconv_fused_v4_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream));
}

extern "C" int kgcn_graphconv_chain_dw_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                           int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                           const float* const* x, const float* const* du, float* const* partial,
                                           const size_t* partial_bytes, void* stream) {
    return kgcn_graphconv_chain_dw_g_f32(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, n_layers, dims, x, du, nullptr, partial,
                                         partial_bytes, stream);
}

extern "C" int kgcn_graphconv_chain_dw_g_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                             int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                             const float* const* x, const float* const* du, const float* const* g,
                                             float* const* partial, const size_t* partial_bytes, void* stream) {
    KGCN_REQUIRE(rowptr_t && col_t && val_t && dims && x && du && partial && partial_bytes, KGCN_ERR_NULL,
                 "graphconv_chain_dw: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && n_layers >= 1 && n_layers <= 8, KGCN_ERR_BAD_SHAPE,
                 "graphconv_chain_dw: bad shape (1..8 layers)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int l = 0;
    while (l < n_layers) {   // as many consecutive layers per launch as tensor memory holds accumulators for
        int f_out[4], n_try = std::min(4, n_layers - l);
        for (int k = 0; k < n_try; ++k) f_out[k] = dims[l + k + 1];
        const int n = std::max(1, fused_dw_jobs_per_launch(channels, f_out, n_try));
        DwJob jobs[4];
        for (int k = 0; k < n; ++k) {
            const int i = l + k;
            KGCN_REQUIRE(x[i] && du[i] && partial[i], KGCN_ERR_NULL, "graphconv_chain_dw: NULL pointer at layer %d", i);
            KGCN_REQUIRE(fused_dw_eligible(n_graphs, channels, n_nodes, dims[i], dims[i + 1], x[i], du[i], rowptr_t, col_t, val_t),
                         KGCN_ERR_UNSUPPORTED, "graphconv_chain_dw: layer %d (%d -> %d) is not supported by the fused weight-gradient kernel",
                         i, dims[i], dims[i + 1]);
            jobs[k] = DwJob{rowptr_t, col_t, val_t, x[i], du[i], partial[i], partial_bytes[i], dims[i], dims[i + 1], g ? g[i] : nullptr};
        }
        const int rc = launch_graphconv_fused_dw_jobs(jobs, n, n_graphs, channels, n_nodes, nullptr, st);
        if (rc) return rc;
        l += n;
    }
    return KGCN_OK;
}

extern "C" int32_t kgcn_gcn_step_chain_grid(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers,
                                            const int32_t* dims, int32_t n_labels) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || n_layers < 1 || n_layers > 4 || dims == nullptr) return 0;
    if (n_labels < 1 || n_labels > 4) return 0;
    const int kind = chain_kind(n_graphs, channels, n_nodes, n_layers, dims);
    if (kind == 2)   // wide layers (v5 kernel): forward jobs + head + dx jobs, at most 4 jobs
        return 2 * n_layers - 1 <= 4 ? fused_v5_head_grid(n_graphs, channels, n_nodes, dims[n_layers - 1], dims[n_layers], n_labels) : 0;
    if (kind != 1) return 0;
    int jobs = n_layers - 1;   // dx jobs
    for (int l = 0; l < n_layers; ++l) {
        const int cg = fused_v4_chain_group(n_graphs, channels, n_nodes, dims[l], dims[l + 1], l == n_layers - 1 ? n_labels : 0);
        if (cg == 0) return 0;
        jobs += (channels + cg - 1) / cg;
    }
    if (jobs > 6) return 0;
    return fused_v4_chain_grid(n_graphs, channels, n_nodes, dims[n_layers - 1], dims[n_layers]);
}

extern "C" int kgcn_gcn_step_chain_f32(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* rowptr_t,
                                       const int32_t* col_t, const float* val_t, int64_t n_graphs, int32_t channels,
                                       int32_t n_nodes, int32_t n_layers, const int32_t* dims, const int32_t* dims_valid,
                                       const float* x, const float* const* w, const float* const* bias, float* const* y,
                                       float* const* du, int32_t act, const float* head_w, const float* head_b,
                                       int32_t n_labels, const float* labels, const float* mask, float inv_batch,
                                       float* logits, float* prediction, float* gathered, float* head_partial,
                                       uint32_t flags, void* stream) {
    return kgcn_gcn_step_chain_g_f32(rowptr, col, val, rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, n_layers, dims, dims_valid, x, w,
                                     bias, y, du, nullptr, act, head_w, head_b, n_labels, labels, mask, inv_batch, logits, prediction,
                                     gathered, head_partial, flags, stream);
}

extern "C" int32_t kgcn_gcn_step_chain_g_supported(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers,
                                                   const int32_t* dims) {
    if (n_graphs <= 0 || channels <= 0 || n_nodes <= 0 || n_layers < 2 || n_layers > 4 || dims == nullptr) return 0;
    return chain_kind(n_graphs, channels, n_nodes, n_layers, dims) != 0 ? 1 : 0;   // the v4 chained kernel or the wide-layer (v5) one
}

extern "C" int kgcn_gcn_step_chain_g_f32(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* rowptr_t,
                                         const int32_t* col_t, const float* val_t, int64_t n_graphs, int32_t channels,
                                         int32_t n_nodes, int32_t n_layers, const int32_t* dims, const int32_t* dims_valid,
                                         const float* x, const float* const* w, const float* const* bias, float* const* y,
                                         float* const* du, float* const* g_save, int32_t act, const float* head_w,
                                         const float* head_b, int32_t n_labels, const float* labels, const float* mask,
                                         float inv_batch, float* logits, float* prediction, float* gathered, float* head_partial,
                                         uint32_t flags, void* stream) {
    KGCN_REQUIRE(rowptr && col && val && rowptr_t && col_t && val_t && dims && x && w && y && du && head_w && labels && head_partial,
                 KGCN_ERR_NULL, "gcn_step_chain: NULL pointer argument");
    KGCN_REQUIRE(g_save == nullptr || kgcn_gcn_step_chain_g_supported(n_graphs, channels, n_nodes, n_layers, dims), KGCN_ERR_UNSUPPORTED,
                 "gcn_step_chain: storing G is not supported for this network (kgcn_gcn_step_chain_g_supported)");
    KGCN_REQUIRE(n_graphs > 0 && channels > 0 && n_nodes > 0 && n_layers >= 1 && n_layers <= 4, KGCN_ERR_BAD_SHAPE,
                 "gcn_step_chain: bad shape (1..4 layers)");
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "gcn_step_chain: unknown act %d", act);
    const int L = n_layers;
    V4ChainJob jobs[6];
    V4Head head{n_labels, head_w, head_b, labels, mask, inv_batch, logits, prediction, gathered, head_partial};
    const float* in = x;
    int k = 0;
    const int kind = chain_kind(n_graphs, channels, n_nodes, n_layers, dims);
    KGCN_REQUIRE(kind != 0, KGCN_ERR_UNSUPPORTED, "gcn_step_chain: network not supported (kgcn_gcn_step_chain_grid)");
    if (kind == 2) {   // wide layers: the v5 kernel, one job per layer
        KGCN_REQUIRE(2 * L - 1 <= 4, KGCN_ERR_UNSUPPORTED, "gcn_step_chain: network not supported (kgcn_gcn_step_chain_grid)");
        for (int l = 0; l < L; ++l, ++k) {
            float* out = (l == L - 1) ? du[L - 1] : y[l];
            KGCN_REQUIRE(w[l] && out, KGCN_ERR_NULL, "gcn_step_chain: NULL weight / output of layer %d", l);
            KGCN_REQUIRE(aligned16(in) && aligned16(out) && aligned16(w[l]) && (!bias || aligned16(bias[l])), KGCN_ERR_MISALIGNED,
                         "gcn_step_chain: 16-byte alignment required");
            jobs[k] = V4ChainJob{rowptr, col, val, in, w[l], bias ? bias[l] : nullptr, out, dims[l], dims[l + 1], act, 0, nullptr, KGCN_ACT_NONE,
                                 dims_valid ? dims_valid[l + 1] : dims[l + 1], (l == L - 1) ? &head : nullptr};
            in = out;
        }
        for (int l = L - 1; l >= 1; --l, ++k) {
            KGCN_REQUIRE(du[l] && du[l - 1] && y[l - 1], KGCN_ERR_NULL, "gcn_step_chain: NULL pointer at dx of layer %d", l);
            jobs[k] = V4ChainJob{rowptr_t, col_t, val_t, du[l], w[l], nullptr, du[l - 1], dims[l + 1], dims[l], KGCN_ACT_NONE, 1,
                                 y[l - 1], act, 0, nullptr};
            if (g_save != nullptr) jobs[k].zsave = g_save[l];   // G_l = A^T . du[l]: the aggregate of this very job
        }
        return launch_graphconv_fused_v5_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream));
    }
    for (int l = 0; l < L; ++l) {
        float* out = (l == L - 1) ? du[L - 1] : y[l];
        KGCN_REQUIRE(w[l] && out, KGCN_ERR_NULL, "gcn_step_chain: NULL weight / output of layer %d", l);
        KGCN_REQUIRE(aligned16(in) && aligned16(out) && aligned16(w[l]) && (!bias || aligned16(bias[l])), KGCN_ERR_MISALIGNED,
                     "gcn_step_chain: 16-byte alignment required");
        k = add_fwd_jobs(jobs, k, 6, rowptr, col, val, in, w[l], bias ? bias[l] : nullptr, out, dims[l], dims[l + 1],
                         dims_valid ? dims_valid[l + 1] : dims[l + 1], act, (l == L - 1) ? &head : nullptr, n_graphs, channels, n_nodes);
        KGCN_REQUIRE(k > 0 && k + (L - 1 - l) <= 6, KGCN_ERR_UNSUPPORTED, "gcn_step_chain: network not supported (kgcn_gcn_step_chain_grid)");
        in = out;
    }
    for (int l = L - 1; l >= 1; --l, ++k) {   // du[l - 1] = (sum_c A_c^T . du[l] . W_l,c^T) (.) act'(y[l - 1])
        KGCN_REQUIRE(du[l] && du[l - 1] && y[l - 1], KGCN_ERR_NULL, "gcn_step_chain: NULL pointer at dx of layer %d", l);
        KGCN_REQUIRE(aligned16(du[l]) && aligned16(du[l - 1]), KGCN_ERR_MISALIGNED, "gcn_step_chain: 16-byte alignment required");
        jobs[k] = V4ChainJob{rowptr_t, col_t, val_t, du[l], w[l], nullptr, du[l - 1], dims[l + 1], dims[l], KGCN_ACT_NONE, 1,
                             y[l - 1], act, 0, nullptr};
        if (g_save != nullptr) jobs[k].zsave = g_save[l];   // G_l = A^T . du[l]: the aggregate of this very job
    }
    return launch_graphconv_fused_v4_chain(jobs, k, n_graphs, channels, n_nodes, static_cast<cudaStream_t>(stream),
                                           (flags & KGCN_FLAG_INPUTS_STABLE) != 0);
}
