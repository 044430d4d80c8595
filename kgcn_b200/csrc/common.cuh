// Shared helpers for libkgcn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/kgcn_b200.h"

namespace kgcn {

// thread-local last-error text (include/kgcn_b200.h: kgcn_last_error)
char* error_buffer();
int fail(int code, const char* fmt, ...);
void count_launch();  // process-wide counter of kernel launches issued by this library (kgcn_launch_count)

#define KGCN_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t err__ = (expr);                                                             \
        if (err__ != cudaSuccess)                                                               \
            return ::kgcn::fail(KGCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                                cudaGetErrorString(err__), __FILE__, __LINE__);                 \
    } while (0)

#define KGCN_LAUNCH_OK(what)                                                                    \
    do {                                                                                        \
        ::kgcn::count_launch();                                                                 \
        cudaError_t err__ = cudaGetLastError();                                                 \
        if (err__ != cudaSuccess)                                                               \
            return ::kgcn::fail(KGCN_ERR_CUDA, "launch of %s failed: %s", what,                 \
                                cudaGetErrorString(err__));                                     \
    } while (0)

#define KGCN_REQUIRE(cond, code, ...)                                                           \
    do {                                                                                        \
        if (!(cond)) return ::kgcn::fail(code, __VA_ARGS__);                                    \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Programmatic dependent launch: every kernel of this library is launched with the
// programmatic-stream-serialization attribute and starts with pdl_prologue(), so inside a stream /
// CUDA graph the next kernel's grid is already resident and past its launch latency when the
// previous one drains (the per-step work is ~15 launches of a few microseconds each).
// KGCN_PDL=0 in the environment disables the attribute (plain stream order).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
    // No explicit launch_dependents: the implicit trigger at block exit lets the next grid's CTAs move in
    // as this grid's tail drains.  (Triggering at kernel entry was measured slower: the waiting grid
    // then competes for SM resources with the running one.)
    pdl_wait();   // all global reads / writes of this kernel come after the producer grid has completed
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Activations.  Accurate expf/tanhf on purpose: parity tolerance is 1e-5 relative and these are
// never the bottleneck (the kernels are HBM-bound).
__device__ __forceinline__ float apply_act(float x, int act) {
    switch (act) {
        case KGCN_ACT_RELU: return fmaxf(x, 0.0f);
        case KGCN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        case KGCN_ACT_TANH: return tanhf(x);
        default: return x;
    }
}

// d act / d pre-activation, expressed through the activation OUTPUT y (what forward stored).
__device__ __forceinline__ float act_grad_from_output(float y, int act) {
    switch (act) {
        case KGCN_ACT_RELU: return y > 0.0f ? 1.0f : 0.0f;
        case KGCN_ACT_SIGMOID: return y * (1.0f - y);
        case KGCN_ACT_TANH: return 1.0f - y * y;
        default: return 1.0f;
    }
}

// ---- internal launchers shared between translation units (all enqueue on `st`) ----

// C[M,N] (ldc) = epilogue( op(A)[M,K] . op(B)[K,N] ), optional accumulate into C.
struct GemmEpilogue {
    const float* bias = nullptr;          // [N], added per column
    int act = KGCN_ACT_NONE;              // applied after bias
    const int32_t* enabled = nullptr;     // [M / n_nodes] rows >= enabled[g] -> 0
    int n_nodes = 0;
    bool accumulate = false;              // C += result (no bias/act allowed then)
};
int launch_sgemm(bool trans_a, bool trans_b, int64_t M, int N, int K, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float* C, int64_t ldc, const GemmEpilogue& ep, cudaStream_t st);

// out[Ka, N] = A[M, Ka]^T . B[M, N], reduced over the (long) M dimension with a deterministic
// two-pass split-K; also emits colsum_b[N] = column sums of B (may be null).
size_t reduce_gemm_workspace_bytes(int64_t M, int Ka, int N);
int launch_reduce_gemm_tn(int64_t M, int Ka, int N, const float* A, int64_t lda, const float* B, int64_t ldb,
                          float* out, float* colsum_b, void* workspace, size_t workspace_bytes, cudaStream_t st);

// out[M, N] (+ colsum_out[N], may be null) = sum over `splits` partial blocks of (M + 1) * N floats, fixed order.
int launch_splitk_reduce(const float* partial, int splits, int64_t M, int N, float* out, float* colsum_out,
                         cudaStream_t st);
// same for partial blocks of (M + 1) x (channels * n_per_ch): out is channel-major [channels][M][n_per_ch]
int launch_splitk_reduce_ch(const float* partial, int splits, int64_t M, int n_per_ch, int channels, float* out,
                            float* colsum_out, cudaStream_t st);

// du = dy * act'(y) (elementwise), optional row mask by enabled_node_nums.
int launch_act_grad(const float* y, const float* dy, float* du, int64_t n, int feat, int act,
                    const int32_t* enabled, int n_nodes, bool dy_bcast, cudaStream_t st);

// Readout head fused into the epilogue of a chain job (training step): GraphGather + Dense(n_labels) + softmax
// cross-entropy; the job then stores dU = d gathered (.) act'(H) instead of H.
struct V4Head {
    int n_labels;
    const float* w;        // [f_out][n_labels]
    const float* b;        // [n_labels] or NULL
    const float* labels;   // [B][n_labels]
    const float* mask;     // [B] or NULL
    float inv_batch;
    float* logits;         // [B][n_labels] or NULL
    float* prediction;     // [B][n_labels] or NULL
    float* gathered;       // [B][f_out] or NULL
    float* partial;        // [grid][f_out * n_labels + 8]
};
// One job of a chained fused-layer launch (graphconv_fused_v4.cu): y = epilogue((A . x) . W) on the job's own widths.
struct V4ChainJob {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* x;         // [B, N, f_in]
    const float* w;         // forward: [C][f_in][f_out]; w_transposed (backward dx): the layer's [C][f_out][f_in]
    const float* bias;      // [C][f_out] or NULL (ignored when w_transposed)
    float* y;               // [B, N, f_out]
    int f_in, f_out, act, w_transposed;
    const float* mul_src;   // y *= act'(mul_src) of activation mul_act
    int mul_act, f_out_valid;
    const V4Head* head = nullptr;
    int c_begin = 0, c_count = 0;   // channel group [c_begin, c_begin + c_count) of the CSR's channels (0 = all)
    int acc_in = 0;                 // add the existing y before the activation (later group of a layer split over channel groups)
    float* zsave = nullptr;         // optional [B, N, C * f_in]: the job also stores its aggregate Z = [A_0 . x | A_1 . x | ..] (for a dx job
                                    // that is G = A^T . dU of the layer, which the weight-gradient kernel then reads instead of gathering it again)
};
// channels per job a forward layer needs in a chain: `channels` (one job), fewer (K = C * f_in beyond tensor memory: channel
// groups, the later ones accumulate), 0 = no single-CTA plan
int fused_v4_chain_group(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int head_labels);
bool soft_jobs_enabled();   // graphconv_fused_dw.cu: KGCN_SOFT_JOBS (A-B knob, default on)
// graphconv_fused_v5.cu: the transposed-product kernel for wide layers (weights resident in tensor memory), same job struct
bool fused_v5_plannable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out);
int fused_v5_head_grid(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int n_labels);   // 0: no fused head for this shape
int launch_graphconv_fused_v5_chain(const V4ChainJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, cudaStream_t st);
bool fused_v4_head_chainable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int n_labels);
int fused_v4_chain_grid(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out);
bool fused_v4_chainable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out);
// early_inputs: job 0's x / CSR may be read before the previous kernel of the stream has completed (KGCN_FLAG_INPUTS_STABLE)
int launch_graphconv_fused_v4_chain(const V4ChainJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, cudaStream_t st,
                                    bool early_inputs = false);

// One job of a multi-layer weight-gradient launch (graphconv_fused_dw.cu)
struct DwJob {
    const int32_t* rowptr_t;
    const int32_t* col_t;
    const float* val_t;
    const float* x;        // [B, N, f_in]   input of the layer
    const float* du;       // [B, N, f_out]  dU of the layer
    float* partial;        // [splits][(f_in + 1)][channels * f_out]
    size_t partial_bytes;
    int f_in, f_out;
    const float* g = nullptr;   // optional [B, N, channels * f_out]: G = [A_0^T . dU | A_1^T . dU | ..] of the layer, already computed (by the dx
                                // job of the chained launch, V4ChainJob::zsave) -- the kernel then copies G rows instead of gathering them
};
int launch_graphconv_fused_dw_jobs(const DwJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, int* splits_out,
                                   cudaStream_t st);
int fused_dw_jobs_per_launch(int channels, const int* f_out, int n_jobs);

int launch_bspmm(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs, int channels,
                 int n_rows, int n_cols, int feat, const float* rhs, int64_t rs_g, int64_t rs_c, float* out,
                 int64_t os_g, int64_t os_c, const float* self_scale, int act, cudaStream_t st);

}  // namespace kgcn
