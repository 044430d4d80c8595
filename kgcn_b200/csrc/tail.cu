// The tail of a training step in ONE launch: fixed-order reduction of the per-CTA weight-gradient partials of every
// GraphConv layer -> (data parallel: one-shot all-reduce over NVLink peer memory) -> Adam.
//
//   reduce     graphconv_fused_dw_kernel leaves one partial block [(f_in + 1), C * f_out] per CTA; a lane owns 4 consecutive
//              elements of the flat parameter buffer (one 16-byte load per partial), the 32 warps of a block take splits
//              w, w + 32, .. and the 32 warp sums are added in warp order, so the result is deterministic.
//              Parameters outside every segment (readout head, GraphDense) already have their gradient in `grad`.
//   all-reduce low-latency PUSH over peer-mapped memory (cudaIpc, kgcn_p2p_*): every rank owns a mailbox
//              ll[2][W][n_pad] of 8-byte slots {fp32 value, step}.  A lane stores its 4 local sums, tagged with the step
//              number, straight into slot [step & 1][my rank] of every PEER's mailbox (posted NVLink writes, no fence, no
//              flag: an aligned 8-byte store is single-copy atomic, so value and tag arrive together -- NCCL's LL idea),
//              then spins on its OWN mailbox (local memory) until the W - 1 peers' slots carry this step's tag and adds
//              the values in rank order: every rank computes bit-identical sums in one NVLink one-way latency.
//              Double-buffered by step parity: a rank can reach step t + 2 only after it consumed every peer's step t + 1
//              slot, which that peer wrote after finishing its step t -- so nobody still reads what is overwritten.
//              The reference has no counterpart (single process); this replaces the NCCL all-reduce between two
//              graph replays of round 1 (kgcn/core.py:121-127 is the optimizer it feeds).
//   Adam       TensorFlow's formulation (kgcn_adam_f32), step counter on the device so the launch replays from a CUDA graph.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace kgcn {
namespace {

constexpr int kMaxSegments = 16;
constexpr int kMaxWorld = 8;
constexpr int kTailWarps = 32;   // (2 warps at <= 64 registers would fit beside a chained-layer CTA on every SM, so the next step
                                 // could start everywhere under this kernel: measured slower, 8.9 vs 6.1 us alone and +6 us per step)
constexpr int kTailThreads = kTailWarps * 32;
constexpr int kTailElems = 128;   // elements per block: 32 lanes x 4

struct TailSegment {
    long long kernel_off, bias_off;   // offsets into the flat buffers; kernel [C][rows][cols], bias [C][cols] (-1: none)
    const float* partial;             // [splits][(rows + 1)][C * cols]
    int splits, rows, cols, channels;
    long long stride;                 // floats between splits
};

struct TailParams {
    float* param;
    float* grad;
    float* m;
    float* v;
    long long n;
    float lr, beta1, beta2, eps, grad_scale;
    int* step_state;                  // [0] steps applied, [1] block ticket
    int n_segments;
    TailSegment seg[kMaxSegments];
    int rank, world;
    unsigned long long* ll[kMaxWorld];   // peer-mapped mailboxes [2][world][n_pad] of {value, step}
    long long n_pad;
    int* error_flag;                  // set when a peer never shows up (bounded spin)
    const float* stats_partial;       // optional: [stats_splits] x {cost_sum, correct_count, ..} every stats_stride floats
    long long stats_stride;
    int stats_splits;
    float* stats_out;                 // [2]
    int n_blocks;                     // blocks that own parameters; one more block reduces the statistics
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kTailThreads) reduce_adam_kernel(const TailParams p) {
    // All CTAs of this small grid are resident at once, so the dependent launch can be released right away: the next step's
    // first kernel (kgcn_gcn_step_chain_f32 with KGCN_FLAG_INPUTS_STABLE) then loads and aggregates its first tiles while
    // this kernel reduces, exchanges and updates; whatever it reads of this kernel's results comes after its own
    // griddepcontrol.wait.  Kernels that wait at their entry simply park there.
    pdl_launch_dependents();
    pdl_prologue();
    __shared__ float4 red[kTailWarps][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = p.step_state[0] + 1;
    if (static_cast<int>(blockIdx.x) >= p.n_blocks) {
        // ---- the extra block: per-CTA statistics of the fused head -> stats_out, summed in CTA order ----
        if (threadIdx.x < 2) {
            float acc = 0.0f;
            for (int k = 0; k < p.stats_splits; k += 16) {
                float val[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) val[j] = (k + j < p.stats_splits) ? __ldcg(p.stats_partial + static_cast<long long>(k + j) * p.stats_stride + threadIdx.x) : 0.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (k + j < p.stats_splits) acc += val[j];
            }
            p.stats_out[threadIdx.x] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(p.step_state + 1, 1) == static_cast<int>(gridDim.x) - 1) {
                p.step_state[0] = t;
                p.step_state[1] = 0;
            }
        }
        return;
    }
    const long long i = (static_cast<long long>(blockIdx.x) * 32 + lane) * 4;   // first of this lane's 4 elements

    // ---- which gradient are elements i .. i + 3?  (segment offsets and widths are multiples of 4) ----
    const float* src = nullptr;
    long long stride = 0;
    int splits = 0;
    if (i < p.n) {
#pragma unroll 1
        for (int s = 0; s < p.n_segments; ++s) {
            const TailSegment& g = p.seg[s];
            const long long total = g.stride;
            const long long k = i - g.kernel_off, kb = i - g.bias_off;
            if (k >= 0 && k < static_cast<long long>(g.channels) * g.rows * g.cols) {
                const int c = static_cast<int>(k / (static_cast<long long>(g.rows) * g.cols));
                const long long r = k - static_cast<long long>(c) * g.rows * g.cols;
                const int row = static_cast<int>(r / g.cols), col = static_cast<int>(r - static_cast<long long>(row) * g.cols);
                src = g.partial + static_cast<long long>(row) * g.channels * g.cols + c * g.cols + col;
                stride = total;
                splits = g.splits;
            } else if (g.bias_off >= 0 && kb >= 0 && kb < static_cast<long long>(g.channels) * g.cols) {
                src = g.partial + static_cast<long long>(g.rows) * g.channels * g.cols + kb;
                stride = total;
                splits = g.splits;
            }
        }
    }
    // Adam operands are requested before anything is waited for
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), v0 = m0, p0 = m0, gd = m0;
    if (warp == 0 && i < p.n) {
        m0 = *reinterpret_cast<const float4*>(p.m + i);
        v0 = *reinterpret_cast<const float4*>(p.v + i);
        p0 = *reinterpret_cast<const float4*>(p.param + i);
        if (src == nullptr) gd = *reinterpret_cast<const float4*>(p.grad + i);
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src != nullptr) {
        for (int z = warp; z < splits; z += 8 * kTailWarps) {   // predicated batches of 8 independent loads per lane
            float4 val[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                val[j] = (z + kTailWarps * j < splits) ? __ldcg(reinterpret_cast<const float4*>(src + static_cast<long long>(z + kTailWarps * j) * stride))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (z + kTailWarps * j < splits) { s.x += val[j].x; s.y += val[j].y; s.z += val[j].z; s.w += val[j].w; }
        }
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float4 g4 = gd;
        if (src != nullptr) {
            g4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < kTailWarps; ++w) { const float4 r = red[w][lane]; g4.x += r.x; g4.y += r.y; g4.z += r.z; g4.w += r.w; }
        }
        float g[4] = {g4.x, g4.y, g4.z, g4.w};
        if (p.world > 1 && i < p.n) {
            // ---- one-shot all-reduce: push {value, step} into every peer's mailbox, spin on the own one ----
            const long long slot = (static_cast<long long>(t & 1) * p.world) * p.n_pad + i;
            const unsigned long long tag = static_cast<unsigned long long>(static_cast<unsigned>(t)) << 32;
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r)
                if (r < p.world && r != p.rank) {
                    unsigned long long* dst = p.ll[r] + slot + static_cast<long long>(p.rank) * p.n_pad;
#pragma unroll
                    for (int j = 0; j < 4; ++j) st_relaxed_sys_u64(dst + j, tag | __float_as_uint(g[j]));
                }
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            bool ok = true;
            const long long t0 = clock64();
#pragma unroll 1
            for (int r = 0; r < p.world; ++r) {
                if (r == p.rank) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tot[j] += g[j];
                    continue;
                }
                const unsigned long long* mine = p.ll[p.rank] + slot + static_cast<long long>(r) * p.n_pad;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned long long w;
                    while (((w = ld_relaxed_sys_u64(mine + j)) >> 32) != static_cast<unsigned>(t)) {
                        if (clock64() - t0 > 4000000000ll) {   // ~2 s: a peer never launched its step; fail instead of hanging the GPU
                            ok = false;
                            break;
                        }
                    }
                    tot[j] += __uint_as_float(static_cast<unsigned>(w));
                }
            }
            if (!ok && p.error_flag != nullptr) atomicExch(p.error_flag, 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = tot[j];
        }
        if (i < p.n) {
            *reinterpret_cast<float4*>(p.grad + i) = make_float4(g[0], g[1], g[2], g[3]);
            // TF AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
            const float lr_t = p.lr * sqrtf(1.0f - powf(p.beta2, static_cast<float>(t))) / (1.0f - powf(p.beta1, static_cast<float>(t)));
            const float mo[4] = {m0.x, m0.y, m0.z, m0.w}, vo[4] = {v0.x, v0.y, v0.z, v0.w}, po[4] = {p0.x, p0.y, p0.z, p0.w};
            float mn[4], vn[4], pn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float gr = g[j] * p.grad_scale;
                mn[j] = p.beta1 * mo[j] + (1.0f - p.beta1) * gr;
                vn[j] = p.beta2 * vo[j] + (1.0f - p.beta2) * gr * gr;
                pn[j] = po[j] - lr_t * mn[j] / (sqrtf(vn[j]) + p.eps);
            }
            *reinterpret_cast<float4*>(p.m + i) = make_float4(mn[0], mn[1], mn[2], mn[3]);
            *reinterpret_cast<float4*>(p.v + i) = make_float4(vn[0], vn[1], vn[2], vn[3]);
            *reinterpret_cast<float4*>(p.param + i) = make_float4(pn[0], pn[1], pn[2], pn[3]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.step_state + 1, 1) == static_cast<int>(gridDim.x) - 1) {
            p.step_state[0] = t;
            p.step_state[1] = 0;
        }
    }
}

}  // namespace
}  // namespace kgcn

using namespace kgcn;

extern "C" int kgcn_reduce_adam_f32(float* param, float* grad, float* m, float* v, int64_t n, const kgcn_grad_segment* segments,
                                    int32_t n_segments, float lr, float beta1, float beta2, float eps, float grad_scale,
                                    int32_t* step_state, const kgcn_p2p_group* group, const float* stats_partial,
                                    int32_t stats_splits, int64_t stats_stride, float* stats_out, void* stream) {
    KGCN_REQUIRE(param && grad && m && v && step_state, KGCN_ERR_NULL, "reduce_adam: NULL pointer argument");
    KGCN_REQUIRE(n >= 0 && n % 4 == 0 && n_segments >= 0 && n_segments <= kMaxSegments && (n_segments == 0 || segments != nullptr),
                 KGCN_ERR_BAD_SHAPE, "reduce_adam: bad n (a multiple of 4) / n_segments (at most %d segments)", kMaxSegments);
    KGCN_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(m) && aligned16(v), KGCN_ERR_MISALIGNED,
                 "reduce_adam: flat buffers must be 16-byte aligned");
    if (n == 0) return KGCN_OK;
    TailParams p{};
    p.param = param; p.grad = grad; p.m = m; p.v = v; p.n = n;
    p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.grad_scale = grad_scale;
    p.step_state = step_state;
    p.n_segments = n_segments;
    for (int s = 0; s < n_segments; ++s) {
        const kgcn_grad_segment& g = segments[s];
        KGCN_REQUIRE(g.partial != nullptr && g.splits > 0 && g.rows >= 0 && g.cols > 0 && g.channels > 0 && g.kernel_off >= 0 &&
                         g.kernel_off + static_cast<int64_t>(g.channels) * g.rows * g.cols <= n &&
                         (g.bias_off < 0 || g.bias_off + static_cast<int64_t>(g.channels) * g.cols <= n),
                     KGCN_ERR_BAD_SHAPE, "reduce_adam: segment %d is out of range", s);
        KGCN_REQUIRE(g.cols % 4 == 0 && g.kernel_off % 4 == 0 && (g.bias_off < 0 || g.bias_off % 4 == 0) && aligned16(g.partial),
                     KGCN_ERR_MISALIGNED, "reduce_adam: segment %d: offsets and width must be multiples of 4 floats", s);
        const long long dflt = static_cast<long long>(g.rows + 1) * g.channels * g.cols;
        KGCN_REQUIRE(g.stride == 0 || (g.stride >= dflt && g.stride % 4 == 0), KGCN_ERR_BAD_SHAPE, "reduce_adam: segment %d: bad stride", s);
        p.seg[s] = TailSegment{g.kernel_off, g.bias_off, g.partial, g.splits, g.rows, g.cols, g.channels, g.stride ? g.stride : dflt};
    }
    unsigned blocks = static_cast<unsigned>(ceil_div<int64_t>(n, kTailElems));
    p.n_blocks = static_cast<int>(blocks);
    if (stats_partial != nullptr) {
        KGCN_REQUIRE(stats_out != nullptr && stats_splits > 0 && stats_stride >= 2, KGCN_ERR_BAD_SHAPE, "reduce_adam: bad statistics partials");
        p.stats_partial = stats_partial; p.stats_splits = stats_splits; p.stats_stride = stats_stride; p.stats_out = stats_out;
        ++blocks;
    }
    p.rank = 0;
    p.world = 1;
    if (group != nullptr && group->world > 1) {
        KGCN_REQUIRE(group->world <= kMaxWorld && group->rank >= 0 && group->rank < group->world, KGCN_ERR_BAD_SHAPE,
                     "reduce_adam: bad rank %d / world %d (at most %d ranks)", group->rank, group->world, kMaxWorld);
        KGCN_REQUIRE(group->n_pad >= n, KGCN_ERR_WORKSPACE, "reduce_adam: mailboxes too small (%lld slots per rank)", (long long)group->n_pad);
        p.rank = group->rank;
        p.world = group->world;
        p.n_pad = group->n_pad;
        p.error_flag = group->error_flag;
        for (int r = 0; r < group->world; ++r) {
            KGCN_REQUIRE(group->mailbox[r] != nullptr, KGCN_ERR_NULL, "reduce_adam: peer %d is not mapped", r);
            p.ll[r] = reinterpret_cast<unsigned long long*>(group->mailbox[r]);
        }
    }
    launch_pdl(reduce_adam_kernel, blocks, kTailThreads, 0, static_cast<cudaStream_t>(stream), p);
    KGCN_LAUNCH_OK("reduce_adam_kernel");
    return KGCN_OK;
}

// ---- peer-mapped buffers (cudaIpc): plain cudaMalloc allocations, so the handle names exactly this buffer ----
extern "C" int kgcn_p2p_alloc(size_t n_bytes, void** device_ptr, unsigned char* handle64) {
    KGCN_REQUIRE(device_ptr != nullptr && handle64 != nullptr && n_bytes > 0, KGCN_ERR_NULL, "p2p_alloc: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void* ptr = nullptr;
    KGCN_CUDA_OK(cudaMalloc(&ptr, n_bytes));
    KGCN_CUDA_OK(cudaMemset(ptr, 0, n_bytes));
    cudaIpcMemHandle_t h;
    KGCN_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    *device_ptr = ptr;
    return KGCN_OK;
}

extern "C" int kgcn_p2p_open(const unsigned char* handle64, void** device_ptr) {
    KGCN_REQUIRE(device_ptr != nullptr && handle64 != nullptr, KGCN_ERR_NULL, "p2p_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    KGCN_CUDA_OK(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KGCN_OK;
}

extern "C" int kgcn_p2p_close(void* device_ptr) {
    if (device_ptr != nullptr) KGCN_CUDA_OK(cudaIpcCloseMemHandle(device_ptr));
    return KGCN_OK;
}

extern "C" int kgcn_p2p_free(void* device_ptr) {
    if (device_ptr != nullptr) KGCN_CUDA_OK(cudaFree(device_ptr));
    return KGCN_OK;
}
