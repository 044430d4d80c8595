// The tail of a training step in ONE launch: fixed-order reduction of the per-CTA weight-gradient partials of every
// GraphConv layer -> (data parallel: one-shot all-reduce over NVLink peer memory) -> Adam.
//
//   reduce     graphconv_fused_dw_kernel leaves one partial block [(f_in + 1), C * f_out] per CTA; element i of the flat
//              parameter buffer sums its `splits` partials in split order (32 warps take splits w, w + 32, ..; the 32 warp
//              sums are then added in warp order), exactly like splitk_reduce_kernel, so the result is deterministic.
//              Parameters outside every segment (readout head, GraphDense) already have their gradient in `grad`.
//   all-reduce every rank owns a peer-mapped exchange buffer (cudaIpc, kgcn_p2p_*): block b writes its 32 local sums,
//              publishes flag[b] = step (release, system scope), then lanes 0..W-1 poll the W ranks' flag[b] and all lanes
//              read the W ranks' values in ONE NVLink round trip and add them in rank order -- every rank computes
//              bit-identical sums.  The buffer is double-buffered by step parity: a rank passes the wait of step t only
//              after every peer has published step t, i.e. has finished reading step t - 1.
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
constexpr int kTailThreads = 1024;

struct TailSegment {
    long long kernel_off, bias_off;   // offsets into the flat buffers; kernel [C][rows][cols], bias [C][cols] (-1: none)
    const float* partial;             // [splits][(rows + 1)][C * cols]
    int splits, rows, cols, channels;
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
    float* xg[kMaxWorld];             // peer-mapped exchange buffers [2][n_pad]
    unsigned* flags[kMaxWorld];       // peer-mapped flags [n_blocks]
    long long n_pad;
    int* error_flag;                  // set when a peer never shows up (bounded spin)
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kTailThreads) reduce_adam_kernel(const TailParams p) {
    pdl_prologue();
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = static_cast<long long>(blockIdx.x) * 32 + lane;
    const int t = p.step_state[0] + 1;

    // ---- which gradient is element i? ----
    const float* src = nullptr;
    long long stride = 0;
    int splits = 0;
    if (i < p.n) {
#pragma unroll 1
        for (int s = 0; s < p.n_segments; ++s) {
            const TailSegment& g = p.seg[s];
            const long long total = static_cast<long long>(g.rows + 1) * g.channels * g.cols;
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
    float m0 = 0.0f, v0 = 0.0f, p0 = 0.0f, gdirect = 0.0f;
    if (warp == 0 && i < p.n) {
        m0 = p.m[i];
        v0 = p.v[i];
        p0 = p.param[i];
        if (src == nullptr) gdirect = p.grad[i];
    }
    float s = 0.0f;
    if (src != nullptr) {
        for (int z = warp; z < splits; z += 256) {   // predicated batches of 8: up to 256 partials in ONE round trip
            float val[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) val[j] = (z + 32 * j < splits) ? __ldcg(src + static_cast<long long>(z + 32 * j) * stride) : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (z + 32 * j < splits) s += val[j];
        }
    }
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0) {
        float g = gdirect;
        if (src != nullptr) {
            g = 0.0f;
#pragma unroll
            for (int w = 0; w < 32; ++w) g += red[w][lane];
        }
        if (p.world > 1) {
            // ---- one-shot all-reduce over peer memory ----
            float* mine = p.xg[p.rank] + static_cast<long long>(t & 1) * p.n_pad;
            if (i < p.n_pad) mine[i] = (i < p.n) ? g : 0.0f;
            __threadfence_system();
            __syncwarp();
            if (lane == 0) st_release_sys(p.flags[p.rank] + blockIdx.x, static_cast<unsigned>(t));
            bool ok = true;
            if (lane < p.world && lane != p.rank) {
                const unsigned* f = p.flags[lane] + blockIdx.x;
                const long long t0 = clock64();
                while (static_cast<int>(ld_acquire_sys(f)) < t) {
                    if (clock64() - t0 > 4000000000ll) {   // ~2 s: a peer never launched its step; fail instead of hanging the GPU
                        ok = false;
                        break;
                    }
                }
            }
            ok = __all_sync(0xffffffffu, ok);
            __threadfence_system();
            if (!ok) {
                if (lane == 0 && p.error_flag != nullptr) atomicExch(p.error_flag, 1);
            } else if (i < p.n) {
                float pv[kMaxWorld];
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r)
                    pv[r] = (r < p.world && r != p.rank) ? ld_relaxed_sys(p.xg[r] + static_cast<long long>(t & 1) * p.n_pad + i) : 0.0f;
                float tot = 0.0f;
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r)
                    if (r < p.world) tot += (r == p.rank) ? g : pv[r];
                g = tot;
            }
        }
        if (i < p.n) {
            p.grad[i] = g;
            // TF AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
            const float lr_t = p.lr * sqrtf(1.0f - powf(p.beta2, static_cast<float>(t))) / (1.0f - powf(p.beta1, static_cast<float>(t)));
            const float gr = g * p.grad_scale;
            const float mi = p.beta1 * m0 + (1.0f - p.beta1) * gr;
            const float vi = p.beta2 * v0 + (1.0f - p.beta2) * gr * gr;
            p.m[i] = mi;
            p.v[i] = vi;
            p.param[i] = p0 - lr_t * mi / (sqrtf(vi) + p.eps);
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
                                    int32_t* step_state, const kgcn_p2p_group* group, void* stream) {
    KGCN_REQUIRE(param && grad && m && v && step_state, KGCN_ERR_NULL, "reduce_adam: NULL pointer argument");
    KGCN_REQUIRE(n >= 0 && n_segments >= 0 && n_segments <= kMaxSegments && (n_segments == 0 || segments != nullptr), KGCN_ERR_BAD_SHAPE,
                 "reduce_adam: bad n / n_segments (at most %d segments)", kMaxSegments);
    if (n == 0) return KGCN_OK;
    TailParams p{};
    p.param = param; p.grad = grad; p.m = m; p.v = v; p.n = n;
    p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.grad_scale = grad_scale;
    p.step_state = step_state;
    p.n_segments = n_segments;
    for (int s = 0; s < n_segments; ++s) {
        const kgcn_grad_segment& g = segments[s];
        KGCN_REQUIRE(g.partial != nullptr && g.splits > 0 && g.rows > 0 && g.cols > 0 && g.channels > 0 && g.kernel_off >= 0 &&
                         g.kernel_off + static_cast<int64_t>(g.channels) * g.rows * g.cols <= n &&
                         (g.bias_off < 0 || g.bias_off + static_cast<int64_t>(g.channels) * g.cols <= n),
                     KGCN_ERR_BAD_SHAPE, "reduce_adam: segment %d is out of range", s);
        p.seg[s] = TailSegment{g.kernel_off, g.bias_off, g.partial, g.splits, g.rows, g.cols, g.channels};
    }
    const unsigned blocks = static_cast<unsigned>(ceil_div<int64_t>(n, 32));
    p.rank = 0;
    p.world = 1;
    if (group != nullptr && group->world > 1) {
        KGCN_REQUIRE(group->world <= kMaxWorld && group->rank >= 0 && group->rank < group->world, KGCN_ERR_BAD_SHAPE,
                     "reduce_adam: bad rank %d / world %d (at most %d ranks)", group->rank, group->world, kMaxWorld);
        KGCN_REQUIRE(group->n_pad >= n && group->n_flags >= static_cast<int64_t>(blocks), KGCN_ERR_WORKSPACE,
                     "reduce_adam: exchange buffers too small (%lld floats, %lld flags)", (long long)group->n_pad, (long long)group->n_flags);
        p.rank = group->rank;
        p.world = group->world;
        p.n_pad = group->n_pad;
        p.error_flag = group->error_flag;
        for (int r = 0; r < group->world; ++r) {
            KGCN_REQUIRE(group->xg[r] != nullptr && group->flags[r] != nullptr, KGCN_ERR_NULL, "reduce_adam: peer %d is not mapped", r);
            p.xg[r] = group->xg[r];
            p.flags[r] = group->flags[r];
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
