// Step-loop kernels around the graph layers: the readout head of the shipped classifiers
// (example_model/model.py:56-69: Dense(label_dim) -> softmax -> mask * softmax_cross_entropy ->
// reduce_mean / reduce_sum / correct_count) with its backward, and Adam (kgcn/core.py:121-127,
// TF defaults beta1 .9, beta2 .999, eps 1e-8) over one flat parameter buffer.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace kgcn {
namespace {

constexpr int kMaxLabels = 32;

// One kernel for the whole head.  Each block owns a contiguous slice of the batch:
//   phase 0: w, labels, mask of the slice -> shared memory; the gathered rows g (either read, or formed here from the
//            node rows: FUSE_GATHER) -> shared memory.  All of these loads are independent, so the head pays ONE
//            global round trip for its inputs instead of a chain of dependent ones;
//   phase 1 (warp per graph): logits, softmax, masked cross-entropy, d logits, d gathered;
//   phase 2 (thread per weight): d out_w[f,l] / d out_b[l] over the slice, in graph order;
// block partials go to `partial`; the last block to finish (ticket in state[2]) adds them in a fixed order (R helper
// threads per output over contiguous block ranges, then the R sums in helper order), so every sum is deterministic
// and no memset / second launch is needed.
constexpr int kReadoutThreads = 512;
constexpr int kReadoutMaxSlice = 128;

// FUSE_GATHER: the GraphGather sums (layers.py:164, rows added in index order exactly like gather_fwd_kernel) are formed
// here from the node rows x_nodes [n_graphs, n_nodes, feat] and written to g (an OUTPUT then) -- one launch and one pass
// over the last layer's activations less per step.
template <bool FUSE_GATHER>
__global__ void __launch_bounds__(kReadoutThreads, 1) readout_kernel(
    const float* __restrict__ x_nodes, int n_nodes,
    float* g, int64_t n_graphs, int feat, const float* __restrict__ w, const float* __restrict__ bias,
    int n_labels, const float* __restrict__ labels, const float* __restrict__ mask, float inv_batch,
    float* __restrict__ logits, float* __restrict__ prediction, float* __restrict__ dlogits, float* __restrict__ dg,
    float* __restrict__ dw, float* __restrict__ dbias, float* __restrict__ partial, float* __restrict__ state,
    float* __restrict__ du_nodes, int act) {
    pdl_prologue();
    extern __shared__ __align__(16) float rd_smem[];   // g_s [slice][feat] | w_s [feat][n_labels] | y_s [slice][n_labels] | m_s [slice]
    __shared__ float dz_s[kReadoutMaxSlice * kMaxLabels];
    __shared__ float cost_s[kReadoutMaxSlice], corr_s[kReadoutMaxSlice];
    __shared__ float red_s[kReadoutThreads];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n_graphs + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = blockIdx.x * per;
    const int n_here = static_cast<int>(max(static_cast<int64_t>(0), min(n_graphs, b0 + per) - b0));
    float* g_s = rd_smem;
    float* w_s = g_s + static_cast<size_t>(per) * feat;
    float* y_s = w_s + static_cast<size_t>(feat) * n_labels;
    float* m_s = y_s + static_cast<size_t>(per) * n_labels;

    // ---- phase 0 ----
    for (int i = threadIdx.x; i < feat * n_labels; i += kReadoutThreads) w_s[i] = w[i];
    for (int i = threadIdx.x; i < n_here * n_labels; i += kReadoutThreads) y_s[i] = labels[b0 * n_labels + i];
    for (int i = threadIdx.x; i < n_here; i += kReadoutThreads) m_s[i] = mask ? mask[b0 + i] : 1.0f;
    if (FUSE_GATHER) {
        const int lpr = feat >> 2;   // lanes per node row when every lane owns 4 consecutive features
        if ((feat & 3) == 0 && lpr <= 32 && (32 % lpr) == 0 && (reinterpret_cast<uintptr_t>(x_nodes) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
            // 32 / lpr graphs per warp; a lane walks ALL node rows of its 4 features in index order (the sum stays
            // bit-identical to gather_fwd_kernel), 16 independent 16-byte loads in flight per lane: the phase is one
            // or two memory round trips instead of n_nodes / 4 dependent ones.
            const int gpw = 32 / lpr, sub = lane / lpr, fl = (lane - sub * lpr) * 4;
            for (int i0 = warp * gpw; i0 < n_here; i0 += (kReadoutThreads / 32) * gpw) {
                const int i = i0 + sub;
                if (i >= n_here) continue;
                const float4* xb = reinterpret_cast<const float4*>(x_nodes + (b0 + i) * n_nodes * feat + fl);
                float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                int r = 0;
                for (; r + 16 <= n_nodes; r += 16) {
                    float4 v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = xb[static_cast<int64_t>(r + j) * lpr];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
                    }
                }
                for (; r < n_nodes; ++r) {
                    const float4 v = xb[static_cast<int64_t>(r) * lpr];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
                *reinterpret_cast<float4*>(g_s + i * feat + fl) = acc;
                *reinterpret_cast<float4*>(g + (b0 + i) * feat + fl) = acc;
            }
        } else {   // any width: warp per graph, lane owns features lane, lane + 32, ...
            for (int i = warp; i < n_here; i += kReadoutThreads / 32) {
                const float* xb = x_nodes + (b0 + i) * n_nodes * feat;
                for (int f = lane; f < feat; f += 32) {
                    float acc = 0.0f;
                    for (int r = 0; r < n_nodes; ++r) acc += xb[static_cast<int64_t>(r) * feat + f];
                    g_s[i * feat + f] = acc;
                    g[(b0 + i) * feat + f] = acc;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < n_here * feat; i += kReadoutThreads) g_s[i] = g[b0 * feat + i];
    }
    __syncthreads();

    // ---- phase 1 ----
    for (int i = warp; i < n_here; i += kReadoutThreads / 32) {
        const int64_t b = b0 + i;
        const float* gb = g_s + i * feat;
        const float* yb = y_s + i * n_labels;
        float z[kMaxLabels];
#pragma unroll 1
        for (int l = 0; l < n_labels; ++l) {
            float acc = 0.0f;
            for (int f = lane; f < feat; f += 32) acc = fmaf(gb[f], w_s[f * n_labels + l], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            z[l] = acc + (bias ? bias[l] : 0.0f);
        }
        float zmax = z[0];
        for (int l = 1; l < n_labels; ++l) zmax = fmaxf(zmax, z[l]);
        float sum = 0.0f;
        for (int l = 0; l < n_labels; ++l) sum += expf(z[l] - zmax);
        const float lse = logf(sum) + zmax;
        const float m = m_s[i];
        float cost = 0.0f, ysum = 0.0f;
        int arg_p = 0, arg_y = 0;
        for (int l = 0; l < n_labels; ++l) {
            const float y = yb[l];
            cost -= y * (z[l] - lse);            // softmax cross-entropy with (possibly soft) labels
            ysum += y;
            if (z[l] > z[arg_p]) arg_p = l;
            if (y > yb[arg_y]) arg_y = l;
        }
        float dz[kMaxLabels];
        for (int l = 0; l < n_labels; ++l) {
            const float pr = expf(z[l] - lse);
            dz[l] = m * inv_batch * (pr * ysum - yb[l]);
            if (lane == 0) {
                if (logits) logits[b * n_labels + l] = z[l];
                if (prediction) prediction[b * n_labels + l] = pr;
                if (dlogits) dlogits[b * n_labels + l] = dz[l];
                dz_s[i * n_labels + l] = dz[l];
            }
        }
        if (dg != nullptr)
            for (int f = lane; f < feat; f += 32) {
                float acc = 0.0f;
                for (int l = 0; l < n_labels; ++l) acc = fmaf(dz[l], w_s[f * n_labels + l], acc);
                dg[b * feat + f] = acc;
            }
        if (lane == 0) {
            cost_s[i] = m * cost;                                   // cost_sum term   (model.py:64)
            corr_s[i] = m * (arg_p == arg_y ? 1.0f : 0.0f);         // correct_count term (model.py:66-69)
        }
    }
    __syncthreads();

    // ---- phase 1b (training step): dU of the last graph layer, du[b, r, :] = dg[b, :] (.) act'(x[b, r, :]) -- the
    // GraphGather gradient (layers.py:164, a broadcast over the node rows) times the activation gradient, so the layer's
    // backward needs no separate activation-gradient pass.  The rows were read a moment ago by the gather phase (L2 hits).
    if (FUSE_GATHER && du_nodes != nullptr && dg != nullptr) {
        const int64_t row_elems = static_cast<int64_t>(n_nodes) * feat;
        if ((feat & 3) == 0 && (reinterpret_cast<uintptr_t>(x_nodes) & 15) == 0 && (reinterpret_cast<uintptr_t>(du_nodes) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(dg) & 15) == 0) {
            const int f4 = feat >> 2;
            const int64_t n4 = static_cast<int64_t>(n_here) * n_nodes * f4;
            const float4* xs = reinterpret_cast<const float4*>(x_nodes + b0 * row_elems);
            float4* us = reinterpret_cast<float4*>(du_nodes + b0 * row_elems);
            const float4* dgs = reinterpret_cast<const float4*>(dg + b0 * feat);
            // batches of 8 independent 16-byte loads per thread: the pass is one or two memory round trips, not n4 / 512
            for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 8 * kReadoutThreads) {
                float4 xv[8], gv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int64_t idx = i0 + static_cast<int64_t>(u) * kReadoutThreads;
                    if (idx < n4) {
                        const int64_t row = idx / f4;
                        xv[u] = xs[idx];
                        gv[u] = dgs[static_cast<int>(row / n_nodes) * f4 + static_cast<int>(idx - row * f4)];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int64_t idx = i0 + static_cast<int64_t>(u) * kReadoutThreads;
                    if (idx < n4)
                        us[idx] = make_float4(gv[u].x * act_grad_from_output(xv[u].x, act), gv[u].y * act_grad_from_output(xv[u].y, act),
                                              gv[u].z * act_grad_from_output(xv[u].z, act), gv[u].w * act_grad_from_output(xv[u].w, act));
                }
            }
        } else {
            const int64_t n1 = static_cast<int64_t>(n_here) * row_elems;
            for (int64_t idx = threadIdx.x; idx < n1; idx += kReadoutThreads) {
                const int64_t row = idx / feat;
                const int f = static_cast<int>(idx - row * feat), i = static_cast<int>(row / n_nodes);
                du_nodes[b0 * row_elems + idx] = dg[(b0 + i) * feat + f] * act_grad_from_output(x_nodes[b0 * row_elems + idx], act);
            }
        }
    }

    // ---- phase 2: block partials, graph order inside the slice ----
    const int n_w = (feat + 1) * n_labels;      // row `feat` is the bias gradient
    const int n_out = n_w + 2;                  // + cost_sum, correct_count
    float* my = partial + static_cast<size_t>(blockIdx.x) * n_out;
    for (int o = threadIdx.x; o < n_out; o += kReadoutThreads) {
        float acc = 0.0f;
        if (o < n_w) {
            if (dw != nullptr) {
                const int f = o / n_labels, l = o - f * n_labels;
                if (f < feat)
                    for (int i = 0; i < n_here; ++i) acc = fmaf(g_s[i * feat + f], dz_s[i * n_labels + l], acc);
                else
                    for (int i = 0; i < n_here; ++i) acc += dz_s[i * n_labels + l];
            }
        } else {
            const float* src = (o == n_w) ? cost_s : corr_s;
            for (int i = 0; i < n_here; ++i) acc += src[i];
        }
        my[o] = acc;
    }
    __threadfence();
    __syncthreads();
    int* ticket = reinterpret_cast<int*>(state + 2);
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- final sum over the block partials: R helpers per output over contiguous block ranges ----
    const int nb = static_cast<int>(gridDim.x);
    const int R = max(1, min(4, kReadoutThreads / n_out));
    const int chunk = (nb + R - 1) / R;
    for (int o0 = 0; o0 < n_out; o0 += kReadoutThreads / R) {
        const int slot = threadIdx.x, h = slot / (kReadoutThreads / R), o = o0 + slot - h * (kReadoutThreads / R);
        float acc = 0.0f;
        if (o < n_out && h < R) {
            const float* src = partial + o;
            const int k_end = min(nb, h * chunk + chunk);
            // predicated batches of 32 independent loads, added in block order: a helper's whole range (<= 50 blocks
            // at 148 CTAs and R = 3) costs one or two L2 round trips, never a chain of dependent ones
            for (int k = h * chunk; k < k_end; k += 32) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (k + j < k_end) ? __ldcg(src + static_cast<size_t>(k + j) * n_out) : 0.0f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (k + j < k_end) acc += v[j];
            }
        }
        red_s[slot] = acc;
        __syncthreads();
        if (h == 0 && o < n_out) {
            acc = red_s[slot];
            for (int r = 1; r < R; ++r) acc += red_s[slot + r * (kReadoutThreads / R)];
            if (o < n_w) {
                const int f = o / n_labels, l = o - f * n_labels;
                if (dw != nullptr) {
                    if (f < feat) dw[o] = acc;
                    else if (dbias != nullptr) dbias[l] = acc;
                }
            } else {
                state[o - n_w] = acc;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// step_state (device, may be null): [0] = number of steps already applied, [1] = block ticket.
// Reading the step on the device keeps the launch parameters constant, so the whole training step
// can be replayed from a CUDA graph; the last block to finish advances the counter.
// d out_w[f, l] = sum_b g[b, f] * dz[b, l], d out_b[l] = sum_b dz[b, l] for the tiny readout head.
// Each block reduces a slice of the batch into `partial[block]`; the last block to finish (ticket)
// sums the partials in block order, so the result is deterministic.  ticket must be zero on entry
// and is reset on exit.
__global__ void __launch_bounds__(256) readout_dw_kernel(const float* __restrict__ g, const float* __restrict__ dz,
                                                         int64_t n_graphs, int feat, int n_labels,
                                                         float* __restrict__ partial, int* __restrict__ ticket,
                                                         float* __restrict__ dw, float* __restrict__ dbias) {
    pdl_prologue();
    const int n_out = (feat + 1) * n_labels;   // row `feat` is the bias gradient
    const int64_t per = (n_graphs + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = blockIdx.x * per, b1 = min(n_graphs, b0 + per);
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        const int f = o / n_labels, l = o - f * n_labels;
        float acc = 0.0f;
        if (f < feat)
            for (int64_t b = b0; b < b1; ++b) acc = fmaf(g[b * feat + f], dz[b * n_labels + l], acc);
        else
            for (int64_t b = b0; b < b1; ++b) acc += dz[b * n_labels + l];
        partial[static_cast<size_t>(blockIdx.x) * n_out + o] = acc;
    }
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1) == static_cast<int>(gridDim.x) - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
        float acc = 0.0f;
        {   // fixed order, 16 independent loads in flight (this loop is a chain of L2 round trips, nothing else)
            const float* src = partial + o;
            const unsigned nb = gridDim.x;
            unsigned k = 0;
            for (; k + 16 <= nb; k += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldcg(src + static_cast<size_t>(k + j) * n_out);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += v[j];
            }
            for (; k < nb; ++k) acc += __ldcg(src + static_cast<size_t>(k) * n_out);
        }
        const int f = o / n_labels, l = o - f * n_labels;
        if (f < feat) dw[o] = acc;
        else if (dbias != nullptr) dbias[l] = acc;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

__global__ void adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps,
                            float grad_scale, int host_step, int* __restrict__ step_state) {
    pdl_prologue();
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // this thread's first element is requested BEFORE the step counter is needed, so the counter and the operands
    // arrive in one memory round trip (the whole parameter buffer is a few KB: the kernel is that round trip)
    float g0 = 0.0f, m0 = 0.0f, v0 = 0.0f, p0 = 0.0f;
    if (i < n) {
        g0 = grad[i];
        m0 = m[i];
        v0 = v[i];
        p0 = param[i];
    }
    const int t = step_state != nullptr ? step_state[0] + 1 : host_step;
    // TF AdamOptimizer: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    const float lr_t = lr * sqrtf(1.0f - powf(beta2, static_cast<float>(t))) / (1.0f - powf(beta1, static_cast<float>(t)));
    for (; i < n; i += stride) {
        const float gr = g0 * grad_scale;
        const float mi = beta1 * m0 + (1.0f - beta1) * gr;
        const float vi = beta2 * v0 + (1.0f - beta2) * gr * gr;
        m[i] = mi;
        v[i] = vi;
        param[i] = p0 - lr_t * mi / (sqrtf(vi) + eps);
        const int64_t nx = i + stride;
        if (nx < n) {
            g0 = grad[nx];
            m0 = m[nx];
            v0 = v[nx];
            p0 = param[nx];
        }
    }
    if (step_state != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(step_state + 1, 1) == static_cast<int>(gridDim.x) - 1) {
                step_state[0] = t;
                step_state[1] = 0;
            }
        }
    }
}

}  // namespace
}  // namespace kgcn

using namespace kgcn;

static int readout_blocks(int64_t n_graphs) {
    // every SM takes a slice: the node-row passes (GraphGather sums, the dU rows of the training step) move the whole
    // activation tensor and want all of them; phase 1 (a warp per graph) is tiny either way
    int64_t nb = std::min<int64_t>(n_graphs, kNumSMs);
    nb = std::max<int64_t>(nb, ceil_div<int64_t>(n_graphs, kReadoutMaxSlice));
    return static_cast<int>(nb);
}

static size_t readout_dyn_smem(int64_t n_graphs, int nb, int feat, int n_labels) {
    const size_t slice = static_cast<size_t>(ceil_div<int64_t>(n_graphs, nb));
    return (slice * (static_cast<size_t>(feat) + n_labels + 1) + static_cast<size_t>(feat) * n_labels) * sizeof(float);
}

extern "C" size_t kgcn_readout_workspace_bytes(int64_t n_graphs, int32_t feat, int32_t n_labels) {
    if (n_graphs <= 0) return 0;
    return static_cast<size_t>(readout_blocks(n_graphs)) * ((static_cast<size_t>(feat) + 1) * n_labels + 2) * sizeof(float);
}

extern "C" int kgcn_readout_xent_f32(const float* g, int64_t n_graphs, int32_t feat, const float* w, const float* bias,
                                     int32_t n_labels, const float* labels, const float* mask, float inv_batch,
                                     float* logits, float* prediction, float* stats, float* dlogits, float* dg,
                                     float* dw, float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
    KGCN_REQUIRE(g && w && labels && stats, KGCN_ERR_NULL, "readout_xent: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && feat > 0 && n_labels > 0 && n_labels <= kMaxLabels, KGCN_ERR_BAD_SHAPE,
                 "readout_xent: bad shape (n_labels <= %d)", kMaxLabels);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= kgcn_readout_workspace_bytes(n_graphs, feat, n_labels),
                 KGCN_ERR_WORKSPACE, "readout_xent: workspace too small");
    const int nb = readout_blocks(n_graphs);
    KGCN_REQUIRE(ceil_div<int64_t>(n_graphs, nb) <= kReadoutMaxSlice, KGCN_ERR_UNSUPPORTED,
                 "readout_xent: batch too large for one launch (%lld graphs)", (long long)n_graphs);
    const size_t dyn = readout_dyn_smem(n_graphs, nb, feat, n_labels);
    KGCN_REQUIRE(dyn <= 160 * 1024, KGCN_ERR_UNSUPPORTED, "readout_xent: feature width %d too large for the staged head", feat);
    KGCN_CUDA_OK(cudaFuncSetAttribute(readout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    launch_pdl(readout_kernel<false>, nb, kReadoutThreads, dyn, static_cast<cudaStream_t>(stream), static_cast<const float*>(nullptr), 0,
               const_cast<float*>(g), n_graphs, feat, w, bias, n_labels, labels, mask, inv_batch, logits, prediction, dlogits, dg,
               dw, dbias, static_cast<float*>(workspace), stats, static_cast<float*>(nullptr), 0);
    KGCN_LAUNCH_OK("readout_kernel");
    return KGCN_OK;
}

static int gather_readout_impl(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* g,
                                            const float* w, const float* bias, int32_t n_labels, const float* labels,
                                            const float* mask, float inv_batch, float* logits, float* prediction,
                                            float* stats, float* dlogits, float* dg, float* dw, float* dbias,
                                            void* workspace, size_t workspace_bytes, void* stream, int32_t act, float* du_nodes) {
    KGCN_REQUIRE(x && g && w && labels && stats, KGCN_ERR_NULL, "gather_readout_xent: NULL pointer argument");
    KGCN_REQUIRE(n_graphs > 0 && n_nodes > 0 && feat > 0 && n_labels > 0 && n_labels <= kMaxLabels, KGCN_ERR_BAD_SHAPE,
                 "gather_readout_xent: bad shape (n_labels <= %d)", kMaxLabels);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= kgcn_readout_workspace_bytes(n_graphs, feat, n_labels),
                 KGCN_ERR_WORKSPACE, "gather_readout_xent: workspace too small");
    const int nb = readout_blocks(n_graphs);
    KGCN_REQUIRE(ceil_div<int64_t>(n_graphs, nb) <= kReadoutMaxSlice, KGCN_ERR_UNSUPPORTED,
                 "gather_readout_xent: batch too large for one launch (%lld graphs)", (long long)n_graphs);
    const size_t dyn = readout_dyn_smem(n_graphs, nb, feat, n_labels);
    KGCN_REQUIRE(dyn <= 160 * 1024, KGCN_ERR_UNSUPPORTED, "gather_readout_xent: feature width %d too large for the staged head", feat);
    KGCN_CUDA_OK(cudaFuncSetAttribute(readout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    launch_pdl(readout_kernel<true>, nb, kReadoutThreads, dyn, static_cast<cudaStream_t>(stream), x, static_cast<int>(n_nodes), g,
               n_graphs, feat, w, bias, n_labels, labels, mask, inv_batch, logits, prediction, dlogits, dg, dw, dbias,
               static_cast<float*>(workspace), stats, du_nodes, static_cast<int>(act));
    KGCN_LAUNCH_OK("readout_kernel(gather)");
    return KGCN_OK;
}

extern "C" int kgcn_gather_readout_xent_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* g,
                                            const float* w, const float* bias, int32_t n_labels, const float* labels,
                                            const float* mask, float inv_batch, float* logits, float* prediction,
                                            float* stats, float* dlogits, float* dg, float* dw, float* dbias,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    return gather_readout_impl(x, n_graphs, n_nodes, feat, g, w, bias, n_labels, labels, mask, inv_batch, logits, prediction, stats,
                               dlogits, dg, dw, dbias, workspace, workspace_bytes, stream, KGCN_ACT_NONE, nullptr);
}

extern "C" int kgcn_gather_readout_xent_du_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* g,
                                               const float* w, const float* bias, int32_t n_labels, const float* labels,
                                               const float* mask, float inv_batch, float* logits, float* prediction,
                                               float* stats, float* dlogits, float* dg, float* dw, float* dbias, int32_t act,
                                               float* du_nodes, void* workspace, size_t workspace_bytes, void* stream) {
    KGCN_REQUIRE(du_nodes != nullptr && dg != nullptr, KGCN_ERR_NULL, "gather_readout_xent_du: du_nodes and dg are required");
    KGCN_REQUIRE(act >= KGCN_ACT_NONE && act <= KGCN_ACT_TANH, KGCN_ERR_BAD_SHAPE, "gather_readout_xent_du: unknown act %d", act);
    return gather_readout_impl(x, n_graphs, n_nodes, feat, g, w, bias, n_labels, labels, mask, inv_batch, logits, prediction, stats,
                               dlogits, dg, dw, dbias, workspace, workspace_bytes, stream, act, du_nodes);
}

extern "C" int kgcn_adam_f32(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int64_t step, float grad_scale, int32_t* step_state, void* stream) {
    KGCN_REQUIRE(param && grad && m && v, KGCN_ERR_NULL, "adam: NULL pointer argument");
    KGCN_REQUIRE(n >= 0 && (step_state != nullptr || step >= 1), KGCN_ERR_BAD_SHAPE, "adam: bad n/step");
    if (n == 0) return KGCN_OK;
    const int64_t blocks = std::min<int64_t>(ceil_div<int64_t>(n, 256), kNumSMs * 8);
    launch_pdl(adam_kernel, static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), 
        param, grad, m, v, n, lr, beta1, beta2, eps, grad_scale, static_cast<int>(step), step_state);
    KGCN_LAUNCH_OK("adam_kernel");
    return KGCN_OK;
}
