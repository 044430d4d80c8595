// Fused GraphConv layer for sm_100a: ONE kernel emits the next layer's features.
//
//     y[g] = act( sum_c (A[g,c] . x[g]) . W_c  +  rowsum(A[g,c]) (x) bias_c )
//
// which equals the reference's sum_c A[g,c].(x[g].W_c + bias_c) (kgcn/layers.py:105-116) re-associated
// "aggregate first" (the bias reaches a node once per incident entry, so it
// is scaled by the row sum of the adjacency values; rows without entries output act(0)).
//
// Per persistent CTA, per tile of G whole graphs (G*N <= 128 rows):
//   1. TMA: one bulk-async copy lands the tile's [G*N, F_in] feature rows in shared memory
//      (cp.async.bulk + mbarrier complete_tx); all threads stage the tile's CSR slice meanwhile.
//   2. CUDA cores: per-graph neighbour aggregation Z = A.X as a segmented sum out of shared memory
//      (lane group per row, broadcast LDS.64 {offset,value} + LDS.128 + 4 FFMA per entry).  Z is
//      written straight into the tensor-core operand layout (K-major, SWIZZLE_128B) as a
//      tf32 hi / lo pair, next to the running row sums of A.
//   3. Tensor cores: Y = Z . [W_1; ...; W_C] with tcgen05.mma kind::tf32, M = 128, accumulator in
//      TMEM.  3xTF32 split (Zhi.Whi + Zlo.Whi + Zhi.Wlo) keeps fp32-level accuracy (~1e-6) while
//      the contraction stays far below the HBM time of the tile.
//   4. Epilogue: tcgen05.ld TMEM -> registers, + rowsum (x) bias, activation, staged through padded
//      shared memory and written with coalesced 16-byte stores.
// HBM traffic per layer = x once + y once + CSR once + W once per CTA: the algorithmic minimum.
#include <algorithm>

#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

constexpr int kThreads = 256;
constexpr int kBM = 128;  // rows per tile = UMMA M

struct FusedParams {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int64_t n_graphs;
    int channels, n_nodes, f_in, f_out, act;
    int graphs_per_tile, n_tiles;
    int Kp, Np;       // K = channels * f_in padded to 32, N = f_out padded to 16
    int cv_cap;       // staged {offset, value} capacity (entries)
    int lpr_log2;     // lanes per row in the aggregation
    uint32_t off_zhi, off_zlo, off_whi, off_wlo, off_x, off_y, off_rp, off_cv, off_deg, off_bias, smem_total;
    uint32_t y_pitch;     // bytes per staged output row
    uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ int2 lds_i2(uint32_t addr) {
    int2 r;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}
template <int VEC>
__device__ __forceinline__ void lds_f(float (&r)[VEC], uint32_t addr);
template <>
__device__ __forceinline__ void lds_f<4>(float (&r)[4], uint32_t addr) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_f<2>(float (&r)[2], uint32_t addr) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r[0]), "=f"(r[1]) : "r"(addr));
}
template <>
__device__ __forceinline__ void lds_f<1>(float (&r)[1], uint32_t addr) {
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[0]) : "r"(addr));
}
template <int VEC>
__device__ __forceinline__ void sts_f(uint32_t addr, const float (&r)[VEC]);
template <>
__device__ __forceinline__ void sts_f<4>(uint32_t addr, const float (&r)[4]) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
}
template <>
__device__ __forceinline__ void sts_f<2>(uint32_t addr, const float (&r)[2]) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(r[0]), "f"(r[1]) : "memory");
}
template <>
__device__ __forceinline__ void sts_f<1>(uint32_t addr, const float (&r)[1]) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(r[0]) : "memory");
}

// epilogue for one 16-column accumulator chunk of this thread's row
template <int ACT>
__device__ __forceinline__ void epilogue_chunk(float (&v)[16], int col0, int f_out, int channels, uint32_t deg_addr,
                                               uint32_t bias_addr, int row, uint32_t yrow_addr) {
    for (int c = 0; c < channels; ++c) {
        const float d = __uint_as_float(lds_u32(deg_addr + 4u * (c * kBM + row)));
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int colj = col0 + j;
            const float b = colj < f_out ? __uint_as_float(lds_u32(bias_addr + 4u * (c * f_out + colj))) : 0.0f;
            v[j] = fmaf(d, b, v[j]);
        }
    }
    if (ACT != KGCN_ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = apply_act(v[j], ACT);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (col0 + 4 * q < f_out) {  // the staged row pitch is padded to 16 B, so a partial last float4 is fine
            const float t[4] = {v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]};
            sts_f<4>(yrow_addr + 4u * (col0 + 4 * q), t);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads, 1) graphconv_fused_fwd_kernel(const FusedParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_x, bar_mma;
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));  // generic pointer to the same place
    const uint32_t zhi = base + p.off_zhi, zlo = base + p.off_zlo, whi = base + p.off_whi, wlo = base + p.off_wlo;
    const uint32_t xs = base + p.off_x, ys = base + p.off_y, rp_addr = base + p.off_rp, cv_addr = base + p.off_cv;
    const uint32_t deg_addr = base + p.off_deg, bias_addr = base + p.off_bias;
    int32_t* rp_s = reinterpret_cast<int32_t*>(gen + p.off_rp);
    int2* cv_s = reinterpret_cast<int2*>(gen + p.off_cv);

    const int C = p.channels, N = p.n_nodes, f_in = p.f_in, f_out = p.f_out;
    const int K = C * f_in, Kp = p.Kp, Np = p.Np;
    const uint32_t z_atom = kBM * 128u, w_atom = static_cast<uint32_t>(Np) * 128u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---------------- one-time setup ----------------
    if (tid == 0) {
        mbar_init(&bar_x, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
    // zero the operand regions once: K / N padding must contribute exact zeros (never NaN garbage)
    {
        const uint32_t n16 = (p.off_x - p.off_zhi) >> 4;  // Zhi, Zlo, Whi, Wlo are contiguous
        const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t i = tid; i < n16; i += kThreads) sts_f<4>(zhi + (i << 4), z4);
    }
    __syncthreads();
    // W -> (Whi, Wlo) in the K-major SWIZZLE_128B B-operand layout: B row n = output column n,
    // k index = c * f_in + k.  One thread per (n, 4 consecutive k): coalesced over n in global,
    // conflict-free 16-byte stores in shared (that is what the swizzle is for).
    {
        const int kq = (K + 3) >> 2;
        for (int idx = tid; idx < kq * f_out; idx += kThreads) {
            const int n = idx % f_out, k4 = (idx / f_out) << 2;
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = k4 + j;
                const float wv = kk < K ? __ldg(p.w + static_cast<size_t>(kk) * f_out + n) : 0.0f;  // w is [C][f_in][f_out]
                hi[j] = tf32_hi(wv);
                lo[j] = wv - hi[j];
            }
            const uint32_t off = sw128_offset(n, k4, w_atom);
            sts_f<4>(whi + off, hi);
            sts_f<4>(wlo + off, lo);
        }
        for (int idx = tid; idx < C * f_out; idx += kThreads) {
            const float b = p.bias ? __ldg(p.bias + idx) : 0.0f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_addr + 4u * idx), "f"(b) : "memory");
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = umma_idesc_tf32_m128(Np);

    const int lpr = 1 << p.lpr_log2;
    const int sub = tid & (lpr - 1);
    const int group = tid >> p.lpr_log2;
    const int n_groups = kThreads >> p.lpr_log2;
    const int chunk_f = lpr * VEC;
    const uint32_t row_pitch = static_cast<uint32_t>(f_in) * 4u;
    const uint32_t tile_bytes_graph = static_cast<uint32_t>(N) * row_pitch;

    uint32_t parity = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, parity ^= 1u) {
        const int64_t g0 = static_cast<int64_t>(tile) * p.graphs_per_tile;
        const int ng = static_cast<int>(min(static_cast<int64_t>(p.graphs_per_tile), p.n_graphs - g0));
        const int rows = ng * N;
        const uint32_t x_bytes = static_cast<uint32_t>(ng) * tile_bytes_graph;
        const float* x_tile = p.x + g0 * N * f_in;
        const bool bulk_ok = (x_bytes & 15u) == 0;

        // ---- 1. features: one TMA bulk copy (or a cooperative copy for an unaligned tail tile) ----
        if (bulk_ok) {
            if (tid == 0) {
                mbar_expect_tx(&bar_x, x_bytes);
                bulk_g2s(gen + p.off_x, x_tile, x_bytes, &bar_x);
            }
        } else {
            for (uint32_t i = tid; i < (x_bytes >> 2); i += kThreads)
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(xs + 4u * i), "f"(__ldg(x_tile + i)) : "memory");
        }
        // ---- CSR slice of the tile ----
        const int rows_csr = ng * C * N;
        const int32_t* rp_g = p.rowptr + g0 * C * N;
        for (int r = tid; r <= rows_csr; r += kThreads) rp_s[r] = __ldg(rp_g + r);
        __syncthreads();
        const int32_t e0 = rp_s[0];
        const int n_entries = rp_s[rows_csr] - e0;
        const bool staged = n_entries <= p.cv_cap;
        if (staged) {
            const int mat_rows = C * N;  // CSR rows per graph: all channels gather from the same feature tile
            for (int k = tid; k < n_entries; k += kThreads) {
                int m = 0;
                while (m + 1 < ng && e0 + k >= rp_s[(m + 1) * mat_rows]) ++m;
                cv_s[k] = make_int2(__ldg(p.col + e0 + k) * static_cast<int>(row_pitch) + m * static_cast<int>(tile_bytes_graph),
                                    __float_as_int(__ldg(p.val + e0 + k)));
            }
        }
        __syncthreads();
        if (bulk_ok) mbar_wait(&bar_x, parity);

        // ---- 2. aggregation Z[w, c*f_in + f] = sum_e val_e * x[col_e, f], written as tf32 hi / lo ----
        {
            int i = group, gl = 0;  // carry counters: row w = gl * N + i
            while (i >= N) { i -= N; ++gl; }
            for (int w = group; w < rows; w += n_groups) {
                for (int c = 0; c < C; ++c) {
                    const int r = (gl * C + c) * N + i;
                    const int s = static_cast<int>(lds_u32(rp_addr + 4u * r)) - e0;
                    const int e = static_cast<int>(lds_u32(rp_addr + 4u * r + 4u)) - e0;
                    float deg = 0.0f;
                    for (int f0 = sub * VEC; f0 < f_in; f0 += chunk_f) {
                        float acc[VEC];
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = 0.0f;
                        deg = 0.0f;
                        const uint32_t xb = xs + 4u * f0;
                        if (staged) {
                            uint32_t a = cv_addr + 8u * s;
                            const uint32_t a_end = cv_addr + 8u * e;
#pragma unroll 2
                            for (; a < a_end; a += 8) {
                                const int2 cv = lds_i2(a);
                                float xv[VEC];
                                lds_f<VEC>(xv, xb + static_cast<uint32_t>(cv.x));
                                const float v = __int_as_float(cv.y);
                                deg += v;
#pragma unroll
                                for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                            }
                        } else {  // unusually dense tile: entries straight from global memory
                            for (int k = s; k < e; ++k) {
                                const uint32_t off = static_cast<uint32_t>(__ldg(p.col + e0 + k)) * row_pitch + gl * tile_bytes_graph;
                                const float v = __ldg(p.val + e0 + k);
                                float xv[VEC];
                                lds_f<VEC>(xv, xb + off);
                                deg += v;
#pragma unroll
                                for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                            }
                        }
                        float hi[VEC], lo[VEC];
#pragma unroll
                        for (int t = 0; t < VEC; ++t) {
                            hi[t] = tf32_hi(acc[t]);
                            lo[t] = acc[t] - hi[t];
                        }
                        const int kk = c * f_in + f0;
                        if (VEC == 4 && (kk & 3) == 0) {
                            const uint32_t off = sw128_offset(w, kk, z_atom);
                            sts_f<VEC>(zhi + off, hi);
                            sts_f<VEC>(zlo + off, lo);
                        } else {
#pragma unroll
                            for (int t = 0; t < VEC; ++t) {
                                const uint32_t off = sw128_offset(w, kk + t, z_atom);
                                const float h1[1] = {hi[t]}, l1[1] = {lo[t]};
                                sts_f<1>(zhi + off, h1);
                                sts_f<1>(zlo + off, l1);
                            }
                        }
                    }
                    if (sub == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(deg_addr + 4u * (c * kBM + w)), "f"(deg) : "memory");
                }
                i += n_groups;
                while (i >= N) { i -= N; ++gl; }
            }
        }
        // generic-proxy writes of Z must be visible to the tensor core (async proxy)
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();

        // ---- 3. Y = Z . W on the tensor cores (3xTF32), accumulator in TMEM ----
        if (tid == 0) {
            tc_fence_after_sync();
            bool acc_flag = false;
            const int n_atoms = Kp >> 5;
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a_base = (pass == 1) ? zlo : zhi;
                const uint32_t b_base = (pass == 2) ? wlo : whi;
                for (int at = 0; at < n_atoms; ++at) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (at * 32 + ks * 8 < K) {  // skip k-steps that only see padding
                            umma_tf32(tmem_d, umma_desc_sw128(a_base + at * z_atom + ks * 32u),
                                      umma_desc_sw128(b_base + at * w_atom + ks * 32u), idesc, acc_flag);
                            acc_flag = true;
                        }
                    }
                }
            }
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, parity);
        tc_fence_after_sync();

        // ---- 4. epilogue: TMEM -> registers -> + rowsum (x) bias -> act -> padded smem ----
        {
            const int q = warp & 3, h = warp >> 2;  // TMEM lane quarter / which 16-column chunks
            const int row = q * 32 + lane;
            const uint32_t yrow = ys + static_cast<uint32_t>(row) * p.y_pitch;
            for (int j = h; j * 16 < f_out; j += 2) {
                float v[16];
                tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 16), v);
                if (row < rows) {
                    switch (p.act) {
                        case KGCN_ACT_RELU: epilogue_chunk<KGCN_ACT_RELU>(v, j * 16, f_out, C, deg_addr, bias_addr, row, yrow); break;
                        case KGCN_ACT_SIGMOID: epilogue_chunk<KGCN_ACT_SIGMOID>(v, j * 16, f_out, C, deg_addr, bias_addr, row, yrow); break;
                        case KGCN_ACT_TANH: epilogue_chunk<KGCN_ACT_TANH>(v, j * 16, f_out, C, deg_addr, bias_addr, row, yrow); break;
                        default: epilogue_chunk<KGCN_ACT_NONE>(v, j * 16, f_out, C, deg_addr, bias_addr, row, yrow);
                    }
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();

        // ---- coalesced copy-out of the tile's [rows, f_out] block ----
        {
            float* y_tile = p.y + g0 * N * f_out;
            if ((f_out & 3) == 0 && (reinterpret_cast<uintptr_t>(y_tile) & 15u) == 0) {
                const int q4 = f_out >> 2;
                for (int idx = tid; idx < rows * q4; idx += kThreads) {
                    const int r = idx / q4, cq = idx - r * q4;
                    float t[4];
                    lds_f<4>(t, ys + static_cast<uint32_t>(r) * p.y_pitch + 16u * cq);
                    *reinterpret_cast<float4*>(y_tile + static_cast<size_t>(r) * f_out + 4 * cq) = make_float4(t[0], t[1], t[2], t[3]);
                }
            } else {
                for (int idx = tid; idx < rows * f_out; idx += kThreads) {
                    const int r = idx / f_out, cc = idx - r * f_out;
                    y_tile[idx] = __uint_as_float(lds_u32(ys + static_cast<uint32_t>(r) * p.y_pitch + 4u * cc));
                }
            }
        }
        // the next iteration's first __syncthreads orders this copy-out before Ystage / Z are rewritten
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, p.tmem_cols);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

bool plan(FusedParams& p, int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    if (n_nodes > kBM || f_out > 256 || f_out < 1) return false;
    const int K = channels * f_in;
    p.Kp = static_cast<int>(up(K, 32));
    p.Np = static_cast<int>(up(f_out, 16));
    p.graphs_per_tile = std::max(1, kBM / n_nodes);
    // keep every SM busy when the batch is small: shrink tiles until there are >= 148 of them
    while (p.graphs_per_tile > 1 && ceil_div<int64_t>(n_graphs, p.graphs_per_tile) < kNumSMs) --p.graphs_per_tile;
    p.n_tiles = static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_tile));
    const uint32_t rows_max = static_cast<uint32_t>(p.graphs_per_tile) * n_nodes;
    const uint32_t n_atoms = p.Kp / 32;
    uint32_t off = 0;
    p.off_zhi = off; off += n_atoms * kBM * 128u;
    p.off_zlo = off; off += n_atoms * kBM * 128u;
    p.off_whi = off; off += n_atoms * p.Np * 128u;
    p.off_wlo = off; off += n_atoms * p.Np * 128u;
    p.off_x = off; off += up(rows_max * f_in * 4u, 128);
    p.y_pitch = up(f_out * 4u, 16) + 16u;
    p.off_y = off; off += up(rows_max * p.y_pitch, 128);
    p.off_rp = off; off += up((rows_max * channels + 2) * 4u, 16);
    p.cv_cap = static_cast<int>(std::max<uint32_t>(256, 6 * rows_max * channels));
    p.off_cv = off; off += static_cast<uint32_t>(p.cv_cap) * 8u;
    p.off_deg = off; off += static_cast<uint32_t>(channels) * kBM * 4u;
    p.off_bias = off; off += up(static_cast<uint32_t>(channels) * f_out * 4u, 16);
    p.smem_total = off + 1024;  // slack for the manual 1024-B alignment
    uint32_t cols = 32;
    while (cols < static_cast<uint32_t>(p.Np)) cols <<= 1;
    p.tmem_cols = cols;
    return p.smem_total <= 227 * 1024 - 256;  // static __shared__ (barriers, TMEM slot) shares the 227 KB
}

}  // namespace

bool fused_fwd_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x,
                        const float* y) {
    FusedParams p{};
    if (n_graphs <= 0 || !plan(p, n_graphs, channels, n_nodes, f_in, f_out)) return false;
    // tiles must start 16-byte aligned for the bulk copy / float4 stores
    const uint64_t tile_x = static_cast<uint64_t>(p.graphs_per_tile) * n_nodes * f_in * 4;
    return aligned16(x) && aligned16(y) && tile_x % 16 == 0 && n_graphs * static_cast<int64_t>(n_nodes) < (1ll << 31);
}

int launch_graphconv_fused_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                               int channels, int n_nodes, const float* x, int f_in, const float* w, const float* bias,
                               int f_out, int act, float* y, cudaStream_t st) {
    FusedParams p{};
    KGCN_REQUIRE(plan(p, n_graphs, channels, n_nodes, f_in, f_out), KGCN_ERR_UNSUPPORTED,
                 "fused GraphConv: shape does not fit one SM's shared memory");
    p.rowptr = rowptr; p.col = col; p.val = val; p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.n_graphs = n_graphs; p.channels = channels; p.n_nodes = n_nodes; p.f_in = f_in; p.f_out = f_out; p.act = act;
    const int vec = (f_in % 4 == 0) ? 4 : ((f_in % 2 == 0) ? 2 : 1);
    int lpr_log2 = 0;
    while ((1 << lpr_log2) < 32 && (1 << lpr_log2) * vec < f_in) ++lpr_log2;
    p.lpr_log2 = lpr_log2;
    const unsigned grid = static_cast<unsigned>(std::min<int>(p.n_tiles, kNumSMs));
    auto go = [&](auto kernel) -> int {
        KGCN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem_total)));
        kernel<<<grid, kThreads, p.smem_total, st>>>(p);
        KGCN_LAUNCH_OK("graphconv_fused_fwd_kernel");
        return KGCN_OK;
    };
    switch (vec) {
        case 4: return go(graphconv_fused_fwd_kernel<4>);
        case 2: return go(graphconv_fused_fwd_kernel<2>);
        default: return go(graphconv_fused_fwd_kernel<1>);
    }
}

}  // namespace kgcn
