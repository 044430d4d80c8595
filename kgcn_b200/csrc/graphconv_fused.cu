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
//   4. Epilogue: tcgen05.ld TMEM -> registers, + rowsum (x) bias, activation, 16-byte global stores
//      (a thread owns 64 contiguous bytes of one output row; L2 merges the sectors).
// Tiles are 64 rows (UMMA M = 64) when that lets two CTAs share an SM -- their phases (TMA wait,
// CUDA-core aggregation, tensor-core contraction, epilogue) then overlap -- else 128 rows.
// HBM traffic per layer = x once + y once + CSR once + W once per CTA: the algorithmic minimum.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "fused_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {


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
    int bm;           // rows per tile = UMMA M (64 or 128)
    int n_stages;     // TMA pipeline depth
    uint32_t off_zhi, off_zlo, off_whi, off_wlo, off_stage, off_cv, off_deg, off_bias, smem_total;
    uint32_t stage_bytes, st_rp, st_col, st_val;   // per-stage layout: features at 0, then the CSR slices
    uint32_t y_pitch;   // bytes per row of the staged output tile (multiple of 128)
    uint32_t tmem_cols;
    long long* dbg;   // optional [grid][8] per-CTA phase cycle sums (thread 0), tuning aid
};

// Activations for the epilogue: ex2.approx / rcp.approx based, |error| ~1e-6 (inside the 1e-5 parity
// tolerance) at 4-6 instructions per element instead of ~30 for expf + IEEE division.
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <int ACT>
__device__ __forceinline__ float fast_act(float x) {
    if (ACT == KGCN_ACT_RELU) return fmaxf(x, 0.0f);
    if (ACT == KGCN_ACT_SIGMOID) return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));  // FMUL, EX2, FADD, RCP
    if (ACT == KGCN_ACT_TANH) {
        const float t = ex2_approx(-2.8853900817779268f * fabsf(x));   // exp(-2|x|) in (0, 1]: no overflow
        return copysignf((1.0f - t) * rcp_approx(1.0f + t), x);
    }
    return x;
}

// ---- aggregation of one tile: Z[w, c*f_in + f] = sum_e val_e * x[col_e, f] as tf32 hi / lo, plus row sums ----
struct AggCtx {
    uint32_t zhi, zlo, xs, rp_addr, cv_addr, deg_addr;
    int e0, rows, C, N, f_in;
    uint32_t row_pitch, tile_bytes_graph;
    const int32_t* col;
    const float* val;
};

// hot path: one channel, the whole feature row covered by one lane group (f_in <= LPR * 4), CSR staged
template <int BM>
__device__ __forceinline__ void aggregate_simple(const AggCtx& a, int group, int n_groups, int sub) {
    // addresses live in registers for the whole tile; the empty asm keeps ptxas from rematerialising
    // them (it otherwise rebuilds the shared-window base from SR_CgaCtaId for every row)
    uint32_t xb = a.xs + 16u * sub, rp = a.rp_addr, cvb = a.cv_addr - 8u * static_cast<uint32_t>(a.e0);
    uint32_t zh = a.zhi + (static_cast<uint32_t>(sub) >> 3) * (BM * 128u), zl = a.zlo + (static_cast<uint32_t>(sub) >> 3) * (BM * 128u);
    uint32_t dg = a.deg_addr;
    asm volatile("" : "+r"(xb), "+r"(rp), "+r"(cvb), "+r"(zh), "+r"(zl), "+r"(dg));
    const uint32_t kchunk = static_cast<uint32_t>(sub) & 7u;  // 16-byte chunk inside the 128-byte atom row
    const bool active = sub * 4 < a.f_in;
    // Two rows (w and w + n_groups) are walked together and the {offset, value} pair of the NEXT
    // entry is fetched before the FFMAs of the current one: the LDS -> LDS -> FFMA chain of this
    // latency-bound loop then overlaps across entries and rows.  Reading one pair past the end of a
    // row is harmless (it is the next row's first pair or staging slack) and is never used.
    for (int w0 = group; w0 < a.rows; w0 += 2 * n_groups) {
        const int w1 = w0 + n_groups;
        const bool has1 = w1 < a.rows;
        uint32_t p0 = cvb + 8u * lds_u32(rp + 4u * w0);
        const uint32_t e0 = cvb + 8u * lds_u32(rp + 4u * w0 + 4u);
        uint32_t p1 = has1 ? cvb + 8u * lds_u32(rp + 4u * w1) : 0u;
        const uint32_t e1 = has1 ? cvb + 8u * lds_u32(rp + 4u * w1 + 4u) : 0u;
        float acc0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float deg0 = 0.0f, deg1 = 0.0f;
        int2 c0 = lds_i2(p0), c1 = lds_i2(has1 ? p1 : p0);
#pragma unroll 1
        while (p0 < e0 && p1 < e1) {
            float x0[4], x1[4];
            lds_f<4>(x0, xb + static_cast<uint32_t>(c0.x));
            lds_f<4>(x1, xb + static_cast<uint32_t>(c1.x));
            const float v0 = __int_as_float(c0.y), v1 = __int_as_float(c1.y);
            p0 += 8;
            p1 += 8;
            c0 = lds_i2(p0);   // prefetch the next pairs before the FFMAs wait on x0 / x1
            c1 = lds_i2(p1);
            deg0 += v0;
            deg1 += v1;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                acc0[t] = fmaf(v0, x0[t], acc0[t]);
                acc1[t] = fmaf(v1, x1[t], acc1[t]);
            }
        }
#pragma unroll 1
        while (p0 < e0) {
            float x0[4];
            lds_f<4>(x0, xb + static_cast<uint32_t>(c0.x));
            const float v0 = __int_as_float(c0.y);
            p0 += 8;
            c0 = lds_i2(p0);
            deg0 += v0;
#pragma unroll
            for (int t = 0; t < 4; ++t) acc0[t] = fmaf(v0, x0[t], acc0[t]);
        }
#pragma unroll 1
        while (p1 < e1) {
            float x1[4];
            lds_f<4>(x1, xb + static_cast<uint32_t>(c1.x));
            const float v1 = __int_as_float(c1.y);
            p1 += 8;
            c1 = lds_i2(p1);
            deg1 += v1;
#pragma unroll
            for (int t = 0; t < 4; ++t) acc1[t] = fmaf(v1, x1[t], acc1[t]);
        }
        if (active) {
            float hi[4], lo[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                hi[t] = tf32_hi(acc0[t]);
                lo[t] = acc0[t] - hi[t];
            }
            const uint32_t u0 = static_cast<uint32_t>(w0);
            const uint32_t off0 = (u0 << 7) + ((kchunk ^ (u0 & 7u)) << 4);   // (w>>3)*1024 + (w&7)*128 == w*128
            sts_f<4>(zh + off0, hi);
            sts_f<4>(zl + off0, lo);
            if (has1) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    hi[t] = tf32_hi(acc1[t]);
                    lo[t] = acc1[t] - hi[t];
                }
                const uint32_t u1 = static_cast<uint32_t>(w1);
                const uint32_t off1 = (u1 << 7) + ((kchunk ^ (u1 & 7u)) << 4);
                sts_f<4>(zh + off1, hi);
                sts_f<4>(zl + off1, lo);
            }
        }
        if (sub == 0) {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(dg + 4u * w0), "f"(deg0) : "memory");
            if (has1) asm volatile("st.shared.f32 [%0], %1;" ::"r"(dg + 4u * w1), "f"(deg1) : "memory");
        }
    }
}

// general path: any channel count / feature width / vector width, optional un-staged CSR
template <int VEC, int BM>
__device__ __noinline__ void aggregate_general(const AggCtx& a, int group, int n_groups, int sub, int chunk_f, bool staged) {
    constexpr uint32_t z_atom = BM * 128u;
    int i = group, gl = 0;  // carry counters: row w = gl * N + i
    while (i >= a.N) { i -= a.N; ++gl; }
    for (int w = group; w < a.rows; w += n_groups) {
        for (int c = 0; c < a.C; ++c) {
            const int r = (gl * a.C + c) * a.N + i;
            const int s = static_cast<int>(lds_u32(a.rp_addr + 4u * r)) - a.e0;
            const int e = static_cast<int>(lds_u32(a.rp_addr + 4u * r + 4u)) - a.e0;
            float deg = 0.0f;
            for (int f0 = sub * VEC; f0 < a.f_in; f0 += chunk_f) {
                float acc[VEC];
#pragma unroll
                for (int t = 0; t < VEC; ++t) acc[t] = 0.0f;
                deg = 0.0f;
                const uint32_t xb = a.xs + 4u * f0;
                if (staged) {
                    for (uint32_t p = a.cv_addr + 8u * s; p < a.cv_addr + 8u * e; p += 8) {
                        const int2 cv = lds_i2(p);
                        float xv[VEC];
                        lds_f<VEC>(xv, xb + static_cast<uint32_t>(cv.x));
                        const float v = __int_as_float(cv.y);
                        deg += v;
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                    }
                } else {  // unusually dense tile: entries straight from global memory
                    for (int k = s; k < e; ++k) {
                        const uint32_t off = static_cast<uint32_t>(__ldg(a.col + a.e0 + k)) * a.row_pitch + gl * a.tile_bytes_graph;
                        const float v = __ldg(a.val + a.e0 + k);
                        float xv[VEC];
                        lds_f<VEC>(xv, xb + off);
                        deg += v;
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
                    }
                }
                const int kk = c * a.f_in + f0;
#pragma unroll
                for (int t = 0; t < VEC; ++t) {
                    const float hi = tf32_hi(acc[t]);
                    const uint32_t off = sw128_offset(w, kk + t, z_atom);
                    const float h1[1] = {hi}, l1[1] = {acc[t] - hi};
                    sts_f<1>(a.zhi + off, h1);
                    sts_f<1>(a.zlo + off, l1);
                }
            }
            if (sub == 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a.deg_addr + 4u * (c * BM + w)), "f"(deg) : "memory");
        }
        i += n_groups;
        while (i >= a.N) { i -= a.N; ++gl; }
    }
}

// ---------------------------------------------------------------------------------------------
// Epilogue.  Accumulators (three per tile, one per 3xTF32 pass) are read from TMEM, summed,
// + rowsum (x) bias, activated, and written into a staging tile in shared memory (the Zhi region,
// free once the MMAs have completed) with a 16-byte-chunk XOR swizzle, so that both the
// row-per-thread writes and the row-major read-out are bank-conflict free; the tile then leaves
// with fully coalesced 16-byte global stores.  (Storing straight from the TMEM register layout
// costs one 16/32-byte sector write per lane: measured 2-3.6k cycles per tile.)
// ---------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float finish(float acc, float degbias) { return fast_act<ACT>(acc + degbias); }

// M = 64: .16x256b loads keep all 32 lanes busy.  Warp w owns tile rows 16*(w&3) .. +15 (TMEM lanes
// 32*(w&3) .. +15) and the 32-column slabs (w>>2), (w>>2)+2, ...
template <int ACT>
__device__ __forceinline__ void epilogue_m64(uint32_t tmem_d, uint32_t np, int warp, int lane, int f_out, int C,
                                             uint32_t deg_addr, uint32_t bias_addr, uint32_t ys, uint32_t ypitch) {
    const int q = warp & 3;
    const uint32_t ra = q * 16 + (lane >> 2), rb = ra + 8;
    for (int slab = warp >> 2; slab * 32 < f_out; slab += 2) {
        float v[16], v1[16], v2[16];
        const uint32_t ta = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(slab * 32);
        tmem_ld_16x256b_x4(ta, v);
        tmem_ld_16x256b_x4(ta + np, v1);
        tmem_ld_16x256b_x4(ta + 2 * np, v2);
        float ba[8], bb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ba[i] = bb[i] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float da = __uint_as_float(lds_u32(deg_addr + 4u * (c * 64 + ra)));
            const float db = __uint_as_float(lds_u32(deg_addr + 4u * (c * 64 + rb)));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float b2[2];
                lds_f<2>(b2, bias_addr + 4u * (c * 256 + slab * 32 + 8 * i + 2 * (lane & 3)));
                ba[2 * i] = fmaf(da, b2[0], ba[2 * i]);
                ba[2 * i + 1] = fmaf(da, b2[1], ba[2 * i + 1]);
                bb[2 * i] = fmaf(db, b2[0], bb[2 * i]);
                bb[2 * i + 1] = fmaf(db, b2[1], bb[2 * i + 1]);
            }
        }
        tmem_ld_wait();
        tmem_ld_fence(v);
        tmem_ld_fence(v1);
        tmem_ld_fence(v2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t col = slab * 32 + 8 * i + 2 * (lane & 3);
            if (col < static_cast<uint32_t>(f_out)) {
                const float oa[2] = {finish<ACT>(v[4 * i] + (v1[4 * i] + v2[4 * i]), ba[2 * i]),
                                     finish<ACT>(v[4 * i + 1] + (v1[4 * i + 1] + v2[4 * i + 1]), ba[2 * i + 1])};
                const float ob[2] = {finish<ACT>(v[4 * i + 2] + (v1[4 * i + 2] + v2[4 * i + 2]), bb[2 * i]),
                                     finish<ACT>(v[4 * i + 3] + (v1[4 * i + 3] + v2[4 * i + 3]), bb[2 * i + 1])};
                sts_f<2>(ys + ystage_off(ra, col, ypitch), oa);
                sts_f<2>(ys + ystage_off(rb, col, ypitch), ob);
            }
        }
    }
}

// M = 128: accumulator row m lives in TMEM lane m; a thread owns one row, 16 columns at a time
template <int ACT>
__device__ __forceinline__ void epilogue_m128(uint32_t tmem_d, uint32_t np, int warp, int lane, int f_out, int C,
                                              uint32_t deg_addr, uint32_t bias_addr, uint32_t ys, uint32_t ypitch) {
    const int q = warp & 3;
    const uint32_t row = q * 32 + lane;
    for (int j = warp >> 2; j * 16 < f_out; j += 4) {   // 16 warps: 4 lane quarters x 4 column phases
        float v[16], v1[16], v2[16];
        const uint32_t ta = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 16);
        tmem_ld16(ta, v);
        tmem_ld16(ta + np, v1);
        tmem_ld16(ta + 2 * np, v2);
        float bsum[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bsum[i] = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float d = __uint_as_float(lds_u32(deg_addr + 4u * (c * 128 + row)));
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float b[4];
                lds_f<4>(b, bias_addr + 4u * (c * 256 + j * 16 + 4 * qd));   // bias rows are padded to 256 floats
#pragma unroll
                for (int i = 0; i < 4; ++i) bsum[4 * qd + i] = fmaf(d, b[i], bsum[4 * qd + i]);
            }
        }
        tmem_ld_wait();
        tmem_ld_fence(v);
        tmem_ld_fence(v1);
        tmem_ld_fence(v2);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const uint32_t col = j * 16 + 4 * qd;
            if (col < static_cast<uint32_t>(f_out)) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = finish<ACT>(v[4 * qd + i] + (v1[4 * qd + i] + v2[4 * qd + i]), bsum[4 * qd + i]);
                sts_f<4>(ys + ystage_off(row, col, ypitch), o);
            }
        }
    }
}

constexpr int kMaxStages = 3;

template <int VEC, int BM>
__global__ void __launch_bounds__(BM * 4 + 32, BM == 64 ? 2 : 1) graphconv_fused_fwd_kernel(const FusedParams p) {
    constexpr int kConsumers = BM * 4;
    constexpr int kBlock = kConsumers + 32;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_mma;
    __shared__ __align__(16) StageInfo sinfo[kMaxStages];
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));  // generic pointer to the same place
    const uint32_t zhi = base + p.off_zhi, zlo = base + p.off_zlo, whi = base + p.off_whi, wlo = base + p.off_wlo;
    const uint32_t cv_addr = base + p.off_cv, deg_addr = base + p.off_deg, bias_addr = base + p.off_bias;

    const int C = p.channels, N = p.n_nodes, f_in = p.f_in, f_out = p.f_out;
    const int K = C * f_in, Kp = p.Kp, Np = p.Np;
    constexpr uint32_t z_atom = BM * 128u;
    const uint32_t w_atom = static_cast<uint32_t>(Np) * 128u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.n_stages;
    const uint32_t row_pitch = static_cast<uint32_t>(f_in) * 4u;
    const uint32_t tile_bytes_graph = static_cast<uint32_t>(N) * row_pitch;

    auto tile_graphs = [&](int t) {
        return static_cast<int>(min(static_cast<int64_t>(p.graphs_per_tile), p.n_graphs - static_cast<int64_t>(t) * p.graphs_per_tile));
    };

    // ---------------- one-time setup (all 9 warps) ----------------
    if (tid == 0) {
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_mma, 3);   // one commit per issuing warp
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
    // zero the operand regions once: K / N padding must contribute exact zeros (never NaN garbage)
    {
        const uint32_t n16 = (p.off_stage - p.off_zhi) >> 4;  // Zhi, Zlo, Whi, Wlo are contiguous
        const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t i = tid; i < n16; i += kBlock) sts_f<4>(zhi + (i << 4), z4);
        for (uint32_t i = tid; i < static_cast<uint32_t>(C) * 64u; i += kBlock) sts_f<4>(bias_addr + (i << 4), z4);
    }
    pdl_wait();   // barrier init, TMEM allocation and the zero fill above overlap the previous kernel's tail
    __syncthreads();

    if (warp == kConsumers / 32) {
        // =============================== TMA producer warp ===============================
        // Runs up to S tiles ahead: per tile one bulk copy each for the feature rows, the row-extent
        // slice, the column slice and the value slice, all completing on full[s].
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int s = it % S;
                if (it >= S) mbar_wait(&bar_empty[s], ((it / S) - 1) & 1);
                const int64_t g0 = static_cast<int64_t>(tile) * p.graphs_per_tile;
                const int ng = tile_graphs(tile);
                const int64_t r0 = g0 * C * N;
                const int rows_csr = ng * C * N;
                const int32_t e_first = __ldg(p.rowptr + r0), e_last = __ldg(p.rowptr + r0 + rows_csr);
                // bulk copies need 16-byte aligned sources and sizes: copy the enclosing aligned slices
                const int64_t rp_lo = r0 & ~3ll;
                const uint32_t rp_cnt = static_cast<uint32_t>((r0 + rows_csr + 1 - rp_lo + 3) & ~3ll);
                const int32_t e_lo = e_first & ~3;
                const uint32_t e_cnt = static_cast<uint32_t>((e_last - e_lo + 3) & ~3);
                const bool staged = e_cnt <= static_cast<uint32_t>(p.cv_cap);
                unsigned char* st = gen + p.off_stage + static_cast<size_t>(s) * p.stage_bytes;
                StageInfo& si = sinfo[s];
                si.e_first = e_first;
                si.n_entries = e_last - e_first;
                si.rp_skip = static_cast<int32_t>(r0 - rp_lo);
                si.e_skip = e_first - e_lo;
                si.staged = staged ? 1 : 0;
                const uint32_t x_bytes = static_cast<uint32_t>(ng) * tile_bytes_graph;
                mbar_expect_tx(&bar_full[s], x_bytes + 4u * rp_cnt + (staged ? 8u * e_cnt : 0u));
                bulk_g2s(st, p.x + g0 * N * f_in, x_bytes, &bar_full[s]);
                bulk_g2s(st + p.st_rp, p.rowptr + rp_lo, 4u * rp_cnt, &bar_full[s]);
                if (staged && e_cnt) {
                    bulk_g2s(st + p.st_col, p.col + e_lo, 4u * e_cnt, &bar_full[s]);
                    bulk_g2s(st + p.st_val, p.val + e_lo, 4u * e_cnt, &bar_full[s]);
                }
            }
        }
    } else {
        // =============================== consumer warps ===============================
        // W -> (Whi, Wlo) in the K-major SWIZZLE_128B B-operand layout: B row n = output column n,
        // k index = c * f_in + k.  One thread per (n, 4 consecutive k): coalesced over n in global,
        // conflict-free 16-byte stores in shared (that is what the swizzle is for).
        {
            const int kq = (K + 3) >> 2;
            for (int idx = tid; idx < kq * f_out; idx += kConsumers) {
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
            if (p.bias != nullptr)
                for (int idx = tid; idx < C * f_out; idx += kConsumers) {
                    const int c = idx / f_out, n = idx - c * f_out;
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_addr + 4u * (c * 256 + n)), "f"(__ldg(p.bias + idx)) : "memory");
                }
        }
        fence_proxy_async_smem();  // W is read by the tensor core through the async proxy
        tc_fence_before_sync();
        consumer_sync<kConsumers>();
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_slot;
        const uint32_t idesc = umma_idesc_tf32(BM, Np);
        // operand descriptors are tile-invariant: build them once, step them by adding to the address field
        const uint64_t desc_zhi = umma_desc_sw128(zhi), desc_zlo = umma_desc_sw128(zlo);
        const uint64_t desc_whi = umma_desc_sw128(whi), desc_wlo = umma_desc_sw128(wlo);
        const int n_atoms = Kp >> 5;

        const int lpr = 1 << p.lpr_log2;
        const int sub = tid & (lpr - 1);
        const int group = tid >> p.lpr_log2;
        const int n_groups = kConsumers >> p.lpr_log2;
        const bool simple = (VEC == 4) && C == 1 && f_in <= lpr * 4;
        const bool y_vec4 = (f_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15u) == 0;
    
        long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tlast = clock64();
        auto mark = [&](int ph) {
            if (p.dbg != nullptr && tid == 0) {
                const long long now = clock64();
                tph[ph] += now - tlast;
                tlast = now;
            }
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int s = it % S;
            const int64_t g0 = static_cast<int64_t>(tile) * p.graphs_per_tile;
            const int ng = tile_graphs(tile);
            const int rows = ng * N;
            const uint32_t st = base + p.off_stage + static_cast<uint32_t>(s) * p.stage_bytes;

            // ---- 1. wait for the stage: features + CSR slices landed by TMA ----
            mbar_wait(&bar_full[s], (it / S) & 1);
            mark(0);
            const StageInfo si = sinfo[s];
            const uint32_t rp_addr = st + p.st_rp + 4u * static_cast<uint32_t>(si.rp_skip);
            // {column, value} -> {byte offset of the neighbour's feature row in the stage, value}
            if (si.staged) {
                const uint32_t col_a = st + p.st_col + 4u * static_cast<uint32_t>(si.e_skip);
                const uint32_t val_a = st + p.st_val + 4u * static_cast<uint32_t>(si.e_skip);
                const int mat_rows = C * N;
                for (int k = tid; k < si.n_entries; k += kConsumers) {
                    int m = 0;
                    while (m + 1 < ng && si.e_first + k >= static_cast<int>(lds_u32(rp_addr + 4u * (m + 1) * mat_rows))) ++m;
                    const uint32_t off = lds_u32(col_a + 4u * k) * row_pitch + static_cast<uint32_t>(m) * tile_bytes_graph;
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(cv_addr + 8u * k), "r"(off), "r"(lds_u32(val_a + 4u * k)) : "memory");
                }
            }
            consumer_sync<kConsumers>();
            mark(1);

            // ---- 2. aggregation on the CUDA cores ----
            {
                AggCtx a{zhi, zlo, st, rp_addr, cv_addr, deg_addr, si.e_first, rows, C, N, f_in, row_pitch, tile_bytes_graph, p.col, p.val};
                if (simple && si.staged) aggregate_simple<BM>(a, group, n_groups, sub);
                else aggregate_general<VEC, BM>(a, group, n_groups, sub, lpr * VEC, si.staged != 0);
            }
            fence_proxy_async_smem();  // generic-proxy writes of Z -> visible to the tensor core (async proxy)
            tc_fence_before_sync();
            consumer_sync<kConsumers>();
            mark(2);

            // ---- 3. Y = Z . W on the tensor cores (3xTF32), accumulator in TMEM ----
            if (warp < 3) {   // whole warp converged, one elected lane issues pass `warp` into accumulator `warp`
                if (elect_one()) {
                    if (warp == 0) mbar_arrive(&bar_empty[s]);  // every consumer is past the barrier: the stage can be refilled
                    tc_fence_after_sync();
                    const uint64_t da = (warp == 1) ? desc_zlo : desc_zhi;
                    const uint64_t db = (warp == 2) ? desc_wlo : desc_whi;
                    const uint32_t d = tmem_d + static_cast<uint32_t>(warp * Np);
                    switch (n_atoms) {
                        case 1: issue_pass<1>(d, da, db, idesc, z_atom >> 4, w_atom >> 4, K); break;
                        case 2: issue_pass<2>(d, da, db, idesc, z_atom >> 4, w_atom >> 4, K); break;
                        case 3: issue_pass<3>(d, da, db, idesc, z_atom >> 4, w_atom >> 4, K); break;
                        case 4: issue_pass<4>(d, da, db, idesc, z_atom >> 4, w_atom >> 4, K); break;
                        default: issue_pass_loop(d, da, db, idesc, z_atom >> 4, w_atom >> 4, K, n_atoms);
                    }
                    umma_commit(&bar_mma);
                }
                __syncwarp();
            }
            mark(3);
            mbar_wait(&bar_mma, it & 1);
            tc_fence_after_sync();
            mark(4);

            // ---- 4. epilogue: TMEM -> registers -> + rowsum (x) bias -> act -> staged tile -> global ----
            {
                const uint32_t ys = zhi;   // the operand region is free once the MMAs have completed
                if (BM == 64) {
                    switch (p.act) {
                        case KGCN_ACT_RELU: epilogue_m64<KGCN_ACT_RELU>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        case KGCN_ACT_SIGMOID: epilogue_m64<KGCN_ACT_SIGMOID>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        case KGCN_ACT_TANH: epilogue_m64<KGCN_ACT_TANH>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        default: epilogue_m64<KGCN_ACT_NONE>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch);
                    }
                } else {
                    switch (p.act) {
                        case KGCN_ACT_RELU: epilogue_m128<KGCN_ACT_RELU>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        case KGCN_ACT_SIGMOID: epilogue_m128<KGCN_ACT_SIGMOID>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        case KGCN_ACT_TANH: epilogue_m128<KGCN_ACT_TANH>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch); break;
                        default: epilogue_m128<KGCN_ACT_NONE>(tmem_d, Np, warp, lane, f_out, C, deg_addr, bias_addr, ys, p.y_pitch);
                    }
                }
                tc_fence_before_sync();
                consumer_sync<kConsumers>();
                copy_out<kConsumers>(ys, p.y_pitch, p.y + g0 * N * f_out, rows, f_out, tid, y_vec4);
            }
            tc_fence_before_sync();
            consumer_sync<kConsumers>();  // TMEM / row sums / cv pairs consumed before the next tile overwrites them
            mark(5);
        }
        if (p.dbg != nullptr && tid == 0)
            for (int i = 0; i < 8; ++i) p.dbg[static_cast<size_t>(blockIdx.x) * 8 + i] = (i == 7) ? it : tph[i];
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, p.tmem_cols);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

bool plan_bm(FusedParams& p, int bm, int max_smem, int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    if (n_nodes > bm || f_out > 160 || f_out < 1) return false;   // 3 * Np TMEM columns <= 512
    const int K = channels * f_in;
    p.bm = bm;
    p.Kp = static_cast<int>(up(K, 32));
    p.Np = static_cast<int>(up(f_out, 16));
    p.graphs_per_tile = std::max(1, bm / n_nodes);
    // keep every SM busy when the batch is small: shrink tiles until there are >= 2 per SM
    while (p.graphs_per_tile > 1 && ceil_div<int64_t>(n_graphs, p.graphs_per_tile) < 2 * kNumSMs) --p.graphs_per_tile;
    // every tile (the last, shorter one included) is fetched with a bulk copy: 16-byte sizes only
    const uint64_t graph_bytes = static_cast<uint64_t>(n_nodes) * f_in * 4;
    while (p.graphs_per_tile > 1 && (p.graphs_per_tile * graph_bytes) % 16 != 0) --p.graphs_per_tile;
    if ((p.graphs_per_tile * graph_bytes) % 16 != 0) return false;
    if (graph_bytes % 16 != 0 && n_graphs % p.graphs_per_tile != 0) return false;
    p.n_tiles = static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_tile));
    const uint32_t rows_max = static_cast<uint32_t>(p.graphs_per_tile) * n_nodes;
    const uint32_t n_atoms = p.Kp / 32;
    uint32_t off = 0;
    p.off_zhi = off; off += n_atoms * bm * 128u;
    p.off_zlo = off; off += n_atoms * bm * 128u;
    // the output tile is staged in the Zhi|Zlo region after the MMAs: rows of y_pitch bytes, whole 128-byte groups
    p.y_pitch = up(f_out * 4u, 128);
    if (p.y_pitch * static_cast<uint32_t>(bm) > 2u * n_atoms * bm * 128u) return false;
    p.off_whi = off; off += n_atoms * p.Np * 128u;
    p.off_wlo = off; off += n_atoms * p.Np * 128u;
    p.off_stage = off;
    p.cv_cap = static_cast<int>(up(std::max<uint32_t>(256, 6 * rows_max * channels), 4));
    p.st_rp = up(rows_max * f_in * 4u, 128);
    p.st_col = p.st_rp + up((rows_max * channels + 8) * 4u, 16);
    p.st_val = p.st_col + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u;
    p.stage_bytes = up(p.st_val + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u, 128);
    const uint32_t fixed = static_cast<uint32_t>(p.cv_cap) * 8u + static_cast<uint32_t>(channels) * bm * 4u +
                           static_cast<uint32_t>(channels) * 256u * 4u + 1024u;
    p.n_stages = 0;
    for (int st = kMaxStages; st >= 1; --st)
        if (off + st * p.stage_bytes + fixed <= static_cast<uint32_t>(max_smem)) { p.n_stages = st; break; }
    if (p.n_stages == 0) return false;
    off += p.n_stages * p.stage_bytes;
    p.off_cv = off; off += static_cast<uint32_t>(p.cv_cap) * 8u;
    p.off_deg = off; off += static_cast<uint32_t>(channels) * bm * 4u;
    p.off_bias = off; off += static_cast<uint32_t>(channels) * 256u * 4u;
    p.smem_total = off + 1024;  // slack for the manual 1024-B alignment
    uint32_t cols = 32;
    while (cols < 3u * static_cast<uint32_t>(p.Np)) cols <<= 1;   // three accumulators (one per 3xTF32 pass)
    p.tmem_cols = cols;
    return true;
}

constexpr int kSmemOneCta = 227 * 1024 - 512;   // static __shared__ (barriers, stage info) shares the 227 KB
constexpr int kSmemTwoCtas = 113 * 1024 - 512;

// 64-row tiles when two CTAs then fit one SM (their phases overlap), else 128-row tiles.
bool plan(FusedParams& p, int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    static const char* force = getenv("KGCN_FUSED_BM");   // tuning knob: 64 / 128 forces the tile height
    if (force != nullptr) {
        const int bm = atoi(force);
        if (bm == 128) return plan_bm(p, 128, kSmemOneCta, n_graphs, channels, n_nodes, f_in, f_out);
        if (bm == 64) return plan_bm(p, 64, kSmemTwoCtas, n_graphs, channels, n_nodes, f_in, f_out) ||
                             plan_bm(p, 64, kSmemOneCta, n_graphs, channels, n_nodes, f_in, f_out);
    }
    FusedParams q = p;
    if (plan_bm(q, 64, kSmemTwoCtas, n_graphs, channels, n_nodes, f_in, f_out) && q.n_stages >= 2) {
        p = q;
        return true;
    }
    if (plan_bm(p, 128, kSmemOneCta, n_graphs, channels, n_nodes, f_in, f_out)) return true;
    return plan_bm(p, 64, kSmemOneCta, n_graphs, channels, n_nodes, f_in, f_out);
}

}  // namespace

bool fused_fwd_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x,
                        const float* y, const int32_t* rowptr, const int32_t* col, const float* val) {
    FusedParams p{};
    if (n_graphs <= 0 || !plan(p, n_graphs, channels, n_nodes, f_in, f_out)) return false;
    return aligned16(x) && aligned16(rowptr) && aligned16(col) && aligned16(val) && (reinterpret_cast<uintptr_t>(y) & 3u) == 0 && n_graphs * static_cast<int64_t>(n_nodes) < (1ll << 31);
}

static long long* g_dbg = nullptr;   // set through kgcn_debug_fused_times (tuning only)

int launch_graphconv_fused_fwd(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                               int channels, int n_nodes, const float* x, int f_in, const float* w, const float* bias,
                               int f_out, int act, float* y, cudaStream_t st) {
    FusedParams p{};
    KGCN_REQUIRE(plan(p, n_graphs, channels, n_nodes, f_in, f_out), KGCN_ERR_UNSUPPORTED,
                 "fused GraphConv: shape does not fit one SM's shared memory");
    // the CSR slices are fetched with 16-byte bulk copies of the enclosing aligned ranges
    KGCN_REQUIRE(aligned16(rowptr) && aligned16(col) && aligned16(val), KGCN_ERR_MISALIGNED,
                 "fused GraphConv: rowptr/col/val must be 16-byte aligned");
    p.rowptr = rowptr; p.col = col; p.val = val; p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.n_graphs = n_graphs; p.channels = channels; p.n_nodes = n_nodes; p.f_in = f_in; p.f_out = f_out; p.act = act;
    p.dbg = g_dbg;
    const int vec = (f_in % 4 == 0) ? 4 : ((f_in % 2 == 0) ? 2 : 1);
    int lpr_log2 = 0;
    while ((1 << lpr_log2) < 32 && (1 << lpr_log2) * vec < f_in) ++lpr_log2;
    p.lpr_log2 = lpr_log2;
    const int ctas_per_sm = (p.bm == 64 && p.smem_total <= static_cast<uint32_t>(kSmemTwoCtas) + 1024) ? 2 : 1;
    const unsigned grid = static_cast<unsigned>(std::min<int>(p.n_tiles, kNumSMs * ctas_per_sm));
    auto go = [&](auto kernel) -> int {
        KGCN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem_total)));
        launch_pdl(kernel, grid, p.bm * 4 + 32, p.smem_total, st, p);
        KGCN_LAUNCH_OK("graphconv_fused_fwd_kernel");
        return KGCN_OK;
    };
    if (p.bm == 64) {
        switch (vec) {
            case 4: return go(graphconv_fused_fwd_kernel<4, 64>);
            case 2: return go(graphconv_fused_fwd_kernel<2, 64>);
            default: return go(graphconv_fused_fwd_kernel<1, 64>);
        }
    }
    switch (vec) {
        case 4: return go(graphconv_fused_fwd_kernel<4, 128>);
        case 2: return go(graphconv_fused_fwd_kernel<2, 128>);
        default: return go(graphconv_fused_fwd_kernel<1, 128>);
    }
}

}  // namespace kgcn

// Tuning hook (not part of the documented ABI): device buffer of [grid][8] int64 that the fused kernel
// fills with per-CTA phase cycle sums {wait stage, convert, aggregate, mma issue, mma wait, epilogue, -, tiles}.
extern "C" void kgcn_debug_fused_times(long long* device_buffer) { kgcn::g_dbg = device_buffer; }
