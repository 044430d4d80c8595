// Fused GraphConv backward for sm_100a (single adjacency channel): ONE kernel per layer computes
//
//     dU = dy (.) act'(y)                       (kgcn/layers.py:115 + the model's activation)
//     G  = A^T . dU                             (adjoint_a=True of bspmm_call.py:44; transposed BatchedCSR)
//     dx = G . W^T                              (only when the layer input needs a gradient)
//     dW = X^T . G ,  dbias = column sums of G  (per-CTA partial sums, reduced by splitk_reduce_kernel)
//
// i.e. the backward of  y = act((A.x).W + rowsum(A) (x) b)  == act(A.(x.W + b))  (layers.py:105-116).
// It replaces the act_grad -> SpMM -> split-K FFMA GEMM -> FFMA GEMM chain of graphconv.cu (three
// extra round trips of a [B*N, F] tensor through HBM) by: x, y, dy read once, dx written once.
//
// Per persistent CTA (one per SM), per tile of 64 rows (whole graphs):
//   1. TMA: bulk-async copies land the tile's x rows, y rows, dy rows (or the per-graph rows of a
//      broadcast dy = GraphGather's gradient) and the transposed-CSR slices; up to 3 stages deep.
//   2. CUDA cores: dU in place; G = A^T.dU as a segmented sum out of shared memory, written as tf32
//      hi / lo pairs into TWO tensor-core operand layouts: K-major SWIZZLE_128B (A operand of dx) and
//      MN-major SWIZZLE_128B_BASE32B (B operand of dW -- the contraction index of dW is the row index,
//      so both of its operands are "transposed" tiles; tcgen05 reads them MN-major, no transposition
//      pass).  x is split the same way.  Column sums of G (dbias) accumulate in registers.
//   3. Tensor cores, 3xTF32 with stacked operands (fewer, larger instructions):
//        dx : Ghi.[Whi;Wlo]^T (N = 2*F_in) and Glo.Whi^T   -> two TMEM accumulators, 8 + 8 MMAs
//        dW : [Xhi;Xlo]^T.[Ghi|Glo] (M = 128, N = 2*F_out) -> one TMEM accumulator that lives across
//             ALL tiles of the CTA (8 MMAs per tile, accumulate flag on after the first tile)
//   4. dx epilogue: TMEM -> registers -> swizzled staging tile -> coalesced 16-byte stores.
// At the end each CTA adds the four hi/lo blocks of its dW accumulator and writes one partial
// [F_in + 1, F_out] block (last row = dbias partial); a fixed-order reduction over CTAs follows.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "fused_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

constexpr int kRows = 64;                 // rows per tile: UMMA M of dx, K of dW
constexpr int kConsumers = 512;           // 16 warps, 8 threads per tile row
constexpr int kBlock = kConsumers + 32;   // + the TMA producer warp
constexpr int kStagesMax = 3;
constexpr uint32_t kMnLbo = kRows * 128u; // bytes between 32-column chunks of an MN-major operand
constexpr uint32_t kMnSbo = 512u;         // 4-row swizzle atoms along K
constexpr uint32_t kZAtom = kRows * 128u; // K-major atom (32 k values) of the 64-row G operand
constexpr int kMi = 64;                   // F_in padded to the UMMA M granule of dW (hi and lo stacked -> M = 128)

struct BwdParams {
    const int32_t* rowptr;   // transposed BatchedCSR
    const int32_t* col;
    const float* val;
    const float* x;
    const float* w;
    const float* y;
    const float* dy;
    float* dx;
    float* partial;          // [grid][(f_in + 1) * f_out]
    int64_t n_graphs;
    int n_nodes, f_in, f_out, act, dy_bcast;
    int graphs_per_tile, n_tiles;
    int graphs_per_cta;      // every CTA owns a contiguous range of graphs (balanced: the last tile of a range may be short)
    int KGp;                 // f_out padded to 32: K of dx, N of dW
    int Fip;                 // f_in padded to 16: N of dx
    int cv_cap, lpr_log2, n_stages;
    uint32_t off_z, off_w, off_z2, off_x, off_stage, off_cv, smem_total;
    uint32_t stage_bytes, st_u, st_d, st_rp, st_col, st_val;
    uint32_t u_src;          // what TMA puts into the U region: 0 nothing, 1 y, 2 dy
    uint32_t d_src;          // what TMA puts into the D region: 0 nothing, 1 dy rows, 2 dy per-graph rows
    uint32_t y_pitch, tmem_cols;
};

struct BwdCtx {
    uint32_t zhi, zlo, z2hi, z2lo, us, rp_addr, cv_addr;
    int e0, rows, N, f_out;
    uint32_t row_pitch;
    const int32_t* col;
    const float* val;
};

// G row w (4 features of lane `sub`) -> the two operand layouts, hi / lo
template <bool HAS_DX>
__device__ __forceinline__ void store_g(const BwdCtx& a, uint32_t w, int sub, const float (&acc)[4]) {
    float hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        hi[t] = tf32_hi(acc[t]);
        lo[t] = acc[t] - hi[t];
    }
    const uint32_t s = static_cast<uint32_t>(sub);
    if (HAS_DX) {
        const uint32_t off = (s >> 3) * kZAtom + (w << 7) + (((s & 7u) ^ (w & 7u)) << 4);
        sts_f<4>(a.zhi + off, hi);
        sts_f<4>(a.zlo + off, lo);
    }
    const uint32_t off2 = mn32_offset(w, s << 2, kMnLbo);
    sts_f<4>(a.z2hi + off2, hi);
    sts_f<4>(a.z2lo + off2, lo);
}

// G = A^T . dU for one tile: a lane group per row, two rows in flight (see aggregate_simple in
// graphconv_fused.cu); csum accumulates this thread's share of the column sums of G.
template <bool HAS_DX>
__device__ __forceinline__ void aggregate_bwd(const BwdCtx& a, int group, int n_groups, int sub, float (&csum)[4]) {
    uint32_t xb = a.us + 16u * sub, rp = a.rp_addr, cvb = a.cv_addr - 8u * static_cast<uint32_t>(a.e0);
    asm volatile("" : "+r"(xb), "+r"(rp), "+r"(cvb));
    const bool active = sub * 4 < a.f_out;
    for (int w0 = group; w0 < a.rows; w0 += 2 * n_groups) {
        const int w1 = w0 + n_groups;
        const bool has1 = w1 < a.rows;
        uint32_t p0 = cvb + 8u * lds_u32(rp + 4u * w0);
        const uint32_t e0 = cvb + 8u * lds_u32(rp + 4u * w0 + 4u);
        uint32_t p1 = has1 ? cvb + 8u * lds_u32(rp + 4u * w1) : 0u;
        const uint32_t e1 = has1 ? cvb + 8u * lds_u32(rp + 4u * w1 + 4u) : 0u;
        float acc0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int2 c0 = lds_i2(p0), c1 = lds_i2(has1 ? p1 : p0);
#pragma unroll 1
        while (p0 < e0 && p1 < e1) {
            float x0[4], x1[4];
            lds_f<4>(x0, xb + static_cast<uint32_t>(c0.x));
            lds_f<4>(x1, xb + static_cast<uint32_t>(c1.x));
            const float v0 = __int_as_float(c0.y), v1 = __int_as_float(c1.y);
            p0 += 8;
            p1 += 8;
            c0 = lds_i2(p0);
            c1 = lds_i2(p1);
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
#pragma unroll
            for (int t = 0; t < 4; ++t) acc1[t] = fmaf(v1, x1[t], acc1[t]);
        }
        if (active) {
            store_g<HAS_DX>(a, static_cast<uint32_t>(w0), sub, acc0);
#pragma unroll
            for (int t = 0; t < 4; ++t) csum[t] += acc0[t];
            if (has1) {
                store_g<HAS_DX>(a, static_cast<uint32_t>(w1), sub, acc1);
#pragma unroll
                for (int t = 0; t < 4; ++t) csum[t] += acc1[t];
            }
        }
    }
}

// tile with more entries than the stage holds: same mapping, entries straight from global memory
template <bool HAS_DX>
__device__ __noinline__ void aggregate_bwd_unstaged(const BwdCtx& a, int group, int n_groups, int sub, float (&csum)[4]) {
    const bool active = sub * 4 < a.f_out;
    for (int w = group; w < a.rows; w += n_groups) {
        const int s = static_cast<int>(lds_u32(a.rp_addr + 4u * w)), e = static_cast<int>(lds_u32(a.rp_addr + 4u * w + 4u));
        const uint32_t gbase = static_cast<uint32_t>(w / a.N) * static_cast<uint32_t>(a.N) * a.row_pitch;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int k = s; k < e; ++k) {
            const uint32_t off = gbase + static_cast<uint32_t>(__ldg(a.col + k)) * a.row_pitch;
            const float v = __ldg(a.val + k);
            float xv[4];
            lds_f<4>(xv, a.us + 16u * sub + off);
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] = fmaf(v, xv[t], acc[t]);
        }
        if (active) {
            store_g<HAS_DX>(a, static_cast<uint32_t>(w), sub, acc);
#pragma unroll
            for (int t = 0; t < 4; ++t) csum[t] += acc[t];
        }
    }
}

template <bool HAS_DX>
__global__ void __launch_bounds__(kBlock, 1) graphconv_fused_bwd_kernel(const BwdParams p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kStagesMax], bar_empty[kStagesMax], bar_mma;
    __shared__ __align__(16) StageInfo sinfo[kStagesMax];
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));
    const int N = p.n_nodes, f_in = p.f_in, f_out = p.f_out, KGp = p.KGp, Fip = p.Fip;
    const int n_katoms = KGp >> 5;
    const uint32_t zhi = base + p.off_z, zlo = zhi + static_cast<uint32_t>(n_katoms) * kZAtom;
    const uint32_t wsm = base + p.off_w, w_atom = 2u * static_cast<uint32_t>(Fip) * 128u;
    const uint32_t z2hi = base + p.off_z2, z2lo = z2hi + static_cast<uint32_t>(n_katoms) * kMnLbo;
    const uint32_t xhi = base + p.off_x, xlo = xhi + (kMi / 32) * kMnLbo;
    const uint32_t cv_addr = base + p.off_cv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.n_stages;
    const uint32_t row_pitch = static_cast<uint32_t>(f_out) * 4u;            // dU rows
    const uint32_t graph_bytes_u = static_cast<uint32_t>(N) * row_pitch;
    const uint32_t graph_bytes_x = static_cast<uint32_t>(N) * static_cast<uint32_t>(f_in) * 4u;

    const int64_t cta_g0 = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
    const int cta_ng = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(p.graphs_per_cta), p.n_graphs - cta_g0)));
    const int cta_tiles = (cta_ng + p.graphs_per_tile - 1) / p.graphs_per_tile;

    if (tid == 0) {
        for (int i = 0; i < kStagesMax; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 1);
        }
        mbar_init(&bar_mma, HAS_DX ? 3 : 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
    {   // operand regions start as exact zeros: K / M / N padding must never contribute NaN garbage
        const uint32_t n16 = (p.off_stage - p.off_z) >> 4;
        const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t i = tid; i < n16; i += kBlock) sts_f<4>(base + p.off_z + (i << 4), z4);
    }
    pdl_wait();
    __syncthreads();

    if (warp == kConsumers / 32) {
        // =============================== TMA producer warp ===============================
        if (lane == 0) {
            for (int it = 0; it < cta_tiles; ++it) {
                const int s = it % S;
                if (it >= S) mbar_wait(&bar_empty[s], ((it / S) - 1) & 1);
                const int64_t g0 = cta_g0 + static_cast<int64_t>(it) * p.graphs_per_tile;
                const int ng = min(p.graphs_per_tile, cta_ng - it * p.graphs_per_tile);
                const int64_t r0 = g0 * N;
                const int rows_csr = ng * N;
                const int32_t e_first = __ldg(p.rowptr + r0), e_last = __ldg(p.rowptr + r0 + rows_csr);
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
                const uint32_t x_bytes = static_cast<uint32_t>(ng) * graph_bytes_x;
                const uint32_t u_bytes = p.u_src ? static_cast<uint32_t>(ng) * graph_bytes_u : 0u;
                const uint32_t d_bytes = p.d_src == 1 ? static_cast<uint32_t>(ng) * graph_bytes_u
                                                      : (p.d_src == 2 ? static_cast<uint32_t>(ng) * row_pitch : 0u);
                mbar_expect_tx(&bar_full[s], x_bytes + u_bytes + d_bytes + 4u * rp_cnt + (staged ? 8u * e_cnt : 0u));
                bulk_g2s(st, p.x + r0 * f_in, x_bytes, &bar_full[s]);
                if (p.u_src) bulk_g2s(st + p.st_u, (p.u_src == 1 ? p.y : p.dy) + r0 * f_out, u_bytes, &bar_full[s]);
                if (p.d_src == 1) bulk_g2s(st + p.st_d, p.dy + r0 * f_out, d_bytes, &bar_full[s]);
                if (p.d_src == 2) bulk_g2s(st + p.st_d, p.dy + g0 * f_out, d_bytes, &bar_full[s]);
                bulk_g2s(st + p.st_rp, p.rowptr + rp_lo, 4u * rp_cnt, &bar_full[s]);
                if (staged && e_cnt) {
                    bulk_g2s(st + p.st_col, p.col + e_lo, 4u * e_cnt, &bar_full[s]);
                    bulk_g2s(st + p.st_val, p.val + e_lo, 4u * e_cnt, &bar_full[s]);
                }
            }
        }
    } else {
        // =============================== consumer warps ===============================
        if (HAS_DX) {
            // B operand of dx: row n = input feature n, k = output feature: W itself ([f_in][f_out]
            // row-major is already "K-major"); rows 0..Fip-1 hold Whi, rows Fip..2Fip-1 hold Wlo.
            const int kq = f_out >> 2;
            for (int idx = tid; idx < kq * f_in; idx += kConsumers) {
                const int n = idx / kq, k4 = (idx - n * kq) << 2;
                const float4 wv = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(n) * f_out + k4));
                const float v[4] = {wv.x, wv.y, wv.z, wv.w};
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hi[j] = tf32_hi(v[j]);
                    lo[j] = v[j] - hi[j];
                }
                sts_f<4>(wsm + sw128_offset(n, k4, w_atom), hi);
                sts_f<4>(wsm + sw128_offset(Fip + n, k4, w_atom), lo);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        consumer_sync<kConsumers>();
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_slot;
        const uint32_t d1a = tmem_d, d1b = tmem_d + 2u * Fip, d2 = HAS_DX ? tmem_d + 3u * Fip : tmem_d;
        const uint32_t idesc_dxa = umma_idesc_tf32(kRows, 2 * Fip), idesc_dxb = umma_idesc_tf32(kRows, Fip);
        const uint32_t idesc_dw = umma_idesc_tf32(128, 2 * KGp) | kUmmaMajorMnA | kUmmaMajorMnB;
        const uint64_t desc_zhi = umma_desc_sw128(zhi), desc_zlo = umma_desc_sw128(zlo), desc_w = umma_desc_sw128(wsm);
        const uint64_t desc_x = umma_desc_mn32(xhi, kMnLbo, kMnSbo), desc_g = umma_desc_mn32(z2hi, kMnLbo, kMnSbo);

        const int lpr = 1 << p.lpr_log2;
        const int sub = tid & (lpr - 1);
        const int group = tid >> p.lpr_log2;
        const int n_groups = kConsumers >> p.lpr_log2;
        const bool dx_vec4 = (reinterpret_cast<uintptr_t>(p.dx) & 15u) == 0;
        const int f4o = f_out >> 2, f4i = f_in >> 2;
        float csum[4] = {0.0f, 0.0f, 0.0f, 0.0f};

        for (int it = 0; it < cta_tiles; ++it) {
            const int s = it % S;
            const int64_t g0 = cta_g0 + static_cast<int64_t>(it) * p.graphs_per_tile;
            const int ng = min(p.graphs_per_tile, cta_ng - it * p.graphs_per_tile);
            const int rows = ng * N;
            const uint32_t st = base + p.off_stage + static_cast<uint32_t>(s) * p.stage_bytes;
            const uint32_t us = st + p.st_u, ds = st + p.st_d;

            // ---- 1. stage landed ----
            mbar_wait(&bar_full[s], (it / S) & 1);
            const StageInfo si = sinfo[s];
            const uint32_t rp_addr = st + p.st_rp + 4u * static_cast<uint32_t>(si.rp_skip);

            // ---- 2a. {column, value} -> {byte offset of the neighbour's dU row, value};  dU in place ----
            if (si.staged) {
                const uint32_t col_a = st + p.st_col + 4u * static_cast<uint32_t>(si.e_skip);
                const uint32_t val_a = st + p.st_val + 4u * static_cast<uint32_t>(si.e_skip);
                for (int k = tid; k < si.n_entries; k += kConsumers) {
                    int m = 0;
                    while (m + 1 < ng && si.e_first + k >= static_cast<int>(lds_u32(rp_addr + 4u * (m + 1) * N))) ++m;
                    const uint32_t off = lds_u32(col_a + 4u * k) * row_pitch + static_cast<uint32_t>(m) * graph_bytes_u;
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(cv_addr + 8u * k), "r"(off), "r"(lds_u32(val_a + 4u * k)) : "memory");
                }
            }
            if (p.d_src != 0) {
                for (int idx = tid; idx < rows * f4o; idx += kConsumers) {
                    float g[4];
                    if (p.d_src == 2) {
                        const int r = idx / f4o, c4 = idx - r * f4o;
                        lds_f<4>(g, ds + 16u * static_cast<uint32_t>((r / N) * f4o + c4));
                    } else {
                        lds_f<4>(g, ds + 16u * idx);
                    }
                    if (p.u_src == 1) {
                        float yv[4];
                        lds_f<4>(yv, us + 16u * idx);
#pragma unroll
                        for (int t = 0; t < 4; ++t) g[t] *= act_grad_from_output(yv[t], p.act);
                    }
                    sts_f<4>(us + 16u * idx, g);
                }
            }
            consumer_sync<kConsumers>();

            // ---- 2b. G = A^T.dU into both operand layouts;  x -> [Xhi;Xlo] (MN-major) ----
            {
                BwdCtx a{zhi, zlo, z2hi, z2lo, us, rp_addr, cv_addr, si.e_first, rows, N, f_out, row_pitch, p.col, p.val};
                if (si.staged) aggregate_bwd<HAS_DX>(a, group, n_groups, sub, csum);
                else aggregate_bwd_unstaged<HAS_DX>(a, group, n_groups, sub, csum);
            }
            for (int idx = tid; idx < rows * f4i; idx += kConsumers) {
                const int r = idx / f4i, c4 = idx - r * f4i;
                float v[4], hi[4], lo[4];
                lds_f<4>(v, st + 16u * idx);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    hi[t] = tf32_hi(v[t]);
                    lo[t] = v[t] - hi[t];
                }
                const uint32_t off = mn32_offset(r, c4 << 2, kMnLbo);
                sts_f<4>(xhi + off, hi);
                sts_f<4>(xlo + off, lo);
            }
            if (rows < kRows) {   // short tile: the missing rows are K entries of dW -> exact zeros in both operands
                const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                const int pad = kRows - rows;
                for (int idx = tid; idx < pad * (kMi / 4); idx += kConsumers) {
                    const uint32_t off = mn32_offset(rows + idx / (kMi / 4), (idx % (kMi / 4)) << 2, kMnLbo);
                    sts_f<4>(xhi + off, z4);
                    sts_f<4>(xlo + off, z4);
                }
                for (int idx = tid; idx < pad * (KGp / 4); idx += kConsumers) {
                    const uint32_t off = mn32_offset(rows + idx / (KGp / 4), (idx % (KGp / 4)) << 2, kMnLbo);
                    sts_f<4>(z2hi + off, z4);
                    sts_f<4>(z2lo + off, z4);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            consumer_sync<kConsumers>();

            // ---- 3. tensor cores ----
            if (warp < (HAS_DX ? 3 : 1)) {
                if (elect_one()) {
                    tc_fence_after_sync();
                    const int role = HAS_DX ? warp : 2;
                    if (role == 2) {
                        uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
                        for (int ks = 0; ks < kRows / 8; ++ks) {   // 8 rows (1024 B) per K step
                            umma_tf32(d2, desc_x + static_cast<uint64_t>(64 * ks), desc_g + static_cast<uint64_t>(64 * ks), idesc_dw, acc);
                            acc = 1;
                        }
                    } else {
                        const uint64_t da = role == 0 ? desc_zhi : desc_zlo;
                        const uint32_t d = role == 0 ? d1a : d1b;
                        const uint32_t idesc = role == 0 ? idesc_dxa : idesc_dxb;
                        switch (n_katoms) {
                            case 1: issue_pass<1>(d, da, desc_w, idesc, kZAtom >> 4, w_atom >> 4, f_out); break;
                            case 2: issue_pass<2>(d, da, desc_w, idesc, kZAtom >> 4, w_atom >> 4, f_out); break;
                            default: issue_pass_loop(d, da, desc_w, idesc, kZAtom >> 4, w_atom >> 4, f_out, n_katoms);
                        }
                    }
                    umma_commit(&bar_mma);
                }
                __syncwarp();
            }
            mbar_wait(&bar_mma, it & 1);
            tc_fence_after_sync();
            if (tid == 0) mbar_arrive(&bar_empty[s]);   // x / dU / CSR of this stage are consumed: refill it

            // ---- 4. dx tile: TMEM -> registers -> staged tile -> global ----
            if (HAS_DX) {
                const uint32_t ys = zhi;   // free once the MMAs have completed
                const int q = warp & 3;
                const uint32_t ra = q * 16 + (lane >> 2), rb = ra + 8;
                const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
                for (int slab = warp >> 2; slab * 16 < f_in; slab += 4) {
                    float a[8], b[8], c[8];
                    tmem_ld_16x256b_x2(d1a + lane_bits + static_cast<uint32_t>(slab * 16), a);
                    tmem_ld_16x256b_x2(d1a + lane_bits + static_cast<uint32_t>(Fip + slab * 16), b);
                    tmem_ld_16x256b_x2(d1b + lane_bits + static_cast<uint32_t>(slab * 16), c);
                    tmem_ld_wait();
                    tmem_ld_fence8(a);
                    tmem_ld_fence8(b);
                    tmem_ld_fence8(c);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint32_t colx = slab * 16 + 8 * i + 2 * (lane & 3);
                        if (colx < static_cast<uint32_t>(f_in)) {
                            const float oa[2] = {a[4 * i] + (b[4 * i] + c[4 * i]), a[4 * i + 1] + (b[4 * i + 1] + c[4 * i + 1])};
                            const float ob[2] = {a[4 * i + 2] + (b[4 * i + 2] + c[4 * i + 2]), a[4 * i + 3] + (b[4 * i + 3] + c[4 * i + 3])};
                            sts_f<2>(ys + ystage_off(ra, colx, p.y_pitch), oa);
                            sts_f<2>(ys + ystage_off(rb, colx, p.y_pitch), ob);
                        }
                    }
                }
                tc_fence_before_sync();
                consumer_sync<kConsumers>();
                copy_out<kConsumers>(ys, p.y_pitch, p.dx + g0 * N * f_in, rows, f_in, tid, dx_vec4);
                tc_fence_before_sync();
                consumer_sync<kConsumers>();   // staging tile and accumulators consumed before the next tile
            }
        }

        // ---- per-CTA partial dW (+ dbias): accumulator rows 0..63 = Xhi^T.[Ghi|Glo], rows 64..127 = Xlo^T.[Ghi|Glo] ----
        {
            tc_fence_after_sync();
            const int q = warp & 3;
            const int fi = (q & 1) * 32 + lane;
            const bool lo_part = q >= 2;
            const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
            const uint32_t spitch = static_cast<uint32_t>(KGp + 4) * 4u;    // padded rows: conflict-free 16-byte accesses
            const uint32_t scratch = z2hi;                                   // the G operand is dead now
            float* part_out = p.partial + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(f_in) + 1) * f_out;
            if (lo_part) {
                for (int j = warp >> 2; j * 16 < f_out; j += 4) {
                    float v[16], v2[16];
                    tmem_ld16(d2 + lane_bits + static_cast<uint32_t>(j * 16), v);
                    tmem_ld16(d2 + lane_bits + static_cast<uint32_t>(KGp + j * 16), v2);
                    tmem_ld_wait();
                    tmem_ld_fence(v);
                    tmem_ld_fence(v2);
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float o[4] = {v[4 * qd] + v2[4 * qd], v[4 * qd + 1] + v2[4 * qd + 1], v[4 * qd + 2] + v2[4 * qd + 2],
                                            v[4 * qd + 3] + v2[4 * qd + 3]};
                        sts_f<4>(scratch + fi * spitch + 4u * (j * 16 + 4 * qd), o);
                    }
                }
            }
            if (sub * 4 < f_out) sts_f<4>(xhi + 4u * static_cast<uint32_t>(group * f_out + sub * 4), csum);   // x operand is dead too
            consumer_sync<kConsumers>();
            if (!lo_part) {
                for (int j = warp >> 2; j * 16 < f_out; j += 4) {
                    float v[16], v2[16];
                    tmem_ld16(d2 + lane_bits + static_cast<uint32_t>(j * 16), v);
                    tmem_ld16(d2 + lane_bits + static_cast<uint32_t>(KGp + j * 16), v2);
                    tmem_ld_wait();
                    tmem_ld_fence(v);
                    tmem_ld_fence(v2);
                    if (fi < f_in) {
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const int colw = j * 16 + 4 * qd;
                            if (colw < f_out) {
                                float lo4[4];
                                lds_f<4>(lo4, scratch + fi * spitch + 4u * colw);
                                *reinterpret_cast<float4*>(part_out + static_cast<size_t>(fi) * f_out + colw) =
                                    make_float4((v[4 * qd] + v2[4 * qd]) + lo4[0], (v[4 * qd + 1] + v2[4 * qd + 1]) + lo4[1],
                                                (v[4 * qd + 2] + v2[4 * qd + 2]) + lo4[2], (v[4 * qd + 3] + v2[4 * qd + 3]) + lo4[3]);
                            }
                        }
                    }
                }
            }
            if (tid < f_out) {
                float sacc = 0.0f;
                for (int g = 0; g < n_groups; ++g) sacc += __uint_as_float(lds_u32(xhi + 4u * static_cast<uint32_t>(g * f_out + tid)));
                part_out[static_cast<size_t>(f_in) * f_out + tid] = sacc;
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, p.tmem_cols);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr int kSmemMax = 227 * 1024 - 512;   // static __shared__ (barriers, stage info) shares the 227 KB

bool plan_bwd(BwdParams& p, int64_t n_graphs, int n_nodes, int f_in, int f_out, int act, bool dy_bcast, bool has_dx) {
    if (n_graphs <= 0 || n_nodes < 1 || n_nodes > kRows) return false;
    if (f_in < 4 || f_in > kMi || f_out < 4 || f_out > 128 || (f_in & 3) || (f_out & 3)) return false;
    p.KGp = static_cast<int>(up(f_out, 32));
    p.Fip = static_cast<int>(up(f_in, 16));
    p.graphs_per_tile = std::max(1, kRows / n_nodes);
    while (p.graphs_per_tile > 1 && ceil_div<int64_t>(n_graphs, p.graphs_per_tile) < kNumSMs) --p.graphs_per_tile;
    p.n_tiles = static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_tile));
    const uint32_t rows_max = static_cast<uint32_t>(p.graphs_per_tile) * n_nodes;
    const uint32_t n_katoms = p.KGp / 32;
    int lpr_log2 = 0;
    while ((1 << lpr_log2) * 4 < f_out) ++lpr_log2;
    p.lpr_log2 = lpr_log2;
    p.u_src = act != KGCN_ACT_NONE ? 1u : (dy_bcast ? 0u : 2u);
    p.d_src = dy_bcast ? 2u : (act != KGCN_ACT_NONE ? 1u : 0u);
    uint32_t off = 0;
    p.off_z = off;  if (has_dx) off += 2 * n_katoms * kZAtom;
    p.off_w = off;  if (has_dx) off += n_katoms * 2u * p.Fip * 128u;
    p.off_z2 = off; off += 2 * n_katoms * kMnLbo;
    p.off_x = off;  off += 2 * (kMi / 32) * kMnLbo;
    p.off_stage = off;
    p.y_pitch = up(f_in * 4u, 128);
    if (has_dx && p.y_pitch * kRows > 2 * n_katoms * kZAtom) return false;
    // end-of-kernel scratch: dW lo block in the G operand region, dbias partials in the x operand region
    if (kMi * (p.KGp + 4u) * 4u > 2 * n_katoms * kMnLbo) return false;
    if ((static_cast<uint32_t>(kConsumers) >> lpr_log2) * f_out * 4u > 2 * (kMi / 32) * kMnLbo) return false;
    p.cv_cap = static_cast<int>(up(std::max<uint32_t>(256, 6 * rows_max), 4));
    p.st_u = up(rows_max * f_in * 4u, 128);
    p.st_d = p.st_u + up(rows_max * f_out * 4u, 128);
    const uint32_t d_bytes = p.d_src == 1 ? rows_max * f_out * 4u : (p.d_src == 2 ? p.graphs_per_tile * f_out * 4u : 0u);
    p.st_rp = p.st_d + up(d_bytes, 128);
    p.st_col = p.st_rp + up((rows_max + 8) * 4u, 16);
    p.st_val = p.st_col + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u;
    p.stage_bytes = up(p.st_val + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u, 128);
    const uint32_t fixed = static_cast<uint32_t>(p.cv_cap) * 8u + 1024u;
    p.n_stages = 0;
    for (int st = kStagesMax; st >= 1; --st)
        if (off + st * p.stage_bytes + fixed <= static_cast<uint32_t>(kSmemMax)) { p.n_stages = st; break; }
    if (p.n_stages == 0) return false;
    off += p.n_stages * p.stage_bytes;
    p.off_cv = off; off += static_cast<uint32_t>(p.cv_cap) * 8u;
    p.smem_total = off + 1024;
    const uint32_t need_cols = (has_dx ? 3u * p.Fip : 0u) + 2u * p.KGp;
    uint32_t cols = 32;
    while (cols < need_cols) cols <<= 1;
    if (cols > 512) return false;
    p.tmem_cols = cols;
    return true;
}

bool fused_bwd_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_FUSED_BWD");   // tuning / A-B knob: 0 forces the decomposed backward
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}

}  // namespace

size_t fused_bwd_partial_bytes(int f_in, int f_out) {
    return static_cast<size_t>(kNumSMs) * (static_cast<size_t>(f_in) + 1) * f_out * sizeof(float);
}

bool fused_bwd_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int act, bool dy_bcast,
                        const float* x, const float* w, const float* y, const float* dy, const float* dx,
                        const int32_t* rowptr, const int32_t* col, const float* val) {
    if (!fused_bwd_enabled() || channels != 1) return false;
    BwdParams p{};
    if (!plan_bwd(p, n_graphs, n_nodes, f_in, f_out, act, dy_bcast, dx != nullptr)) return false;
    if (n_graphs * static_cast<int64_t>(n_nodes) >= (1ll << 31)) return false;
    return aligned16(x) && aligned16(w) && aligned16(dy) && (y == nullptr || aligned16(y)) &&
           (reinterpret_cast<uintptr_t>(dx) & 3u) == 0 && aligned16(rowptr) && aligned16(col) && aligned16(val);
}

int launch_graphconv_fused_bwd(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                               int n_nodes, const float* x, int f_in, const float* w, int f_out, int act, const float* y,
                               const float* dy, bool dy_bcast, float* dx, float* dw, float* dbias, void* workspace,
                               size_t workspace_bytes, cudaStream_t st) {
    BwdParams p{};
    KGCN_REQUIRE(plan_bwd(p, n_graphs, n_nodes, f_in, f_out, act, dy_bcast, dx != nullptr), KGCN_ERR_UNSUPPORTED,
                 "fused GraphConv backward: shape does not fit one SM's shared memory");
    const int64_t grid0 = std::min<int64_t>(p.n_tiles, kNumSMs);
    p.graphs_per_cta = static_cast<int>(ceil_div<int64_t>(n_graphs, grid0));
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta));   // every CTA owns >= 1 graph
    const size_t need = static_cast<size_t>(grid) * (static_cast<size_t>(f_in) + 1) * f_out * sizeof(float);
    KGCN_REQUIRE(workspace != nullptr && workspace_bytes >= need && aligned16(workspace), KGCN_ERR_WORKSPACE,
                 "fused GraphConv backward: workspace %zu < %zu bytes", workspace_bytes, need);
    p.rowptr = rowptr_t; p.col = col_t; p.val = val_t; p.x = x; p.w = w; p.y = y; p.dy = dy; p.dx = dx;
    p.partial = static_cast<float*>(workspace);
    p.n_graphs = n_graphs; p.n_nodes = n_nodes; p.f_in = f_in; p.f_out = f_out; p.act = act; p.dy_bcast = dy_bcast ? 1 : 0;
    auto go = [&](auto kernel) -> int {
        KGCN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem_total)));
        launch_pdl(kernel, grid, kBlock, p.smem_total, st, p);
        KGCN_LAUNCH_OK("graphconv_fused_bwd_kernel");
        return KGCN_OK;
    };
    const int rc = dx != nullptr ? go(graphconv_fused_bwd_kernel<true>) : go(graphconv_fused_bwd_kernel<false>);
    if (rc) return rc;
    return launch_splitk_reduce(p.partial, static_cast<int>(grid), f_in, f_out, dw, dbias, st);
}

}  // namespace kgcn
