// Fused GraphConv weight gradient for sm_100a: ONE kernel computes, for all adjacency channels c of a layer,
//
//     G_c   = A_c^T . dU                      (adjoint_a=True of bspmm_call.py:44 / bconv_call.py:46-57; transposed CSR)
//     dW_c  = X^T . G_c ,  dbias_c = column sums of G_c            (backward of kgcn/layers.py:112-115)
//
// where dU = dy (.) act'(y) comes from the caller.  The input gradient dx = sum_c G_c . W_c^T of the same layer is
// the fused layer kernel (graphconv_fused_v4.cu) run on (A^T, dU, W^T), so the backward of a layer is two streaming
// launches that share the concurrent-role design: no phase of a CTA waits for another phase of the same tile.
//
// Per persistent CTA (one per SM, contiguous graph range), roles on different chunks at once:
//   TMA producer warp   ring of stages: the tile's x rows, dU rows and transposed-CSR slices (cp.async.bulk)
//   8 worker warps      one THREAD per (tile row, 32-column slab): G rows by a gather over the row's CSR entries (8 LDS.128
//                       per entry, 16-byte chunk order rotated per lane -> conflict-free although every dU row starts at
//                       the same bank), x rows by a plain copy; both are split into tf32 hi / lo and written MN-major
//                       (SWIZZLE_128B, 32-byte base: the contraction index of X^T.G is the ROW index, so both operands are
//                       "transposed" tiles that tcgen05 reads without a transposition pass).  Double-buffered operands.
//   MMA warp            3xTF32 into ONE fp32 accumulator [f_in x C*f_out] in tensor memory that lives across all tiles of
//                       the CTA; a second accumulator takes ones^T.[Ghi|Glo] = the column sums of G (dbias).
// At the end the CTA writes one partial block [(f_in + 1), C*f_out] (last row = dbias); a fixed-order reduction over
// CTAs follows (splitk_reduce_kernel / the fused reduce + Adam tail).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "fused_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

constexpr int kWorkWarps = 8;
constexpr int kWarpMma = kWorkWarps, kWarpTma = kWorkWarps + 1;
constexpr int kBlock = (kWorkWarps + 2) * 32;
constexpr int kStagesMax = 4;

struct DwParams {
    const int32_t* rowptr;   // transposed BatchedCSR
    const int32_t* col;
    const float* val;
    const float* x;          // [B, N, f_in]
    const float* du;         // [B, N, f_out]
    const float* g;          // optional [B, N, C * f_out]: G = [A_0^T . dU | A_1^T . dU | ..] already computed -- its rows are copied, not gathered
    float* partial;          // [grid][(f_in + 1) * Ng]
    int64_t n_graphs;
    int C, N, f_in, f_out, Ng;
    int G, graphs_per_cta;
    int R;                   // operand chunk: rows (= K of one MMA batch), multiple of 8, <= 64
    int stacked;             // f_in <= 64: [Xhi ; Xlo] is one M = 128 operand
    int n_xs, n_gs, spc;     // 32-column slabs of x, of G = [G_0 | G_1 | ..], slabs per channel
    int n_stages, opbufs, cv_cap;
    uint32_t off_ones, off_op, op_bytes, op_g, lbo, off_stage, stage_bytes, st_du, st_rp, st_col, st_val, smem_total;
    uint32_t tmem_cols;
    uint32_t tm_off;         // first tensor-memory column of this job's accumulators (dW at tm_off, dbias at tm_off + Ng)
    int fresh;               // 0: same plan as the previous job of the launch -- no job boundary at all (soft boundary)
};

struct Ring {
    int idx = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
        if (++idx == n) {
            idx = 0;
            phase ^= 1u;
        }
    }
};

__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {   // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}

// 32 values of one (row, slab) -> tf32 hi / lo, MN-major: the row's 128-byte line of slab `chunk`; inside the line the
// 32-byte granule g sits at g ^ (row & 3).  acc block i holds 16-byte chunk q = i ^ s7 of the slab (the gather's rotation),
// so no un-rotation is needed: the 8 lanes of a store phase hit 8 different 16-byte positions (bijective in lane & 7).
__device__ __forceinline__ void store_split(uint32_t line_hi, uint32_t line_lo, uint32_t row, uint32_t s7, const float (&acc)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t q = static_cast<uint32_t>(i) ^ s7;
        const uint32_t pos = ((((q >> 1) ^ (row & 3u)) << 1) | (q & 1u)) << 4;
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __uint_as_float(__float_as_uint(acc[4 * i + j]) & 0xFFFFE000u);   // truncation: lo = x - hi is exact
            lo[j] = acc[4 * i + j] - hi[j];
        }
        sts_f<4>(line_hi + pos, hi);
        sts_f<4>(line_lo + pos, lo);
    }
}

// Several layers' weight gradients in ONE launch ("jobs"): after the dx chain every layer's dU exists, so the CTA walks
// its graph range once per layer, each job with its own accumulators in tensor memory (columns tm_off .. tm_off + 2 Ng).
// Job boundary: the MMA warp waits for its last MMAs, then one CTA barrier -- shared-memory layouts may differ per job.
// Consecutive jobs with the SAME plan (the equal-width hidden layers of a network) have no boundary: the jobs are
// independent (each reads its own x / dU and owns its accumulators), the padding and the ones operand stay valid, and
// every ring continues with its per-slot phase bits -- the producer runs ahead into the next job while the last MMAs
// of the previous one are still in flight, so the chain pays fill and drain once.
constexpr int kDwMaxJobs = 4;
struct DwBatch {
    int n_jobs;
    uint32_t tmem_cols;
    long long* dbg;   // tuning aid (kgcn_debug_dw_times): [CTA][256] clock64 stamps, see tools/dw_timeline.py
    DwParams job[kDwMaxJobs];
};

// stamp slots per CTA: 0 kernel entry, 1 workers left the job loop, 2 partials written, 3 after the setup barrier;
// 8 + 8 * T + e for the CTA's T-th tile over all jobs: e = 0 producer issued the tile's copies, 1 worker warp 0 saw the stage full,
// 2 worker warp 0 finished the tile's last chunk, 3 MMA warp saw the tile's first chunk, 4 MMA warp issued the tile's last chunk
// (compiled in only with -DKGCN_DW_TIMELINE: `make TIMELINE=1`; the product build carries no stamp code)
#ifdef KGCN_DW_TIMELINE
#define DW_STAMP(slot) do { if (b.dbg != nullptr && lane == 0) b.dbg[static_cast<size_t>(blockIdx.x) * 256 + (slot)] = clock64(); } while (0)
#else
#define DW_STAMP(slot) do { } while (0)
#endif

__device__ __forceinline__ void bar_cta_roles() { asm volatile("bar.sync 2, %0;" ::"n"(kBlock) : "memory"); }

__global__ void __launch_bounds__(kBlock, 1) graphconv_fused_dw_kernel(const DwBatch b) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kStagesMax], bar_empty[kStagesMax], bar_opfull[2], bar_opempty[2], bar_done;
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_jobs = b.n_jobs;
    if (warp == 0) DW_STAMP(0);
    [[maybe_unused]] int tile_no = 0;   // tiles of all earlier jobs (stamps only)

    if (tid == 0) {
        for (int i = 0; i < kStagesMax; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], kWorkWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_opfull[i], kWorkWarps);
            mbar_init(&bar_opempty[i], 1);
        }
        mbar_init(&bar_done, 1);
        fence_mbar_init();
    }
    if (warp == kWarpMma) tmem_alloc(&tmem_slot, b.tmem_cols);
    pdl_wait();   // everything above overlaps the previous kernel's tail
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    // one phase bit per barrier slot and role (ring sizes may differ per job)
    uint32_t ph_a = 0, ph_b = 0, ph_done = 0;

    for (int jb = 0; jb < n_jobs; ++jb) {
        const DwParams& p = b.job[jb];
        const int C = p.C, N = p.N, f_in = p.f_in, f_out = p.f_out, Ng = p.Ng, S = p.n_stages, R = p.R;
        const uint32_t pitch_x = static_cast<uint32_t>(f_in) * 4u, pitch_u = static_cast<uint32_t>(f_out) * 4u;
        const int64_t g_begin = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
        const int64_t left = p.n_graphs - g_begin;
        const int n_graphs_cta = static_cast<int>(left < p.graphs_per_cta ? (left > 0 ? left : 0) : p.graphs_per_cta);
        const int n_tiles = (n_graphs_cta + p.G - 1) / p.G;
        const int last_ng = n_graphs_cta - (n_tiles - 1) * p.G;
        const uint32_t tm = tmem + p.tm_off;

        if (jb > 0 && p.fresh) bar_cta_roles();   // the previous job has finished with shared memory (its MMAs have completed)
        if (warp != kWarpTma && p.fresh) {
            // Only what no worker ever writes has to be zeroed: the ones operand (M = 64: two 32-row chunks; column 0 of every K
            // row is 1.0: element m = 0 sits in granule 0 -> position (row & 3) << 5 of the row's line) and the M-padding chunks
            // of the X operand (f_in < 64 stacked / f_in < 128 split).  Chunks that hold data are fully rewritten for every
            // operand chunk, rows past a short chunk are written as zeros by the workers, the MMAs read whole K steps only.
            constexpr int kSetup = kBlock - 32;
            const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const uint32_t n16_ones = (2u * p.lbo) >> 4;
            for (uint32_t i = tid; i < n16_ones; i += kSetup) sts_f<4>(base + p.off_ones + (i << 4), z4);
            const int half = p.stacked ? 2 : 4;       // chunks per hi / lo half of the X operand
            const uint32_t lbo16 = p.lbo >> 4;
            for (int ob = 0; ob < p.opbufs; ++ob)
                for (int ch = 0; ch < 2 * half; ++ch)
                    if ((ch % half) >= p.n_xs) {
                        const uint32_t cb = base + p.off_op + static_cast<uint32_t>(ob) * p.op_bytes + static_cast<uint32_t>(ch) * p.lbo;
                        for (uint32_t i = tid; i < lbo16; i += kSetup) sts_f<4>(cb + (i << 4), z4);
                    }
            asm volatile("bar.sync 3, %0;" ::"n"(kSetup) : "memory");
            if (tid < R) {
                const float one[1] = {1.0f};
                sts_f<1>(base + p.off_ones + static_cast<uint32_t>(tid) * 128u + ((static_cast<uint32_t>(tid) & 3u) << 5), one);
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync 3, %0;" ::"n"(kSetup) : "memory");
            if (warp == 0 && jb == 0) DW_STAMP(3);
        }

        if (warp == kWarpTma) {
            // =============================== TMA producer ===============================
            if (lane == 0) {
                int s = 0;
                for (int it = 0; it < n_tiles; ++it) {
                    mbar_wait_relaxed(&bar_empty[s], ((ph_a >> s) & 1u) ^ 1u);
                    ph_a ^= 1u << s;
                    const int64_t g0 = g_begin + static_cast<int64_t>(it) * p.G;
                    const int ng = (it == n_tiles - 1) ? last_ng : p.G;
                    const int64_t r0 = g0 * C * N;
                    const int rows_csr = ng * C * N;
                    unsigned char* st = gen + p.off_stage + static_cast<size_t>(s) * p.stage_bytes;
                    uint64_t* full = &bar_full[s];
                    const int64_t rp_lo = r0 & ~3ll;
                    const uint32_t rp_cnt = static_cast<uint32_t>((r0 + rows_csr + 1 - rp_lo + 3) & ~3ll);
                    const uint32_t x_bytes = static_cast<uint32_t>(ng * N) * pitch_x, u_bytes = static_cast<uint32_t>(ng * N) * pitch_u;
                    if (p.g != nullptr) {   // G rows take the place of the dU rows; no CSR slices
                        const uint32_t g_bytes = u_bytes * static_cast<uint32_t>(C);
                        mbar_expect_tx(full, x_bytes + g_bytes);   // the one arrival of the phase
                        bulk_g2s(st + p.st_du, p.g + g0 * N * Ng, g_bytes, full);
                        bulk_g2s(st, p.x + g0 * N * f_in, x_bytes, full);
                        if (tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it));
                        if (++s == S) s = 0;
                        continue;
                    }
                    mbar_expect_tx_only(full, x_bytes + u_bytes + 4u * rp_cnt);
                    bulk_g2s(st + p.st_du, p.du + g0 * N * f_out, u_bytes, full);
                    bulk_g2s(st, p.x + g0 * N * f_in, x_bytes, full);
                    bulk_g2s(st + p.st_rp, p.rowptr + rp_lo, 4u * rp_cnt, full);
                    const int32_t e_first = __ldg(p.rowptr + r0), e_last = __ldg(p.rowptr + r0 + rows_csr);
                    const int32_t e_lo = e_first & ~3;
                    const uint32_t e_cnt = static_cast<uint32_t>((e_last - e_lo + 3) & ~3);
                    const bool staged = e_cnt <= static_cast<uint32_t>(p.cv_cap) && e_cnt != 0;
                    mbar_expect_tx(full, staged ? 8u * e_cnt : 0u);   // the one arrival of the phase
                    if (staged) {
                        bulk_g2s(st + p.st_col, p.col + e_lo, 4u * e_cnt, full);
                        bulk_g2s(st + p.st_val, p.val + e_lo, 4u * e_cnt, full);
                    }
                    if (tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it));
                    if (++s == S) s = 0;
                }
            }
            __syncwarp();
        } else if (warp == kWarpMma) {
            // =============================== MMA issuer ===============================
            const uint32_t idesc_w = umma_idesc_tf32(128, Ng) | kUmmaMajorMnA | kUmmaMajorMnB;
            const uint32_t idesc_b = umma_idesc_tf32(64, Ng) | kUmmaMajorMnA | kUmmaMajorMnB;
            const uint32_t d_w = tm, d_b = tm + static_cast<uint32_t>(Ng);
            const uint64_t desc_ones = umma_desc_mn32(base + p.off_ones, p.lbo, 512u);
            const uint32_t lbo16 = p.lbo >> 4;
            int ob = 0;
            uint32_t acc = 0;
            for (int it = 0; it < n_tiles; ++it) {
                const int rows_t = ((it == n_tiles - 1) ? last_ng : p.G) * N;
                for (int c0 = 0; c0 < rows_t; c0 += R) {
                    const int ksteps = (min(R, rows_t - c0) + 7) >> 3;
                    mbar_wait(&bar_opfull[ob], (ph_a >> ob) & 1u);
                    ph_a ^= 1u << ob;
                    tc_fence_after_sync();
                    __syncwarp();
                    if (c0 == 0 && tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it) + 3);
                    if (elect_one()) {
                        const uint32_t op = base + p.off_op + static_cast<uint32_t>(ob) * p.op_bytes;
                        const uint64_t dxh = umma_desc_mn32(op, p.lbo, 512u);                         // Xhi (stacked: [Xhi ; Xlo])
                        const uint64_t dxl = dxh + static_cast<uint64_t>(4u * lbo16);                 // Xlo (not stacked)
                        const uint64_t dgh = umma_desc_mn32(op + p.op_g, p.lbo, 512u);                // Ghi
                        const uint64_t dgl = dgh + static_cast<uint64_t>(static_cast<uint32_t>(p.n_gs) * lbo16);   // Glo
                        for (int ks = 0; ks < ksteps; ++ks) {   // 8 rows (1024 B) per K step
                            const uint64_t o = static_cast<uint64_t>(64 * ks);
                            umma_tf32(d_w, dxh + o, dgh + o, idesc_w, acc);
                            umma_tf32(d_b, desc_ones + o, dgh + o, idesc_b, acc);
                            acc = 1;
                            umma_tf32(d_w, dxh + o, dgl + o, idesc_w, 1);
                            umma_tf32(d_b, desc_ones + o, dgl + o, idesc_b, 1);
                            if (!p.stacked) umma_tf32(d_w, dxl + o, dgh + o, idesc_w, 1);
                        }
                        umma_commit(&bar_opempty[ob]);   // the operand buffer may be overwritten once these MMAs have read it
                    }
                    __syncwarp();
                    if (c0 + R >= rows_t && tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it) + 4);
                    if (++ob == p.opbufs) ob = 0;
                }
            }
            if (jb + 1 == n_jobs || b.job[jb + 1].fresh) {
                if (elect_one()) umma_commit(&bar_done);
                __syncwarp();
                mbar_wait(&bar_done, ph_done & 1u);   // all MMAs so far have completed (operands + accumulators final)
                ph_done ^= 1u;
            }
        } else {
            // =============================== worker warps ===============================
            const uint32_t s7 = static_cast<uint32_t>(lane) & 7u;
            const int n_cs = p.n_gs + p.n_xs;
            const uint32_t lo_x = (p.stacked ? 2u : 4u) * p.lbo, lo_g = static_cast<uint32_t>(p.n_gs) * p.lbo;
            const uint32_t r0_step = static_cast<uint32_t>(p.G * C * N);
            uint32_t r0_lo = static_cast<uint32_t>((g_begin * C * N) & 3);
            int s = 0, ob = 0;
            for (int it = 0; it < n_tiles; ++it) {
                const bool last = it == n_tiles - 1;
                const int rows_t = (last ? last_ng : p.G) * N;
                const int rows_csr = last ? last_ng * C * N : static_cast<int>(r0_step);
                const uint32_t st = base + p.off_stage + static_cast<uint32_t>(s) * p.stage_bytes;
                mbar_wait(&bar_full[s], (ph_a >> s) & 1u);
                ph_a ^= 1u << s;
                if (warp == 0 && tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it) + 1);
                const bool copy_g = p.g != nullptr;   // the stage holds finished G rows where the dU rows would be
                const uint32_t rp_addr = st + p.st_rp + 4u * (r0_lo & 3u);
                const int e_first = copy_g ? 0 : static_cast<int>(lds_u32(rp_addr));
                const int e_last = copy_g ? 0 : static_cast<int>(lds_u32(rp_addr + 4u * static_cast<uint32_t>(rows_csr)));
                const int e_lo = e_first & ~3;
                const bool staged = static_cast<uint32_t>((e_last - e_lo + 3) & ~3) <= static_cast<uint32_t>(p.cv_cap);
                const uint32_t col_addr = st + p.st_col - 4u * static_cast<uint32_t>(e_lo);   // entry e at col_addr + 4 e
                const uint32_t val_addr = st + p.st_val - 4u * static_cast<uint32_t>(e_lo);

                for (int c0 = 0; c0 < rows_t; c0 += R) {
                    const int rc8 = (min(R, rows_t - c0) + 7) & ~7;   // rows the MMAs of this chunk read (zero rows past the tile)
                    const int n_rb = (rc8 + 31) >> 5;
                    mbar_wait(&bar_opempty[ob], ((ph_b >> ob) & 1u) ^ 1u);
                    ph_b ^= 1u << ob;
                    tc_fence_after_sync();
                    const uint32_t op = base + p.off_op + static_cast<uint32_t>(ob) * p.op_bytes;
                    for (int j = warp; j < n_cs * n_rb; j += kWorkWarps) {
                        const int cs = j / n_rb, rb = j - cs * n_rb;
                        const int rl = rb * 32 + lane;          // row inside the chunk
                        const int r = c0 + rl;                  // row inside the tile
                        const bool valid = r < rows_t;
                        float acc[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
                        uint32_t line_hi, line_lo;
                        if (cs < p.n_gs && copy_g) {
                            // ---- G slab, already computed: the row's 128 bytes, same rotated chunk order as the gather leaves ----
                            if (valid) {
                                const uint32_t ga = st + p.st_du + static_cast<uint32_t>(r) * (pitch_u * static_cast<uint32_t>(C)) + static_cast<uint32_t>(cs) * 128u + (s7 << 4);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float t[4];
                                    lds_f<4>(t, ga ^ (static_cast<uint32_t>(i) << 4));
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = t[jj];
                                }
                            }
                            line_hi = op + p.op_g + static_cast<uint32_t>(cs) * p.lbo + static_cast<uint32_t>(rl) * 128u;
                            line_lo = line_hi + lo_g;
                        } else if (cs < p.n_gs) {
                            // ---- G slab: gather over the row's entries of channel c ----
                            const int c = cs / p.spc, sl = cs - c * p.spc;
                            const int gl = r / N, node = r - gl * N;
                            int e = e_first, e_end = e_first;
                            if (valid) {
                                const uint32_t ra = rp_addr + 4u * static_cast<uint32_t>((gl * C + c) * N + node);
                                e = static_cast<int>(lds_u32(ra));
                                e_end = static_cast<int>(lds_u32(ra + 4u));
                            }
                            const uint32_t ubase = st + p.st_du + static_cast<uint32_t>(gl * N) * pitch_u + static_cast<uint32_t>(sl) * 128u + (s7 << 4);
                            if (staged) {
                                uint32_t ce = col_addr + 4u * static_cast<uint32_t>(e), ve = val_addr + 4u * static_cast<uint32_t>(e);
                                const uint32_t cend = col_addr + 4u * static_cast<uint32_t>(e_end);
                                uint32_t cn = lds_u32(ce);   // one entry of look-ahead; reading one past the row is harmless (slack)
                                float vn = lds_f32(ve);
#pragma unroll 1
                                while (ce < cend) {
                                    const uint32_t ua = ubase + cn * pitch_u;
                                    const float v = vn;
                                    ce += 4;
                                    ve += 4;
                                    cn = lds_u32(ce);
                                    vn = lds_f32(ve);
                                    float uv[8][4];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) lds_f<4>(uv[i], ua ^ (static_cast<uint32_t>(i) << 4));
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
#pragma unroll
                                        for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = fmaf(v, uv[i][jj], acc[4 * i + jj]);
                                }
                            } else {   // unusually dense tile: the CSR slice did not fit the stage, entries come from global memory
                                for (; e < e_end; ++e) {
                                    const uint32_t ua = ubase + static_cast<uint32_t>(__ldg(p.col + e)) * pitch_u;
                                    const float v = __ldg(p.val + e);
                                    float uv[8][4];
#pragma unroll
                                    for (int i = 0; i < 8; ++i) lds_f<4>(uv[i], ua ^ (static_cast<uint32_t>(i) << 4));
#pragma unroll
                                    for (int i = 0; i < 8; ++i)
#pragma unroll
                                        for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = fmaf(v, uv[i][jj], acc[4 * i + jj]);
                                }
                            }
                            line_hi = op + p.op_g + static_cast<uint32_t>(cs) * p.lbo + static_cast<uint32_t>(rl) * 128u;
                            line_lo = line_hi + lo_g;
                        } else {
                            // ---- x slab: the row's 128 bytes, same rotated chunk order ----
                            const int xs = cs - p.n_gs;
                            if (valid) {
                                const uint32_t xa = st + static_cast<uint32_t>(r) * pitch_x + static_cast<uint32_t>(xs) * 128u + (s7 << 4);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    float t[4];
                                    lds_f<4>(t, xa ^ (static_cast<uint32_t>(i) << 4));
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = t[jj];
                                }
                            }
                            line_hi = op + static_cast<uint32_t>(xs) * p.lbo + static_cast<uint32_t>(rl) * 128u;
                            line_lo = line_hi + lo_x;
                        }
                        __syncwarp();
                        if (rl < rc8) store_split(line_hi, line_lo, static_cast<uint32_t>(rl), s7, acc);
                    }
                    fence_proxy_async_smem();   // operands are read by the tensor core through the async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_opfull[ob]);
                    if (++ob == p.opbufs) ob = 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);   // this warp is done reading the stage
                if (warp == 0 && tile_no + it < 30) DW_STAMP(8 + 8 * (tile_no + it) + 2);
                if (++s == S) s = 0;
                r0_lo += r0_step;
            }
        }
        tile_no += n_tiles;
    }
    if (warp == 0) DW_STAMP(1);

    // ---- per-CTA partials: accumulator rows -> [(f_in + 1), Ng] per job (last row = column sums of G = dbias partial) ----
    bar_cta_roles();          // the MMA warp has seen the last job's MMAs complete
    tc_fence_after_sync();
    if (warp < kWorkWarps) {
        for (int jb = 0; jb < n_jobs; ++jb) {
            const DwParams& p = b.job[jb];
            const int f_in = p.f_in, Ng = p.Ng;
            const uint32_t tm = tmem + p.tm_off;
            float* part_out = p.partial + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(f_in) + 1) * Ng;
            const int q = warp & 3, h = warp >> 2;            // TMEM lane quarter, column phase
            const uint32_t lane_bits = static_cast<uint32_t>(q * 32) << 16;
            const int row = q * 32 + lane;                    // accumulator row = TMEM lane
            const uint32_t scratch = base + p.off_op;         // operands are dead now: [64][Ng + 4] floats for the Xlo half
            const uint32_t spitch = static_cast<uint32_t>(Ng + 4) * 4u;
            if (p.stacked) {
                if (q >= 2) {
                    for (int j = h; j * 16 < Ng; j += kWorkWarps / 4) {
                        float v[16];
                        tmem_ld16(tm + lane_bits + static_cast<uint32_t>(j * 16), v);
                        tmem_ld_wait();
                        tmem_ld_fence(v);
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float o[4] = {v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]};
                            sts_f<4>(scratch + static_cast<uint32_t>(row - 64) * spitch + 4u * static_cast<uint32_t>(j * 16 + 4 * qd), o);
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kWorkWarps * 32) : "memory");
                if (q < 2) {
                    for (int j = h; j * 16 < Ng; j += kWorkWarps / 4) {
                        float v[16];
                        tmem_ld16(tm + lane_bits + static_cast<uint32_t>(j * 16), v);
                        tmem_ld_wait();
                        tmem_ld_fence(v);
                        if (row < f_in) {
#pragma unroll
                            for (int qd = 0; qd < 4; ++qd) {
                                float l4[4];
                                lds_f<4>(l4, scratch + static_cast<uint32_t>(row) * spitch + 4u * static_cast<uint32_t>(j * 16 + 4 * qd));
                                *reinterpret_cast<float4*>(part_out + static_cast<size_t>(row) * Ng + j * 16 + 4 * qd) =
                                    make_float4(v[4 * qd] + l4[0], v[4 * qd + 1] + l4[1], v[4 * qd + 2] + l4[2], v[4 * qd + 3] + l4[3]);
                            }
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kWorkWarps * 32) : "memory");   // scratch is reused by the next job
            } else {
                for (int j = h; j * 16 < Ng; j += kWorkWarps / 4) {
                    float v[16];
                    tmem_ld16(tm + lane_bits + static_cast<uint32_t>(j * 16), v);
                    tmem_ld_wait();
                    tmem_ld_fence(v);
                    if (row < f_in) {
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            *reinterpret_cast<float4*>(part_out + static_cast<size_t>(row) * Ng + j * 16 + 4 * qd) =
                                make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    }
                }
            }
            if (q == 0) {   // dbias accumulator: row 0 of the M = 64 block = TMEM lane 0
                for (int j = h; j * 16 < Ng; j += kWorkWarps / 4) {
                    float v[16];
                    tmem_ld16(tm + static_cast<uint32_t>(Ng + j * 16), v);
                    tmem_ld_wait();
                    tmem_ld_fence(v);
                    if (lane == 0) {
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd)
                            *reinterpret_cast<float4*>(part_out + static_cast<size_t>(f_in) * Ng + j * 16 + 4 * qd) =
                                make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
                    }
                }
            }
        }
    }

    if (warp == 0) DW_STAMP(2);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem, b.tmem_cols);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr int kSmemMax = 227 * 1024 - 1024;   // static __shared__ (barriers) shares the 227 KB

// g_wide: the stage holds finished G rows ([rows, C * f_out]) where the dU rows and the CSR slices would be (C > 1 only: for one
// channel G and dU have the same shape and the layout stays exactly the gather layout, so such a job can follow a gathering
// job of the same widths without a job boundary)
bool plan_dw_try(DwParams& p, int64_t n_graphs, int C, int N, int f_in, int f_out, int G, int R, int opbufs, int min_stages, bool g_wide = false) {
    p.C = C; p.N = N; p.f_in = f_in; p.f_out = f_out; p.Ng = C * f_out; p.n_graphs = n_graphs;
    p.stacked = f_in <= 64 ? 1 : 0;
    p.n_xs = f_in / 32;
    p.spc = f_out / 32;
    p.n_gs = p.Ng / 32;
    p.G = G; p.R = R; p.opbufs = opbufs;
    const int64_t grid0 = std::min<int64_t>(kNumSMs, n_graphs);
    const int64_t gpc = ceil_div<int64_t>(n_graphs, grid0);
    if (gpc > (1 << 24)) return false;
    p.graphs_per_cta = static_cast<int>(gpc);
    if (p.G > p.graphs_per_cta) p.G = p.graphs_per_cta;
    const uint32_t rows_max = static_cast<uint32_t>(p.G) * N;
    if (static_cast<uint32_t>(R) > up(rows_max, 8)) p.R = static_cast<int>(up(rows_max, 8));
    p.lbo = static_cast<uint32_t>(p.R) * 128u;
    uint32_t off = 0;
    p.off_ones = off; off += 2u * p.lbo;                              // M = 64: two 32-row chunks (the second stays zero)
    p.off_op = off;
    p.op_g = (p.stacked ? 4u : 8u) * p.lbo;                           // X operand: M = 128 (4 chunks), hi | lo
    p.op_bytes = p.op_g + 2u * static_cast<uint32_t>(p.n_gs) * p.lbo;  // G operand: Ghi | Glo
    off += static_cast<uint32_t>(opbufs) * p.op_bytes;
    if (p.stacked && 64u * (static_cast<uint32_t>(p.Ng) + 4u) * 4u > static_cast<uint32_t>(opbufs) * p.op_bytes) return false;   // end-of-kernel scratch
    p.off_stage = off;
    p.cv_cap = g_wide ? 4 : static_cast<int>(up(std::max<uint32_t>(256, 6 * rows_max * C), 4));
    p.st_du = up(rows_max * f_in * 4u, 128);
    p.st_rp = p.st_du + up(rows_max * static_cast<uint32_t>(g_wide ? p.Ng : f_out) * 4u, 128);
    p.st_col = p.st_rp + up((rows_max * C + 8) * 4u, 16);
    p.st_val = p.st_col + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u;
    p.stage_bytes = up(p.st_val + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u, 128);
    p.n_stages = 0;
    for (int st = kStagesMax; st >= min_stages; --st)
        if (off + st * p.stage_bytes + 1024 <= static_cast<uint32_t>(kSmemMax)) { p.n_stages = st; break; }
    if (p.n_stages == 0) return false;
    p.smem_total = off + p.n_stages * p.stage_bytes + 1024;
    uint32_t cols = 32;
    while (cols < 2u * static_cast<uint32_t>(p.Ng)) cols <<= 1;
    p.tmem_cols = cols;
    return cols <= 512;
}

bool plan_dw(DwParams& p, int64_t n_graphs, int C, int N, int f_in, int f_out, bool g_wide = false) {
    if (n_graphs <= 0 || C < 1 || C > 8 || N < 1 || N > 128) return false;
    if (f_in % 32 != 0 || f_out % 32 != 0 || f_in > 128 || f_in < 32 || C * f_out > 256) return false;
    // preference: >= 2 stages always; double-buffered operands as long as a chunk still fills the warps (>= 32 rows), then
    // single-buffered 64 / 32-row chunks (wide layers: F = 128), 16-row chunks last
    const int g_max = std::max(1, 64 / N);
    static const char* forced = getenv("KGCN_DW_PLAN");   // tuning / A-B knob: "opbufs:R" tried first (e.g. 2:32)
    if (forced != nullptr) {
        int ob = 0, r = 0;
        if (sscanf(forced, "%d:%d", &ob, &r) == 2 && (ob == 1 || ob == 2) && (r == 16 || r == 32 || r == 64))
            for (int G = g_max; G >= 1; G = (G > 1 ? G / 2 : 0))
                if (plan_dw_try(p, n_graphs, C, N, f_in, f_out, G, r, ob, 2, g_wide)) return true;
    }
    const int order[6][2] = {{2, 64}, {2, 32}, {1, 64}, {1, 32}, {2, 16}, {1, 16}};
    for (const auto& o : order)
        for (int G = g_max; G >= 1; G = (G > 1 ? G / 2 : 0))
            if (plan_dw_try(p, n_graphs, C, N, f_in, f_out, G, o[1], o[0], 2, g_wide)) return true;
    return plan_dw_try(p, n_graphs, C, N, f_in, f_out, 1, 32, 1, 1, g_wide);
}

}  // namespace

bool fused_dw_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_FUSED_DW");   // tuning / A-B knob: 0 forces the older backward paths
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}

bool soft_jobs_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_SOFT_JOBS");   // A-B knob: 0 restores the hard (barrier) boundary between all jobs of a chain
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}

size_t fused_dw_partial_bytes(int f_in, int n_total) {
    return static_cast<size_t>(kNumSMs) * (static_cast<size_t>(f_in) + 1) * n_total * sizeof(float);
}

bool fused_dw_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x, const float* du,
                       const int32_t* rowptr, const int32_t* col, const float* val) {
    if (!fused_dw_enabled()) return false;
    DwParams p{};
    if (!plan_dw(p, n_graphs, channels, n_nodes, f_in, f_out)) return false;
    if (n_graphs * static_cast<int64_t>(n_nodes) * channels >= (1ll << 31)) return false;
    return aligned16(x) && aligned16(du) && aligned16(rowptr) && aligned16(col) && aligned16(val);
}

// number of partial blocks the kernel writes for this shape (= its grid), 0 when the shape is not supported
int fused_dw_splits(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    DwParams p{};
    if (!fused_dw_enabled() || !plan_dw(p, n_graphs, channels, n_nodes, f_in, f_out)) return 0;
    return static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta));
}

static long long* g_dbg_dw = nullptr;

// jobs[k]: partial_k[grid][(f_in_k + 1)][channels * f_out_k]; no reduction.  All jobs share (n_graphs, channels, n_nodes).
int launch_graphconv_fused_dw_jobs(const DwJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, int* splits_out,
                                   cudaStream_t st) {
    KGCN_REQUIRE(n_jobs >= 1 && n_jobs <= kDwMaxJobs, KGCN_ERR_BAD_SHAPE, "fused GraphConv weight gradient: 1..%d jobs", kDwMaxJobs);
    DwBatch b{};
    b.n_jobs = n_jobs;
    b.dbg = g_dbg_dw;
    bool wide_job[kDwMaxJobs] = {};
    uint32_t cols = 0, smem = 0;
    for (int k = 0; k < n_jobs; ++k) {
        DwParams& p = b.job[k];
        const DwJob& j = jobs[k];
        // a precomputed G is used when its (for C > 1: wider) stage layout has a plan; else the job gathers from dU as without it
        const bool wide = j.g != nullptr && channels > 1;
        const bool use_g = j.g != nullptr && aligned16(j.g) && (!wide || plan_dw(p, n_graphs, channels, n_nodes, j.f_in, j.f_out, true));
        KGCN_REQUIRE((use_g && wide) || plan_dw(p, n_graphs, channels, n_nodes, j.f_in, j.f_out), KGCN_ERR_UNSUPPORTED,
                     "fused GraphConv weight gradient: unsupported shape");
        wide_job[k] = use_g && wide;
        KGCN_REQUIRE(p.graphs_per_cta == b.job[0].graphs_per_cta, KGCN_ERR_UNSUPPORTED, "fused GraphConv weight gradient: graph ranges differ");
        const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta));
        const size_t need = static_cast<size_t>(grid) * (static_cast<size_t>(j.f_in) + 1) * p.Ng * sizeof(float);
        KGCN_REQUIRE(j.partial != nullptr && j.partial_bytes >= need && aligned16(j.partial), KGCN_ERR_WORKSPACE,
                     "fused GraphConv weight gradient: workspace %zu < %zu bytes", j.partial_bytes, need);
        p.rowptr = j.rowptr_t; p.col = j.col_t; p.val = j.val_t; p.x = j.x; p.du = j.du;
        p.g = use_g ? j.g : nullptr;
        p.partial = j.partial;
        p.tm_off = cols;
        p.fresh = (k == 0 || !soft_jobs_enabled() || j.f_in != jobs[k - 1].f_in || j.f_out != jobs[k - 1].f_out || wide_job[k] != wide_job[k - 1]) ? 1 : 0;
        cols += 2u * static_cast<uint32_t>(p.Ng);
        smem = std::max(smem, p.smem_total);
    }
    KGCN_REQUIRE(cols <= 512, KGCN_ERR_UNSUPPORTED, "fused GraphConv weight gradient: %u tensor-memory columns for %d jobs", cols, n_jobs);
    uint32_t alloc = 32;
    while (alloc < cols) alloc <<= 1;
    b.tmem_cols = alloc;
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_graphs, b.job[0].graphs_per_cta));
    KGCN_CUDA_OK(cudaFuncSetAttribute(graphconv_fused_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    launch_pdl(graphconv_fused_dw_kernel, grid, kBlock, smem, st, b);
    KGCN_LAUNCH_OK("graphconv_fused_dw_kernel");
    if (splits_out != nullptr) *splits_out = static_cast<int>(grid);
    return KGCN_OK;
}

// how many consecutive jobs (same batch) fit one launch: 2 * channels * f_out tensor-memory columns each, 512 in total
int fused_dw_jobs_per_launch(int channels, const int* f_out, int n_jobs) {
    int cols = 0, k = 0;
    for (; k < n_jobs && k < kDwMaxJobs; ++k) {
        cols += 2 * channels * f_out[k];
        if (cols > 512) break;
    }
    return k;
}

int launch_graphconv_fused_dw_partial(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                      int channels, int n_nodes, const float* x, int f_in, const float* du, int f_out,
                                      float* partial, size_t partial_bytes, int* splits_out, cudaStream_t st) {
    const DwJob job{rowptr_t, col_t, val_t, x, du, partial, partial_bytes, f_in, f_out};
    return launch_graphconv_fused_dw_jobs(&job, 1, n_graphs, channels, n_nodes, splits_out, st);
}

int launch_graphconv_fused_dw(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs, int channels,
                              int n_nodes, const float* x, int f_in, const float* du, int f_out, float* dw, float* dbias,
                              void* workspace, size_t workspace_bytes, cudaStream_t st) {
    int splits = 0;
    const int rc = launch_graphconv_fused_dw_partial(rowptr_t, col_t, val_t, n_graphs, channels, n_nodes, x, f_in, du, f_out,
                                                     static_cast<float*>(workspace), workspace_bytes, &splits, st);
    if (rc) return rc;
    return launch_splitk_reduce_ch(static_cast<const float*>(workspace), splits, f_in, f_out, channels, dw, dbias, st);
}

}  // namespace kgcn

// Tuning hook (not part of the documented ABI): device buffer of [148][256] int64 clock stamps of the weight-gradient kernel
// (see DW_STAMP and tools/dw_timeline.py).
extern "C" void kgcn_debug_dw_times(long long* device_buffer) { kgcn::g_dbg_dw = device_buffer; }
