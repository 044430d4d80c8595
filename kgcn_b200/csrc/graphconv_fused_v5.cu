// Fused GraphConv layer for wide features (F = 128: config 5), "v5": the TRANSPOSED product with the weights resident in
// tensor memory.
//
//     y^T[n, r] = act( sum_k Wt[n, k] . Z[r, k]  +  sum_c rowsum(A_c)[r] . bias_c[n] ),   Z = [A_0.x | A_1.x | ..]   (kgcn/layers.py:105-116)
//
// At F = 128 the [W ; bias] hi / lo operand of the v4 kernel is 160 KB of shared memory -- it only fits as two output-column
// slices, each re-aggregating the tile on half the SMs.  Here the roles of the operands are swapped:
//   A operand  = W^T (hi | lo), M = f_out lanes x K columns of TENSOR MEMORY, loaded once per job and kept for all tiles;
//   B operand  = the aggregated tile Z (hi, lo) in shared memory, K-major SWIZZLE_128B, N = tile rows (64), double-buffered:
//                one thread per (row, 32-feature slab) gathers out of the TMA stage (8 LDS.128 per entry) and writes its 128-byte
//                line with 16 STS.128.  The 16-byte chunk order of both the gather and the store is permuted per lane by two
//                invertible GF(2) maps (t = T.(lane & 7) for the loads, t ^ (lane & 7) for the stores), so neither the loads
//                (every neighbour row starts at the same bank) nor the swizzled stores conflict -- and no un-rotation is needed;
//   D          = y^T in tensor memory, lanes = output feature, columns = tile rows, double-buffered; the epilogue reads 16 rows
//                per tcgen05.ld and each store instruction writes 32 consecutive features of one row = one 128-byte line.
// Shared memory is left for the TMA ring and Z; x and y cross HBM once, one CTA per SM computes all f_out columns.
// Same warp roles as v4 (TMA producer, 8 aggregation warps, MMA issuer, 8 epilogue warps), same multi-job chaining (a CTA
// owns the same graphs in every job; job boundary = CTA barrier + async-proxy fence), 3xTF32 (Whi.Zhi + Wlo.Zhi + Whi.Zlo).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "fused_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

constexpr int kMaxStages = 4;
constexpr int kAggWarps = 8, kEpiWarps = 8, kWarpEpi0 = 8, kWarpMma = 16, kWarpTma = 19;
constexpr int kBlock = 20 * 32;
constexpr int kRegsAgg = 128, kRegsEpi = 80, kRegsMisc = 56;
constexpr int kMaxJobs = 4;
constexpr int kMaxC = 4;
constexpr int kHeadTiles = 4;   // fused head: tiles per CTA whose column sums / d gathered are kept in shared memory at once

struct V5Params {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    const float* mul_src;   // y *= act'(mul_src) of mul_act (backward dx)
    float* zsave;           // optional [rows, K]: the aggregate Z = A . x in fp32 is also stored (a dx job's G = A^T . dU, see V4ChainJob)
    int64_t n_graphs;
    int C, N, f_in, f_out, act, mul_act, w_trans, f_valid;
    int K;                  // C * f_in
    int G, R;               // graphs per tile, MMA N = tile rows rounded up to 16
    int graphs_per_cta, n_slabs, slabs_per_ch, n_stages, cv_cap;
    uint32_t z_atom, z_half, z_buf;   // bytes: one 32-k atom (R x 128), hi -> lo, buffer -> buffer
    uint32_t off_z, off_deg, off_stage, stage_bytes, st_rp, st_col, st_val, smem_total;
    uint32_t tm_acc;        // first TMEM column of accumulator 0 (W^T hi at 0, lo at K)
    // fused readout head (last forward job of a training step, see graphconv_fused_v4.cu): GraphGather + Dense(n_labels) +
    // softmax cross-entropy on the epilogue's own tiles; the job then writes dU = dg (.) act'(H) instead of H
    int head, n_labels;
    const float* head_w;       // [f_out][n_labels]
    const float* head_b;       // [n_labels] or NULL
    const float* labels;       // [B][n_labels]
    const float* mask;         // [B] or NULL
    float inv_batch;
    float* logits;             // [B][n_labels] (may be NULL)
    float* prediction;         // [B][n_labels] (may be NULL)
    float* gathered;           // [B][f_out]    (may be NULL)
    float* head_partial;       // [grid][f_out * n_labels + 8]: dW_dense | db_dense (4) | cost_sum, correct_count, 0, 0
    uint32_t off_head;
};
struct V5Batch {
    int n_jobs;
    V5Params job[kMaxJobs];
};

template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
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
// CTA-wide barrier of all warp roles at a job boundary.  A named bar.sync: legal PTX from the roles' different program
// counters (compute-sanitizer's synccheck reports exactly that pattern as divergence, see tools/ubench/synccheck_probe.cu and
// profiles/r02b_sanitizer.txt; the mbarrier form the chained v4 kernel uses costs this kernel 70 bytes of spills in the
// aggregation loop: 47 vs 35 us for the two forward layers of config 5).
#define CTA_ROLE_BARRIER() asm volatile("bar.sync 2, %0;" ::"n"(kBlock) : "memory")
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]^T ; A: lane = M row, one tf32 per 32-bit column, 8 columns per K-step
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
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
__device__ __forceinline__ float fast_act_rt(float x, int act) {
    if (act == KGCN_ACT_RELU) return fmaxf(x, 0.0f);
    if (act == KGCN_ACT_SIGMOID) return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    if (act == KGCN_ACT_TANH) {
        const float t = ex2_approx(-2.8853900817779268f * fabsf(x));
        return copysignf((1.0f - t) * rcp_approx(1.0f + t), x);
    }
    return x;
}

struct Range {
    int64_t g_begin;
    int n_tiles, last_ng;
};
__device__ __forceinline__ Range cta_range(const V5Params& p) {
    Range t;
    t.g_begin = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
    const int64_t left = p.n_graphs - t.g_begin;
    const int n = static_cast<int>(left < p.graphs_per_cta ? (left > 0 ? left : 0) : p.graphs_per_cta);
    t.n_tiles = (n + p.G - 1) / p.G;
    t.last_ng = n - (t.n_tiles - 1) * p.G;
    return t;
}

__global__ void __maxnreg__(96) graphconv_fused_v5_kernel(const V5Batch b) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_zfull[2], bar_zempty[2], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_wready;   // W^T of the job is in tensor memory (written by the epilogue warps)
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_jobs = b.n_jobs;

    if (tid == 0) {
        mbar_init(&bar_wready, kEpiWarps);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], kAggWarps);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_zfull[i], kAggWarps);
            mbar_init(&bar_zempty[i], 1);
            mbar_init(&bar_tfull[i], 1);
            mbar_init(&bar_tempty[i], kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == kWarpMma) tmem_alloc(&tmem_slot, 512);
    pdl_wait();   // everything above overlaps the previous kernel's tail
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == kWarpTma) {
        // =============================== TMA producer ===============================
        reg_dec<kRegsMisc>();
        uint32_t ph_empty = 0;
        for (int j = 0; j < n_jobs; ++j) {
            const V5Params& p = b.job[j];
            if (j > 0) CTA_ROLE_BARRIER();   // the previous job's outputs are complete and visible to the async proxy
            if (lane == 0) {
                const int C = p.C, N = p.N, f_in = p.f_in, S = p.n_stages;
                const uint32_t pitch = static_cast<uint32_t>(f_in) * 4u;
                const Range tr = cta_range(p);
                int s = 0;
                for (int it = 0; it < tr.n_tiles; ++it) {
                    mbar_wait_relaxed(&bar_empty[s], ((ph_empty >> s) & 1u) ^ 1u);
                    ph_empty ^= 1u << s;
                    const int64_t g0 = tr.g_begin + static_cast<int64_t>(it) * p.G;
                    const int ng = (it == tr.n_tiles - 1) ? tr.last_ng : p.G;
                    const int64_t r0 = g0 * C * N;
                    const int rows_csr = ng * C * N;
                    unsigned char* st = gen + p.off_stage + static_cast<size_t>(s) * p.stage_bytes;
                    uint64_t* full = &bar_full[s];
                    const int64_t rp_lo = r0 & ~3ll;
                    const uint32_t rp_cnt = static_cast<uint32_t>((r0 + rows_csr + 1 - rp_lo + 3) & ~3ll);
                    const uint32_t x_bytes = static_cast<uint32_t>(ng) * static_cast<uint32_t>(N) * pitch;
                    mbar_expect_tx_only(full, x_bytes + 4u * rp_cnt);
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
                    if (++s == S) s = 0;
                }
            }
            __syncwarp();
        }
    } else if (warp < kAggWarps) {
        // =============================== aggregation warps ===============================
        reg_inc<kRegsAgg>();
        uint32_t ph_full = 0, ph_zempty = 0;
        // chunk permutations (GF(2)-linear in lane & 7): loads read chunk i ^ t, stores write position i ^ u, u = t ^ (lane & 7)
        const uint32_t s7 = static_cast<uint32_t>(lane) & 7u;
        const uint32_t b0 = s7 & 1u, b1 = (s7 >> 1) & 1u, b2 = (s7 >> 2) & 1u;
        const uint32_t t7 = b1 | (b2 << 1) | ((b0 ^ b1) << 2);
        const uint32_t u7 = t7 ^ s7;
        for (int j = 0; j < n_jobs; ++j) {
            const V5Params& p = b.job[j];
            if (j > 0) CTA_ROLE_BARRIER();
            const int N = p.N, C = p.C, f_in = p.f_in, S = p.n_stages, R = p.R;
            const uint32_t pitch = static_cast<uint32_t>(f_in) * 4u;
            const Range tr = cta_range(p);
            const int n_items_rb = (R + 31) >> 5;                  // 32-row blocks per tile
            const uint32_t r0_step = static_cast<uint32_t>(p.G * C * N);
            uint32_t r0_lo = static_cast<uint32_t>((tr.g_begin * C * N) & 3);
            int64_t zrow0 = tr.g_begin * N;   // global row index of the current tile's first row (zsave only)
            int s = 0, zi = 0;
            for (int it = 0; it < tr.n_tiles; ++it) {
                const bool last = it == tr.n_tiles - 1;
                const int rows = (last ? tr.last_ng : p.G) * N;
                const int rows_csr = last ? tr.last_ng * C * N : static_cast<int>(r0_step);
                const uint32_t st = base + p.off_stage + static_cast<uint32_t>(s) * p.stage_bytes;
                mbar_wait(&bar_full[s], (ph_full >> s) & 1u);
                ph_full ^= 1u << s;
                const uint32_t rp_addr = st + p.st_rp + 4u * (r0_lo & 3u);
                const int e_first = static_cast<int>(lds_u32(rp_addr));
                const int e_last = static_cast<int>(lds_u32(rp_addr + 4u * static_cast<uint32_t>(rows_csr)));
                const int e_lo = e_first & ~3;
                const bool staged = static_cast<uint32_t>((e_last - e_lo + 3) & ~3) <= static_cast<uint32_t>(p.cv_cap);
                const uint32_t col_addr = st + p.st_col - 4u * static_cast<uint32_t>(e_lo);
                const uint32_t val_addr = st + p.st_val - 4u * static_cast<uint32_t>(e_lo);
                mbar_wait(&bar_zempty[zi], ((ph_zempty >> zi) & 1u) ^ 1u);
                ph_zempty ^= 1u << zi;
                tc_fence_after_sync();
                const uint32_t zb = base + p.off_z + static_cast<uint32_t>(zi) * p.z_buf;
                const uint32_t degb = base + p.off_deg + static_cast<uint32_t>((it & 3) * kMaxC * R) * 4u;
                // warp-items: (slab, 32-row block); lanes = consecutive rows
                for (int wi = warp; wi < p.n_slabs * n_items_rb; wi += kAggWarps) {
                    const int slab = wi / n_items_rb, rb = wi - slab * n_items_rb;
                    const int c = slab / p.slabs_per_ch, fs = slab - c * p.slabs_per_ch;
                    const int r = rb * 32 + lane;
                    const bool valid = r < rows;
                    const int gl = r / N, node = r - gl * N;
                    int e = e_first, e_end = e_first;
                    if (valid) {
                        const uint32_t ra = rp_addr + 4u * static_cast<uint32_t>((gl * C + c) * N + node);
                        e = static_cast<int>(lds_u32(ra));
                        e_end = static_cast<int>(lds_u32(ra + 4u));
                    }
                    const uint32_t xbase = st + static_cast<uint32_t>(gl * N) * pitch + static_cast<uint32_t>(fs) * 128u + (t7 << 4);
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
                    float deg = 0.0f;
                    if (staged) {
                        uint32_t ce = col_addr + 4u * static_cast<uint32_t>(e), ve = val_addr + 4u * static_cast<uint32_t>(e);
                        const uint32_t cend = col_addr + 4u * static_cast<uint32_t>(e_end);
                        uint32_t cn = lds_u32(ce);     // one entry of look-ahead; reading one past the row is harmless (slack)
                        float vn = lds_f32(ve);
#pragma unroll 1
                        while (ce < cend) {
                            const uint32_t xa = xbase + cn * pitch;
                            const float v = vn;
                            ce += 4;
                            ve += 4;
                            cn = lds_u32(ce);
                            vn = lds_f32(ve);
                            float xv[8][4];
#pragma unroll
                            for (int i = 0; i < 8; ++i) lds_f<4>(xv[i], xa ^ (static_cast<uint32_t>(i) << 4));
                            deg += v;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = fmaf(v, xv[i][jj], acc[4 * i + jj]);
                        }
                    } else {   // unusually dense tile: the CSR slice did not fit the stage, entries come from global memory
#pragma unroll 1
                        for (; e < e_end; ++e) {
                            const uint32_t xa = xbase + static_cast<uint32_t>(__ldg(p.col + e)) * pitch;
                            const float v = __ldg(p.val + e);
                            deg += v;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float xv[4];
                                lds_f<4>(xv, xa ^ (static_cast<uint32_t>(i) << 4));
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) acc[4 * i + jj] = fmaf(v, xv[jj], acc[4 * i + jj]);
                            }
                        }
                    }
                    __syncwarp();
                    if (p.zsave != nullptr && valid) {
                        // the row's slab of Z in fp32: block i holds the 16-byte chunk i ^ t of the slab's 128-byte line, so the blocks
                        // (i, i + 1), i even, are the chunk pair starting at (i ^ t) & 6 -- in swapped order when t is odd.  Four
                        // 32-byte stores, one full sector each (every lane writes another row: 32 requests per store whatever its width)
                        const uint64_t zrow = reinterpret_cast<uint64_t>(p.zsave) +
                                              (static_cast<uint64_t>(zrow0 + r) * static_cast<uint64_t>(p.K) + static_cast<uint64_t>(slab * 32)) * 4u + ((t7 & 6u) << 4);
                        const bool odd = (t7 & 1u) != 0;
#pragma unroll
                        for (int i = 0; i < 8; i += 2) {
                            float lo4[4], hi4[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                lo4[jj] = odd ? acc[4 * (i + 1) + jj] : acc[4 * i + jj];
                                hi4[jj] = odd ? acc[4 * i + jj] : acc[4 * (i + 1) + jj];
                            }
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(zrow ^ static_cast<uint64_t>(i << 4)),
                                         "f"(lo4[0]), "f"(lo4[1]), "f"(lo4[2]), "f"(lo4[3]), "f"(hi4[0]), "f"(hi4[1]), "f"(hi4[2]), "f"(hi4[3]) : "memory");
                        }
                    }
                    if (r < R) {
                        // the row's 128-byte line of atom `slab`: block i holds chunk i ^ t -> position (i ^ t) ^ (r & 7) = i ^ u
                        const uint32_t line = zb + static_cast<uint32_t>(slab) * p.z_atom + static_cast<uint32_t>(r) * 128u + (u7 << 4);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float hi[4], lo[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                hi[jj] = __uint_as_float(__float_as_uint(acc[4 * i + jj]) & 0xFFFFE000u);   // truncation: lo = x - hi is exact
                                lo[jj] = acc[4 * i + jj] - hi[jj];
                            }
                            sts_f<4>(line ^ (static_cast<uint32_t>(i) << 4), hi);
                            sts_f<4>((line + p.z_half) ^ (static_cast<uint32_t>(i) << 4), lo);
                        }
                        if (fs == 0) {
                            const float dv[1] = {deg};
                            sts_f<1>(degb + 4u * static_cast<uint32_t>(c * R + r), dv);
                        }
                    }
                }
                fence_proxy_async_smem();   // Z is read by the tensor core through the async proxy
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bar_empty[s]);    // this warp is done reading the stage
                    mbar_arrive(&bar_zfull[zi]);
                }
                if (++s == S) s = 0;
                zi ^= 1;
                r0_lo += r0_step;
                zrow0 += static_cast<int64_t>(p.G) * N;
            }
        }
    } else if (warp >= kWarpMma) {
        reg_dec<kRegsMisc>();
        uint32_t ph_zfull = 0, ph_tempty = 0, ph_wready = 0;
        for (int j = 0; j < n_jobs; ++j) {
            const V5Params& p = b.job[j];
            if (j > 0) CTA_ROLE_BARRIER();
            if (warp != kWarpMma) continue;
            // =============================== MMA issuer ===============================
            if (lane == 0) mbar_wait_relaxed(&bar_wready, ph_wready & 1u);   // W^T is in tensor memory (epilogue warps); one polling lane, back-off
            __syncwarp();
            ph_wready ^= 1u;
            tc_fence_after_sync();
            const int K = p.K, R = p.R;
            const Range tr = cta_range(p);
            const uint32_t idesc = umma_idesc_tf32(128, R);
            const uint32_t whi = tmem, wlo = tmem + static_cast<uint32_t>(K);
            const uint32_t atom16 = p.z_atom >> 4, half16 = p.z_half >> 4;
            const int ks = K >> 3;
            int zi = 0, ai = 0;
            for (int it = 0; it < tr.n_tiles; ++it) {
                mbar_wait(&bar_zfull[zi], (ph_zfull >> zi) & 1u);
                ph_zfull ^= 1u << zi;
                mbar_wait(&bar_tempty[ai], ((ph_tempty >> ai) & 1u) ^ 1u);
                ph_tempty ^= 1u << ai;
                tc_fence_after_sync();
                __syncwarp();
                if (elect_one()) {
                    const uint32_t d = tmem + p.tm_acc + static_cast<uint32_t>(ai * R);
                    const uint64_t dzhi = umma_desc_sw128(base + p.off_z + static_cast<uint32_t>(zi) * p.z_buf);
                    uint32_t acc = 0;
#pragma unroll 1
                    for (int pass = 0; pass < 3; ++pass) {   // Whi.Zhi + Wlo.Zhi + Whi.Zlo
                        const uint32_t wa = (pass == 1) ? wlo : whi;
                        const uint64_t dz = (pass == 2) ? dzhi + half16 : dzhi;
#pragma unroll 4
                        for (int k8 = 0; k8 < ks; ++k8) {
                            umma_tf32_ts(d, wa + 8u * k8, dz + static_cast<uint64_t>((k8 >> 2) * atom16 + 2 * (k8 & 3)), idesc, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(&bar_zempty[zi]);
                    umma_commit(&bar_tfull[ai]);
                }
                __syncwarp();
                zi ^= 1;
                ai ^= 1;
            }
        }
    } else {
        // =============================== epilogue warps ===============================
        reg_dec<kRegsEpi>();
        uint32_t ph_tfull = 0;
        const int e = warp - kWarpEpi0;
        const int wq = e & 3, h = e >> 2;                 // TMEM lane quarter (32 output features), column half
        const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
        const int n = wq * 32 + lane;                     // this thread's output feature
        for (int j = 0; j < n_jobs; ++j) {
            const V5Params& p = b.job[j];
            if (j > 0) CTA_ROLE_BARRIER();
            const int N = p.N, C = p.C, f_in = p.f_in, f_out = p.f_out, K = p.K, R = p.R;
            // ---- W^T -> tensor memory (hi at column k, lo at K + k): lane = output feature, warp (q, h) takes every second 32-k chunk
            {
                const uint32_t wt = tmem + lane_sel;
                for (int kc = h; kc * 32 < K; kc += 2) {
                    // forward: Wt[n][k] = w[k][n] (k = c * f_in + f; coalesced over the lanes).  backward dx: the layer's
                    // w[c][n][f_o] with k = c * f_in + f_o (this job's f_in is the layer's f_out): contiguous along k
                    const int k0 = kc * 32, c0 = k0 / f_in;
                    const float* src = p.w_trans == 0 ? p.w + static_cast<size_t>(k0) * f_out + n
                                                      : p.w + (static_cast<size_t>(c0) * f_out + n) * f_in + (k0 - c0 * f_in);
                    const size_t stride = p.w_trans == 0 ? static_cast<size_t>(f_out) : 1;
                    uint32_t hi[32], lo[32];
                    float wv[32];
                    if (p.w_trans == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) wv[i] = n < f_out ? __ldg(src + static_cast<size_t>(i) * stride) : 0.0f;
                    } else {   // the thread's own row: 8 x 16-byte loads (rows are 16-byte aligned: f_in % 32 == 0)
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 q = n < f_out ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                            wv[4 * i] = q.x; wv[4 * i + 1] = q.y; wv[4 * i + 2] = q.z; wv[4 * i + 3] = q.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        hi[i] = __float_as_uint(tf32_hi(wv[i]));
                        lo[i] = __float_as_uint(wv[i] - __uint_as_float(hi[i]));
                    }
                    tmem_st32(wt + static_cast<uint32_t>(k0), hi);
                    tmem_st32(wt + static_cast<uint32_t>(K + k0), lo);
                }
                tmem_st_wait();
                tc_fence_before_sync();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_wready);
            float bias_c[kMaxC];
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
                bias_c[c] = (c < C && p.bias != nullptr && p.w_trans == 0 && n < f_out) ? __ldg(p.bias + static_cast<size_t>(c) * f_out + n) : 0.0f;
            const Range tr = cta_range(p);
            const bool col_ok = n < f_out;
            const bool keep = n < p.f_valid;
            if (p.head) {
                // ======== fused head.  Lane = output feature n (32 wq + lane), this warp's tile rows = columns [c_lo, c_hi) of the
                // accumulator.  Tiles are taken in groups (see gmax): (a) activation + per-graph column sums
                // for every tile of the group, (b) one warp per graph: gather, Dense, softmax cross-entropy, d logits, d gathered,
                // (c) dU = dg (.) act'(H) with H re-read from tensor memory.  G <= 4 graphs per tile (N >= 32). ========
                const int L = p.n_labels, Fs = f_out, G = p.G;
                const int te = tid - kWarpEpi0 * 32;
                const uint32_t hs_gsum = base + p.off_head;                                           // [kHeadTiles][2 halves][G][Fs]
                const uint32_t hs_dg = hs_gsum + static_cast<uint32_t>(2 * kHeadTiles * G * Fs) * 4u; // [kHeadTiles][G][Fs]
                const uint32_t hs_wd = hs_dg + static_cast<uint32_t>(kHeadTiles * G * Fs) * 4u;       // [Fs][L] + bias [4]
                const uint32_t hs_hp = hs_wd + static_cast<uint32_t>(Fs * L + 4) * 4u;                // [8 warps][Fs * L + 8]
                for (int i = te; i < Fs * L + 4; i += 256) {
                    const int bi = i - Fs * L;
                    const float wv[1] = {bi < 0 ? __ldg(p.head_w + i) : ((bi < L && p.head_b != nullptr) ? __ldg(p.head_b + bi) : 0.0f)};
                    sts_f<1>(hs_wd + 4u * static_cast<uint32_t>(i), wv);
                }
                const float z1[1] = {0.0f};
                for (int i = te; i < kEpiWarps * (Fs * L + 8); i += 256) sts_f<1>(hs_hp + 4u * static_cast<uint32_t>(i), z1);
                asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                const uint32_t my_hp = hs_hp + static_cast<uint32_t>(e * (Fs * L + 8)) * 4u;
                const int half = (R / 2 + 15) & ~15;
                const int c_lo = h * half, c_hi = min(R, c_lo + half);
                const int nq = Fs >> 5;                                   // 32-feature groups a lane of phase (b) owns (<= 4)
                // DEFERRED mode (a CTA's tiles <= kHeadTiles: the latency-bound case): (a) is the plain epilogue -- H goes to y, the
                // accumulator is handed back at once, so aggregation and MMAs of the next tiles overlap it -- plus the column sums of
                // ALL tiles; then ONE (b) for all graphs of the CTA and (c) as an element-wise sweep over the CTA's rows of y (read H
                // back from L2, write dU in place).  Otherwise groups of two tiles with H kept in tensor memory between (a) and (c).
                const bool deferred = tr.n_tiles <= kHeadTiles;
                const int gmax = deferred ? tr.n_tiles : 2;
                int ai = 0;
                for (int it0 = 0; it0 < tr.n_tiles; it0 += gmax) {
                    const int gsz = min(gmax, tr.n_tiles - it0);
                    const int64_t g0_grp = tr.g_begin + static_cast<int64_t>(it0) * G;
                    const int64_t left_cta = p.n_graphs - tr.g_begin;
                    const int n_cta = left_cta < p.graphs_per_cta ? static_cast<int>(left_cta) : p.graphs_per_cta;
                    const int ng_grp = min(gsz * G, n_cta - it0 * G);
                    float y_lane = 0.0f, m0 = 1.0f;     // label l (lane l) and mask of the graph this warp finishes in (b): requested early
                    if (e < ng_grp) {
                        if (lane < L) y_lane = __ldg(p.labels + (g0_grp + e) * L + lane);
                        if (p.mask) m0 = __ldg(p.mask + g0_grp + e);
                    }
                    // (a)
                    int aa = ai;
                    for (int u = 0; u < gsz; ++u) {
                        const int it = it0 + u;
                        const int rows = ((it == tr.n_tiles - 1) ? tr.last_ng : G) * N;
                        mbar_wait_relaxed(&bar_tfull[aa], (ph_tfull >> aa) & 1u);
                        ph_tfull ^= 1u << aa;
                        tc_fence_after_sync();
                        const uint32_t ta = tmem + lane_sel + p.tm_acc + static_cast<uint32_t>(aa * R);
                        const uint32_t degb = base + p.off_deg + static_cast<uint32_t>((it & 3) * kMaxC * R) * 4u;
                        const uint32_t gs_u = hs_gsum + static_cast<uint32_t>((u * 2 + h) * G * Fs) * 4u + 4u * static_cast<uint32_t>(n);
                        if (col_ok)
                            for (int g = 0; g < G; ++g) sts_f<1>(gs_u + 4u * static_cast<uint32_t>(g * Fs), z1);
#pragma unroll 1
                        for (int r0 = c_lo; r0 < c_hi; r0 += 16) {
                            float v[16];
                            tmem_ld16(ta + static_cast<uint32_t>(r0), v);
                            tmem_ld_wait();
                            tmem_ld_fence(v);
#pragma unroll
                            for (int c = 0; c < kMaxC; ++c)
                                if (c < C && bias_c[c] != 0.0f) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) v[i] = fmaf(lds_f32(degb + 4u * static_cast<uint32_t>(c * R + r0 + i)), bias_c[c], v[i]);
                                }
                            const int ga = r0 / N;                         // a 16-row chunk touches at most two graphs (N >= 32)
                            const int split = (ga + 1) * N - r0;           // first row of the chunk that belongs to graph ga + 1
                            float s0 = 0.0f, s1 = 0.0f;
                            float* hrow = p.y + ((tr.g_begin + static_cast<int64_t>(it) * G) * N + r0) * f_out + n;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float hv = (r0 + i < rows && keep) ? fast_act_rt(v[i], p.act) : 0.0f;
                                if (i < split) s0 += hv; else s1 += hv;
                                if (deferred && r0 + i < rows && col_ok) hrow[static_cast<size_t>(i) * f_out] = hv;   // H, replaced by dU in (c)
                            }
                            if (col_ok) {
                                const float a0[1] = {lds_f32(gs_u + 4u * static_cast<uint32_t>(ga * Fs)) + s0};
                                sts_f<1>(gs_u + 4u * static_cast<uint32_t>(ga * Fs), a0);
                                if (split < 16 && ga + 1 < G) {
                                    const float a1[1] = {lds_f32(gs_u + 4u * static_cast<uint32_t>((ga + 1) * Fs)) + s1};
                                    sts_f<1>(gs_u + 4u * static_cast<uint32_t>((ga + 1) * Fs), a1);
                                }
                            }
                        }
                        if (deferred) {   // the accumulator and the row sums of this tile are consumed
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_tempty[aa]);
                        }
                        aa ^= 1;
                    }
                    if (deferred) __threadfence_block();   // (c) reads H written by other threads of the CTA
                    asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                    // (b) one warp per graph of the group; a lane owns features lane + 32 q
                    for (int gi = e; gi < ng_grp; gi += kEpiWarps) {
                        const int u = gi / G, g = gi - u * G;
                        const int64_t bg = g0_grp + gi;
                        float yl = y_lane, m = m0;
                        if (gi != e) {
                            yl = lane < L ? __ldg(p.labels + bg * L + lane) : 0.0f;
                            m = p.mask ? __ldg(p.mask + bg) : 1.0f;
                        }
                        float gv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < nq) {
                                const uint32_t a = hs_gsum + 4u * static_cast<uint32_t>((u * 2 * G + g) * Fs + q * 32 + lane);
                                gv[q] = lds_f32(a) + lds_f32(a + 4u * static_cast<uint32_t>(G * Fs));
                                if (p.gathered != nullptr) p.gathered[bg * Fs + q * 32 + lane] = gv[q];
                            }
                        float zl = -3.0e38f;
#pragma unroll 1
                        for (int l = 0; l < L; ++l) {
                            float acc = 0.0f;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (q < nq) acc = fmaf(gv[q], lds_f32(hs_wd + 4u * static_cast<uint32_t>((q * 32 + lane) * L + l)), acc);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                            if (lane == l) zl = acc + lds_f32(hs_wd + 4u * static_cast<uint32_t>(Fs * L + l));
                        }
                        float zmax = zl, ymax = lane < L ? yl : -3.0e38f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                            ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
                        }
                        float ex = lane < L ? __expf(zl - zmax) : 0.0f, ysum = lane < L ? yl : 0.0f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            ex += __shfl_xor_sync(0xffffffffu, ex, o);
                            ysum += __shfl_xor_sync(0xffffffffu, ysum, o);
                        }
                        const float lse = __logf(ex) + zmax;
                        float cost = lane < L ? -yl * (zl - lse) : 0.0f;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
                        const int arg_p = __ffs(__ballot_sync(0xffffffffu, lane < L && zl == zmax)) - 1;
                        const int arg_y = __ffs(__ballot_sync(0xffffffffu, lane < L && yl == ymax)) - 1;
                        const float pr = lane < L ? __expf(zl - lse) : 0.0f;
                        const float dzl = m * p.inv_batch * (pr * ysum - yl);
                        if (lane < L) {
                            if (p.logits) p.logits[bg * L + lane] = zl;
                            if (p.prediction) p.prediction[bg * L + lane] = pr;
                            const float bacc[1] = {lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + lane)) + dzl};
                            sts_f<1>(my_hp + 4u * static_cast<uint32_t>(Fs * L + lane), bacc);
                        }
                        if (lane == 0) {
                            const float c2[2] = {lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + 4)) + m * cost,
                                                 lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + 5)) + m * (arg_p == arg_y ? 1.0f : 0.0f)};
                            sts_f<2>(my_hp + 4u * static_cast<uint32_t>(Fs * L + 4), c2);
                        }
                        float dg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                        for (int l = 0; l < L; ++l) {   // d gathered and this warp's share of dW_dense
                            const float dz = __shfl_sync(0xffffffffu, dzl, l);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (q < nq) {
                                    const uint32_t wi = 4u * static_cast<uint32_t>((q * 32 + lane) * L + l);
                                    dg[q] = fmaf(dz, lds_f32(hs_wd + wi), dg[q]);
                                    const float a0[1] = {fmaf(gv[q], dz, lds_f32(my_hp + wi))};
                                    sts_f<1>(my_hp + wi, a0);
                                }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q < nq) {
                                const float o0[1] = {dg[q]};
                                sts_f<1>(hs_dg + 4u * static_cast<uint32_t>((u * G + g) * Fs + q * 32 + lane), o0);
                            }
                        __syncwarp();
                    }
                    asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                    // (c)
                    if (deferred) {   // element-wise sweep over the CTA's rows of y: 16-byte accesses, 8 independent loads per thread in flight
                        ai = aa;
                        const int n_cta_rows = n_cta * N;
                        const int f4 = Fs >> 2;
                        float* y0 = p.y + tr.g_begin * N * static_cast<int64_t>(f_out);
                        for (int i0 = te; i0 < n_cta_rows * f4; i0 += 8 * 256) {
                            float4 hv[8];
#pragma unroll
                            for (int k8 = 0; k8 < 8; ++k8) {
                                const int idx = i0 + k8 * 256;
                                hv[k8] = idx < n_cta_rows * f4 ? *reinterpret_cast<const float4*>(y0 + static_cast<size_t>(idx) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                            for (int k8 = 0; k8 < 8; ++k8) {
                                const int idx = i0 + k8 * 256;
                                if (idx < n_cta_rows * f4) {
                                    const int row = idx / f4, c4 = (idx - row * f4) << 2;
                                    float d[4];
                                    lds_f<4>(d, hs_dg + 4u * static_cast<uint32_t>((row / N) * Fs + c4));   // tile-major = graph-major: [tile][g][Fs]
                                    const float o[4] = {d[0] * act_grad_from_output(hv[k8].x, p.act), d[1] * act_grad_from_output(hv[k8].y, p.act),
                                                        d[2] * act_grad_from_output(hv[k8].z, p.act), d[3] * act_grad_from_output(hv[k8].w, p.act)};
                                    *reinterpret_cast<float4*>(y0 + static_cast<size_t>(idx) * 4) = make_float4(o[0], o[1], o[2], o[3]);
                                }
                            }
                        }
                    }
                    for (int u = 0; u < (deferred ? 0 : gsz); ++u) {
                        const int it = it0 + u;
                        const int rows = ((it == tr.n_tiles - 1) ? tr.last_ng : G) * N;
                        const int64_t row_base = (tr.g_begin + static_cast<int64_t>(it) * G) * N;
                        const uint32_t ta = tmem + lane_sel + p.tm_acc + static_cast<uint32_t>(ai * R);
                        const uint32_t degb = base + p.off_deg + static_cast<uint32_t>((it & 3) * kMaxC * R) * 4u;
                        const uint32_t dg_u = hs_dg + 4u * static_cast<uint32_t>(u * G * Fs + n);
#pragma unroll 1
                        for (int r0 = c_lo; r0 < c_hi; r0 += 16) {
                            float v[16];
                            tmem_ld16(ta + static_cast<uint32_t>(r0), v);
                            float* yrow = p.y + (row_base + r0) * f_out + n;
                            const int ga = r0 / N;
                            const int split = (ga + 1) * N - r0;
                            const float d0 = col_ok ? lds_f32(dg_u + 4u * static_cast<uint32_t>(ga * Fs)) : 0.0f;
                            const float d1 = (col_ok && split < 16 && ga + 1 < G) ? lds_f32(dg_u + 4u * static_cast<uint32_t>((ga + 1) * Fs)) : 0.0f;
                            tmem_ld_wait();
                            tmem_ld_fence(v);
#pragma unroll
                            for (int c = 0; c < kMaxC; ++c)
                                if (c < C && bias_c[c] != 0.0f) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) v[i] = fmaf(lds_f32(degb + 4u * static_cast<uint32_t>(c * R + r0 + i)), bias_c[c], v[i]);
                                }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float hv = keep ? fast_act_rt(v[i], p.act) : 0.0f;
                                const float du = (i < split ? d0 : d1) * act_grad_from_output(hv, p.act);
                                if (r0 + i < rows && col_ok) yrow[static_cast<size_t>(i) * f_out] = keep ? du : 0.0f;
                            }
                        }
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_tempty[ai]);   // accumulator and row sums of this tile are consumed
                        ai ^= 1;
                    }
                }
                // per-CTA partial of the head's parameter gradients and statistics: the 8 warps' sums, added in warp order
                asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                float* hp_out = p.head_partial + static_cast<size_t>(blockIdx.x) * (Fs * L + 8);
                for (int i = te; i < Fs * L + 8; i += 256) {
                    float acc = 0.0f;
                    for (int w8 = 0; w8 < kEpiWarps; ++w8) acc += lds_f32(hs_hp + 4u * static_cast<uint32_t>(w8 * (Fs * L + 8) + i));
                    hp_out[i] = acc;
                }
            } else {
                int ai = 0;
                for (int it = 0; it < tr.n_tiles; ++it) {
                    const int rows = ((it == tr.n_tiles - 1) ? tr.last_ng : p.G) * N;
                    const int64_t row_base = (tr.g_begin + static_cast<int64_t>(it) * p.G) * N;
                    mbar_wait_relaxed(&bar_tfull[ai], (ph_tfull >> ai) & 1u);
                    ph_tfull ^= 1u << ai;
                    tc_fence_after_sync();
                    const uint32_t ta = tmem + lane_sel + p.tm_acc + static_cast<uint32_t>(ai * R);
                    // row sums of tile `it`: four buffers -- the writer of tile it + 4 runs only after this tile's accumulator has
                    // been handed back (below, after the last read of degb)
                    const uint32_t degb = base + p.off_deg + static_cast<uint32_t>((it & 3) * kMaxC * R) * 4u;
                    const int half = (R / 2 + 15) & ~15;           // columns (tile rows) of this warp: [h * half, min(R, (h + 1) * half))
                    const int c_lo = h * half, c_hi = min(R, c_lo + half);
    #pragma unroll 1
                    for (int r0 = c_lo; r0 < c_hi; r0 += 16) {
                        float v[16];
                        tmem_ld16(ta + static_cast<uint32_t>(r0), v);
                        float* yrow = p.y + (row_base + r0) * f_out + n;
                        tmem_ld_wait();
                        tmem_ld_fence(v);
    #pragma unroll
                        for (int c = 0; c < kMaxC; ++c)
                            if (c < C && bias_c[c] != 0.0f) {
    #pragma unroll
                                for (int i = 0; i < 16; ++i) v[i] = fmaf(lds_f32(degb + 4u * static_cast<uint32_t>(c * R + r0 + i)), bias_c[c], v[i]);
                            }
                        if (p.act != KGCN_ACT_NONE) {
    #pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fast_act_rt(v[i], p.act);
                        }
    #pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (r0 + i < rows && col_ok) yrow[static_cast<size_t>(i) * f_out] = keep ? v[i] : 0.0f;
                    }
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tempty[ai]);   // accumulator and row sums of this tile are consumed
                    ai ^= 1;
                }
                if (p.mul_src != nullptr) {
                    // backward dx: y *= act'(mul_src) as an element-wise sweep over the CTA's rows once all tiles are stored (16-byte
                    // accesses, 2 x 4 independent loads per thread in flight) -- inside the tile loop the 16 strided loads per chunk
                    // were a round trip per chunk on the epilogue's critical path
                    __threadfence_block();
                    asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                    const int te = tid - kWarpEpi0 * 32;
                    const int64_t left_cta = p.n_graphs - tr.g_begin;
                    const int n_cta = left_cta < p.graphs_per_cta ? static_cast<int>(left_cta) : p.graphs_per_cta;
                    const int total4 = n_cta * N * (f_out >> 2);
                    float* y0 = p.y + tr.g_begin * N * static_cast<int64_t>(f_out);
                    const float* m0p = p.mul_src + tr.g_begin * N * static_cast<int64_t>(f_out);
                    for (int i0 = te; i0 < total4; i0 += 4 * 256) {
                        float4 yv[4], mv4[4];
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int idx = i0 + k4 * 256;
                            const bool ok = idx < total4;
                            yv[k4] = ok ? *reinterpret_cast<const float4*>(y0 + static_cast<size_t>(idx) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            mv4[k4] = ok ? __ldg(reinterpret_cast<const float4*>(m0p) + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const int idx = i0 + k4 * 256;
                            if (idx < total4)
                                *reinterpret_cast<float4*>(y0 + static_cast<size_t>(idx) * 4) =
                                    make_float4(yv[k4].x * act_grad_from_output(mv4[k4].x, p.mul_act), yv[k4].y * act_grad_from_output(mv4[k4].y, p.mul_act),
                                                yv[k4].z * act_grad_from_output(mv4[k4].z, p.mul_act), yv[k4].w * act_grad_from_output(mv4[k4].w, p.mul_act));
                        }
                    }
                }
            }
            // this CTA's outputs of the job: visible to the next job's TMA loads (async proxy) after the role barrier
            __threadfence();
            fence_proxy_async_all();
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem, 512);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
constexpr int kSmemMax = 227 * 1024 - 1024;

bool plan_v5_try(V5Params& p, int64_t n_graphs, int C, int N, int f_in, int f_out, int G, int head_labels = 0) {
    p.C = C; p.N = N; p.f_in = f_in; p.f_out = f_out; p.n_graphs = n_graphs;
    p.K = C * f_in;
    p.n_slabs = p.K / 32;
    p.slabs_per_ch = f_in / 32;
    const int64_t grid0 = std::min<int64_t>(kNumSMs, n_graphs);
    const int64_t gpc = ceil_div<int64_t>(n_graphs, grid0);
    if (gpc > (1 << 24)) return false;
    p.graphs_per_cta = static_cast<int>(gpc);
    p.G = std::min(G, p.graphs_per_cta);
    const uint32_t rows_max = static_cast<uint32_t>(p.G) * N;
    p.R = static_cast<int>(up(rows_max, 16));
    if (p.R > 128) return false;
    if (2 * p.K + 2 * p.R > 512) return false;          // W^T hi | lo + two accumulators of R columns
    p.tm_acc = static_cast<uint32_t>(2 * p.K);
    p.z_atom = static_cast<uint32_t>(p.R) * 128u;
    p.z_half = static_cast<uint32_t>(p.n_slabs) * p.z_atom;
    p.z_buf = 2u * p.z_half;
    uint32_t off = 0;
    p.off_z = off; off += 2u * p.z_buf;
    p.off_deg = off; off += up(4u * kMaxC * p.R * 4u, 128);
    p.off_head = off;
    p.head = 0;
    if (head_labels > 0) {   // gsum [kHeadTiles][2][G][F] + dg [kHeadTiles][G][F] + Dense weights + 8 per-warp accumulator blocks
        if (p.G > 4 || N < 32 || f_out % 32 != 0 || head_labels > 4) return false;
        off += up(static_cast<uint32_t>(3 * kHeadTiles * p.G * f_out + f_out * head_labels + 4 + kEpiWarps * (f_out * head_labels + 8)) * 4u, 128);
    }
    p.off_stage = off;
    p.cv_cap = static_cast<int>(up(std::max<uint32_t>(256, 6 * rows_max * C), 4));
    p.st_rp = up(rows_max * f_in * 4u, 128);
    p.st_col = p.st_rp + up((rows_max * C + 8) * 4u, 16);
    p.st_val = p.st_col + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u;
    p.stage_bytes = up(p.st_val + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u, 128);
    p.n_stages = 0;
    for (int st = kMaxStages; st >= 2; --st)
        if (off + st * p.stage_bytes + 1024 <= static_cast<uint32_t>(kSmemMax)) { p.n_stages = st; break; }
    if (p.n_stages == 0) return false;
    p.smem_total = off + p.n_stages * p.stage_bytes + 1024;
    return true;
}

bool plan_v5(V5Params& p, int64_t n_graphs, int C, int N, int f_in, int f_out, int head_labels = 0) {
    if (n_graphs <= 0 || N > 128 || N < 1 || f_in % 32 != 0 || f_out < 1 || f_out > 128 || C < 1 || C > kMaxC) return false;
    for (int G = std::max(1, 128 / N); G >= 1; G = (G > 1 ? G / 2 : 0))
        if (plan_v5_try(p, n_graphs, C, N, f_in, f_out, G, head_labels)) return true;
    return false;
}

bool v5_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_FUSED_V5");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

}  // namespace

bool fused_v5_plannable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    V5Params p{};
    return v5_enabled() && plan_v5(p, n_graphs, channels, n_nodes, f_in, f_out);
}

// grid of a chained launch whose last forward layer (f_in -> f_out) carries the fused head; 0 = not supported
int fused_v5_head_grid(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int n_labels) {
    V5Params p{};
    if (!v5_enabled() || n_labels < 1 || !plan_v5(p, n_graphs, channels, n_nodes, f_in, f_out, n_labels)) return 0;
    return static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta));
}

int launch_graphconv_fused_v5_chain(const V4ChainJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, cudaStream_t st) {
    KGCN_REQUIRE(n_jobs >= 1 && n_jobs <= kMaxJobs, KGCN_ERR_BAD_SHAPE, "fused GraphConv v5 chain: 1..%d jobs", kMaxJobs);
    V5Batch b{};
    b.n_jobs = n_jobs;
    uint32_t smem = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const V4ChainJob& j = jobs[k];
        V5Params& p = b.job[k];
        KGCN_REQUIRE(plan_v5(p, n_graphs, channels, n_nodes, j.f_in, j.f_out, j.head != nullptr ? j.head->n_labels : 0), KGCN_ERR_UNSUPPORTED,
                     "fused GraphConv v5: job %d (%d -> %d) unsupported", k, j.f_in, j.f_out);
        KGCN_REQUIRE(p.graphs_per_cta == b.job[0].graphs_per_cta, KGCN_ERR_UNSUPPORTED, "fused GraphConv v5 chain: graph ranges differ");
        p.rowptr = j.rowptr; p.col = j.col; p.val = j.val; p.x = j.x; p.y = j.y;
        p.w = j.w;
        p.bias = j.w_transposed ? nullptr : j.bias;
        p.w_trans = j.w_transposed ? 1 : 0;
        p.act = j.act;
        KGCN_REQUIRE(j.mul_src == nullptr || j.f_out % 4 == 0, KGCN_ERR_UNSUPPORTED, "fused GraphConv v5: act' epilogue needs f_out %% 4 == 0");
        p.mul_src = j.mul_src;
        p.mul_act = j.mul_act;
        KGCN_REQUIRE(j.zsave == nullptr || ((reinterpret_cast<uintptr_t>(j.zsave) & 127u) == 0 && (p.K & 31) == 0), KGCN_ERR_UNSUPPORTED,
                     "fused GraphConv v5: the aggregate can only be stored into a 128-byte aligned buffer");
        p.zsave = j.zsave;
        p.f_valid = (j.f_out_valid > 0 && j.f_out_valid < j.f_out) ? j.f_out_valid : j.f_out;
        if (j.head != nullptr) {
            const V4Head& hd = *j.head;
            KGCN_REQUIRE(j.mul_src == nullptr && !j.w_transposed && hd.w && hd.labels && hd.partial && hd.n_labels >= 1 && hd.n_labels <= 4,
                         KGCN_ERR_BAD_SHAPE, "fused GraphConv v5: bad head (1..4 labels)");
            p.head = 1;
            p.n_labels = hd.n_labels;
            p.head_w = hd.w; p.head_b = hd.b; p.labels = hd.labels; p.mask = hd.mask; p.inv_batch = hd.inv_batch;
            p.logits = hd.logits; p.prediction = hd.prediction; p.gathered = hd.gathered; p.head_partial = hd.partial;
        }
        smem = std::max(smem, p.smem_total);
    }
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_graphs, b.job[0].graphs_per_cta));
    KGCN_CUDA_OK(cudaFuncSetAttribute(graphconv_fused_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    launch_pdl(graphconv_fused_v5_kernel, grid, kBlock, smem, st, b);
    KGCN_LAUNCH_OK("graphconv_fused_v5_kernel");
    return KGCN_OK;
}

}  // namespace kgcn
