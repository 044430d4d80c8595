// Fused GraphConv layer for sm_100a, warp-specialised pipeline ("v4"):
//
//     y[g] = act( sum_c (A[g,c] . x[g]) . W_c  +  rowsum(A[g,c]) (x) bias_c )        (kgcn/layers.py:105-116,
//                                                          re-associated aggregate-first)
//
// One persistent CTA per SM owns a contiguous range of graphs and walks it in tiles of G whole graphs
// (G * N <= 128 rows = the 128 TMEM lanes).  Four roles run concurrently on DIFFERENT tiles:
//
//   producer warp    TMA bulk copies of the tile's feature rows and CSR slices into a ring of smem stages
//   aggregation      one THREAD per (tile row, 32-feature slab): walks the row's CSR entries and gathers the
//   warps            neighbours' 128-byte slab segments out of the stage with 8 LDS.128 per entry.  The
//                    16-byte chunk order is rotated per lane (chunk i ^ (lane & 7)) so that the 8 lanes of an
//                    LDS phase always hit 8 different bank groups although every neighbour row starts at the
//                    same bank.  The row of Z = A.X is split into tf32 hi / lo and written with tcgen05.st
//                    straight into TENSOR MEMORY (lane = row), next to the row sum of the adjacency values.
//   MMA warp         one thread issues tcgen05.mma kind::tf32 with the A operand in TMEM and B = [W ; bias]
//                    in shared memory (K-major SWIZZLE_128B): 3xTF32 (Zhi.Whi + Zlo.Whi + Zhi.Wlo) into ONE
//                    fp32 accumulator in TMEM.  The bias rides in the GEMM: Z carries rowsum(A_c) in column
//                    K + c and B carries bias_c in row K + c, so the epilogue is activation only.
//   epilogue warps   tcgen05.ld the accumulator (double-buffered), activation, swizzled per-warp staging
//                    tile, coalesced 16-byte global stores (four full 128-byte lines per store instruction).
// 20 warps = 5 warpgroups (2 aggregation, 2 epilogue, 1 MMA + TMA); setmaxnreg moves registers from the
// epilogue / MMA / TMA warpgroups to the aggregation threads (32 accumulators + 32 values in flight each).
//
// Z (in TMEM) and the accumulator are double-buffered, so aggregation of tile i+1, the contraction of
// tile i and the epilogue of tile i-1 overlap; shared memory is left for a deep (3-4 tile) TMA ring.
// HBM traffic per layer = x once + y once + CSR once + W once per CTA: the algorithmic minimum.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "fused_common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace kgcn {
namespace {

constexpr int kV4MaxStages = 4;

struct V4Params {
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int64_t n_graphs;
    int C, N, f_in, f_out, act;
    int G;                // graphs per tile
    int graphs_per_cta;   // contiguous graph range per CTA
    int K, Kp, Np;        // K = C * f_in; Kp = K + 8 (row-sum columns); Np = f_out padded to 16
    int n_slabs, slabs_per_ch;
    int n_stages, cv_cap;
    int zbufs;            // Z buffers in TMEM (2 when they fit, else 1)
    int abufs;            // accumulators in TMEM (2 when they fit, else 1)
    int w_trans;          // 0: w[c][k][n] (+ bias) -- the forward layer;  1: w[c][n][k], no bias -- dx = G . W^T of the backward
    int w_ld, w_cstride;  // row pitch and channel stride of w (floats)
    int y_ld;             // row pitch of y (floats)
    int n_split;          // output-column slices (blockIdx.y): slice s computes columns [s * f_out, (s + 1) * f_out) of y_ld
    int C_csr, c_begin;   // the CSR holds C_csr channels per graph; this launch contracts channels [c_begin, c_begin + C)
    int acc_in;           // add the existing y before the activation (second launch of a layer split over channel groups)
    int f_valid;          // output columns >= f_valid (of this slice) are written as exact zeros: feature padding stays inert
    // fused readout head (last forward job of a training step): GraphGather + Dense(n_labels) + softmax cross-entropy on the
    // epilogue's own tiles (example_model/model.py:56-69); the job then writes dU = dg (.) act'(H) instead of H
    int head, n_labels;
    const float* head_w;       // [f_out][n_labels] dense/kernel (padded rows are zero)
    const float* head_b;       // [n_labels] or NULL
    const float* labels;       // [B][n_labels]
    const float* mask;         // [B] or NULL
    float inv_batch;
    float* logits;             // [B][n_labels] (may be NULL)
    float* prediction;         // [B][n_labels] (may be NULL)
    float* gathered;           // [B][f_out]    (may be NULL)
    float* head_partial;       // [grid][f_out * n_labels + 8]: dW_dense | db_dense (4) | cost_sum, correct_count, 0, 0
    uint32_t off_head;
    const float* mul_src; // optional [rows, y_ld]: the output is multiplied by act'(mul_src) of activation mul_act (backward:
    int mul_act;          // dx . act'(x) = the dU of the layer below, so no separate activation-gradient pass is needed)
    uint32_t off_whi, off_wlo, off_ystage, off_stage, stage_bytes, st_rp, st_col, st_val, smem_total;
    uint32_t w_atom;      // bytes between K atoms of the B operand
    uint32_t tm_z;        // first TMEM column of Z buffer 0 (accumulators at columns 0 and Np)
    int soft;             // chained job with the same plan as its predecessor: no CTA barrier, tile-wise hand-over
    int soft_next;        // the next job of the chain is soft: publish every finished tile (fence + counter)
    int wbufs, wsel;      // B operand buffers (chain: 2 = the next soft job's [W ; bias] is staged during this job), buffer of this job
    int prestaged;        // this job's B operand was staged by the stager warps during the previous job
    int hard_ord;         // jobs with a hard boundary: 1, 2, .. in launch order (phase of the role barrier)
    int x_prev;           // soft job whose x IS the previous job's y (same tiles): when a CTA's tiles fit the stage ring, the previous
                          // job's epilogue writes each output tile straight into this job's stage slot (on-chip hand-over)
    uint32_t w_pair;      // bytes of one (hi, lo) B operand buffer
    float* zsave;         // chained kernel, optional [rows, K]: the aggregate Z = A . x (fp32, before the tf32 split) is also stored
    long long* dbg;
};

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
    if (ACT == KGCN_ACT_SIGMOID) return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
    if (ACT == KGCN_ACT_TANH) {
        const float t = ex2_approx(-2.8853900817779268f * fabsf(x));
        return copysignf((1.0f - t) * rcp_approx(1.0f + t), x);
    }
    return x;
}

// D[tmem] (+)= A[tmem] . B[smem]^T ; A: lane = row, one tf32 per 32-bit column, 8 columns per K-step
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}

// tuning aid: per-role phase timers (lane 0 of the first warp of each role), enabled by kgcn_debug_v4_times
struct PhaseTimer {
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    long long last = 0;
    bool on;
    __device__ __forceinline__ explicit PhaseTimer(bool enabled) : on(enabled) { if (on) last = clock64(); }
    __device__ __forceinline__ void mark(int ph) {
        if (on) {
            const long long now = clock64();
            acc[ph] += now - last;
            last = now;
        }
    }
};

// ---- tile bookkeeping shared by every role ----
struct TileRange {
    int64_t g_begin;
    int n_graphs_cta, n_tiles;
};
__device__ __forceinline__ TileRange cta_range(const V4Params& p) {
    TileRange t;
    t.g_begin = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
    const int64_t left = p.n_graphs - t.g_begin;
    t.n_graphs_cta = static_cast<int>(left < p.graphs_per_cta ? (left > 0 ? left : 0) : p.graphs_per_cta);
    t.n_tiles = (t.n_graphs_cta + p.G - 1) / p.G;
    return t;
}

// One (row, slab) of Z = A.X: gather + FFMA over the row's entries; acc block i holds chunk (i ^ s7).
template <bool STAGED>
__device__ __forceinline__ void gather_row(float (&acc)[32], float& deg, int rs, int re, uint32_t col_addr, uint32_t val_addr,
                                           const int32_t* gcol, const float* gval, uint32_t xbase, uint32_t pitch) {
    if (STAGED) {
        uint32_t ce = col_addr + 4u * static_cast<uint32_t>(rs), ve = val_addr + 4u * static_cast<uint32_t>(rs);
        const uint32_t cend = col_addr + 4u * static_cast<uint32_t>(re);
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
                for (int j = 0; j < 4; ++j) acc[4 * i + j] = fmaf(v, xv[i][j], acc[4 * i + j]);
        }
    } else {   // unusually dense tile: the CSR slice did not fit the stage, entries come from global memory
        for (int e = rs; e < re; ++e) {
            const uint32_t xa = xbase + static_cast<uint32_t>(__ldg(gcol + e)) * pitch;
            const float v = __ldg(gval + e);
            float xv[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) lds_f<4>(xv[i], xa ^ (static_cast<uint32_t>(i) << 4));
            deg += v;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[4 * i + j] = fmaf(v, xv[i][j], acc[4 * i + j]);
        }
    }
}

template <int ACT>
__device__ __forceinline__ void act16(float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fast_act<ACT>(v[i]);
}
__device__ __forceinline__ void act16_rt(float (&v)[16], int act) {
    switch (act) {
        case KGCN_ACT_RELU: act16<KGCN_ACT_RELU>(v); break;
        case KGCN_ACT_SIGMOID: act16<KGCN_ACT_SIGMOID>(v); break;
        case KGCN_ACT_TANH: act16<KGCN_ACT_TANH>(v); break;
        default: break;
    }
}

// t[i] *= act'(y[i]) with the activation switch outside the element loop (compact code: the epilogue paths run once or
// twice per job at small batches, so every extra instruction is an instruction-cache miss, not an issue slot)
template <int ACT>
__device__ __forceinline__ void mul_act_grad4(float (&t)[4], const float (&y)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (ACT == KGCN_ACT_RELU) t[i] = y[i] > 0.0f ? t[i] : 0.0f;
        else if (ACT == KGCN_ACT_SIGMOID) t[i] *= y[i] * (1.0f - y[i]);
        else if (ACT == KGCN_ACT_TANH) t[i] *= 1.0f - y[i] * y[i];
    }
}
__device__ __forceinline__ void mul_act_grad4_rt(float (&t)[4], const float (&y)[4], int act) {
    switch (act) {
        case KGCN_ACT_RELU: mul_act_grad4<KGCN_ACT_RELU>(t, y); break;
        case KGCN_ACT_SIGMOID: mul_act_grad4<KGCN_ACT_SIGMOID>(t, y); break;
        case KGCN_ACT_TANH: mul_act_grad4<KGCN_ACT_TANH>(t, y); break;
        default: break;
    }
}

// Chained kernel: the activation of a network is a template parameter (ACT = the one non-trivial activation of all jobs,
// -1 = mixed: runtime switch), so only ONE activation's code is in the instruction stream -- at the chain's sizes most
// paths run once or twice per launch and instruction fetch, not issue, bounds them.
template <int ACT>
__device__ __forceinline__ void act16_sel(float (&v)[16], int act) {
    if (ACT == -1) act16_rt(v, act);
    else if (ACT != KGCN_ACT_NONE && act != KGCN_ACT_NONE) act16<ACT>(v);
}
template <int ACT>
__device__ __forceinline__ void mul_act_grad4_sel(float (&t)[4], const float (&y)[4], int act) {
    if (ACT == -1) mul_act_grad4_rt(t, y, act);
    else if (ACT != KGCN_ACT_NONE && act != KGCN_ACT_NONE) mul_act_grad4<ACT>(t, y);
}
template <int ACT>
__device__ __forceinline__ float act1_sel(float x, int act) {
    if (ACT == -1) {
        switch (act) {
            case KGCN_ACT_RELU: return fast_act<KGCN_ACT_RELU>(x);
            case KGCN_ACT_SIGMOID: return fast_act<KGCN_ACT_SIGMOID>(x);
            case KGCN_ACT_TANH: return fast_act<KGCN_ACT_TANH>(x);
            default: return x;
        }
    }
    return (ACT != KGCN_ACT_NONE && act != KGCN_ACT_NONE) ? fast_act<(ACT < 0 ? 0 : ACT)>(x) : x;
}

// KS K-steps of one 3xTF32 pass, unrolled so every operand address is a constant add
template <int KS>
__device__ __forceinline__ void issue_tile(uint32_t d, uint32_t zhi, uint32_t zlo, uint64_t dwhi, uint64_t dwlo, uint32_t idesc,
                                           uint32_t w_atom16) {
    uint32_t acc = 0;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t za = (pass == 1) ? zlo : zhi;
        const uint64_t db = (pass == 2) ? dwlo : dwhi;
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            umma_tf32_ts(d, za + 8u * j, db + static_cast<uint64_t>((j >> 2) * w_atom16 + 2 * (j & 3)), idesc, acc);
            acc = 1;
        }
    }
}
__device__ __noinline__ void issue_tile_loop(uint32_t d, uint32_t zhi, uint32_t zlo, uint64_t dwhi, uint64_t dwlo, uint32_t idesc,
                                             uint32_t w_atom16, int ks) {
    uint32_t acc = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t za = (pass == 1) ? zlo : zhi;
        const uint64_t db = (pass == 2) ? dwlo : dwhi;
        for (int j = 0; j < ks; ++j) {
            umma_tf32_ts(d, za + 8u * j, db + static_cast<uint64_t>((j >> 2) * w_atom16 + 2 * (j & 3)), idesc, acc);
            acc = 1;
        }
    }
}

// Warp roles.  20 warps = 5 warpgroups; registers are re-balanced per warpgroup with setmaxnreg (the
// gather threads hold a 32-float accumulator block plus 32 floats in flight).
constexpr int kAggWarps = 8;      // warps 0-7   (warpgroups 0-1): 4 TMEM lane quarters x 2 slab phases
constexpr int kEpiWarps = 8;      // warps 8-15  (warpgroups 2-3): 4 TMEM lane quarters x 2 column phases
constexpr int kWarpEpi0 = 8;
constexpr int kWarpMma = 16;      // warpgroup 4: MMA issuer, two B-operand stager warps (chained kernel; idle in the single-layer one), TMA producer
constexpr int kStagers = 64;      // threads of warps 17, 18
constexpr int kWarpTma = 19;
constexpr int kBlock = 20 * 32;
constexpr int kRegsAgg = 128, kRegsEpi = 80, kRegsMisc = 64;   // 256 x 128 + 256 x 80 + 128 x 64 = the 640 x 96 of the launch

template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// ring position + mbarrier phase bit of a role walking a ring of `n` slots
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
// waits that are not latency-critical back off, so the polling does not take issue slots from the working warps
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t* bar, uint32_t bytes) {   // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// EPI: 0 = activation epilogue; 1 = multiply by act'(mul_src) (backward dx -> dU of the layer below);
//      2 = add the existing y, then activation (second launch of a layer split over channel groups)
template <int EPI>
__global__ void __maxnreg__(96) graphconv_fused_v4_kernel(const V4Params p) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kV4MaxStages], bar_empty[kV4MaxStages];
    __shared__ __align__(8) uint64_t bar_zfull[2], bar_zempty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = p.C, N = p.N, f_in = p.f_in, f_out = p.f_out, K = p.K, Kp = p.Kp, Np = p.Np, S = p.n_stages;
    const uint32_t pitch = static_cast<uint32_t>(f_in) * 4u;
    const TileRange tr = cta_range(p);
    const int n_tiles = tr.n_tiles;
    const int last_ng = tr.n_graphs_cta - (n_tiles - 1) * p.G;   // graphs in the CTA's last tile
    const long long t_start = p.dbg ? clock64() : 0;
    long long* dbg = p.dbg ? p.dbg + static_cast<size_t>(blockIdx.x) * 16 : nullptr;

    if (tid == 0) {
        for (int i = 0; i < kV4MaxStages; ++i) {
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
    if ((f_out & 15) != 0) {   // B rows n >= f_out (up to Np) are read by the MMAs: they must be exact zeros.  Every other byte
        // the MMAs read (n < f_out, k < Kp) is written by the staging below, zeros for the K padding included.
        const uint32_t n16 = (p.off_ystage - p.off_whi) >> 4;
        const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t i = tid; i < n16; i += kBlock) sts_f<4>(base + p.off_whi + (i << 4), z4);
    }
    pdl_wait();   // everything above overlaps the previous kernel's tail
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == kWarpTma) {
        // =============================== TMA producer ===============================
        reg_dec<kRegsMisc>();
        if (lane == 0) {
            PhaseTimer pt(dbg != nullptr);
            Ring rs;
            for (int it = 0; it < n_tiles; ++it) {
                pt.mark(1);
                mbar_wait_relaxed(&bar_empty[rs.idx], rs.phase ^ 1u);   // a fresh barrier passes the parity-1 wait
                pt.mark(0);
                const int64_t g0 = tr.g_begin + static_cast<int64_t>(it) * p.G;
                const int ng = (it == n_tiles - 1) ? last_ng : p.G;
                const int64_t r0 = g0 * p.C_csr * N;
                const int rows_csr = ng * p.C_csr * N;
                unsigned char* st = gen + p.off_stage + static_cast<size_t>(rs.idx) * p.stage_bytes;
                uint64_t* full = &bar_full[rs.idx];
                // the feature rows and the row extents do not depend on anything: they go first; the column / value
                // slices need the tile's first and last entry index (one dependent global load)
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
                rs.advance(S);
            }
            if (dbg) dbg[10] = pt.acc[0];
        }
    } else {
        // B operand [W ; bias] -> (hi, lo), K-major SWIZZLE_128B: B row n = output column col0 + n, k = c * f_in + f for
        // the weights, k = K + c for bias_c (it meets rowsum(A_c) in column K + c of Z).  All non-producer warps.
        {
            constexpr int kStagers = kBlock - 32;
            const int col0 = static_cast<int>(blockIdx.y) * f_out;
            if (p.w_trans == 0) {
                const int nq_n = f_out >> 2, kq_n = Kp >> 2;   // one thread per 4 (k) x 4 (n) block: 4 coalesced 16-byte loads
                for (int idx = tid; idx < kq_n * nq_n; idx += kStagers) {
                    const int n0 = (idx % nq_n) << 2, k4 = (idx / nq_n) << 2;
                    float4 r[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int kk = k4 + j;
                        r[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (kk < K) {
                            const int c = kk / f_in, f = kk - c * f_in;
                            r[j] = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(c) * p.w_cstride +
                                                                         static_cast<size_t>(f) * p.w_ld + col0 + n0));
                        } else if (kk - K < C && p.bias != nullptr) {
                            r[j] = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(kk - K) * p.w_ld + col0 + n0));
                        }
                    }
                    const float t[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                                           {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
                    for (int nn = 0; nn < 4; ++nn) {
                        float hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            hi[j] = tf32_hi(t[nn][j]);
                            lo[j] = t[nn][j] - hi[j];
                        }
                        const uint32_t off = sw128_offset(n0 + nn, k4, p.w_atom);
                        sts_f<4>(base + p.off_whi + off, hi);
                        sts_f<4>(base + p.off_wlo + off, lo);
                    }
                }
            } else {
                // transposed weights (dx = G . W^T): B(n, c * f_in + f) = w[c][col0 + n][f] -- k is the contiguous index
                const int kq = f_in >> 2;
                for (int idx = tid; idx < f_out * C * kq; idx += kStagers) {
                    const int n = idx / (C * kq), r = idx - n * (C * kq), c = r / kq, f4 = (r - c * kq) << 2;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(c) * p.w_cstride +
                                                                           static_cast<size_t>(col0 + n) * p.w_ld + f4));
                    const float t[4] = {v.x, v.y, v.z, v.w};
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hi[j] = tf32_hi(t[j]);
                        lo[j] = t[j] - hi[j];
                    }
                    const uint32_t off = sw128_offset(n, c * f_in + f4, p.w_atom);
                    sts_f<4>(base + p.off_whi + off, hi);
                    sts_f<4>(base + p.off_wlo + off, lo);
                }
                const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // no bias: the row-sum columns K .. K + 7 of Z meet exact zeros
                for (int idx = tid; idx < f_out * 2; idx += kStagers) {
                    const uint32_t off = sw128_offset(idx >> 1, K + ((idx & 1) << 2), p.w_atom);
                    sts_f<4>(base + p.off_whi + off, z4);
                    sts_f<4>(base + p.off_wlo + off, z4);
                }
            }
        }
        fence_proxy_async_smem();   // B is read by the tensor core through the async proxy
        if (warp < 4) {   // the unused row-sum columns K + C .. K + 7 of every Z buffer stay zero for the whole kernel
            const uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int b = 0; b < p.zbufs; ++b) {
                const uint32_t zc = tmem + (static_cast<uint32_t>(warp * 32) << 16) + p.tm_z + static_cast<uint32_t>(b * 2 * Kp);
                tmem_st8(zc + K, z8);
                tmem_st8(zc + Kp + K, z8);
            }
            tmem_st_wait();
            tc_fence_before_sync();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kBlock - 32) : "memory");
        tc_fence_after_sync();

        if (warp < kAggWarps) {
            // =============================== aggregation warps ===============================
            reg_inc<kRegsAgg>();
            const int wq = warp & 3;          // TMEM lane quarter of this warp
            const int phase = warp >> 2;      // slab phase: slabs phase, phase + 2, ...
            const uint32_t s7 = static_cast<uint32_t>(lane) & 7u;
            const bool p1 = (s7 & 1u) != 0, p2 = (s7 & 2u) != 0, p4 = (s7 & 4u) != 0;
            const int w = wq * 32 + lane;     // tile row = TMEM lane
            const int gl = w / N, node = w - gl * N;
            const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
            const int full_rows = p.G * N;
            const uint32_t r0_step = static_cast<uint32_t>(p.G * p.C_csr * N);
            uint32_t r0_lo = static_cast<uint32_t>((tr.g_begin * p.C_csr * N) & 3);
            const uint32_t z_stride = static_cast<uint32_t>(2 * Kp);
            const uint32_t row_rp_off = 4u * static_cast<uint32_t>((gl * p.C_csr + p.c_begin) * N + node);
            const uint32_t row_x_off = static_cast<uint32_t>(gl * N) * pitch + (s7 << 4);
            Ring rs, rz;
            PhaseTimer pt(dbg != nullptr && warp == 0 && lane == 0);
            if (pt.on) dbg[12] = pt.last - t_start;
            for (int it = 0; it < n_tiles; ++it) {
                const bool last = it == n_tiles - 1;
                const int rows = last ? last_ng * N : full_rows;
                const int rows_csr = last ? last_ng * p.C_csr * N : static_cast<int>(r0_step);
                const uint32_t st = base + p.off_stage + static_cast<uint32_t>(rs.idx) * p.stage_bytes;
                mbar_wait(&bar_full[rs.idx], rs.phase);
                pt.mark(0);
                const uint32_t rp_addr = st + p.st_rp + 4u * (r0_lo & 3u);
                const int e_first = static_cast<int>(lds_u32(rp_addr));
                const int e_last = static_cast<int>(lds_u32(rp_addr + 4u * static_cast<uint32_t>(rows_csr)));
                const int e_lo = e_first & ~3;
                const bool staged = static_cast<uint32_t>((e_last - e_lo + 3) & ~3) <= static_cast<uint32_t>(p.cv_cap);
                const uint32_t col_addr = st + p.st_col - 4u * static_cast<uint32_t>(e_lo);   // entry e at col_addr + 4 e
                const uint32_t val_addr = st + p.st_val - 4u * static_cast<uint32_t>(e_lo);
                const bool valid = w < rows;

                mbar_wait(&bar_zempty[rz.idx], rz.phase ^ 1u);
                tc_fence_after_sync();
                pt.mark(1);
                const uint32_t zc = tmem + lane_sel + p.tm_z + static_cast<uint32_t>(rz.idx) * z_stride;
                const uint32_t row_rp = rp_addr + row_rp_off;
                const uint32_t row_x = st + row_x_off;
                int c = 0, fs = phase;
                for (int slab = phase; slab < p.n_slabs; slab += 2) {
                    while (fs >= p.slabs_per_ch) { fs -= p.slabs_per_ch; ++c; }
                    int rs_ = e_first, re_ = e_first;   // rows beyond the tile: empty
                    if (valid) {
                        const uint32_t ra = row_rp + 4u * static_cast<uint32_t>(c * N);
                        rs_ = static_cast<int>(lds_u32(ra));
                        re_ = static_cast<int>(lds_u32(ra + 4u));
                    }
                    const uint32_t xbase = row_x + static_cast<uint32_t>(fs) * 128u;
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
                    float deg = 0.0f;
                    if (staged) gather_row<true>(acc, deg, rs_, re_, col_addr, val_addr, p.col, p.val, xbase, pitch);
                    else gather_row<false>(acc, deg, rs_, re_, col_addr, val_addr, p.col, p.val, xbase, pitch);
                    // undo the per-lane chunk rotation: block i holds chunk i ^ s7 -> three conditional butterflies
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float a0 = acc[4 * i + j], a1 = acc[4 * (i + 1) + j];
                            acc[4 * i + j] = p1 ? a1 : a0;
                            acc[4 * (i + 1) + j] = p1 ? a0 : a1;
                        }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if ((i & 2) == 0)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float a0 = acc[4 * i + j], a1 = acc[4 * (i + 2) + j];
                                acc[4 * i + j] = p2 ? a1 : a0;
                                acc[4 * (i + 2) + j] = p2 ? a0 : a1;
                            }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float a0 = acc[4 * i + j], a1 = acc[4 * (i + 4) + j];
                            acc[4 * i + j] = p4 ? a1 : a0;
                            acc[4 * (i + 4) + j] = p4 ? a0 : a1;
                        }
                    __syncwarp();   // the gather loop is divergent; tcgen05.st is warp-collective
                    pt.mark(2);
                    // tf32 split by truncation: hi keeps the top 19 bits, lo = x - hi is exact (|lo| < 2^-10 |x|) and is
                    // itself read as tf32 by the tensor core -> 21 mantissa bits survive, 2 instructions per element
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        hi[i] = __float_as_uint(acc[i]) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(acc[i] - __uint_as_float(hi[i]));
                    }
                    tmem_st32(zc + static_cast<uint32_t>(slab * 32), hi);
                    tmem_st32(zc + static_cast<uint32_t>(Kp + slab * 32), lo);
                    if (fs == 0) {   // rowsum(A_c) rides in column K + c and meets bias_c in the contraction
                        const float h = tf32_hi(deg);
                        tmem_st1(zc + static_cast<uint32_t>(K + c), __float_as_uint(h));
                        tmem_st1(zc + static_cast<uint32_t>(Kp + K + c), __float_as_uint(deg - h));
                    }
                    fs += 2;
                }
                pt.mark(3);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[rs.idx]);   // this warp is done reading the stage
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_zfull[rz.idx]);
                pt.mark(4);
                rs.advance(S);
                rz.advance(p.zbufs);
                r0_lo += r0_step;
            }
            if (pt.on) {
                for (int i = 0; i < 5; ++i) dbg[i] = pt.acc[i];
                dbg[11] = clock64() - t_start;
                dbg[13] = n_tiles;
            }
        } else if (warp >= kWarpMma) {
            reg_dec<kRegsMisc>();
            if (warp == kWarpMma) {
                // =============================== MMA issuer ===============================
                // the whole warp walks the tile loop converged; one elected lane issues (operands stay on the uniform datapath)
                const uint32_t idesc = umma_idesc_tf32(128, Np);
                const uint64_t dwhi = umma_desc_sw128(base + p.off_whi), dwlo = umma_desc_sw128(base + p.off_wlo);
                const uint32_t w_atom16 = p.w_atom >> 4;
                const int ks = Kp >> 3;
                Ring rz, ra;
                PhaseTimer pt(dbg != nullptr && lane == 0);
                for (int it = 0; it < n_tiles; ++it) {
                    mbar_wait(&bar_zfull[rz.idx], rz.phase);
                    pt.mark(0);
                    mbar_wait(&bar_tempty[ra.idx], ra.phase ^ 1u);
                    tc_fence_after_sync();
                    pt.mark(1);
                    const uint32_t d = tmem + static_cast<uint32_t>(ra.idx * Np);
                    const uint32_t zhi = tmem + p.tm_z + static_cast<uint32_t>(rz.idx * 2 * Kp), zlo = zhi + static_cast<uint32_t>(Kp);
                    __syncwarp();
                    if (elect_one()) {
                        switch (ks) {
                            case 5: issue_tile<5>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                            case 9: issue_tile<9>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                            case 13: issue_tile<13>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                            case 17: issue_tile<17>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                            default: issue_tile_loop(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16, ks);
                        }
                        umma_commit(&bar_zempty[rz.idx]);   // Z buffer may be overwritten once these MMAs have read it
                        umma_commit(&bar_tfull[ra.idx]);    // accumulator is complete
                    }
                    __syncwarp();
                    pt.mark(2);
                    rz.advance(p.zbufs);
                    ra.advance(p.abufs);
                }
                if (pt.on) for (int i = 0; i < 3; ++i) dbg[5 + i] = pt.acc[i];
            }
        } else {
            // =============================== epilogue warps ===============================
            // warp (q, h): TMEM lanes 32 q .. 32 q + 31, 32-column slabs h, h + 2, ...  The activated 32 x 32 block goes
            // through a swizzled per-warp staging tile and leaves as 16-byte stores, 128 contiguous bytes per row.
            reg_dec<kRegsEpi>();
            const int e = warp - kWarpEpi0;
            const int wq = e & 3, h = e >> 2;
            const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
            const uint32_t ys = base + p.off_ystage + static_cast<uint32_t>(e) * 4096u;   // 32 rows x 128 B
            const uint32_t l7 = static_cast<uint32_t>(lane) & 7u;
            const int n_cslabs = (Np + 31) >> 5;
            const int full_rows = p.G * N;
            const uint32_t yrow = ys + static_cast<uint32_t>(lane) * 128u;   // this lane's staging row; chunk c at c ^ (lane & 7)
            const uint32_t ysrc = ys + (static_cast<uint32_t>(lane) >> 3) * 128u;   // copy-out: rows 4 k + lane / 8, chunk lane & 7
            const int colq = static_cast<int>(l7) * 4;
            const int row0 = wq * 32 + (lane >> 3);   // first of the 8 tile rows (stride 4) this lane copies out
            const size_t y_ld = static_cast<size_t>(p.y_ld);
            float* y_tile = p.y + (tr.g_begin * N + row0) * y_ld + static_cast<size_t>(blockIdx.y) * f_out + colq;
            const size_t y_step = static_cast<size_t>(full_rows) * y_ld;
            Ring ra;
            PhaseTimer pt(dbg != nullptr && e == 0 && lane == 0);
            for (int it = 0; it < n_tiles; ++it) {
                const int rows = (it == n_tiles - 1) ? last_ng * N : full_rows;
                mbar_wait_relaxed(&bar_tfull[ra.idx], ra.phase);
                tc_fence_after_sync();
                pt.mark(0);
                const uint32_t ta = tmem + lane_sel + static_cast<uint32_t>(ra.idx * Np);
                for (int cs = h; cs < n_cslabs; cs += 2) {
                    float v0[16], v1[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v1[i] = 0.0f;
                    tmem_ld16(ta + static_cast<uint32_t>(cs * 32), v0);
                    if (cs * 32 + 16 < Np) tmem_ld16(ta + static_cast<uint32_t>(cs * 32 + 16), v1);
                    tmem_ld_wait();
                    tmem_ld_fence(v0);
                    tmem_ld_fence(v1);
                    pt.mark(2);
                    if (cs + 2 >= n_cslabs) {   // this warp has read all its columns: hand the accumulator back
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_tempty[ra.idx]);
                    }
                    if (EPI != 2) {
                        act16_rt(v0, p.act);
                        act16_rt(v1, p.act);
                    }
                    if (EPI != 2 && cs * 32 + 32 > p.f_valid) {   // padded output columns: act(0) may be non-zero (sigmoid) -- force exact zeros
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            if (cs * 32 + i >= p.f_valid) v0[i] = 0.0f;
                            if (cs * 32 + 16 + i >= p.f_valid) v1[i] = 0.0f;
                        }
                    }
                    tmem_ld_fence(v0);
                    tmem_ld_fence(v1);
                    pt.mark(3);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float t0[4] = {v0[4 * c4], v0[4 * c4 + 1], v0[4 * c4 + 2], v0[4 * c4 + 3]};
                        const float t1[4] = {v1[4 * c4], v1[4 * c4 + 1], v1[4 * c4 + 2], v1[4 * c4 + 3]};
                        sts_f<4>(yrow + ((static_cast<uint32_t>(c4) ^ l7) << 4), t0);
                        sts_f<4>(yrow + ((static_cast<uint32_t>(c4 + 4) ^ l7) << 4), t1);
                    }
                    __syncwarp();
                    pt.mark(4);
                    // copy-out: each store instruction writes 4 rows x 128 contiguous bytes
                    const bool col_ok = cs * 32 + colq < f_out;
                    float* ycs = y_tile + cs * 32;
                    if (EPI == 0) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                            float t[4];
                            lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                            if (row0 + 4 * k < rows && col_ok)
                                *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) = make_float4(t[0], t[1], t[2], t[3]);
                        }
                    } else if (EPI == 2) {
                        float4 ov[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            ov[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            if (row0 + 4 * k < rows && col_ok) ov[k] = *reinterpret_cast<const float4*>(ycs + static_cast<size_t>(4 * k) * y_ld);
                        }
                        const int cbase = cs * 32 + colq;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                            float t[4];
                            lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                            const float o[4] = {ov[k].x, ov[k].y, ov[k].z, ov[k].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                t[j] += o[j];
                                switch (p.act) {
                                    case KGCN_ACT_RELU: t[j] = fast_act<KGCN_ACT_RELU>(t[j]); break;
                                    case KGCN_ACT_SIGMOID: t[j] = fast_act<KGCN_ACT_SIGMOID>(t[j]); break;
                                    case KGCN_ACT_TANH: t[j] = fast_act<KGCN_ACT_TANH>(t[j]); break;
                                    default: break;
                                }
                                if (cbase + j >= p.f_valid) t[j] = 0.0f;
                            }
                            if (row0 + 4 * k < rows && col_ok)
                                *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) = make_float4(t[0], t[1], t[2], t[3]);
                        }
                    } else {
                        const float* mcs = p.mul_src + (ycs - p.y);   // same [rows, y_ld] layout as the output
                        float4 mv[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            mv[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            if (row0 + 4 * k < rows && col_ok) mv[k] = __ldg(reinterpret_cast<const float4*>(mcs + static_cast<size_t>(4 * k) * y_ld));
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                            float t[4];
                            lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                            if (row0 + 4 * k < rows && col_ok)
                                *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) =
                                    make_float4(t[0] * act_grad_from_output(mv[k].x, p.mul_act), t[1] * act_grad_from_output(mv[k].y, p.mul_act),
                                                t[2] * act_grad_from_output(mv[k].z, p.mul_act), t[3] * act_grad_from_output(mv[k].w, p.mul_act));
                        }
                    }
                    __syncwarp();
                    pt.mark(5);
                }
                if (h >= n_cslabs) {   // no columns for this warp (f_out <= 32): it still takes part in the hand-back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_tempty[ra.idx]);
                }
                ra.advance(p.abufs);
                y_tile += y_step;
            }
            if (pt.on) {
                dbg[8] = pt.acc[0]; dbg[9] = pt.acc[1]; dbg[14] = pt.acc[2]; dbg[15] = pt.acc[3];
                p.dbg[148 * 16 + blockIdx.x * 2] = pt.acc[4]; p.dbg[148 * 16 + blockIdx.x * 2 + 1] = pt.acc[5];
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem, 512);
}


// ================================================================================================================
// Chained layers: ONE launch runs several fused layers back to back over the same batch.  A CTA owns the same graph
// range in every job, and a graph's rows never leave the CTA (the aggregation is graph-local), so job j + 1 may start
// on a CTA as soon as THAT CTA has finished job j: no grid-wide dependency, no launch / drain / fill per layer.
//   forward chain   x -> GraphConv_0 -> .. -> GraphConv_{L-1}                      (EPI 0: activation epilogue)
//   dx chain        dU_{L-1} -> dU_{L-2} -> ..  each job (A^T, dU_l, W_l^T) x act'(x_l)   (EPI 1)
// HARD job boundary = one CTA-wide barrier: the epilogue warps have fenced their global stores towards the async proxy (the
// next job's TMA reads them), all MMAs that read the B operand have completed, every stage has been consumed.  The
// mbarriers are NOT re-initialised: every role tracks one phase bit per barrier slot, so ring sizes may differ per job.
// SOFT job boundary (consecutive jobs with the same plan, i.e. identical shared / tensor memory layout and tiling): no
// barrier at all.  Tile t of job j + 1 depends only on tile t of job j (same graphs), so the epilogue warps publish every
// finished tile -- stores, fence.proxy.async, release-increment of a shared counter -- and the TMA producer of job j + 1
// acquires the counter before it loads tile t.  All rings simply continue; the aggregation of (j + 1, 0) overlaps the
// epilogue of (j, last).  With two B operand buffers the two stager warps write [W ; bias] of job j + 1 into the free one
// during job j (as soon as the MMAs of job j - 1 have completed), so the MMAs of job j + 1 wait for nothing but Z.
// ON-CHIP HAND-OVER (a soft job that reads its predecessor's output, CTA's tiles <= stages): the epilogue writes the output
// tile it stores to HBM also into the successor's stage slot (slot = tile index; the slot's last reader was this job's
// own aggregation of the same tile) and arrives on the slot's full barrier -- 8 epilogue-warp arrivals + the producer's one,
// which only fetches the CSR slices.  No proxy fence, no reload: the successor's aggregation starts when the tile is stored.
constexpr int kV4MaxJobs = 6;
struct V4Batch {
    int n_jobs;
    int head_job, head_setup_job;   // the job with the fused head (-1: none) and the job at whose start its shared memory is set up
    int early;            // job 0's inputs are stable: only the epilogue warps (parameters, all global writes) wait for the
                          // previous grid; the TMA producer and the aggregation start while it is still running
    int tile_fence_gpu;   // A-B knob (KGCN_CHAIN_TILE_FENCE=1): membar.gl before the proxy fence of a published tile
    long long* dbg;   // tuning aid (kgcn_debug_v4_chain_times): [CTA][64] clock64 stamps, see tools/chain_timeline.py
    V4Params job[kV4MaxJobs];
};
// stamp slot: job * 16 + event; events: 0 job start (agg), 1 first stage landed, 2 agg tile 0 done, 3 agg last tile done,
// 4 MMA first issue done, 5 MMA last issue done, 6 epi first accumulator ready, 7 epi first tile stored, 8 epi last tile stored,
// 9 TMA first issue, 10 B operand staged, 11 epi job end (after fences)
#define V4_STAMP(ev) do { if (b.dbg != nullptr && lane == 0) b.dbg[static_cast<size_t>(blockIdx.x) * 128 + j * 16 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t ld_acquire_cta_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_cta_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// end of one tile's epilogue (any variant): publish the tile to a soft successor's TMA loads (stores -> async proxy), count it,
// advance the accumulator ring and the output pointer
#define PUBLISH_TILE()                                                     \
    do {                                                                   \
        if (handover) {                                                    \
            __syncwarp();                                                  \
            if (lane == 0) mbar_arrive(&bar_full[it]);                     \
        } else if (p.soft_next) {                                          \
            if (b.tile_fence_gpu) __threadfence();                         \
            fence_proxy_async_all();                                       \
            __syncwarp();                                                  \
        }                                                                  \
        if (lane == 0) red_release_cta_add_u32(&tiles_done, 1u);           \
        if (e == 0 && it == 0) V4_STAMP(7);                                \
        if (e == 0 && it == n_tiles - 1) V4_STAMP(8);                      \
        if (++ai == p.abufs) ai = 0;                                       \
        y_tile += y_step;                                                  \
    } while (0)

__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t cnt) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt) : "memory");
}
// CTA-wide barrier of all warp roles on an mbarrier (count = warps of the CTA): every thread keeps its own phase bit.  (A named
// bar.sync reached from the roles' different program counters is legal, but compute-sanitizer's synccheck reports it.)
#define CTA_ROLE_BARRIER()                                  \
    do {                                                    \
        __syncwarp();                                       \
        if (lane == 0) {   /* one polling lane per warp, with back-off: waiting roles must not take issue slots */ \
            mbar_arrive(&bar_roles);                        \
            while (!mbar_try_wait(&bar_roles, static_cast<uint32_t>(p.hard_ord - 1) & 1u)) {   /* phase = ordinal of the hard boundary */ \
            }                                               \
        }                                                   \
        __syncwarp();                                       \
    } while (0)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ TileRange cta_range_of(const V4Params& p) {
    TileRange t;
    t.g_begin = static_cast<int64_t>(blockIdx.x) * p.graphs_per_cta;
    const int64_t left = p.n_graphs - t.g_begin;
    t.n_graphs_cta = static_cast<int>(left < p.graphs_per_cta ? (left > 0 ? left : 0) : p.graphs_per_cta);
    t.n_tiles = (t.n_graphs_cta + p.G - 1) / p.G;
    return t;
}

// [W ; bias] (or W^T) of one job -> (hi, lo) K-major SWIZZLE_128B B operand at `base` (+ off_whi / off_wlo), by `n_thr`
// threads (index t).  ONE copy of this code in the kernel (it runs once or twice per job: a real call, scalar arguments --
// a reference to the kernel parameters would force a local copy of the whole parameter block).
struct StageB {
    const float* w;
    const float* bias;
    int C, f_in, f_out, K, Kp, w_trans, w_ld, w_cstride;
    uint32_t w_atom, off_whi, off_wlo, w_pair;
};
__device__ __forceinline__ StageB stage_args(const V4Params& p) {
    return StageB{p.w, p.bias, p.C, p.f_in, p.f_out, p.K, p.Kp, p.w_trans, p.w_ld, p.w_cstride, p.w_atom, p.off_whi, p.off_wlo, p.w_pair};
}
// (CALLER: one instance per calling role -- ptxas wants all callers of a function under the same setmaxnreg budget)
template <int CALLER>
__device__ __noinline__ void stage_b_operand(const StageB p, uint32_t base, int t, int n_thr) {
    const int C = p.C, f_in = p.f_in, f_out = p.f_out, K = p.K, Kp = p.Kp;
    // B rows n in [f_out, Np) are read by the MMAs: the loops below run over Np rows and write them as exact zeros
    const int Np = (f_out + 15) & ~15;
    if (p.w_trans == 0) {
        const int nq_n = Np >> 2, kq_n = Kp >> 2;
        for (int idx = t; idx < kq_n * nq_n; idx += n_thr) {
            const int n0 = (idx % nq_n) << 2, k4 = (idx / nq_n) << 2;
            float4 r[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = k4 + j;
                r[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (n0 >= f_out) continue;   // f_out % 4 == 0: a quad of output columns is inside or outside as a whole
                if (kk < K) {
                    const int c = kk / f_in, f = kk - c * f_in;
                    r[j] = __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(c) * p.w_cstride + static_cast<size_t>(f) * p.w_ld + n0));
                } else if (kk - K < C && p.bias != nullptr) {
                    r[j] = __ldg(reinterpret_cast<const float4*>(p.bias + static_cast<size_t>(kk - K) * p.w_ld + n0));
                }
            }
            const float tt[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                                    {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    hi[j] = tf32_hi(tt[nn][j]);
                    lo[j] = tt[nn][j] - hi[j];
                }
                const uint32_t off = sw128_offset(n0 + nn, k4, p.w_atom);
                sts_f<4>(base + p.off_whi + off, hi);
                sts_f<4>(base + p.off_wlo + off, lo);
            }
        }
    } else {
        const int kq = f_in >> 2;
        for (int idx = t; idx < Np * C * kq; idx += n_thr) {
            const int n = idx / (C * kq), r = idx - n * (C * kq), c = r / kq, f4 = (r - c * kq) << 2;
            const float4 v = n < f_out ? __ldg(reinterpret_cast<const float4*>(p.w + static_cast<size_t>(c) * p.w_cstride + static_cast<size_t>(n) * p.w_ld + f4))
                                       : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float tt[4] = {v.x, v.y, v.z, v.w};
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                hi[j] = tf32_hi(tt[j]);
                lo[j] = tt[j] - hi[j];
            }
            const uint32_t off = sw128_offset(n, c * f_in + f4, p.w_atom);
            sts_f<4>(base + p.off_whi + off, hi);
            sts_f<4>(base + p.off_wlo + off, lo);
        }
        const float z4[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // no bias: the row-sum columns K .. K + 7 of Z meet exact zeros
        for (int idx = t; idx < Np * 2; idx += n_thr) {
            const uint32_t off = sw128_offset(idx >> 1, K + ((idx & 1) << 2), p.w_atom);
            sts_f<4>(base + p.off_whi + off, z4);
            sts_f<4>(base + p.off_wlo + off, z4);
        }
    }
    fence_proxy_async_smem();   // B is read by the tensor core through the async proxy
}

template <int ACT>
__global__ void __maxnreg__(96) graphconv_fused_v4_chain_kernel(const V4Batch b) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t bar_full[kV4MaxStages], bar_empty[kV4MaxStages];
    __shared__ __align__(8) uint64_t bar_zfull[2], bar_zempty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar_wfull;    // B operand of a prestaged (soft) job written by the two stager warps
    __shared__ __align__(8) uint64_t bar_wempty;   // all MMAs of a job have completed (its B operand buffer is free)
    __shared__ __align__(8) uint64_t bar_w0full;   // B operand staged at the job's own start by the epilogue + stager warps
    __shared__ __align__(8) uint64_t bar_roles;    // CTA-wide role barrier (hard job boundaries)
    __shared__ uint32_t tiles_done;   // epilogue-warp arrivals: 8 per finished tile, over all jobs (soft job boundaries)

    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* gen = smem_dyn + (base - smem_u32(smem_dyn));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_jobs = b.n_jobs;
    if (b.dbg != nullptr && tid == 0) b.dbg[static_cast<size_t>(blockIdx.x) * 128 + 126] = clock64();

    if (tid == 0) {
        tiles_done = 0;
        mbar_init(&bar_wfull, kStagers / 32);
        mbar_init(&bar_wempty, 1);
        mbar_init(&bar_w0full, kEpiWarps + kStagers / 32);
        mbar_init(&bar_roles, kBlock / 32);
        for (int i = 0; i < kV4MaxStages; ++i) {
            mbar_init(&bar_full[i], 1 + kEpiWarps);   // producer + (hand-over) the epilogue warps; the producer arrives for them otherwise
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
    if (!b.early) pdl_wait();   // everything above overlaps the previous kernel's tail
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == kWarpTma) {
        // =============================== TMA producer ===============================
        reg_dec<kRegsMisc>();
        uint32_t ph_empty = 0;   // one phase bit per barrier slot (ring sizes may change between jobs)
        uint32_t tiles_before = 0;   // tiles of all earlier jobs
        for (int j = 0; j < n_jobs; ++j) {
            const V4Params& p = b.job[j];
            if (j > 0 && !p.soft) CTA_ROLE_BARRIER();   // the previous job's outputs are complete and visible to the async proxy
            if (lane == 0) {
                const int C_csr = p.C_csr, N = p.N, f_in = p.f_in, S = p.n_stages;
                const uint32_t pitch = static_cast<uint32_t>(f_in) * 4u;
                const TileRange tr = cta_range_of(p);
                const int n_tiles = tr.n_tiles;
                const int last_ng = tr.n_graphs_cta - (n_tiles - 1) * p.G;
                int s = 0;
                for (int it = 0; it < n_tiles; ++it) {
                    mbar_wait_relaxed(&bar_empty[s], ((ph_empty >> s) & 1u) ^ 1u);
                    ph_empty ^= 1u << s;
                    const bool handover = p.x_prev && n_tiles <= S;   // the previous job's epilogue fills x of this slot
                    if (p.soft && !handover) {   // tile `it` of the previous job (same graphs) has been stored and fenced by all epilogue warps
                        const uint32_t need = static_cast<uint32_t>(kEpiWarps) * (tiles_before - static_cast<uint32_t>(n_tiles) + static_cast<uint32_t>(it) + 1u);
                        while (ld_acquire_cta_u32(&tiles_done) < need) __nanosleep(20);
                    }
                    const int64_t g0 = tr.g_begin + static_cast<int64_t>(it) * p.G;
                    const int ng = (it == n_tiles - 1) ? last_ng : p.G;
                    const int64_t r0 = g0 * C_csr * N;
                    const int rows_csr = ng * C_csr * N;
                    unsigned char* st = gen + p.off_stage + static_cast<size_t>(s) * p.stage_bytes;
                    uint64_t* full = &bar_full[s];
                    const int64_t rp_lo = r0 & ~3ll;
                    const uint32_t rp_cnt = static_cast<uint32_t>((r0 + rows_csr + 1 - rp_lo + 3) & ~3ll);
                    const uint32_t x_bytes = static_cast<uint32_t>(ng) * static_cast<uint32_t>(N) * pitch;
                    if (it == 0) V4_STAMP(9);
                    mbar_expect_tx_only(full, (handover ? 0u : x_bytes) + 4u * rp_cnt);
                    if (!handover) bulk_g2s(st, p.x + g0 * N * f_in, x_bytes, full);
                    bulk_g2s(st + p.st_rp, p.rowptr + rp_lo, 4u * rp_cnt, full);
                    const int32_t e_first = __ldg(p.rowptr + r0), e_last = __ldg(p.rowptr + r0 + rows_csr);
                    const int32_t e_lo = e_first & ~3;
                    const uint32_t e_cnt = static_cast<uint32_t>((e_last - e_lo + 3) & ~3);
                    const bool staged = e_cnt <= static_cast<uint32_t>(p.cv_cap) && e_cnt != 0;
                    mbar_expect_tx(full, staged ? 8u * e_cnt : 0u);   // the producer's arrival of the phase
                    if (!handover) mbar_arrive_cnt(full, kEpiWarps);   // nobody else fills this slot
                    if (staged) {
                        bulk_g2s(st + p.st_col, p.col + e_lo, 4u * e_cnt, full);
                        bulk_g2s(st + p.st_val, p.val + e_lo, 4u * e_cnt, full);
                    }
                    if (++s == S) s = 0;
                }
                tiles_before += static_cast<uint32_t>(n_tiles);
            }
            __syncwarp();
        }
    } else if (warp < kAggWarps) {
        // =============================== aggregation warps ===============================
        reg_inc<kRegsAgg>();
        uint32_t ph_full = 0, ph_zempty = 0;
        const int wq = warp & 3;          // TMEM lane quarter of this warp
        const int phase = warp >> 2;      // slab phase: slabs phase, phase + 2, ...
        const uint32_t s7 = static_cast<uint32_t>(lane) & 7u;
        const bool p1 = (s7 & 1u) != 0, p2 = (s7 & 2u) != 0, p4 = (s7 & 4u) != 0;
        const int w = wq * 32 + lane;     // tile row = TMEM lane
        const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
        for (int j = 0; j < n_jobs; ++j) {
            const V4Params& p = b.job[j];
            if (j > 0 && !p.soft) CTA_ROLE_BARRIER();
            if (warp == 0) V4_STAMP(0);
            // (the B operand is staged by the epilogue warps, which have nothing else to do until the first accumulator is
            // ready; the aggregation starts as soon as the first tile has landed)
            const int N = p.N, f_in = p.f_in, K = p.K, Kp = p.Kp, S = p.n_stages;
            const uint32_t pitch = static_cast<uint32_t>(f_in) * 4u;
            if (warp < 4 && !p.soft) {   // the unused row-sum columns K + C .. K + 7 of every Z buffer stay zero for the whole job (and its soft successors)
                const uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int zb = 0; zb < p.zbufs; ++zb) {
                    const uint32_t zc = tmem + (static_cast<uint32_t>(warp * 32) << 16) + p.tm_z + static_cast<uint32_t>(zb * 2 * Kp);
                    tmem_st8(zc + K, z8);
                    tmem_st8(zc + Kp + K, z8);
                }
                tmem_st_wait();   // ordered before this warp's first zfull arrival, which the MMA warp waits for
            }
            const TileRange tr = cta_range_of(p);
            const int n_tiles = tr.n_tiles;
            const int last_ng = tr.n_graphs_cta - (n_tiles - 1) * p.G;
            const int gl = w / N, node = w - gl * N;
            const int full_rows = p.G * N;
            const uint32_t r0_step = static_cast<uint32_t>(p.G * p.C_csr * N);
            uint32_t r0_lo = static_cast<uint32_t>((tr.g_begin * p.C_csr * N) & 3);
            const uint32_t z_stride = static_cast<uint32_t>(2 * Kp);
            const uint32_t row_rp_off = 4u * static_cast<uint32_t>((gl * p.C_csr + p.c_begin) * N + node);
            const uint32_t row_x_off = static_cast<uint32_t>(gl * N) * pitch + (s7 << 4);
            int64_t zrow0 = tr.g_begin * N;   // global row index of the current tile's first row (zsave only)
            int s = 0, zi = 0;
            for (int it = 0; it < n_tiles; ++it) {
                const bool last = it == n_tiles - 1;
                const int rows = last ? last_ng * N : full_rows;
                const int rows_csr = last ? last_ng * p.C_csr * N : static_cast<int>(r0_step);
                const uint32_t st = base + p.off_stage + static_cast<uint32_t>(s) * p.stage_bytes;
                mbar_wait(&bar_full[s], (ph_full >> s) & 1u);
                ph_full ^= 1u << s;
                if (warp == 0 && it == 0) V4_STAMP(1);
                const uint32_t rp_addr = st + p.st_rp + 4u * (r0_lo & 3u);
                const int e_first = static_cast<int>(lds_u32(rp_addr));
                const int e_last = static_cast<int>(lds_u32(rp_addr + 4u * static_cast<uint32_t>(rows_csr)));
                const int e_lo = e_first & ~3;
                const bool staged = static_cast<uint32_t>((e_last - e_lo + 3) & ~3) <= static_cast<uint32_t>(p.cv_cap);
                const uint32_t col_addr = st + p.st_col - 4u * static_cast<uint32_t>(e_lo);
                const uint32_t val_addr = st + p.st_val - 4u * static_cast<uint32_t>(e_lo);
                const bool valid = w < rows;

                mbar_wait(&bar_zempty[zi], ((ph_zempty >> zi) & 1u) ^ 1u);
                ph_zempty ^= 1u << zi;
                tc_fence_after_sync();
                const uint32_t zc = tmem + lane_sel + p.tm_z + static_cast<uint32_t>(zi) * z_stride;
                const uint32_t row_rp = rp_addr + row_rp_off;
                const uint32_t row_x = st + row_x_off;
                int c = 0, fs = phase;
                for (int slab = phase; slab < p.n_slabs; slab += 2) {
                    while (fs >= p.slabs_per_ch) { fs -= p.slabs_per_ch; ++c; }
                    int rs_ = e_first, re_ = e_first;
                    if (valid) {
                        const uint32_t ra = row_rp + 4u * static_cast<uint32_t>(c * N);
                        rs_ = static_cast<int>(lds_u32(ra));
                        re_ = static_cast<int>(lds_u32(ra + 4u));
                    }
                    const uint32_t xbase = row_x + static_cast<uint32_t>(fs) * 128u;
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.0f;
                    float deg = 0.0f;
                    if (staged) gather_row<true>(acc, deg, rs_, re_, col_addr, val_addr, p.col, p.val, xbase, pitch);
                    else gather_row<false>(acc, deg, rs_, re_, col_addr, val_addr, p.col, p.val, xbase, pitch);
#pragma unroll
                    for (int i = 0; i < 8; i += 2)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float a0 = acc[4 * i + jj], a1 = acc[4 * (i + 1) + jj];
                            acc[4 * i + jj] = p1 ? a1 : a0;
                            acc[4 * (i + 1) + jj] = p1 ? a0 : a1;
                        }
                    if (p.zsave != nullptr && valid) {
                        // the row's slab of Z in fp32 (for a dx job: G = A^T . dU, which the weight-gradient launch then reads instead
                        // of gathering it again).  After the first un-rotation stage the blocks (i, i + 1), i even, hold the 16-byte
                        // chunks (q, q + 1), q = i ^ (s7 & 6), of the slab's 128-byte line: four 32-byte stores, one full sector each
                        // (every lane writes another row, so a store instruction is 32 requests whatever its width)
                        const uint64_t zrow = reinterpret_cast<uint64_t>(p.zsave) + (static_cast<uint64_t>(zrow0 + w) * static_cast<uint64_t>(K) + static_cast<uint64_t>(slab * 32)) * 4u +
                                              ((s7 & 6u) << 4);
#pragma unroll
                        for (int i = 0; i < 8; i += 2)
                            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(zrow ^ static_cast<uint64_t>(i << 4)),
                                         "f"(acc[4 * i]), "f"(acc[4 * i + 1]), "f"(acc[4 * i + 2]), "f"(acc[4 * i + 3]), "f"(acc[4 * i + 4]),
                                         "f"(acc[4 * i + 5]), "f"(acc[4 * i + 6]), "f"(acc[4 * i + 7]) : "memory");
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if ((i & 2) == 0)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float a0 = acc[4 * i + jj], a1 = acc[4 * (i + 2) + jj];
                                acc[4 * i + jj] = p2 ? a1 : a0;
                                acc[4 * (i + 2) + jj] = p2 ? a0 : a1;
                            }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float a0 = acc[4 * i + jj], a1 = acc[4 * (i + 4) + jj];
                            acc[4 * i + jj] = p4 ? a1 : a0;
                            acc[4 * (i + 4) + jj] = p4 ? a0 : a1;
                        }
                    __syncwarp();
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        hi[i] = __float_as_uint(acc[i]) & 0xFFFFE000u;
                        lo[i] = __float_as_uint(acc[i] - __uint_as_float(hi[i]));
                    }
                    tmem_st32(zc + static_cast<uint32_t>(slab * 32), hi);
                    tmem_st32(zc + static_cast<uint32_t>(Kp + slab * 32), lo);
                    if (fs == 0) {
                        const float h = tf32_hi(deg);
                        tmem_st1(zc + static_cast<uint32_t>(K + c), __float_as_uint(h));
                        tmem_st1(zc + static_cast<uint32_t>(Kp + K + c), __float_as_uint(deg - h));
                    }
                    fs += 2;
                }
                zrow0 += full_rows;
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_zfull[zi]);
                if (warp == 0 && it == 0) V4_STAMP(2);
                if (warp == 0 && it == n_tiles - 1) V4_STAMP(3);
                if (++s == S) s = 0;
                if (++zi == p.zbufs) zi = 0;
                r0_lo += r0_step;
            }
        }
    } else if (warp >= kWarpMma) {
        reg_dec<kRegsMisc>();
        uint32_t ph_zfull = 0, ph_tempty = 0, ph_wfull = 0, ph_wempty = 0, ph_w0full = 0;
        if (warp == kWarpMma + 1 || warp == kWarpMma + 2) {
            // =============================== B operand stagers ===============================
            // [W ; bias] of a soft job is written into the free B operand buffer one job ahead, off the epilogue warps'
            // time line; jobs that are staged at their own start (the first one, hard boundaries, one buffer) get these
            // 64 threads on top of the 256 epilogue threads.
            const int ts = tid - (kWarpMma + 1) * 32;
            if (b.early) pdl_wait();   // parameters belong to the previous grid until it completes
            for (int j = 0; j < n_jobs; ++j) {
                const V4Params& p = b.job[j];
                if (j > 0 && !p.soft) CTA_ROLE_BARRIER();
                if (j > 0) {   // every MMA of job j - 1 has completed: the buffer it read may be overwritten
                    mbar_wait_relaxed(&bar_wempty, ph_wempty & 1u);
                    ph_wempty ^= 1u;
                }
                if (!p.prestaged) {
                    stage_b_operand<1>(stage_args(p), base + static_cast<uint32_t>(p.wsel) * p.w_pair, 256 + ts, 256 + kStagers);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_w0full);
                }
                if (j + 1 < n_jobs && b.job[j + 1].prestaged) {
                    stage_b_operand<1>(stage_args(b.job[j + 1]), base + static_cast<uint32_t>(b.job[j + 1].wsel) * p.w_pair, ts, kStagers);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_wfull);
                }
            }
        }
        for (int j = 0; j < n_jobs; ++j) {
            if (warp != kWarpMma) break;
            const V4Params& p = b.job[j];
            if (j > 0 && !p.soft) CTA_ROLE_BARRIER();
            // =============================== MMA issuer ===============================
            if (!p.prestaged) {   // B operand staged at this job's start (epilogue + stager warps)
                if (lane == 0) mbar_wait_relaxed(&bar_w0full, ph_w0full & 1u);   // one polling lane, with back-off: the stagers share its scheduler
                __syncwarp();
                ph_w0full ^= 1u;
            } else {   // staged into the other buffer while the previous job ran
                if (lane == 0) mbar_wait_relaxed(&bar_wfull, ph_wfull & 1u);
                __syncwarp();
                ph_wfull ^= 1u;
            }
            tc_fence_after_sync();
            const int Kp = p.Kp, Np = p.Np;
            const TileRange tr = cta_range_of(p);
            const int n_tiles = tr.n_tiles;
            const uint32_t idesc = umma_idesc_tf32(128, Np);
            const uint32_t wb = base + static_cast<uint32_t>(p.wsel) * p.w_pair;
            const uint64_t dwhi = umma_desc_sw128(wb + p.off_whi), dwlo = umma_desc_sw128(wb + p.off_wlo);
            const uint32_t w_atom16 = p.w_atom >> 4;
            const int ks = Kp >> 3;
            int zi = 0, ai = 0;
            for (int it = 0; it < n_tiles; ++it) {
                mbar_wait(&bar_zfull[zi], (ph_zfull >> zi) & 1u);
                ph_zfull ^= 1u << zi;
                mbar_wait(&bar_tempty[ai], ((ph_tempty >> ai) & 1u) ^ 1u);
                ph_tempty ^= 1u << ai;
                tc_fence_after_sync();
                const uint32_t d = tmem + static_cast<uint32_t>(ai * Np);
                const uint32_t zhi = tmem + p.tm_z + static_cast<uint32_t>(zi * 2 * Kp), zlo = zhi + static_cast<uint32_t>(Kp);
                __syncwarp();
                if (elect_one()) {
                    switch (ks) {
                        case 5: issue_tile<5>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                        case 9: issue_tile<9>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                        case 13: issue_tile<13>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                        case 17: issue_tile<17>(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16); break;
                        default: issue_tile_loop(d, zhi, zlo, dwhi, dwlo, idesc, w_atom16, ks);
                    }
                    umma_commit(&bar_zempty[zi]);
                    umma_commit(&bar_tfull[ai]);
                }
                __syncwarp();
                if (it == 0) V4_STAMP(4);
                if (it == n_tiles - 1) V4_STAMP(5);
                if (++zi == p.zbufs) zi = 0;
                if (++ai == p.abufs) ai = 0;
            }
            if (elect_one()) umma_commit(&bar_wempty);   // arrives when every MMA of this job has completed
            __syncwarp();
        }
    } else {
        // =============================== epilogue warps ===============================
        reg_dec<kRegsEpi>();   // the registers the aggregation warps take come from here and from the MMA / TMA warpgroup
        uint32_t ph_tfull = 0;
        const int e = warp - kWarpEpi0;
        const int wq = e & 3, h = e >> 2;
        const uint32_t lane_sel = static_cast<uint32_t>(wq * 32) << 16;
        const uint32_t l7 = static_cast<uint32_t>(lane) & 7u;
        const int colq = static_cast<int>(l7) * 4;
        const int row0 = wq * 32 + (lane >> 3);
        const int te = tid - kWarpEpi0 * 32;   // 0 .. 255 inside the epilogue group
        if (b.early) pdl_wait();   // parameters (and every buffer this kernel writes) belong to the previous grid until it completes
        for (int j = 0; j < n_jobs; ++j) {
            const V4Params& p = b.job[j];
            if (j > 0 && !p.soft) CTA_ROLE_BARRIER();
            if (!p.prestaged) {   // (a prestaged job's operand was written by the stager warps during the previous job)
                stage_b_operand<0>(stage_args(p), base + static_cast<uint32_t>(p.wsel) * p.w_pair, te, 256 + kStagers);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_w0full);
            }
            if (e == 0) V4_STAMP(10);
            const int N = p.N, f_out = p.f_out, Np = p.Np;
            const TileRange tr = cta_range_of(p);
            const int n_tiles = tr.n_tiles;
            const int last_ng = tr.n_graphs_cta - (n_tiles - 1) * p.G;
            const uint32_t ys = base + p.off_ystage + static_cast<uint32_t>(e) * 4096u;
            const int n_cslabs = (Np + 31) >> 5;
            const int full_rows = p.G * N;
            const uint32_t yrow = ys + static_cast<uint32_t>(lane) * 128u;
            const uint32_t ysrc = ys + (static_cast<uint32_t>(lane) >> 3) * 128u;
            const size_t y_ld = static_cast<size_t>(p.y_ld);
            float* y_tile = p.y + (tr.g_begin * N + row0) * y_ld + colq;
            const size_t y_step = static_cast<size_t>(full_rows) * y_ld;
            const bool head = p.head != 0;
            const bool mul = p.mul_src != nullptr;
            const bool accin = p.acc_in != 0;
            // on-chip hand-over to the next job: its stage slot `it` takes this job's output tile `it` (pitch = f_out floats)
            const bool handover = j + 1 < n_jobs && b.job[j + 1].x_prev && n_tiles <= b.job[j + 1].n_stages;
            const uint32_t ho_pitch = static_cast<uint32_t>(f_out) * 4u;
            const uint32_t ho_lane = static_cast<uint32_t>(row0) * ho_pitch + static_cast<uint32_t>(colq) * 4u;   // + slot + slab + 4 k rows
            // ---- head state (training step, last forward job) ----
            const int L = p.n_labels, Fs = f_out;
            const uint32_t hs_gsum = base + p.off_head;                                            // [2 tiles][4][G][Fs]
            const uint32_t hs_dg = hs_gsum + static_cast<uint32_t>(2 * 4 * p.G * Fs) * 4u;         // [2][G][Fs]
            const uint32_t hs_wd = hs_dg + static_cast<uint32_t>(2 * p.G * Fs) * 4u;               // [Fs][L] + bias [4]
            const uint32_t hs_hp = hs_wd + static_cast<uint32_t>(Fs * L + 4) * 4u;                 // [8][Fs * L + 8] per-warp sums
            const uint32_t hs_zs = hs_hp + static_cast<uint32_t>(kEpiWarps * (Fs * L + 8)) * 4u;   // [8][4] d logits of the warp's graph
            const uint32_t my_hp = hs_hp + static_cast<uint32_t>(e * (Fs * L + 8)) * 4u;
            if (j == b.head_setup_job) {
                // the head's Dense weights and zeroed accumulators, placed at the start of the soft run that ends in the head job
                // (same plan, so the head's shared memory is reserved and untouched): two cold global round trips that the head
                // job itself would otherwise wait for
                const V4Params& hp = b.job[b.head_job];
                const int hL = hp.n_labels, hF = hp.f_out;
                const uint32_t h_wd = base + hp.off_head + static_cast<uint32_t>(2 * 5 * hp.G * hF) * 4u;
                const uint32_t h_hp = h_wd + static_cast<uint32_t>(hF * hL + 4) * 4u;
                for (int i = te; i < hF * hL + 4; i += 256) {
                    const int bi = i - hF * hL;
                    const float wv[1] = {bi < 0 ? __ldg(hp.head_w + i) : ((bi < hL && hp.head_b != nullptr) ? __ldg(hp.head_b + bi) : 0.0f)};
                    sts_f<1>(h_wd + 4u * static_cast<uint32_t>(i), wv);
                }
                const float z1[1] = {0.0f};
                for (int i = te; i < kEpiWarps * (hF * hL + 8); i += 256) sts_f<1>(h_hp + 4u * static_cast<uint32_t>(i), z1);
                asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
            }
            int ai = 0;
            if (!head) {
                for (int it = 0; it < n_tiles; ++it) {
                    const int ng = (it == n_tiles - 1) ? last_ng : p.G;
                    const int rows = ng * N;
                    mbar_wait_relaxed(&bar_tfull[ai], (ph_tfull >> ai) & 1u);
                    ph_tfull ^= 1u << ai;
                    tc_fence_after_sync();
                    if (e == 0 && it == 0) V4_STAMP(6);
                    const uint32_t ta = tmem + lane_sel + static_cast<uint32_t>(ai * Np);
                    const uint32_t ho_slot = handover ? base + b.job[j + 1].off_stage + static_cast<uint32_t>(it) * b.job[j + 1].stage_bytes : 0u;
                    for (int cs = h; cs < n_cslabs; cs += 2) {
                        float v0[16], v1[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v1[i] = 0.0f;
                        tmem_ld16(ta + static_cast<uint32_t>(cs * 32), v0);
                        if (cs * 32 + 16 < Np) tmem_ld16(ta + static_cast<uint32_t>(cs * 32 + 16), v1);
                        tmem_ld_wait();
                        tmem_ld_fence(v0);
                        tmem_ld_fence(v1);
                        if (cs + 2 >= n_cslabs) {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_tempty[ai]);
                        }
                        if (!accin) {
                            act16_sel<ACT>(v0, p.act);
                            act16_sel<ACT>(v1, p.act);
                            if (cs * 32 + 32 > p.f_valid) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (cs * 32 + i >= p.f_valid) v0[i] = 0.0f;
                                    if (cs * 32 + 16 + i >= p.f_valid) v1[i] = 0.0f;
                                }
                            }
                        }
                        tmem_ld_fence(v0);
                        tmem_ld_fence(v1);
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const float t0[4] = {v0[4 * c4], v0[4 * c4 + 1], v0[4 * c4 + 2], v0[4 * c4 + 3]};
                            const float t1[4] = {v1[4 * c4], v1[4 * c4 + 1], v1[4 * c4 + 2], v1[4 * c4 + 3]};
                            sts_f<4>(yrow + ((static_cast<uint32_t>(c4) ^ l7) << 4), t0);
                            sts_f<4>(yrow + ((static_cast<uint32_t>(c4 + 4) ^ l7) << 4), t1);
                        }
                        __syncwarp();
                        const bool col_ok = cs * 32 + colq < f_out;
                        float* ycs = y_tile + cs * 32;
                        if (accin) {   // later channel group of a layer: add the partial pre-activation already in y, then activate
                            const int cbase = cs * 32 + colq;
#pragma unroll 2
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                                float t[4];
                                lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                                if (row0 + 4 * k < rows && col_ok) {
                                    float4* dst = reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld);
                                    const float4 o = *dst;
                                    t[0] += o.x; t[1] += o.y; t[2] += o.z; t[3] += o.w;
#pragma unroll
                                    for (int jj = 0; jj < 4; ++jj) {
                                        t[jj] = act1_sel<ACT>(t[jj], p.act);
                                        if (cbase + jj >= p.f_valid) t[jj] = 0.0f;
                                    }
                                    *dst = make_float4(t[0], t[1], t[2], t[3]);
                                    if (handover) sts_f<4>(ho_slot + ho_lane + static_cast<uint32_t>(cs) * 128u + static_cast<uint32_t>(4 * k) * ho_pitch, t);
                                }
                            }
                        } else if (!mul) {
#pragma unroll 2   // (rolled: this code runs once or twice per launch, its instruction fetch costs more than the loop overhead)
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                                float t[4];
                                lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                                if (row0 + 4 * k < rows && col_ok) {
                                    *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) = make_float4(t[0], t[1], t[2], t[3]);
                                    if (handover) sts_f<4>(ho_slot + ho_lane + static_cast<uint32_t>(cs) * 128u + static_cast<uint32_t>(4 * k) * ho_pitch, t);
                                }
                            }
                        } else {
                            const float* mcs = p.mul_src + (ycs - p.y);   // same [rows, y_ld] layout as the output
                            float4 mv[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                mv[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                                if (row0 + 4 * k < rows && col_ok) mv[k] = __ldg(reinterpret_cast<const float4*>(mcs + static_cast<size_t>(4 * k) * y_ld));
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                                float t[4];
                                lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                                const float yv[4] = {mv[k].x, mv[k].y, mv[k].z, mv[k].w};
                                mul_act_grad4_sel<ACT>(t, yv, p.mul_act);
                                if (row0 + 4 * k < rows && col_ok) {
                                    *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) = make_float4(t[0], t[1], t[2], t[3]);
                                    if (handover) sts_f<4>(ho_slot + ho_lane + static_cast<uint32_t>(cs) * 128u + static_cast<uint32_t>(4 * k) * ho_pitch, t);
                                }
                            }
                        }
                        __syncwarp();
                    }
                    if (h >= n_cslabs) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_tempty[ai]);
                    }

                    PUBLISH_TILE();
                }
            } else {
                // ======== fused head.  This warp owns tile rows 32 wq .. 32 wq + 31 and columns 32 h .. 32 h + 31.  Tiles are taken
                // in groups: the whole job when its tiles fit the accumulators (the latency-bound case: (b) and both barriers are
                // paid once per job), else one tile at a time.  H stays in tensor memory between (a) and (c). ========
                const int cs = h;
                const bool has_cols = cs < n_cslabs;
                const int gmax = n_tiles <= p.abufs ? n_tiles : 1;
                const uint32_t gs_stride = static_cast<uint32_t>(4 * p.G * Fs) * 4u, dg_stride = static_cast<uint32_t>(p.G * Fs) * 4u;
                for (int it0 = 0; it0 < n_tiles; it0 += gmax) {
                    const int gsz = min(gmax, n_tiles - it0);
                    const int64_t g0_grp = tr.g_begin + static_cast<int64_t>(it0) * p.G;
                    const int ng_grp = min(gsz * p.G, tr.n_graphs_cta - it0 * p.G);   // graphs of the group
                    // label l (lane l) and mask of the graph this warp will finish (cold HBM lines): requested now, needed after (a)
                    float y_lane = 0.0f, m0 = 1.0f;
                    if (e < ng_grp) {
                        if (lane < L) y_lane = __ldg(p.labels + (g0_grp + e) * L + lane);
                        if (p.mask) m0 = __ldg(p.mask + g0_grp + e);
                    }
                    // (a) per tile: activation, per-graph column sums over this warp's rows (four independent partial sums per lane)
                    int aa = ai;
                    for (int u = 0; u < gsz; ++u) {
                        const int it = it0 + u;
                        const int rows = ((it == n_tiles - 1) ? last_ng : p.G) * N;
                        mbar_wait_relaxed(&bar_tfull[aa], (ph_tfull >> aa) & 1u);
                        ph_tfull ^= 1u << aa;
                        tc_fence_after_sync();
                        if (e == 0 && it == 0) V4_STAMP(6);
                        if (has_cols) {
                            const uint32_t ta = tmem + lane_sel + static_cast<uint32_t>(aa * Np);
                            float v0[16], v1[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v1[i] = 0.0f;
                            tmem_ld16(ta + static_cast<uint32_t>(cs * 32), v0);
                            if (cs * 32 + 16 < Np) tmem_ld16(ta + static_cast<uint32_t>(cs * 32 + 16), v1);
                            tmem_ld_wait();
                            tmem_ld_fence(v0);
                            tmem_ld_fence(v1);
                            act16_sel<ACT>(v0, p.act);
                            act16_sel<ACT>(v1, p.act);
                            if (cs * 32 + 32 > p.f_valid) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (cs * 32 + i >= p.f_valid) v0[i] = 0.0f;
                                    if (cs * 32 + 16 + i >= p.f_valid) v1[i] = 0.0f;
                                }
                            }
                            tmem_ld_fence(v0);
                            tmem_ld_fence(v1);
#pragma unroll
                            for (int c4 = 0; c4 < 4; ++c4) {
                                const float t0[4] = {v0[4 * c4], v0[4 * c4 + 1], v0[4 * c4 + 2], v0[4 * c4 + 3]};
                                const float t1[4] = {v1[4 * c4], v1[4 * c4 + 1], v1[4 * c4 + 2], v1[4 * c4 + 3]};
                                sts_f<4>(yrow + ((static_cast<uint32_t>(c4) ^ l7) << 4), t0);
                                sts_f<4>(yrow + ((static_cast<uint32_t>(c4 + 4) ^ l7) << 4), t1);
                            }
                            __syncwarp();
                            const int r_lo = wq * 32, r_hi = min(rows, r_lo + 32);
                            if (r_lo < r_hi) {
                                const int g_first = r_lo / N, g_last = (r_hi - 1) / N;
                                const uint32_t cbase = ys + ((static_cast<uint32_t>(lane) & 3u) << 2);
                                const uint32_t cch = static_cast<uint32_t>(lane) >> 2;   // lane c reads column c: chunk c >> 2 sits at (c >> 2) ^ (r & 7)
                                for (int g = g_first; g <= g_last; ++g) {
                                    const int ra = max(g * N, r_lo) - r_lo, rb = min(g * N + N, r_hi) - r_lo;
                                    float s4[4] = {0.f, 0.f, 0.f, 0.f};
                                    int r = ra;
                                    for (; r + 4 <= rb; r += 4) {
#pragma unroll
                                        for (int q4 = 0; q4 < 4; ++q4)
                                            s4[q4] += lds_f32(cbase + static_cast<uint32_t>(r + q4) * 128u + ((cch ^ (static_cast<uint32_t>(r + q4) & 7u)) << 4));
                                    }
                                    for (; r < rb; ++r) s4[0] += lds_f32(cbase + static_cast<uint32_t>(r) * 128u + ((cch ^ (static_cast<uint32_t>(r) & 7u)) << 4));
                                    const float o[1] = {(s4[0] + s4[1]) + (s4[2] + s4[3])};
                                    sts_f<1>(hs_gsum + static_cast<uint32_t>(u) * gs_stride + 4u * static_cast<uint32_t>((wq * p.G + g) * Fs + cs * 32 + lane), o);
                                }
                            }
                            __syncwarp();   // the staging tile is reused by the next tile of the group
                        }
                        if (++aa == p.abufs) aa = 0;
                    }
                    if (e == 0 && it0 == 0) V4_STAMP(12);
                    asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                    if (e == 0 && it0 == 0) V4_STAMP(13);
                    // (b) one warp per graph of the group: GraphGather sum, logits, softmax cross-entropy, d logits, d gathered.  Lane l
                    // holds label l's quantities; loops over labels are rolled (this code runs once per group: keep it small)
                    for (int gi = e; gi < ng_grp; gi += kEpiWarps) {
                        const int u = gi / p.G, g = gi - u * p.G;
                        const int64_t bg = g0_grp + gi;
                        const uint32_t gsum_u = hs_gsum + static_cast<uint32_t>(u) * gs_stride;
                        const int q_lo = (g * N) >> 5, q_hi = (g * N + N - 1) >> 5;
                        float yl = y_lane, m = m0;
                        if (gi != e) {   // more than 8 graphs per group (small graphs): the later ones load their own
                            yl = lane < L ? __ldg(p.labels + bg * L + lane) : 0.0f;
                            m = p.mask ? __ldg(p.mask + bg) : 1.0f;
                        }
                        float gv0 = 0.0f, gv1 = 0.0f;
                        for (int q = q_lo; q <= q_hi; ++q) {
                            gv0 += lds_f32(gsum_u + 4u * static_cast<uint32_t>((q * p.G + g) * Fs + lane));
                            if (Fs > 32) gv1 += lds_f32(gsum_u + 4u * static_cast<uint32_t>((q * p.G + g) * Fs + lane + 32));
                        }
                        if (p.gathered != nullptr) {
                            p.gathered[bg * Fs + lane] = gv0;
                            if (Fs > 32) p.gathered[bg * Fs + lane + 32] = gv1;
                        }
                        float zl = -3.0e38f;
#pragma unroll 1
                        for (int l = 0; l < L; ++l) {
                            float acc = gv0 * lds_f32(hs_wd + 4u * static_cast<uint32_t>(lane * L + l));
                            if (Fs > 32) acc = fmaf(gv1, lds_f32(hs_wd + 4u * static_cast<uint32_t>((lane + 32) * L + l)), acc);
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
                        const int arg_p = __ffs(__ballot_sync(0xffffffffu, lane < L && zl == zmax)) - 1;   // first maximum wins
                        const int arg_y = __ffs(__ballot_sync(0xffffffffu, lane < L && yl == ymax)) - 1;
                        const float pr = lane < L ? __expf(zl - lse) : 0.0f;
                        const float dzl = m * p.inv_batch * (pr * ysum - yl);
                        if (lane < L) {
                            if (p.logits) p.logits[bg * L + lane] = zl;
                            if (p.prediction) p.prediction[bg * L + lane] = pr;
                            const float o[1] = {dzl};
                            sts_f<1>(hs_zs + 4u * static_cast<uint32_t>(e * 4 + lane), o);
                            const float bacc[1] = {lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + lane)) + dzl};
                            sts_f<1>(my_hp + 4u * static_cast<uint32_t>(Fs * L + lane), bacc);
                        }
                        if (lane == 0) {
                            const float c2[2] = {lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + 4)) + m * cost,
                                                 lds_f32(my_hp + 4u * static_cast<uint32_t>(Fs * L + 5)) + m * (arg_p == arg_y ? 1.0f : 0.0f)};
                            sts_f<2>(my_hp + 4u * static_cast<uint32_t>(Fs * L + 4), c2);
                        }
                        __syncwarp();
                        float dg0 = 0.0f, dg1 = 0.0f;
#pragma unroll 1
                        for (int l = 0; l < L; ++l) {   // d gathered and this warp's share of dW_dense (lane owns features lane, lane + 32)
                            const float dz = lds_f32(hs_zs + 4u * static_cast<uint32_t>(e * 4 + l));
                            const uint32_t i0 = 4u * static_cast<uint32_t>(lane * L + l), i1 = 4u * static_cast<uint32_t>((lane + 32) * L + l);
                            dg0 = fmaf(dz, lds_f32(hs_wd + i0), dg0);
                            const float a0[1] = {fmaf(gv0, dz, lds_f32(my_hp + i0))};
                            sts_f<1>(my_hp + i0, a0);
                            if (Fs > 32) {
                                dg1 = fmaf(dz, lds_f32(hs_wd + i1), dg1);
                                const float a1[1] = {fmaf(gv1, dz, lds_f32(my_hp + i1))};
                                sts_f<1>(my_hp + i1, a1);
                            }
                        }
                        const float o0[1] = {dg0}, o1[1] = {dg1};
                        sts_f<1>(hs_dg + static_cast<uint32_t>(u) * dg_stride + 4u * static_cast<uint32_t>(g * Fs + lane), o0);
                        if (Fs > 32) sts_f<1>(hs_dg + static_cast<uint32_t>(u) * dg_stride + 4u * static_cast<uint32_t>(g * Fs + lane + 32), o1);
                        __syncwarp();
                    }
                    if (e == 0 && it0 == 0) V4_STAMP(14);
                    asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                    if (e == 0 && it0 == 0) V4_STAMP(15);
                    // (c) per tile: dU = dg[graph of the row] (.) act'(H) with H re-read from tensor memory (the accumulator is handed
                    // back here), transposed through the warp's staging tile; the tile leaves like an activation tile
                    for (int u = 0; u < gsz; ++u) {
                        const int it = it0 + u;
                        const int rows = ((it == n_tiles - 1) ? last_ng : p.G) * N;
                        const uint32_t ho_slot = handover ? base + b.job[j + 1].off_stage + static_cast<uint32_t>(it) * b.job[j + 1].stage_bytes : 0u;
                        if (has_cols) {
                            const uint32_t ta = tmem + lane_sel + static_cast<uint32_t>(ai * Np);
                            float v0[16], v1[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) v1[i] = 0.0f;
                            tmem_ld16(ta + static_cast<uint32_t>(cs * 32), v0);
                            if (cs * 32 + 16 < Np) tmem_ld16(ta + static_cast<uint32_t>(cs * 32 + 16), v1);
                            tmem_ld_wait();
                            tmem_ld_fence(v0);
                            tmem_ld_fence(v1);
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_tempty[ai]);       // the accumulator is in registers: hand it back
                            act16_sel<ACT>(v0, p.act);
                            act16_sel<ACT>(v1, p.act);
                            if (cs * 32 + 32 > p.f_valid) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    if (cs * 32 + i >= p.f_valid) v0[i] = 0.0f;
                                    if (cs * 32 + 16 + i >= p.f_valid) v1[i] = 0.0f;
                                }
                            }
                            tmem_ld_fence(v0);
                            tmem_ld_fence(v1);
                            const int r = wq * 32 + lane;
                            const int g = (r < rows) ? r / N : 0;
                            const uint32_t dga = hs_dg + static_cast<uint32_t>(u) * dg_stride + 4u * static_cast<uint32_t>(g * Fs + cs * 32);
#pragma unroll
                            for (int c4 = 0; c4 < 4; ++c4) {
                                float d0[4], d1[4];
                                lds_f<4>(d0, dga + 16u * static_cast<uint32_t>(c4));
                                lds_f<4>(d1, dga + 16u * static_cast<uint32_t>(c4 + 4));
                                const float h0[4] = {v0[4 * c4], v0[4 * c4 + 1], v0[4 * c4 + 2], v0[4 * c4 + 3]};
                                const float h1[4] = {v1[4 * c4], v1[4 * c4 + 1], v1[4 * c4 + 2], v1[4 * c4 + 3]};
                                mul_act_grad4_sel<ACT>(d0, h0, p.act);
                                mul_act_grad4_sel<ACT>(d1, h1, p.act);
                                sts_f<4>(yrow + ((static_cast<uint32_t>(c4) ^ l7) << 4), d0);
                                sts_f<4>(yrow + ((static_cast<uint32_t>(c4 + 4) ^ l7) << 4), d1);
                            }
                            __syncwarp();
                            const bool col_ok = cs * 32 + colq < f_out;
                            float* ycs = y_tile + cs * 32;
#pragma unroll 2
                            for (int k = 0; k < 8; ++k) {
                                const uint32_t r7 = (static_cast<uint32_t>(4 * k) + (static_cast<uint32_t>(lane) >> 3)) & 7u;
                                float t[4];
                                lds_f<4>(t, ysrc + static_cast<uint32_t>(k) * 512u + ((l7 ^ r7) << 4));
                                if (row0 + 4 * k < rows && col_ok) {
                                    *reinterpret_cast<float4*>(ycs + static_cast<size_t>(4 * k) * y_ld) = make_float4(t[0], t[1], t[2], t[3]);
                                    if (handover) sts_f<4>(ho_slot + ho_lane + static_cast<uint32_t>(cs) * 128u + static_cast<uint32_t>(4 * k) * ho_pitch, t);
                                }
                            }
                            __syncwarp();
                        } else {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bar_tempty[ai]);
                        }
                        PUBLISH_TILE();
                    }
                }
            }
            if (head) {
                // per-CTA partial of the head's parameter gradients and statistics: the 8 warps' sums, added in warp order
                asm volatile("bar.sync 4, %0;" ::"n"(256) : "memory");
                float* hp_out = p.head_partial + static_cast<size_t>(blockIdx.x) * (Fs * L + 8);
                for (int i = te; i < Fs * L + 8; i += 256) {
                    float acc = 0.0f;
                    for (int w8 = 0; w8 < kEpiWarps; ++w8) acc += lds_f32(hs_hp + 4u * static_cast<uint32_t>(w8 * (Fs * L + 8) + i));
                    hp_out[i] = acc;
                }
            }
            // this CTA's outputs of the job: visible to the next job's TMA loads (async proxy) after the role barrier
            if (j + 1 < n_jobs && !p.soft_next) {
                __threadfence();
                fence_proxy_async_all();
            }
            if (e == 0) V4_STAMP(11);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (b.dbg != nullptr && tid == 0) b.dbg[static_cast<size_t>(blockIdx.x) * 128 + 127] = clock64();
    if (warp == kWarpMma) tmem_dealloc(tmem, 512);
}

inline uint32_t up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

constexpr int kSmemMax = 227 * 1024 - 1024;   // static __shared__ (barriers) shares the 227 KB

// One candidate plan: `n_split` output-column slices of f_out_total / n_split columns (each slice is its own CTA row of
// the grid with its own [W ; bias] slice resident in shared memory), G graphs per tile.
uint32_t head_smem_bytes(int G, int f_out, int n_labels) {
    return static_cast<uint32_t>(2 * 5 * G * f_out + f_out * n_labels + 4 + kEpiWarps * (f_out * n_labels + 8) + kEpiWarps * 4) * 4u;   // gsum / dg of up to 2 tiles (a group)
}

bool plan_v4_try(V4Params& p, int64_t n_graphs, int C, int N, int f_in, int f_out_total, int n_split, int G, int C_csr,
                 int head_labels = 0, int wbufs = 1) {
    if (f_out_total % n_split != 0) return false;
    const int f_out = f_out_total / n_split;
    if (f_out % 4 != 0 || f_out > 256 || (n_split > 1 && f_out % 32 != 0)) return false;
    p.C = C; p.N = N; p.f_in = f_in; p.f_out = f_out; p.n_graphs = n_graphs; p.n_split = n_split;
    p.K = C * f_in;
    p.Kp = p.K + 8;
    p.Np = static_cast<int>(up(f_out, 16));
    p.n_slabs = p.K / 32;
    p.slabs_per_ch = f_in / 32;
    // TMEM: abufs accumulators + zbufs x (Zhi | Zlo)
    if (2 * p.Np + 4 * p.Kp <= 512) { p.abufs = 2; p.zbufs = 2; }
    else if (2 * p.Np + 2 * p.Kp <= 512) { p.abufs = 2; p.zbufs = 1; }
    else if (p.Np + 2 * p.Kp <= 512) { p.abufs = 1; p.zbufs = 1; }
    else return false;
    p.tm_z = static_cast<uint32_t>(p.abufs * p.Np);
    p.G = G;
    // contiguous graph ranges, one per CTA of a slice; small batches are spread over all SMs
    const int64_t grid0 = std::min<int64_t>(std::max(1, kNumSMs / n_split), n_graphs);
    const int64_t gpc = ceil_div<int64_t>(n_graphs, grid0);
    if (gpc > (1 << 24)) return false;
    p.graphs_per_cta = static_cast<int>(gpc);
    if (p.G > p.graphs_per_cta) p.G = p.graphs_per_cta;
    const uint32_t rows_max = static_cast<uint32_t>(p.G) * N;
    const uint32_t n_watoms = (static_cast<uint32_t>(p.Kp) + 31) / 32;
    p.w_atom = static_cast<uint32_t>(p.Np) * 128u;
    uint32_t off = 0;
    p.off_whi = off; off += n_watoms * p.w_atom;
    p.off_wlo = off; off += n_watoms * p.w_atom;
    p.w_pair = off;
    p.wbufs = wbufs; p.wsel = 0; p.prestaged = 0; p.soft = 0; p.soft_next = 0; p.x_prev = 0;
    off += static_cast<uint32_t>(wbufs - 1) * p.w_pair;
    p.off_ystage = off; off += kEpiWarps * 4096u;
    p.off_head = off;
    if (head_labels > 0) {
        if (n_split != 1 || f_out > 64 || f_out % 32 != 0 || head_labels > 4 || p.G > 16) return false;
        off += (head_smem_bytes(p.G, f_out, head_labels) + 127u) & ~127u;
    }
    p.off_stage = off;
    p.C_csr = C_csr;
    p.cv_cap = static_cast<int>(up(std::max<uint32_t>(256, 6 * rows_max * C_csr), 4));
    p.st_rp = up(rows_max * f_in * 4u, 128);
    p.st_col = p.st_rp + up((rows_max * C_csr + 8) * 4u, 16);
    p.st_val = p.st_col + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u;
    p.stage_bytes = up(p.st_val + (static_cast<uint32_t>(p.cv_cap) + 4) * 4u, 128);
    p.n_stages = 0;
    for (int st = kV4MaxStages; st >= 2; --st)
        if (off + st * p.stage_bytes + 1024 <= static_cast<uint32_t>(kSmemMax)) { p.n_stages = st; break; }
    if (p.n_stages == 0) return false;
    p.smem_total = off + p.n_stages * p.stage_bytes + 1024;
    return true;
}

// Preference: whole output width in one CTA and full 128-row tiles; wide layers (F = 128: [W ; bias] hi / lo alone is
// 160 KB) fall back to column slices -- each slice aggregates the tile again (the second reader hits L2) -- and to
// fewer graphs per tile until at least two TMA stages fit.
int g_cap() {
    static const int cap = [] {
        const char* e = getenv("KGCN_V4_G");   // tuning knob: at most this many graphs per tile
        return (e != nullptr && atoi(e) > 0) ? atoi(e) : 1 << 20;
    }();
    return cap;
}
bool plan_v4(V4Params& p, int64_t n_graphs, int C, int N, int f_in, int f_out, int C_csr, int head_labels = 0, int wbufs = 1) {
    if (n_graphs <= 0 || N > 128 || N < 1 || f_in % 32 != 0 || f_out % 4 != 0 || C_csr > 8 || C < 1 || C > C_csr) return false;
    for (int n_split = 1; n_split <= 4; n_split *= 2)
        for (int G = std::min(g_cap(), std::max(1, 128 / N)); G >= 1; G = (G > 1 ? G / 2 : 0))
            if (plan_v4_try(p, n_graphs, C, N, f_in, f_out, n_split, G, C_csr, head_labels, wbufs)) return true;
    return false;
}

// A layer whose K = C * f_in does not fit tensor memory (three bond types x 96 features) is run as several launches over
// channel groups: the first writes the partial pre-activation, the following ones add to it and the last activates.
// Returns the largest group size that has a plan (0: none).
int plan_v4_group(V4Params& p, int64_t n_graphs, int C, int N, int f_in, int f_out) {
    for (int cg = C; cg >= 1; --cg)
        if (plan_v4(p, n_graphs, cg, N, f_in, f_out, C)) return (cg == C || p.n_split == 1) ? cg : 0;
    return 0;
}

}  // namespace

bool fused_v4_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_FUSED_V4");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

bool fused_v4_eligible(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, const float* x, const float* y,
                       const int32_t* rowptr, const int32_t* col, const float* val, const float* w, const float* bias) {
    V4Params p{};
    if (!fused_v4_enabled() || plan_v4_group(p, n_graphs, channels, n_nodes, f_in, f_out) == 0) return false;
    return aligned16(x) && aligned16(y) && aligned16(rowptr) && aligned16(col) && aligned16(val) && aligned16(w) && aligned16(bias) &&
           n_graphs * static_cast<int64_t>(n_nodes) * channels < (1ll << 31);
}

bool fused_v4_plannable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    V4Params p{};
    return fused_v4_enabled() && plan_v4_group(p, n_graphs, channels, n_nodes, f_in, f_out) > 0;
}

static bool handover_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_CHAIN_HANDOVER");   // A-B knob: 0 = soft jobs always reload their input through HBM / L2
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}
static bool wbufs2_enabled() {
    static const bool on = [] {
        const char* e = getenv("KGCN_CHAIN_WBUFS");   // A-B knob: 1 keeps a single B operand buffer in chained launches
        return e == nullptr || atoi(e) != 1;
    }();
    return on;
}

static long long* g_dbg_v4 = nullptr;
static long long* g_dbg_chain = nullptr;

int launch_graphconv_fused_v4(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs, int channels,
                              int n_nodes, const float* x, int f_in, const float* w, const float* bias, int f_out, int act,
                              float* y, cudaStream_t st, bool w_transposed, const float* mul_src, int mul_act, int f_out_valid) {
    V4Params p0{};
    const int cg = plan_v4_group(p0, n_graphs, channels, n_nodes, f_in, f_out);
    KGCN_REQUIRE(cg > 0, KGCN_ERR_UNSUPPORTED, "fused GraphConv v4: unsupported shape");
    const bool want_mul = mul_act != KGCN_ACT_NONE && mul_src != nullptr;
    for (int c0 = 0; c0 < channels; c0 += cg) {
        const int cn = std::min(cg, channels - c0);
        const bool first = c0 == 0, last = c0 + cn == channels;
        KGCN_REQUIRE(!(want_mul && !(first && last)), KGCN_ERR_UNSUPPORTED, "fused GraphConv v4: act' epilogue with channel groups");
        V4Params p{};
        KGCN_REQUIRE(plan_v4(p, n_graphs, cn, n_nodes, f_in, f_out, channels), KGCN_ERR_UNSUPPORTED, "fused GraphConv v4: unsupported shape");
        p.c_begin = c0;
        p.rowptr = rowptr; p.col = col; p.val = val; p.x = x; p.y = y;
        p.act = last ? act : KGCN_ACT_NONE;
        p.acc_in = first ? 0 : 1;
        p.y_ld = f_out;
        p.w_trans = w_transposed ? 1 : 0;
        // forward: w[c][f_in][f_out];  transposed (backward dx): the layer's w[c][f_out (= this call's n)][f_in (= this call's k)]
        p.w_ld = w_transposed ? f_in : f_out;
        p.w_cstride = f_in * f_out;
        p.w = w + static_cast<size_t>(c0) * p.w_cstride;
        p.bias = (w_transposed || bias == nullptr) ? nullptr : bias + static_cast<size_t>(c0) * f_out;
        // columns >= f_valid are written as exact zeros by the launch that activates (feature padding stays inert); only
        // supported without column slices (F = 128 layers are never padded)
        p.f_valid = (last && f_out_valid > 0 && f_out_valid < f_out) ? f_out_valid : p.f_out;
        KGCN_REQUIRE(p.f_valid == p.f_out || p.n_split == 1, KGCN_ERR_UNSUPPORTED, "fused GraphConv v4: padded outputs with column slices");
        p.mul_src = want_mul ? mul_src : nullptr;
        p.mul_act = mul_act;
        p.dbg = g_dbg_v4;
        const dim3 grid(static_cast<unsigned>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta)), static_cast<unsigned>(p.n_split));
        auto go = [&](auto kernel) -> int {
            KGCN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem_total)));
            launch_pdl(kernel, grid, kBlock, p.smem_total, st, p);
            KGCN_LAUNCH_OK("graphconv_fused_v4_kernel");
            return KGCN_OK;
        };
        const int rc = p.acc_in ? go(graphconv_fused_v4_kernel<2>) : (p.mul_src != nullptr ? go(graphconv_fused_v4_kernel<1>) : go(graphconv_fused_v4_kernel<0>));
        if (rc) return rc;
    }
    return KGCN_OK;
}

// ---- chained launch: jobs[k] over the same (n_graphs, channels, n_nodes); job k + 1 usually reads job k's output ----
bool fused_v4_chainable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    V4Params p{};
    return fused_v4_enabled() && plan_v4(p, n_graphs, channels, n_nodes, f_in, f_out, channels) && p.n_split == 1;
}

int launch_graphconv_fused_v4_chain(const V4ChainJob* jobs, int n_jobs, int64_t n_graphs, int channels, int n_nodes, cudaStream_t st,
                                    bool early_inputs) {
    KGCN_REQUIRE(n_jobs >= 1 && n_jobs <= kV4MaxJobs, KGCN_ERR_BAD_SHAPE, "fused GraphConv chain: 1..%d jobs", kV4MaxJobs);
    V4Batch b{};
    b.n_jobs = n_jobs;
    b.dbg = g_dbg_chain;
    b.early = early_inputs ? 1 : 0;
    uint32_t smem = 0;
    {
        const char* e = getenv("KGCN_CHAIN_TILE_FENCE");
        b.tile_fence_gpu = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    int head_k = -1;
    for (int k = 0; k < n_jobs; ++k)
        if (jobs[k].head != nullptr) head_k = k;
    auto cn_of = [&](const V4ChainJob& j) { return j.c_count > 0 ? j.c_count : channels; };
    int labels_prev = 0, n_hard = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const V4ChainJob& j = jobs[k];
        V4Params& p = b.job[k];
        const int own_labels = j.head != nullptr ? j.head->n_labels : 0;
        const int cn = cn_of(j);
        KGCN_REQUIRE(j.c_begin >= 0 && j.c_begin + cn <= channels, KGCN_ERR_BAD_SHAPE, "fused GraphConv chain: bad channel group");
        KGCN_REQUIRE(plan_v4(p, n_graphs, cn, n_nodes, j.f_in, j.f_out, channels, own_labels) && p.n_split == 1, KGCN_ERR_UNSUPPORTED,
                     "fused GraphConv chain: job %d (%d -> %d) has no single-CTA plan", k, j.f_in, j.f_out);
        // Richer layouts, taken only when they cost neither tile height nor more than the stages a short job needs: (1) the
        // head's shared memory also in the other jobs of the head job's shape, so that their plans are identical and the
        // boundaries between them can be soft; (2) a second B operand buffer, so that a soft job's [W ; bias] is staged
        // during its predecessor (latency-bound regime only: few tiles per CTA and job).
        int labels = own_labels;
        if (soft_jobs_enabled()) {
            const bool head_shape = head_k >= 0 && cn == cn_of(jobs[head_k]) && j.f_in == jobs[head_k].f_in && j.f_out == jobs[head_k].f_out;
            const int want_labels = head_shape ? jobs[head_k].head->n_labels : own_labels;
            const int tiles_per_cta = (p.graphs_per_cta + p.G - 1) / p.G;
            V4Params q{};
            for (int wb = (tiles_per_cta <= 4 && wbufs2_enabled() ? 2 : 1); wb >= 1; --wb) {
                if ((wb > 1 || want_labels != own_labels) && plan_v4(q, n_graphs, cn, n_nodes, j.f_in, j.f_out, channels, want_labels, wb) &&
                    q.n_split == 1 && q.G == p.G && q.abufs == p.abufs && q.zbufs == p.zbufs && q.n_stages >= std::min(p.n_stages, 2)) {
                    p = q;
                    labels = want_labels;
                    break;
                }
            }
        }
        KGCN_REQUIRE(p.graphs_per_cta == b.job[0].graphs_per_cta, KGCN_ERR_UNSUPPORTED, "fused GraphConv chain: graph ranges differ");
        KGCN_REQUIRE(!(j.acc_in && (j.mul_src != nullptr || j.head != nullptr)), KGCN_ERR_UNSUPPORTED, "fused GraphConv chain: accumulate + other epilogue");
        p.c_begin = j.c_begin;
        p.rowptr = j.rowptr; p.col = j.col; p.val = j.val; p.x = j.x; p.y = j.y;
        p.act = j.act;
        p.acc_in = j.acc_in;
        p.y_ld = j.f_out;
        p.w_trans = j.w_transposed ? 1 : 0;
        p.w_ld = j.w_transposed ? j.f_in : j.f_out;
        p.w_cstride = j.f_in * j.f_out;
        p.w = j.w + static_cast<size_t>(j.c_begin) * p.w_cstride;
        p.bias = (j.w_transposed || j.bias == nullptr) ? nullptr : j.bias + static_cast<size_t>(j.c_begin) * j.f_out;
        p.f_valid = (j.f_out_valid > 0 && j.f_out_valid < j.f_out) ? j.f_out_valid : p.f_out;
        p.mul_src = j.mul_src;
        p.mul_act = j.mul_act;
        KGCN_REQUIRE(j.zsave == nullptr || (cn == channels && (reinterpret_cast<uintptr_t>(j.zsave) & 127u) == 0 && (p.K & 31) == 0), KGCN_ERR_UNSUPPORTED,
                     "fused GraphConv chain: the aggregate can only be stored for a whole-layer job into a 128-byte aligned buffer");
        p.zsave = j.zsave;
        p.head = 0;
        if (j.head != nullptr) {
            const V4Head& hd = *j.head;
            KGCN_REQUIRE(j.mul_src == nullptr && hd.w && hd.labels && hd.partial && hd.n_labels >= 1 && hd.n_labels <= 4,
                         KGCN_ERR_BAD_SHAPE, "fused GraphConv chain: bad head (1..4 labels)");
            p.head = 1;
            p.n_labels = hd.n_labels;
            p.head_w = hd.w; p.head_b = hd.b; p.labels = hd.labels; p.mask = hd.mask; p.inv_batch = hd.inv_batch;
            p.logits = hd.logits; p.prediction = hd.prediction; p.gathered = hd.gathered; p.head_partial = hd.partial;
        }
        p.dbg = nullptr;
        p.soft = (k > 0 && soft_jobs_enabled() && cn == cn_of(jobs[k - 1]) && j.f_in == jobs[k - 1].f_in && j.f_out == jobs[k - 1].f_out &&
                  labels == labels_prev && p.G == b.job[k - 1].G && p.n_stages == b.job[k - 1].n_stages && p.off_stage == b.job[k - 1].off_stage &&
                  p.wbufs == b.job[k - 1].wbufs)
                     ? 1 : 0;
        p.soft_next = 0;
        if (p.soft) b.job[k - 1].soft_next = 1;
        p.wsel = (p.soft && p.wbufs == 2) ? (b.job[k - 1].wsel ^ 1) : 0;
        p.prestaged = (p.soft && p.wbufs == 2) ? 1 : 0;
        p.hard_ord = (k > 0 && !p.soft) ? ++n_hard : 0;
        p.x_prev = (p.soft && handover_enabled() && j.x == jobs[k - 1].y && j.f_in == jobs[k - 1].f_out) ? 1 : 0;
        labels_prev = labels;
        smem = std::max(smem, p.smem_total);
    }
    b.head_job = head_k;
    b.head_setup_job = head_k;
    while (b.head_setup_job > 0 && b.job[b.head_setup_job].soft) --b.head_setup_job;   // first job of the soft run (same plan)
    const unsigned grid = static_cast<unsigned>(ceil_div<int64_t>(n_graphs, b.job[0].graphs_per_cta));
    // the one non-trivial activation of the chain (forward act, act' of the dx jobs, the head's act'), else the generic kernel
    int chain_act = KGCN_ACT_NONE;
    for (int k = 0; k < n_jobs; ++k) {
        const V4Params& p = b.job[k];
        for (const int a : {p.act, p.mul_src != nullptr ? p.mul_act : KGCN_ACT_NONE}) {
            if (a == KGCN_ACT_NONE) continue;
            chain_act = (chain_act == KGCN_ACT_NONE || chain_act == a) ? a : -1;
        }
        if (chain_act == -1) break;
    }
    auto go = [&](auto kernel) -> int {
        KGCN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        launch_pdl(kernel, grid, kBlock, smem, st, b);
        KGCN_LAUNCH_OK("graphconv_fused_v4_chain_kernel");
        return KGCN_OK;
    };
    switch (chain_act) {
        case KGCN_ACT_NONE: return go(graphconv_fused_v4_chain_kernel<KGCN_ACT_NONE>);
        case KGCN_ACT_RELU: return go(graphconv_fused_v4_chain_kernel<KGCN_ACT_RELU>);
        case KGCN_ACT_SIGMOID: return go(graphconv_fused_v4_chain_kernel<KGCN_ACT_SIGMOID>);
        case KGCN_ACT_TANH: return go(graphconv_fused_v4_chain_kernel<KGCN_ACT_TANH>);
        default: return go(graphconv_fused_v4_chain_kernel<-1>);
    }
}

int fused_v4_chain_group(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int head_labels) {
    if (!fused_v4_enabled()) return 0;
    V4Params p{};
    for (int cg = channels; cg >= 1; --cg) {
        // the head rides on the LAST group's job; earlier groups are plain jobs (they need no head shared memory)
        if (plan_v4(p, n_graphs, cg, n_nodes, f_in, f_out, channels, head_labels) && p.n_split == 1) return cg;
        if (head_labels > 0) break;   // the fused head needs the whole layer in one job's epilogue state only on the last group: keep it simple
    }
    return 0;
}

bool fused_v4_head_chainable(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out, int n_labels) {
    V4Params p{};
    return fused_v4_enabled() && n_labels >= 1 && plan_v4(p, n_graphs, channels, n_nodes, f_in, f_out, channels, n_labels) && p.n_split == 1;
}

int fused_v4_chain_grid(int64_t n_graphs, int channels, int n_nodes, int f_in, int f_out) {
    V4Params p{};
    if (!plan_v4(p, n_graphs, channels, n_nodes, f_in, f_out, channels)) return 0;
    return static_cast<int>(ceil_div<int64_t>(n_graphs, p.graphs_per_cta));
}

}  // namespace kgcn

// Tuning hook (not part of the documented ABI): device buffer of [148][16] + [148][2] int64 phase cycle sums of the v4
// kernel (aggregation warp 0, MMA warp, epilogue warp 0, producer; see tools/phase_times_v4.py for the slot names).
extern "C" void kgcn_debug_v4_times(long long* device_buffer) { kgcn::g_dbg_v4 = device_buffer; }
// same for the chained kernel: [grid][128] clock64 stamps (slot = job * 16 + event, 126 = kernel entry, 127 = kernel exit)
extern "C" void kgcn_debug_v4_chain_times(long long* device_buffer) { kgcn::g_dbg_chain = device_buffer; }
