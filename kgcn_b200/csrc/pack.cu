// COO -> BatchedCSR ingest (include/kgcn_b200.h: kgcn_pack_coo_host / kgcn_pack_coo_device).
//
// Replaces the reference's per-step feed of B*C SparseTensorValue triples
// (kgcn/feed.py:112-126 -> kgcn/default_model.py:10 placeholders -> kgcn/core.py:269): the whole
// batch becomes three flat arrays.  The sort is a STABLE counting sort by row, so the entries of
// one CSR row keep their COO storage order -- that is the per-output-element accumulation order of
// tf.sparse_tensor_dense_matmul's CPU kernel -- and duplicates are preserved.
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace kgcn {
namespace {

template <typename IdxT>
int pack_range_host(int64_t m_begin, int64_t m_end, int32_t out_rows, int32_t other_dim, const int64_t* nnz_off,
                    const IdxT* idx, const float* values, bool transpose, int32_t* rowptr, int32_t* col, float* val,
                    int32_t* perm, int64_t* bad_entry) {
    std::vector<int32_t> cursor(static_cast<size_t>(out_rows) + 1);
    for (int64_t m = m_begin; m < m_end; ++m) {
        const int64_t s = nnz_off[m], e = nnz_off[m + 1];
        std::fill(cursor.begin(), cursor.end(), 0);
        for (int64_t k = s; k < e; ++k) {
            const int64_t r = transpose ? idx[2 * k + 1] : idx[2 * k];
            const int64_t c = transpose ? idx[2 * k] : idx[2 * k + 1];
            if (r < 0 || r >= out_rows || c < 0 || c >= other_dim) {
                *bad_entry = k;
                return KGCN_ERR_INDEX_RANGE;
            }
            ++cursor[static_cast<size_t>(r) + 1];
        }
        int32_t run = static_cast<int32_t>(s);
        int32_t* rp = rowptr + m * out_rows;
        for (int32_t i = 0; i < out_rows; ++i) {
            const int32_t cnt = cursor[static_cast<size_t>(i) + 1];
            rp[i] = run;
            cursor[i] = run;
            run += cnt;
        }
        for (int64_t k = s; k < e; ++k) {
            const int64_t r = transpose ? idx[2 * k + 1] : idx[2 * k];
            const int64_t c = transpose ? idx[2 * k] : idx[2 * k + 1];
            const int32_t pos = cursor[r]++;
            col[pos] = static_cast<int32_t>(c);
            val[pos] = values[k];
            if (perm) perm[pos] = static_cast<int32_t>(k);
        }
    }
    return KGCN_OK;
}

template <typename IdxT>
int pack_host(int64_t n_mat, int32_t out_rows, int32_t other_dim, const int64_t* nnz_off, const IdxT* idx,
              const float* values, bool transpose, int32_t* rowptr, int32_t* col, float* val, int32_t* perm) {
    const int64_t nnz = nnz_off[n_mat];
    rowptr[n_mat * out_rows] = static_cast<int32_t>(nnz);
    unsigned n_threads = 1;
    if (nnz > (1 << 16) && n_mat >= 64) n_threads = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (n_threads == 1) {
        int64_t bad = -1;
        int rc = pack_range_host(0, n_mat, out_rows, other_dim, nnz_off, idx, values, transpose, rowptr, col, val,
                                 perm, &bad);
        if (rc) return fail(rc, "pack_coo: sparse index out of range at COO entry %lld (dense_shape [%d, %d])",
                            (long long)bad, transpose ? other_dim : out_rows, transpose ? out_rows : other_dim);
        return KGCN_OK;
    }
    std::vector<std::thread> pool;
    std::vector<int> rcs(n_threads, 0);
    std::vector<int64_t> bads(n_threads, -1);
    const int64_t per = (n_mat + n_threads - 1) / n_threads;
    for (unsigned t = 0; t < n_threads; ++t) {
        const int64_t b = std::min<int64_t>(n_mat, t * per), e = std::min<int64_t>(n_mat, (t + 1) * per);
        pool.emplace_back([=, &rcs, &bads]() {
            rcs[t] = pack_range_host(b, e, out_rows, other_dim, nnz_off, idx, values, transpose, rowptr, col, val, perm,
                                     &bads[t]);
        });
    }
    for (auto& th : pool) th.join();
    for (unsigned t = 0; t < n_threads; ++t)
        if (rcs[t])
            return fail(rcs[t], "pack_coo: sparse index out of range at COO entry %lld (dense_shape [%d, %d])",
                        (long long)bads[t], transpose ? other_dim : out_rows, transpose ? out_rows : other_dim);
    return KGCN_OK;
}

// One CTA per matrix; stable counting sort with the per-row counters in shared memory.
constexpr int kPackThreads = 128;
constexpr int kPackQuadraticLimit = 4096;

__global__ void __launch_bounds__(kPackThreads) pack_coo_kernel(int64_t n_mat, int out_rows, int other_dim,
                                                                const int64_t* __restrict__ nnz_off,
                                                                const int32_t* __restrict__ idx,
                                                                const float* __restrict__ values, int transpose,
                                                                int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                                                float* __restrict__ val, int32_t* __restrict__ perm,
                                                                int32_t* __restrict__ status) {
    pdl_prologue();
    extern __shared__ int32_t start[];  // [out_rows + 1]
    const int64_t m = blockIdx.x;
    const int64_t s = nnz_off[m], e = nnz_off[m + 1];
    const int ri = transpose ? 1 : 0, ci = transpose ? 0 : 1;
    for (int i = threadIdx.x; i <= out_rows; i += kPackThreads) start[i] = 0;
    __syncthreads();
    for (int64_t k = s + threadIdx.x; k < e; k += kPackThreads) {
        const int r = idx[2 * k + ri], c = idx[2 * k + ci];
        if (r < 0 || r >= out_rows || c < 0 || c >= other_dim)
            atomicExch(status, KGCN_ERR_INDEX_RANGE);
        else
            atomicAdd(&start[r + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive scan; out_rows is tens..thousands, once per ingest
        int run = 0;
        for (int i = 1; i <= out_rows; ++i) {
            run += start[i];
            start[i] = run;
        }
    }
    __syncthreads();
    int32_t* rp = rowptr + m * out_rows;
    for (int i = threadIdx.x; i < out_rows; i += kPackThreads) rp[i] = static_cast<int32_t>(s) + start[i];
    if (m == n_mat - 1 && threadIdx.x == 0) rowptr[n_mat * out_rows] = static_cast<int32_t>(e);

    if (e - s <= kPackQuadraticLimit) {
        // rank among earlier same-row entries keeps the storage order (stable)
        for (int64_t k = s + threadIdx.x; k < e; k += kPackThreads) {
            const int r = idx[2 * k + ri], c = idx[2 * k + ci];
            if (r < 0 || r >= out_rows || c < 0 || c >= other_dim) continue;
            int rank = 0;
            for (int64_t j = s; j < k; ++j) rank += (idx[2 * j + ri] == r);
            const int64_t pos = s + start[r] + rank;
            col[pos] = c;
            val[pos] = values[k];
            if (perm) perm[pos] = static_cast<int32_t>(k);
        }
    } else if (threadIdx.x == 0) {  // one huge matrix (block-diagonal B=1 ingest): sequential cursor walk
        for (int64_t k = s; k < e; ++k) {
            const int r = idx[2 * k + ri], c = idx[2 * k + ci];
            if (r < 0 || r >= out_rows || c < 0 || c >= other_dim) continue;
            const int64_t pos = s + start[r]++;
            col[pos] = c;
            val[pos] = values[k];
            if (perm) perm[pos] = static_cast<int32_t>(k);
        }
    }
}

}  // namespace
}  // namespace kgcn

using namespace kgcn;

extern "C" int kgcn_pack_coo_host(int64_t n_mat, int32_t n_rows, int32_t n_cols, const int64_t* nnz_off,
                                  const void* indices, int32_t idx_is_i64, const float* values, int32_t transpose,
                                  int32_t* rowptr, int32_t* col, float* val, int32_t* perm) {
    KGCN_REQUIRE(nnz_off && rowptr, KGCN_ERR_NULL, "pack_coo: NULL pointer argument");
    KGCN_REQUIRE(n_mat >= 0 && n_rows > 0 && n_cols > 0, KGCN_ERR_BAD_SHAPE, "pack_coo: bad shape n_mat=%lld [%d,%d]",
                 (long long)n_mat, n_rows, n_cols);
    KGCN_REQUIRE(nnz_off[0] == 0, KGCN_ERR_BAD_SHAPE, "pack_coo: nnz_off[0] must be 0");
    for (int64_t m = 0; m < n_mat; ++m)
        KGCN_REQUIRE(nnz_off[m + 1] >= nnz_off[m], KGCN_ERR_BAD_SHAPE, "pack_coo: nnz_off not monotone at %lld", (long long)m);
    const int64_t nnz = nnz_off[n_mat];
    KGCN_REQUIRE(nnz < (1ll << 31), KGCN_ERR_BAD_SHAPE, "pack_coo: nnz %lld does not fit int32 offsets", (long long)nnz);
    KGCN_REQUIRE(nnz == 0 || (indices && values && col && val), KGCN_ERR_NULL, "pack_coo: NULL pointer argument");
    const int32_t out_rows = transpose ? n_cols : n_rows;
    const int32_t other = transpose ? n_rows : n_cols;
    KGCN_REQUIRE(n_mat * static_cast<int64_t>(out_rows) < (1ll << 31), KGCN_ERR_BAD_SHAPE, "pack_coo: too many rows");
    if (idx_is_i64)
        return pack_host(n_mat, out_rows, other, nnz_off, static_cast<const int64_t*>(indices), values, transpose != 0,
                         rowptr, col, val, perm);
    return pack_host(n_mat, out_rows, other, nnz_off, static_cast<const int32_t*>(indices), values, transpose != 0,
                     rowptr, col, val, perm);
}

extern "C" int kgcn_pack_coo_device(int64_t n_mat, int32_t n_rows, int32_t n_cols, const int64_t* nnz_off,
                                    const int32_t* indices, const float* values, int32_t transpose, int32_t* rowptr,
                                    int32_t* col, float* val, int32_t* perm, int32_t* status_flag, void* stream) {
    KGCN_REQUIRE(nnz_off && rowptr && status_flag, KGCN_ERR_NULL, "pack_coo_device: NULL pointer argument");
    KGCN_REQUIRE(n_mat > 0 && n_rows > 0 && n_cols > 0, KGCN_ERR_BAD_SHAPE, "pack_coo_device: bad shape");
    KGCN_REQUIRE(n_mat < (1ll << 31), KGCN_ERR_BAD_SHAPE, "pack_coo_device: too many matrices");
    const int32_t out_rows = transpose ? n_cols : n_rows;
    const int32_t other = transpose ? n_rows : n_cols;
    const size_t smem = (static_cast<size_t>(out_rows) + 1) * sizeof(int32_t);
    KGCN_REQUIRE(smem <= 200 * 1024, KGCN_ERR_UNSUPPORTED, "pack_coo_device: %d rows exceed the shared-memory sort", out_rows);
    if (smem > 48 * 1024)
        KGCN_CUDA_OK(cudaFuncSetAttribute(pack_coo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    launch_pdl(pack_coo_kernel, static_cast<unsigned>(n_mat), kPackThreads, smem, static_cast<cudaStream_t>(stream), 
        n_mat, out_rows, other, nnz_off, indices, values, transpose, rowptr, col, val, perm, status_flag);
    KGCN_LAUNCH_OK("pack_coo_kernel");
    return KGCN_OK;
}
