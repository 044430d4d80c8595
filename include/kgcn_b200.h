/*
 * kgcn_b200.h -- C ABI of libkgcn_b200.so: the B200 (sm_100a) batched graph-convolution path.
 *
 * This is the drop-in boundary for the three custom-op libraries kGCN loads with
 * tf.load_op_library but does not ship (citations relative to clinfo/kGCN @ 32328d5):
 *
 *     ./bspmm.so    op "Bspmm"              kgcn/bspmm_call.py:8,15,18
 *     ./bconv.so    op "Bconv"              kgcn/bconv_call.py:8,21,24-25
 *     ./batched.so  ops "Bspmm", "Bspmdt"   kgcn/batched_call.py:8,14,26,29
 *
 * and for the TensorFlow ops the default branch of GraphConv / GraphDense / GraphGather lowers
 * to (kgcn/layers.py:105-116, 243-262, 163-164).  INTEGRATION.md shows the binding a kGCN
 * maintainer would add on top of it.
 *
 * Conventions
 *  - Plain pointers and sizes only; no C++ / torch types.  Unless a parameter says "host", every
 *    pointer is a DEVICE pointer valid on the current CUDA device.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream).  Nothing synchronises the device; nothing allocates or frees device memory: the
 *    caller owns every buffer, scratch space is passed in explicitly (kgcn_*_workspace_bytes).
 *  - Every entry point returns a kgcn_status (0 = OK).  No exceptions cross the ABI, nothing
 *    calls exit().  kgcn_last_error() returns a thread-local message for the last failure.
 *  - Re-entrant; no global mutable state except that thread-local string and one-time
 *    cudaFuncSetAttribute calls.  One host thread per GPU is the intended use.
 *  - float = IEEE fp32.  Index arrays are int32 unless stated.
 *
 * Batched adjacency layout ("BatchedCSR")
 *  A batch holds n_graphs * channels sparse matrices, each [n_rows, n_cols], ordered
 *  graph-major / channel-minor exactly like the flattened lists of kgcn/bconv_call.py:11-15.
 *  Matrix m = g*channels + c owns CSR rows [m*n_rows, (m+1)*n_rows):
 *      rowptr[n_graphs*channels*n_rows + 1]  offsets into col/val, rowptr[0] = 0
 *      col[nnz]                              column inside the matrix, 0 <= col < n_cols
 *      val[nnz]                              fp32 value
 *  Within a row, entries keep their COO *storage order* (stable sort), so per-output-element
 *  summation order equals tf.sparse_tensor_dense_matmul's CPU kernel; duplicates are kept and
 *  therefore accumulate.  The transpose ("adjoint_a") is a second BatchedCSR with n_rows and
 *  n_cols swapped, built by the same packer.
 */
#ifndef KGCN_B200_H_
#define KGCN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGCN_B200_ABI_VERSION 1

typedef enum kgcn_status {
    KGCN_OK = 0,
    KGCN_ERR_BAD_SHAPE = 1,     /* negative / zero / inconsistent sizes                        */
    KGCN_ERR_MISALIGNED = 2,    /* pointer not aligned as the entry point requires             */
    KGCN_ERR_INDEX_RANGE = 3,   /* sparse index outside dense_shape (TF: InvalidArgumentError) */
    KGCN_ERR_CUDA = 4,          /* a CUDA runtime call failed; text in kgcn_last_error()       */
    KGCN_ERR_WORKSPACE = 5,     /* workspace too small / NULL                                  */
    KGCN_ERR_UNSUPPORTED = 6,   /* valid request this build has no kernel for                  */
    KGCN_ERR_NULL = 7           /* required pointer is NULL                                    */
} kgcn_status;

/* Activation fused into epilogues; the set the shipped models apply after GraphConv /
 * GraphDense (example_model/model.py:43-53 sigmoid, sparse_infer.py:43-58 relu, tanh). */
typedef enum kgcn_act { KGCN_ACT_NONE = 0, KGCN_ACT_RELU = 1, KGCN_ACT_SIGMOID = 2, KGCN_ACT_TANH = 3 } kgcn_act;

/* `flags` bits of the GraphConv entry points. */
#define KGCN_FLAG_DEFAULT 0          /* library picks the fastest kernel that meets fp32 parity        */
#define KGCN_FLAG_REFERENCE_ORDER 1  /* force the decomposed W-first path: (X.W + b) with exact-fp32    */
                                     /* FFMA, then A.( ), i.e. the operation order of layers.py:112-113 */
#define KGCN_FLAG_INPUTS_STABLE 4    /* kgcn_gcn_step_chain_f32 only: x and the CSR arrays were complete before the   */
                                    /* PREVIOUS kernel on this stream was launched (a resident batch, not one a packer */
                                    /* or a copy just produced), so the launch may read them while that kernel -- e.g.  */
                                    /* the previous step's reduce + all-reduce + Adam -- is still running (programmatic */
                                    /* dependent launch); parameters are read only after it has completed            */
#define KGCN_FLAG_DY_BROADCAST 2     /* backward only: dy is [n_graphs, f_out] and stands for the same   */
                                     /* row repeated over all n_nodes (gradient of GraphGather,         */
                                     /* layers.py:164): fuses the broadcast into the backward           */

int kgcn_abi_version(void);
/* Thread-local, never NULL, valid until the next failing call on this thread. */
const char* kgcn_last_error(void);
/* Number of CUDA kernels this library has launched in this process so far (monotonic, relaxed
 * atomic; launches recorded during CUDA-graph capture count once).  Diagnostics only. */
uint64_t kgcn_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Ingest: COO (as fed through kgcn/feed.py:112-126 SparseTensorValue triples) -> BatchedCSR.
 * HOST function, HOST pointers.  Replaces the per-placeholder feed of B*C sparse tensors
 * (kgcn/default_model.py:10, kgcn/core.py:267-269).
 *
 *   n_mat          number of matrices (= n_graphs*channels), graph-major / channel-minor
 *   n_rows,n_cols  dense_shape shared by all matrices (kgcn/data_util.py:30-37 align_size)
 *   nnz_off[n_mat+1] exclusive prefix sum of per-matrix nnz (int64)
 *   indices        [nnz,2] (row, col) pairs; idx_is_i64 != 0 -> int64 (TF placeholder dtype),
 *                  else int32 (what data_util.py:40-45 produces)
 *   values[nnz]    fp32
 * Outputs (caller-allocated, host): rowptr[n_mat*n_rows+1], col[nnz], val[nnz], and
 *   perm[nnz] (optional, may be NULL): CSR position -> original COO position.
 * Pass transpose != 0 to build the adjoint instead (rowptr then has n_mat*n_cols+1 entries).
 * Stable counting sort: entries of one CSR row keep COO storage order.
 * Errors: KGCN_ERR_INDEX_RANGE for an index outside [0,n_rows) x [0,n_cols).
 */
int kgcn_pack_coo_host(int64_t n_mat, int32_t n_rows, int32_t n_cols, const int64_t* nnz_off,
                       const void* indices, int32_t idx_is_i64, const float* values, int32_t transpose,
                       int32_t* rowptr, int32_t* col, float* val, int32_t* perm);

/* Same, on the device: inputs/outputs are DEVICE pointers, int32 indices only, one CTA per
 * matrix does a stable counting sort in shared memory.  `status_flag` (device int32, caller
 * zeroes it) is set to KGCN_ERR_INDEX_RANGE if any index is out of range. */
int kgcn_pack_coo_device(int64_t n_mat, int32_t n_rows, int32_t n_cols, const int64_t* nnz_off,
                         const int32_t* indices, const float* values, int32_t transpose,
                         int32_t* rowptr, int32_t* col, float* val, int32_t* perm,
                         int32_t* status_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Batched SpMM -- the "Bspmm" / "Bconv" / "Bspmdt" ops.
 *
 * For every graph g and channel c:   OUT(g,c) (+)= A[g,c] . RHS(g,c)
 *     RHS(g,c) = rhs + g*rhs_stride_g + c*rhs_stride_c   row-major [n_cols, feat]
 *     OUT(g,c) = out + g*out_stride_g + c*out_stride_c   row-major [n_rows, feat]
 * (strides in floats).  With out_stride_c == 0 the channel contributions of one graph are
 * summed in channel order into the same output (Bconv, kgcn/bconv_call.py:10-21; the tf.add_n
 * of layers.py:115); with rhs_stride_c == 0 all channels share one right-hand side
 * (GINAggregate, layers.py:468; the aggregate-first GraphConv).  Bspmm
 * (kgcn/bspmm_call.py:10-15) is channels = 1; Bspmdt (kgcn/batched_call.py:21-26) is
 * channels = 1 with rhs_stride_g = n_cols*feat over the stacked dense matrix.
 * `self_scale` (device, [channels], may be NULL): adds self_scale[c] * RHS(g,c) row-wise
 * (requires n_rows == n_cols) -- the epsilon term of GINAggregate (layers.py:469).
 * adjoint_a is expressed by passing the transposed BatchedCSR.
 * Output is fully overwritten (rows without entries become 0).
 */
int kgcn_bspmm_f32(const int32_t* rowptr, const int32_t* col, const float* val,
                   int64_t n_graphs, int32_t channels, int32_t n_rows, int32_t n_cols, int32_t feat,
                   const float* rhs, int64_t rhs_stride_g, int64_t rhs_stride_c,
                   float* out, int64_t out_stride_g, int64_t out_stride_c,
                   const float* self_scale, void* stream);

/* Gradient w.r.t. the sparse values (kgcn/bspmm_call.py:49-54):
 *   dval[perm ? perm[e] : e] = < DY(g,c)[row_e, :], RHS(g,c)[col_e, :] >   for every CSR entry e.
 * DY strides follow the OUT convention above (stride_c = 0 replicates dY over channels as
 * kgcn/bconv_call.py:46 does). */
int kgcn_bspmm_dvalues_f32(const int32_t* rowptr, const int32_t* col, const int32_t* perm,
                           int64_t n_graphs, int32_t channels, int32_t n_rows, int32_t n_cols, int32_t feat,
                           const float* dy, int64_t dy_stride_g, int64_t dy_stride_c,
                           const float* rhs, int64_t rhs_stride_g, int64_t rhs_stride_c,
                           float* dval, void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphConv forward (kgcn/layers.py:64-116, all four branches compute the same function):
 *     y[g,i,:] = act( sum_c sum_{(i,j,v) in A[g,c]} v * ( x[g,j,:] . w[c] + bias[c] ) )
 *   x [n_graphs, n_nodes, f_in]; w [channels, f_in, f_out]; bias [channels, f_out] (may be
 *   NULL = zeros); y [n_graphs, n_nodes, f_out]; act = kgcn_act (KGCN_ACT_NONE reproduces the
 *   layer itself; the others fuse the tf.sigmoid / tf.nn.relu the models apply next).
 * The bias is applied BEFORE aggregation as the reference does (layers.py:112-113), i.e. a
 * node receives rowsum(A[g,c])[i] * bias[c], and a row without entries outputs act(0).
 * `workspace` must hold kgcn_graphconv_workspace_bytes(...) bytes.
 */
size_t kgcn_graphconv_workspace_bytes(int64_t n_graphs, int32_t channels, int32_t n_nodes,
                                      int32_t f_in, int32_t f_out);
int kgcn_graphconv_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val,
                           int64_t n_graphs, int32_t channels, int32_t n_nodes,
                           const float* x, int32_t f_in, const float* w, const float* bias, int32_t f_out,
                           int32_t act, float* y, int32_t flags, void* workspace, size_t workspace_bytes,
                           void* stream);

/* GraphConv backward (TF autodiff through layers.py:112-116; the sparse half is
 * kgcn/bspmm_call.py:44: dB = Bspmm(A, dY, adjoint_a=True)).
 *   rowptr_t/col_t/val_t  transposed BatchedCSR of the forward adjacency
 *   y, dy                 forward output (post-activation) and its gradient
 *   dx [n_graphs,n_nodes,f_in] (may be NULL: first layer), dw [channels,f_in,f_out],
 *   dbias [channels,f_out] -- all overwritten, not accumulated.
 */
int kgcn_graphconv_bwd_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t,
                           int64_t n_graphs, int32_t channels, int32_t n_nodes,
                           const float* x, int32_t f_in, const float* w, int32_t f_out,
                           int32_t act, const float* y, const float* dy,
                           float* dx, float* dw, float* dbias, int32_t flags,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphDense (kgcn/layers.py:223-265): y = act(x . kernel + bias) on all n_graphs*n_nodes rows
 * (padding included, layers.py:255-262).  If enabled_node_nums (device int32 [n_graphs]) is
 * non-NULL, rows i >= enabled_node_nums[g] are written as exact zeros (layers.py:243-254).
 * kernel [f_in,f_out], bias [f_out] or NULL.
 */
int kgcn_graphdense_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t f_in,
                            const float* kernel, const float* bias, int32_t f_out, int32_t act,
                            const int32_t* enabled_node_nums, float* y, void* stream);
size_t kgcn_graphdense_workspace_bytes(int64_t n_graphs, int32_t n_nodes, int32_t f_in, int32_t f_out);
int kgcn_graphdense_bwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t f_in,
                            const float* kernel, int32_t f_out, int32_t act,
                            const int32_t* enabled_node_nums, const float* y, const float* dy,
                            float* dx, float* dkernel, float* dbias,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphGather (kgcn/layers.py:156-167): out[g,:] = sum_i x[g,i,:] over ALL n_nodes rows
 * (padding included).  Backward broadcasts: dx[g,i,:] = dout[g,:].
 */
int kgcn_gather_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* out, void* stream);
int kgcn_gather_bwd_f32(const float* dout, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphMaxPooling (kgcn/layers.py:122-153): per molecule, channel and feature the reference forms the sparse
 * matrix A * x[:, k] (column j scaled by x[j, k]), densifies it (tf.sparse_tensor_to_dense: absent entries
 * become 0) and takes tf.reduce_max over axis 1; channels are summed (tf.add_n):
 *   y[g,i,k] = sum_c max( {A_c[i,j] * x[g,j,k] : (i,j) stored} U {0 if row i stores fewer than n_nodes entries} ).
 * workspace (kgcn_maxpool_workspace_bytes; NULL for inference) receives the per-channel maxima and tie counts
 * the backward needs.  Backward = TF's gradient of that op chain (ties share the gradient evenly, implicit
 * zeros included) as a gather over the TRANSPOSED BatchedCSR: deterministic.  Square matrices (n_nodes x n_nodes).
 */
size_t kgcn_maxpool_workspace_bytes(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t feat);
int kgcn_maxpool_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                         int32_t channels, int32_t n_nodes, const float* x, int32_t feat, float* y,
                         void* workspace, size_t workspace_bytes, void* stream);
int kgcn_maxpool_bwd_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                         int32_t channels, int32_t n_nodes, const float* x, int32_t feat, const float* dy,
                         const void* workspace, size_t workspace_bytes, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-molecule readout of the block-diagonal model (example_model/sparse.py:79-90: tf.scan of reduce_sum over
 * the row range of every molecule): out[m,:] = sum of rows start[m] .. start[m]+size[m]-1 of x [rows, feat].
 * start / size: device int64[n_segments].  Backward broadcasts dout[m,:] to those rows (rows outside every
 * segment are left untouched).
 */
int kgcn_segment_sum_fwd_f32(const float* x, const int64_t* start, const int64_t* size, int64_t n_segments,
                             int32_t feat, float* out, void* stream);
int kgcn_segment_sum_bwd_f32(const float* dout, const int64_t* start, const int64_t* size, int64_t n_segments,
                             int32_t feat, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphBatchNormalization (kgcn/layers.py:170-220, kgcn/legacy/layers.py:170-218): per-feature
 *   y = gamma * (x - mean) / sqrt(var + eps) + beta   on the first enabled_node_nums[g] rows of molecule g,
 * exact zeros on the rest (the extract / normalise / split / pad sequence of layers.py:202-214);
 * enabled_node_nums == NULL: every row (layers.py:216-219).
 * mode 0: mean / var are INPUTS (moving statistics -- what the Keras layer uses under the reference trainer);
 * mode 1: mean / var are OUTPUTS, the batch statistics over the enabled rows (biased variance; the legacy
 * tf.layers.batch_normalization(training=True) variant).  The legacy variant without enabled_node_nums
 * normalises per (node, feature) pair over the batch: call with n_nodes = 1, feat = N * F.
 * gamma / beta may be NULL (1 / 0).  Backward returns dx (may be NULL), dgamma, dbeta (may be NULL).
 * workspace: kgcn_graph_bn_workspace_bytes (+ 2 * feat floats for the backward).
 */
size_t kgcn_graph_bn_workspace_bytes(int64_t n_graphs, int32_t feat);
int kgcn_graph_bn_fwd_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat,
                          const int32_t* enabled_node_nums, const float* gamma, const float* beta, float* mean,
                          float* var, float eps, int32_t mode, float* y, void* workspace, size_t workspace_bytes,
                          void* stream);
int kgcn_graph_bn_bwd_f32(const float* x, const float* dy, int64_t n_graphs, int32_t n_nodes, int32_t feat,
                          const int32_t* enabled_node_nums, const float* gamma, const float* mean, const float* var,
                          float eps, int32_t mode, float* dx, float* dgamma, float* dbeta, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Step-loop helpers (the minimal trainer around the layers; kgcn/core.py:121-127, 267-269).
 *
 * Readout head of the shipped classifiers (example_model/model.py:56-69):
 *   logits = g . w + bias;  prediction = softmax(logits);
 *   cost[b] = mask[b] * softmax_cross_entropy(labels[b], logits[b]);
 *   stats[0] = cost_sum = sum_b cost[b];  stats[1] = correct_count (argmax match, masked);
 *   `stats` is a float[4] device buffer the caller zero-initialises ONCE: [2] holds an internal block
 *   ticket that the kernel restores to zero, [3] is reserved (deterministic last-block reductions,
 *   no memset and no second launch per call);
 *   gradients of cost_opt = inv_batch * sum_b cost[b] (reduce_mean, inv_batch = 1/batch_size):
 *   dlogits [B,L], dg [B,F] = dlogits . w^T, dw [F,L] = g^T . dlogits, dbias [L].
 * g [n_graphs, feat]; w [feat, n_labels]; labels [n_graphs, n_labels] (one-hot or soft);
 * mask [n_graphs] or NULL.  Any output pointer except stats may be NULL.  n_labels <= 32;
 * any n_graphs (every block stages a slice of <= 128 graphs in shared memory).
 */
size_t kgcn_readout_workspace_bytes(int64_t n_graphs, int32_t feat, int32_t n_labels);
int kgcn_readout_xent_f32(const float* g, int64_t n_graphs, int32_t feat, const float* w, const float* bias,
                          int32_t n_labels, const float* labels, const float* mask, float inv_batch,
                          float* logits, float* prediction, float* stats, float* dlogits, float* dg,
                          float* dw, float* dbias, void* workspace, size_t workspace_bytes, void* stream);

/* GraphGather + the readout head in one launch: g [n_graphs, feat] is an OUTPUT here, formed from the node rows
 * x [n_graphs, n_nodes, feat] exactly like kgcn_gather_fwd_f32 (rows added in index order); everything else as
 * kgcn_readout_xent_f32.  Saves one launch and one pass over the last layer's activations per step. */
int kgcn_gather_readout_xent_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* g,
                                 const float* w, const float* bias, int32_t n_labels, const float* labels,
                                 const float* mask, float inv_batch, float* logits, float* prediction, float* stats,
                                 float* dlogits, float* dg, float* dw, float* dbias, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* Forward of a layer whose feature widths were padded to multiples of 32 by the caller (zero columns in x, zero rows /
 * columns in w, zero bias entries) so that Tox21-style widths (75 -> 50) run on the tcgen05 layer kernel: output columns
 * >= f_out_valid are written as exact zeros whatever the activation (sigmoid(0) = 0.5 would otherwise leak into the
 * gradients of the next layer's zero rows).  Same mathematics as kgcn_graphconv_fwd_f32 on the un-padded problem.
 * kgcn_graphconv_fwd_fused(shape) != 0 says the fused layer kernel takes the (padded) shape. */
int32_t kgcn_graphconv_fwd_fused(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t f_in, int32_t f_out);
int kgcn_graphconv_fwd_padded_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                  int32_t channels, int32_t n_nodes, const float* x, int32_t f_in, const float* w,
                                  const float* bias, int32_t f_out, int32_t f_out_valid, int32_t act, float* y,
                                  void* stream);

/* kgcn_gather_readout_xent_f32 that also emits the dU of the last graph layer for the step loop:
 *   du_nodes[b, r, :] = dg[b, :] (.) act'(x[b, r, :])      (the GraphGather gradient of layers.py:164 -- a broadcast over the
 * node rows -- times the gradient of the activation `act` that produced x), so the layer's backward
 * (kgcn_graphconv_bwd_partial_f32) needs no separate activation-gradient pass.  du_nodes and dg are required. */
int kgcn_gather_readout_xent_du_f32(const float* x, int64_t n_graphs, int32_t n_nodes, int32_t feat, float* g,
                                    const float* w, const float* bias, int32_t n_labels, const float* labels,
                                    const float* mask, float inv_batch, float* logits, float* prediction, float* stats,
                                    float* dlogits, float* dg, float* dw, float* dbias, int32_t act, float* du_nodes,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* Step-loop form of the GraphConv backward (same mathematics as kgcn_graphconv_bwd_f32; kgcn/layers.py:105-116 with the
 * registered gradients of bspmm_call.py:28-52 / bconv_call.py:28-70), two launches:
 *   dx      = sum_c (A_c^T . du) . W_c^T              the fused layer kernel on (A^T, du, W^T); skipped when dx == NULL.
 *             With act_below != KGCN_ACT_NONE the result is multiplied by act_below'(x) in the epilogue: x is the output
 *             of the layer below, so dx then IS that layer's dU and the chain needs no activation-gradient launches.
 *   partial = per-CTA blocks [(f_in + 1), channels * f_out] of X^T . [A_0^T.du | A_1^T.du | ..] (last row: column sums =
 *             dbias) -- NOT reduced: kgcn_reduce_adam_f32 sums them in fixed order inside the optimizer launch.
 * du [n_graphs, n_nodes, f_out] = dy (.) act'(y) of this layer (from the head, or the dx of the layer above).
 * kgcn_graphconv_bwd_splits returns the number of partial blocks for a shape (0: shape not supported by the fused
 * kernels -- feature widths must be multiples of 32, f_in <= 128, channels * f_out <= 256; use kgcn_graphconv_bwd_f32);
 * partial must hold splits * (f_in + 1) * channels * f_out floats. */
int32_t kgcn_graphconv_bwd_splits(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t f_in, int32_t f_out,
                                  int32_t need_dx);
int kgcn_graphconv_bwd_partial_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                   int32_t channels, int32_t n_nodes, const float* x, int32_t f_in, const float* w,
                                   int32_t f_out, const float* du, float* dx, int32_t act_below, float* partial,
                                   size_t partial_bytes, void* stream);

/* Chained launches for the step loop: ALL GraphConv layers of a network in one launch.  A CTA owns the same graph range in
 * every layer and the aggregation never leaves a graph, so layer l + 1 starts on a CTA as soon as that CTA has finished layer l
 * (CTA-local barrier + proxy fence; no grid-wide dependency, no launch / drain / fill per layer).  dims[n_layers + 1] are the
 * STORED widths (multiples of 32), dims_valid the logical ones (NULL = same; columns beyond are written as exact zeros).
 *   kgcn_graphconv_chain_fwd_f32   y[l] = act(GraphConv_l(y[l - 1])), y[-1] = x;  w / bias / y are HOST arrays of n_layers
 *                                  device pointers (w[l] [C][dims[l]][dims[l+1]], bias[l] [C][dims[l+1]] or bias == NULL).
 *   kgcn_graphconv_chain_dx_f32    du[l - 1] = (sum_c A_c^T . du[l] . W_l,c^T) (.) act'(x[l]) for l = n_layers - 1 .. 1, where
 *                                  x[l] is the input of layer l (= y[l - 1]); du[n_layers - 1] is the input (from the head),
 *                                  du[l] is [B, N, dims[l + 1]].  x, w, du are HOST arrays of n_layers device pointers
 *                                  (x[0] is not read).  Same mathematics as kgcn_graphconv_bwd_partial_f32's dx, layer by layer.
 * kgcn_graphconv_chain_supported(...) != 0: every layer (and every dx) has a single-CTA plan on these widths (<= 4 layers). */
int32_t kgcn_graphconv_chain_supported(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims);
int kgcn_graphconv_chain_fwd_f32(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n_graphs,
                                 int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                 const int32_t* dims_valid, const float* x, const float* const* w, const float* const* bias,
                                 float* const* y, int32_t act, void* stream);
int kgcn_graphconv_chain_dx_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                const float* const* x, const float* const* w, float* const* du, int32_t act, void* stream);

/* The graph-local part of a whole training step in ONE launch: forward layers 0 .. L-1, the readout head fused into the last
 * layer's epilogue (GraphGather + Dense(n_labels) + softmax cross-entropy on the CTA's own tiles, example_model/model.py:56-69;
 * kgcn_gather_readout_xent_du_f32's mathematics), and the dx chain of kgcn_graphconv_chain_dx_f32.  The last layer's
 * activations are never written: its epilogue stores du[L-1] = d gathered (.) act'(H_{L-1}) directly.
 *   y[l] (l < L-1) layer outputs; du[l] dU of layer l (all L written); w / bias as in kgcn_graphconv_chain_fwd_f32;
 *   head_w [dims[L]][n_labels] (rows beyond the logical width zero), head_b [n_labels] or NULL, labels [B][n_labels],
 *   mask [B] or NULL; outputs logits / prediction [B][n_labels], gathered [B][dims[L]] (each may be NULL);
 *   head_partial [kgcn_gcn_step_chain_grid(...)][dims[L] * n_labels + 8]: per-CTA {dW_dense | db_dense (padded to 4) |
 *   cost_sum, correct_count, 0, 0}, reduced by kgcn_reduce_adam_f32 (segments with rows = 0 + stats_partial).
 * kgcn_gcn_step_chain_grid returns the number of CTAs (= partial blocks), 0 when the network is not supported
 * (needs kgcn_graphconv_chain_supported, dims[L] <= 64, n_labels <= 4, 2 L - 1 <= 6 jobs).
 * flags: KGCN_FLAG_DEFAULT or KGCN_FLAG_INPUTS_STABLE. */
int32_t kgcn_gcn_step_chain_grid(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                 int32_t n_labels);
int kgcn_gcn_step_chain_f32(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* rowptr_t,
                            const int32_t* col_t, const float* val_t, int64_t n_graphs, int32_t channels, int32_t n_nodes,
                            int32_t n_layers, const int32_t* dims, const int32_t* dims_valid, const float* x,
                            const float* const* w, const float* const* bias, float* const* y, float* const* du, int32_t act,
                            const float* head_w, const float* head_b, int32_t n_labels, const float* labels, const float* mask,
                            float inv_batch, float* logits, float* prediction, float* gathered, float* head_partial,
                            uint32_t flags, void* stream);

/* The same launch, which ALSO stores what its dx jobs aggregate: g_save (HOST array of n_layers device pointers, entry 0 unused)
 * receives G_l = [A_0^T . du[l] | A_1^T . du[l] | ..] ([B, N, channels * dims[l + 1]], fp32, 128-byte aligned) for l = n_layers - 1 .. 1 -- the `adjoint_a=True` product of
 * the registered gradient (kgcn/bspmm_call.py:44), which the dx job computes anyway as the aggregate of (A^T, du[l], W_l^T).
 * kgcn_graphconv_chain_dw_g_f32 then reads G_l instead of gathering it a second time; the results are bit-identical to the
 * launches without g (same per-row accumulation order, same tf32 split).  Networks whose step launch can do it:
 * kgcn_gcn_step_chain_g_supported(...) != 0.  g_save == NULL: exactly kgcn_gcn_step_chain_f32. */
int32_t kgcn_gcn_step_chain_g_supported(int64_t n_graphs, int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims);
int kgcn_gcn_step_chain_g_f32(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* rowptr_t,
                              const int32_t* col_t, const float* val_t, int64_t n_graphs, int32_t channels, int32_t n_nodes,
                              int32_t n_layers, const int32_t* dims, const int32_t* dims_valid, const float* x,
                              const float* const* w, const float* const* bias, float* const* y, float* const* du,
                              float* const* g_save, int32_t act, const float* head_w, const float* head_b, int32_t n_labels,
                              const float* labels, const float* mask, float inv_batch, float* logits, float* prediction,
                              float* gathered, float* head_partial, uint32_t flags, void* stream);

/* Weight-gradient partials of ALL layers (kgcn_graphconv_bwd_partial_f32 with dx == NULL, layer by layer) in as few launches
 * as tensor memory allows (2 * channels * dims[l + 1] accumulator columns per layer, 512 per launch): x[l] = input of layer l,
 * du[l] = its dU, partial[l] / partial_bytes[l] its partial blocks.  HOST arrays of n_layers entries. */
int kgcn_graphconv_chain_dw_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                const float* const* x, const float* const* du, float* const* partial,
                                const size_t* partial_bytes, void* stream);
/* g (HOST array of n_layers device pointers or NULL): where g[l] != NULL, layer l's G_l = A^T . du[l] is read from
 * there (written by kgcn_gcn_step_chain_g_f32) instead of being gathered from du[l] and the transposed CSR. */
int kgcn_graphconv_chain_dw_g_f32(const int32_t* rowptr_t, const int32_t* col_t, const float* val_t, int64_t n_graphs,
                                  int32_t channels, int32_t n_nodes, int32_t n_layers, const int32_t* dims,
                                  const float* const* x, const float* const* du, const float* const* g, float* const* partial,
                                  const size_t* partial_bytes, void* stream);

/* The reduction alone (gradient checks, the NCCL cross-check path): dw [channels][f_in][f_out] and dbias [channels][f_out]
 * (may be NULL) from `splits` partial blocks, summed in split order (deterministic). */
int kgcn_reduce_partials_f32(const float* partial, int32_t splits, int32_t f_in, int32_t f_out, int32_t channels,
                             float* dw, float* dbias, void* stream);

/* The tail of a training step in ONE launch: fixed-order reduction of the weight-gradient partials -> (data parallel:
 * one-shot all-reduce over NVLink peer memory) -> Adam (kgcn_adam_f32's formulation, device-side step counter).
 * `segments` (HOST array) says which ranges of the flat buffers are fed from partial blocks: kernel [channels][rows][cols]
 * at kernel_off, bias [channels][cols] at bias_off (-1: none), partial [splits][(rows + 1)][channels * cols].  Every other
 * element takes its gradient from grad[i] as it is.  The reduced (and, with a group, all-reduced) gradient is written back
 * to grad.  n, every segment offset and every segment width must be multiples of 4 (16-byte vector accesses).
 * `group` (HOST struct, may be NULL = single GPU): rank r owns a mailbox of 2 * world * n_pad 8-byte slots {fp32 value,
 * step tag} (n_pad >= n), zero-initialised, and all W mailboxes are mapped into this process (kgcn_p2p_alloc on the owning
 * rank, kgcn_p2p_open on the others).  Each rank PUSHES its tagged sums into every peer's mailbox (posted NVLink stores) and
 * spins on its own, local one; sums are formed in rank order, so all ranks hold bit-identical gradients and parameters.
 * Every rank must enqueue the call once per step; a rank that waits ~2 s for a peer sets *error_flag (device int32, may be
 * NULL) instead of hanging the GPU. */
typedef struct kgcn_grad_segment {
    int64_t kernel_off, bias_off;
    const float* partial;
    int32_t splits, rows, cols, channels;
    int64_t stride;      /* floats between the splits' blocks; 0 = (rows + 1) * channels * cols.  rows = 0: a flat
                            gradient of `cols` floats at bias_off (kernel_off ignored) -- the fused head's Dense kernel */
} kgcn_grad_segment;
typedef struct kgcn_p2p_group {
    int32_t rank, world;
    int64_t n_pad;
    void* mailbox[8];
    int32_t* error_flag;
} kgcn_p2p_group;
/* stats_partial (may be NULL): per-CTA {cost_sum, correct_count} of the fused head every stats_stride floats, summed in CTA
 * order into stats_out[0..1] by one extra block of the same launch. */
int kgcn_reduce_adam_f32(float* param, float* grad, float* m, float* v, int64_t n, const kgcn_grad_segment* segments,
                         int32_t n_segments, float lr, float beta1, float beta2, float eps, float grad_scale,
                         int32_t* step_state, const kgcn_p2p_group* group, const float* stats_partial,
                         int32_t stats_splits, int64_t stats_stride, float* stats_out, void* stream);
/* Peer-mapped device buffers for the exchange above (cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle; the 64-byte
 * handle travels between the ranks' processes through any host channel, e.g. torch.distributed.all_gather).  HOST calls. */
int kgcn_p2p_alloc(size_t n_bytes, void** device_ptr, unsigned char* handle64);
int kgcn_p2p_open(const unsigned char* handle64, void** device_ptr);
int kgcn_p2p_close(void* device_ptr);
int kgcn_p2p_free(void* device_ptr);

/* Adam with TensorFlow's formulation (tf.train.AdamOptimizer, kgcn/core.py:121-127):
 *   g = grad * grad_scale;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   param -= lr * sqrt(1-b2^step)/(1-b1^step) * m / (sqrt(v) + eps);     step counts from 1.
 * One launch over a flat buffer of n floats (all layers' kernels and biases back to back).
 * step_state (device int32[2], zero-initialised by the caller, may be NULL): when given, the step
 * number is read from and advanced on the DEVICE (step = step_state[0] + 1) and the host `step`
 * argument is ignored -- launch parameters then never change, so a training step can be captured
 * once in a CUDA graph and replayed. */
int kgcn_adam_f32(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, float grad_scale, int32_t* step_state, void* stream);

/* ------------------------------------------------------------------------------------------
 * Record IO for the block-diagonal ("sparse") ingest and the checkpoint reader.  HOST functions,
 * HOST pointers, no CUDA.  They replace what the reference gets from TensorFlow's C++ runtime:
 * tf.data.TFRecordDataset (task_sparse_gcn.py:119) and tf.io.parse_single_example with the
 * feature_spec of task_sparse_gcn.py:153-166, over files written by
 * kgcn/preprocessing/utils.py:178-226 (convert_to_example / save_tfrecords); and the CRC-32C
 * (Castagnoli) that guards both TFRecord files and tensor-bundle checkpoints
 * (model/reaction/model.best.ckpt.*, restored by kgcn/core.py's tf.train.Saver).
 */
/* CRC-32C of n_bytes at data (0 for an empty range); _masked applies TensorFlow's storage mask
 * rotr(crc, 15) + 0xa282ead8. */
uint32_t kgcn_crc32c(const void* data, size_t n_bytes);
uint32_t kgcn_crc32c_masked(const void* data, size_t n_bytes);

/* Walks a whole TFRecord file image (u64le length | u32le masked crc of the length | data |
 * u32le masked crc of the data, repeated).  Writes the byte offset and length of each record's
 * data into rec_off / rec_len (first `capacity` records) and the number of records found into
 * *n_records; call with capacity 0 to count.  verify_crc != 0 checks both checksums of every
 * record.  KGCN_ERR_BAD_SHAPE for a truncated or corrupted file (TF: DataLossError). */
int kgcn_tfrecord_scan(const void* file, size_t n_bytes, int32_t verify_crc, int64_t* rec_off,
                       int64_t* rec_len, int64_t capacity, int64_t* n_records);

/* For each of n_rec serialized tensorflow.Example messages (file + rec_off[r], rec_len[r] bytes)
 * extracts the feature stored under `key` and appends its values to `values`: kind 1 = float_list
 * -> fp32, kind 2 = int64_list -> int64 (packed or unpacked encoding).  counts[r] (optional) = number of values record r
 * contributed (0 if the key is absent: VarLenFeature semantics; the FixedLenFeature shape check is
 * the caller's), *total = their sum.  `values` may be NULL (count only).  KGCN_ERR_WORKSPACE if
 * `capacity` values are not enough (*total then says how many are needed), KGCN_ERR_BAD_SHAPE for
 * a malformed message, KGCN_ERR_UNSUPPORTED if the feature holds another kind than requested
 * (TF: InvalidArgumentError) or for kind 0 (bytes lists: not used by the feature_spec).
 * This is exactly the concatenated layout construct_batched_adjacency_and_feature_matrices
 * (kgcn/data_util.py:698-845) takes after tf.data batching + tf.sparse.to_dense flattening. */
int kgcn_tfexample_gather(const void* file, const int64_t* rec_off, const int64_t* rec_len, int64_t n_rec,
                          const char* key, int32_t kind, void* values, int64_t capacity, int64_t* counts,
                          int64_t* total);

#ifdef __cplusplus
}
#endif
#endif /* KGCN_B200_H_ */
