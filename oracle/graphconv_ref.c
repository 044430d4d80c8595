/*
 * graphconv_ref.c -- CPU restatement (plain C + OpenMP over molecules) of kGCN's default
 * GraphConv path and of one training step of the classifier family of example_model/model.py.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-checked against oracle/ref_layers.py), by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * kgcn_b200/ links or loads it.
 *
 * PARITY UNPINNED for the arithmetic (the reference ships no golden outputs for these layers and
 * TensorFlow is not installable here); pinned instead to the hand-derived
 * KAT vectors and to oracle/ref_layers.py (tests/test_oracle_c.py).
 *
 * Structure follows the reference op for op, per molecule and per channel
 * (clinfo/kGCN @ 32328d5, kgcn/layers.py:105-116):
 *     fw = matmul(x[b], W_c) + bias_c                      layers.py:112
 *     el = sparse_tensor_dense_matmul(adj[b][c], fw)       layers.py:113  (per nnz, storage order)
 *     o[b] = add_n(el over c)                              layers.py:115
 * then the model's activation (example_model/model.py:43), GraphGather = reduce_sum over nodes
 * (layers.py:164), Dense(label_dim), mask * softmax cross-entropy, reduce_mean
 * (model.py:56-61), TF autodiff for the backward (sparse part: kgcn/bspmm_call.py:44) and
 * tf.train.AdamOptimizer (kgcn/core.py:121-127).  It is a PORT of the algorithm without
 * TensorFlow's per-op dispatch, so it is a conservative (fast) stand-in for the TF-CPU path.
 *
 * Flat parameter layout (identical to kgcn_b200/trainer.py, every tensor padded to 4 floats):
 *     for each layer l: kernel[C][F_l][F_{l+1}], bias[C][F_{l+1}];  then out_w[F_L][labels], out_b[labels]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float act_fwd(float x, int act) {
    switch (act) {
        case 1: return x > 0.0f ? x : 0.0f;
        case 2: return 1.0f / (1.0f + expf(-x));
        case 3: return tanhf(x);
        default: return x;
    }
}
static inline float act_bwd_from_out(float y, int act) {
    switch (act) {
        case 1: return y > 0.0f ? 1.0f : 0.0f;
        case 2: return y * (1.0f - y);
        case 3: return 1.0f - y * y;
        default: return 1.0f;
    }
}
static inline int64_t pad4(int64_t n) { return (n + 3) / 4 * 4; }

/* out[N,Fo] = x[N,Fi] . w[Fi,Fo] + bias[Fo] */
static void dense_bias(const float* x, const float* w, const float* bias, int N, int Fi, int Fo, float* out) {
    for (int i = 0; i < N; ++i) {
        float* o = out + (size_t)i * Fo;
        for (int j = 0; j < Fo; ++j) o[j] = 0.0f;
        for (int k = 0; k < Fi; ++k) {
            const float a = x[(size_t)i * Fi + k];
            const float* wr = w + (size_t)k * Fo;
            for (int j = 0; j < Fo; ++j) o[j] += a * wr[j];
        }
        if (bias)
            for (int j = 0; j < Fo; ++j) o[j] += bias[j];
    }
}

/* GraphConv forward for molecules [b_lo, b_hi); y = act(sum_c A_bc . (x_b W_c + bias_c)). */
static void graphconv_fwd_range(int64_t b_lo, int64_t b_hi, int C, int N, int Fi, int Fo, const int64_t* nnz_off,
                                const int32_t* idx, const float* val, const float* x, const float* w, const float* bias,
                                int act, float* y, float* fw /* scratch N*Fo */) {
    for (int64_t b = b_lo; b < b_hi; ++b) {
        float* yb = y + (size_t)b * N * Fo;
        memset(yb, 0, sizeof(float) * (size_t)N * Fo);
        for (int c = 0; c < C; ++c) {
            dense_bias(x + (size_t)b * N * Fi, w + (size_t)c * Fi * Fo, bias ? bias + (size_t)c * Fo : NULL, N, Fi, Fo, fw);
            for (int64_t e = nnz_off[b * C + c]; e < nnz_off[b * C + c + 1]; ++e) {
                const int i = idx[2 * e], j = idx[2 * e + 1];
                const float v = val[e];
                float* o = yb + (size_t)i * Fo;
                const float* f = fw + (size_t)j * Fo;
                for (int k = 0; k < Fo; ++k) o[k] += v * f[k];
            }
        }
        if (act)
            for (size_t k = 0; k < (size_t)N * Fo; ++k) yb[k] = act_fwd(yb[k], act);
    }
}

int kgcn_ref_graphconv_fwd(int64_t B, int C, int N, int Fi, int Fo, const int64_t* nnz_off, const int32_t* idx,
                           const float* val, const float* x, const float* w, const float* bias, int act, float* y,
                           int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        float* fw = (float*)malloc(sizeof(float) * (size_t)N * Fo);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b)
            graphconv_fwd_range(b, b + 1, C, N, Fi, Fo, nnz_off, idx, val, x, w, bias, act, y, fw);
        free(fw);
    }
    return 0;
}

int kgcn_ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* One training step.  dims[0..n_layers] = feature widths (dims[0] = input features).
 * stats[0] = cost_sum, stats[1] = correct_count.  grads is overwritten with d cost_opt / d params
 * where cost_opt = inv_batch * sum_b mask_b * xent_b.  If apply_update, Adam step `step` (>= 1). */
int kgcn_ref_train_step(int64_t B, int N, int C, int n_layers, const int* dims, int n_labels, int act,
                        const int64_t* nnz_off, const int32_t* idx, const float* val, const float* x,
                        const float* labels, const float* mask, float inv_batch, float* params, float* grads,
                        float* adam_m, float* adam_v, int step, float lr, int apply_update, int want_grads,
                        float* stats, float* logits_out, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    const int T = omp_get_max_threads();
#else
    const int T = 1;
#endif
    int64_t off_w[64], off_b[64], n_params = 0;
    if (n_layers > 60 || n_labels > 64 || dims[n_layers] > 1024) return 1;
    for (int l = 0; l < n_layers; ++l) {
        off_w[l] = n_params; n_params += pad4((int64_t)C * dims[l] * dims[l + 1]);
        off_b[l] = n_params; n_params += pad4((int64_t)C * dims[l + 1]);
    }
    const int FL = dims[n_layers];
    const int64_t off_ow = n_params; n_params += pad4((int64_t)FL * n_labels);
    const int64_t off_ob = n_params; n_params += pad4(n_labels);
    int fmax = 0;
    size_t act_elems = 0;
    for (int l = 0; l <= n_layers; ++l) { if (dims[l] > fmax) fmax = dims[l]; act_elems += (size_t)N * dims[l]; }

    float* tgrads = want_grads ? (float*)calloc((size_t)T * n_params, sizeof(float)) : NULL;
    double cost_sum = 0.0, correct = 0.0;

#pragma omp parallel reduction(+ : cost_sum, correct)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        float* g = tgrads ? tgrads + (size_t)tid * n_params : NULL;
        float* h = (float*)malloc(sizeof(float) * act_elems);          /* activations of one molecule */
        float* fw = (float*)malloc(sizeof(float) * (size_t)N * fmax);
        float* dh = (float*)malloc(sizeof(float) * (size_t)N * fmax);
        float* dfw = (float*)malloc(sizeof(float) * (size_t)N * fmax);
        float* dprev = (float*)malloc(sizeof(float) * (size_t)N * fmax);
        float gathered[1024], z[64], dz[64];
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            /* ---- forward ---- */
            float* hl = h;
            memcpy(hl, x + (size_t)b * N * dims[0], sizeof(float) * (size_t)N * dims[0]);
            for (int l = 0; l < n_layers; ++l) {
                const int Fi = dims[l], Fo = dims[l + 1];
                float* hn = hl + (size_t)N * Fi;
                memset(hn, 0, sizeof(float) * (size_t)N * Fo);
                for (int c = 0; c < C; ++c) {
                    dense_bias(hl, params + off_w[l] + (size_t)c * Fi * Fo, params + off_b[l] + (size_t)c * Fo, N, Fi, Fo, fw);
                    for (int64_t e = nnz_off[b * C + c]; e < nnz_off[b * C + c + 1]; ++e) {
                        const int i = idx[2 * e], j = idx[2 * e + 1];
                        const float v = val[e];
                        for (int k = 0; k < Fo; ++k) hn[(size_t)i * Fo + k] += v * fw[(size_t)j * Fo + k];
                    }
                }
                if (act)
                    for (size_t k = 0; k < (size_t)N * Fo; ++k) hn[k] = act_fwd(hn[k], act);
                hl = hn;
            }
            for (int k = 0; k < FL; ++k) gathered[k] = 0.0f;
            for (int i = 0; i < N; ++i)
                for (int k = 0; k < FL; ++k) gathered[k] += hl[(size_t)i * FL + k];
            float zmax = -INFINITY;
            for (int c = 0; c < n_labels; ++c) {
                float a = 0.0f;
                for (int k = 0; k < FL; ++k) a += gathered[k] * params[off_ow + (size_t)k * n_labels + c];
                z[c] = a + params[off_ob + c];
                if (z[c] > zmax) zmax = z[c];
                if (logits_out) logits_out[b * n_labels + c] = z[c];
            }
            float se = 0.0f;
            for (int c = 0; c < n_labels; ++c) se += expf(z[c] - zmax);
            const float lse = logf(se) + zmax;
            const float m = mask ? mask[b] : 1.0f;
            float cost = 0.0f, ysum = 0.0f;
            int ap = 0, ay = 0;
            for (int c = 0; c < n_labels; ++c) {
                const float yv = labels[b * n_labels + c];
                cost -= yv * (z[c] - lse);
                ysum += yv;
                if (z[c] > z[ap]) ap = c;
                if (yv > labels[b * n_labels + ay]) ay = c;
            }
            cost_sum += (double)(m * cost);
            correct += (double)(m * (ap == ay ? 1.0f : 0.0f));
            if (!want_grads) continue;
            /* ---- backward ---- */
            for (int c = 0; c < n_labels; ++c) {
                dz[c] = m * inv_batch * (expf(z[c] - lse) * ysum - labels[b * n_labels + c]);
                g[off_ob + c] += dz[c];
            }
            for (int k = 0; k < FL; ++k) {
                float a = 0.0f;
                for (int c = 0; c < n_labels; ++c) {
                    g[off_ow + (size_t)k * n_labels + c] += gathered[k] * dz[c];
                    a += dz[c] * params[off_ow + (size_t)k * n_labels + c];
                }
                for (int i = 0; i < N; ++i) dh[(size_t)i * FL + k] = a;      /* grad of reduce_sum(axis=1) */
            }
            for (int l = n_layers - 1; l >= 0; --l) {
                const int Fi = dims[l], Fo = dims[l + 1];
                float* hin = hl - (size_t)N * Fi;
                for (size_t k = 0; k < (size_t)N * Fo; ++k) dh[k] *= act_bwd_from_out(hl[k], act);
                if (l > 0) memset(dprev, 0, sizeof(float) * (size_t)N * Fi);
                for (int c = 0; c < C; ++c) {
                    memset(dfw, 0, sizeof(float) * (size_t)N * Fo);
                    for (int64_t e = nnz_off[b * C + c]; e < nnz_off[b * C + c + 1]; ++e) {   /* A^T . du */
                        const int i = idx[2 * e], j = idx[2 * e + 1];
                        const float v = val[e];
                        for (int k = 0; k < Fo; ++k) dfw[(size_t)j * Fo + k] += v * dh[(size_t)i * Fo + k];
                    }
                    float* gw = g + off_w[l] + (size_t)c * Fi * Fo;
                    float* gb = g + off_b[l] + (size_t)c * Fo;
                    const float* wc = params + off_w[l] + (size_t)c * Fi * Fo;
                    for (int i = 0; i < N; ++i) {
                        const float* d = dfw + (size_t)i * Fo;
                        for (int k = 0; k < Fo; ++k) gb[k] += d[k];
                        for (int a = 0; a < Fi; ++a) {
                            const float xv = hin[(size_t)i * Fi + a];
                            float* gr = gw + (size_t)a * Fo;
                            for (int k = 0; k < Fo; ++k) gr[k] += xv * d[k];
                        }
                        if (l > 0)
                            for (int a = 0; a < Fi; ++a) {
                                const float* wr = wc + (size_t)a * Fo;
                                float s = 0.0f;
                                for (int k = 0; k < Fo; ++k) s += d[k] * wr[k];
                                dprev[(size_t)i * Fi + a] += s;
                            }
                    }
                }
                if (l > 0) memcpy(dh, dprev, sizeof(float) * (size_t)N * Fi);
                hl = hin;
            }
        }
        free(h); free(fw); free(dh); free(dfw); free(dprev);
    }
    if (stats) { stats[0] = (float)cost_sum; stats[1] = (float)correct; }
    if (want_grads) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_params; ++i) {
            float s = 0.0f;
            for (int t = 0; t < T; ++t) s += tgrads[(size_t)t * n_params + i];
            grads[i] = s;
        }
        free(tgrads);
        if (apply_update) {
            const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
            const float lr_t = lr * sqrtf(1.0f - powf(b2, (float)step)) / (1.0f - powf(b1, (float)step));
            for (int64_t i = 0; i < n_params; ++i) {
                adam_m[i] = b1 * adam_m[i] + (1.0f - b1) * grads[i];
                adam_v[i] = b2 * adam_v[i] + (1.0f - b2) * grads[i] * grads[i];
                params[i] -= lr_t * adam_m[i] / (sqrtf(adam_v[i]) + eps);
            }
        }
    }
    return 0;
}
