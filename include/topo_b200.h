/*
 * topo_b200.h -- C ABI of the B200-native simplicial-complex stage.
 *
 * One shared library (libtopo_b200.so, built by topo_audio_autoencoder_b200/csrc/build.py with
 * nvcc -gencode arch=compute_100a,code=sm_100a).  Plain pointers and sizes only; no torch types.
 * The reference (Monlarc/topo-audio-autoencoder) is pure Python and has no FFI layer, so each
 * entry point cites the Python function whose arithmetic it replaces; INTEGRATION.md shows the
 * ctypes binding that puts it behind the reference's own signatures.
 *
 * Conventions
 *   - every pointer named dev_* / without a host_ prefix is DEVICE memory on the current device;
 *     host_* pointers are host memory.
 *   - outputs are caller-allocated; entry points neither allocate nor synchronise (the two
 *     topo_tables_* constructors/destructors excepted) and enqueue on `stream`.
 *   - return value: TOPO_OK or an error code; topo_last_error() gives a thread-local message.
 *   - all floating point is fp32 (IEEE, denormals kept: the sparsity pattern of the reference's
 *     operators is "whatever is non-zero in fp32", complex_builder.py:82-85).
 *
 * Layout of one batch of complexes ("simplex axis")
 *   A complex on n vertices has n_r = C(n, r+1) candidate simplices of rank r = 0..3, listed in
 *   itertools.combinations (lexicographic) order (rectifier.py:28-30).  Per sample all ranks share
 *   one axis of length N = n_0+n_1+n_2+n_3 with rank r in [off_r, off_r + n_r).  A batch is a
 *   row-major [B, N] array.  Feature matrices are COMPACT: rank r holds only the active simplices
 *   of every sample, samples concatenated, rows in ascending simplex id -- exactly the row order
 *   of the reference's per-sample [n_r_active, C] tensors (encoder.py:230-247).
 */
#ifndef TOPO_B200_H
#define TOPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOPO_OK 0
#define TOPO_ERR_INVALID 1      /* bad argument */
#define TOPO_ERR_CUDA 2         /* CUDA runtime error (message in topo_last_error) */
#define TOPO_ERR_UNSUPPORTED 3  /* valid request outside what the kernels are instantiated for */

typedef void* topo_stream_t;            /* a cudaStream_t */
typedef struct topo_tables topo_tables; /* opaque: static combinatorial tables of one n, on one device */

int topo_version(void);
const char* topo_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * T1. Constraint tables.   Replaces ConstraintMatrices.create(n)  (rectifier.py:24-64).
 * Built on the host with closed-form lexicographic rank/unrank (no linear searches) and
 * uploaded once.  Holds: vertex lists, face ids, coface ids and sorted adjacency neighbour lists.
 * ------------------------------------------------------------------------------------------- */
int topo_tables_create(int n_vertices, topo_tables** out);
/* upload_to_device == 0 builds the host tables only (no CUDA call; for inspection and CPU tests) */
int topo_tables_create_ex(int n_vertices, int upload_to_device, topo_tables** out);
void topo_tables_destroy(topo_tables* t);
/* counts[r] = n_r; offsets[r] = off_r (5 entries, offsets[4] = N) */
int topo_tables_sizes(const topo_tables* t, int64_t host_counts[4], int64_t host_offsets[5]);
/* [n_r, r+1] int64 vertex ids of every rank-r simplex == SimplexIndices.{edges,triangles,tetra} */
int topo_tables_simplex_vertices(const topo_tables* t, int rank, int64_t* host_out);
/* [n_r, r+1] int32 ids (within rank r-1) of the faces of every rank-r simplex, ascending; rank>=1 */
int topo_tables_faces(const topo_tables* t, int rank, int32_t* host_out);
/* [n_r, n-1-r] int32 ids (within rank r+1) of the cofaces of every rank-r simplex, ascending; rank<=2 */
int topo_tables_cofaces(const topo_tables* t, int rank, int32_t* host_out);
/* dense 0/1 [n_r, n_{r-1}] fp32 == vertex_to_edge / edge_to_triangle / triangle_to_tetra */
int topo_tables_face_matrix(const topo_tables* t, int rank, float* dev_out, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G1. Gates.
 * Hard Concrete (README.md:15-18 names it; the reference ships no code -- spec in DESIGN.md):
 *   x = logits + loc[rank];  training: x = (log u - log(1-u) + x) / beta
 *   s = sigmoid(x);  z = clamp(s*(zeta-gamma)+gamma, 0, 1);  ste: value (z > 0.5), gradient of z
 * params = device float[7] {beta, gamma, zeta, loc0, loc1, loc2, loc3}.  128-bit vectorised.
 * grad_params (device float[7]) is overwritten.  u may be NULL when training == 0.  bwd workspace: NULL (the seven sums are
 * accumulated with floating-point atomics) or topo_hard_concrete_bwd_workspace_floats floats (per-CTA sums added in CTA order:
 * bit-reproducible).
 * Binary Gumbel: BinaryGumbel.forward training branch (encoder.py:34-41),
 *   softmax(([l, 1-l] + g) / temp, dim 0)[0] with g = dev_gumbels [2, count].
 * ------------------------------------------------------------------------------------------- */
int topo_hard_concrete_fwd(const float* logits, const float* u, const float* params,
                           const int64_t host_offsets[5], int64_t batch, int training, int ste,
                           float* z, topo_stream_t stream);
int topo_hard_concrete_bwd(const float* logits, const float* u, const float* params,
                           const int64_t host_offsets[5], int64_t batch, int training,
                           const float* grad_z, float* grad_logits, float* grad_params,
                           float* workspace, topo_stream_t stream);
int64_t topo_hard_concrete_bwd_workspace_floats(const int64_t host_offsets[5], int64_t batch);
int topo_binary_gumbel_fwd(const float* logits, const float* gumbels, float temp, int64_t count,
                           float* probs, topo_stream_t stream);
int topo_binary_gumbel_bwd(const float* logits, const float* gumbels, float temp, int64_t count,
                           const float* grad_probs, float* grad_logits, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * R2. Geometric-mean face rectifier.   Replaces enforce_constraints  (rectifier.py:75-127).
 *   c_s = exp(sum_{f face of s} log(p'_f + eps) / (r+1)), forced to exact 0 when any face is 0;
 *   p'_s = min(p_s, c_s), level by level on the rectified level below.
 * Backward follows torch autograd of that expression, including torch.minimum's 50/50 split on
 * ties and the zero gradient path of the masked branch.  workspace: [batch, N] floats.
 * ------------------------------------------------------------------------------------------- */
int topo_rectify_fwd(const topo_tables* t, const float* probs_in, float eps, int64_t batch,
                     float* probs_out, topo_stream_t stream);
int topo_rectify_bwd(const topo_tables* t, const float* probs_in, const float* probs_out,
                     const float* grad_out, float eps, int64_t batch, float* grad_in,
                     float* workspace, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G3. Active sets.   Replaces the four nonzero() calls of get_active_simplex_embeddings
 * (encoder.py:230-233) for a whole batch.
 *   pos     [B, N] int32: position of the simplex in its sample's ascending active list, or -1
 *   act_idx [B, N] int32: act_idx[b, off_r + i] = id (within rank r) of the i-th active simplex
 *   counts  [B, 4] int32
 *   row_off [4, B+1] int32: exclusive scan of counts over the batch (compact row of sample b
 *           rank r starts at row_off[r][b]; row_off[r][B] = total active rows of rank r)
 * ------------------------------------------------------------------------------------------- */
int topo_active_sets(const topo_tables* t, const float* probs, int64_t batch, int32_t* pos,
                     int32_t* act_idx, int32_t* counts, int32_t* row_off, topo_stream_t stream);

/* A batch of complexes as the SCCN / embedding / operator kernels see it (device pointers). */
typedef struct {
    const float* probs;      /* [B, N] rectified probabilities */
    const int32_t* pos;      /* [B, N] */
    const int32_t* act_idx;  /* [B, N] */
    const int32_t* counts;   /* [B, 4]; 16-byte aligned (a sample's four counts are one 128-bit load) */
    const int32_t* row_off;  /* [4, B+1] */
    int64_t batch;
} topo_complex_view;

/* ---------------------------------------------------------------------------------------------
 * L1 / L2. Structural penalties.  compute_vertex_penalty (encoder.py:199-203) and
 * compute_entropy_loss (encoder.py:205-221; line 223 raises in the reference and is omitted).
 * One value per sample.
 * ------------------------------------------------------------------------------------------- */
int topo_penalties_fwd(const topo_tables* t, const float* probs, int64_t batch, float min_active,
                       float max_active, float* vertex_penalty, float* entropy_loss,
                       topo_stream_t stream);
int topo_penalties_bwd(const topo_tables* t, const float* probs, int64_t batch, float min_active,
                       float max_active, const float* grad_vertex_penalty,
                       const float* grad_entropy_loss, float* grad_probs, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G3 (cont). Active embeddings:  X_r[row] = lne_r[id] * p[id]   (encoder.py:242-247), where
 * lne_r = LayerNorm(Embedding table) is computed once per step with topo_layernorm_fwd.
 * bwd: grad_probs[b, off_r+id] += <g, lne[id]>;  grad_lne[id] = sum_b p * g  (overwritten).
 * ------------------------------------------------------------------------------------------- */
int topo_embed_fwd(const topo_tables* t, const topo_complex_view* cv, int rank, int channels,
                   const float* lne, float* x_out, topo_stream_t stream);
int topo_embed_bwd(const topo_tables* t, const topo_complex_view* cv, int rank, int channels,
                   const float* lne, const float* grad_x, float* grad_lne, float* grad_probs,
                   topo_stream_t stream);

/* Row-wise LayerNorm over [rows, channels] (nn.LayerNorm, eps inside the sqrt, biased variance).
 * bwd overwrites grad_x and ACCUMULATES into grad_gamma / grad_beta: with floating-point atomics when workspace is NULL, from
 * per-CTA column sums added in CTA order (bit-reproducible) when it holds topo_layernorm_bwd_workspace_floats floats. */
int topo_layernorm_fwd(int64_t rows, int channels, const float* x, const float* gamma,
                       const float* beta, float eps, float* y, topo_stream_t stream);
int topo_layernorm_bwd(int64_t rows, int channels, const float* x, const float* gamma, float eps,
                       const float* grad_y, float* grad_x, float* grad_gamma, float* grad_beta,
                       float* workspace, topo_stream_t stream);
int64_t topo_layernorm_bwd_workspace_floats(int64_t rows, int channels);

/* ---------------------------------------------------------------------------------------------
 * B2. Weighted incidence / adjacency operators.  Replaces build_sparse_matrices
 * (complex_builder.py:23-115) for ONE sample.  op: 0..3 = adjacency rank_0..rank_3,
 * 4..6 = incidence rank_1..rank_3.  Active sets come from the caller (pos/act_idx/counts of one
 * sample, as produced by topo_active_sets or derived from user index lists).
 *   count: row_nnz[op][i] for every compact row          -> dev_row_nnz  [7, max_rows]
 *   scan : exclusive scan per operator                   -> dev_row_ptr  [7, max_rows + 1]
 *   fill : COO (row-major sorted) int64 indices + values -> per-operator caller buffers
 *   bwd  : grad_probs[N] += d(values)/d(probs) . grad_values
 * Values are single fp32 products (p_e | p_t*p_t | p_s*p_s | p_s*p_s' | p_coface), so they are
 * bit-identical to the reference's dense products; an entry exists iff its value != 0.
 * ------------------------------------------------------------------------------------------- */
int topo_operators_count(const topo_tables* t, const float* probs, const int32_t* pos,
                         const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                         int32_t* row_ptr /* [7, max_rows+1] */, topo_stream_t stream);
int topo_operators_fill(const topo_tables* t, const float* probs, const int32_t* pos,
                        const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                        const int32_t* row_ptr, int64_t* const dev_rows[7], int64_t* const dev_cols[7],
                        float* const dev_vals[7], topo_stream_t stream);
int topo_operators_bwd(const topo_tables* t, const float* probs, const int32_t* pos,
                       const int32_t* act_idx, const int32_t* counts, int64_t max_rows,
                       const int32_t* row_ptr, const float* const dev_grad_vals[7],
                       float* grad_probs, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * S1 (a). Matrix-free neighbourhood aggregation for a whole batch -- the SpMM half of
 * convs_*(x_source, neighborhood) (custom_sccn.py:78-81, 95-98, 113-116), using
 * N @ (X @ W) == (N @ X) @ W and the factorisation of every adjacency through the incidences
 * (complex_builder.py:62-64):
 *   down[r] = I_{r+1} X_{r+1}          r = 0..2   (rows: rank r)      "from above"
 *   up[r]   = I_r^T  X_{r-1}           r = 1..3   (rows: rank r)      "from below"
 *   same[0] = A_0 X_0
 *   same[1] = I_2 up[2] - diag.,  same[2] = I_3 up[3] - diag.,  same[3] = I_3^T down[2] - diag.
 * No operator is materialised: neighbours come from the static tables, weights from probs.
 * x / down / up / same: arrays of per-rank compact [rows_r, channels] device pointers
 * (down[3] and up[0] unused, may be NULL).
 * bwd: g_down / g_up / g_same are the gradients w.r.t. the aggregates (g_down[2], g_up[2],
 * g_up[3] are updated in place to their totals); g_x is ACCUMULATED (+=), g_probs [B, N] too.
 * ------------------------------------------------------------------------------------------- */
int topo_sccn_aggregate_fwd(const topo_tables* t, const topo_complex_view* cv, int channels,
                            const float* const x[4], float* const down[4], float* const up[4],
                            float* const same[4], topo_stream_t stream);
int topo_sccn_aggregate_bwd(const topo_tables* t, const topo_complex_view* cv, int channels,
                            const float* const x[4], const float* const down[4],
                            const float* const up[4], float* const g_down[4], float* const g_up[4],
                            const float* const g_same[4], float* const g_x[4], float* g_probs,
                            topo_stream_t stream);

/* Generic CSR path for caller-supplied sparse operators (GradientSCCN.forward accepts arbitrary
 * matrices, test_sccn.py:15-35):  y = A x;  g_val[e] = <g_y[row(e)], x[col(e)]>. */
int topo_spmm_csr(int64_t rows, const int32_t* row_ptr, const int32_t* col_idx, const float* vals,
                  const float* x, int channels, float* y, topo_stream_t stream);
int topo_sddmm_csr(int64_t rows, const int32_t* row_ptr, const int32_t* col_idx, const float* g_y,
                   const float* x, int channels, float* g_vals, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * S1 (b). Message combine for one rank -- the dense half of GradientSCCNLayer.forward
 * (custom_sccn.py:73-136) on a tile of target rows, fused in one kernel:
 *   m_k  = scale_k * (agg_k @ W_k) + x                    (conv weight, scale, residual)
 *   a    = softmax_k( Linear2(GELU(Linear1(m_k))) )       (message attention, :128-130)
 *   out  = sum_k a_k m_k;  out = LayerNorm(out) if apply_ln (:132-134)
 * n_msgs in 1..3.  n_rows_dev (nullable) is a device int32 holding the live row count
 * (<= rows); rows past it are not touched.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int channels;
    int n_msgs;
    const float* agg[3];     /* [rows, C] aggregate per message */
    const float* w[3];       /* [C, C] conv weight (in, out) per message */
    const float* scale[3];   /* device scalar per message */
    const float* x;          /* [rows, C] current features (residual); NULL when residual is off */
    const float* att_w1;     /* [C, C]  nn.Linear(C, C).weight  (out, in) */
    const float* att_b1;     /* [C] */
    const float* att_w2;     /* [C]     nn.Linear(C, 1).weight */
    const float* att_b2;     /* [1] */
    const float* ln_gamma;   /* [C] */
    const float* ln_beta;    /* [C] */
    float ln_eps;
    int apply_ln;
    /* Optional saved activations, one [rows, C] buffer per message (all NULL = not used).  The forward
     * kernels WRITE them; topo_sccn_combine_bwd_attention READS them and then skips both recompute GEMMs:
     *   saved_m[k]   = scale_k (agg_k W_k) + x          (the message)
     *   saved_pre[k] = Linear1(m_k) = W1 m_k + b1       (pre-GELU attention hidden layer) */
    float* saved_m[3];
    float* saved_pre[3];
    /* Optional [3, rows] attention scores Linear2(GELU(pre_k)) (NULL = not used).  Written by
     * topo_sccn_combine_fwd_tc; required by topo_sccn_combine_bwd_tc, which then evaluates erf once. */
    float* saved_score;
    /* Memory layout of saved_m / saved_pre.  TOPO_SAVED_ROW_MAJOR: [rows, C].  TOPO_SAVED_TILE_FRAGMENT: a
     * permutation inside every 128-row tile that makes the tensor-core kernels' accesses contiguous
     * (csrc/layout.cuh); the buffers then hold rows rounded up to a multiple of 128.  topo_sccn_combine_fwd_tc2
     * writes (only) the tile-fragment layout; topo_sccn_combine_bwd_tc reads either. */
    int saved_layout;
    /* Optional (channels == 64): this rank's weight images from topo_sccn_prepare_images, laid out as
     *   [ W1 image | message 0: W image, V image | message 1: ... ]   (24 KB each)
     * When NULL the tensor-core kernels build the images themselves (about 30 us of set-up per launch). */
    const void* weight_images;
    /* Upper bound on the CTAs (= SMs) the tensor-core kernels of this call may occupy; 0 = all of them.  The four
     * ranks of a layer are independent, so the host launches them on four streams with the SMs divided in
     * proportion to their work: they run side by side instead of paying four set-ups and four tails in turn. */
    int max_ctas;
} topo_combine_params;

#define TOPO_SAVED_ROW_MAJOR 0
#define TOPO_SAVED_TILE_FRAGMENT 1

typedef struct {
    float* g_agg[3];         /* [rows, C] overwritten */
    float* g_x;              /* [rows, C] overwritten (residual path only) */
    float* g_wprod[3];       /* [C, C] ACCUMULATED: agg_k^T (dL/dm_k); dW_k = scale_k * this,
                                dscale_k = <W_k, this> (finished by the caller, 2 tiny ops) */
    float* g_att_w1;         /* ACCUMULATED */
    float* g_att_b1;
    float* g_att_w2;
    float* g_att_b2;
    float* g_ln_gamma;
    float* g_ln_beta;
    /* topo_sccn_combine_bwd_tc only.  NULL: the parameter gradients above are ACCUMULATED with floating-point atomics (their
     * low bits depend on the order the CTAs finish in).  Otherwise: [CTAs of the launch][TOPO_CTA_PARTIAL_FLOATS] scratch --
     * CTA b stores its own partial sums at slot b ([w1 C*C][wprod_0..2 C*C each][b1 C][w2 C][gamma C][beta C][b2 1]) with plain
     * stores, the seven pointers above are NOT touched, and topo_sccn_finish_weight_grads adds the slots in CTA order:
     * bit-reproducible parameter gradients.  The launch uses topo_sccn_combine_grid(rows, max_ctas) CTAs. */
    float* cta_partials;
} topo_combine_grads;
#define TOPO_CTA_PARTIAL_FLOATS 16704

/* bf16x3 operand images of the weights (csrc/weight_images.cu), built once per layer and step and shared by the
 * forward and backward tensor-core kernels of all ranks.  A job with `w` writes the image of W_k [in][out]
 * (24,576 bytes) followed by the image of V_k = s_k W_k W1^T (24,576 bytes) to dst; a job with w == NULL writes
 * the image of W1 [out][in] (24,576 bytes).  At most 96 jobs per call (all layers of a 6-layer SCCN), one CTA each. */
typedef struct {
    const float* w;
    const float* scale;
    const float* att_w1;
    void* dst;
} topo_image_job;
#define TOPO_WEIGHT_IMAGE_BYTES 24576
int topo_sccn_prepare_images(const topo_image_job* jobs, int n_jobs, int channels, topo_stream_t stream);

/* The tail of the parameter-gradient chain for up to 96 jobs in one launch (one CTA each, deterministic).  With p = wprod
 * (accumulated by the backward) or, when `partials` is set, p[i] = sum over the launch's active CTAs b, in order, of
 * partials[b * TOPO_CTA_PARTIAL_FLOATS + partial_offset + i]  (active = min(topo_sccn_combine_grid(rows, max_ctas),
 * ceil(min(*n_rows_dev, rows) / 128)); n_rows_dev may be NULL):
 *   w != NULL:  g_w = scale * p  ([count] elements),  g_scale[0] = <w, p>      (conv weights: count = C * C)
 *   w == NULL:  g_w = p                                                          (attention / LayerNorm parameters) */
typedef struct {
    const float* wprod;
    const float* w;
    const float* scale;
    float* g_w;
    float* g_scale;
    const float* partials;
    const int32_t* n_rows_dev;
    int64_t rows;
    int32_t partial_offset;
    int32_t count;           /* 0 = channels * channels */
    int32_t max_ctas;
    int32_t reserved_;
} topo_wgrad_job;
int topo_sccn_finish_weight_grads(const topo_wgrad_job* jobs, int n_jobs, int channels, topo_stream_t stream);
/* CTAs topo_sccn_combine_fwd_tc2 / topo_sccn_combine_bwd_tc launch for `rows` rows under the max_ctas bound of topo_combine_params */
int topo_sccn_combine_grid(int64_t rows, int max_ctas);

int topo_sccn_combine_fwd(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                          float* out, topo_stream_t stream);
/* Same contract, channels == 64 only: the GEMMs run on the tensor cores (tcgen05.mma kind::tf32, 3xTF32
 * operand splitting for fp32-level accuracy, accumulators in tensor memory), 128-row tiles. */
int topo_sccn_combine_fwd_tc(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                             float* out, topo_stream_t stream);
/* Second-generation tensor-core forward (channels == 64): bf16x3 operand images, ONE 128 x 128 x 64 product
 * per message ([agg_k W_k | agg_k (s_k W_k W1^T)], no dependent GEMM chain), double-buffered operand slots and
 * accumulators, saved activations in the tile-fragment layout (always written: saved_m, saved_pre,
 * saved_score must be set and saved_layout == TOPO_SAVED_TILE_FRAGMENT).  csrc/combine_fwd16.cu */
int topo_sccn_combine_fwd_tc2(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                              float* out, topo_stream_t stream);
/* workspace: n_msgs * rows * C floats (dL/dm_k between the two backward kernels).
 * _attention: LayerNorm, softmax and attention-MLP backward -> dL/dm_k (workspace), g_x, attention and
 *             LayerNorm parameter gradients.  _conv: dL/dagg_k and g_wprod from the workspace.
 * topo_sccn_combine_bwd runs both. */
int topo_sccn_combine_bwd_attention(const topo_combine_params* p, int64_t rows,
                                    const int32_t* n_rows_dev, const float* grad_out,
                                    const topo_combine_grads* g, float* workspace,
                                    topo_stream_t stream);
int topo_sccn_combine_bwd_conv(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                               const topo_combine_grads* g, const float* workspace,
                               topo_stream_t stream);
int topo_sccn_combine_bwd(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                          const float* grad_out, const topo_combine_grads* g, float* workspace,
                          topo_stream_t stream);
/* topo_sccn_combine_bwd_conv on the tensor cores (channels == 64): both the input gradient and the
 * weight-gradient product (a contraction over rows, staged as transposed K-major tiles) are 3xTF32
 * tcgen05 GEMMs; the 64 x 64 product accumulates in tensor memory across the CTA's tiles. */
int topo_sccn_combine_bwd_conv_tc(const topo_combine_params* p, int64_t rows,
                                  const int32_t* n_rows_dev, const topo_combine_grads* g,
                                  const float* workspace, topo_stream_t stream);
/* The whole combine backward in one tensor-core kernel (channels == 64; needs saved_m, saved_pre and
 * saved_score from topo_sccn_combine_fwd_tc): LayerNorm / softmax / attention-MLP backward, g_agg, g_x and
 * all weight-gradient products as bf16x3 tcgen05 GEMMs on operand images staged once and read in both
 * majors (csrc/combine_bwd_tc.cu).  Same outputs as topo_sccn_combine_bwd; no workspace. */
int topo_sccn_combine_bwd_tc(const topo_combine_params* p, int64_t rows, const int32_t* n_rows_dev,
                             const float* grad_out, const topo_combine_grads* g, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Unit-test and measurement entry points.  NOT part of the product library: libtopo_b200.so neither declares nor
 * exports them.  They exist in libtopo_b200_debug.so, the same sources compiled with -DTOPO_DEBUG_KERNELS=1 plus
 * csrc/gemm16_debug.cu (csrc/build.py build_debug()), which tests/test_gpu_tc.py and scripts/ablate_*.py load.
 * ------------------------------------------------------------------------------------------- */
#ifdef TOPO_DEBUG_KERNELS
/* Unit-test entry of the tensor-core path (tcgen05.mma kind::tf32, 3xTF32 operand splitting):
 * out[rows, 64] = a[rows, 64] @ w[64, 64].  mode 0: both operands in shared memory (K-major SWIZZLE_128B);
 * mode 3: the A operand in tensor memory (tcgen05.st by the row threads). */
int topo_debug_gemm_tf32x3(const float* a, const float* w, int64_t rows, int mode, float* out,
                           topo_stream_t stream);

/* Unit-test entry of the bf16x3 tensor-core path (tcgen05.mma kind::f16 on three bf16 parts per fp32
 * operand, six part products, fp32 accumulation; csrc/tc16.cuh).  One 16-bit SWIZZLE_128B image serves
 * both operand majors:
 *   mode 0: out[rows, 64] = a @ w     (w stored [in][out], read MN-major)
 *   mode 1: out[rows, 64] = a @ w^T   (w stored [out][in], read K-major)
 *   mode 2: out[64, 64]  += a^T w     (w is a second [rows, 64] matrix; both read MN-major, contraction over rows)
 *   mode 3: layout probe (where the rows of an M = 64 accumulator land in tensor memory)
 *   mode 4: mode 1 issued as tcgen05.mma.cta_group::2 by clusters of two CTAs that each hold half of the weight image
 * mn_lbo / mn_sbo / mn_kstep: descriptor fields of the MN-major operands in bytes (16384 / 1024 / 2048 for
 * the layouts in tc16.cuh; exposed so the test can pin them). */
int topo_debug_gemm_bf16x3(const float* a, const float* w, int64_t rows, int mode, int mn_lbo, int mn_sbo,
                           int mn_kstep, float* out, topo_stream_t stream);

/* Measurement hooks of the two tensor-core combine kernels (scripts/ablate_fwd16.py, scripts/ablate_bwd.py).  They only
 * act in a library built with TOPO_DEBUG_KERNELS=1; the shipped build ignores them.
 *   mask:   bit set = one part of the forward switched off for differential timing (results are then wrong)
 *   stamps: device buffer of uint64 that CTA 0 / thread 0 fills with %globaltimer values at phase boundaries */
void topo_debug_fwd16_mask(int mask);
void topo_debug_fwd16_stamps(unsigned long long* device_buffer);
void topo_debug_bwd_stamps(unsigned long long* device_buffer);
#endif  /* TOPO_DEBUG_KERNELS */

/* ---------------------------------------------------------------------------------------------
 * D1-D3. Tiled pairwise spectral distance.  Replaces the pair loop of compute_distances
 * (precompute_distances.py:89-115) and BatchAudioDistance.forward (:33-49) on precomputed
 * magnitude spectrograms:  spec [n, d], d = sum of the n_scales segment lengths (one segment per
 * STFT scale, flattened),
 *   d(i,j) = sum_s [ mean((x_s-y_s)^2) / (mean(x_s^2) + 1e-7) + mean|log(x_s+eps) - log(y_s+eps)| ]
 * with x = the LOWER-index clip (so the normaliser comes from it, :89, :106-110), mirrored
 * (:114-115), zero diagonal.  The squared-difference term is evaluated as mean x^2 + mean y^2 - 2 <x, y> / len with the
 * Gram term <x, y> on the tensor cores (bf16x3, fp32-accurate, two-level accumulation); the L1-of-logs term on the integer
 * pipe (one VABSDIFF per pair-element) from Q6.20 fixed-point logs q(v) = round((v + 40) 2^20): exact for |v| >= 8, 9.5e-7
 * resolution below, window [-40, 24) in log units (log_eps >= 1e-17, magnitudes below 2.6e10; values outside saturate).
 *   padded_size     : row length Dp after every segment is padded to a multiple of 64 bins
 *   image_bytes     : size of the bf16x3 operand image of n clips ([ceil(n / 128)][Dp / 64] tiles of 48 KB)
 *   logq_words      : 32-bit words of the fixed-point log operand of n clips ([ceil(n / 64)][Dp][64], k-major inside
 *                     64-clip blocks so that a stage of the L1 kernel is three contiguous 8 KB pieces)
 *   workspace_floats: scratch of one rows / block call (Gram and L1 partials per scale)
 *   prepare         : spec -> logq (16-byte aligned), image (1024-byte aligned; rows past n in the last 128-clip block must
 *                     be zero: allocate it zero-filled), sq_mean [n, n_scales]
 *   rows            : the block [row_begin, row_end) x [col_begin, col_end) of the symmetric matrix of ONE prepared set
 *                     (row-block sharding across ranks needs no collective)
 * ------------------------------------------------------------------------------------------- */
int64_t topo_distance_padded_size(const int64_t* host_seg_len, int n_scales);
int64_t topo_distance_image_bytes(int64_t n, const int64_t* host_seg_len, int n_scales);
int64_t topo_distance_logq_words(int64_t n, const int64_t* host_seg_len, int n_scales);
int64_t topo_distance_workspace_floats(int64_t n_rows, int64_t n_cols, int n_scales);
int topo_distance_prepare(const float* spec, int64_t n, int64_t d, const int64_t* host_seg_len,
                          int n_scales, float log_eps, uint32_t* logq, void* image,
                          float* sq_mean, topo_stream_t stream);
int topo_distance_rows(const uint32_t* logq, const void* image, const float* sq_mean, int64_t n,
                       const int64_t* host_seg_len, int n_scales, int64_t row_begin, int64_t row_end,
                       int64_t col_begin, int64_t col_end, float* workspace, float* out, topo_stream_t stream);

/* One block of the sweep with rows and columns taken from DIFFERENT prepared sets (streaming: a resident row block
 * against column blocks prepared on the fly, precompute_distances.py:89-115 at collection sizes whose spectra do not fit
 * in HBM).  row_* hold clips [row_global0, row_global0 + n_rows) of the collection, col_* clips [col_global0, ...); the
 * collection-wide indices decide which clip of a pair is the lower one (its mean square is the normaliser, :106-110) and
 * where the diagonal is.  out[i * ld_out + j], i < n_rows, j < n_cols. */
int topo_distance_block(const uint32_t* row_logq, const void* row_image, const float* row_sq_mean, int64_t n_rows,
                        int64_t row_global0, const uint32_t* col_logq, const void* col_image, const float* col_sq_mean,
                        int64_t n_cols, int64_t col_global0, const int64_t* seg_len, int n_scales, float* workspace,
                        float* out, int64_t ld_out, topo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * (f)1. The decoder's cross-attention over the stage's compact rows (decoder.py:58-63, 144-162: nn.MultiheadAttention,
 * head dimension 16, queries = the q_len rows made from a sample's vertices, memory = its active edges, triangles and
 * tetrahedra).  q [batch, q_len, C], C = 16 * heads, already projected; k, v [rows, C] = the projected memory rows of ALL
 * samples in the concatenated compact layout (rank 1 rows of every sample, then rank 2, then rank 3); seg [batch][3][2] int32 =
 * (first row, row count) of sample b's run inside each rank: the memory is neither padded nor copied and no mask exists.
 *   fwd: out [batch, q_len, C] = softmax(q k^T / 4) v per (sample, head); lse2 [batch, heads, q_len] = log2 of the softmax
 *        denominator in the scaled log2 domain (kept for the backward)
 *   bwd: dq, dk, dv (dk / dv rows outside every run are not written); d_row [batch, heads, q_len] scratch; max_run_len = the
 *        longest run of any sample (sizes the key-parallel grid).  Two deterministic passes, no atomics.
 * All buffers 16-byte aligned, fp32.
 * ------------------------------------------------------------------------------------------- */
int topo_cross_attention_fwd(const float* q, const float* k, const float* v, const int32_t* seg, int64_t batch,
                             int64_t q_len, int heads, float* out, float* lse2, topo_stream_t stream);
int topo_cross_attention_bwd(const float* q, const float* k, const float* v, const int32_t* seg, const float* out,
                             const float* lse2, const float* d_out, int64_t batch, int64_t q_len, int heads,
                             int64_t max_run_len, float* d_row, float* dq, float* dk, float* dv, topo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TOPO_B200_H */
