/* b200rec.h -- C ABI of libb200rec.so: the ReChorus training hot path as sm_100a CUDA kernels.
 *
 * The reference (THUwangcy/ReChorus) has no FFI: its hot path is Python calling ATen.  Each entry point below
 * names the reference code it replaces (paths relative to the reference's src/).  The Python plugin layer
 * (rechorus_b200/) binds these with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (e.g. torch.Tensor.data_ptr()), row-major,
 *     contiguous; tables/activations float32, ids int64 exactly as the reference's collate emits them
 *     (models/BaseModel.py:135-152); row pointers must be 16-byte aligned and d % 4 == 0.
 *   - all work is enqueued on `stream` (a cudaStream_t); no entry point synchronises or allocates.
 *     Scratch memory is an explicit caller-provided workspace sized by the matching *_workspace_bytes().
 *   - return value: 0 = ok; >0 = cudaError_t; <0 = B2R_E_* below.  b2r_last_error() gives the message
 *     (thread-local).  Out-of-range ids never read out of bounds: they are clamped to row 0 and counted in
 *     the device-side int32 `err_flag` the caller passes (may be NULL); the reference raises an ATen index
 *     error in that case, the plugin raises IndexError when it polls the flag.
 */
#ifndef B200REC_H
#define B200REC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2R_VERSION 101          /* major*10000 + minor*100 + patch */

#define B2R_E_BADARG   (-1)      /* null pointer, d % 4 != 0, negative size ... */
#define B2R_E_WORKSPACE (-2)     /* workspace too small */
#define B2R_E_UNSUPPORTED (-3)   /* shape outside what the kernels were built for */

typedef void* b2r_stream_t;      /* cudaStream_t */

#if defined(__GNUC__)
#define B2R_API __attribute__((visibility("default")))
#else
#define B2R_API
#endif

B2R_API int         b2r_version(void);
B2R_API const char* b2r_last_error(void);
/* sm count / compute capability of the current device (148 / 10.0 on B200) */
B2R_API int         b2r_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Measurement hooks for bench.py (no effect on results):
 *  - b2r_launch_count(): kernels launched by this library so far (a CUB device-wide call counts as one);
 *  - b2r_profile_arm(tag, ev_start, ev_stop): the next launch of the tagged kernel inside a whole-step entry
 *    point is bracketed by cudaEventRecord on its own stream (one-shot; pass NULLs to disarm). */
#define B2R_PROF_TAGS 8
#define B2R_PROF_SCORE_FWD   0
#define B2R_PROF_SCORE_BWDQ  1
#define B2R_PROF_SEGMENT_I   2
#define B2R_PROF_SEGMENT_U   3
#define B2R_PROF_PLAN_I      4
#define B2R_PROF_LOSS        5
B2R_API long long b2r_launch_count(void);
B2R_API int       b2r_profile_arm(int tag, void* ev_start, void* ev_stop);

/* ------------------------------------------------------------------------------------------------
 * Scoring: pred[b,c] = < Q[qid[b]], T[ids[b,c]] >
 * replaces  models/general/BPRMF.py:39-42  (embedding x2, mul, sum)  with Q=u_embeddings, qid=user_id
 * and       models/sequential/SASRec.py:80-81                        with Q=user state [B,d], qid=NULL
 * ids [B,C] int64, pred [B,C] float32.  n_q / n_t = number of rows in Q / T (for range checks).
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_rowdot_fwd(const float* Q, const int64_t* qid, int64_t n_q,
                   const float* T, const int64_t* ids, int64_t n_t,
                   float* pred, int B, int C, int d, int32_t* err_flag, b2r_stream_t stream);

/* Backward of the scoring w.r.t. the query side: dQ[b,:] = sum_c g[b,c] * T[ids[b,c],:]   (dense [B,d];
 * replaces the mul/sum backward half of loss.backward(), helpers/BaseRunner.py:205).  Fixed summation
 * order -> bit-reproducible. */
B2R_API int b2r_rowdot_bwd_query(const float* g, const float* T, const int64_t* ids, int64_t n_t,
                         float* dQ, int B, int C, int d, b2r_stream_t stream);

/* Plain row gather out[r,:] = T[ids[r],:]  (nn.Embedding forward, e.g. NeuMF.py:63-66, SASRec.py:59) */
B2R_API int b2r_gather_rows(const float* T, const int64_t* ids, int64_t n_t, float* out, int64_t n, int d,
                    int32_t* err_flag, b2r_stream_t stream);
/* out[r*out_ld + 0..d) = T[ids[r / ids_div]]: writes into a column block of a wider activation matrix and
 * repeats each id ids_div times (NeuMF.py:61 repeats user ids over the candidates; :69 concatenates) */
B2R_API int b2r_gather_rows_strided(const float* T, const int64_t* ids, int64_t n_t, float* out, int out_ld,
                            int64_t n, int d, int ids_div, int32_t* err_flag, b2r_stream_t stream);

/* Flat pair scoring out[e] = < Q[qidx[e]], T[rows[e]] >, rows[e] < 0 = unused slot (score 0).  The shard owner's
 * half of BPRMF.py:39-42 when the item table is row-range sharded across GPUs (BASELINE config 5): the pairs are
 * what the all-to-all delivered, Q is the all-gathered user-vector block. */
B2R_API int b2r_pairdot_fwd(const float* Q, const int64_t* qidx, int64_t n_q, const float* T, const int64_t* rows,
                    int64_t n_t, float* out, int64_t n, int d, int32_t* err_flag, b2r_stream_t stream);

/* Groundwork, opt-in (B2R_SHARD_P2P=1), not used by the default paths of this round: b2r_pairdot_fwd whose result for
 * pair e is stored at out_tab[e / seg][e % seg] -- out_tab is a DEVICE array of n / seg pointers, typically the
 * peer-mapped receive buffers of the ranks the pairs came from (torch.distributed._symmetric_memory), so the scores
 * reach their requesters by NVLink stores from inside the kernel instead of an all-to-all. */
B2R_API int b2r_pairdot_fwd_p2p(const float* Q, const int64_t* qidx, int64_t n_q, const float* T, const int64_t* rows,
                                int64_t n_t, float* const* out_tab, int64_t seg, int64_t n, int d, int32_t* err_flag,
                                b2r_stream_t stream);

/* out[key[e],:] = sum over each run of consecutive valid pairs sharing key[e] of coef[e] * T[rows[e],:]  (rows < 0 =
 * unused slot; every key occupies one contiguous run, as the stable owner-bucketing of the sharded exchange
 * guarantees; out rows without pairs are left untouched).  The shard owner's half of dQ = sum_c g * I[id]. */
B2R_API int b2r_pair_runs_sum(const int64_t* key, const int64_t* rows, const float* coef, const float* T, int64_t n_t,
                      float* out, int64_t n_out, int64_t n, int d, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BPR loss + closed-form gradient (models/BaseModel.py:175-189; formula SURVEY.md A.4).
 * pred [B,C] (column 0 = positive).  loss_out: 1 float.  grad_pred [B,C] = d loss / d pred (may be NULL).
 * row_ws: B floats of scratch.  Deterministic (fixed-order reduction).
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_bpr_loss(const float* pred, float* loss_out, float* grad_pred, float* row_ws,
                 int B, int C, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Index plan: the sorted, de-duplicated view of the ids a batch touches in one table.  It replaces the
 * accumulation half of ATen's embedding_dense_backward (the autograd node behind nn.Embedding,
 * BPRMF.py:31-32) and is what makes the scatter deterministic: every unique row gets ONE owner that sums
 * its contributions in ascending position order.
 *   in : ids[n] int64 in [0, n_rows)
 *   out: sorted_key[n] uint32 (row ids ascending), sorted_pos[n] uint32 (original flat position, stable),
 *        seg_start[n] int32 (index into sorted_* where unique row u starts), n_uniq (device int32)
 * ---------------------------------------------------------------------------------------------- */
B2R_API size_t b2r_plan_workspace_bytes(int64_t n, int64_t n_rows);
B2R_API int b2r_plan_build(const int64_t* ids, int64_t n, int64_t n_rows,
                   uint32_t* sorted_key, uint32_t* sorted_pos, int32_t* seg_start, int32_t* n_uniq,
                   void* ws, size_t ws_bytes, int32_t* err_flag, b2r_stream_t stream);
/* same, but positions p < ignore_n whose id equals ignore_id are dropped from the gradient (they get the
 * sentinel key n_rows, which b2r_segment_apply skips): the right-padding of SASRec histories, whose
 * gradient is exactly zero in the reference (SASRec.py:74 masks them) */
B2R_API int b2r_plan_build_ex(const int64_t* ids, int64_t n, int64_t n_rows, int64_t ignore_id, int64_t ignore_n,
                      uint32_t* sorted_key, uint32_t* sorted_pos, int32_t* seg_start, int32_t* n_uniq,
                      void* ws, size_t ws_bytes, int32_t* err_flag, b2r_stream_t stream);

/* One contribution stream into a table gradient: position p (0 <= p < n) contributes
 *      coef[p] * src[row(p), :]      with   row(p) = src_id ? src_id[p / div] : p / div
 * BPRMF item table : src = u_embeddings, coef = grad_pred, src_id = user_id, div = C     (dI = g * u)
 * BPRMF user table : src = dQ [B,d],     coef = NULL(=1),  src_id = NULL,    div = 1
 * NeuMF / SASRec   : src = dense gradient of the gathered activations, coef = NULL, div = 1            */
typedef struct {
    const float*   src;
    const float*   coef;
    const int64_t* src_id;
    int64_t        n;
    int32_t        div;
    int32_t        ld;             /* leading dimension of src in floats (0 = d) */
} b2r_grad_source;

/* Optimizer applied to the touched rows (helpers/BaseRunner.py:110-114 builds torch.optim.<name>;
 * here the update is row-sparse / lazy: untouched rows do not move).  kind: 0 SGD, 1 Adam, 2 Adagrad.
 * Coupled L2 (weight_decay) as torch.optim does: g += wd * w.  bc1/bc2 = 1 - beta^t for Adam. */
typedef struct {
    int32_t kind;
    float   lr, beta1, beta2, eps, weight_decay, bc1, bc2;
    int32_t state_ld;   /* row stride (floats) of the m / v arrays in the row-sparse kernels; 0 = d.  2*d lets a table
                           keep m and v interleaved per row ([n_rows][2][d]): one 2*4d-byte burst instead of two */
    const float* clock; /* NULL: bc1 / bc2 above are used.  Else a DEVICE array float[4] kept by b2r_optim_tick:
                           [0] = step count t, [1] = lr / (1 - beta1^t), [2] = 1 / sqrt(1 - beta2^t); the kernels read the
                           Adam step size and bias correction from there, so an enqueued / graph-captured optimizer
                           launch needs no per-step host parameters */
} b2r_optim;

/* advance a device-side optimizer clock by one step (see b2r_optim.clock); one thread, double precision.  The betas are
 * doubles (torch computes 1 - beta ** t from the python floats); the bias corrections are rounded to float before use,
 * as the host-parameter route (bc1 / bc2 above) does, so both routes take bit-identical step sizes */
B2R_API int b2r_optim_tick(float* clock, float lr, double beta1, double beta2, b2r_stream_t stream);

/* Segment reduce over a plan built on the concatenation of up to two sources' ids
 * (positions [0, s0.n) belong to s0, [s0.n, s0.n + s1.n) to s1; s1 may be NULL).
 *   mode 0: write row-sparse gradient: uniq_rows[u] (int64) and grad_rows[u,:]  (u < *n_uniq)
 *   mode 1: add into a dense gradient table dense[row,:] += (one writer per row, no atomics)
 *   mode 2: apply `opt` in place to W (and state m, v) for the touched rows -- fused backward+optimizer
 * replaces embedding_dense_backward + grad zero-fill + the embedding part of optimizer.step()
 * (helpers/BaseRunner.py:193,205,206). */
B2R_API int b2r_segment_apply(const uint32_t* sorted_key, const uint32_t* sorted_pos, const int32_t* seg_start,
                      const int32_t* n_uniq, int64_t n, int64_t n_rows, int d,
                      const b2r_grad_source* s0, const b2r_grad_source* s1,
                      int mode, int64_t* uniq_rows, float* grad_rows, float* dense,
                      float* W, float* m, float* v, const b2r_optim* opt, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Bucketed form of plan + segment for the hot path (modes 1 and 2 only; d in {32, 64, 128}):
 *   b2r_bucket_partition : two passes over the ids partition the (row id, position) pairs into buckets of
 *                          2^shift consecutive table rows (~200 pairs each) with L2 atomics;
 *   b2r_bucket_apply     : one CTA per bucket sorts its pairs by (row, position) in shared memory and applies
 *                          mode 1 (dense[row] += gradient row) or mode 2 (optimizer in place), each row's
 *                          contributions summed in ascending position -> same bits on every run.
 * The workspace must have been initialised once with b2r_bucket_workspace_init (zeroed counters); partition
 * leaves the counters zero again, so the same workspace is reusable step after step.
 * ---------------------------------------------------------------------------------------------- */
B2R_API size_t b2r_bucket_workspace_bytes(int64_t n, int64_t n_rows);
B2R_API int b2r_bucket_workspace_init(void* ws, size_t ws_bytes, int64_t n, int64_t n_rows, b2r_stream_t stream);
B2R_API int b2r_bucket_partition(const int64_t* ids, int64_t n, int64_t n_rows, int64_t ignore_id, int64_t ignore_n,
                         void* ws, size_t ws_bytes, int32_t* err_flag, b2r_stream_t stream);
B2R_API int b2r_bucket_apply(const void* ws, int64_t n, int64_t n_rows, int d, const b2r_grad_source* s0,
                     const b2r_grad_source* s1, int mode, float* dense, float* W, float* m, float* v,
                     const b2r_optim* opt, b2r_stream_t stream);

/* One table's share of an apply launch (same meaning as the b2r_bucket_apply arguments of those names). */
typedef struct b2r_apply_job {
    const void* ws;              /* the table's plan: workspace filled by b2r_bucket_partition */
    int64_t n, n_rows;
    const b2r_grad_source* s0;
    const b2r_grad_source* s1;   /* optional second contribution stream (NULL: none) */
    float* dense;                /* mode 1 */
    float* W;                    /* mode 2 */
    float* m;
    float* v;
} b2r_apply_job;

/* b2r_bucket_apply for two tables in ONE launch (job b may be NULL): the two updates must be independent -- neither
 * job's sources may alias the other job's W -- e.g. the user and the item table of one BPRMF step, the item gradient
 * reading a saved copy of the user rows.  The small table's update then runs underneath the large one's. */
B2R_API int b2r_bucket_apply_pair(const b2r_apply_job* a, const b2r_apply_job* b, int d, int mode, const b2r_optim* opt,
                                  b2r_stream_t stream);

/* "Direct" plans: the same plan as b2r_bucket_partition in two launches instead of four -- every (row, position) pair is
 * dropped into the fixed-capacity region of its row range with one L2 atomic (pairs beyond a region's capacity go to a
 * spill list), then one kernel sorts every bucket and lists the row heads.  b2r_direct_plan_apply == b2r_bucket_apply on
 * such a plan.  The workspace must be initialised once with b2r_direct_plan_init; build/apply may then alternate. */
B2R_API size_t b2r_direct_plan_workspace_bytes(int64_t n, int64_t n_rows);
B2R_API int b2r_direct_plan_init(void* ws, size_t ws_bytes, int64_t n, int64_t n_rows, b2r_stream_t stream);
B2R_API int b2r_direct_plan_build(const int64_t* ids, int64_t n, int64_t n_rows, int64_t ignore_id, int64_t ignore_n,
                                  void* ws, size_t ws_bytes, int32_t* err_flag, b2r_stream_t stream);
B2R_API int b2r_direct_plan_apply(const void* ws, int64_t n, int64_t n_rows, int d, const b2r_grad_source* s0,
                                  const b2r_grad_source* s1, int mode, float* dense, float* W, float* m, float* v,
                                  const b2r_optim* opt, b2r_stream_t stream);

/* Fast, order-nondeterministic alternative to plan+segment for mode 1 (dense += via red.global.add.v4.f32) */
B2R_API int b2r_scatter_add_atomic(const int64_t* ids, int64_t n_rows, const b2r_grad_source* s, int d,
                           float* dense, int32_t* err_flag, b2r_stream_t stream);

/* Dense optimizer step over a whole parameter tensor (exact torch.optim semantics; used for the small dense
 * parameters of NeuMF/SASRec and for the exact-reference dense-Adam mode of the embedding tables). */
B2R_API int b2r_dense_optim(float* W, const float* grad, float* m, float* v, int64_t numel,
                    const b2r_optim* opt, b2r_stream_t stream);


/* ------------------------------------------------------------------------------------------------
 * Dense layers (fp32 CUDA-core SGEMM with fused epilogues) -- the MLP tower of models/general/NeuMF.py:69-75
 * and the q/k/v + feed-forward Linear layers of utils/layers.py:26-28,106-107.  W is a torch nn.Linear
 * weight [N, K] row-major (its leading dimension may exceed K only through the pointer/ld pairs of X, Y).
 *   fwd        : Y[M,N] = act(X[M,K] W^T + bias),  relu != 0 applies ReLU
 *   bwd_input  : dX[M,K] = (dY * [relu_out > 0]) W          (relu_out = the layer's saved output, or NULL)
 *   bwd_weight : dW[N,K] = (dY * [relu_out > 0])^T X ; dbias[N] = column sums (NULL to skip); split over M
 *                in fixed chunks whose partial sums are added in order -> bit-reproducible
 * relu_out, when given, must share dY's leading dimension.
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_linear_fwd(const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy,
                   int64_t M, int N, int K, int relu, b2r_stream_t stream);
/* Same forward on the tensor cores: tcgen05.mma kind::tf32 with the accumulator in TMEM and a 4-product hi/lo operand
 * split (error ~2^-21 relative, i.e. fp32 class).  K % 32 == 0 (<= 128), N % 16 == 0 (<= 256), contiguous W.  Returns
 * B2R_E_UNSUPPORTED outside that class. */
B2R_API int b2r_linear_fwd_tc(const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy,
                      int64_t M, int N, int K, int relu, b2r_stream_t stream);
/* general form: x_mask (same layout as X, may be NULL) zeroes X where x_mask <= 0 -- with X = dY, x_mask = the layer's
 * saved ReLU output and W = the transposed weight this is the input gradient dX = (dY * [y > 0]) W of a Linear layer */
B2R_API int b2r_linear_tc(const float* X, int ldx, const float* x_mask, const float* W, const float* bias, float* Y,
                  int ldy, int64_t M, int N, int K, int relu, b2r_stream_t stream);
B2R_API int b2r_linear_bwd_input(const float* dY, int lddy, const float* relu_out, const float* W, float* dX, int lddx,
                         int64_t M, int N, int K, b2r_stream_t stream);
B2R_API size_t b2r_linear_bwd_weight_workspace_bytes(int64_t M, int N, int K);
B2R_API int b2r_linear_bwd_weight(const float* dY, int lddy, const float* relu_out, const float* X, int ldx,
                          float* dW, float* dbias, int64_t M, int N, int K, void* ws, size_t ws_bytes,
                          b2r_stream_t stream);

/* The same weight gradient on the tcgen05 tensor cores (kind::tf32 with a hi/lo operand split, accumulator in TMEM; the
 * batch-row index is the MMA's reduction dimension, activations are transposed on their way into shared memory).
 * Shape class: N % 4 == 0, N <= 128, K % 16 == 0, K <= 240; returns B2R_E_UNSUPPORTED otherwise. */
B2R_API size_t b2r_linear_bwd_weight_tc_workspace_bytes(int64_t M, int N, int K);
B2R_API int b2r_linear_bwd_weight_tc(const float* dY, int lddy, const float* relu_out, const float* X, int ldx,
                                     float* dW, float* dbias, int64_t M, int N, int K, void* ws, size_t ws_bytes,
                                     b2r_stream_t stream);

/* y = LayerNorm(x + res) * gamma + beta, eps inside the sqrt, biased variance (utils/layers.py:113,117 via
 * nn.LayerNorm); mean/rstd [rows] are saved for the backward.  Backward returns dz (gradient of x + res, to be
 * used for both addends) and dgamma/dbeta (fixed-order two-stage reduction).  d <= 256 in the backward. */
B2R_API int b2r_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                          float* mean, float* rstd, int64_t rows, int d, float eps, b2r_stream_t stream);
B2R_API size_t b2r_add_layernorm_bwd_workspace_bytes(int64_t rows, int d);
B2R_API int b2r_add_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                          const float* mean, const float* rstd, float* dz, float* dgamma, float* dbeta,
                          int64_t rows, int d, void* ws, size_t ws_bytes, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SASRec sequence kernels (models/sequential/SASRec.py:51-86, utils/layers.py:34-63).
 *   embed_history : x[b,t,:] = I[hist[b,t]] + P[(len[b]-t) * (hist[b,t] > 0)]; pos_out (optional) gets the
 *                   position index used per element                                    (SASRec.py:58-66)
 *   attention_fwd : causal multi-head attention, heads = contiguous d/H chunks, scores / sqrt(d/H),
 *                   softmax over keys j <= i, no output projection; q,k,v rows at ld stride (layers.py:52-63)
 *   attention_bwd : recomputes the probabilities; writes dq, dk, dv rows at ldg stride.  L <= 128.
 *   select_last   : h[b] = y[b, len[b]-1] * (hist[b, len[b]-1] > 0)                    (SASRec.py:74-76)
 *   small_table_grad : dense gradient of a small table shared by many positions (the position table):
 *                   dense_out[id] = sum_{r: ids[r]=id} src[r], fixed order, no atomics.
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_embed_history(const float* I, int64_t n_items, const float* P, int64_t n_pos, const int64_t* hist,
                      const int64_t* lengths, float* x, int64_t* pos_out, int B, int L, int d,
                      int32_t* err_flag, b2r_stream_t stream);
B2R_API int b2r_attention_fwd(const float* q, const float* k, const float* v, int ld, float* ctx, int B, int L, int d,
                      int H, b2r_stream_t stream);
B2R_API int b2r_attention_bwd(const float* q, const float* k, const float* v, int ld, const float* dctx, float* dq,
                      float* dk, float* dv, int ldg, int B, int L, int d, int H, b2r_stream_t stream);
/* The same with dead rows skipped: live[b] (int64, NULL = L) is the number of leading positions of sequence b whose
 * output anything downstream reads -- SASRec uses position len-1 only (SASRec.py:74-81) and attention is causal, so the
 * positions >= len are exact dead work (forward and backward).  Dead rows of ctx / dq / dk / dv are written as zeros. */
B2R_API int b2r_attention_fwd_live(const float* q, const float* k, const float* v, int ld, const int64_t* live, float* ctx,
                                   int B, int L, int d, int H, b2r_stream_t stream);
B2R_API int b2r_attention_bwd_live(const float* q, const float* k, const float* v, int ld, const int64_t* live,
                                   const float* dctx, float* dq, float* dk, float* dv, int ldg, int B, int L, int d, int H,
                                   b2r_stream_t stream);
/* The same function with the per-(row, head) state in registers (csrc/attention_rt.cu): one lane owns one query row (forward,
 * dQ) or one key row (dK, dV) of one head, the other operand arrives as broadcast shared-memory loads.  The forward also
 * writes lse [B, L, H]: log2 of the softmax denominator plus the running max (base-2 domain); the backward recomputes the
 * probabilities from it and takes the forward's output ctx for the softmax-Jacobian row term dO_i . O_i.
 * Covers d/H in {8, 16, 32}, L <= 128, d and ld multiples of 4, 16-byte aligned pointers; anything else returns
 * B2R_E_UNSUPPORTED (callers fall back to b2r_attention_*_live).  Same dead-row contract. */
B2R_API int b2r_attention_fwd_rt(const float* q, const float* k, const float* v, int ld, const int64_t* live, float* ctx,
                                 float* lse, int B, int L, int d, int H, b2r_stream_t stream);
B2R_API int b2r_attention_bwd_rt(const float* q, const float* k, const float* v, int ld, const int64_t* live,
                                 const float* ctx, const float* lse, const float* dctx, float* dq, float* dk, float* dv,
                                 int ldg, int B, int L, int d, int H, b2r_stream_t stream);
/* Attention for ONE query per
 * sequence -- the query at position t* = clamp(lengths[b]-1, 0, L-1), the only position of SASRec's last block whose
 * output is used (models/sequential/SASRec.py:74-81) -- against keys/values 0..t* (the causal row of
 * utils/layers.py:52-63).  q_last, ctx_last, dq_last are [B, d]; k, v, dk, dv are [B, L, d] rows with leading
 * dimension ld / ldg; prob [B, H, L] keeps the softmax row for the backward.  Identical results to b2r_attention_fwd/
 * bwd restricted to that query; rows of dk, dv beyond t* are written as zeros. */
B2R_API int b2r_attention_last_fwd(const float* q_last, const float* k, const float* v, int ld, const int64_t* lengths,
                                   float* ctx_last, float* prob, int B, int L, int d, int H, b2r_stream_t stream);
B2R_API int b2r_attention_last_bwd(const float* q_last, const float* k, const float* v, int ld, const int64_t* lengths,
                                   const float* prob, const float* dctx_last, float* dq_last, float* dk, float* dv,
                                   int ldg, int B, int L, int d, int H, b2r_stream_t stream);
B2R_API int b2r_select_last(const float* y, const int64_t* hist, const int64_t* lengths, float* h, int B, int L,
                    int d, b2r_stream_t stream);
B2R_API int b2r_select_last_bwd(const float* dh, const int64_t* hist, const int64_t* lengths, float* dy, int B, int L,
                        int d, b2r_stream_t stream);
B2R_API size_t b2r_small_table_grad_workspace_bytes(int64_t n, int n_rows, int d);
B2R_API int b2r_small_table_grad(const float* src, int ld, const int64_t* ids, int64_t n, int n_rows, int d,
                         float* dense_out, void* ws, size_t ws_bytes, b2r_stream_t stream);

/* out[r,k] = a[r,k] * w[k]   and   out[k] = sum_r a[r,k] * b[r,k]   (the GMF half of NeuMF's final Linear,
 * models/general/NeuMF.py:68,74-75, and its weight gradient) */
B2R_API int b2r_colscale(const float* a, const float* w, float* out, int64_t rows, int d, b2r_stream_t stream);
B2R_API int b2r_colsum_prod(const float* a, const float* b, float* out, int64_t rows, int d, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row f3, batch production on the device (opt-in): the per-epoch negative sampling of
 * models/BaseModel.py:206-214.  out[i*K + j] is uniform over the items of [1, n_items) that are not in
 * the training clicks of user_ids[i] (CSR: clicked_ptr [n_users+1], clicked_items sorted ascending and unique per user,
 * all within [1, n_items)).  No rejection loop: with m clicks the user has A = n_items-1-m allowed items; r =
 * floor(x * A / 2^32), x = word 0 of Philox4x32-10(counter = (lo32(i*K+j), hi32(i*K+j), 0, epoch), key = seed), and
 * the (r+1)-th allowed item is found by one binary search over the click row.  Same distribution as the reference,
 * NOT the same stream (NumPy's global Mersenne Twister consumed by a data-dependent Python loop).
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_sample_negatives(const int64_t* user_ids, int64_t N, int K, const int64_t* clicked_ptr,
                                 const int64_t* clicked_items, int64_t n_users, int64_t n_items, uint64_t seed,
                                 uint32_t epoch, int64_t* out, int32_t* err_flag, b2r_stream_t stream);
/* One training batch of a GeneralModel assembled on the device (GeneralModel.Dataset._get_feed_dict + collate_batch,
 * models/BaseModel.py:192-203,135-152): batch row t is training row perm[start + t] (perm NULL: start + t);
 * out_uid[t] = users[row], out_iid[t, 0] = items[row], out_iid[t, 1 + k] = neg[row, k]. */
B2R_API int b2r_collate_general(const int64_t* users, const int64_t* items, const int64_t* neg, const int64_t* perm,
                                int64_t start, int Bn, int K, int64_t* out_uid, int64_t* out_iid, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Groundwork for the exact dense-Adam mode (not used by the default paths of this round): advance rows of a table whose
 * Adam state is kept row-sparsely through the optimizer steps they skipped, exactly as torch.optim.Adam over the whole
 * table (helpers/BaseRunner.py:110-114,206) would have moved them with a zero data gradient (g = weight_decay * w).
 * last [n_rows] int32 holds the step each row is up to date with.  rows == NULL: all n_rows rows (flush); otherwise n
 * unique row ids.  Every processed row is brought to step `upto` and stamped max(upto, stamp) -- pass stamp = upto + 1
 * when the step-(upto+1) update with b2r_bucket_apply / b2r_segment_apply follows.  opt carries lr, betas, eps,
 * weight_decay, state_ld (kind must be Adam; bc1/bc2 are not used: every skipped step has its own).
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_adam_exact_advance(const int64_t* rows, int64_t n, int64_t n_rows, int d, float* W, float* m, float* v,
                                   int32_t* last, int upto, int stamp, const b2r_optim* opt, int32_t* err_flag,
                                   b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation (row f2): integer ranks of the ground-truth item on the device.
 *
 * b2r_gt_rank:        rank[r] = #{c : pred[r*ld + c] >= pred[r*ld]}  -- helpers/BaseRunner.py:63 (column 0 is the
 *                     ground truth, ties count against it, rank >= 1; NaN compares false as in NumPy).
 * b2r_rank_histogram: hist[k] = #{r : rank[r] == k} for k <= kmax, hist[kmax+1] = #{rank > kmax}; hist has kmax+2
 *                     int64 slots.  HR@k and NDCG@k (BaseRunner.py:66-74) are sums over hist[1..k].
 * b2r_rank_all_items: the test_all protocol (BaseModel.py:194-198: candidates = [target] + arange(1, n_items),
 *                     scored by a dot product with the query row -- BPRMF.py:42, SASRec.py:81 -- then
 *                     preds[row, clicked item] = -inf, BaseRunner.py:244-251) without materialising the
 *                     [B, n_items] score matrix: rank[b] = 1 + #{1 <= j < n_items, (b,j) not masked :
 *                     <Q[b], I[j]> >= <Q[b], I[target[b]]>}.  Q [B, ldq] are the query rows (user vectors or
 *                     sequence states), (mask_row[e], mask_item[e]) the n_mask unique (batch row, item id) pairs to
 *                     exclude, s0 [B] scratch that receives the targets' scores.  The unmasked target column counts
 *                     (it scores equal to column 0), exactly as in the reference.
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_gt_rank(const float* pred, int64_t N, int64_t C, int64_t ld, int64_t* rank, b2r_stream_t stream);
B2R_API int b2r_rank_histogram(const int64_t* rank, int64_t N, int kmax, int64_t* hist, b2r_stream_t stream);
B2R_API int b2r_rank_all_items(const float* Q, int ldq, const float* I, const int64_t* target, int B, int64_t n_items,
                               int d, const int64_t* mask_row, const int64_t* mask_item, int64_t n_mask, float* s0,
                               int64_t* rank, int32_t* err_flag, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused BPRMF forward + loss + query-side backward: every candidate row is read from HBM once and stays in
 * registers between scoring and the gradient.  Outputs: grad_pred [B,C] (= d loss / d pred, loss = mean of the
 * per-sample losses), row_loss [B] (per-sample -log S), dQ [B,d] (= sum_c g[b,c] I[iid[b,c]]), pred [B,C]
 * optional (NULL to skip).  Returns B2R_E_UNSUPPORTED when no fused variant exists for (d, C) -- callers then
 * use b2r_rowdot_fwd / b2r_bpr_loss / b2r_rowdot_bwd_query.  Replaces BPRMF.py:39-42, BaseModel.py:182-185
 * and their autograd (BaseRunner.py:205) in one kernel.
 * ---------------------------------------------------------------------------------------------- */
B2R_API int b2r_bprmf_fused_fwd_bwd(const float* U, const int64_t* uid, int64_t n_users, const float* I,
                            const int64_t* iid, int64_t n_items, float* pred, float* grad_pred, float* row_loss,
                            float* dQ, int B, int C, int d, int32_t* err_flag, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * One whole BPRMF training step enqueued from C (no Python between kernels) on a step context:
 *   fused gather/score/loss/dQ -> row-sparse fused optimizer on I, then on U (lazy rule, see b2r_optim).
 * Replaces one iteration of the hot loop helpers/BaseRunner.py:193-206 for models/general/BPRMF.py (the
 * per-row candidate shuffle of :187-191 is a mathematical no-op for this model and is not performed).
 * The context owns a side stream on which the index plans (sorts) are built; passing the NEXT batch's ids
 * (next_uid/next_iid, device pointers that must stay unmodified until the next call) lets the sort of step
 * t+1 overlap the HBM-bound kernels of step t, like a data loader's prefetch.  Pass NULLs to disable.
 * The workspace (b2r_bprmf_step_workspace_bytes) is caller-owned and must outlive the context.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float*  U;  float* I;            /* [n_users, d], [n_items, d]                       */
    float*  Um; float* Uv;           /* optimizer state for U (Adam: m, v; Adagrad: v)   */
    float*  Im; float* Iv;           /* optimizer state for I                            */
    int64_t n_users, n_items;
    int32_t d, _pad;
} b2r_bprmf_tables;

B2R_API size_t b2r_bprmf_step_workspace_bytes(int B, int C, int d, int64_t n_users, int64_t n_items);
B2R_API int b2r_bprmf_ctx_create(void** ctx_out, int B, int C, int d, int64_t n_users, int64_t n_items, void* ws,
                         size_t ws_bytes);
B2R_API int b2r_bprmf_ctx_destroy(void* ctx);
/* forget the prefetched plan (call at the start of an epoch and after an aborted one): a prefetched plan is matched to
 * the next batch by its id pointers, which a caching allocator may hand out again for a different batch */
B2R_API int b2r_bprmf_ctx_reset(void* ctx);
B2R_API int b2r_bprmf_train_step(void* ctx, const b2r_bprmf_tables* t, const int64_t* uid, const int64_t* iid,
                         const int64_t* next_uid, const int64_t* next_iid, const b2r_optim* opt,
                         float* loss_out, int32_t* err_flag, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * List-wise ranking losses of the impression models: value and closed-form gradient of ImpressionModel.loss
 * (models/BaseImpressionModel.py:44-128).  pred [B, Cn] float32, target [B, Cn] int64 (1 clicked, 0 shown, -1 padding);
 * columns < max_pos are the positive slots (:55-58; the reference also reads column max_pos: max_pos < Cn).
 * kind: 0 'BPR' (reweight between sigmoid and log, :82-85), 1 'BPR..after' (:73-75), 2 'BPR..before' (:76-78), each
 * with hard = 0/1 ('hard' in the name, :63-68); 3 'listnet' (:88-97), 4 'softmaxCE' (:99-110), 5 'attention_rank'
 * (:112-128).  loss_out: 1 float; grad_pred [B, Cn] = d loss / d pred (may be NULL).  ws: b2r_listwise_workspace_bytes(B).
 * ---------------------------------------------------------------------------------------------- */
B2R_API size_t b2r_listwise_workspace_bytes(int B);
B2R_API int b2r_listwise_loss(const float* pred, const int64_t* target, int B, int Cn, int max_pos, int kind, int hard,
                              float* loss_out, float* grad_pred, void* ws, size_t ws_bytes, b2r_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded tables (BASELINE config 5): the exchange done by kernels over peer-mapped memory (symmetric-memory
 * buffers of the ranks of one box; NVLink loads / stores) instead of collectives between kernels.  No reference
 * counterpart (the reference is single-device); contract: SURVEY.md 8(e).  Pointer tables are HOST arrays of W device
 * pointers (W <= 16).  Ordering between ranks is the caller's (signal-pad barriers between the phases).
 *
 * b2r_route_ids: stable partition of one table's ids [B, C] by owner (id / rows_per): pair (b, c) goes to slot
 *   s of its owner o, slots counted in (sample, candidate) order; written through peer_rows[o][s] = local row and
 *   peer_q[o][s] = q_base + b (the row of the replicated query block); the unused tail of each region gets row -1;
 *   slot_of[b*C + c] = o * cap + s; dest_total[o] = pairs sent to o; *overflow += 1 if some dest_total > cap.
 * b2r_serve_rows: for every request e with req_rows[e] >= 0: row req_rows[e] of T is stored into row req_q[e] of EVERY
 *   rank's block peer_dst[k] (gather + all-gather in one pass).
 * b2r_scatter_f32_to_peers / b2r_scatter_rows_to_peers: val[i] * scale (resp. row i of src) -> peer_dst[slot_of[i] / cap]
 *   at element (row) slot_of[i] % cap.
 * b2r_sum_rows_from_peers: out[i] = sum_k peer_src[k][offset + i], k ascending (fixed order -> deterministic).
 * ---------------------------------------------------------------------------------------------- */
B2R_API size_t b2r_route_workspace_bytes(int B, int W);
B2R_API int b2r_route_ids(const int64_t* ids, int B, int C, int W, int64_t rows_per, int64_t n_rows,
                          const void* const* peer_rows, const void* const* peer_q, int64_t q_base, int cap,
                          int* slot_of, int* dest_total, int* overflow, void* ws, size_t ws_bytes, int32_t* err_flag,
                          b2r_stream_t stream);
B2R_API int b2r_serve_rows(const float* T, int64_t n_t, const int64_t* req_rows, const int64_t* req_q, int64_t n,
                           const void* const* peer_dst, int W, int d, int32_t* err_flag, b2r_stream_t stream);
B2R_API int b2r_scatter_f32_to_peers(const float* val, const int* slot_of, int64_t n, const void* const* peer_dst, int W,
                                     int cap, float scale, b2r_stream_t stream);
B2R_API int b2r_scatter_rows_to_peers(const float* src, const int* slot_of, int64_t n, const void* const* peer_dst, int W,
                                      int cap, int d, b2r_stream_t stream);
B2R_API int b2r_sum_rows_from_peers(const void* const* peer_src, int W, int64_t offset_floats, float* out,
                                    int64_t n_floats, b2r_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H */
