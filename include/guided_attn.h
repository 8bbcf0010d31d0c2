/* guided_attn.h -- C ABI of libguidedattn.so: the B200 (sm_100a) kernels of the cross-attention guidance path.
 *
 * This is the drop-in boundary.  Every entry point replaces one reference op sequence (paths relative to the
 * jackBonadies/Guided-Attention tree) and is what a maintainer's FFI stub (ctypes) binds.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All `*_dev` / unqualified data pointers are DEVICE pointers owned
 *     by the caller (PyTorch); the library never allocates, frees or synchronises.  `*_host` pointers are small HOST
 *     arrays that are copied into kernel parameters at launch (no H2D memcpy, no staging buffer).
 *   - every kernel is enqueued on the caller-supplied stream (`cudaStream_t` passed as void*).
 *   - return value: GA_OK or a negative error code; `ga_last_error()` gives a thread-local message.  No C++ exception
 *     crosses the boundary.
 *   - tensors are dense row-major; attention operands use the projection layout (batch, tokens, heads*head_dim), i.e.
 *     exactly what `attn.to_q/to_k/to_v` produce -- the `head_to_batch_dim` permutes of the reference
 *     (utils/ptp_utils.py:77-79) are folded into the kernels' indexing.
 */
#ifndef GUIDED_ATTN_H_
#define GUIDED_ATTN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GA_ABI_VERSION 7

typedef void* ga_stream_t; /* cudaStream_t */

enum ga_status { GA_OK = 0, GA_ERR_BAD_ARG = -1, GA_ERR_UNSUPPORTED = -2, GA_ERR_ALIGNMENT = -3, GA_ERR_CUDA = -4 };
enum ga_dtype { GA_F32 = 0, GA_F16 = 1, GA_BF16 = 2 };
/* kernel variant: AUTO picks TCGEN05 for 16-bit operands it supports, SIMT otherwise (fp32 is SIMT-only: exact fp32).
 * TCGEN05 itself picks between the single-shot kernel (few work items: latency) and the persistent pipelined kernel
 * (many work items: bandwidth); _SINGLE / _PIPE force one of the two (tests, A/B measurements). */
enum ga_impl { GA_IMPL_AUTO = 0, GA_IMPL_SIMT = 1, GA_IMPL_TCGEN05 = 2, GA_IMPL_TCGEN05_SINGLE = 3,
               GA_IMPL_TCGEN05_PIPE = 4 };
/* reference utils/helpers.py:10-13 (AnnotationType) */
enum ga_token_kind { GA_TOKEN_COOR = 0, GA_TOKEN_BOX = 1, GA_TOKEN_KEYWORD = 2 };

#define GA_MAX_ACC_SLICES 32 /* per-layer accumulators one tail launch can reduce */
#define GA_MAX_TOKENS 24     /* tracked tokens per launch */
#define GA_MAX_BOXES 32      /* boxes per rasteriser launch */
#define GA_MAX_CTX 128       /* text-context length the attention kernels accept (SD: 77) */

/* per-token outputs of the guidance tail, `stats[token][GA_STATS]` */
enum ga_stat {
  GA_STAT_MAX = 0,      /* max of the smoothed map          (reference pipeline_guided_attention.py:255)            */
  GA_STAT_SUM = 1,      /* sum of the smoothed map          (:263)                                                  */
  GA_STAT_COL = 2,      /* sum (jj+.5) p                    (:264-268)                                              */
  GA_STAT_ROW = 3,      /* sum (ii+.5) p                                                                             */
  GA_STAT_INSIDE = 4,   /* inside-box loss                  (utils/helpers.py:247-277)                              */
  GA_STAT_OUTSIDE = 5,  /* outside-box loss                                                                          */
  GA_STAT_SCALED = 6,   /* token's term of the total loss   (pipeline_guided_attention.py:405-438)                  */
  GA_STAT_UNSCALED = 7, /* token's threshold quantity       (:424, :411)                                            */
  GA_STAT_HINGE_IN = 8, /* strict mode only: sum_in  w [p < 1/n_in] p   (saved for the backward)                    */
  GA_STAT_HINGE_OUT = 9,/* strict mode only: sum_out w [p > 0] p                                                    */
  GA_STAT_NINSIDE = 10, /* number of inside pixels                                                                  */
  GA_STAT_CENTER = 11,  /* centring loss |col-cx*res|/(res-1) + 4|row-cy*res|/(res-1)   (:390-395)                  */
  /* statistics of the UN-smoothed renormalised map A[:, :, token] -- what the custom-loss plug-ins read
   * (run.py:157-161, 203-225: CustomLossBase.get_map_for_token + calc_weighted_center); differentiable, so the built-in
   * toLeftOf loss is a few scalar operations on them and its gradient re-enters the tail backward through g_stats   */
  GA_STAT_RAW_SUM = 12, /* sum of the raw map                                                                         */
  GA_STAT_RAW_COL = 13, /* sum (jj+.5) A / sum A                                                                      */
  GA_STAT_RAW_ROW = 14, /* sum (ii+.5) A / sum A                                                                      */
  GA_STATS = 16
};

/* One tracked text token = one entry of the reference's `config.token_dict` (run.py:81-91). */
typedef struct ga_token {
  int32_t column;      /* token index - first: column of the renormalised map (pipeline_guided_attention.py:228)    */
  int32_t kind;        /* ga_token_kind                                                                              */
  int32_t box;         /* index into masks / weights for BOX tokens, -1 otherwise                                    */
  int32_t group;       /* sub-prompt id (threshold grouping, :358-387)                                               */
  float target_x;      /* float32(cx * res): box centre or crosshair x at map resolution (:392)                      */
  float target_y;      /* float32(cy * res)                                                                          */
  float center_weight; /* weight of the centring term in the scaled loss: bb_center_weight (BOX) or 1 (COOR)         */
  float group_weight;  /* 1, or 1/len(sub-prompt) when sub_prompt_avg_within (:376-379)                              */
} ga_token_t;

typedef struct ga_tail_params {
  int32_t res;          /* map side: res*res == n_query of the accumulators                                          */
  int32_t n_ctx;        /* text-context length T (77)                                                                */
  int32_t first, last;  /* tokens [first, last) enter the renormalisation softmax; reference: 1 and T-1 (or n_ids-1) */
  int32_t n_tokens;     /* <= GA_MAX_TOKENS                                                                          */
  int32_t n_groups;
  int32_t strict;       /* curHyperParams["strict"] (utils/helpers.py:250)                                           */
  int32_t smooth;       /* smooth_attentions (pipeline_guided_attention.py:251)                                      */
  float w1d[3];         /* separable smoothing taps, normalised to sum 1 (utils/gaussian_smoothing.py:37-43)         */
  float temperature;    /* 100 (:218)                                                                                */
  float inv_count;      /* 1 / (number of (n_query, n_ctx) head-maps summed into the accumulators)                   */
  float inside_scale;   /* curHyperParams["inside_loss_scale"]                                                       */
  float outside_scale;  /* curHyperParams["outside_loss_scale"] * 3 (:426)                                           */
  int32_t n_samples;    /* independent samples evaluated by one launch; 1 = the reference's semantics (everything
                           averaged into one map).  Sample s owns slices [s*slices_i, (s+1)*slices_i) of accumulator i */
} ga_tail_params_t;

/* ---- library ---------------------------------------------------------------------------------------------------- */
int ga_version(void);               /* GA_ABI_VERSION */
const char* ga_last_error(void);    /* thread-local, never NULL */
int ga_device_supported(int device); /* 1 if `device` is compute capability 10.x, 0 otherwise, <0 on error */

/* ---- K1: fused cross-attention forward ------------------------------------------------------------------------------
 * Replaces AttendExciteCrossAttnProcessor.__call__'s  baddbmm -> softmax -> controller(P) -> bmm  sequence
 * (utils/ptp_utils.py:77-85, 97-146) together with AttentionStore.forward (:226-230) and the per-layer part of
 * aggregate_attention (:273-289): P is never written to HBM.
 *   q (B, N, H*d)   k, v (B, T, H*d)   o (B, N, H*d)   all `dtype`
 *   lse (B, H, N) fp32: row log-sum-exp of scale*q.k, saved for the backward
 *   acc (B, N, T) fp32 or NULL: acc[b] = sum over heads of P[b, h]  (deterministic, no atomics)
 * T <= GA_MAX_CTX, d % 8 == 0, d <= 256. */
int ga_cross_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, float* acc, int batch,
                      int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype, int impl,
                      ga_stream_t stream);

/* ---- K2: fused cross-attention backward ------------------------------------------------------------------------------
 * autograd of K1 with the attention-map gradient injected:
 *   dP = dO V^T + d_acc[b]      dS = P o (dP - rowsum(P o dP))      dQ = scale dS K
 *   d_o (B, N, H*d) `dtype`;  d_acc fp32 (N, T) slices, rows `d_acc_row_stride` elements apart (>= T; the tail kernel
 *   pads rows to a multiple of 4 floats so they can be read with 16-byte loads), slices `d_acc_batch_stride` elements
 *   apart (0 = one slice broadcast over the batch), or NULL;  d_q (B, N, H*d) `dtype`.
 *   d_k, d_v: fp32 (B, T, H*d), ACCUMULATED with atomics (caller zero-fills), or NULL (the reference only
 *   differentiates w.r.t. the latents, pipeline_guided_attention.py:466). */
int ga_cross_attn_bwd(const void* q, const void* k, const void* v, const float* lse, const void* d_o,
                      const float* d_acc, int64_t d_acc_batch_stride, int d_acc_row_stride, void* d_q, float* d_k,
                      float* d_v, int batch,
                      int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype, int impl,
                      ga_stream_t stream);

/* ---- optional additive score bias of the cross-attention kernels ------------------------------------------------------
 * Two independent parts, either may be absent; S' = scale*q.k + mask + pww, softmax and everything after it use S'.
 *  - `mask`: the processor's `attention_mask` (utils/ptp_utils.py:135-136), fp32 DEVICE, already broadcast-indexable:
 *    element (b*H + h, n, t) at mask[(b*H + h)*mask_stride_bh + n*mask_stride_n + t] (strides may be 0); NULL = none.
 *  - paint-with-words (utils/ptp_utils.py:113-138, `paint_with_words_stop` > cur_time_step_iter, 77-key layers):
 *    S'[bh, n, t] += w * 0.4 * max(S) * ln(1 + sigma_t) for every BOX token t whose box, rasterised at THIS layer's
 *    resolution (`rect.of_size(hw)`, K5), contains pixel n.  `pww_masks` (pww_count, n_query) u8 DEVICE = those masks,
 *    `pww_column[i]` = the token index (column of S) mask i biases, `pww_coef` = DEVICE scalar w*0.4*ln(1+sigma_t)
 *    (device-resident so that a captured CUDA graph follows the denoising step; 0 switches the bias off),
 *    `pww_smax` = DEVICE 8-byte word written by ga_cross_attn_smax: the global max of S (after `mask`) over
 *    (b, h, n, t), taken in the same launch-wide sense as the reference's `attention_scores.max()`.
 *    The max stays differentiable like in the reference: the backward adds  scale * coef * sum(mask o dS') * K[b*,t*,h*]
 *    to dQ of the one query row (b*, n*, h*) that owns it.
 * Only the SIMT variant implements the bias; GA_IMPL_AUTO selects it when `bias` is non-NULL. */
typedef struct ga_score_bias {
  const float* mask;
  int64_t mask_stride_bh, mask_stride_n;
  const uint8_t* pww_masks;
  const float* pww_coef;
  const unsigned long long* pww_smax;
  int32_t pww_count;                 /* 0 = paint-with-words off; <= GA_MAX_TOKENS */
  int32_t pww_column[GA_MAX_TOKENS];
} ga_score_bias_t;

/* max over (b, h, n, t) of scale*q.k (+ bias->mask when given; `bias` may be NULL) -> *smax (packed: order-preserving
 * float bits << 32 | ~flat index, ties to the lowest index).  Zero-fills *smax on the stream first. */
int ga_cross_attn_smax(const void* q, const void* k, unsigned long long* smax, const ga_score_bias_t* bias_host,
                       int batch, int heads, int n_query, int n_ctx, int head_dim, float scale, int dtype,
                       ga_stream_t stream);

/* ga_cross_attn_fwd / ga_cross_attn_bwd / ga_attn_probs with the score bias (`bias_host` NULL = identical to the
 * plain entry points).  `pww_partials`: caller-owned DEVICE workspace of batch*heads*ceil(n_query/64) floats, needed
 * when bias_host->pww_count > 0 (per-CTA partial sums of the max gradient, reduced in a fixed order: deterministic).
 * With paint-with-words d_k / d_v receive the softmax terms only (the max term's dK is not produced). */
int ga_cross_attn_fwd_ex(const void* q, const void* k, const void* v, void* o, float* lse, float* acc,
                         const ga_score_bias_t* bias_host, int batch, int heads, int n_query, int n_ctx, int head_dim,
                         float scale, int dtype, int impl, ga_stream_t stream);
int ga_cross_attn_bwd_ex(const void* q, const void* k, const void* v, const float* lse, const void* d_o,
                         const float* d_acc, int64_t d_acc_batch_stride, int d_acc_row_stride, void* d_q, float* d_k,
                         float* d_v, const ga_score_bias_t* bias_host, float* pww_partials, int batch, int heads,
                         int n_query, int n_ctx, int head_dim, float scale, int dtype, int impl, ga_stream_t stream);
int ga_attn_probs_ex(const void* q, const void* k, void* probs, const ga_score_bias_t* bias_host, int batch, int heads,
                     int n_query, int n_ctx, int head_dim, float scale, int dtype, ga_stream_t stream);

/* Materialise P (B*H, N, T) in `dtype`, batch-major rows b*H+h, for API compatibility with code that reads the
 * reference's per-head maps (AttentionStore.get_average_attention, utils/ptp_utils.py:245-247).  Off the hot path. */
int ga_attn_probs(const void* q, const void* k, void* probs, int batch, int heads, int n_query, int n_ctx,
                  int head_dim, float scale, int dtype, ga_stream_t stream);

/* ---- self-attention ------------------------------------------------------------------------------
 * Exact softmax(scale Q K^T) V where queries and keys are the same N image tokens; q, k, v, o (B, N, H*d) 16-bit,
 * lse (B, H, N) fp32.  Replaces the attn1 branch of the reference processor (utils/ptp_utils.py:66-93 with
 * encoder_hidden_states=None): the (B*H, N, N) probability tensor is never materialised.  d % 8 == 0, d <= 160. */
int ga_self_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int batch, int heads,
                     int n_tokens, int head_dim, float scale, int dtype, ga_stream_t stream);

/* Backward of ga_self_attn_fwd (autograd of the reference's attn1 branch): P is recomputed from `lse`; two launches
 * (dQ, then dK/dV), no atomics.  o, d_o, d_q, d_k, d_v (B, N, H*d) 16-bit; dvec (B, H, N) fp32 caller-owned workspace
 * (receives rowsum(dO o O)). */
int ga_self_attn_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* d_o,
                     void* d_q, void* d_k, void* d_v, float* dvec, int batch, int heads, int n_tokens, int head_dim,
                     float scale, int dtype, ga_stream_t stream);

/* ---- K5: box-mask rasteriser ------------------------------------------------------------------------------------------
 * masks[i, ii, jj] = inside_box(jj, ii, Rect(box_i, size 1).of_size(res))   (utils/helpers.py:164-173, 28-30)
 * float64, no FMA contraction, the reference's operation order: bit-exact.  boxes_host: n x (x, y, w, h). */
int ga_rasterize_boxes(const double* boxes_host, int n_boxes, int res, double shrink, uint8_t* masks,
                       ga_stream_t stream);

/* ---- K3+K4+K6: guidance tail -------------------------------------------------------------------------------------------
 * Replaces aggregate_attention's mean (utils/ptp_utils.py:287-288), _compute_max_attention_per_index
 * (pipeline_guided_attention.py:201-296), helpers.calculate_bounding_box_losses (utils/helpers.py:215-277) and
 * _compute_loss (:398-451):
 *   Abar = inv_count * sum of accumulator slices;  A = softmax(temperature * Abar[:, first:last]);
 *   per token: 3x3 separable Gaussian with reflect padding, max/argmax, normalise, centre of mass, box losses, loss.
 * acc_host[i] points to n_samples * slices_host[i] consecutive (res*res, n_ctx) fp32 slices (one K1 `acc` tensor).
 * Outputs (each with a leading n_samples dimension): attn_text (res*res, last-first) fp32;  smoothed (n_tokens, res*res) fp32;  stats (n_tokens, GA_STATS) fp32;
 *          argmax (n_tokens) int32 first-occurrence row-major;  total (1) fp32 = sum group_weight * scaled.
 * masks (n_boxes, res, res) u8 and weights (n_boxes, res, res) fp32 (strict mode only, may be NULL otherwise).
 * ticket: n_samples x 4-byte device workspace that hands the per-pixel stage over to the per-token stage inside the
 *         single launch; zero it once, every launch leaves it zero; launches sharing a ticket must be stream-ordered. */
int ga_guidance_tail_fwd(const float* const* acc_host, const int32_t* slices_host, int n_acc,
                         const ga_tail_params_t* params_host, const ga_token_t* tokens_host, const uint8_t* masks,
                         const float* weights, float* attn_text, float* smoothed, float* stats, int32_t* argmax,
                         float* total, uint32_t* ticket, ga_stream_t stream);

/* Backward of the tail, one launch:  d_abar (res*res, n_ctx) fp32 = d loss / d (each accumulator slice), i.e. already
 * multiplied by inv_count; rows are `d_abar_row_stride` (>= n_ctx) floats apart; one (res*res, stride) slice per
 * sample.  Upstream gradients (all DEVICE, any may be NULL = zero): g_total (1), g_stats
 * (n_tokens, GA_STATS), g_attn_text (res*res, last-first) -- each with a leading n_samples dimension. */
int ga_guidance_tail_bwd(const ga_tail_params_t* params_host, const ga_token_t* tokens_host, const uint8_t* masks,
                         const float* weights, const float* attn_text, const float* smoothed, const float* stats,
                         const int32_t* argmax, const float* g_total, const float* g_stats, const float* g_attn_text,
                         float* d_abar, int d_abar_row_stride, ga_stream_t stream);

/* ---- device-side driver of one denoising step -----------------------------------------------------------
 * Replaces the HOST control flow of one denoising step of reference pipeline_guided_attention.py:925-1053 -- recursion
 * rounds, `_perform_iterative_refinement_step` (:475-581), the threshold tests `meets_threshold` (:1074-1088) on the
 * sub-prompt-grouped unscaled losses (:358-387) and the re-noising (:1046-1050) -- by ONE CUDA graph with conditional
 * (WHILE / IF) nodes built from five device programs the caller captured (raw `cudaGraph_t` handles; cloned into the
 * driver's graph, the caller keeps ownership and must keep the memory they reference alive):
 *     eval, update, cfg : the three programs of the guided loop (reads `latents`, update/cfg write `latents_out`)
 *     advance           : latents <- latents_out
 *     renoise           : latents <- sqrt(Bt) latents + sqrt(1 - Bt) noise[n_draws++]
 * The tail kernel's `stats` of the eval / update programs (static buffers, (n_tokens, GA_STATS) fp32) and the optional
 * custom-loss scalars are read by the driver's decision kernels; the same UNet passes run, in the same order, as in the
 * host-driven loop, and no value is read back to the host. */
#define GA_STEP_CTL_BYTES 256
typedef struct ga_step_programs {
  void *eval, *update, *cfg, *advance, *renoise; /* cudaGraph_t */
  void* refine_update; /* cudaGraph_t or NULL: the update program of the refinement loop when it differs from `update`
                          (`use_optimizer`: SGD with momentum, reference :495-497, :545-547); NULL = `update` */
} ga_step_programs_t;

typedef struct ga_step_params {     /* everything that changes from one denoising step to the next */
  double thr_call;        /* thresholds[i] of the __call__ argument (used for the first test of a round, :963)         */
  double thr_cfg;         /* state.config.thresholds[i] (refinement loop, :501)                                          */
  double thr_last;        /* last value of state.config.thresholds (the `-1` test of :1001)                              */
  int32_t has_thr_call, has_thr_cfg, has_thr_last; /* key present / dict non-empty                                       */
  int32_t check;          /* (i in thresholds) or update_cond: the round's first evaluation is tested at all            */
  int32_t update_cond;    /* (not only_update_on_threshold_steps and i < max_iter_to_alter) or i in config.thresholds   */
  int32_t recurse_ok;     /* not (i > recurse_until)                                                                      */
  int32_t renoise_ok;     /* prev_timestep > 0 (:1047)                                                                    */
  int32_t recurse_steps;  /* >= 1                                                                                          */
  int32_t max_refine;     /* max_refinement_steps (10)                                                                    */
  int64_t timestep;       /* -> *t_dev                                                                                     */
  float step_size;        /* scale_factor * sqrt(scale_range[i]) -> *step_dev                                             */
  float ddim[4];          /* DDIM coefficients of the step -> ddim_dev[0..3]                                              */
  float renoise[2];       /* sqrt(Bt), sqrt(1 - Bt) -> renoise_dev[0..1]                                                  */
} ga_step_params_t;

/* ctl_dev: GA_STEP_CTL_BYTES of zero-initialised DEVICE memory (state + lifetime counters, see ga_step_counters).
 * tokens_host: the tail's token table (group ids and kinds are used); n_groups, avg_within as in the tail spec.
 * t_dev (int64), step_dev (float), ddim_dev (float[4]), renoise_dev (float[2]): the DEVICE scalars the captured programs
 * read; ga_step_driver_run fills them from `params_host` on the stream before launching the graph.
 * stats_refine / custom_refine: outputs of `programs->refine_update` (ignored when that is NULL).
 * refine_first_dev: optional DEVICE float the driver sets to 1 when a refinement is about to start and to 0 after each
 * of its iterations (the momentum program uses it to start from a fresh optimizer state); may be NULL. */
int ga_step_driver_create(void** driver_out, const ga_step_programs_t* programs, void* ctl_dev,
                          const float* stats_eval, const float* stats_update, const float* custom_eval,
                          const float* custom_update, const float* stats_refine, const float* custom_refine,
                          float* refine_first_dev, const ga_token_t* tokens_host, int n_tokens, int n_groups,
                          int avg_within, int64_t* t_dev, float* step_dev, float* ddim_dev, float* renoise_dev);
int ga_step_driver_run(void* driver, const ga_step_params_t* params_host, ga_stream_t stream);
/* Only the parameter kernel of ga_step_driver_run (fills *t_dev, *step_dev, ddim_dev, renoise_dev and the control block):
 * for denoising steps that test no threshold the caller can replay `eval`, `cfg` and `advance` itself -- there is nothing
 * to decide, and a conditional node costs ~50 us per executed body (measured, tools/diag_cond_graph.py). */
int ga_step_driver_set_params(void* driver, const ga_step_params_t* params_host, ga_stream_t stream);
int ga_step_driver_destroy(void* driver);
/* int32 offsets into ctl_dev of the lifetime counters (UNet passes by program, refinement iterations, rounds, re-noise) */
enum ga_step_counter { GA_STEP_N_EVAL = 0, GA_STEP_N_UPDATE = 1, GA_STEP_N_CFG = 2, GA_STEP_N_REFINE = 3,
                       GA_STEP_N_ROUNDS = 4, GA_STEP_N_RENOISE = 5 };
#define GA_STEP_COUNTER_BASE 19 /* counters start at ((int32_t*)ctl_dev)[GA_STEP_COUNTER_BASE] */

/* ---- stand-alone stages (same device code as the tail), for the module-level API -----------------------------------
 * GaussianSmoothing.forward on reflect-padded maps (utils/gaussian_smoothing.py:63-71 + pipeline :253):
 *   maps (n, res, res) fp32 -> out (n, res, res) */
int ga_smooth_fwd(const float* maps, float* out, int n_maps, int res, const float* w1d_host, ga_stream_t stream);
int ga_smooth_bwd(const float* g_out, float* g_maps, int n_maps, int res, const float* w1d_host, ga_stream_t stream);

/* helpers.calculate_bounding_box_losses (utils/helpers.py:215-277) on one normalised map p (res, res):
 *   out2 = (loss_inside, loss_outside);  bwd: g_p = g_out2[0] d inside/dp + g_out2[1] d outside/dp */
int ga_box_loss_fwd(const float* p, const uint8_t* mask, const float* weights, int res, int strict, float* out2,
                    ga_stream_t stream);
int ga_box_loss_bwd(const float* p, const uint8_t* mask, const float* weights, int res, int strict,
                    const float* g_out2, float* g_p, ga_stream_t stream);

/* ---- the caller either side of the attention layers: GroupNorm (+ SiLU) of the UNet blocks -------------------------
 * The guidance path runs the UNet forward AND backward on every guided step (pipeline_guided_attention.py:455-470
 * `_update_latent`; the UNet forward the pipeline drives: :583-743); the norms of its ResNet / transformer blocks
 * (diffusers `ResnetBlock2D`: norm1 -> SiLU -> conv1, norm2 -> SiLU -> conv2; `Transformer2DModel.norm`) are
 * `torch.nn.GroupNorm(32, C)` on (n, C, h, w) activations.  These entry points replace that op pair for CHANNELS-LAST
 * 16-bit activations: x, y, d_y, d_x are (n, hw, channels) dense (NHWC memory), gamma / beta (channels) in the same
 * dtype, `stats` (n, groups, 2) fp32 = (mean, rstd) written by the forward and read by the backward.
 *   y = act(GroupNorm(x)),  act = SiLU when `silu` != 0, identity otherwise;   d_x = d loss / d x  (no d gamma / d beta:
 *   the UNet is frozen on the guidance path).
 * `ws`: scratch of ga_group_norm_ws_bytes() bytes (contents irrelevant on entry, per call in flight).  Two launches per
 * call, no atomics: results are bit-stable run to run.
 * ga_group_norm_ws_bytes returns -1 when the shape is not supported (channels % 8, channels % groups, a group
 * narrower than the kernels' 8-channel vectors allow). */
int64_t ga_group_norm_ws_bytes(int n, int hw, int channels, int groups);
/* `shift` (n rows of `channels` values, `shift_stride` elements apart), same dtype, or NULL: the kernels normalise
 * x + shift[n, c] -- the bias of the convolution that produced x and the block's time-embedding projection
 * (ResnetBlock2D: `conv1(..) + time_emb_proj(..)[:, :, None, None]`) folded into the norm instead of two broadcast-add
 * launches; the stride lets the rows be a column slice of ONE projection computed for all blocks of the UNet. */
int ga_group_norm_fwd(const void* x, const void* shift, int64_t shift_stride, const void* gamma, const void* beta,
                      void* y, float* stats, float* ws, int n, int hw, int channels, int groups, float eps, int silu,
                      int dtype, ga_stream_t stream);
int ga_group_norm_bwd(const void* x, const void* shift, int64_t shift_stride, const void* d_y, const void* gamma,
                      const void* beta, const float* stats, void* d_x, float* ws, int n, int hw, int channels, int groups,
                      int silu, int dtype, ga_stream_t stream);
/* out = a + bias[c] (+ b when b != NULL) on channels-last (n_pixels, channels) 16-bit tensors: a convolution's bias and
 * the block's residual connection (`x + conv2(..)`) in one vectorised pass.  `out` may alias `a`. */
int ga_add_bias_residual(const void* a, const void* b, const void* bias, void* out, int64_t n_pixels, int channels,
                         int dtype, ga_stream_t stream);

/* The gate of the transformer blocks' feed-forward (diffusers `GEGLU.forward`: `h, gate = proj(x).chunk(2, dim=-1);
 * return h * gelu(gate)`, exact erf GELU) on the projection output `proj` (rows, 2 * inner), dense 16-bit:
 *   out (rows, inner) = proj[:, :inner] * gelu(proj[:, inner:]);   d_proj (rows, 2 * inner) from d_out (rows, inner).
 * One vectorised launch per direction instead of PyTorch's strided gelu + product (+ gelu_backward, two products and a
 * concatenating copy in the backward). */
int ga_geglu_fwd(const void* proj, void* out, int64_t rows, int inner, int dtype, ga_stream_t stream);
int ga_geglu_bwd(const void* proj, const void* d_out, void* d_proj, int64_t rows, int inner, int dtype,
                 ga_stream_t stream);

/* LayerNorm forward of the transformer blocks (diffusers `BasicTransformerBlock.norm1/2/3`: `torch.nn.LayerNorm(C)`) on
 * dense 16-bit tokens x (rows, channels): y = (x - mean) * rstd * gamma + beta, one warp per row; `mean`, `rstd` (rows)
 * fp32 are written for the backward (which stays PyTorch's `native_layer_norm_backward`). */
int ga_layer_norm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                      int64_t rows, int channels, float eps, int dtype, ga_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GUIDED_ATTN_H_ */
