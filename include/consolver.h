/*
 * consolver.h — C ABI of libconsolver.so: the B200 (sm_100a) implementation of the ConsistencySolver
 * sampling step (G-U-N/consolver `PPOScheduler.step` / `FMPPOScheduler.step` + `FactorNetPPO.sample_action`
 * + the caller's CFG combine).
 *
 * Boundary rules (every entry point):
 *   - plain pointers + sizes + a CUDA stream handle; no torch / C++ types;
 *   - all data pointers are DEVICE pointers unless the comment says "host";
 *   - the library never allocates, never synchronises and never throws; the caller owns all memory;
 *   - returns 0 on success, a negative CONSOLVER_ERR_* for a rejected argument, or a positive
 *     cudaError_t if the launch failed.
 *
 * The reference has no native interface for this path (it is pure PyTorch), so each entry point cites the
 * reference Python lines it replaces (paths relative to the reference root).
 */
#ifndef CONSOLVER_H_
#define CONSOLVER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bumped on EVERY change of a signature, struct layout or flag meaning in this header.  Loaders must also compare
 * consolver_abi_hash() with the hash of the header they were written against (see consolver_abi_hash below). */
#define CONSOLVER_ABI_VERSION 8

/* element type of latents / model outputs */
#define CONSOLVER_F32  0
#define CONSOLVER_F16  1
#define CONSOLVER_BF16 2

#define CONSOLVER_MAX_ORDER   8     /* order_dim  <= 8  (reference default 4, FLUX config 2)          */
#define CONSOLVER_MAX_HIDDEN  1024  /* hidden_dim <= 1024 (reference default 256)                     */
#define CONSOLVER_MAX_LOGITS  4096  /* action_dims * num_actions <= 4096 (reference default 5*161)    */
#define CONSOLVER_MAX_IN      16    /* MLP input width (2, or 2+order_dim-1 with use_conv)            */

/* per-sample coefficient record written by consolver_policy_f32 and read by the step kernels:
 *   coef[b*stride + i]            i < order_dim : multiplier of history entry i (0 = newest)
 *   coef[b*stride + order_dim]    (1 + s0) multiplier of the combined model output (1 if scaler_dim == 0)
 *   coef[b*stride + order_dim+1]  (1 + s1) multiplier of the sample               (1 if scaler_dim <  2)
 * stride = order_dim + 2. */
#define CONSOLVER_COEF_STRIDE(order_dim) ((order_dim) + 2)

/* flags of the step entry points */
#define CONSOLVER_FLAG_VPRED        1   /* prediction_type == "v_prediction" (scheduler_ppo.py:316-317)            */
#define CONSOLVER_FLAG_EFF_SCALE    2   /* scaler_dim >= 1: eff *= coef[order_dim]     (scheduler_ppo.py:274-277)  */
#define CONSOLVER_FLAG_X_SCALE      4   /* scaler_dim == 2: sample *= coef[order_dim+1] (scheduler_ppo.py:278)     */
#define CONSOLVER_FLAG_PDL          8   /* launch with programmatic dependent launch: the bulk loads are issued
                                           before waiting on the preceding (policy) kernel's coefficients.  Only
                                           valid when the kernel launched right before on this stream is the one that
                                           writes `coef` (or writes nothing this step reads): e0 / cond / hist / x are
                                           loaded BEFORE the wait, `coef` after it                                  */

#define CONSOLVER_FLAG_CHAIN       16   /* back-to-back solver steps on one stream: this step is a programmatic
                                           dependent launch of the PREVIOUS STEP; loads of e0/cond and of history
                                           entries older than the newest are issued before waiting on it, x and the
                                           newest history entry after.  Only valid when e0/cond/older entries are not
                                           written by the immediately preceding kernel (solver-only replays)        */
#define CONSOLVER_FLAG_LOWP_COMBINE 32  /* consolver_step_fm, 16-bit model outputs only: form the history combination
                                           and its product with dt in the MODEL dtype, the way torch evaluates the
                                           baseline solvers' `0.5 * dt * (v_old + v_new)` on bf16/fp16 tensors
                                           (edit_ppo/scheduler_fm.py:430): sum rounded to the model dtype, dt rounded
                                           to it, product rounded to it, then the fp32 add with the sample.  Without
                                           the flag the combination is fp32 (per-sample fp32 coefficients promote,
                                           edit_ppo/scheduler_fmppo.py:413-429).  Ignored for fp32 model outputs     */

#define CONSOLVER_FLAG_X_F32        64   /* consolver_step_sd with a 16-bit dtype: x, x_out and x_out2 are fp32 (the model
                                           outputs, history and ring slot stay `dtype`).  This is the layout of the
                                           reference's mixed-precision training loop (train_ppo.py:353, accelerate
                                           fp16/bf16 autocast): fp32 latents, 16-bit U-Net output, and torch promotion
                                           keeps every later product and the returned latent in fp32
                                           (scheduler_ppo.py:272,:323-330).  Ignored when dtype is CONSOLVER_F32    */

#define CONSOLVER_FLAG_X_WAS_LOWP  128   /* with CONSOLVER_FLAG_X_F32: the fp32 latent handed in is an exact upcast of a
                                           `dtype` latent (the step at which torch promotion turns a 16-bit pipeline's
                                           latent fp32).  The one place the reference multiplies the still-16-bit
                                           sample by a scalar — v-prediction's sqrt(1-abar_t)*sample,
                                           scheduler_ppo.py:317 — is then a 16-bit product; ignored with X_SCALE     */

/* WHERE THE REFERENCE RUNS.  The reference keeps its schedule scalars as 0-d fp32 tensors on the HOST
 * (scheduler_ppo.py:110-114,:309-312) and ATen combines such a scalar with a tensor differently per device.  By default
 * the kernels follow ATen's CUDA rules, i.e. the reference executed on a GPU the way its drivers run it (pinned by the
 * tests/golden/cuda_* fixtures, produced by the unmodified reference on a B200):
 *     scalar * t   the scalar enters as an fp32 value: one rounding, to t's dtype
 *     t / scalar   t * (1/scalar), reciprocal taken once in fp32 — this moves fp32 results by an ulp here and there
 * CONSOLVER_FLAG_HOST_SCALARS selects torch's CPU rules instead (the reference executed on CPU tensors; fixtures
 * tests/golden/{sd,sd16,fm}_*): scalar rounded to a 16-bit t's dtype before the product, true division. */
#define CONSOLVER_FLAG_HOST_SCALARS 256

#define CONSOLVER_FLAG_LOWP_COEF    512  /* consolver_step_sd, 16-bit dtype: the per-sample coefficients are 16-bit TENSORS
                                           in the reference — the policy, bin buffer included, was cast to the pipeline
                                           dtype (gen_ppo.py:193-195) — so `c_i * e_i` and the partial sums are 16-bit
                                           ops, EXCEPT the closing coefficient 1 - sum(...), which is fp32 because
                                           torch.sum returns fp32 under autocast (gen_ppo.py:309) and promotes the sum
                                           from there on; `(1+s0)` / `(1+s1)` are 16-bit too, so a depth-1 estimate and a
                                           16-bit sample stay 16-bit through their scalings.  coef[] holds the values
                                           consolver_policy_* wrote with CONSOLVER_POLICY_COEF_F16/_BF16 */

/* policy_flags of the policy entry points */
#define CONSOLVER_POLICY_HOST_DIV    1   /* x / x_div and logits / temp as true divisions (torch on CPU tensors); default:
                                           multiplications by the fp32 reciprocals (ATen's CUDA division by a scalar).
                                           Also selects the ORDER in which the closing coefficient's
                                           torch.sum(torch.stack(...), dim=0) adds its terms (scheduler_ppo.py:172):
                                           left to right on CPU tensors (exact for up to 4 terms, order_dim <= 5; torch's
                                           CPU reduction of more terms is not restated); on CUDA tensors ATen's reduce
                                           kernel order —
                                           per-sample 4-accumulator form for B >= 2, a two-level tree when B == 1
                                           (three terms: (c0 + a2) + a1), see csrc/mlp_device.cuh::sum_terms            */
#define CONSOLVER_POLICY_ACT_F16     2   /* the MLP runs under torch.autocast(fp16) (gen_ppo.py:309, train_ppo.py:353) or
                                           with parameters cast to fp16: the input row, every Linear output and
                                           logits/temp are rounded to fp16 (fp32 accumulation, bias added before the
                                           rounding); softmax in fp32.  The caller passes weights already rounded      */
#define CONSOLVER_POLICY_ACT_BF16    4   /* same with bfloat16                                                         */
#define CONSOLVER_POLICY_COEF_F16    8   /* the bin values are fp16 tensors (policy cast to the pipeline dtype): a0+1,
                                           1+s0, 1+s1 are rounded to fp16, the closing coefficient is the fp32
                                           1 - sum (see CONSOLVER_FLAG_LOWP_COEF)                                      */
#define CONSOLVER_POLICY_COEF_BF16  16   /* same with bfloat16                                                         */

#define CONSOLVER_ERR_NULL        (-1)
#define CONSOLVER_ERR_SIZE        (-2)
#define CONSOLVER_ERR_UNSUPPORTED (-3)
#define CONSOLVER_ERR_DTYPE       (-4)

typedef void* consolver_stream_t; /* a cudaStream_t */

#if defined(__GNUC__)
#define CONSOLVER_API __attribute__((visibility("default")))
#else
#define CONSOLVER_API
#endif

/* In-kernel Exp(1) draw that reproduces torch's CUDA `Tensor.exponential_(1)` stream bit-for-bit, so the sample
 * kernel needs no q buffer and no separate RNG launch: the value written at linear index i of a contiguous fp32
 * tensor of `numel` elements for default-generator state (seed, offset).  nthreads and the generator's offset
 * increment come from consolver_torch_philox_plan(numel).  `state` (nullable, device uint64[2] = {seed, offset})
 * overrides `seed` and is ADDED to `offset`: CUDA-graph replays refresh it instead of re-recording parameters. */
typedef struct consolver_rng {
  uint64_t seed;
  uint64_t offset;
  const uint64_t* state;
  uint32_t nthreads;
} consolver_rng_t;

CONSOLVER_API int consolver_abi_version(void);
/* FNV-1a 64-bit hash of this header's text as it was when the library was compiled (csrc/abi_hash.h is generated from
 * include/consolver.h by the build).  A loader hashes the header it binds against the same way and refuses a library
 * whose hash differs: a stale .so built from an older header cannot be called with the wrong argument list. */
CONSOLVER_API uint64_t consolver_abi_hash(void);
/* human-readable text for any return value of this library (static storage). */
CONSOLVER_API const char* consolver_error_string(int err);

/*
 * Policy step: FactorNetPPO.forward_ + sample_action (factor_net_ppo.py:137-168; FM variant
 * edit_ppo/factor_net_ppo.py:149-180) and the scheduler's mask / coefficient assembly
 * (scheduler_ppo.py:248-259 + set_default_coefficients :165-175; edit_ppo/scheduler_fmppo.py:403-410).
 *
 * The MLP input is the same for every sample of the batch (scheduler_ppo.py:207-210), so the MLP + softmax is
 * evaluated ONCE per launch (per CTA) on x = (x0, x1) / x_div; then every sample b draws its A indices as
 * argmax_k probs[a,k] / q[b,a,k] — exactly what torch.multinomial(probs.view(-1,K), 1) computes with
 * q = empty_like(probs).exponential_(1) (factor_net_ppo.py:161) — lowest index on ties.
 *
 *   w1 [H,in_dim] b1 [H] w2 [H,H] b2 [H] w3 [A*K,H] b3 [A*K]   row-major fp32 (nn.Linear layout: mlp.{0,2,4})
 *   action_values [A,K]      the `action_values` buffer of the state_dict (bin values)
 *   x0,x1                    (t, prev_t) for SD / (sigma, sigma_next) for FM, already rounded through the model
 *                            dtype as the reference's torch.tensor(..., dtype=model_output.dtype) does
 *   x_div                    999.0 for SD (normalize_input), 1.0 for FM
 *   temp                     softmax temperature divisor: 1.0 for SD, 0.01 for FM (logits / 0.01)
 *   feat [B,n_feat] or NULL  extra per-sample MLP inputs (use_conv cosine features); in_dim = 2 + n_feat.
 *                            When non-NULL the MLP is evaluated per sample.
 *   q [B*A,K] or NULL        the Exp(1) draw; NULL => use idx_in
 *   idx_in [B,A] or NULL     forced bin indices (replay / PPO update); exactly one of q, idx_in is given
 *   n_hist                   number of model outputs in the history INCLUDING the current one (1..order_dim)
 *   policy_flags             CONSOLVER_POLICY_* (0 = fp32 policy evaluated by the reference on CUDA tensors)
 * outputs (any may be NULL except coef):
 *   probs_table [A,K]        full softmax table of the shared row ([B,A,K], one table per sample, when feat != NULL)
 *   idx [B,A] int64          sampled bin indices
 *   actions [B,A]            bin values (the reference's `actions`)
 *   act_probs [B,A]          probabilities of the sampled bins (the reference's `probs`)
 *   act_logp [B,A]           log(act_probs + 1e-9) (train_ppo.py:410-411)
 *   masks [B,A]              1 everywhere except columns n_hist-1 .. order_dim-2 (scheduler_ppo.py:248-249)
 *   coef [B,order_dim+2]     see CONSOLVER_COEF_STRIDE
 */
CONSOLVER_API int consolver_policy_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* w3, const float* b3, const float* action_values,
                         float x0, float x1, float x_div, float temp,
                         const float* feat, int n_feat,
                         const float* q, const int64_t* idx_in,
                         int B, int H, int A, int K, int order_dim, int scaler_dim, int n_hist, int policy_flags,
                         float* probs_table, int64_t* idx, float* actions, float* act_probs, float* act_logp,
                         float* masks, float* coef, consolver_stream_t stream);

/*
 * Fused SD solver step: CFG combine (denoise_ppo.py:96-100) + linear-multistep combine over the history
 * (scheduler_ppo.py:263-280) + DDIM update (scheduler_ppo.py:306-332), one pass over HBM.
 *
 *   dtype            CONSOLVER_F32 / F16 / BF16: element type of the model outputs, history and ring slot, and of
 *                    x / x_out / x_out2 unless CONSOLVER_FLAG_X_F32 is set.  F32: bit-identical to the reference's
 *                    op-by-op fp32 arithmetic (on CUDA tensors; CONSOLVER_FLAG_HOST_SCALARS: on CPU tensors).
 *                    16-bit: the arithmetic torch performs on those dtypes —
 *                      CFG combine: a rounding to `dtype` after each op, guidance kept in fp32;
 *                      n_hist == 1 without scaler flags (the estimate is the raw 16-bit output): every
 *                        `scalar * tensor` product of scheduler_ppo.py:316-330 is a 16-bit product (one rounding;
 *                        with HOST_SCALARS the scalar is rounded first as well); with a 16-bit latent all
 *                        intermediates are rounded too, with CONSOLVER_FLAG_X_F32 the rest is fp32;
 *                      otherwise (n_hist > 1 or scalers): fp32 arithmetic on the upcast values — exactly torch's
 *                        promotion when the latent is fp32 (X_F32); with a 16-bit latent (a layout the reference
 *                        never produces there) the fp32 result is rounded once
 *   e0               newest model output [B,N]; if `cond` != NULL it is the UNCONDITIONAL half and
 *                    eps = e0 + guidance*(cond - e0) is formed in the kernel
 *   slot_out         NULL, or [B,N]: receives eps (the in-place history-ring slot of this step)
 *   hist (host)      array of n_hist-1 device pointers to the older model outputs, newest first
 *   x, x_out         current / next latent [B,N]
 *   x_out2           NULL, or a second destination for the next latent with `out2_stride` elements between
 *                    samples (0 = N): lets the caller have x' land directly in the other half of the next
 *                    CFG-doubled denoiser input (denoise_ppo.py:66) or inside a wider packed sequence
 *                    (edit_ppo/denoise_diffusion.py:102) instead of running torch.cat
 *   coef             [B,coef_stride] from consolver_policy_f32; n_hist == 1 bypasses the coefficient
 *                    (scheduler_ppo.py:263-265)
 *   sa_t,sb_t,sa_p,sb_p   sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev), sqrt(1-abar_prev) (fp32 values)
 *   n_per_sample     N = C*H*W elements of one sample
 */
CONSOLVER_API int consolver_step_sd(int dtype, const void* e0, const void* cond, float guidance, void* slot_out,
                      const void* const* hist, int n_hist, const void* x, void* x_out,
                      void* x_out2, int64_t out2_stride,
                      const float* coef, int coef_stride, int order_dim,
                      float sa_t, float sb_t, float sa_p, float sb_p, int flags,
                      int B, int64_t n_per_sample, consolver_stream_t stream);

/*
 * Fused flow-matching solver step (edit_ppo/scheduler_fmppo.py:354,:413-436): fp32 upcast of the sample,
 * multistep combine, x' = x*(1+s1) + dt*eff*(1+s0), cast to the model dtype.
 *   dtype     element type of model outputs, history and x_out
 *   x_dtype   element type of the incoming sample (dtype or CONSOLVER_F32)
 *   dt        sigma_next - sigma (fp32 value)
 * With n_hist == 1, scaler_dim == 0 and a 16-bit dtype the reference's `dt * model_output` product is rounded
 * to the model dtype before the fp32 add (0-d fp32 tensor times bf16 tensor); the kernel reproduces that.
 */
CONSOLVER_API int consolver_step_fm(int dtype, int x_dtype, const void* e0, void* slot_out,
                      const void* const* hist, int n_hist, const void* x, void* x_out,
                      void* x_out2, int64_t out2_stride,
                      const float* coef, int coef_stride, int order_dim, float dt, int flags,
                      int B, int64_t n_per_sample, consolver_stream_t stream);

/*
 * consolver_step_fm with sample-strided model outputs: e0 and every hist entry are views with `e_stride` elements
 * between consecutive samples (each sample contiguous; e_stride >= n_per_sample, a multiple of 8 for the vector path;
 * 0 = n_per_sample).  This is `noise_pred[:, :L]` of a transformer output that covers [latents | image latents]
 * (edit_ppo/denoise_diffusion.py:140): the slice is consumed in place instead of being copied out first.
 */
CONSOLVER_API int consolver_step_fm_strided(int dtype, int x_dtype, const void* e0, int64_t e_stride, void* slot_out,
                      const void* const* hist, int n_hist, const void* x, void* x_out,
                      void* x_out2, int64_t out2_stride,
                      const float* coef, int coef_stride, int order_dim, float dt, int flags,
                      int B, int64_t n_per_sample, consolver_stream_t stream);

/*
 * Fused multistep DPM-Solver / DPM-Solver++ step with AMED direction scaling — replaces, per step,
 * the caller's CFG combine (gen_pretrain/pipeline.py:1069-1071), diffusers' convert_model_output (diffusers 0.26.3
 * DPMSolverMultistepScheduler; not part of the reference tree) and the plugin's first/second/third-order updates
 * (diffusers_amed_plugin_dpmpp.py:70-138, :140-262, :264-348) with one pass over HBM.  Scalars are computed by the host the way
 * the plugin computes them (0-d fp32 tensors) and handed in:
 *   m0 = e                            convert == CONSOLVER_DPM_CONVERT_NONE
 *      = (x - ck0*e) / ck1                        CONSOLVER_DPM_CONVERT_DIV / _DIV_RECIP   (dpmsolver++ / epsilon: ck0 = sigma_s, ck1 = alpha_s)
 *      = ck1*x + ck0*e                            CONSOLVER_DPM_CONVERT_LIN   (v-prediction forms)
 *   x'  = cx*x - a0*m0                                          m1 == NULL            (first order,  :121/:123)
 *       = (cx*x - a0*m0) - a1*D                                 m1 != NULL, m2 == NULL (second order, :201-208)
 *             D   = rinv*(m0 - m1)
 *       = ((cx*x - a0*m0) - a1*D1) - a2*D2                      m1, m2 != NULL         (third order,  :326-346)
 *             D10 = rinv*(m0 - m1), D11 = rinv1*(m1 - m2), D1 = D10 + w*(D10 - D11), D2 = rs*(D10 - D11)
 * (a term the plugin ADDS is passed with its sign flipped: a - (-b) == a + b exactly.)
 * e = e0, or e0 + guidance*(cond - e0) when cond != NULL.  m0 is written to slot_out when non-NULL (the caller's
 * ring of converted outputs).  fp32 arithmetic in exactly this order; 16-bit dtypes round e, m0 and x' once each.
 */
#define CONSOLVER_DPM_CONVERT_NONE 0
#define CONSOLVER_DPM_CONVERT_DIV  1
#define CONSOLVER_DPM_CONVERT_LIN  2
#define CONSOLVER_DPM_CONVERT_DIV_RECIP 3   /* DIV as ATen evaluates it on CUDA tensors: (x - ck0*e) * (1/ck1), the
                                              reciprocal of the host-resident scalar taken once in fp32 (DIV itself is
                                              the true division of a run on CPU tensors)                              */
typedef struct consolver_dpm_update {   /* host struct, fp32 values of the plugin's 0-d tensors */
  float cx, a0;              /* all orders                                                                 */
  float a1, rinv;            /* second and third order: rinv = 1/r0                                        */
  float a2, rinv1, w, rs;    /* third order: rinv1 = 1/r1, w = r0/(r0+r1), rs = 1/(r0+r1)                  */
} consolver_dpm_update_t;
CONSOLVER_API int consolver_step_dpm(int dtype, int x_dtype, const void* e0, const void* cond, float guidance,
                      void* slot_out, const void* m1, const void* m2, const void* x, void* x_out, void* x_out2,
                      int64_t out2_stride, int convert, float ck0, float ck1, const consolver_dpm_update_t* upd,
                      int B, int64_t n_per_sample, consolver_stream_t stream);

/*
 * Probability tables only: MLP + softmax (factor_net_ppo.py:137-157) for `rows` input rows in one launch (one CTA
 * per row).  The schedulers call it once per trajectory with the (t, prev_t) / (sigma, sigma_next) rows of the
 * whole timestep grid — the policy input does not depend on the sample — and then run only
 * consolver_policy_sample_f32 per step.
 *   x_rows [rows,2] fp32 (device)      probs_tables [rows,A,K] (device, out)
 */
CONSOLVER_API int consolver_policy_table_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                             const float* w3, const float* b3, const float* x_rows, int rows,
                                             float x_div, float temp, int H, int A, int K, int policy_flags,
                                             float* probs_tables, consolver_stream_t stream);

/*
 * Sampling only, from a given probability table probs_in [A,K]: the draw, gathers, masks and coefficient assembly
 * of consolver_policy_f32 (factor_net_ppo.py:159-168; scheduler_ppo.py:248-259,:165-175).  Same outputs.
 * Exactly one draw source: q [B*A,K] (given Exp(1) values), idx_in [B,A] (forced bins) or rng (host struct: the
 * kernel generates torch's exponential_ values itself; q_out, nullable [B*A,K], receives them for checking).
 */
CONSOLVER_API int consolver_torch_philox_plan(int64_t numel, uint32_t* nthreads, uint64_t* offset_increment);
/* state[1] += amount on the stream (one thread): put it at the end of a captured graph so the device-resident
 * generator state advances by the graph's total consumption and replays need no host refresh. */
CONSOLVER_API int consolver_rng_state_advance(uint64_t* state, uint64_t amount, consolver_stream_t stream);
CONSOLVER_API int consolver_policy_sample_f32(const float* probs_in, const float* action_values, const float* q,
                                              const int64_t* idx_in, const consolver_rng_t* rng, float* q_out,
                                              int B, int A, int K, int order_dim,
                                              int scaler_dim, int n_hist, int policy_flags, int64_t* idx, float* actions,
                                              float* act_probs, float* act_logp, float* masks, float* coef,
                                              consolver_stream_t stream);

/*
 * CONTINUOUS (Gaussian) policy — ppo_type != "discrete".  EXTENSION, PARITY UNPINNED: the reference instantiates
 * `FactorNetPPOContinous` (scheduler_ppo.py:23,:139) but ships no source for it, so these semantics are this library's
 * own (csrc/policy_gauss.cu states them); masks and coefficient assembly are the discrete policy's
 * (scheduler_ppo.py:248-259,:165-175).  Same trunk, last layer w3 [2A,H] -> raw mean / raw log-std per action dim;
 *   mean = mid + half*tanh(raw_mean), std = half*exp(clamp(raw_logstd,-7,1)), [lo,hi] = action_range[a] (host-chosen,
 *   by default the discrete bins' ranges); action = mean + std*z; logp = -z^2/2 - log std - log(2 pi)/2.
 * Exactly one draw source: z [B*A] (given N(0,1) values), actions_in [B*A] (forced actions: z is recovered), or rng (the
 * kernel regenerates what torch.randn([B,A]) on the CUDA default generator would draw; plan =
 * consolver_torch_philox_plan(B*A)).  Outputs: mean_std [2,A] (nullable), z_out [B*A] (nullable), actions / act_probs
 * (= exp(logp), a density) / act_logp / masks [B,A], coef [B,order_dim+2].
 */
CONSOLVER_API int consolver_policy_gauss_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                             const float* w3, const float* b3, const float* action_range,
                                             float x0, float x1, float x_div,
                                             const float* z, const float* actions_in, const consolver_rng_t* rng,
                                             int B, int H, int A, int order_dim, int scaler_dim, int n_hist,
                                             int policy_flags, float* mean_std, float* z_out, float* actions,
                                             float* act_probs, float* act_logp, float* masks, float* coef,
                                             consolver_stream_t stream);

/*
 * One call per scheduler step for the SD path: the policy (consolver_policy_sample_f32 when probs_in != NULL,
 * else consolver_policy_f32 with the weights) followed by consolver_step_sd on the same stream.  Saves one host
 * crossing; with CONSOLVER_FLAG_PDL the step kernel is a programmatic dependent launch whose bulk loads overlap
 * the policy kernel.  Arguments as in the functions above.
 */
CONSOLVER_API int consolver_sd_policy_and_step(const float* w1, const float* b1, const float* w2, const float* b2,
                                 const float* w3, const float* b3, const float* action_values,
                                 const float* probs_in,
                                 float x0, float x1, float x_div, float temp,
                                 const float* q, const int64_t* idx_in, const consolver_rng_t* rng,
                                 int H, int A, int K, int scaler_dim, int policy_flags,
                                 float* probs_table, int64_t* idx, float* actions, float* act_probs,
                                 float* act_logp, float* masks, float* coef,
                                 int dtype, const void* e0, const void* cond, float guidance, void* slot_out,
                                 const void* const* hist, int n_hist, const void* x, void* x_out,
                                 void* x_out2, int64_t out2_stride,
                                 int order_dim, float sa_t, float sb_t, float sa_p, float sb_p, int flags,
                                 int B, int64_t n_per_sample, consolver_stream_t stream);

/*
 * use_conv=True first pass: per-sample cosine similarity between the newest model output and each older history
 * slot (factor_net_ppo.py:108-130; zero for slots not yet filled), written as feat [B, order_dim-1] for the
 * `feat` argument of consolver_policy_f32.  e0/cond/guidance/hist/n_hist as in consolver_step_sd (with `cond` the
 * newest output is formed as e0 + guidance*(cond - e0) on the fly).  `workspace` is caller-owned scratch of
 * consolver_cosine_features_workspace(B, order_dim) bytes (8-byte aligned); it is zeroed on the stream here.
 */
CONSOLVER_API size_t consolver_cosine_features_workspace(int B, int order_dim);
CONSOLVER_API int consolver_cosine_features(int dtype, const void* e0, const void* cond, float guidance,
                                            const void* const* hist, int n_hist, int order_dim, int B,
                                            int64_t n_per_sample, void* workspace, float* feat,
                                            consolver_stream_t stream);

/*
 * PPO update of the policy: forward + clipped-ratio loss + entropy bonus + backward of
 * `factor_net(conds, actions)` (factor_net_ppo.py:170-184) and the loss of train_ppo.py:406-427, evaluated on the
 * `rows` DISTINCT condition rows of a rollout instead of B*(n-1) replicas.
 *   x_rows [rows,2]              (t, prev_t) / (sigma, sigma_next) of steps 1..n-1
 *   idx, old_probs, advantages   [rows,B,A]: sampled bins, their probabilities at rollout time, advantages * masks
 *   workspace                    consolver_ppo_workspace(rows,H,A,K) bytes of scratch (per-row partial gradients)
 *   grad_flat [P]                d loss / d parameters in nn.Module.parameters() order: mlp.0.weight [H,2], mlp.0.bias,
 *                                mlp.2.weight [H,H], mlp.2.bias, mlp.4.weight [A*K,H], mlp.4.bias   (overwritten)
 *   stats [4]                    {loss, policy_loss, mean normalised entropy, mean ratio}
 * The sum over rows runs in a fixed order: results are deterministic run to run.
 */
CONSOLVER_API size_t consolver_ppo_workspace(int rows, int H, int A, int K);
CONSOLVER_API int consolver_ppo_loss_grad_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                              const float* w3, const float* b3, const float* x_rows, int rows,
                                              float x_div, float temp, int H, int A, int K,
                                              const int64_t* idx, const float* old_probs, const float* advantages,
                                              int B, float clip_range, float entropy_coef,
                                              void* workspace, float* grad_flat, float* stats,
                                              consolver_stream_t stream);

/*
 * The same update with the data-parallel gradient exchange FUSED into the reduction kernel: a one-shot all-reduce (AVG)
 * over NVLink peer memory instead of a separate NCCL call — replaces DDP's gradient all-reduce of train_ppo.py:257,:430
 * and edit_ppo/train_ppo.py:382 (300 KB per PPO epoch, latency-bound).  `peers` == NULL or world == 1: identical to
 * consolver_ppo_loss_grad_f32.  Otherwise every rank must call this with the same `epoch`; on return (in stream order)
 * grad_flat holds (sum over ranks, added in rank order — bit-identical on every rank) / world.
 *   buffer_ptrs_dev   device array [world] of peer-mapped base pointers of the ranks' exchange buffers, each
 *                     2 * stride_floats floats (two epoch parities); entry [rank] is this rank's own buffer
 *                     (stride_floats >= P, a multiple of 64)
 *   signal_ptrs_dev   device array [world] of peer-mapped flag pads, pad_words uint32 each, zero before the first call;
 *                     pad_words >= consolver_ppo_exchange_pad_words(H, A, K, world) (one flag per rank and CTA)
 *   epoch             1, 2, 3, ... incremented by one per call, identical on all ranks
 * The buffers are obtained by the caller from any symmetric-memory allocator (the Python host uses
 * torch.distributed._symmetric_memory); the library only needs the pointers.
 */
typedef struct consolver_peers {
  void* const* buffer_ptrs_dev;
  void* const* signal_ptrs_dev;
  int rank, world;
  uint32_t epoch;
  int64_t stride_floats;
  int64_t pad_words;
} consolver_peers_t;
CONSOLVER_API int64_t consolver_ppo_exchange_pad_words(int H, int A, int K, int world);
CONSOLVER_API int consolver_ppo_loss_grad_allreduce_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                              const float* w3, const float* b3, const float* x_rows, int rows,
                                              float x_div, float temp, int H, int A, int K,
                                              const int64_t* idx, const float* old_probs, const float* advantages,
                                              int B, float clip_range, float entropy_coef,
                                              void* workspace, float* grad_flat, float* stats,
                                              const consolver_peers_t* peers, consolver_stream_t stream);

/* Tuning knobs for benchmarking (process-global; not part of the numerical contract).
 *   threads: CTA size of the step kernels (32..512, multiple of 32; 0 = default)
 *   unroll : 16-byte vectors per thread per stream; only the one-vector form is compiled in (round 2: the
 *            two-vector form was never ahead), so 0, 1 and 2 all run it — kept so that callers need not change */
CONSOLVER_API int consolver_set_step_launch(int threads, int unroll);

#ifdef __cplusplus
}
#endif
#endif /* CONSOLVER_H_ */
