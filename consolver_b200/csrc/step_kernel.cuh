// step_kernel.cuh — the fused solver-step kernel template (SD DDIM-form and FM Euler-form) and its launcher.
//
// One launch reads, per sample, the newest model output (or the CFG pair), the n_hist-1 older history
// slots and the current latent ONCE and writes the next latent (and, optionally, the CFG-combined model
// output into its history-ring slot) ONCE:   bytes/sample = (n_hist + 4) * N * sizeof(T)   with a CFG pair.
// Reference arithmetic being fused: denoise_ppo.py:96-100, scheduler_ppo.py:263-280 and :306-332,
// edit_ppo/scheduler_fmppo.py:354,:413-436.
#pragma once
#include <type_traits>

#include "step_common.cuh"

namespace consolver {

// kModeFMStrided: the FM step with sample-strided model outputs (a separate instantiation, so that the contiguous
// form pays nothing for it: the extra address arithmetic cost ~0.3 us per launch at small batches when it was
// folded into kModeFM)
enum : int { kModeSD = 0, kModeFM = 1, kModeFMStrided = 2 };

// NH  > 0 : history depth known at compile time (1..4), loads fully unrolled
// NH == 0 : runtime depth (5..8), guarded loads
// T  : element type of model outputs, history and ring slot;  TX : element type of the incoming latent.
// The next latent is written as T by the FM step (the reference casts back to the model dtype,
// edit_ppo/scheduler_fmppo.py:436) and as TX by the SD step (torch promotion: fp32 latents with 16-bit model outputs
// under accelerator.autocast stay fp32, train_ppo.py:353).
// HOST : CONSOLVER_FLAG_HOST_SCALARS resolved at compile time (SD only), so that the default instantiations — the ones
//        that run in production — carry no IEEE-division slow path; HOST = true exists only with U = 1.
template <typename T, typename TX, int NH, int MODE, int E, int U, bool HOST = false>
__global__ void __launch_bounds__(512) step_kernel(const StepParams p) {
  using TO = typename std::conditional<MODE == kModeSD, TX, T>::type;
  constexpr int kOlder = NH ? NH - 1 : kMaxOlder;
  const int b = blockIdx.x / p.chunks_per_sample;
  const int chunk = blockIdx.x - b * p.chunks_per_sample;
  const long long base = (long long)b * p.n_per_sample;
  const long long v0 = (long long)chunk * ((long long)blockDim.x * U) + threadIdx.x;
  const int nh = NH ? NH : p.n_hist;
  const bool pair = p.cond != nullptr;
  // kModeFMStrided: model outputs (e0 and the history) are sample-strided views, e.g. the first L tokens of a wider
  // packed transformer output (edit_ppo/denoise_diffusion.py:140); every other instantiation keeps one offset
  const long long edelta = (MODE == kModeFMStrided) ? (long long)b * (p.e_stride - p.n_per_sample) : 0;

  Raw<T, E> r_e0[U], r_c[U];
  Raw<T, E> r_h[U][kOlder > 0 ? kOlder : 1];
  Raw<TX, E> r_x[U];
  long long off[U];
  bool live[U];

  // ---- issue every load before the first use ------------------------------------------------------------
  // CHAIN (back-to-back solver steps, each a programmatic dependent launch of the previous one): the latent x and
  // the newest history slot are the previous step's outputs; everything else (the CFG pair, older slots) is not,
  // so it is requested before waiting on the previous step.
  const bool chain = p.flags & CONSOLVER_FLAG_CHAIN;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long v = v0 + (long long)u * blockDim.x;
    live[u] = v < p.nvec_per_sample;
    off[u] = base + v * E;
    if (live[u]) {
      r_e0[u].load(static_cast<const T*>(p.e0) + off[u] + edelta);
      if (pair) r_c[u].load(static_cast<const T*>(p.cond) + off[u]);
      if (!chain) r_x[u].load(static_cast<const TX*>(p.x) + off[u]);
#pragma unroll
      for (int j = 0; j < kOlder; ++j)
        if ((NH || j < nh - 1) && !(chain && j == 0)) r_h[u][j].load(static_cast<const T*>(p.hist[j]) + off[u] + edelta);
    }
  }

  // ---- per-sample coefficients (CTA-uniform).  Under PDL the preceding policy kernel may still be running:
  //      the bulk loads above do not depend on it, only these few floats do. ------------------------------------
  if (p.flags & (CONSOLVER_FLAG_PDL | CONSOLVER_FLAG_CHAIN)) grid_dependency_wait();
  if (chain) {
    // every predecessor is complete now: let the NEXT step start its independent loads, then fetch what the
    // previous step produced
    grid_launch_dependents();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (live[u]) {
        r_x[u].load_produced(static_cast<const TX*>(p.x) + off[u]);
        if (kOlder > 0 && (NH || 0 < nh - 1))
          r_h[u][0].load_produced(static_cast<const T*>(p.hist[0]) + off[u] + edelta);
      }
    }
  }
  const float* cf = p.coef + (long long)b * p.coef_stride;
  float c[kOlder + 1];
#pragma unroll
  for (int j = 0; j < kOlder + 1; ++j) c[j] = (j < nh && nh > 1) ? ld_produced_f32(cf + j) : 0.f;
  const bool eff_scale = p.flags & CONSOLVER_FLAG_EFF_SCALE;
  const bool x_scale = p.flags & CONSOLVER_FLAG_X_SCALE;
  const float cs0 = eff_scale ? ld_produced_f32(cf + p.order_dim) : 1.f;
  const float cs1 = x_scale ? ld_produced_f32(cf + p.order_dim + 1) : 1.f;
  const bool vpred = p.flags & CONSOLVER_FLAG_VPRED;
  const bool lowp = p.flags & CONSOLVER_FLAG_LOWP_COMBINE;
  const float g = p.guidance;
  // ---- which torch rules the reference's expressions were evaluated under (see consolver.h) ----------------------
  constexpr bool k16 = Elem<T>::k16;
  constexpr bool host = HOST;                                  // CPU-torch: true division, scalars rounded to 16 bit
  const bool lowc = k16 && (p.flags & CONSOLVER_FLAG_LOWP_COEF);   // coefficients are 16-bit tensors (all but the last)
  const float inv_k0 = __fdiv_rn(1.f, p.k0);                   // ATen CUDA: t / scalar == t * (1/scalar)
  // Is the combined estimate / the sample still a 16-bit TENSOR in the reference when _get_prev_sample runs?  Uniform
  // over the launch.  The estimate is promoted by the first fp32 per-sample multiplier it meets: any coefficient of an
  // fp32 policy (n_hist > 1, or a scaler), the closing coefficient of a 16-bit policy (n_hist > 1).  The sample is a
  // 16-bit tensor when the latent is stored as one, or was one before this step's promotion (X_WAS_LOWP), and stays
  // one through a 16-bit (1+s1) scaling.
  const bool e16_in = k16 && nh == 1 && (!eff_scale || lowc);
  const bool x16_in = k16 && (Elem<TX>::k16 || (p.flags & CONSOLVER_FLAG_X_WAS_LOWP)) && (!x_scale || lowc);
  // scalar (0-d host tensor) times tensor: a 16-bit product when the tensor is 16-bit
  auto smul = [&](float k, float t, bool is16) -> float {
    if (k16 && is16) return round_to<T>(__fmul_rn(host ? round_to<T>(k) : k, t));
    return __fmul_rn(k, t);
  };
  auto r16 = [&](float v, bool is16) -> float { return (k16 && is16) ? round_to<T>(v) : v; };

#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!live[u]) continue;
    Raw<T, E> r_slot;
    Raw<TO, E> r_out;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      // CFG: u + g*(c - u)                                            denoise_ppo.py:100
      float eps = r_e0[u].get(i);
      if (pair) {
        if (Elem<T>::k16) {
          // the caller's combine on 16-bit tensors rounds after every op; the python scalar g stays fp32
          const float d = round_to<T>(__fsub_rn(r_c[u].get(i), eps));
          eps = round_to<T>(__fadd_rn(eps, round_to<T>(__fmul_rn(g, d))));
        } else {
          eps = __fadd_rn(eps, __fmul_rn(g, __fsub_rn(r_c[u].get(i), eps)));
        }
        r_slot.set(i, eps);
      }
      // eff = ((0 + c0*e0) + c1*e1) + ...                              scheduler_ppo.py:263-272
      float eff;
      if (nh == 1) {
        eff = eps;
      } else if (MODE == kModeSD && lowc) {
        // 16-bit coefficient tensors: 16-bit products and partial sums, until the closing coefficient — the fp32
        // `1 - torch.sum(...)` of set_default_coefficients under autocast — promotes the sum
        float acc = round_to<T>(__fadd_rn(0.f, round_to<T>(__fmul_rn(c[0], eps))));
        eff = acc;
#pragma unroll
        for (int j = 0; j < kOlder; ++j) {
          if (NH || j < nh - 1) {
            if (j + 1 < nh - 1) {
              acc = round_to<T>(__fadd_rn(acc, round_to<T>(__fmul_rn(c[j + 1], r_h[u][j].get(i)))));
            } else {
              eff = __fadd_rn(acc, __fmul_rn(c[j + 1], r_h[u][j].get(i)));
            }
          }
        }
      } else {
        eff = __fadd_rn(0.f, __fmul_rn(c[0], eps));
#pragma unroll
        for (int j = 0; j < kOlder; ++j)
          if (NH || j < nh - 1) eff = __fadd_rn(eff, __fmul_rn(c[j + 1], r_h[u][j].get(i)));
      }
      float xs = r_x[u].get(i);
      float out;
      if (MODE == kModeSD) {
        bool e16 = e16_in;
        if (eff_scale) eff = r16(__fmul_rn(eff, cs0), e16);               // :274-277
        if (x_scale) xs = r16(__fmul_rn(xs, cs1), x16_in);                // :278
        // _get_prev_sample (scheduler_ppo.py:306-332): every `0-d scalar * tensor` is a product in the tensor's dtype,
        // every tensor-tensor op is 16-bit only when both sides are
        if (vpred) {                                                      // :316-317
          const bool b16 = e16 && x16_in;
          eff = r16(__fadd_rn(smul(p.k0, eff, e16), smul(p.k1, xs, x16_in)), b16);
          e16 = b16;
        }
        const bool b16 = e16 && x16_in;
        const float t2 = r16(__fsub_rn(xs, smul(p.k1, eff, e16)), b16);
        const float x0 = r16(host ? __fdiv_rn(t2, p.k0) : __fmul_rn(t2, inv_k0), b16);           // :323
        out = r16(__fadd_rn(smul(p.k2, x0, b16), smul(p.k3, eff, e16)), b16);                     // :329-330
      } else {
        if (eff_scale) eff = __fmul_rn(eff, cs0);                         // edit_ppo/scheduler_fmppo.py:419-424
        if (x_scale) xs = __fmul_rn(xs, cs1);
        // edit_ppo/scheduler_fmppo.py:429.  First step without scalers: `dt * model_output` is a 0-d fp32
        // tensor times a 16-bit tensor, which torch evaluates in the 16-bit dtype — dt is rounded to it, the
        // product is formed in fp32 and rounded to it — before the fp32 add with the upcast sample.
        // LOWP_COMBINE extends that to the baseline solvers' multi-term form (edit_ppo/scheduler_fm.py:430),
        // where the history sum itself is a 16-bit tensor expression.
        float prod;
        if (Elem<T>::k16 && !eff_scale && (nh == 1 || lowp)) {
          if (nh > 1) eff = Elem<T>::to_f(Elem<T>::from_f(eff));
          const float dt16 = Elem<T>::to_f(Elem<T>::from_f(p.k0));
          prod = Elem<T>::to_f(Elem<T>::from_f(__fmul_rn(dt16, eff)));
        } else {
          prod = __fmul_rn(p.k0, eff);
        }
        out = __fadd_rn(xs, prod);
      }
      r_out.set(i, out);
    }
    r_out.store(static_cast<TO*>(p.x_out) + off[u]);
    if (p.x_out2) r_out.store(static_cast<TO*>(p.x_out2) + (long long)b * p.out2_stride + (off[u] - base));
    if (pair && p.slot_out) r_slot.store(static_cast<T*>(p.slot_out) + off[u]);
  }
  // plain (non-pair) step with a ring slot requested: copy e0 through
  if (!pair && p.slot_out) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (live[u]) r_e0[u].store(static_cast<T*>(p.slot_out) + off[u]);
  }
}

template <typename T, typename TX, int NH, int MODE, int E, int U, bool HOST = false>
static int launch_one(StepParams& p, int threads, cudaStream_t stream) {
  const long long per_cta = (long long)threads * U;
  p.chunks_per_sample = (int)((p.nvec_per_sample + per_cta - 1) / per_cta);
  const long long grid = (long long)p.chunks_per_sample * p.B;
  if (grid <= 0 || grid > 0x7fffffffLL) return CONSOLVER_ERR_SIZE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (p.flags & (CONSOLVER_FLAG_PDL | CONSOLVER_FLAG_CHAIN)) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, step_kernel<T, TX, NH, MODE, E, U, HOST>, (const StepParams)p);
  return (int)e;
}

template <typename T, typename TX, int NH, int MODE>
static int launch_nh(StepParams& p, cudaStream_t stream) {
  StepLaunchCfg lc = step_launch_cfg();
  int threads = lc.threads > 0 ? lc.threads : 256;
  constexpr int E = Elem<T>::kPerVec;
  p.nvec_per_sample = p.n_per_sample / E;
  if constexpr (MODE == kModeSD) {
    if (p.flags & CONSOLVER_FLAG_HOST_SCALARS) return launch_one<T, TX, NH, MODE, E, 1, true>(p, threads, stream);
  }
  // One 16-byte vector per thread and stream at every size.  Round 1 switched to two vectors per thread once the grid
  // reached ~8 CTAs per SM; re-measured on the final kernel (profiles/launch_shape_sweep_r02.jsonl, batch 32..4096, SD
  // fp32 and FM bf16) the one-vector form is never behind and 1.5-2 % ahead at batch 256 (0.98 vs 0.967 of the copy peak),
  // so the two-vector instantiations are no longer compiled (consolver_set_step_launch still accepts unroll = 2 and runs
  // this form): 165 -> 105 step kernels in the library.
  return launch_one<T, TX, NH, MODE, E, 1>(p, threads, stream);
}

// Instantiation matrix (kept small: every entry is a separate SASS kernel in libconsolver.so):
//   128-bit path   depths 1..4 fully unrolled (NH = n_hist), deeper histories through the guarded runtime-depth form
//                  (NH = 0); the sample-strided FM form only as NH = 2 (the FLUX configuration) and NH = 0
//   scalar path    ragged sizes / unaligned pointers: only the runtime-depth form
template <typename T, typename TX, int MODE>
static int launch_step(StepParams& p, bool vec_ok, cudaStream_t stream) {
  if (!vec_ok) {
    StepLaunchCfg lc = step_launch_cfg();
    p.nvec_per_sample = p.n_per_sample;
    const int threads = lc.threads > 0 ? lc.threads : 256;
    if constexpr (MODE == kModeSD) {
      if (p.flags & CONSOLVER_FLAG_HOST_SCALARS) return launch_one<T, TX, 0, MODE, 1, 1, true>(p, threads, stream);
    }
    return launch_one<T, TX, 0, MODE, 1, 1>(p, threads, stream);
  }
  if constexpr (MODE == kModeFMStrided) {
    return p.n_hist == 2 ? launch_nh<T, TX, 2, MODE>(p, stream) : launch_nh<T, TX, 0, MODE>(p, stream);
  } else {
    switch (p.n_hist) {
      case 1: return launch_nh<T, TX, 1, MODE>(p, stream);
      case 2: return launch_nh<T, TX, 2, MODE>(p, stream);
      case 3: return launch_nh<T, TX, 3, MODE>(p, stream);
      case 4: return launch_nh<T, TX, 4, MODE>(p, stream);
      default: return launch_nh<T, TX, 0, MODE>(p, stream);
    }
  }
}

// argument validation shared by both entry points; fills p.hist
inline int fill_common(StepParams& p, const void* e0, const void* cond, void* slot_out, const void* const* hist,
                       int n_hist, const void* x, void* x_out, void* x_out2, long long out2_stride,
                       const float* coef, int coef_stride, int order_dim, int flags, int B, long long n_per_sample) {
  if (!e0 || !x || !x_out || !coef) return CONSOLVER_ERR_NULL;
  if (order_dim < 2 || order_dim > CONSOLVER_MAX_ORDER || n_hist < 1 || n_hist > order_dim) return CONSOLVER_ERR_SIZE;
  if (coef_stride < order_dim + 2 || B <= 0 || n_per_sample <= 0) return CONSOLVER_ERR_SIZE;
  if (n_hist > 1 && !hist) return CONSOLVER_ERR_NULL;
  p = StepParams{};
  p.e0 = e0; p.cond = cond; p.slot_out = slot_out; p.x = x; p.x_out = x_out;
  p.x_out2 = x_out2; p.out2_stride = out2_stride > 0 ? out2_stride : n_per_sample;
  if (x_out2 && p.out2_stride < n_per_sample) return CONSOLVER_ERR_SIZE;
  for (int j = 0; j < n_hist - 1; ++j) {
    if (!hist[j]) return CONSOLVER_ERR_NULL;
    p.hist[j] = hist[j];
  }
  p.coef = coef; p.coef_stride = coef_stride; p.order_dim = order_dim; p.n_hist = n_hist; p.flags = flags;
  p.n_per_sample = n_per_sample; p.B = B;
  p.e_stride = n_per_sample;
  return 0;
}

inline bool all_aligned(const StepParams& p) {
  bool ok = aligned16(p.e0) && aligned16(p.x) && aligned16(p.x_out);
  if (p.cond) ok = ok && aligned16(p.cond);
  if (p.slot_out) ok = ok && aligned16(p.slot_out);
  if (p.x_out2) ok = ok && aligned16(p.x_out2) && (p.out2_stride % 8 == 0);
  for (int j = 0; j < p.n_hist - 1; ++j) ok = ok && aligned16(p.hist[j]);
  return ok;
}

}  // namespace consolver
