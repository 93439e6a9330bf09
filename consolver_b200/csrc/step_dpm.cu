// step_dpm.cu — fused multistep DPM-Solver(++) step with AMED direction scaling: the baseline the reference ships as
// diffusers_amed_plugin_dpmpp.py (first-order update :70-138, second-order :140-262, step :350-436), whose model-output
// conversion lives in diffusers' DPMSolverMultistepScheduler.convert_model_output (diffusers 0.26.3, not in the
// reference tree; restated in oracle/consolver_oracle.py).
//
// One launch per solver step: CFG combine of the denoiser pair, conversion of the noise prediction to the quantity the
// solver integrates (m0; the data prediction for dpmsolver++), the first/second-order update, and the write of m0 into
// the two-slot ring — every latent-sized operand read once and written once:
//   reads  u, c, x (+ m1, + m2 for the second / third-order step)    writes x', m0      => 5, 6 or 7 tensors per step
// against 3 (CFG) + 4 (convert) + 9 (update) + casts for the op-by-op version.  Same streaming skeleton as step_kernel.cuh.
#include "step_common.cuh"

namespace consolver {

struct DpmParams {
  const void* e0;
  const void* cond;   // non-null: CFG pair, e0 is the unconditional half
  void* slot_out;     // nullable: where m0 goes
  const void* m1;     // nullable: previous step's m (second-order update when present)
  const void* m2;     // nullable: the one before (third-order update when present; needs m1)
  const void* x;
  void* x_out;
  void* x_out2;
  long long out2_stride;
  float guidance;
  int convert;
  float ck0, ck1;     // DIV: m0 = (x - ck0 e) / ck1     LIN: m0 = ck1 x + ck0 e
  consolver_dpm_update_t k;  // see include/consolver.h
  long long n_per_sample, nvec_per_sample;
  int chunks_per_sample;
  int B;
};

template <typename T, typename TX, int E, int U>
__global__ void __launch_bounds__(512) dpm_step_kernel(const DpmParams p) {
  const int b = blockIdx.x / p.chunks_per_sample;
  const int chunk = blockIdx.x - b * p.chunks_per_sample;
  const long long base = (long long)b * p.n_per_sample;
  const long long v0 = (long long)chunk * ((long long)blockDim.x * U) + threadIdx.x;
  const bool pair = p.cond != nullptr;
  const bool second = p.m1 != nullptr;
  const bool third = second && p.m2 != nullptr;

  Raw<T, E> r_e0[U], r_c[U], r_m1[U], r_m2[U];
  Raw<TX, E> r_x[U];
  long long off[U];
  bool live[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long v = v0 + (long long)u * blockDim.x;
    live[u] = v < p.nvec_per_sample;
    off[u] = base + v * E;
    if (live[u]) {
      r_e0[u].load(static_cast<const T*>(p.e0) + off[u]);
      if (pair) r_c[u].load(static_cast<const T*>(p.cond) + off[u]);
      r_x[u].load(static_cast<const TX*>(p.x) + off[u]);
      if (second) r_m1[u].load(static_cast<const T*>(p.m1) + off[u]);
      if (third) r_m2[u].load(static_cast<const T*>(p.m2) + off[u]);
    }
  }
  const float g = p.guidance;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!live[u]) continue;
    Raw<T, E> r_m0, r_out;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      float eps = r_e0[u].get(i);
      if (pair) {
        eps = __fadd_rn(eps, __fmul_rn(g, __fsub_rn(r_c[u].get(i), eps)));      // gen_pretrain/pipeline.py:1069-1071
        if (Elem<T>::k16) eps = Elem<T>::to_f(Elem<T>::from_f(eps));
      }
      const float xs = r_x[u].get(i);
      float m0;
      if (p.convert == CONSOLVER_DPM_CONVERT_DIV) {
        m0 = __fdiv_rn(__fsub_rn(xs, __fmul_rn(p.ck0, eps)), p.ck1);
      } else if (p.convert == CONSOLVER_DPM_CONVERT_DIV_RECIP) {
        m0 = __fmul_rn(__fsub_rn(xs, __fmul_rn(p.ck0, eps)), __fdiv_rn(1.f, p.ck1));   // ATen CUDA: t / scalar
      } else if (p.convert == CONSOLVER_DPM_CONVERT_LIN) {
        m0 = __fadd_rn(__fmul_rn(p.ck1, xs), __fmul_rn(p.ck0, eps));
      } else {
        m0 = eps;
      }
      if (Elem<T>::k16) m0 = Elem<T>::to_f(Elem<T>::from_f(m0));                // the value the ring keeps
      r_m0.set(i, m0);
      float out = __fsub_rn(__fmul_rn(p.k.cx, xs), __fmul_rn(p.k.a0, m0));     // plugin :121 / :205-207 / :335-336
      if (third) {
        const float m1 = r_m1[u].get(i);
        const float d10 = __fmul_rn(p.k.rinv, __fsub_rn(m0, m1));                // plugin :328
        const float d11 = __fmul_rn(p.k.rinv1, __fsub_rn(m1, r_m2[u].get(i)));
        const float dd = __fsub_rn(d10, d11);
        const float d1 = __fadd_rn(d10, __fmul_rn(p.k.w, dd));                   // :329
        const float d2 = __fmul_rn(p.k.rs, dd);                                  // :330
        out = __fsub_rn(__fsub_rn(out, __fmul_rn(p.k.a1, d1)), __fmul_rn(p.k.a2, d2));   // :337-338 / :345-346
      } else if (second) {
        const float d1 = __fmul_rn(p.k.rinv, __fsub_rn(m0, r_m1[u].get(i)));     // plugin :201
        out = __fsub_rn(out, __fmul_rn(p.k.a1, d1));                              // plugin :208
      }
      r_out.set(i, out);
    }
    r_out.store(static_cast<T*>(p.x_out) + off[u]);
    if (p.x_out2) r_out.store(static_cast<T*>(p.x_out2) + (long long)b * p.out2_stride + (off[u] - base));
    if (p.slot_out) r_m0.store(static_cast<T*>(p.slot_out) + off[u]);
  }
}

template <typename T, typename TX, int E, int U>
static int launch_dpm_one(DpmParams& p, int threads, cudaStream_t stream) {
  const long long per_cta = (long long)threads * U;
  p.chunks_per_sample = (int)((p.nvec_per_sample + per_cta - 1) / per_cta);
  const long long grid = (long long)p.chunks_per_sample * p.B;
  if (grid <= 0 || grid > 0x7fffffffLL) return CONSOLVER_ERR_SIZE;
  dpm_step_kernel<T, TX, E, U><<<(unsigned)grid, threads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename T, typename TX>
static int launch_dpm(DpmParams& p, bool vec_ok, cudaStream_t stream) {
  StepLaunchCfg lc = step_launch_cfg();
  const int threads = lc.threads > 0 ? lc.threads : 256;
  if (!vec_ok) {
    p.nvec_per_sample = p.n_per_sample;
    return launch_dpm_one<T, TX, 1, 1>(p, threads, stream);
  }
  constexpr int E = Elem<T>::kPerVec;
  p.nvec_per_sample = p.n_per_sample / E;
  const long long ctas_u2 = ((p.nvec_per_sample + 2LL * threads - 1) / (2LL * threads)) * p.B;
  const int unroll = lc.unroll > 0 ? lc.unroll : (ctas_u2 >= (long long)sm_count() * 8 ? 2 : 1);
  if (unroll >= 2) return launch_dpm_one<T, TX, E, 2>(p, threads, stream);
  return launch_dpm_one<T, TX, E, 1>(p, threads, stream);
}

}  // namespace consolver

using namespace consolver;

extern "C" int consolver_step_dpm(int dtype, int x_dtype, const void* e0, const void* cond, float guidance,
                                  void* slot_out, const void* m1, const void* m2, const void* x, void* x_out,
                                  void* x_out2, int64_t out2_stride, int convert, float ck0, float ck1,
                                  const consolver_dpm_update_t* upd, int B, int64_t n_per_sample,
                                  consolver_stream_t stream) {
  if (!e0 || !x || !x_out || !upd) return CONSOLVER_ERR_NULL;
  if (m2 && !m1) return CONSOLVER_ERR_NULL;
  if (B <= 0 || n_per_sample <= 0) return CONSOLVER_ERR_SIZE;
  if (convert < CONSOLVER_DPM_CONVERT_NONE || convert > CONSOLVER_DPM_CONVERT_DIV_RECIP) return CONSOLVER_ERR_UNSUPPORTED;
  if (x_dtype != dtype && x_dtype != CONSOLVER_F32) return CONSOLVER_ERR_DTYPE;
  DpmParams p{};
  p.e0 = e0; p.cond = cond; p.slot_out = slot_out; p.m1 = m1; p.m2 = m2; p.x = x; p.x_out = x_out; p.x_out2 = x_out2;
  p.out2_stride = out2_stride > 0 ? (long long)out2_stride : (long long)n_per_sample;
  if (x_out2 && p.out2_stride < (long long)n_per_sample) return CONSOLVER_ERR_SIZE;
  p.guidance = guidance; p.convert = convert; p.ck0 = ck0; p.ck1 = ck1;
  p.k = *upd;
  p.n_per_sample = (long long)n_per_sample; p.B = B;
  bool al = aligned16(e0) && aligned16(x) && aligned16(x_out) && (!cond || aligned16(cond)) &&
            (!slot_out || aligned16(slot_out)) && (!m1 || aligned16(m1)) && (!m2 || aligned16(m2)) &&
            (!x_out2 || (aligned16(x_out2) && p.out2_stride % 8 == 0));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case CONSOLVER_F32:
      return launch_dpm<float, float>(p, al && n_per_sample % 4 == 0, s);
    case CONSOLVER_F16:
      if (x_dtype == CONSOLVER_F32) return launch_dpm<__half, float>(p, al && n_per_sample % 8 == 0, s);
      return launch_dpm<__half, __half>(p, al && n_per_sample % 8 == 0, s);
    case CONSOLVER_BF16:
      if (x_dtype == CONSOLVER_F32) return launch_dpm<__nv_bfloat16, float>(p, al && n_per_sample % 8 == 0, s);
      return launch_dpm<__nv_bfloat16, __nv_bfloat16>(p, al && n_per_sample % 8 == 0, s);
    default:
      return CONSOLVER_ERR_DTYPE;
  }
}
