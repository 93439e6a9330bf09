// step_sd.cu — C-ABI entry point of the fused SD (DDIM-form) solver step.  See include/consolver.h.
#include "step_kernel.cuh"

namespace consolver {
static StepLaunchCfg g_cfg = {0, 0};
StepLaunchCfg step_launch_cfg() { return g_cfg; }
}  // namespace consolver

using namespace consolver;

extern "C" int consolver_set_step_launch(int threads, int unroll) {
  if (threads != 0 && (threads < 32 || threads > 512 || (threads & 31))) return CONSOLVER_ERR_SIZE;
  if (unroll != 0 && unroll != 1 && unroll != 2) return CONSOLVER_ERR_SIZE;
  g_cfg.threads = threads;
  g_cfg.unroll = unroll;
  return 0;
}

extern "C" int consolver_step_sd(int dtype, const void* e0, const void* cond, float guidance, void* slot_out,
                                 const void* const* hist, int n_hist, const void* x, void* x_out,
                                 void* x_out2, int64_t out2_stride,
                                 const float* coef, int coef_stride, int order_dim,
                                 float sa_t, float sb_t, float sa_p, float sb_p, int flags,
                                 int B, int64_t n_per_sample, consolver_stream_t stream) {
  StepParams p;
  int rc = fill_common(p, e0, cond, slot_out, hist, n_hist, x, x_out, x_out2, (long long)out2_stride, coef,
                       coef_stride, order_dim, flags, B, (long long)n_per_sample);
  if (rc) return rc;
  p.guidance = guidance;
  p.k0 = sa_t; p.k1 = sb_t; p.k2 = sa_p; p.k3 = sb_p;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool al = all_aligned(p);
  const bool x32 = flags & CONSOLVER_FLAG_X_F32;      // fp32 latents with 16-bit model outputs (autocast pipelines)
  switch (dtype) {
    case CONSOLVER_F32:
      return launch_step<float, float, kModeSD>(p, al && n_per_sample % Elem<float>::kPerVec == 0, s);
    case CONSOLVER_F16:
      if (x32) return launch_step<__half, float, kModeSD>(p, al && n_per_sample % Elem<__half>::kPerVec == 0, s);
      return launch_step<__half, __half, kModeSD>(p, al && n_per_sample % Elem<__half>::kPerVec == 0, s);
    case CONSOLVER_BF16:
      if (x32)
        return launch_step<__nv_bfloat16, float, kModeSD>(p, al && n_per_sample % Elem<__nv_bfloat16>::kPerVec == 0, s);
      return launch_step<__nv_bfloat16, __nv_bfloat16, kModeSD>(
          p, al && n_per_sample % Elem<__nv_bfloat16>::kPerVec == 0, s);
    default:
      return CONSOLVER_ERR_DTYPE;
  }
}
