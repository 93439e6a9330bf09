// step_fm.cu — C-ABI entry point of the fused flow-matching (Euler-form) solver step.  See include/consolver.h.
#include "step_kernel.cuh"

using namespace consolver;

template <int MODE>
static int dispatch_fm(StepParams& p, int dtype, int x_dtype, bool al, cudaStream_t s) {
  const long long n = p.n_per_sample;
  switch (dtype) {
    case CONSOLVER_F32:
      return launch_step<float, float, MODE>(p, al && n % 4 == 0, s);
    case CONSOLVER_F16:
      if (x_dtype == CONSOLVER_F32) return launch_step<__half, float, MODE>(p, al && n % 8 == 0, s);
      return launch_step<__half, __half, MODE>(p, al && n % 8 == 0, s);
    case CONSOLVER_BF16:
      if (x_dtype == CONSOLVER_F32) return launch_step<__nv_bfloat16, float, MODE>(p, al && n % 8 == 0, s);
      return launch_step<__nv_bfloat16, __nv_bfloat16, MODE>(p, al && n % 8 == 0, s);
    default:
      return CONSOLVER_ERR_DTYPE;
  }
}

extern "C" int consolver_step_fm(int dtype, int x_dtype, const void* e0, void* slot_out,
                                 const void* const* hist, int n_hist, const void* x, void* x_out,
                                 void* x_out2, int64_t out2_stride,
                                 const float* coef, int coef_stride, int order_dim, float dt, int flags,
                                 int B, int64_t n_per_sample, consolver_stream_t stream) {
  return consolver_step_fm_strided(dtype, x_dtype, e0, 0, slot_out, hist, n_hist, x, x_out, x_out2, out2_stride, coef,
                                   coef_stride, order_dim, dt, flags, B, n_per_sample, stream);
}

extern "C" int consolver_step_fm_strided(int dtype, int x_dtype, const void* e0, int64_t e_stride, void* slot_out,
                                         const void* const* hist, int n_hist, const void* x, void* x_out,
                                         void* x_out2, int64_t out2_stride,
                                         const float* coef, int coef_stride, int order_dim, float dt, int flags,
                                         int B, int64_t n_per_sample, consolver_stream_t stream) {
  StepParams p;
  int rc = fill_common(p, e0, nullptr, slot_out, hist, n_hist, x, x_out, x_out2, (long long)out2_stride, coef,
                       coef_stride, order_dim, flags & ~CONSOLVER_FLAG_VPRED, B, (long long)n_per_sample);
  if (rc) return rc;
  if (e_stride != 0 && e_stride < n_per_sample) return CONSOLVER_ERR_SIZE;
  if (e_stride > 0) p.e_stride = (long long)e_stride;
  p.k0 = dt;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool al = all_aligned(p) && p.e_stride % 8 == 0;
  if (x_dtype != dtype && x_dtype != CONSOLVER_F32) return CONSOLVER_ERR_DTYPE;
  return p.e_stride != p.n_per_sample ? dispatch_fm<kModeFMStrided>(p, dtype, x_dtype, al, s)
                                      : dispatch_fm<kModeFM>(p, dtype, x_dtype, al, s);
}
