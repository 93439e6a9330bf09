// features.cu — cosine-similarity features of the `use_conv=True` policy input (factor_net_ppo.py:108-130,:146-149):
// for every sample, cos(history slot j, newest model output) for j = 1 .. order_dim-1 over all C*H*W elements
// (zero-padded slots give 0).  This is the first pass of the two-pass use_conv step: a streaming reduction that
// reads the newest output (or forms it from the CFG pair on the fly) and the older slots once; the policy kernel
// then runs its MLP per sample on [t, t_prev, cos_1..] and the fused step kernel makes the second pass.
//
// One CTA reduces one chunk of one sample (fp32 per-thread partials, fp64 across threads), adds its partial dot
// products / squared norms to a per-sample fp64 accumulator with atomics, and the last CTA of a sample (ticket
// counter) turns the sums into cosines:  dot / (max(|a|, eps) * max(|b|, eps)),  eps = 1e-8 as
// torch.nn.functional.cosine_similarity.
#include "step_common.cuh"

namespace consolver {

constexpr int kFeatThreads = 256;

struct FeatParams {
  const void* e0;
  const void* cond;
  const void* hist[kMaxOlder];
  float guidance;
  int n_hist, order_dim, B, chunks_per_sample;
  long long n_per_sample, nvec_per_sample;
  double* acc;          // [B, 2*order_dim - 1]: norm0, then (dot_j, norm_j) for j = 1..order_dim-1
  unsigned int* ticket; // [B]
  float* feat;          // [B, order_dim-1]
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

template <typename T, int E>
__global__ void __launch_bounds__(kFeatThreads) cosine_features_kernel(const FeatParams p) {
  __shared__ double sh[kFeatThreads / 32];
  __shared__ bool last;
  const int b = blockIdx.x / p.chunks_per_sample;
  const int chunk = blockIdx.x - b * p.chunks_per_sample;
  const long long base = (long long)b * p.n_per_sample;
  const int nold = p.n_hist - 1;
  const bool pair = p.cond != nullptr;
  float n0 = 0.f, dot[kMaxOlder], nj[kMaxOlder];
#pragma unroll
  for (int j = 0; j < kMaxOlder; ++j) dot[j] = nj[j] = 0.f;

  for (long long v = (long long)chunk * blockDim.x + threadIdx.x; v < p.nvec_per_sample;
       v += (long long)p.chunks_per_sample * blockDim.x) {
    const long long off = base + v * E;
    Raw<T, E> a, c, h[kMaxOlder];
    a.load(static_cast<const T*>(p.e0) + off);
    if (pair) c.load(static_cast<const T*>(p.cond) + off);
#pragma unroll
    for (int j = 0; j < kMaxOlder; ++j)
      if (j < nold) h[j].load(static_cast<const T*>(p.hist[j]) + off);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      float e = a.get(i);
      if (pair) {
        e = __fadd_rn(e, __fmul_rn(p.guidance, __fsub_rn(c.get(i), e)));
        if (Elem<T>::k16) e = Elem<T>::to_f(Elem<T>::from_f(e));
      }
      n0 = fmaf(e, e, n0);
#pragma unroll
      for (int j = 0; j < kMaxOlder; ++j)
        if (j < nold) {
          const float hv = h[j].get(i);
          dot[j] = fmaf(e, hv, dot[j]);
          nj[j] = fmaf(hv, hv, nj[j]);
        }
    }
  }
  double* acc = p.acc + (size_t)b * (2 * p.order_dim - 1);
  double s = block_sum((double)n0, sh);
  if (threadIdx.x == 0) atomicAdd(acc, s);
  for (int j = 0; j < nold; ++j) {
    s = block_sum((double)dot[j], sh);
    if (threadIdx.x == 0) atomicAdd(acc + 1 + 2 * j, s);
    s = block_sum((double)nj[j], sh);
    if (threadIdx.x == 0) atomicAdd(acc + 2 + 2 * j, s);
  }
  // last CTA of this sample finalises
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket + b, 1u);
    last = (t == (unsigned)p.chunks_per_sample - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < p.order_dim - 1) {
    __threadfence();
    const int j = threadIdx.x;
    float f = 0.f;
    if (j < nold) {
      const volatile double* va = acc;
      const double na = fmax(sqrt(va[0]), 1e-8), nb = fmax(sqrt(va[2 + 2 * j]), 1e-8);
      f = (float)(va[1 + 2 * j] / (na * nb));
    }
    p.feat[(size_t)b * (p.order_dim - 1) + j] = f;
  }
}

template <typename T>
static int launch_feat(FeatParams& p, bool vec_ok, cudaStream_t stream) {
  constexpr int E = Elem<T>::kPerVec;
  p.nvec_per_sample = vec_ok ? p.n_per_sample / E : p.n_per_sample;
  const long long per_cta = (long long)kFeatThreads * 4;                 // >= 4 vectors per thread
  long long chunks = (p.nvec_per_sample + per_cta - 1) / per_cta;
  const long long want = (2LL * sm_count() + p.B - 1) / p.B;                    // enough CTAs for ~2 per SM ...
  if (chunks > want) chunks = want < 1 ? 1 : want;                        // ... but no more than that
  p.chunks_per_sample = (int)chunks;
  const long long grid = chunks * p.B;
  if (grid > 0x7fffffffLL) return CONSOLVER_ERR_SIZE;
  if (vec_ok)
    cosine_features_kernel<T, E><<<(unsigned)grid, kFeatThreads, 0, stream>>>(p);
  else
    cosine_features_kernel<T, 1><<<(unsigned)grid, kFeatThreads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace consolver

using namespace consolver;

extern "C" size_t consolver_cosine_features_workspace(int B, int order_dim) {
  return (size_t)B * (2 * order_dim - 1) * sizeof(double) + (size_t)B * sizeof(unsigned int);
}

extern "C" int consolver_cosine_features(int dtype, const void* e0, const void* cond, float guidance,
                                         const void* const* hist, int n_hist, int order_dim, int B,
                                         int64_t n_per_sample, void* workspace, float* feat,
                                         consolver_stream_t stream) {
  if (!e0 || !workspace || !feat) return CONSOLVER_ERR_NULL;
  if (order_dim < 2 || order_dim > CONSOLVER_MAX_ORDER || n_hist < 1 || n_hist > order_dim || B <= 0 ||
      n_per_sample <= 0)
    return CONSOLVER_ERR_SIZE;
  if (n_hist > 1 && !hist) return CONSOLVER_ERR_NULL;
  FeatParams p = {};
  p.e0 = e0; p.cond = cond; p.guidance = guidance; p.n_hist = n_hist; p.order_dim = order_dim; p.B = B;
  p.n_per_sample = n_per_sample;
  p.feat = feat;
  bool al = aligned16(e0) && (!cond || aligned16(cond));
  for (int j = 0; j < n_hist - 1; ++j) {
    if (!hist[j]) return CONSOLVER_ERR_NULL;
    p.hist[j] = hist[j];
    al = al && aligned16(hist[j]);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t acc_bytes = (size_t)B * (2 * order_dim - 1) * sizeof(double);
  p.acc = static_cast<double*>(workspace);
  p.ticket = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + acc_bytes);
  cudaError_t e = cudaMemsetAsync(workspace, 0, consolver_cosine_features_workspace(B, order_dim), s);
  if (e != cudaSuccess) return (int)e;
  switch (dtype) {
    case CONSOLVER_F32: return launch_feat<float>(p, al && n_per_sample % 4 == 0, s);
    case CONSOLVER_F16: return launch_feat<__half>(p, al && n_per_sample % 8 == 0, s);
    case CONSOLVER_BF16: return launch_feat<__nv_bfloat16>(p, al && n_per_sample % 8 == 0, s);
    default: return CONSOLVER_ERR_DTYPE;
  }
}
