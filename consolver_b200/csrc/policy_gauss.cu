// policy_gauss.cu — the CONTINUOUS (Gaussian) coefficient policy: ppo_type != "discrete".
//
// EXTENSION, PARITY UNPINNED.  The reference instantiates `FactorNetPPOContinous` for this mode (scheduler_ppo.py:23,
// :139) but ships no source for it (edit_ppo/scheduler_fmppo.py:169-170 is `assert 0`), so there is nothing to
// restate and nothing to pin against.  The semantics below are this repo's own, chosen so that everything downstream —
// masks, coefficient assembly (scheduler_ppo.py:248-259,:165-175), the fused step — is the code the discrete policy
// uses, and they are tested against closed forms (torch.distributions.Normal), not against the reference:
//
//   head      the same trunk 2 -> H -> H, last layer -> 2*A raw outputs: r_mean[a], r_logstd[a]
//   range     per action dim [lo_a, hi_a] = the value range of the discrete policy's bins (factor_net_ppo.py:87-102)
//   mean_a    = mid_a + half_a * tanh(r_mean[a])                  mid = (lo+hi)/2, half = (hi-lo)/2
//   std_a     = half_a * exp(clamp(r_logstd[a], -7, 1))
//   draw      action = mean + std * z,   z ~ N(0,1): ONE torch.randn([B, A]) per step from the default CUDA generator,
//             regenerated in the kernel (Philox4x32-10 + curand's Box-Muller, ATen's thread->element mapping), so a seed
//             gives the actions torch.randn would give
//   logp      = -z^2/2 - log(std) - log(2*pi)/2                   `probs` returned to the caller = exp(logp) (a density)
//
// One fused kernel: every CTA evaluates the MLP once (matrix-vector work on the CUDA cores, see policy.cu) and serves its
// slice of the batch.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>
#include <stdint.h>

#include "../../include/consolver.h"
#include "mlp_device.cuh"
#include "step_common.cuh"

namespace consolver {

constexpr int kGaussThreads = 512;

struct GaussParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const float* range;       // [A,2] = {lo, hi}
  float x0, x1, x_div;
  const float* z_in;        // nullable [B*A]: supplied N(0,1) values
  const float* actions_in;  // nullable [B*A]: forced actions (PPO replay): z = (a - mean) / std
  int rng_mode;
  unsigned long long rng_seed, rng_offset;
  const unsigned long long* rng_state;
  unsigned int rng_nthreads;
  int B, H, A, order_dim, scaler_dim, n_hist, flags, samples_per_cta;
  float* mean_std;          // nullable [2,A]: mean row, std row
  float* z_out;             // nullable [B*A]
  float *actions, *act_probs, *act_logp, *masks, *coef;
};

// The value torch.randn(numel, device="cuda") writes at linear index `li` for generator state (seed, offset): ATen's
// normal_ -> distribution_nullary_kernel (unroll 4) calls curand_normal4 once per 4 elements; thread `idx` of `nthreads`
// serves elements idx + nthreads*(4*j + ii), its j-th call uses Philox counter {offset/4 + j, idx}, and curand_normal4
// maps the 4 outputs to two Box-Muller pairs: (n0,n1) = BM(o.x,o.y), (n2,n3) = BM(o.z,o.w).
__device__ __forceinline__ float torch_randn_at(unsigned long long seed, unsigned long long offset, unsigned int nthreads,
                                                unsigned long long li) {
  const unsigned long long idx = li % nthreads;
  const unsigned long long r = li / nthreads;
  const unsigned long long lo = (offset >> 2) + (r >> 2);
  const uint4 ctr = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)idx, (unsigned)(idx >> 32));
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const uint4 o = curand_Philox4x32_10(ctr, key);
  const unsigned ii = (unsigned)(r & 3);
  const float2 bm = (ii < 2) ? _curand_box_muller(o.x, o.y) : _curand_box_muller(o.z, o.w);
  return (ii & 1) ? bm.y : bm.x;
}

__global__ void __launch_bounds__(kGaussThreads) policy_gauss_kernel(const GaussParams p) {
  grid_launch_dependents();   // a dependent (PDL) step kernel may start its bulk loads right away
  extern __shared__ __align__(16) float smem[];
  const int A = p.A, A2 = 2 * p.A;
  float* x_s = smem;
  float* h1_s = x_s + CONSOLVER_MAX_IN;
  float* h2_s = h1_s + ((p.H + 3) & ~3);
  float* raw_s = h2_s + ((p.H + 3) & ~3);          // [2A] raw head outputs, then {mean[A], std[A]}
  float* lstd_s = raw_s + ((A2 + 3) & ~3);         // [A] log(std)
  float* act_s = lstd_s + ((A + 3) & ~3);          // [spc*A]
  const int b_begin = blockIdx.x * p.samples_per_cta;
  const int nb = min(p.B - b_begin, p.samples_per_cta);

  prefetch_range_l2(p.w2, (size_t)p.H * p.H * sizeof(float));
  if (threadIdx.x == 0) {
    x_s[0] = policy_input(p.x0, p.x_div, p.flags);
    x_s[1] = policy_input(p.x1, p.x_div, p.flags);
  }
  __syncthreads();
  const MlpView mv{p.w1, p.b1, p.w2, p.b2, p.w3, p.b3, p.H, A2, 1, 1.f, p.flags};
  mlp_logits(mv, 2, x_s, h1_s, h2_s, raw_s);
  if (threadIdx.x < A) {
    const int a = threadIdx.x;
    const float lo = __ldg(p.range + 2 * a), hi = __ldg(p.range + 2 * a + 1);
    const float mid = 0.5f * (lo + hi), half = 0.5f * (hi - lo);
    const float mean = fmaf(half, tanhf(raw_s[a]), mid);
    const float ls = fminf(fmaxf(raw_s[A + a], -7.f), 1.f) + logf(half);
    lstd_s[a] = ls;
    raw_s[a] = mean;                 // each thread overwrites only the two raw entries it has just read
    raw_s[A + a] = expf(ls);
    if (blockIdx.x == 0 && p.mean_std) {
      p.mean_std[a] = mean;
      p.mean_std[A + a] = expf(ls);
    }
  }
  __syncthreads();
  const unsigned long long seed = p.rng_state ? p.rng_state[0] : p.rng_seed;
  const unsigned long long off = p.rng_offset + (p.rng_state ? p.rng_state[1] : 0ull);
  for (int pr = threadIdx.x; pr < nb * A; pr += blockDim.x) {
    const int a = pr % A;
    const size_t o = (size_t)b_begin * A + pr;
    const float mean = raw_s[a], std = raw_s[A + a];
    float z, act;
    if (p.actions_in) {
      act = __ldg(p.actions_in + o);
      z = (act - mean) / std;
    } else {
      z = p.z_in ? __ldg(p.z_in + o) : torch_randn_at(seed, off, p.rng_nthreads, o);
      act = fmaf(std, z, mean);
    }
    const float lp = -0.5f * z * z - lstd_s[a] - 0.918938533204672742f;
    act_s[pr] = act;
    if (p.z_out) p.z_out[o] = z;
    if (p.actions) p.actions[o] = act;
    if (p.act_logp) p.act_logp[o] = lp;
    if (p.act_probs) p.act_probs[o] = expf(lp);
    if (p.masks) p.masks[o] = (a >= p.n_hist - 1 && a < p.order_dim - 1) ? 0.f : 1.f;   // scheduler_ppo.py:248-249
  }
  __syncthreads();
  for (int bl = threadIdx.x; bl < nb; bl += blockDim.x)
    write_coef_record(act_s + (size_t)bl * A, p.coef + (size_t)(b_begin + bl) * (p.order_dim + 2), p.n_hist, p.order_dim,
                      p.scaler_dim, 0,
                      (p.flags & CONSOLVER_POLICY_HOST_DIV) ? kSumSequential : (p.B == 1 ? kSumCudaSingle : kSumCudaBatch));
}

}  // namespace consolver

using namespace consolver;

extern "C" int consolver_policy_gauss_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                          const float* w3, const float* b3, const float* action_range,
                                          float x0, float x1, float x_div,
                                          const float* z, const float* actions_in, const consolver_rng_t* rng,
                                          int B, int H, int A, int order_dim, int scaler_dim, int n_hist,
                                          int policy_flags, float* mean_std, float* z_out, float* actions,
                                          float* act_probs, float* act_logp, float* masks, float* coef,
                                          consolver_stream_t stream) {
  if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !action_range || !coef) return CONSOLVER_ERR_NULL;
  if ((z != nullptr) + (actions_in != nullptr) + (rng != nullptr) != 1) return CONSOLVER_ERR_NULL;   // exactly one source
  if (B <= 0 || H <= 0 || H > CONSOLVER_MAX_HIDDEN || A <= 0 || 2 * A > CONSOLVER_MAX_LOGITS || A > kGaussThreads ||
      !(x_div != 0.f))
    return CONSOLVER_ERR_SIZE;
  if (order_dim < 2 || order_dim > CONSOLVER_MAX_ORDER || scaler_dim < 0 || scaler_dim > 2 || n_hist < 1 ||
      n_hist > order_dim || A < order_dim + scaler_dim - 1)
    return CONSOLVER_ERR_SIZE;
  if (policy_flags & (CONSOLVER_POLICY_COEF_F16 | CONSOLVER_POLICY_COEF_BF16)) return CONSOLVER_ERR_UNSUPPORTED;
  if (rng && rng->nthreads == 0) return CONSOLVER_ERR_SIZE;
  GaussParams p = {};
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.range = action_range;
  p.x0 = x0; p.x1 = x1; p.x_div = x_div; p.z_in = z; p.actions_in = actions_in;
  if (rng) {
    p.rng_mode = 1; p.rng_seed = rng->seed; p.rng_offset = rng->offset;
    p.rng_state = reinterpret_cast<const unsigned long long*>(rng->state); p.rng_nthreads = rng->nthreads;
  }
  p.B = B; p.H = H; p.A = A; p.order_dim = order_dim; p.scaler_dim = scaler_dim; p.n_hist = n_hist;
  p.flags = policy_flags;
  p.mean_std = mean_std; p.z_out = z_out; p.actions = actions; p.act_probs = act_probs; p.act_logp = act_logp;
  p.masks = masks; p.coef = coef;
  int spc = (B + sm_count() - 1) / sm_count();
  spc = spc < 64 ? (B < 64 ? B : 64) : (spc > 1024 ? 1024 : spc);
  p.samples_per_cta = spc;
  const int grid = (B + spc - 1) / spc;
  auto r4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  const size_t smem = (CONSOLVER_MAX_IN + 2 * r4(H) + r4(2 * A) + r4(A) + r4((size_t)spc * A)) * sizeof(float);
  if (smem > 48 * 1024) return CONSOLVER_ERR_SIZE;
  policy_gauss_kernel<<<grid, kGaussThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return (int)cudaGetLastError();
}
