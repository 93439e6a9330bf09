// policy.cu — the coefficient policy of the ConsistencySolver step.
//
//   table kernel   MLP (in -> H -> H -> A*K, ReLU) + softmax per action dim for R input rows, one CTA per row
//   sample kernel  per-sample categorical draw argmax_k p[a,k]/q[b,a,k] + bin-value / probability gather + mask
//                  and multistep-coefficient assembly, from a given probability table
//   fused kernel   both in one launch (every CTA evaluates the MLP once, then serves its slice of the batch)
//
// Reference: factor_net_ppo.py:137-168 (FM: edit_ppo/factor_net_ppo.py:149-180), scheduler_ppo.py:248-259,:165-175.
// The reference evaluates the MLP on B identical rows per step (scheduler_ppo.py:207-210).  The row depends only
// on the timestep grid, so the schedulers evaluate all n rows of a trajectory in ONE table launch and each step
// then runs only the sample kernel.
//
// The MLP is matrix-VECTOR work (75k MACs, every weight used once per row): it runs on the CUDA cores, one warp
// per output row with 8 rows of 128-bit weight loads in flight per warp, L2 prefetch of the whole weight set at
// kernel start, activations in shared memory, fp32 FMAs inside a lane and an fp64 butterfly across lanes.
// Tensor cores / shared-memory staging of the weights would add a pass with no reuse.
// The Exp(1) slab of a CTA is contiguous; it is prefetched into shared memory with cp.async while the MLP runs.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <math.h>
#include <stdint.h>

#include "../../include/consolver.h"
#include "mlp_device.cuh"
#include "step_common.cuh"

namespace consolver {

constexpr int kPolicyThreads = 512;
constexpr int kMaxSamplesPerCta = 512;
constexpr int kQSlabBytes = 96 * 1024;

struct PolicyParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *action_values;
  float x0, x1, x_div, temp;
  const float* x_rows;      // table kernel: [R,2] rows (overrides x0,x1)
  const float* feat;
  int n_feat;
  const float* probs_in;    // sample kernel: [A,K] table
  const float* q;
  const long long* idx_in;
  int B, H, A, K, order_dim, scaler_dim, n_hist;
  int flags;                // CONSOLVER_POLICY_*
  float* probs_table;
  long long* idx;
  float *actions, *act_probs, *act_logp, *masks, *coef;
  int samples_per_cta;
  int stage_q;              // 1: the CTA's q slab fits the shared-memory budget
  // in-kernel Exp(1) draw reproducing torch's CUDA `exponential_` stream (rng_mode != 0; q and idx_in are NULL)
  int rng_mode;
  unsigned long long rng_seed, rng_offset;
  const unsigned long long* rng_state;   // nullable device [2] = {seed, offset}: overrides rng_seed, ADDS to rng_offset
  unsigned int rng_nthreads;              // grid*block of the torch launch being reproduced
  float* q_out;                           // nullable [B*A,K]: the draw, for checking
};

// The value torch.empty(numel).exponential_(1) writes at linear index `li` for generator state (seed, offset).
// ATen's distribution_elementwise_grid_stride_kernel (unroll 4): thread `idx` of `nthreads` initialises
// Philox4x32-10 with curand_init(seed, idx, offset) and its j-th curand_uniform4 feeds elements
// idx + nthreads*(4*j + ii), ii = 0..3; the uniform u in (0,1] becomes -log(u), with log(u) replaced by -eps/2 when
// u >= 1 - eps/2 (ATen transformation::exponential).  Philox counter = {offset/4 + j (64 bit), idx (64 bit)}.
__device__ __forceinline__ float torch_exponential_at(unsigned long long seed, unsigned long long offset,
                                                      unsigned int nthreads, unsigned long long li) {
  const unsigned long long idx = li % nthreads;
  const unsigned long long r = li / nthreads;
  const unsigned long long lo = (offset >> 2) + (r >> 2);
  const uint4 ctr = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)idx, (unsigned)(idx >> 32));
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  const uint4 o = curand_Philox4x32_10(ctr, key);
  const unsigned ii = (unsigned)(r & 3);
  const unsigned x = ii == 0 ? o.x : ii == 1 ? o.y : ii == 2 ? o.z : o.w;
  const float u = _curand_uniform(x);
  const float eps = 1.1920928955078125e-07f;
  const float lg = (u >= 1.0f - eps / 2) ? -eps / 2 : __logf(u);   // at::log is __logf on the device (ATen/NumericUtils.h)
  return -lg;
}

__device__ __forceinline__ MlpView mlp_view(const PolicyParams& p) {
  return MlpView{p.w1, p.b1, p.w2, p.b2, p.w3, p.b3, p.H, p.A, p.K, p.temp, p.flags};
}

// Start the asynchronous copy of this CTA's contiguous Exp(1) slab q[b_begin*A*K ...] into shared memory.
__device__ __forceinline__ void stage_q_begin(const PolicyParams& p, int b_begin, int nb, float* q_s) {
  if (!p.stage_q) return;
  const size_t n = (size_t)nb * p.A * p.K;
  if (p.rng_mode) {
    // generate the slab instead of copying it: one Philox block per element, spread over the whole CTA
    const unsigned long long seed = p.rng_state ? p.rng_state[0] : p.rng_seed;
    const unsigned long long off = p.rng_offset + (p.rng_state ? p.rng_state[1] : 0ull);
    const unsigned long long first = (unsigned long long)b_begin * p.A * p.K;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
      const float qv = torch_exponential_at(seed, off, p.rng_nthreads, first + i);
      q_s[i] = qv;
      if (p.q_out) p.q_out[first + i] = qv;
    }
    return;
  }
  const float* src = p.q + (size_t)b_begin * p.A * p.K;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    const size_t n4 = n >> 2;
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(q_s + 4 * i, src + 4 * i);
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) cp_async4(q_s + i, src + i);
  } else {
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) cp_async4(q_s + i, src + i);
  }
  cp_async_commit();
}

// draw / gather for every (sample, action-dim) pair of the CTA, then coefficient assembly per sample.
// p_s: probability table [A,K] in shared memory; q_s: staged Exp(1) slab (or unused); act_s: [nb*A] scratch.
__device__ __forceinline__ void sample_phase(const PolicyParams& p, int b_begin, int nb, const float* p_s,
                                             const float* q_s, float* act_s, float* av_s) {
  const int A = p.A, K = p.K, od = p.order_dim;
  // bin values into shared memory (av_s aliases the logits scratch, dead after the softmax): the gather below
  // then has no dependent global load
  for (int i = threadIdx.x; i < A * K; i += blockDim.x) av_s[i] = __ldg(p.action_values + i);
  if (p.q && p.stage_q) cp_async_wait_all();
  __syncthreads();
  for (int pr = threadIdx.x; pr < nb * A; pr += blockDim.x) {
    const int a = pr % A;
    const size_t o = (size_t)b_begin * A + pr;
    int best = 0;
    if (p.stage_q) {                                               // slab staged (copied or generated) in smem
      const float* pa = p_s + a * K;
      const float* qr = q_s + (size_t)pr * K;
      float bv = -INFINITY;
      for (int k = 0; k < K; ++k) {
        const float r = __fdiv_rn(pa[k], qr[k]);                   // p / q, argmax, first index on ties
        if (k == 0 || r > bv) { bv = r; best = k; }
      }
    } else if (p.rng_mode) {
      const float* pa = p_s + a * K;
      const unsigned long long seed = p.rng_state ? p.rng_state[0] : p.rng_seed;
      const unsigned long long off = p.rng_offset + (p.rng_state ? p.rng_state[1] : 0ull);
      float bv = -INFINITY;
      for (int k = 0; k < K; ++k) {
        const float qv = torch_exponential_at(seed, off, p.rng_nthreads, (unsigned long long)o * K + k);
        if (p.q_out) p.q_out[o * K + k] = qv;
        const float r = __fdiv_rn(pa[k], qv);
        if (k == 0 || r > bv) { bv = r; best = k; }
      }
    } else if (p.q) {
      const float* pa = p_s + a * K;
      float bv = -INFINITY;
      const float* qr = p.q + o * K;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float r = __fdiv_rn(pa[k], __ldg(qr + k));
        if (k == 0 || r > bv) { bv = r; best = k; }
      }
    } else {
      best = (int)p.idx_in[o];
      best = best < 0 ? 0 : (best >= K ? K - 1 : best);
    }
    const float av = av_s[a * K + best];
    const float prb = p_s[a * K + best];
    act_s[pr] = av;
    if (p.idx) p.idx[o] = best;
    if (p.actions) p.actions[o] = av;
    if (p.act_probs) p.act_probs[o] = prb;
    if (p.act_logp) p.act_logp[o] = logf(__fadd_rn(prb, 1e-9f));
    if (p.masks) p.masks[o] = (a >= p.n_hist - 1 && a < od - 1) ? 0.f : 1.f;   // scheduler_ppo.py:248-249
  }
  __syncthreads();
  for (int bl = threadIdx.x; bl < nb; bl += blockDim.x)
    write_coef_record(act_s + (size_t)bl * A, p.coef + (size_t)(b_begin + bl) * (od + 2), p.n_hist, od, p.scaler_dim,
                      p.flags & (CONSOLVER_POLICY_COEF_F16 | CONSOLVER_POLICY_COEF_BF16),
                      (p.flags & CONSOLVER_POLICY_HOST_DIV) ? kSumSequential : (p.B == 1 ? kSumCudaSingle : kSumCudaBatch));
}

struct Smem {
  float *x, *h1, *h2, *lg, *p, *act, *q;
};
__host__ __device__ __forceinline__ size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }  // 16-byte granules
__device__ __forceinline__ Smem carve(float* base, int H, int AK, int spc, int A) {
  Smem s;
  s.x = base;
  s.h1 = s.x + CONSOLVER_MAX_IN;
  s.h2 = s.h1 + r4(H);
  s.lg = s.h2 + r4(H);
  s.p = s.lg + r4(AK);
  s.act = s.p + r4(AK);
  s.q = s.act + r4((size_t)spc * A);
  return s;
}
static size_t smem_floats(int H, int AK, int spc, int A, size_t q_floats) {
  return CONSOLVER_MAX_IN + 2 * r4(H) + 2 * r4(AK) + r4((size_t)spc * A) + q_floats;
}

// ---- fused: MLP + sampling --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolicyThreads) policy_kernel(const PolicyParams p) {
  grid_launch_dependents();   // a dependent (PDL) step kernel may start its bulk loads right away
  extern __shared__ __align__(16) float smem[];
  const int AK = p.A * p.K;
  const Smem s = carve(smem, p.H, AK, p.samples_per_cta, p.A);
  const int in_dim = 2 + p.n_feat;
  const int b_begin = blockIdx.x * p.samples_per_cta;
  const int nb = min(p.B - b_begin, p.samples_per_cta);

  prefetch_range_l2(p.w2, (size_t)p.H * p.H * sizeof(float));
  prefetch_range_l2(p.w3, (size_t)AK * p.H * sizeof(float));
  stage_q_begin(p, b_begin, nb, s.q);

  if (p.feat == nullptr) {
    if (threadIdx.x == 0) {
      s.x[0] = policy_input(p.x0, p.x_div, p.flags);   // normalize_input: x.float() / 999.0 (identity for FM)
      s.x[1] = policy_input(p.x1, p.x_div, p.flags);
    }
    __syncthreads();
    mlp_softmax(mlp_view(p), in_dim, s.x, s.h1, s.h2, s.lg, s.p);
    if (blockIdx.x == 0 && p.probs_table)
      for (int i = threadIdx.x; i < AK; i += blockDim.x) p.probs_table[i] = s.p[i];
    sample_phase(p, b_begin, nb, s.p, s.q, s.act, s.lg);
  } else {
    // use_conv: per-sample features -> per-sample MLP (samples_per_cta is small here)
    for (int bl = 0; bl < nb; ++bl) {
      const int b = b_begin + bl;
      if (threadIdx.x == 0) {
        s.x[0] = policy_input(p.x0, p.x_div, p.flags);
        s.x[1] = policy_input(p.x1, p.x_div, p.flags);
      }
      if (threadIdx.x < p.n_feat)
        s.x[2 + threadIdx.x] = round_act(__ldg(p.feat + (size_t)b * p.n_feat + threadIdx.x),
                                         p.flags & (CONSOLVER_POLICY_ACT_F16 | CONSOLVER_POLICY_ACT_BF16));
      __syncthreads();
      mlp_softmax(mlp_view(p), in_dim, s.x, s.h1, s.h2, s.lg, s.p);
      if (p.probs_table)   // per-sample tables [B,A,K]
        for (int i = threadIdx.x; i < AK; i += blockDim.x) p.probs_table[(size_t)b * AK + i] = s.p[i];
      __syncthreads();     // sample_phase overwrites the logits scratch aliased by the bin-value stage
      PolicyParams one = p;
      one.stage_q = 0;
      sample_phase(one, b, 1, s.p, s.q, s.act, s.lg);
      __syncthreads();
    }
  }
}

// ---- table only: one CTA per input row -------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolicyThreads) policy_table_kernel(const PolicyParams p) {
  grid_launch_dependents();
  extern __shared__ __align__(16) float smem[];
  const int AK = p.A * p.K;
  const Smem s = carve(smem, p.H, AK, 0, p.A);
  prefetch_range_l2(p.w2, (size_t)p.H * p.H * sizeof(float));
  prefetch_range_l2(p.w3, (size_t)AK * p.H * sizeof(float));
  if (threadIdx.x < 2) s.x[threadIdx.x] = policy_input(__ldg(p.x_rows + 2 * blockIdx.x + threadIdx.x), p.x_div, p.flags);
  __syncthreads();
  mlp_softmax(mlp_view(p), 2, s.x, s.h1, s.h2, s.lg, s.p);
  for (int i = threadIdx.x; i < AK; i += blockDim.x) p.probs_table[(size_t)blockIdx.x * AK + i] = s.p[i];
}

// ---- sampling only, from a given table -------------------------------------------------------------------------
__global__ void __launch_bounds__(kPolicyThreads) policy_sample_kernel(const PolicyParams p) {
  grid_launch_dependents();
  extern __shared__ __align__(16) float smem[];
  const int AK = p.A * p.K;
  const Smem s = carve(smem, 0, AK, p.samples_per_cta, p.A);
  const int b_begin = blockIdx.x * p.samples_per_cta;
  const int nb = min(p.B - b_begin, p.samples_per_cta);
  stage_q_begin(p, b_begin, nb, s.q);
  for (int i = threadIdx.x; i < AK; i += blockDim.x) s.p[i] = __ldg(p.probs_in + i);
  if (blockIdx.x == 0 && p.probs_table && p.probs_table != p.probs_in)
    for (int i = threadIdx.x; i < AK; i += blockDim.x) p.probs_table[i] = __ldg(p.probs_in + i);
  sample_phase(p, b_begin, nb, s.p, s.q, s.act, s.lg);
}

enum : int { kLaunchFused = 0, kLaunchTable = 1, kLaunchSample = 2 };

static int check_dims(const PolicyParams& p) {
  if (p.B <= 0 || p.A <= 0 || p.K <= 0 || (long long)p.A * p.K > CONSOLVER_MAX_LOGITS) return CONSOLVER_ERR_SIZE;
  if (p.order_dim < 2 || p.order_dim > CONSOLVER_MAX_ORDER || p.scaler_dim < 0 || p.scaler_dim > 2 ||
      p.n_hist < 1 || p.n_hist > p.order_dim || p.A < p.order_dim + p.scaler_dim - 1)
    return CONSOLVER_ERR_SIZE;
  return 0;
}

static int launch_policy(const PolicyParams& pp, int mode, int rows, cudaStream_t stream) {
  PolicyParams p = pp;
  const bool need_mlp = mode != kLaunchSample;
  if (need_mlp) {
    if (!p.w1 || !p.b1 || !p.w2 || !p.b2 || !p.w3 || !p.b3) return CONSOLVER_ERR_NULL;
    if (p.H <= 0 || p.H > CONSOLVER_MAX_HIDDEN) return CONSOLVER_ERR_SIZE;
    if (!(p.temp > 0.f) || !(p.x_div != 0.f)) return CONSOLVER_ERR_SIZE;
  }
  if (mode == kLaunchTable) {
    if (!p.x_rows || !p.probs_table) return CONSOLVER_ERR_NULL;
    if (rows <= 0 || p.A <= 0 || p.K <= 0 || (long long)p.A * p.K > CONSOLVER_MAX_LOGITS) return CONSOLVER_ERR_SIZE;
    const size_t smem = smem_floats(p.H, p.A * p.K, 0, p.A, 0) * sizeof(float);
    policy_table_kernel<<<rows, kPolicyThreads, smem, stream>>>(p);
    return (int)cudaGetLastError();
  }
  if (!p.action_values || !p.coef) return CONSOLVER_ERR_NULL;
  if (mode == kLaunchSample && !p.probs_in) return CONSOLVER_ERR_NULL;
  if ((p.q != nullptr) + (p.idx_in != nullptr) + (p.rng_mode != 0) != 1) return CONSOLVER_ERR_NULL;  // exactly one
  if (p.rng_mode && p.rng_nthreads == 0) return CONSOLVER_ERR_SIZE;
  int rc = check_dims(p);
  if (rc) return rc;
  if (p.n_feat < 0 || 2 + p.n_feat > CONSOLVER_MAX_IN || (p.n_feat > 0 && !p.feat)) return CONSOLVER_ERR_SIZE;
  if (p.n_feat == 0) p.feat = nullptr;

  const int AK = p.A * p.K;
  int spc;
  if (p.feat) {
    spc = (p.B + sm_count() - 1) / sm_count();     // per-sample MLP: spread the batch over the SMs
  } else {
    spc = kQSlabBytes / (AK * (int)sizeof(float));
    // sampling-only launches carry no per-CTA MLP cost: use more, smaller CTAs
    if (mode == kLaunchSample) spc = min(spc, max(p.rng_mode ? 16 : 64, (p.B + sm_count() - 1) / sm_count()));
    spc = max(1, min(spc, kMaxSamplesPerCta));
    spc = min(spc, p.B);
    if (spc >= 4) spc &= ~3;                       // keeps every CTA's q slab 16-byte aligned
  }
  p.samples_per_cta = spc;
  p.stage_q = ((p.q != nullptr || p.rng_mode) && !p.feat &&
               (size_t)spc * AK * sizeof(float) <= (size_t)kQSlabBytes) ? 1 : 0;
  const size_t q_floats = p.stage_q ? (size_t)spc * AK : 0;
  const int grid = (p.B + spc - 1) / spc;
  const int H = mode == kLaunchSample ? 0 : p.H;
  const size_t smem = smem_floats(H, AK, spc, p.A, q_floats) * sizeof(float);
  // per DEVICE (the attribute lives in the device's context): one process may drive several GPUs
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaFuncSetAttribute(policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(policy_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (mode == kLaunchSample)
    policy_sample_kernel<<<grid, kPolicyThreads, smem, stream>>>(p);
  else
    policy_kernel<<<grid, kPolicyThreads, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace consolver

using namespace consolver;

#include "abi_hash.h"   // generated from include/consolver.h by consolver_b200/build.py
extern "C" int consolver_abi_version(void) { return CONSOLVER_ABI_VERSION; }
extern "C" uint64_t consolver_abi_hash(void) { return CONSOLVER_ABI_HEADER_HASH; }

extern "C" const char* consolver_error_string(int err) {
  switch (err) {
    case 0: return "ok";
    case CONSOLVER_ERR_NULL: return "consolver: required pointer is NULL (or both/neither of q, idx_in given)";
    case CONSOLVER_ERR_SIZE: return "consolver: size / dimension argument out of range";
    case CONSOLVER_ERR_UNSUPPORTED: return "consolver: unsupported configuration";
    case CONSOLVER_ERR_DTYPE: return "consolver: unsupported element type";
    default: return err > 0 ? cudaGetErrorString(static_cast<cudaError_t>(err)) : "consolver: unknown error";
  }
}

extern "C" int consolver_policy_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                    const float* w3, const float* b3, const float* action_values,
                                    float x0, float x1, float x_div, float temp,
                                    const float* feat, int n_feat,
                                    const float* q, const int64_t* idx_in,
                                    int B, int H, int A, int K, int order_dim, int scaler_dim, int n_hist,
                                    int policy_flags,
                                    float* probs_table, int64_t* idx, float* actions, float* act_probs,
                                    float* act_logp, float* masks, float* coef, consolver_stream_t stream) {
  PolicyParams p = {};
  p.flags = policy_flags;
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.action_values = action_values;
  p.x0 = x0; p.x1 = x1; p.x_div = x_div; p.temp = temp;
  p.feat = feat; p.n_feat = n_feat; p.q = q; p.idx_in = reinterpret_cast<const long long*>(idx_in);
  p.B = B; p.H = H; p.A = A; p.K = K; p.order_dim = order_dim; p.scaler_dim = scaler_dim; p.n_hist = n_hist;
  p.probs_table = probs_table; p.idx = reinterpret_cast<long long*>(idx); p.actions = actions;
  p.act_probs = act_probs; p.act_logp = act_logp; p.masks = masks; p.coef = coef;
  return launch_policy(p, kLaunchFused, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int consolver_policy_table_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                          const float* w3, const float* b3, const float* x_rows, int rows,
                                          float x_div, float temp, int H, int A, int K, int policy_flags,
                                          float* probs_tables, consolver_stream_t stream) {
  PolicyParams p = {};
  p.flags = policy_flags;
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
  p.x_rows = x_rows; p.x_div = x_div; p.temp = temp; p.H = H; p.A = A; p.K = K; p.probs_table = probs_tables;
  return launch_policy(p, kLaunchTable, rows, static_cast<cudaStream_t>(stream));
}

static void set_rng(PolicyParams& p, const consolver_rng_t* rng, float* q_out) {
  if (!rng) return;
  p.rng_mode = 1;
  p.rng_seed = rng->seed; p.rng_offset = rng->offset;
  p.rng_state = reinterpret_cast<const unsigned long long*>(rng->state);
  p.rng_nthreads = rng->nthreads;
  p.q_out = q_out;
}

__global__ void rng_advance_kernel(unsigned long long* state, unsigned long long amount) { state[1] += amount; }

extern "C" int consolver_rng_state_advance(uint64_t* state, uint64_t amount, consolver_stream_t stream) {
  if (!state) return CONSOLVER_ERR_NULL;
  rng_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(state),
                                                                      (unsigned long long)amount);
  return (int)cudaGetLastError();
}

extern "C" int consolver_torch_philox_plan(int64_t numel, uint32_t* nthreads, uint64_t* offset_increment) {
  // at::cuda::detail calc_execution_policy (ATen/native/cuda/DistributionTemplates.h): block 256, grid =
  // min(ceil(numel/256), SMs * maxThreadsPerSM/256), increment = ceil(numel / (256*grid*4)) * 4
  if (numel <= 0 || !nthreads || !offset_increment) return CONSOLVER_ERR_SIZE;
  int sms = 0, per_sm = 0, dev = 0;     // of the CURRENT device (two attribute reads; not on the per-step path)
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  if (sms <= 0 || per_sm <= 0) return CONSOLVER_ERR_UNSUPPORTED;
  const uint64_t n = (uint64_t)numel, block = 256;
  uint64_t grid = (n + block - 1) / block;
  const uint64_t cap = (uint64_t)sms * (uint64_t)(per_sm / (int)block);
  if (grid > cap) grid = cap;
  *nthreads = (uint32_t)(grid * block);
  *offset_increment = ((n - 1) / (block * grid * 4) + 1) * 4;
  return 0;
}

extern "C" int consolver_policy_sample_f32(const float* probs_in, const float* action_values, const float* q,
                                           const int64_t* idx_in, const consolver_rng_t* rng, float* q_out,
                                           int B, int A, int K, int order_dim,
                                           int scaler_dim, int n_hist, int policy_flags, int64_t* idx, float* actions,
                                           float* act_probs, float* act_logp, float* masks, float* coef,
                                           consolver_stream_t stream) {
  PolicyParams p = {};
  p.flags = policy_flags;
  p.probs_in = probs_in; p.action_values = action_values; p.q = q;
  p.idx_in = reinterpret_cast<const long long*>(idx_in);
  p.B = B; p.A = A; p.K = K; p.order_dim = order_dim; p.scaler_dim = scaler_dim; p.n_hist = n_hist;
  p.idx = reinterpret_cast<long long*>(idx); p.actions = actions; p.act_probs = act_probs; p.act_logp = act_logp;
  p.masks = masks; p.coef = coef;
  if (!q && !idx_in) set_rng(p, rng, q_out);
  return launch_policy(p, kLaunchSample, 0, static_cast<cudaStream_t>(stream));
}

extern "C" int consolver_sd_policy_and_step(const float* w1, const float* b1, const float* w2, const float* b2,
                                            const float* w3, const float* b3, const float* action_values,
                                            const float* probs_in,
                                            float x0, float x1, float x_div, float temp,
                                            const float* q, const int64_t* idx_in, const consolver_rng_t* rng,
                                            int H, int A, int K, int scaler_dim, int policy_flags,
                                            float* probs_table, int64_t* idx, float* actions, float* act_probs,
                                            float* act_logp, float* masks, float* coef,
                                            int dtype, const void* e0, const void* cond, float guidance,
                                            void* slot_out, const void* const* hist, int n_hist, const void* x,
                                            void* x_out, void* x_out2, int64_t out2_stride, int order_dim,
                                            float sa_t, float sb_t, float sa_p,
                                            float sb_p, int flags, int B, int64_t n_per_sample,
                                            consolver_stream_t stream) {
  int rc;
  if (probs_in) {
    rc = consolver_policy_sample_f32(probs_in, action_values, q, idx_in, rng, nullptr, B, A, K, order_dim,
                                     scaler_dim, n_hist, policy_flags, idx, actions, act_probs, act_logp, masks, coef,
                                     stream);
  } else {
    rc = consolver_policy_f32(w1, b1, w2, b2, w3, b3, action_values, x0, x1, x_div, temp, nullptr, 0, q, idx_in,
                              B, H, A, K, order_dim, scaler_dim, n_hist, policy_flags, probs_table, idx, actions,
                              act_probs, act_logp, masks, coef, stream);
  }
  if (rc) return rc;
  int f = flags;
  if (scaler_dim >= 1) f |= CONSOLVER_FLAG_EFF_SCALE;
  if (scaler_dim >= 2) f |= CONSOLVER_FLAG_X_SCALE;
  return consolver_step_sd(dtype, e0, cond, guidance, slot_out, hist, n_hist, x, x_out, x_out2, out2_stride, coef,
                           CONSOLVER_COEF_STRIDE(order_dim), order_dim, sa_t, sb_t, sa_p, sb_p, f, B, n_per_sample,
                           stream);
}
