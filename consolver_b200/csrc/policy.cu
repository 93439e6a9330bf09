// policy.cu — the coefficient policy of the ConsistencySolver step as ONE kernel launch:
//   MLP (in -> H -> H -> A*K, ReLU) + softmax per action dim + per-sample categorical draw
//   argmax_k p[a,k]/q[b,a,k] + bin-value / probability gather + mask and multistep-coefficient assembly.
// Reference: factor_net_ppo.py:137-168 (FM: edit_ppo/factor_net_ppo.py:149-180), scheduler_ppo.py:248-259,
// :165-175.  The reference evaluates the MLP on B identical rows (scheduler_ppo.py:207-210); here every CTA
// evaluates it once (75k MACs, weights stream from L2 with 128-bit loads, activations live in shared memory)
// and then serves its slice of the batch.  This is matrix-VECTOR work: every weight is used once per CTA, so
// it runs on the CUDA cores with warp-shuffle reductions — tensor cores / smem staging of the weights would
// add a pass with no reuse.  Dot products accumulate in fp64 and round once, which puts the logits within
// half an ulp of the exact value (any fp32 summation order the reference's BLAS may use is an ulp-level
// perturbation of that).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/consolver.h"
#include "step_common.cuh"

namespace consolver {

constexpr int kPolicyThreads = 512;

struct PolicyParams {
  const float *w1, *b1, *w2, *b2, *w3, *b3, *action_values;
  float x0, x1, x_div, temp;
  const float* feat;
  int n_feat;
  const float* q;
  const long long* idx_in;
  int B, H, A, K, order_dim, scaler_dim, n_hist;
  float* probs_table;
  long long* idx;
  float *actions, *act_probs, *act_logp, *masks, *coef;
  int samples_per_cta;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// y[r] = act(b[r] + W[r,:] . x) for r in [0,R): one warp per row, ROWS rows in flight per warp.
template <bool RELU>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                            const float* x_s, float* y_s, int R, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const bool vec = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0);
  constexpr int ROWS = 4;
  for (int r0 = warp * ROWS; r0 < R; r0 += nwarp * ROWS) {
    double acc[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) acc[i] = 0.0;
    if (vec) {
      const int C4 = C >> 2;
      for (int c = lane; c < C4; c += 32) {
        float4 w[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
          w[i] = (r0 + i < R) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(r0 + i) * C) + c)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 xv = reinterpret_cast<const float4*>(x_s)[c];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          acc[i] = fma((double)w[i].x, (double)xv.x, acc[i]);
          acc[i] = fma((double)w[i].y, (double)xv.y, acc[i]);
          acc[i] = fma((double)w[i].z, (double)xv.z, acc[i]);
          acc[i] = fma((double)w[i].w, (double)xv.w, acc[i]);
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
          if (r0 + i < R) acc[i] = fma((double)__ldg(W + (size_t)(r0 + i) * C + c), (double)x_s[c], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const double s = warp_sum(acc[i]);
      if (lane == 0 && r0 + i < R) {
        float v = (float)(s + (double)__ldg(bias + r0 + i));
        y_s[r0 + i] = RELU ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// MLP + softmax for one input row held in x_s[0..in_dim); leaves probs in p_s[0..A*K).
__device__ __forceinline__ void mlp_softmax(const PolicyParams& p, int in_dim, const float* x_s, float* h1_s,
                                            float* h2_s, float* lg_s, float* p_s) {
  const int H = p.H, AK = p.A * p.K;
  // layer 0: in_dim is 2 (or 2 + order_dim - 1): one thread per hidden unit
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < in_dim; ++i) acc = fma((double)__ldg(p.w1 + j * in_dim + i), (double)x_s[i], acc);
    h1_s[j] = fmaxf((float)(acc + (double)__ldg(p.b1 + j)), 0.f);
  }
  __syncthreads();
  dense_layer<true>(p.w2, p.b2, h1_s, h2_s, H, H);
  __syncthreads();
  dense_layer<false>(p.w3, p.b3, h2_s, lg_s, AK, H);
  __syncthreads();
  // softmax(logits / temp) per action dim: one warp per dim
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int a = warp; a < p.A; a += nwarp) {
    float* l = lg_s + a * p.K;
    float m = -INFINITY;
    for (int k = lane; k < p.K; k += 32) {
      const float v = __fdiv_rn(l[k], p.temp);
      l[k] = v;
      m = fmaxf(m, v);
    }
    m = warp_max(m);
    double s = 0.0;
    for (int k = lane; k < p.K; k += 32) {
      const float e = expf(__fsub_rn(l[k], m));
      l[k] = e;
      s += (double)e;
    }
    const float sum = (float)warp_sum(s);
    for (int k = lane; k < p.K; k += 32) p_s[a * p.K + k] = __fdiv_rn(l[k], sum);
  }
  __syncthreads();
}

// draw / gather / coefficient assembly for sample b, probabilities in p_s
__device__ __forceinline__ void serve_sample(const PolicyParams& p, int b, const float* p_s) {
  const int A = p.A, K = p.K, od = p.order_dim;
  float act[CONSOLVER_MAX_ORDER + 4];  // first order_dim+1 action values are all the coefficients need
  for (int a = 0; a < A; ++a) {
    int best = 0;
    if (p.q) {
      const float* qr = p.q + ((size_t)b * A + a) * K;
      float bv = -INFINITY;
      for (int k = 0; k < K; ++k) {
        const float r = __fdiv_rn(p_s[a * K + k], __ldg(qr + k));   // p / q, argmax, first index on ties
        if (k == 0 || r > bv) { bv = r; best = k; }
      }
    } else {
      best = (int)p.idx_in[(size_t)b * A + a];
      best = best < 0 ? 0 : (best >= K ? K - 1 : best);
    }
    const float av = __ldg(p.action_values + a * K + best);
    const float pr = p_s[a * K + best];
    const size_t o = (size_t)b * A + a;
    if (p.idx) p.idx[o] = best;
    if (p.actions) p.actions[o] = av;
    if (p.act_probs) p.act_probs[o] = pr;
    if (p.act_logp) p.act_logp[o] = logf(__fadd_rn(pr, 1e-9f));
    if (p.masks) p.masks[o] = (a >= p.n_hist - 1 && a < od - 1) ? 0.f : 1.f;   // scheduler_ppo.py:248-249
    if (a < od + 1) act[a] = av;
  }
  // set_default_coefficients (scheduler_ppo.py:165-175): c0 = a0 + 1, c_{n-1} = 1 - sum(c_0..c_{n-2})
  float* c = p.coef + (size_t)b * (od + 2);
  const int n = p.n_hist;
  float c0 = __fadd_rn(act[0], 1.f);
  float run = c0;
  for (int i = 0; i < od; ++i) {
    float v = 0.f;
    if (n == 1) {
      v = (i == 0) ? 1.f : 0.f;                 // the step kernel bypasses the coefficient when n_hist == 1
    } else if (i == 0) {
      v = c0;
    } else if (i < n - 1) {
      v = act[i];
      run = __fadd_rn(run, v);
    } else if (i == n - 1) {
      v = __fsub_rn(1.f, run);
    }
    c[i] = v;
  }
  c[od] = p.scaler_dim >= 1 ? __fadd_rn(act[od - 1], 1.f) : 1.f;
  c[od + 1] = p.scaler_dim >= 2 ? __fadd_rn(act[od], 1.f) : 1.f;
}

__global__ void __launch_bounds__(kPolicyThreads) policy_kernel(const PolicyParams p) {
  // let a dependent (PDL) step kernel start its bulk loads right away
  grid_launch_dependents();
  extern __shared__ float smem[];
  const int H = p.H, AK = p.A * p.K;
  float* x_s = smem;                       // [CONSOLVER_MAX_IN]
  float* h1_s = x_s + CONSOLVER_MAX_IN;    // [H]
  float* h2_s = h1_s + H;                  // [H]
  float* lg_s = h2_s + H;                  // [AK]
  float* p_s = lg_s + AK;                  // [AK]
  const int in_dim = 2 + p.n_feat;
  const int b_begin = blockIdx.x * p.samples_per_cta;
  const int b_end = min(p.B, b_begin + p.samples_per_cta);

  if (p.feat == nullptr) {
    if (threadIdx.x == 0) {
      x_s[0] = __fdiv_rn(p.x0, p.x_div);   // normalize_input: x.float() / 999.0 (identity for FM)
      x_s[1] = __fdiv_rn(p.x1, p.x_div);
    }
    __syncthreads();
    mlp_softmax(p, in_dim, x_s, h1_s, h2_s, lg_s, p_s);
    if (blockIdx.x == 0 && p.probs_table)
      for (int i = threadIdx.x; i < AK; i += blockDim.x) p.probs_table[i] = p_s[i];
    for (int b = b_begin + threadIdx.x; b < b_end; b += blockDim.x) serve_sample(p, b, p_s);
  } else {
    // use_conv: per-sample features -> per-sample MLP (one sample at a time per CTA)
    for (int b = b_begin; b < b_end; ++b) {
      if (threadIdx.x == 0) {
        x_s[0] = __fdiv_rn(p.x0, p.x_div);
        x_s[1] = __fdiv_rn(p.x1, p.x_div);
      }
      if (threadIdx.x < p.n_feat) x_s[2 + threadIdx.x] = __ldg(p.feat + (size_t)b * p.n_feat + threadIdx.x);
      __syncthreads();
      mlp_softmax(p, in_dim, x_s, h1_s, h2_s, lg_s, p_s);
      if (threadIdx.x == 0) serve_sample(p, b, p_s);
      __syncthreads();
    }
  }
}

static int launch_policy(const PolicyParams& pp, cudaStream_t stream) {
  PolicyParams p = pp;
  if (!p.w1 || !p.b1 || !p.w2 || !p.b2 || !p.w3 || !p.b3 || !p.action_values || !p.coef) return CONSOLVER_ERR_NULL;
  if ((p.q == nullptr) == (p.idx_in == nullptr)) return CONSOLVER_ERR_NULL;  // exactly one of them
  if (p.B <= 0 || p.H <= 0 || p.H > CONSOLVER_MAX_HIDDEN || p.A <= 0 || p.K <= 0 ||
      (long long)p.A * p.K > CONSOLVER_MAX_LOGITS)
    return CONSOLVER_ERR_SIZE;
  if (p.order_dim < 2 || p.order_dim > CONSOLVER_MAX_ORDER || p.scaler_dim < 0 || p.scaler_dim > 2 ||
      p.n_hist < 1 || p.n_hist > p.order_dim || p.A < p.order_dim + p.scaler_dim - 1)
    return CONSOLVER_ERR_SIZE;
  if (p.n_feat < 0 || 2 + p.n_feat > CONSOLVER_MAX_IN || (p.n_feat > 0 && !p.feat)) return CONSOLVER_ERR_SIZE;
  if (p.n_feat == 0) p.feat = nullptr;
  if (!(p.temp > 0.f) || !(p.x_div != 0.f)) return CONSOLVER_ERR_SIZE;
  // shared row: 512 samples per CTA (every CTA re-derives the 33-float table from L2-resident weights);
  // per-sample MLP: spread the batch over the SMs
  p.samples_per_cta = p.feat ? (p.B + 147) / 148 : kPolicyThreads;
  if (p.samples_per_cta < 1) p.samples_per_cta = 1;
  const int grid = (p.B + p.samples_per_cta - 1) / p.samples_per_cta;
  const size_t smem = (size_t)(CONSOLVER_MAX_IN + 2 * p.H + 2 * p.A * p.K) * sizeof(float);
  policy_kernel<<<grid, kPolicyThreads, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace consolver

using namespace consolver;

extern "C" int consolver_abi_version(void) { return CONSOLVER_ABI_VERSION; }

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
                                    float* probs_table, int64_t* idx, float* actions, float* act_probs,
                                    float* act_logp, float* masks, float* coef, consolver_stream_t stream) {
  PolicyParams p = {};
  p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3; p.action_values = action_values;
  p.x0 = x0; p.x1 = x1; p.x_div = x_div; p.temp = temp;
  p.feat = feat; p.n_feat = n_feat; p.q = q; p.idx_in = reinterpret_cast<const long long*>(idx_in);
  p.B = B; p.H = H; p.A = A; p.K = K; p.order_dim = order_dim; p.scaler_dim = scaler_dim; p.n_hist = n_hist;
  p.probs_table = probs_table; p.idx = reinterpret_cast<long long*>(idx); p.actions = actions;
  p.act_probs = act_probs; p.act_logp = act_logp; p.masks = masks; p.coef = coef;
  return launch_policy(p, static_cast<cudaStream_t>(stream));
}

extern "C" int consolver_sd_policy_and_step(const float* w1, const float* b1, const float* w2, const float* b2,
                                            const float* w3, const float* b3, const float* action_values,
                                            float x0, float x1, float x_div, float temp,
                                            const float* q, const int64_t* idx_in,
                                            int H, int A, int K, int scaler_dim,
                                            float* probs_table, int64_t* idx, float* actions, float* act_probs,
                                            float* act_logp, float* masks, float* coef,
                                            int dtype, const void* e0, const void* cond, float guidance,
                                            void* slot_out, const void* const* hist, int n_hist, const void* x,
                                            void* x_out, int order_dim, float sa_t, float sb_t, float sa_p,
                                            float sb_p, int flags, int B, int64_t n_per_sample,
                                            consolver_stream_t stream) {
  int rc = consolver_policy_f32(w1, b1, w2, b2, w3, b3, action_values, x0, x1, x_div, temp, nullptr, 0, q, idx_in,
                                B, H, A, K, order_dim, scaler_dim, n_hist, probs_table, idx, actions, act_probs,
                                act_logp, masks, coef, stream);
  if (rc) return rc;
  int f = flags;
  if (scaler_dim >= 1) f |= CONSOLVER_FLAG_EFF_SCALE;
  if (scaler_dim >= 2) f |= CONSOLVER_FLAG_X_SCALE;
  return consolver_step_sd(dtype, e0, cond, guidance, slot_out, hist, n_hist, x, x_out, coef,
                           CONSOLVER_COEF_STRIDE(order_dim), order_dim, sa_t, sb_t, sa_p, sb_p, f, B, n_per_sample,
                           stream);
}
