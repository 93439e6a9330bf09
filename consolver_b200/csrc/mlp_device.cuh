// mlp_device.cuh — device-side building blocks of the policy MLP shared by the sampling kernels (policy.cu) and the
// PPO update kernel (ppo.cu): warp reductions, L2 prefetch / cp.async helpers, the one-warp-per-row dense layer and
// the MLP + softmax forward for one input row.  See policy.cu for the design notes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/consolver.h"

namespace consolver {

struct MlpView {
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int H, A, K;
  float temp;
  int flags;   // CONSOLVER_POLICY_*
};

// the value a 16-bit torch tensor would hold: mode = CONSOLVER_POLICY_ACT_F16 / _BF16 bits (0: fp32, unchanged)
__device__ __forceinline__ float round_act(float v, int mode) {
  if (mode & (CONSOLVER_POLICY_ACT_F16 | CONSOLVER_POLICY_COEF_F16)) return __half2float(__float2half_rn(v));
  if (mode & (CONSOLVER_POLICY_ACT_BF16 | CONSOLVER_POLICY_COEF_BF16)) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}
// normalize_input (factor_net_ppo.py:104-106): x.float() / 999.0 — on CUDA tensors ATen multiplies by the fp32
// reciprocal of the python scalar; under autocast the first Linear then rounds its input to the autocast dtype
__device__ __forceinline__ float policy_input(float x, float x_div, int flags) {
  const float v = (flags & CONSOLVER_POLICY_HOST_DIV) ? __fdiv_rn(x, x_div) : __fmul_rn(x, __fdiv_rn(1.f, x_div));
  return round_act(v, flags & (CONSOLVER_POLICY_ACT_F16 | CONSOLVER_POLICY_ACT_BF16));
}

// `torch.sum(torch.stack(terms), dim=0)` of set_default_coefficients (scheduler_ppo.py:172) in the order ATen adds.
// fp32 addition is not associative, and which order the reference gets depends on where it runs (ATen/native/cuda/
// Reduce.cuh of the pinned torch, restated here; m = n_hist - 1 <= 7 terms):
//   kSumSequential  ((s0 + s1) + s2) + ...                        the reference on CPU tensors
//   kSumCudaBatch   B >= 2: the batch is the fastest dimension, every thread reduces its own sample with vt0 = 4
//                   accumulators: acc[i % 4] += s_i, result ((acc0 + acc1) + acc2) + acc3 (== sequential up to 4 terms)
//   kSumCudaSingle  B == 1: the reduced dimension is the fastest one, so block.x = last_pow2(m) threads share ONE sum:
//                   thread x accumulates s_x, s_{x+W}, ... (4 accumulators, combined in order), then a shuffle tree
//                   with DECREASING offsets — three terms give (s0 + s2) + s1
enum : int { kSumSequential = 0, kSumCudaBatch = 1, kSumCudaSingle = 2 };
static_assert(CONSOLVER_MAX_ORDER == 8, "sum_terms is unrolled for at most 7 terms");

// one thread's share of the B == 1 form: terms x, x+W, x+2W, x+3W into four accumulators, combined in order
template <int W>
__device__ __forceinline__ float sum_single_lane(const float (&s)[8], int m, int x) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (x + k * W < 8 && x + k * W < m) acc[k] = __fadd_rn(acc[k], s[(x + k * W) & 7]);
  return __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]);
}

// every loop has a constant trip count and every array index is a compile-time constant after unrolling, so s[] and the
// accumulators stay in registers (a version with run-time indices put a 96-byte stack frame into the sample kernel)
__device__ __forceinline__ float sum_terms(const float (&s)[8], int m, int mode) {
  if (mode == kSumSequential) {
    float run = s[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
      if (i < m) run = __fadd_rn(run, s[i]);
    return run;
  }
  if (mode == kSumCudaBatch) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < m) acc[i & 3] = __fadd_rn(acc[i & 3], s[i]);
    return __fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]);
  }
  if (m >= 4) {                                    // W = last_pow2(m) = 4: lanes 0..3, tree offsets 2 then 1
    const float t0 = sum_single_lane<4>(s, m, 0), t1 = sum_single_lane<4>(s, m, 1);
    const float t2 = sum_single_lane<4>(s, m, 2), t3 = sum_single_lane<4>(s, m, 3);
    return __fadd_rn(__fadd_rn(t0, t2), __fadd_rn(t1, t3));
  }
  if (m >= 2) return __fadd_rn(sum_single_lane<2>(s, m, 0), sum_single_lane<2>(s, m, 1));      // W = 2
  return sum_single_lane<1>(s, m, 0);                                                            // W = 1
}

// set_default_coefficients (scheduler_ppo.py:165-175) for one sample: act[0..A) are the sampled action values, c the
// coefficient record of include/consolver.h: c0 = a0 + 1, c_{n-1} = 1 - sum(c_0..c_{n-2}), then (1+s0), (1+s1).
// cm = CONSOLVER_POLICY_COEF_F16/_BF16: the action values are 16-bit tensors, so a0 + 1 and s + 1 are rounded to that
// dtype; torch.sum returns fp32 under autocast, so the sum and the closing coefficient are fp32.
__device__ __forceinline__ void write_coef_record(const float* act, float* c, int n, int od, int scaler_dim, int cm,
                                                  int sum_mode) {
  float terms[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    terms[i] = (i < n - 1) ? (i == 0 ? round_act(__fadd_rn(act[0], 1.f), cm) : act[i]) : 0.f;
  const float last = (n > 1) ? __fsub_rn(1.f, sum_terms(terms, n - 1, sum_mode)) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < od) {
      float v = 0.f;
      if (n == 1) {
        v = (i == 0) ? 1.f : 0.f;             // the step kernel bypasses the coefficient when n_hist == 1
      } else if (i < n - 1) {
        v = terms[i];
      } else if (i == n - 1) {
        v = last;
      }
      c[i] = v;
    }
  }
  c[od] = scaler_dim >= 1 ? round_act(__fadd_rn(act[od - 1], 1.f), cm) : 1.f;
  c[od + 1] = scaler_dim >= 2 ? round_act(__fadd_rn(act[od], 1.f), cm) : 1.f;
}

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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// pull `bytes` starting at p into L2, one 128-byte line per thread per round
__device__ __forceinline__ void prefetch_range_l2(const void* p, size_t bytes) {
  const char* c = static_cast<const char*>(p);
  for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128) prefetch_l2(c + o);
}

// y[r] = act(b[r] + W[r,:] . x) for r in [0,R): one warp per row, ROWS rows of loads in flight per warp.
template <bool RELU>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias,
                                            const float* x_s, float* y_s, int R, int C, int act_mode = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const bool vec = ((C & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0);
  constexpr int ROWS = 8;
  for (int r0 = warp * ROWS; r0 < R; r0 += nwarp * ROWS) {
    float acc[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; ++i) acc[i] = 0.f;
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
          acc[i] = fmaf(w[i].x, xv.x, acc[i]);
          acc[i] = fmaf(w[i].y, xv.y, acc[i]);
          acc[i] = fmaf(w[i].z, xv.z, acc[i]);
          acc[i] = fmaf(w[i].w, xv.w, acc[i]);
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
#pragma unroll
        for (int i = 0; i < ROWS; ++i)
          if (r0 + i < R) acc[i] = fmaf(__ldg(W + (size_t)(r0 + i) * C + c), x_s[c], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; ++i) {
      const double s = warp_sum((double)acc[i]);
      if (lane == 0 && r0 + i < R) {
        float v = round_act((float)(s + (double)__ldg(bias + r0 + i)), act_mode);
        y_s[r0 + i] = RELU ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// The three Linear layers for one input row held in x_s[0..in_dim); leaves the A*K raw outputs in lg_s.
__device__ __forceinline__ void mlp_logits(const MlpView& p, int in_dim, const float* x_s, float* h1_s, float* h2_s,
                                           float* lg_s) {
  const int H = p.H, AK = p.A * p.K;
  const int act = p.flags & (CONSOLVER_POLICY_ACT_F16 | CONSOLVER_POLICY_ACT_BF16);
  for (int j = threadIdx.x; j < H; j += blockDim.x) {     // layer 0: in_dim is 2 (or 2 + order_dim - 1)
    double acc = 0.0;
    for (int i = 0; i < in_dim; ++i) acc = fma((double)__ldg(p.w1 + j * in_dim + i), (double)x_s[i], acc);
    h1_s[j] = fmaxf(round_act((float)(acc + (double)__ldg(p.b1 + j)), act), 0.f);
  }
  __syncthreads();
  dense_layer<true>(p.w2, p.b2, h1_s, h2_s, H, H, act);
  __syncthreads();
  dense_layer<false>(p.w3, p.b3, h2_s, lg_s, AK, H, act);
  __syncthreads();
}

// MLP + softmax for one input row held in x_s[0..in_dim); leaves probs in p_s[0..A*K).
__device__ __forceinline__ void mlp_softmax(const MlpView& p, int in_dim, const float* x_s, float* h1_s,
                                            float* h2_s, float* lg_s, float* p_s) {
  const int act = p.flags & (CONSOLVER_POLICY_ACT_F16 | CONSOLVER_POLICY_ACT_BF16);
  mlp_logits(p, in_dim, x_s, h1_s, h2_s, lg_s);
  // logits / temp: a true division on CPU tensors, a multiplication by the fp32 reciprocal on CUDA tensors
  const bool host_div = p.flags & CONSOLVER_POLICY_HOST_DIV;
  const float inv_temp = __fdiv_rn(1.f, p.temp);
  // softmax(logits / temp) per action dim: one warp per dim
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int a = warp; a < p.A; a += nwarp) {
    float* l = lg_s + a * p.K;
    float m = -INFINITY;
    for (int k = lane; k < p.K; k += 32) {
      const float v = round_act(host_div ? __fdiv_rn(l[k], p.temp) : __fmul_rn(l[k], inv_temp), act);
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

}  // namespace consolver
