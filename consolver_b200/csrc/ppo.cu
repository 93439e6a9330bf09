// ppo.cu — PPO update of the coefficient policy as hand-written CUDA: forward + loss + backward of
//   curr_probs, entropy = factor_net(conds, actions)                       factor_net_ppo.py:170-184
//   ratio = exp(sum_a log(curr+1e-9) - sum_a log(old+1e-9)); clipped = clamp(ratio, 1-c, 1+c)
//   loss  = -min(adv*ratio, adv*clipped).mean() - entropy_coef * entropy.mean()      train_ppo.py:406-427
// in two launches, producing the gradient in the flat parameter layout (mlp.0.weight, mlp.0.bias, mlp.2.weight,
// mlp.2.bias, mlp.4.weight, mlp.4.bias) that the data-parallel all-reduce operates on.
//
// The reference evaluates the MLP on B*(n-1) replicated condition rows; only the R = n-1 rows of the timestep grid are
// distinct (scheduler_ppo.py:207-210).  Kernel 1 runs one CTA per distinct row: MLP + softmax forward (shared with
// the sampling kernels), a pass over that row's B samples accumulating d(loss)/d(prob) into a [A,K] table in shared
// memory, the entropy term, softmax backward and the three-layer MLP backward, written as a per-row partial
// gradient.  Kernel 2 sums the R partials in a fixed order (deterministic: replicas on different ranks stay
// bit-identical) and reduces the loss statistics.  Everything is latency-bound matrix-vector work on the CUDA cores.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/consolver.h"
#include "mlp_device.cuh"

namespace consolver {

constexpr int kPpoThreads = 512;
constexpr int kPpoChunk = 2048;   // samples per deterministic accumulation chunk
constexpr int kPpoStats = 4;   // per-row tail of the partial buffer: policy-loss sum, entropy sum, ratio sum, spare

struct PpoParams {
  MlpView m;
  const float* x_rows;          // [R,2]
  float x_div;
  const long long* idx;         // [R,B,A] sampled bins
  const float* old_probs;       // [R,B,A] probabilities at rollout time
  const float* adv;             // [R,B,A] advantages (already multiplied by the masks)
  int R, B;
  float clip, ent_coef;
  float* partial;               // [R, P + kPpoStats]
  long long P;
  int parts;                    // threads per (a,k) bin in the deterministic accumulation
};

__device__ __forceinline__ float block_sum_f(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (blockDim.x >> 5) ? sh[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

__global__ void __launch_bounds__(kPpoThreads) ppo_row_kernel(const PpoParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.m.H, A = p.m.A, K = p.m.K, AK = A * K;
  const int r = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  auto r4 = [](int n) { return (n + 3) & ~3; };
  float* x_s = smem;                 // [16]
  float* h1_s = x_s + 16;            // [H]   post-ReLU
  float* h2_s = h1_s + r4(H);        // [H]   post-ReLU
  float* lg_s = h2_s + r4(H);        // [AK]  scratch (exp values after the forward), then d(loss)/d(raw logit)
  float* p_s = lg_s + r4(AK);        // [AK]  probabilities
  float* dp_s = p_s + r4(AK);        // [AK]  d(loss)/d(prob)
  float* dz2_s = dp_s + r4(AK);      // [H]
  float* dz1_s = dz2_s + r4(H);      // [H]
  float* red_s = dz1_s + r4(H);      // [32]
  float* g_s = red_s + 32;           // [kPpoChunk] d loss / d log-prob per sample of the current chunk
  float* part_s = g_s + kPpoChunk;   // [AK * parts] partial sums per bin
  float* part = p.partial + (size_t)r * (p.P + kPpoStats);

  prefetch_range_l2(p.m.w2, (size_t)H * H * sizeof(float));
  prefetch_range_l2(p.m.w3, (size_t)AK * H * sizeof(float));
  if (tid < 2) x_s[tid] = policy_input(__ldg(p.x_rows + 2 * r + tid), p.x_div, p.m.flags);
  __syncthreads();
  mlp_softmax(p.m, 2, x_s, h1_s, h2_s, lg_s, p_s);

  // ---- clipped-ratio loss over this row's B samples --------------------------------------------------------------
  // Deterministic two-pass accumulation of d(loss)/d(prob[a,k]) (no float atomics: the order of the sum is fixed, so
  // the gradient is bit-reproducible run to run).  Per chunk of kChunk samples: pass 1 (one thread per sample)
  // computes the ratio, the loss terms and G_b = d loss / d log-prob; pass 2 (PARTS threads per bin) adds
  // G_b / (p + 1e-9) for the samples of its sub-range whose sampled bin it owns.
  const float M = (float)p.B * (float)p.R * (float)A;        // .mean() runs over B*R*A entries
  const float lo = 1.f - p.clip, hi = 1.f + p.clip;
  const int parts = p.parts;
  float loss_local = 0.f, ratio_local = 0.f;
  for (int i = tid; i < AK * parts; i += nt) part_s[i] = 0.f;
  for (int c0 = 0; c0 < p.B; c0 += kPpoChunk) {
    const int nb = min(kPpoChunk, p.B - c0);
    __syncthreads();
    for (int bl = tid; bl < nb; bl += nt) {
      const size_t o = ((size_t)r * p.B + c0 + bl) * A;
      float lp_new = 0.f, lp_old = 0.f;
      for (int a = 0; a < A; ++a) {
        const int k = (int)p.idx[o + a];
        lp_new += logf(p_s[a * K + k] + 1e-9f);
        lp_old += logf(__ldg(p.old_probs + o + a) + 1e-9f);
      }
      const float ratio = expf(lp_new - lp_old);
      const float clipped = fminf(fmaxf(ratio, lo), hi);
      const bool inside = ratio >= lo && ratio <= hi;          // clamp passes the gradient on the closed interval
      float gsum = 0.f;
      for (int a = 0; a < A; ++a) {
        const float ad = __ldg(p.adv + o + a);
        const float t1 = ad * ratio, t2 = ad * clipped;
        loss_local -= fminf(t1, t2);
        // d min(t1,t2)/d ratio: through t1 when it is the smaller (ties split evenly, and inside the clip range the
        // other half arrives through the clamp), nothing when the clipped branch wins outside the range
        if (inside || t1 < t2) gsum += ad;
      }
      g_s[bl] = -gsum * ratio / M;                            // d loss / d (sum_a log(p_sel + 1e-9))
      ratio_local += ratio;
    }
    __syncthreads();
    for (int t = tid; t < AK * parts; t += nt) {
      const int bin = t / parts, part = t - bin * parts;
      const int a = bin / K, k = bin - a * K;
      const float inv = 1.f / (p_s[bin] + 1e-9f);
      const int per = (nb + parts - 1) / parts;
      const int b_lo = part * per, b_hi = min(nb, b_lo + per);
      float acc = 0.f;
      for (int bl = b_lo; bl < b_hi; ++bl)
        if ((int)p.idx[((size_t)r * p.B + c0 + bl) * A + a] == k) acc += g_s[bl] * inv;
      part_s[t] += acc;
    }
  }
  __syncthreads();
  for (int bin = tid; bin < AK; bin += nt) {
    float acc = 0.f;
    for (int q = 0; q < parts; ++q) acc += part_s[bin * parts + q];
    dp_s[bin] = acc;
  }
  __syncthreads();
  const float loss_row = block_sum_f(loss_local, red_s);
  const float ratio_row = block_sum_f(ratio_local, red_s);
  __syncthreads();

  // ---- entropy bonus: -ent_coef * mean_{r,a} H(p[r,a]) / log K, H via Categorical's clamped log -------------------
  const float eps = 1.1920928955078125e-07f;
  const float ent_scale = p.ent_coef / ((float)p.R * (float)A * logf((float)K));
  float ent_local = 0.f;
  for (int i = tid; i < AK; i += nt) {
    const float pv = p_s[i];
    const float lc = logf(fminf(fmaxf(pv, eps), 1.f - eps));
    ent_local -= pv * lc;
    const float inr = (pv >= eps && pv <= 1.f - eps) ? 1.f : 0.f;
    dp_s[i] += ent_scale * (lc + inr);                          // d(-ent_scale * H)/dp = +ent_scale * (log p + 1)
  }
  const float ent_row = block_sum_f(ent_local, red_s);
  if (tid == 0) {
    part[p.P + 0] = loss_row / M;
    part[p.P + 1] = ent_row / ((float)A * logf((float)K));      // this row's mean normalised entropy
    part[p.P + 2] = ratio_row;
    part[p.P + 3] = 0.f;
  }
  __syncthreads();

  // ---- softmax backward (per action dim), temperature included: raw -> z = raw / temp -> softmax ------------------
  const int warp = tid >> 5, lane = tid & 31, nwarp = nt >> 5;
  for (int a = warp; a < A; a += nwarp) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += p_s[a * K + k] * dp_s[a * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (int k = lane; k < K; k += 32) lg_s[a * K + k] = p_s[a * K + k] * (dp_s[a * K + k] - s) / p.m.temp;
  }
  __syncthreads();

  // ---- layer 3 backward ---------------------------------------------------------------------------------------------
  const long long oW1 = 0, ob1 = 2LL * H, oW2 = 3LL * H, ob2 = oW2 + (long long)H * H, oW3 = ob2 + H,
                  ob3 = oW3 + (long long)AK * H;
  for (int i = tid; i < AK; i += nt) part[ob3 + i] = lg_s[i];
  for (int e = tid; e < AK * H; e += nt) part[oW3 + e] = lg_s[e / H] * h2_s[e % H];       // 32-bit index math
  for (int j = tid; j < H; j += nt) {
    float acc = 0.f;
    for (int q = 0; q < AK; ++q) acc = fmaf(__ldg(p.m.w3 + (size_t)q * H + j), lg_s[q], acc);
    dz2_s[j] = h2_s[j] > 0.f ? acc : 0.f;
  }
  __syncthreads();
  // ---- layer 2 backward ---------------------------------------------------------------------------------------------
  for (int j = tid; j < H; j += nt) part[ob2 + j] = dz2_s[j];
  for (int e = tid; e < H * H; e += nt) part[oW2 + e] = dz2_s[e / H] * h1_s[e % H];
  for (int i = tid; i < H; i += nt) {
    float acc = 0.f;
    for (int j = 0; j < H; ++j) acc = fmaf(__ldg(p.m.w2 + (size_t)j * H + i), dz2_s[j], acc);
    dz1_s[i] = h1_s[i] > 0.f ? acc : 0.f;
  }
  __syncthreads();
  // ---- layer 1 backward ---------------------------------------------------------------------------------------------
  for (int i = tid; i < H; i += nt) {
    part[ob1 + i] = dz1_s[i];
    part[oW1 + 2 * i] = dz1_s[i] * x_s[0];
    part[oW1 + 2 * i + 1] = dz1_s[i] * x_s[1];
  }
}

// grad[p] = sum_r partial[r][p] in a fixed order; stats = {loss, policy_loss, mean entropy, mean ratio}
__global__ void ppo_reduce_kernel(const float* __restrict__ partial, int R, long long P, int B, float ent_coef,
                                  float* __restrict__ grad, float* __restrict__ stats) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = P + kPpoStats;
  if (i < P) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += partial[(size_t)r * stride + i];
    grad[i] = s;
  } else if (i == P) {
    float pl = 0.f, ent = 0.f, ratio = 0.f;
    for (int r = 0; r < R; ++r) {
      pl += partial[(size_t)r * stride + P];
      ent += partial[(size_t)r * stride + P + 1];
      ratio += partial[(size_t)r * stride + P + 2];
    }
    ent /= (float)R;
    stats[0] = pl - ent_coef * ent;
    stats[1] = pl;
    stats[2] = ent;
    stats[3] = ratio / ((float)R * (float)B);
  }
}

// ---- fused reduce + data-parallel exchange over NVLink peer memory ------------------------------------------------------
// The reference averages the policy gradient over ranks with one DDP all-reduce per PPO epoch (train_ppo.py:257,:430;
// edit_ppo/train_ppo.py:382): 75 041 floats = 300 KB, far below the size where a ring/tree pays off — NCCL's cost there is
// launch + protocol latency.  Here the exchange is folded into the kernel that produces the gradient, as a ONE-SHOT
// all-reduce over peer-mapped ("symmetric") buffers with PER-CTA flags (no grid-wide step anywhere):
//   phase 1  CTA c sums its 1024 elements over the per-row partials (fixed order) and stores them into THIS rank's
//            symmetric buffer (local HBM, peer-readable);
//   phase 2  `world` threads of the CTA publish `epoch` into flag [rank][c] of every peer's pad (system-scope release
//            store through NVLink, cumulative over the CTA barrier before it), then spin with system-scope acquire loads
//            on the LOCAL pad until flags [0..world)[c] carry the epoch: CTA c only ever waits for CTA c of the peers;
//   phase 3  every thread loads its 4 elements from all `world` buffers (128-bit peer loads over NVLink/NVSwitch, all
//            in flight together: one round trip), adds them IN RANK ORDER — every rank forms the bit-identical sum — and
//            writes sum * (1/world) into the flat gradient.
// Buffers are double-buffered by epoch parity: a rank overwrites parity p again two epochs later, which requires every
// peer to have published the epoch in between for the same chunk, i.e. to have finished reading it.  Flags are monotonic
// (compared with >=), one word per (rank, CTA).
struct PeerView {
  float* const* bufs;          // device array [world]: peer-mapped base pointers; each buffer = [2][stride] floats
  unsigned int* const* sig;    // device array [world]: peer-mapped flag pads, [world][gridDim.x] words each
  int rank, world;
  unsigned int epoch;
  long long parity_offset;     // floats: (epoch & 1) * stride
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)
               : "memory");
  return v;
}

constexpr int kExThreads = 256;
constexpr int kExPerCta = kExThreads * 4;     // elements per CTA

// grid = ceil((P + 1) / 1024): thread t of CTA c owns elements [1024c + 4t, +4); element index P is the statistics slot
__global__ void __launch_bounds__(kExThreads) ppo_reduce_allreduce_kernel(const float* __restrict__ partial, int R,
                                                                         long long P, int B, float ent_coef,
                                                                         float* __restrict__ grad,
                                                                         float* __restrict__ stats, const PeerView pv) {
  const long long e = (long long)blockIdx.x * kExPerCta + 4 * threadIdx.x;
  const long long stride = P + kPpoStats;
  float* mine = pv.bufs[pv.rank] + pv.parity_offset;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < R; ++r) {
    const float* row = partial + (size_t)r * stride;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (e + k < P) s[k] += row[e + k];
  }
  if (e + 3 < P) {
    *reinterpret_cast<float4*>(mine + e) = make_float4(s[0], s[1], s[2], s[3]);      // stride is a multiple of 64 floats
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (e + k < P) mine[e + k] = s[k];
  }
  if (e <= P && P < e + 4) {          // the thread whose 4-element span contains index P also reduces the statistics
    float pl = 0.f, ent = 0.f, ratio = 0.f;
    for (int r = 0; r < R; ++r) {
      pl += partial[(size_t)r * stride + P];
      ent += partial[(size_t)r * stride + P + 1];
      ratio += partial[(size_t)r * stride + P + 2];
    }
    ent /= (float)R;
    stats[0] = pl - ent_coef * ent;      // this rank's loss statistics (the reference logs them per rank as well)
    stats[1] = pl;
    stats[2] = ent;
    stats[3] = ratio / ((float)R * (float)B);
  }
  __syncthreads();
  if (threadIdx.x < pv.world) {
    const int r = threadIdx.x;
    st_release_sys(pv.sig[r] + (size_t)pv.rank * gridDim.x + blockIdx.x, pv.epoch);     // "chunk c of rank `rank` is ready"
    const unsigned int* flag = pv.sig[pv.rank] + (size_t)r * gridDim.x + blockIdx.x;
    while ((int)(ld_acquire_sys(flag) - pv.epoch) < 0) {}                               // chunk c of rank r is ready
  }
  __syncthreads();
  const float inv_world = 1.f / (float)pv.world;
  if (e + 3 < P) {
    float4 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < pv.world) v[r] = ld_relaxed_sys_v4(pv.bufs[r] + pv.parity_offset + e);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (r < pv.world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
    acc.x *= inv_world; acc.y *= inv_world; acc.z *= inv_world; acc.w *= inv_world;
    if ((reinterpret_cast<uintptr_t>(grad) & 15u) == 0) {
      *reinterpret_cast<float4*>(grad + e) = acc;
    } else {
      grad[e] = acc.x; grad[e + 1] = acc.y; grad[e + 2] = acc.z; grad[e + 3] = acc.w;
    }
  } else {
    for (long long j = e; j < P && j < e + 4; ++j) {
      float acc = 0.f;
      for (int r = 0; r < pv.world; ++r) acc += ld_relaxed_sys(pv.bufs[r] + pv.parity_offset + j);
      grad[j] = acc * inv_world;
    }
  }
}

}  // namespace consolver

using namespace consolver;

extern "C" int64_t consolver_ppo_exchange_pad_words(int H, int A, int K, int world) {
  const long long P = 4LL * H + (long long)H * H + (long long)A * K * H + (long long)A * K;
  return (int64_t)world * ((P + 1 + kExPerCta - 1) / kExPerCta);
}

extern "C" size_t consolver_ppo_workspace(int rows, int H, int A, int K) {
  const long long P = 4LL * H + (long long)H * H + (long long)A * K * H + (long long)A * K;
  return (size_t)rows * (size_t)(P + kPpoStats) * sizeof(float);
}

extern "C" int consolver_ppo_loss_grad_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                           const float* w3, const float* b3, const float* x_rows, int rows,
                                           float x_div, float temp, int H, int A, int K,
                                           const int64_t* idx, const float* old_probs, const float* advantages,
                                           int B, float clip_range, float entropy_coef,
                                           void* workspace, float* grad_flat, float* stats,
                                           consolver_stream_t stream) {
  return consolver_ppo_loss_grad_allreduce_f32(w1, b1, w2, b2, w3, b3, x_rows, rows, x_div, temp, H, A, K, idx, old_probs,
                                               advantages, B, clip_range, entropy_coef, workspace, grad_flat, stats,
                                               nullptr, stream);
}

extern "C" int consolver_ppo_loss_grad_allreduce_f32(const float* w1, const float* b1, const float* w2, const float* b2,
                                                     const float* w3, const float* b3, const float* x_rows, int rows,
                                                     float x_div, float temp, int H, int A, int K,
                                                     const int64_t* idx, const float* old_probs,
                                                     const float* advantages, int B, float clip_range,
                                                     float entropy_coef, void* workspace, float* grad_flat, float* stats,
                                                     const consolver_peers_t* peers, consolver_stream_t stream) {
  if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !x_rows || !idx || !old_probs || !advantages || !workspace ||
      !grad_flat || !stats)
    return CONSOLVER_ERR_NULL;
  if (rows <= 0 || B <= 0 || H <= 0 || H > CONSOLVER_MAX_HIDDEN || A <= 0 || K <= 0 ||
      (long long)A * K > CONSOLVER_MAX_LOGITS || !(temp > 0.f) || x_div == 0.f)
    return CONSOLVER_ERR_SIZE;
  PpoParams p = {};
  p.m = MlpView{w1, b1, w2, b2, w3, b3, H, A, K, temp, 0};   // the update runs on CUDA tensors in the reference
  p.x_rows = x_rows; p.x_div = x_div;
  p.idx = reinterpret_cast<const long long*>(idx); p.old_probs = old_probs; p.adv = advantages;
  p.R = rows; p.B = B; p.clip = clip_range; p.ent_coef = entropy_coef;
  p.partial = static_cast<float*>(workspace);
  p.P = 4LL * H + (long long)H * H + (long long)A * K * H + (long long)A * K;
  auto r4 = [](int n) { return (n + 3) & ~3; };
  p.parts = 8192 / (A * K) >= 8 ? 8 : (8192 / (A * K) >= 1 ? 8192 / (A * K) : 1);
  const size_t smem = (size_t)(16 + 4 * r4(H) + 3 * r4(A * K) + 32 + kPpoChunk + A * K * p.parts) * sizeof(float);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static bool attr_set[64] = {};     // per device: the attribute lives in the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaFuncSetAttribute(ppo_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  ppo_row_kernel<<<rows, kPpoThreads, smem, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const long long n = p.P + 1;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (peers && peers->world > 1) {
    if (!peers->buffer_ptrs_dev || !peers->signal_ptrs_dev) return CONSOLVER_ERR_NULL;
    const unsigned xgrid = (unsigned)((n + kExPerCta - 1) / kExPerCta);
    if (peers->world > 16 || peers->rank < 0 || peers->rank >= peers->world || peers->stride_floats < p.P ||
        peers->stride_floats % 4 != 0 || peers->pad_words < (int64_t)peers->world * xgrid)
      return CONSOLVER_ERR_SIZE;
    PeerView pv;
    pv.bufs = reinterpret_cast<float* const*>(peers->buffer_ptrs_dev);
    pv.sig = reinterpret_cast<unsigned int* const*>(peers->signal_ptrs_dev);
    pv.rank = peers->rank; pv.world = peers->world; pv.epoch = peers->epoch;
    pv.parity_offset = (long long)(peers->epoch & 1u) * peers->stride_floats;
    ppo_reduce_allreduce_kernel<<<xgrid, kExThreads, 0, s>>>(p.partial, rows, p.P, B, entropy_coef, grad_flat, stats, pv);
    return (int)cudaGetLastError();
  }
  ppo_reduce_kernel<<<grid, 256, 0, s>>>(p.partial, rows, p.P, B, entropy_coef, grad_flat, stats);
  return (int)cudaGetLastError();
}
