// Experiment (not product): what bounds ONE isolated fused-step launch at small batch?
//
// Measures, as nodes of a CUDA graph (so host launch gaps do not enter), all on one stream, serialised:
//   null      an empty kernel with the step kernel's grid                      -> per-node launch/drain floor
//   touch     one 16-byte load + one 16-byte store per thread, same grid       -> floor + one DRAM round trip
//   step      the product's fp32 step (CFG pair, n_hist 4, 6 reads + 2 writes) at B = 8 .. 4096
//   copy      cudaMemcpyAsync device-to-device moving the same number of bytes (half read, half written)
// and prints the per-launch times; tools/exp/launch_floor_fit.py fits  T(B) = a + bytes(B) / BW  to both series.
// If the intercept `a` of the step kernel equals the null/touch floor and the copy's intercept, the small-batch fraction
// is a property of launching a kernel, not of this kernel's code.
// Also: launch-shape variants at B = 64 (threads per CTA, L2 prefetch hint) to show none of them moves the number.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_floor launch_floor.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct P { const float *u, *c, *x, *h1, *h2, *h3; float *out, *slot; const float* coef; long long n_per_sample; int chunks; };

enum { HINT_NONE = 0, HINT_L2_128 = 1, HINT_L2_256 = 2 };
template <int HINT> __device__ __forceinline__ float4 ld(const float* p) {
  float4 r;
  if (HINT == HINT_NONE)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (HINT == HINT_L2_128)
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void math(float u, float c, float x, float h1, float h2, float h3, const float* cf, float g,
                                     float inv, float& out, float& eps) {
  const float sb_t = 0.5460f, sa_p = 0.9151f, sb_p = 0.4033f;
  eps = __fadd_rn(u, __fmul_rn(g, __fsub_rn(c, u)));
  float eff = __fadd_rn(0.f, __fmul_rn(cf[0], eps));
  eff = __fadd_rn(eff, __fmul_rn(cf[1], h1));
  eff = __fadd_rn(eff, __fmul_rn(cf[2], h2));
  eff = __fadd_rn(eff, __fmul_rn(cf[3], h3));
  float x0 = __fmul_rn(__fsub_rn(x, __fmul_rn(sb_t, eff)), inv);
  out = __fadd_rn(__fmul_rn(sa_p, x0), __fmul_rn(sb_p, eff));
}

template <int U, int HINT>
__global__ void __launch_bounds__(1024) k_step(const P p) {
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x - b * p.chunks;
  const long long base = (long long)b * p.n_per_sample;
  const long long v0 = (long long)chunk * (blockDim.x * U) + threadIdx.x;
  float4 ru[U], rc[U], rx[U], r1[U], r2[U], r3[U];
  long long off[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    off[i] = base + (v0 + (long long)i * blockDim.x) * 4;
    ru[i] = ld<HINT>(p.u + off[i]); rc[i] = ld<HINT>(p.c + off[i]); rx[i] = ld<HINT>(p.x + off[i]);
    r1[i] = ld<HINT>(p.h1 + off[i]); r2[i] = ld<HINT>(p.h2 + off[i]); r3[i] = ld<HINT>(p.h3 + off[i]);
  }
  float cf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cf[j] = __ldg(p.coef + b * 6 + j);
  const float inv = __fdiv_rn(1.f, 0.8378f);
#pragma unroll
  for (int i = 0; i < U; ++i) {
    float4 o, e;
    math(ru[i].x, rc[i].x, rx[i].x, r1[i].x, r2[i].x, r3[i].x, cf, 3.f, inv, o.x, e.x);
    math(ru[i].y, rc[i].y, rx[i].y, r1[i].y, r2[i].y, r3[i].y, cf, 3.f, inv, o.y, e.y);
    math(ru[i].z, rc[i].z, rx[i].z, r1[i].z, r2[i].z, r3[i].z, cf, 3.f, inv, o.z, e.z);
    math(ru[i].w, rc[i].w, rx[i].w, r1[i].w, r2[i].w, r3[i].w, cf, 3.f, inv, o.w, e.w);
    st(p.out + off[i], o);
    st(p.slot + off[i], e);
  }
}
__global__ void k_null(const P p) {}
__global__ void k_touch(const P p) {
  const long long off = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  st(p.out + off, ld<HINT_NONE>(p.u + off));
}

struct Set { float* t[8]; };

static float time_graph(cudaStream_t s, int iters, const std::function<void(int)>& enqueue) {
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
  for (int i = 0; i < iters; ++i) enqueue(i);
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> ts;
  for (int rep = 0; rep < 9; ++rep) {
    CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(e1, s)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms * 1e3f / iters);
  }
  std::sort(ts.begin(), ts.end());
  CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g)); CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  return ts[4];
}

int main() {
  const long long N = 4 * 64 * 64;
  cudaStream_t s; CK(cudaStreamCreate(&s));
  for (int B : {8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096}) {
    const size_t tensor_bytes = (size_t)B * N * 4, per_launch = 8 * tensor_bytes;
    int nsets = (int)std::max<size_t>(2, std::min<size_t>(48, (3ull * 126 * 1024 * 1024 + per_launch - 1) / per_launch));
    std::vector<Set> sets(nsets);
    for (auto& st_ : sets) for (int i = 0; i < 8; ++i) { CK(cudaMalloc(&st_.t[i], tensor_bytes)); CK(cudaMemset(st_.t[i], 0x3c, tensor_bytes)); }
    float* coef; CK(cudaMalloc(&coef, (size_t)B * 6 * 4)); CK(cudaMemset(coef, 0, (size_t)B * 6 * 4));
    float *ca, *cb; CK(cudaMalloc(&ca, per_launch / 2 * 3)); CK(cudaMalloc(&cb, per_launch / 2 * 3));   // 3 rotating halves
    const int iters = B >= 1024 ? 24 : 64;
    auto mkp = [&](const Set& q, int threads, int U) {
      P p; p.u = q.t[0]; p.c = q.t[1]; p.x = q.t[2]; p.h1 = q.t[3]; p.h2 = q.t[4]; p.h3 = q.t[5]; p.out = q.t[6]; p.slot = q.t[7];
      p.coef = coef; p.n_per_sample = N; p.chunks = (int)((N / 4 + threads * U - 1) / (threads * U)); return p;
    };
    const int U = (long long)B * (N / 4 / 512) >= 148 * 8 ? 2 : 1;             // the product's rule
    auto step = [&](int threads, int hint) {
      return time_graph(s, iters, [&](int i) {
        P p = mkp(sets[i % nsets], threads, U);
        const int grid = p.chunks * B;
        if (U == 2) k_step<2, HINT_NONE><<<grid, threads, 0, s>>>(p);
        else if (hint == HINT_L2_128) k_step<1, HINT_L2_128><<<grid, threads, 0, s>>>(p);
        else if (hint == HINT_L2_256) k_step<1, HINT_L2_256><<<grid, threads, 0, s>>>(p);
        else k_step<1, HINT_NONE><<<grid, threads, 0, s>>>(p);
      });
    };
    const float t_step = step(256, HINT_NONE);
    const float t_copy = time_graph(s, iters, [&](int i) {
      const size_t half = per_launch / 2;
      CK(cudaMemcpyAsync((char*)cb + (i % 3) * half, (char*)ca + (i % 3) * half, half, cudaMemcpyDeviceToDevice, s));
    });
    const int grid256 = (int)(N / 4 / 256 / U) * B;
    const float t_null = time_graph(s, iters, [&](int i) { k_null<<<grid256, 256, 0, s>>>(mkp(sets[i % nsets], 256, U)); });
    const float t_touch = B * (N / 4) >= grid256 * 256 ? time_graph(s, iters, [&](int i) {
      k_touch<<<std::min(grid256, (int)(B * N / 4 / 256)), 256, 0, s>>>(mkp(sets[i % nsets], 256, U)); }) : 0.f;
    printf("FLOOR B=%d bytes=%zu U=%d grid=%d step_us=%.3f copy_us=%.3f null_us=%.3f touch_us=%.3f\n", B, per_launch, U,
           grid256, t_step, t_copy, t_null, t_touch);
    if (B == 64 || B == 256) {
      for (int threads : {128, 256, 512, 1024})
        printf("SHAPE B=%d threads=%d hint=none step_us=%.3f\n", B, threads, step(threads, HINT_NONE));
      if (U == 1) {
        printf("SHAPE B=%d threads=256 hint=L2::128B step_us=%.3f\n", B, step(256, HINT_L2_128));
        printf("SHAPE B=%d threads=256 hint=L2::256B step_us=%.3f\n", B, step(256, HINT_L2_256));
      }
    }
    fflush(stdout);
    for (auto& q : sets) for (int i = 0; i < 8; ++i) cudaFree(q.t[i]);
    cudaFree(coef); cudaFree(ca); cudaFree(cb);
  }
  return 0;
}
