// step_common.cuh — shared pieces of the fused solver-step kernels (SD and FM).
//
// Streaming design (HBM-bound, ~1.5 flop/byte): every latent-sized operand is read exactly once with
// 128-bit non-coherent loads that bypass L1 allocation, all loads of a thread are issued before the first
// use (memory-level parallelism = UNROLL x number of streams), math is fp32 with explicitly rounded
// intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn are never contracted into FMAs, so the fp32 path is
// bit-identical to the reference's op-by-op torch arithmetic), results are written once with 128-bit stores.
// A CTA never straddles two samples, so the per-sample coefficients are CTA-uniform broadcast loads.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/consolver.h"

namespace consolver {

constexpr int kMaxOlder = CONSOLVER_MAX_ORDER - 1;

struct StepParams {
  const void* e0;
  const void* cond;      // non-null: CFG pair, e0 is the unconditional half
  void* slot_out;        // nullable
  const void* hist[kMaxOlder];
  const void* x;
  void* x_out;
  void* x_out2;           // nullable: second copy of x' (e.g. the other half of the next CFG-doubled denoiser input)
  long long out2_stride;  // elements between consecutive samples in x_out2
  const float* coef;
  int coef_stride;
  int order_dim;
  int n_hist;
  int flags;
  float guidance;
  float k0, k1, k2, k3;  // SD: sqrt(abar_t), sqrt(1-abar_t), sqrt(abar_prev), sqrt(1-abar_prev); FM: k0 = dt
  long long e_stride;       // FM: elements between consecutive samples of e0 / hist (>= n_per_sample)
  long long n_per_sample;   // elements
  long long nvec_per_sample;  // thread-vectors per sample (n_per_sample / ELEMS)
  int chunks_per_sample;
  int B;
};

// ---- element traits -------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kPerVec = 4;
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static constexpr bool k16 = false;
};
template <> struct Elem<__half> {
  static constexpr int kPerVec = 8;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static constexpr bool k16 = true;
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kPerVec = 8;
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static constexpr bool k16 = true;
};

// v rounded to T and back: the value a torch op with a T result would hold
template <typename T> __device__ __forceinline__ float round_to(float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); }

// ---- 128-bit streaming accessors ------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st16(void* p, const uint4& v) {
  // streaming (evict-first) stores: measured +4.3 % over default-policy stores on B200 for this 6-read / 2-write
  // mix (profiles/step_variants_r01.txt): 7246 vs 6948 GB/s at B=4096
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- loads of data the PDL PRIMARY produces (the coefficient record; under CHAIN the latent and the newest slot) ----
// They sit after griddepcontrol.wait and must stay there.  `__ldg` / ld.global.nc declare the memory read-only for the
// kernel's lifetime, which (a) is not true of such data — the primary is still writing it when this grid starts — and
// (b) lets nvcc treat the load as invariant and HOIST IT ABOVE THE WAIT: seen in SASS as LDG.E.CONSTANT of the
// coefficient record before ACQBULK in 12 instantiations (16-bit depth-1 and the runtime-depth forms), where the live
// differential fuzzing caught stale coefficients (profiles/live_fuzz_r02.md).  These are ordinary (weak, coherent-path)
// global loads — what the PDL contract covers: after the wait, the prerequisite grids' writes are visible to them —
// written as volatile asm, which the compiler keeps in order with the wait.  (.cg / STRONG.GPU loads are also correct
// but cost 1.8 us per launch at batch 64: every CTA's coefficient read then goes to the L2 point of coherence.)
__device__ __forceinline__ float ld_produced_f32(const float* p) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_produced16(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

// A thread-vector of E elements of type T held as raw 16-byte words (E*sizeof(T) is 16 or 32 bytes).
template <typename T, int E> struct Raw {
  static constexpr int kWords = (E * (int)sizeof(T)) / 16;
  uint4 w[kWords];
  __device__ __forceinline__ void load(const T* p) {
#pragma unroll
    for (int i = 0; i < kWords; ++i) w[i] = ld_stream16(reinterpret_cast<const char*>(p) + 16 * i);
  }
  // data written by the PDL primary (see ld_produced16)
  __device__ __forceinline__ void load_produced(const T* p) {
#pragma unroll
    for (int i = 0; i < kWords; ++i) w[i] = ld_produced16(reinterpret_cast<const char*>(p) + 16 * i);
  }
  __device__ __forceinline__ void store(T* p) const {
#pragma unroll
    for (int i = 0; i < kWords; ++i) st16(reinterpret_cast<char*>(p) + 16 * i, w[i]);
  }
  __device__ __forceinline__ float get(int i) const { return Elem<T>::to_f(reinterpret_cast<const T*>(w)[i]); }
  __device__ __forceinline__ void set(int i, float v) { reinterpret_cast<T*>(w)[i] = Elem<T>::from_f(v); }
};
// scalar fallback (E == 1): ragged sizes / unaligned pointers
template <typename T> struct Raw<T, 1> {
  T v;
  __device__ __forceinline__ void load(const T* p) { v = __ldg(p); }
  __device__ __forceinline__ void load_produced(const T* p) {
    if constexpr (sizeof(T) == 4) {
      unsigned r;
      asm volatile("ld.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
      v = *reinterpret_cast<const T*>(&r);
    } else {
      unsigned short r;
      asm volatile("ld.global.u16 %0, [%1];" : "=h"(r) : "l"(p) : "memory");
      v = *reinterpret_cast<const T*>(&r);
    }
  }
  __device__ __forceinline__ void store(T* p) const { *p = v; }
  __device__ __forceinline__ float get(int) const { return Elem<T>::to_f(v); }
  __device__ __forceinline__ void set(int, float f) { v = Elem<T>::from_f(f); }
};

// process-global launch tuning (consolver_set_step_launch)
struct StepLaunchCfg {
  int threads;
  int unroll;
};
StepLaunchCfg step_launch_cfg();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// SM count of the CURRENT device (148 on B200), queried once per device; grid heuristics are multiples of it.
inline int sm_count() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace consolver
