// Experiment harness (not product): variants of the fused SD step (f32, CFG pair, n_hist = 4) to find what limits
// the streaming rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o step_variants step_variants.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct P { const float *u, *c, *x, *h1, *h2, *h3; float *out, *slot; const float* coef; long long n_per_sample, nvec; int chunks; };

enum { LD_NC_NOALLOC = 0, LD_PLAIN = 1, LD_EVICT_FIRST = 2 };
enum { ST_PLAIN = 0, ST_CS = 1 };

template <int LD> __device__ __forceinline__ float4 ld(const float* p) {
  float4 r;
  if (LD == LD_NC_NOALLOC)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (LD == LD_PLAIN)
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else
    asm volatile("ld.global.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
template <int ST> __device__ __forceinline__ void st(float* p, float4 v) {
  if (ST == ST_PLAIN) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool EXACT> __device__ __forceinline__ void math(float u, float c, float x, float h1, float h2, float h3,
                                                             const float* cf, float g, float& out, float& eps) {
  const float sa_t = 0.8378f, sb_t = 0.5460f, sa_p = 0.9151f, sb_p = 0.4033f;
  if (EXACT) {
    eps = __fadd_rn(u, __fmul_rn(g, __fsub_rn(c, u)));
    float eff = __fadd_rn(0.f, __fmul_rn(cf[0], eps));
    eff = __fadd_rn(eff, __fmul_rn(cf[1], h1));
    eff = __fadd_rn(eff, __fmul_rn(cf[2], h2));
    eff = __fadd_rn(eff, __fmul_rn(cf[3], h3));
    float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sb_t, eff)), sa_t);
    out = __fadd_rn(__fmul_rn(sa_p, x0), __fmul_rn(sb_p, eff));
  } else {
    eps = u + g * (c - u);
    float eff = cf[0] * eps + cf[1] * h1 + cf[2] * h2 + cf[3] * h3;
    float x0 = (x - sb_t * eff) * (1.0f / sa_t);
    out = sa_p * x0 + sb_p * eff;
  }
}

template <int U, int LD, int ST, bool EXACT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_step(const P p) {
  const int b = blockIdx.x / p.chunks, chunk = blockIdx.x - b * p.chunks;
  const long long base = (long long)b * p.n_per_sample;
  const long long v0 = (long long)chunk * (blockDim.x * U) + threadIdx.x;
  float4 ru[U], rc[U], rx[U], r1[U], r2[U], r3[U];
  long long off[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    off[i] = base + (v0 + (long long)i * blockDim.x) * 4;
    ru[i] = ld<LD>(p.u + off[i]); rc[i] = ld<LD>(p.c + off[i]); rx[i] = ld<LD>(p.x + off[i]);
    r1[i] = ld<LD>(p.h1 + off[i]); r2[i] = ld<LD>(p.h2 + off[i]); r3[i] = ld<LD>(p.h3 + off[i]);
  }
  float cf[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cf[j] = __ldg(p.coef + b * 6 + j);
#pragma unroll
  for (int i = 0; i < U; ++i) {
    float4 o, e;
    math<EXACT>(ru[i].x, rc[i].x, rx[i].x, r1[i].x, r2[i].x, r3[i].x, cf, 3.f, o.x, e.x);
    math<EXACT>(ru[i].y, rc[i].y, rx[i].y, r1[i].y, r2[i].y, r3[i].y, cf, 3.f, o.y, e.y);
    math<EXACT>(ru[i].z, rc[i].z, rx[i].z, r1[i].z, r2[i].z, r3[i].z, cf, 3.f, o.z, e.z);
    math<EXACT>(ru[i].w, rc[i].w, rx[i].w, r1[i].w, r2[i].w, r3[i].w, cf, 3.f, o.w, e.w);
    st<ST>(p.out + off[i], o);
    st<ST>(p.slot + off[i], e);
  }
}

// persistent grid-stride variant: grid = 148 * k CTAs, each thread loops over vectors of the whole [B*N] range
template <int LD, int ST, bool EXACT>
__global__ void __launch_bounds__(256) k_persist(const P p, long long total_vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total_vec; v += stride) {
    const long long off = v * 4;
    const int b = (int)(off / p.n_per_sample);
    float4 ru = ld<LD>(p.u + off), rc = ld<LD>(p.c + off), rx = ld<LD>(p.x + off), r1 = ld<LD>(p.h1 + off),
           r2 = ld<LD>(p.h2 + off), r3 = ld<LD>(p.h3 + off);
    float cf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cf[j] = __ldg(p.coef + b * 6 + j);
    float4 o, e;
    math<EXACT>(ru.x, rc.x, rx.x, r1.x, r2.x, r3.x, cf, 3.f, o.x, e.x);
    math<EXACT>(ru.y, rc.y, rx.y, r1.y, r2.y, r3.y, cf, 3.f, o.y, e.y);
    math<EXACT>(ru.z, rc.z, rx.z, r1.z, r2.z, r3.z, cf, 3.f, o.z, e.z);
    math<EXACT>(ru.w, rc.w, rx.w, r1.w, r2.w, r3.w, cf, 3.f, o.w, e.w);
    st<ST>(p.out + off, o);
    st<ST>(p.slot + off, e);
  }
}

// ---- TMA variant: 1-D bulk copies (cp.async.bulk) global -> shared, mbarrier-completed, STAGES-deep ring per CTA, persistent
// grid.  One elected thread produces; all 256 threads consume from shared memory and store with st.global.cs.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int U, int STAGES>
__global__ void __launch_bounds__(256) k_tma(const P p, long long total_tiles, int tiles_per_sample) {
  constexpr int TILE = 256 * 4 * U;                      // floats per tensor per tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw);       // [STAGES][6][TILE]
  __shared__ unsigned long long bar[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const float* src[6] = {p.u, p.c, p.x, p.h1, p.h2, p.h3};
  const long long first = blockIdx.x;
  const long long n_my = first < total_tiles ? (total_tiles - first + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](long long k) {
    const int s = (int)(k % STAGES);
    const long long off = (first + k * gridDim.x) * TILE;
    mbar_expect(&bar[s], 6u * TILE * 4u);
#pragma unroll
    for (int j = 0; j < 6; ++j) bulk_g2s(buf + ((size_t)s * 6 + j) * TILE, src[j] + off, TILE * 4u, &bar[s]);
  };
  if (threadIdx.x == 0)
    for (long long k = 0; k < STAGES && k < n_my; ++k) issue(k);
  for (long long k = 0; k < n_my; ++k) {
    const int s = (int)(k % STAGES);
    mbar_wait(&bar[s], (unsigned)((k / STAGES) & 1));
    const long long tile = first + k * gridDim.x;
    const int b = (int)(tile / tiles_per_sample);
    float cf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cf[j] = __ldg(p.coef + b * 6 + j);
    const float* sb = buf + (size_t)s * 6 * TILE;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int e = (threadIdx.x + i * 256) * 4;
      const float4 ru = *reinterpret_cast<const float4*>(sb + 0 * TILE + e), rc = *reinterpret_cast<const float4*>(sb + 1 * TILE + e),
                   rx = *reinterpret_cast<const float4*>(sb + 2 * TILE + e), r1 = *reinterpret_cast<const float4*>(sb + 3 * TILE + e),
                   r2 = *reinterpret_cast<const float4*>(sb + 4 * TILE + e), r3 = *reinterpret_cast<const float4*>(sb + 5 * TILE + e);
      float4 o, ee;
      math<true>(ru.x, rc.x, rx.x, r1.x, r2.x, r3.x, cf, 3.f, o.x, ee.x);
      math<true>(ru.y, rc.y, rx.y, r1.y, r2.y, r3.y, cf, 3.f, o.y, ee.y);
      math<true>(ru.z, rc.z, rx.z, r1.z, r2.z, r3.z, cf, 3.f, o.z, ee.z);
      math<true>(ru.w, rc.w, rx.w, r1.w, r2.w, r3.w, cf, 3.f, o.w, ee.w);
      st<ST_CS>(p.out + tile * TILE + e, o);
      st<ST_CS>(p.slot + tile * TILE + e, ee);
    }
    __syncthreads();                                     // stage s fully consumed
    if (threadIdx.x == 0 && k + STAGES < n_my) issue(k + STAGES);
  }
}

struct Set { float* t[8]; };

int main(int argc, char** argv) {
  const long long N = 4 * 64 * 64;
  std::vector<int> batches = {64, 256, 1024, 4096};
  if (argc > 1) { batches.clear(); for (int i = 1; i < argc; ++i) batches.push_back(atoi(argv[i])); }
  for (int B : batches) {
    const size_t tensor_bytes = (size_t)B * N * 4;
    const size_t per_launch = 8 * tensor_bytes;
    int nsets = (int)std::max<size_t>(2, std::min<size_t>(24, (3ull * 126 * 1024 * 1024 + per_launch - 1) / per_launch));
    std::vector<Set> sets(nsets);
    for (auto& s : sets) for (int i = 0; i < 8; ++i) { CK(cudaMalloc(&s.t[i], tensor_bytes)); CK(cudaMemset(s.t[i], 0x3c, tensor_bytes)); }
    float* coef; CK(cudaMalloc(&coef, (size_t)B * 6 * 4)); CK(cudaMemset(coef, 0, (size_t)B * 6 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
      for (int i = 0; i < nsets + 2; ++i) launch(sets[i % nsets]);
      CK(cudaDeviceSynchronize());
      std::vector<float> ts;
      const int iters = B >= 1024 ? 20 : 60;
      for (int rep = 0; rep < 7; ++rep) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < iters; ++i) launch(sets[i % nsets]);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ts.push_back(ms * 1e3f / iters);
      }
      std::sort(ts.begin(), ts.end());
      printf("B=%-5d %-44s %8.2f us  %7.1f GB/s\n", B, name, ts[3], per_launch / ts[3] / 1e3);
      CK(cudaGetLastError());
    };
    auto mkp = [&](const Set& s, int U) {
      P p; p.u = s.t[0]; p.c = s.t[1]; p.x = s.t[2]; p.h1 = s.t[3]; p.h2 = s.t[4]; p.h3 = s.t[5]; p.out = s.t[6]; p.slot = s.t[7];
      p.coef = coef; p.n_per_sample = N; p.nvec = N / 4; p.chunks = (int)((N / 4 + 256 * U - 1) / (256 * U)); return p;
    };
#define RUN_STEP(NAME, U, LD, ST, EX, MINB) run(NAME, [&](const Set& s) { P p = mkp(s, U); k_step<U, LD, ST, EX, MINB><<<p.chunks * B, 256>>>(p); })
    RUN_STEP("U2 nc.noalloc st.plain exact (product)", 2, LD_NC_NOALLOC, ST_PLAIN, true, 1);
    RUN_STEP("U1 nc.noalloc st.plain exact", 1, LD_NC_NOALLOC, ST_PLAIN, true, 1);
    RUN_STEP("U4 nc.noalloc st.plain exact", 4, LD_NC_NOALLOC, ST_PLAIN, true, 1);
    RUN_STEP("U2 plain-ld st.plain exact", 2, LD_PLAIN, ST_PLAIN, true, 1);
    RUN_STEP("U2 evict_first-ld st.plain exact", 2, LD_EVICT_FIRST, ST_PLAIN, true, 1);
    RUN_STEP("U2 nc.noalloc st.cs exact", 2, LD_NC_NOALLOC, ST_CS, true, 1);
    RUN_STEP("U2 nc.noalloc st.plain FAST-math", 2, LD_NC_NOALLOC, ST_PLAIN, false, 1);
    RUN_STEP("U2 plain-ld st.cs FAST-math", 2, LD_PLAIN, ST_CS, false, 1);
    RUN_STEP("U1 nc.noalloc st.plain exact minb6", 1, LD_NC_NOALLOC, ST_PLAIN, true, 6);
    RUN_STEP("U2 nc.noalloc st.plain exact minb5", 2, LD_NC_NOALLOC, ST_PLAIN, true, 5);
    for (int k : {2, 4, 8}) {
      char nm[64]; snprintf(nm, sizeof nm, "persistent grid=148x%d exact", k);
      run(nm, [&](const Set& s) { P p = mkp(s, 1); k_persist<LD_NC_NOALLOC, ST_PLAIN, true><<<148 * k, 256>>>(p, (long long)B * N / 4); });
    }
#define RUN_TMA(NAME, U, STG, CTAS_PER_SM) do { \
      const int smem = STG * 6 * 256 * 4 * U * 4; \
      CK(cudaFuncSetAttribute(k_tma<U, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      const long long tiles = (long long)B * N / (256 * 4 * U); \
      const int grid = (int)std::min<long long>(tiles, 148LL * CTAS_PER_SM); \
      run(NAME, [&](const Set& s) { P p = mkp(s, U); k_tma<U, STG><<<grid, 256, smem>>>(p, tiles, (int)(N / (256 * 4 * U))); }); \
    } while (0)
    RUN_TMA("TMA bulk U1 (4 KiB/tensor) 4 stages, 2 CTA/SM st.cs", 1, 4, 2);
    RUN_TMA("TMA bulk U2 (8 KiB/tensor) 2 stages, 2 CTA/SM st.cs", 2, 2, 2);
    RUN_TMA("TMA bulk U2 (8 KiB/tensor) 4 stages, 1 CTA/SM st.cs", 2, 4, 1);
    RUN_TMA("TMA bulk U1 (4 KiB/tensor) 2 stages, 4 CTA/SM st.cs", 1, 2, 4);
    for (auto& s : sets) for (int i = 0; i < 8; ++i) cudaFree(s.t[i]);
    cudaFree(coef);
  }
  return 0;
}
