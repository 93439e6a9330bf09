/* A plain-C host of the C ABI (include/consolver.h): no Python, no torch — the shape of a binding from any compiled
 * host language.  Runs a 4-step SD solver loop with a CFG pair through consolver_step_sd with fixed
 * coefficients on the GPU, the same loop through the plain-C oracle (oracle/consolver_oracle.c) on the CPU, and
 * requires the latents to be bit-identical.  Test infrastructure: built and run by tests/test_gpu_c_host.py.
 *
 *   gcc -std=c99 abi_host.c -I<repo>/include -I$CUDA/include -L<repo>/consolver_b200 -lconsolver \
 *       -L<repo>/oracle/_build -loracle -L$CUDA/lib64 -lcudart -lm
 */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "consolver.h"

void oracle_sd_step_f32(const float* const* hist, int n_hist, const float* x, float* x_out, const float* coef,
                        int order_dim, float sa_t, float sb_t, float sa_p, float sb_p, int flags, int B, int64_t N);
void oracle_cfg_f32(const float* u, const float* c, float g, float* out, int64_t n);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

static uint32_t rng_state = 12345u;
static float frand(void) { /* xorshift, uniform in (-2, 2) */
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5;
  return ((float)(rng_state >> 8) / 8388608.0f - 1.0f) * 2.0f;
}

int main(void) {
  enum { B = 3, OD = 4, STEPS = 4 };
  const int64_t N = 4 * 16 * 16 + 3;               /* ragged on purpose: scalar path */
  const size_t n = (size_t)B * N, bytes = n * sizeof(float);
  const float guidance = 3.0f;
  const float sc[STEPS][4] = {{0.068f, 0.9977f, 0.31f, 0.9507f}, {0.31f, 0.9507f, 0.55f, 0.8352f},
                              {0.55f, 0.8352f, 0.78f, 0.6258f}, {0.78f, 0.6258f, 0.9996f, 0.0292f}};
  if (consolver_abi_version() != CONSOLVER_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }

  float* h_x = malloc(bytes); float* h_ref = malloc(bytes); float* h_got = malloc(bytes);
  float* h_pair[STEPS]; float* h_eps[STEPS];
  float h_coef[B * (OD + 2)];
  for (size_t i = 0; i < n; ++i) h_x[i] = frand();
  for (int s = 0; s < STEPS; ++s) {
    h_pair[s] = malloc(2 * bytes); h_eps[s] = malloc(bytes);
    for (size_t i = 0; i < 2 * n; ++i) h_pair[s][i] = frand();
  }
  float *d_x, *d_xn, *d_pair[STEPS], *d_ring, *d_coef;
  CK(cudaMalloc((void**)&d_x, bytes)); CK(cudaMalloc((void**)&d_xn, bytes));
  CK(cudaMalloc((void**)&d_ring, OD * bytes)); CK(cudaMalloc((void**)&d_coef, sizeof h_coef));
  CK(cudaMemcpy(d_x, h_x, bytes, cudaMemcpyHostToDevice));
  for (int s = 0; s < STEPS; ++s) {
    CK(cudaMalloc((void**)&d_pair[s], 2 * bytes));
    CK(cudaMemcpy(d_pair[s], h_pair[s], 2 * bytes, cudaMemcpyHostToDevice));
  }
  cudaStream_t stream; CK(cudaStreamCreate(&stream));
  memcpy(h_ref, h_x, bytes);

  for (int s = 0; s < STEPS; ++s) {
    const int n_hist = s + 1 < OD ? s + 1 : OD;
    /* per-sample multipliers (newest first) that sum to one, then the two scalers */
    for (int b = 0; b < B; ++b) {
      float* c = h_coef + b * (OD + 2);
      float rest = 1.0f;
      for (int j = 0; j < OD; ++j) c[j] = 0.0f;
      for (int j = 1; j < n_hist; ++j) { c[j] = 0.1f * (float)(j + b) - 0.25f; rest -= c[j]; }
      c[0] = rest; c[OD] = 1.0f; c[OD + 1] = 1.0f;
    }
    CK(cudaMemcpyAsync(d_coef, h_coef, sizeof h_coef, cudaMemcpyHostToDevice, stream));
    /* GPU: the pair goes in raw; eps is written into ring slot s % OD; older slots are read by pointer */
    const void* hist[OD];
    for (int j = 1; j < n_hist; ++j) hist[j - 1] = d_ring + (size_t)((s - j) % OD) * n;
    /* odd steps with CONSOLVER_FLAG_HOST_SCALARS (the reference on CPU tensors), even steps with the default rules
     * (the reference on CUDA tensors): the oracle gets the same flag */
    const int sem = (s & 1) ? CONSOLVER_FLAG_HOST_SCALARS : 0;
    int rc = consolver_step_sd(CONSOLVER_F32, d_pair[s], d_pair[s] + n, guidance, d_ring + (size_t)(s % OD) * n, hist,
                               n_hist, d_x, d_xn, NULL, 0, d_coef, CONSOLVER_COEF_STRIDE(OD), OD, sc[s][0], sc[s][1],
                               sc[s][2], sc[s][3], sem, B, N, stream);
    if (rc != 0) { fprintf(stderr, "consolver_step_sd: %s\n", consolver_error_string(rc)); return 3; }
    { float* t = d_x; d_x = d_xn; d_xn = t; }
    /* CPU oracle: the reference's sequence — CFG combine, then the step on the newest-first history */
    oracle_cfg_f32(h_pair[s], h_pair[s] + n, guidance, h_eps[s], (int64_t)n);
    const float* ohist[OD];
    for (int j = 0; j < n_hist; ++j) ohist[j] = h_eps[s - j];
    oracle_sd_step_f32(ohist, n_hist, h_ref, h_got, h_coef, OD, sc[s][0], sc[s][1], sc[s][2], sc[s][3], sem, B, N);
    memcpy(h_ref, h_got, bytes);
    CK(cudaStreamSynchronize(stream));
    CK(cudaMemcpy(h_got, d_x, bytes, cudaMemcpyDeviceToHost));
    if (memcmp(h_got, h_ref, bytes) != 0) { fprintf(stderr, "step %d: latents differ from the oracle\n", s); return 4; }
  }
  /* argument errors come back as codes, never as crashes */
  if (consolver_step_sd(CONSOLVER_F32, NULL, NULL, 0.f, NULL, NULL, 1, d_x, d_xn, NULL, 0, d_coef, OD + 2, OD, 1.f, 0.f,
                        1.f, 0.f, 0, B, N, stream) != CONSOLVER_ERR_NULL) return 5;
  if (consolver_step_sd(9, d_x, NULL, 0.f, NULL, NULL, 1, d_x, d_xn, NULL, 0, d_coef, OD + 2, OD, 1.f, 0.f, 1.f, 0.f, 0,
                        B, N, stream) != CONSOLVER_ERR_DTYPE) return 6;
  printf("c host ok: %d steps, B=%d, N=%lld, latents bit-identical to the plain-C oracle\n", STEPS, B, (long long)N);
  return 0;
}
