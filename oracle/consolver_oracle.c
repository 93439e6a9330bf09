/*
 * TEST INFRASTRUCTURE ONLY — plain-C restatement of the per-element arithmetic and the index work of the
 * ConsistencySolver step, independent of torch.  It exists to cross-check the torch-CPU oracle
 * (oracle/consolver_oracle.py) and the CUDA kernels; the product never links or calls it.
 * Compile with -ffp-contract=off: the reference rounds after every fp32 operation (one torch op each).
 *
 * Parity pin: tests/test_oracle_c.py runs these functions on the golden vectors produced by the unmodified
 * reference (tests/golden, .npz files) and demands bit-identical latents / indices.
 *
 * Reference lines (relative to the reference root):
 *   CFG combine                 denoise_ppo.py:96-100
 *   multistep combine           scheduler_ppo.py:263-280
 *   DDIM update                 scheduler_ppo.py:306-332
 *   FM Euler update             edit_ppo/scheduler_fmppo.py:354,:413-436
 *   categorical draw            factor_net_ppo.py:161  (torch.multinomial == argmax(p / q), q ~ Exp(1))
 *   masks / coefficients        scheduler_ppo.py:248-259, :165-175
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* history: hist[0] is the newest model output (already CFG-combined), hist[j] older; coef: [B, order_dim+2] with
 * the layout of include/consolver.h.  flags: 1 v-prediction, 2 eff scale, 4 x scale, 256 (the value of
 * CONSOLVER_FLAG_HOST_SCALARS) the reference executed on CPU tensors: `t / sqrt(abar_t)` is a true division.
 * Without it the reference executed on CUDA tensors: ATen computes t * (1 / scalar), the reciprocal taken once in
 * fp32 (pinned by tests/golden/cuda_sd_*). */
void oracle_sd_step_f32(const float* const* hist, int n_hist, const float* x, float* x_out, const float* coef,
                        int order_dim, float sa_t, float sb_t, float sa_p, float sb_p, int flags, int B, int64_t N) {
  const int stride = order_dim + 2;
  for (int b = 0; b < B; ++b) {
    const float* c = coef + (size_t)b * stride;
    for (int64_t i = 0; i < N; ++i) {
      const size_t o = (size_t)b * N + i;
      float eff;
      if (n_hist == 1) {
        eff = hist[0][o];
      } else {
        eff = 0.0f;
        for (int j = 0; j < n_hist; ++j) {
          float m = c[j] * hist[j][o];
          eff = eff + m;
        }
      }
      if (flags & 2) eff = eff * c[order_dim];
      float xs = x[o];
      if (flags & 4) xs = xs * c[order_dim + 1];
      if (flags & 1) {
        float a = sa_t * eff, bb = sb_t * xs;
        eff = a + bb;
      }
      float t0 = sb_t * eff;
      float t1 = xs - t0;
      float inv = 1.0f / sa_t;
      float x0 = (flags & 256) ? t1 / sa_t : t1 * inv;
      float p0 = sa_p * x0, p1 = sb_p * eff;
      x_out[o] = p0 + p1;
    }
  }
}

void oracle_cfg_f32(const float* u, const float* c, float g, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    float d = c[i] - u[i];
    float m = g * d;
    out[i] = u[i] + m;
  }
}

/* fp32 flow-matching step (16-bit I/O is covered by the torch oracle only) */
void oracle_fm_step_f32(const float* const* hist, int n_hist, const float* x, float* x_out, const float* coef,
                        int order_dim, float dt, int flags, int B, int64_t N) {
  const int stride = order_dim + 2;
  for (int b = 0; b < B; ++b) {
    const float* c = coef + (size_t)b * stride;
    for (int64_t i = 0; i < N; ++i) {
      const size_t o = (size_t)b * N + i;
      float eff;
      if (n_hist == 1) {
        eff = hist[0][o];
      } else {
        eff = 0.0f;
        for (int j = 0; j < n_hist; ++j) {
          float m = c[j] * hist[j][o];
          eff = eff + m;
        }
      }
      if (flags & 2) eff = eff * c[order_dim];
      float xs = x[o];
      if (flags & 4) xs = xs * c[order_dim + 1];
      float pr = dt * eff;
      x_out[o] = xs + pr;
    }
  }
}

/* torch.sum(torch.stack(terms), dim=0) of set_default_coefficients (scheduler_ppo.py:172) in the order ATen adds:
 * sum_mode 0: left to right (CPU tensors); 1: CUDA tensors, B >= 2 — four accumulators acc[i%4] += s_i, then
 * ((acc0+acc1)+acc2)+acc3; 2: CUDA tensors, B == 1 — last_pow2(m) threads, each with four accumulators over its strided
 * terms, then a shuffle tree with decreasing offsets (ATen/native/cuda/Reduce.cuh; three terms: (s0+s2)+s1). */
static float oracle_sum_terms(const float* s, int m, int sum_mode) {
  if (m <= 0) return 0.0f;
  if (sum_mode == 0) {
    float run = s[0];
    for (int i = 1; i < m; ++i) run = run + s[i];
    return run;
  }
  if (sum_mode == 1) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = 0; i < m; ++i) acc[i & 3] = acc[i & 3] + s[i];
    float r = acc[0] + acc[1];
    r = r + acc[2];
    return r + acc[3];
  }
  int W = 1;
  while (W * 2 <= m) W *= 2;
  float t[8] = {0};
  for (int x = 0; x < W; ++x) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int k = 0; k < 4 && x + k * W < m; ++k) acc[k] = acc[k] + s[x + k * W];
    float r = acc[0] + acc[1];
    r = r + acc[2];
    t[x] = r + acc[3];
  }
  for (int off = W / 2; off > 0; off /= 2)
    for (int x = 0; x < off; ++x) t[x] = t[x] + t[x + off];
  return t[0];
}

/* draw + gather + masks + coefficients from a probability table probs [A,K] and q [B*A,K]; sum_mode: see above */
void oracle_policy_sample(const float* probs, const float* action_values, const float* q, int B, int A, int K,
                          int order_dim, int scaler_dim, int n_hist, int sum_mode, int64_t* idx, float* actions,
                          float* act_probs, float* masks, float* coef) {
  for (int b = 0; b < B; ++b) {
    float act[64];
    for (int a = 0; a < A; ++a) {
      const float* qr = q + ((size_t)b * A + a) * K;
      int best = 0;
      float bv = 0.0f;
      for (int k = 0; k < K; ++k) {
        float r = probs[a * K + k] / qr[k];
        if (k == 0 || r > bv) { bv = r; best = k; }
      }
      const size_t o = (size_t)b * A + a;
      idx[o] = best;
      actions[o] = action_values[a * K + best];
      act_probs[o] = probs[a * K + best];
      masks[o] = (a >= n_hist - 1 && a < order_dim - 1) ? 0.0f : 1.0f;
      if (a < 64) act[a] = actions[o];
    }
    float* c = coef + (size_t)b * (order_dim + 2);
    float terms[64];
    terms[0] = act[0] + 1.0f;
    for (int i = 1; i < n_hist - 1; ++i) terms[i] = act[i];
    for (int i = 0; i < order_dim; ++i) {
      float v = 0.0f;
      if (n_hist == 1) v = (i == 0) ? 1.0f : 0.0f;
      else if (i < n_hist - 1) v = terms[i];
      else if (i == n_hist - 1) v = 1.0f - oracle_sum_terms(terms, n_hist - 1, sum_mode);
      c[i] = v;
    }
    c[order_dim] = scaler_dim >= 1 ? act[order_dim - 1] + 1.0f : 1.0f;
    c[order_dim + 1] = scaler_dim >= 2 ? act[order_dim] + 1.0f : 1.0f;
  }
}

/* MLP + softmax for one row (plain left-to-right fp32 sums; BLAS orders differ by ulps) */
void oracle_policy_table(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                         const float* b3, float x0, float x1, float x_div, float temp, int H, int A, int K,
                         float* h1, float* h2, float* probs) {
  float xin[2] = {x0 / x_div, x1 / x_div};
  for (int j = 0; j < H; ++j) {
    float s = b1[j];
    for (int i = 0; i < 2; ++i) s += w1[j * 2 + i] * xin[i];
    h1[j] = s > 0.0f ? s : 0.0f;
  }
  for (int j = 0; j < H; ++j) {
    double s = b2[j];
    for (int i = 0; i < H; ++i) s += (double)w2[(size_t)j * H + i] * h1[i];
    h2[j] = s > 0.0 ? (float)s : 0.0f;
  }
  for (int a = 0; a < A; ++a) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const int r = a * K + k;
      double s = b3[r];
      for (int i = 0; i < H; ++i) s += (double)w3[(size_t)r * H + i] * h2[i];
      probs[r] = (float)s / temp;
      if (probs[r] > m) m = probs[r];
    }
    double z = 0.0;
    for (int k = 0; k < K; ++k) { probs[a * K + k] = expf(probs[a * K + k] - m); z += probs[a * K + k]; }
    for (int k = 0; k < K; ++k) probs[a * K + k] = probs[a * K + k] / (float)z;
  }
}
