/*
 * sb_oracle.c -- plain C restatement of the reference's CPU fake-quant path and GPTQ matvec.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a second, independent checker next to the
 * numpy oracle, compiled by oracle/c/Makefile (and by __graft_entry__.build()).
 *
 * Follows
 *   sparsebit/quantization/quantizers/quant_tensor.py:181-184   (QDQ, per-tensor / per-channel)
 *   sparsebit/quantization/observers/minmax.py:22                (min / max)
 *   large_language_models/llama/quantization/cuda/cuda_kernel_4bit.cu:146-156 (dequant-matvec)
 * Build with -ffp-contract=off so the float ops stay separate IEEE operations like ATen's.
 */
#include <math.h>
#include <stdint.h>

static float clampf_keep_nan(float v, float lo, float hi) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

/* x viewed as [outer, C, inner]; C == 1 is the per-tensor case. */
void sbo_qdq(const float* x, const float* scale, const float* zero_point, float* out, int64_t outer,
             int64_t channels, int64_t inner, int qmin, int qmax) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t o = 0; o < outer; ++o) {
    for (int64_t c = 0; c < channels; ++c) {
      const float s = scale[c];
      const float zp = nearbyintf(zero_point[c]);
      const float* xr = x + (o * channels + c) * inner;
      float* yr = out + (o * channels + c) * inner;
      for (int64_t i = 0; i < inner; ++i) {
        float q = nearbyintf(xr[i] / s) + zp;
        q = clampf_keep_nan(q, (float)qmin, (float)qmax);
        yr[i] = (q - zp) * s;
      }
    }
  }
}

void sbo_minmax(const float* x, int64_t n, float* out_min, float* out_max) {
  float lo = INFINITY, hi = -INFINITY;
#pragma omp parallel for reduction(min : lo) reduction(max : hi)
  for (int64_t i = 0; i < n; ++i) {
    if (x[i] < lo) lo = x[i];
    if (x[i] > hi) hi = x[i];
  }
  *out_min = lo;
  *out_max = hi;
}

/* out[m, n] += sum_k (scales[n*G + k/gs] * nib(qw[k/8, n], k%8) - zeros[n*G + k/gs]) * x[m, k], fp64 accumulate */
void sbo_gptq4(const float* x, const int32_t* qw, float* out, const float* scales, const float* zeros, int64_t M,
               int64_t K, int64_t N, int group_size) {
  const int64_t gs = group_size > 0 ? group_size : K;
  const int64_t G = (K + gs - 1) / gs;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t m = 0; m < M; ++m) {
    for (int64_t n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int64_t k = 0; k < K; ++k) {
        const uint32_t w = (uint32_t)qw[(k >> 3) * N + n];
        const float q = (float)((w >> (4 * (k & 7))) & 0xF);
        const float wv = scales[n * G + k / gs] * q - zeros[n * G + k / gs];
        acc += (double)wv * (double)x[m * K + k];
      }
      out[m * N + n] = (float)((double)out[m * N + n] + acc);
    }
  }
}
