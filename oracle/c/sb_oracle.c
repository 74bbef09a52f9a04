/*
 * sb_oracle.c -- plain C restatement of the reference's CPU fake-quant path and GPTQ matvec.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): a second, independent checker next to the
 * numpy oracle, compiled by oracle/c/Makefile (and by __graft_entry__.build()).
 *
 * Follows
 *   sparsebit/quantization/quantizers/quant_tensor.py:181-184   (QDQ, per-tensor / per-channel)
 *   sparsebit/quantization/observers/minmax.py:22                (min / max)
 *   large_language_models/llama/quantization/cuda/cuda_kernel_4bit.cu:146-156 (dequant-matvec; 3 / 2-bit: utils/quant.py:210-258)
 *   sparsebit/quantization/quantizers/adaround.py:46-54         (AdaRound evaluation branch)
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

/* ---- SURVEY 8(f) rows -------------------------------------------------------------------------
 * Integer weight of input channel k from a packed column, for bit in {2, 3, 4}
 * (QuantLinear.pack, large_language_models/llama/quantization/utils/quant.py:210-258): 2 / 4-bit hold
 * 32/bit values per word; 3-bit holds 32 values in 3 words, value 10 split 2 + 1 bits over words 0 / 1 and
 * value 21 split 1 + 2 bits over words 1 / 2. */
static uint32_t sbo_unpack(const int32_t* qw, int64_t N, int64_t n, int64_t k, int bit) {
  if (bit == 2 || bit == 4) {
    const int per = 32 / bit;
    const uint32_t w = (uint32_t)qw[(k / per) * N + n];
    return (w >> (bit * (k % per))) & ((1u << bit) - 1u);
  }
  const int64_t u = k / 32;
  const int j = (int)(k % 32);
  const uint32_t w0 = (uint32_t)qw[(3 * u) * N + n];
  if (j < 10) return (w0 >> (3 * j)) & 7u;
  const uint32_t w1 = (uint32_t)qw[(3 * u + 1) * N + n];
  if (j == 10) return (w0 >> 30) | ((w1 & 1u) << 2);
  if (j < 21) return (w1 >> (3 * (j - 11) + 1)) & 7u;
  const uint32_t w2 = (uint32_t)qw[(3 * u + 2) * N + n];
  if (j == 21) return (w1 >> 31) | ((w2 & 3u) << 1);
  return (w2 >> (3 * (j - 22) + 2)) & 7u;
}

/* VecQuant{2,3,4}MatMulKernel contract (cuda_kernel_{2,3,4}bit.cu), fp64 accumulate */
void sbo_gptq_bits(const float* x, const int32_t* qw, float* out, const float* scales, const float* zeros, int64_t M,
                   int64_t K, int64_t N, int group_size, int bit) {
  const int64_t gs = group_size > 0 ? group_size : K;
  const int64_t G = (K + gs - 1) / gs;
  for (int64_t m = 0; m < M; ++m) {
    for (int64_t n = 0; n < N; ++n) {
      double acc = 0.0;
      for (int64_t k = 0; k < K; ++k) {
        const float q = (float)sbo_unpack(qw, N, n, k, bit);
        const float wv = scales[n * G + k / gs] * q - zeros[n * G + k / gs];
        acc += (double)wv * (double)x[m * K + k];
      }
      out[m * N + n] = (float)((double)out[m * N + n] + acc);
    }
  }
}

/* AdaRound evaluation branch (sparsebit/quantization/quantizers/adaround.py:46-54 with self.training == False):
 * x_q = clamp(floor(x / s) + (v >= 0) + zp, qmin, qmax); out = (x_q - zp) * s.  [outer, C, inner] view. */
void sbo_adaround_hard(const float* x, const float* v, const float* scale, const float* zero_point, float* out,
                       int64_t outer, int64_t C, int64_t inner, int qmin, int qmax) {
  for (int64_t o = 0; o < outer; ++o)
    for (int64_t c = 0; c < C; ++c) {
      const float s = scale[c], zp = zero_point[c];
      const int64_t base = (o * C + c) * inner;
      for (int64_t i = 0; i < inner; ++i) {
        const float fl = floorf(x[base + i] / s);
        const float up = v[base + i] >= 0.0f ? 1.0f : 0.0f;
        const float t = (fl + up) + zp;
        const float q = (t != t) ? t : clampf_keep_nan(t, (float)qmin, (float)qmax);
        out[base + i] = (q - zp) * s;
      }
    }
}
