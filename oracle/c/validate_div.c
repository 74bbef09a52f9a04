/*
 * validate_div.c -- CPU re-enactment of the two division schemes the QDQ kernels use instead of
 * `div.rn.f32` (sparsebit_b200/csrc/common.cuh: div_exact, quant_round).  TEST INFRASTRUCTURE ONLY.
 *
 *   (A) RN32(RN64(x * RN64(1/s)))  ==  RN32(x / s)                       for every finite x, s != 0
 *   (B) fast path: q0 = x*r, e = fma(-q0, s, x), q1 = fma(e, r, q0), t = (q1 + 1.5*2^23) - 1.5*2^23;
 *       accepted iff 0.5 - |q1 - t| > |q1| * 2^-22 (and |s| in [2^-60, 2^60]); then t == rint(x / s).
 *
 * x86-64 SSE float / double arithmetic and fmaf are the same correctly rounded IEEE operations as the PTX
 * .rn instructions, so a mismatch here is a mismatch on the GPU.  Inputs: uniformly random bit patterns,
 * "typical" scales with activations-like x, and adversarial x = (k + 0.5) * s nudged by a few ulps.
 *
 * usage: validate_div [cases_per_family (default 2e7)] [seed]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t rng_state;
static uint64_t rng(void) {
  uint64_t x = rng_state;
  x ^= x << 13;
  x ^= x >> 7;
  x ^= x << 17;
  return rng_state = x;
}
static float bits_to_f(uint32_t b) {
  float f;
  memcpy(&f, &b, 4);
  return f;
}
static uint32_t f_to_bits(float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  return b;
}
static int same_value(float a, float b) { return (a != a && b != b) || a == b; }  /* NaN == NaN, +0 == -0 */

static long long bad_a = 0, bad_b = 0, accepted = 0, tested_b = 0, total = 0;

static void check(float x, float s) {
  if (s == 0.0f || s != s || isinf(s) || x != x || isinf(x)) return;
  ++total;
  const float want = x / s;
  const double rs = 1.0 / (double)s;
  const float got = (float)((double)x * rs);
  if (!same_value(got, want) || (got == want && got != 0.0f && f_to_bits(got) != f_to_bits(want))) {
    if (++bad_a <= 5) fprintf(stderr, "A mismatch: x=%a s=%a exact=%a trick=%a\n", x, s, want, got);
  }
  const float a = fabsf(s);
  if (!(a >= 0x1p-60f && a <= 0x1p60f)) return;
  ++tested_b;
  const float r = (float)rs;
  const float q0 = x * r;
  const float e = fmaf(-q0, s, x);
  const float q1 = fmaf(e, r, q0);
  const float t1 = (q1 + 12582912.0f) + -12582912.0f;
  const float d = q1 - t1;
  const float c = fmaf(fabsf(q1), -0x1p-22f, 0.5f - fabsf(d));
  if (c > 0.0f) {
    ++accepted;
    if (!same_value(t1, rintf(want))) {
      if (++bad_b <= 5) fprintf(stderr, "B mismatch: x=%a s=%a rint(x/s)=%a fast=%a (q1=%a)\n", x, s, rintf(want), t1, q1);
    }
  }
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 20000000LL;
  rng_state = argc > 2 ? strtoull(argv[2], NULL, 10) : 0x9E3779B97F4A7C15ull;
  if (!rng_state) rng_state = 1;
  /* family 1: arbitrary bit patterns */
  for (long long i = 0; i < n; ++i) {
    const uint64_t v = rng();
    check(bits_to_f((uint32_t)v), bits_to_f((uint32_t)(v >> 32)));
  }
  /* family 2: quantizer-like scales (2^-20 .. 2^4), activations within +-300 quantisation steps */
  for (long long i = 0; i < n; ++i) {
    const uint64_t v = rng();
    const float s = ldexpf(1.0f + (float)(v & 0x7FFFFF) * 0x1p-23f, (int)((v >> 23) % 25) - 20);
    const float k = (float)((int)((v >> 32) % 601) - 300) + (float)((v >> 44) & 0xFFFFF) * 0x1p-20f;
    check(k * s, ((v >> 63) & 1) ? -s : s);
  }
  /* family 3: adversarial half-way points (k + 0.5) * s moved by -3..3 ulps */
  for (long long i = 0; i < n; ++i) {
    const uint64_t v = rng();
    const float s = ldexpf(1.0f + (float)(v & 0x7FFFFF) * 0x1p-23f, (int)((v >> 23) % 41) - 30);
    const float k = (float)((int)((v >> 32) % 65537) - 32768) + 0.5f;
    float x = k * s;
    int nudge = (int)((v >> 50) % 7) - 3;
    while (nudge > 0) { x = nextafterf(x, INFINITY); --nudge; }
    while (nudge < 0) { x = nextafterf(x, -INFINITY); ++nudge; }
    check(x, s);
  }
  printf("{\"cases\": %lld, \"exact_trick_mismatches\": %lld, \"fast_path_tested\": %lld, \"fast_path_accepted\": %lld, "
         "\"fast_path_mismatches\": %lld}\n",
         total, bad_a, tested_b, accepted, bad_b);
  return (bad_a || bad_b) ? 1 : 0;
}
