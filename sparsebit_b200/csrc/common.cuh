// common.cuh -- shared device/host helpers for the sm_100a kernels behind include/sparsebit_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/sparsebit_b200.h"

namespace sb200 {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
int sm_count();
extern std::atomic<long long> g_launches;
extern int g_variant;

#define SB_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      sb200::set_error(__VA_ARGS__);   \
      return SB200_E_INVALID;          \
    }                                  \
  } while (0)

#define SB_CUDA(call)                                                                    \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      sb200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                        \
      return SB200_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define SB_LAUNCHED()                                                                     \
  do {                                                                                    \
    sb200::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t e__ = cudaPeekAtLastError();                                              \
    if (e__ != cudaSuccess) {                                                             \
      (void)cudaGetLastError();                                                           \
      sb200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                         \
      return SB200_E_CUDA;                                                                \
    }                                                                                     \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------- order-preserving float key
// enc() maps fp32 to uint32 such that a < b  <=>  enc(a) < enc(b) (with -0 < +0, NaN at the ends).
__host__ __device__ __forceinline__ uint32_t enc_f32(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_f32(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
// Running min/max state: {enc(min), enc(max)} ; "empty" = {0xFFFFFFFF, 0}.  A NaN input poisons
// both slots with the canonical NaN marker below (torch.min / torch.max propagate NaN).
#define SB_MM_EMPTY_MIN 0xFFFFFFFFu
#define SB_MM_EMPTY_MAX 0x00000000u

#ifdef __CUDACC__
// ---------------------------------------------------------------- streaming 128-bit access
__device__ __forceinline__ float4 ld_stream4(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream4(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ float ld_stream1(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream1(float* p, float v) { __stcs(p, v); }

// ---------------------------------------------------------------- QDQ scalar math (oracle order)
// Follows sparsebit/quantization/quantizers/quant_tensor.py:181-184 exactly:
//   zp = round(zero_point); x_q = clamp(round(x / scale) + zp, qmin, qmax); x_dq = (x_q - zp) * scale
// with IEEE semantics for every step, no FMA contraction, NaN propagating through the clamp
// (torch.clamp semantics).
//
// Division.  The IEEE quotient RN32(x/s) is obtained as RN32(RN64(x * RN64(1/s))): the quotient of
// two binary32 numbers is at least 2^-49 (relative) away from every binary32 rounding midpoint
// (|X*2^24 - M*S| >= 1 for 24-bit X, S and a 25-bit odd M), while the double-precision product is
// within 2^-52 of it, so both round to the same float for every input including zero, subnormal,
// infinite and NaN operands.  This is 3 instructions per element (cvt, DMUL, cvt) with the
// reciprocal hoisted out of the loop, instead of the ~35-instruction div.rn.f32 expansion (MUFU +
// Newton + range check + slow-path call) per element.
__device__ __forceinline__ float fmin_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

struct QP {
  float s, zp, qmin, qmax;
  double rs;  // RN64(1 / s)
  __device__ __forceinline__ void set(float scale, float zero_point_raw) {
    s = scale;
    zp = rintf(zero_point_raw);
    rs = __drcp_rn((double)scale);
  }
};

__device__ __forceinline__ float div_exact(float x, const QP& p) {
  return __double2float_rn(__dmul_rn((double)x, p.rs));
}

// ROUNDING: 0 half-to-even (torch.round / nearbyint); -1 = runtime `rounding` in {0, 1, 2}
// (1: floor(v + .5), 2: ceil(v - .5), torch_extensions/common.cuh:66-74).
template <int ROUNDING>
__device__ __forceinline__ float round_q(float v, int rounding) {
  if (ROUNDING == 0) return rintf(v);
  if (rounding == 1) return floorf(__fadd_rn(v, 0.5f));
  if (rounding == 2) return ceilf(__fsub_rn(v, 0.5f));
  return rintf(v);
}

template <int ROUNDING>
__device__ __forceinline__ float qdq1(float x, const QP& p, int rounding) {
  float v = __fadd_rn(round_q<ROUNDING>(div_exact(x, p), rounding), p.zp);
  v = fmax_nan(fmin_nan(v, p.qmax), p.qmin);
  return __fmul_rn(__fsub_rn(v, p.zp), p.s);
}

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-thread running min/max.  min.NaN / max.NaN propagate NaN like torch.min / torch.max.
struct MinMaxAcc {
  float lo, hi;
  __device__ __forceinline__ void init() { lo = __int_as_float(0x7f800000); hi = __int_as_float(0xff800000); }
  __device__ __forceinline__ void add(float v) {
    lo = fmin_nan(lo, v);
    hi = fmax_nan(hi, v);
  }
  __device__ __forceinline__ void add4(const float4& v) { add(v.x); add(v.y); add(v.z); add(v.w); }
  __device__ __forceinline__ void warp_reduce() {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin_nan(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax_nan(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
  }
  // Publish into the {enc(min), enc(max)} state of one channel.  NaN -> both slots get the
  // extreme keys, which decode to NaN (see minmax_read).
  __device__ __forceinline__ void publish(uint32_t* st) const {
    if (lo != lo || hi != hi) {
      atomicMin(st, 0u);
      atomicMax(st + 1, 0xFFFFFFFFu);
    } else if (lo <= hi) {
      atomicMin(st, enc_f32(lo));
      atomicMax(st + 1, enc_f32(hi));
    }
  }
};

// Block-level reduce of a MinMaxAcc; result valid in warp 0.  `sm` needs 2*32 words.
__device__ __forceinline__ void block_reduce_minmax(MinMaxAcc& a, float* sm) {
  a.warp_reduce();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) {
    sm[wid] = a.lo;
    sm[32 + wid] = a.hi;
  }
  __syncthreads();
  if (wid == 0) {
    a.lo = lane < nw ? sm[lane] : __int_as_float(0x7f800000);
    a.hi = lane < nw ? sm[32 + lane] : __int_as_float(0xff800000);
    a.warp_reduce();
  }
  __syncthreads();
}
#endif  // __CUDACC__

}  // namespace sb200
