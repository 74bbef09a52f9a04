// adaround.cu -- AdaRound weight quantizer: the one quantizer of the reference that
// does not go through STE.  Replaces the ATen op chains of
//   sparsebit/quantization/quantizers/adaround.py:26-32 (init_variables), :40-43 (soft round values),
//   :46-54 (_forward)
// with one pass each.  Weights are small (<= 2.4 M elements per tensor), so these are plain
// grid-stride kernels: 4 B read per operand, 4 B written, float4 when the channel run allows it.
//
//   x_floor = floor(x / scale)                      (IEEE fp32 quotient -- div_exact)
//   h(v)    = clamp(sigmoid(v) * (zeta - gamma) + gamma, 0, 1),   zeta = 1.1, gamma = -0.1
//   soft:   x_q = clamp(x_floor + h(v) + zp, qmin, qmax)          (training)
//   hard:   x_q = clamp(x_floor + (v >= 0) + zp, qmin, qmax)      (evaluation; exact)
//   out     = (x_q - zp) * scale
// zero_point is NOT rounded here (the reference only rounds it inside STE.forward).
#include "common.cuh"

namespace sb200 {
namespace {

constexpr int kThreads = 256;
// zeta - gamma evaluated in Python doubles (1.1 - (-0.1) = 1.2000000000000002) then narrowed to the
// tensor's fp32, and gamma narrowed the same way -- what ATen does with a Python scalar operand.
__device__ __forceinline__ float k_stretch() { return (float)1.2000000000000002; }
__device__ __forceinline__ float k_gamma() { return (float)-0.1; }

__device__ __forceinline__ float sigmoid_f(float v) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-v))); }

// un-clamped soft-round value sigmoid(v) * 1.2 - 0.1 (two roundings, no FMA contraction: ATen runs
// mul and add as separate ops)
__device__ __forceinline__ float soft_raw(float v) { return __fadd_rn(__fmul_rn(sigmoid_f(v), k_stretch()), k_gamma()); }

struct ChanQP {
  float s, zp;
  double rs;
};
__device__ __forceinline__ ChanQP load_qp(const float* scale, const float* zp, long long c) {
  ChanQP p;
  p.s = __ldg(scale + c);
  p.zp = __ldg(zp + c);
  p.rs = __drcp_rn((double)p.s);
  return p;
}
__device__ __forceinline__ float quot(float x, const ChanQP& p) { return __double2float_rn(__dmul_rn((double)x, p.rs)); }

enum { kHard = 0, kSoft = 1 };

template <int MODE>
__device__ __forceinline__ float fwd1(float x, float v, const ChanQP& p, float qmin, float qmax) {
  const float fl = floorf(quot(x, p));
  float r;
  if (MODE == kSoft) {
    const float raw = soft_raw(v);
    r = (raw != raw) ? raw : fminf(fmaxf(raw, 0.f), 1.f);  // torch.clamp propagates NaN
  } else {
    r = v >= 0.f ? 1.f : 0.f;
  }
  const float qr = __fadd_rn(__fadd_rn(fl, r), p.zp);
  const float q = (qr != qr) ? qr : fminf(fmaxf(qr, qmin), qmax);
  return __fmul_rn(__fsub_rn(q, p.zp), p.s);
}

// d out / d v for the soft mode (autograd of adaround.py:40-54): clamp passes the gradient on the
// closed interval, floor has none, so only v receives one.
__device__ __forceinline__ float bwd1(float x, float v, float gy, const ChanQP& p, float qmin, float qmax) {
  const float fl = floorf(quot(x, p));
  const float sg = sigmoid_f(v);
  const float raw = __fadd_rn(__fmul_rn(sg, k_stretch()), k_gamma());
  const float r = fminf(fmaxf(raw, 0.f), 1.f);
  const float q = __fadd_rn(__fadd_rn(fl, r), p.zp);
  const bool pass = (q >= qmin) && (q <= qmax) && (raw >= 0.f) && (raw <= 1.f);
  // gy * scale -> through clamp -> through (+gamma) -> * stretch -> sigmoid'
  const float g = __fmul_rn(__fmul_rn(gy, p.s), k_stretch());
  return pass ? __fmul_rn(g, __fmul_rn(sg, __fsub_rn(1.f, sg))) : 0.f;
}

// v = -log(stretch / (rest - gamma) - 1), rest = x/scale - floor(x/scale); ATen evaluates
// `scalar / tensor` as reciprocal(tensor) * scalar.
__device__ __forceinline__ float init1(float x, const ChanQP& p) {
  const float qv = quot(x, p);
  const float rest = __fsub_rn(qv, floorf(qv));
  const float den = __fsub_rn(rest, k_gamma());
  const float ratio = __fmul_rn(__frcp_rn(den), k_stretch());
  return -logf(__fsub_rn(ratio, 1.f));
}

// One CTA walks chunks of (row = outer*C index, 1024 consecutive elements of that row): the
// channel's parameters are loaded once per chunk.
template <int OP, bool VEC>
__global__ void __launch_bounds__(kThreads) adaround_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                            const float* __restrict__ gy, const float* __restrict__ scale,
                                                            const float* __restrict__ zp, float* __restrict__ out,
                                                            long long rows, long long C, long long inner, float qmin,
                                                            float qmax) {
  constexpr int kChunk = kThreads * 4;
  const long long chunks_per_row = (inner + kChunk - 1) / kChunk;
  const long long total = rows * chunks_per_row;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const long long row = t / chunks_per_row;
    const long long k0 = (t - row * chunks_per_row) * kChunk;
    const ChanQP p = load_qp(scale, zp, row % C);
    const long long base = row * inner + k0;
    const long long left = inner - k0;
    if (VEC) {
      const long long k = (long long)threadIdx.x * 4;
      if (k < left) {  // inner % 4 == 0: whole float4 in range
        const float4 xv = *reinterpret_cast<const float4*>(x + base + k);
        float4 vv = make_float4(0.f, 0.f, 0.f, 0.f), gv = vv, o;
        if (OP != 3) vv = *reinterpret_cast<const float4*>(v + base + k);
        if (OP == 2) gv = *reinterpret_cast<const float4*>(gy + base + k);
        if (OP == 0 || OP == 1) {
          o.x = fwd1<OP>(xv.x, vv.x, p, qmin, qmax); o.y = fwd1<OP>(xv.y, vv.y, p, qmin, qmax);
          o.z = fwd1<OP>(xv.z, vv.z, p, qmin, qmax); o.w = fwd1<OP>(xv.w, vv.w, p, qmin, qmax);
        } else if (OP == 2) {
          o.x = bwd1(xv.x, vv.x, gv.x, p, qmin, qmax); o.y = bwd1(xv.y, vv.y, gv.y, p, qmin, qmax);
          o.z = bwd1(xv.z, vv.z, gv.z, p, qmin, qmax); o.w = bwd1(xv.w, vv.w, gv.w, p, qmin, qmax);
        } else {
          o.x = init1(xv.x, p); o.y = init1(xv.y, p); o.z = init1(xv.z, p); o.w = init1(xv.w, p);
        }
        *reinterpret_cast<float4*>(out + base + k) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long k = threadIdx.x + (long long)j * kThreads;
        if (k < left) {
          const float xv = x[base + k];
          const float vv = (OP != 3) ? v[base + k] : 0.f;
          float o;
          if (OP == 0 || OP == 1) o = fwd1<OP>(xv, vv, p, qmin, qmax);
          else if (OP == 2) o = bwd1(xv, vv, gy[base + k], p, qmin, qmax);
          else o = init1(xv, p);
          out[base + k] = o;
        }
      }
    }
  }
}

template <int OP>
int launch(const float* x, const float* v, const float* gy, const float* scale, const float* zp, float* out,
           long long outer, long long C, long long inner, int qmin, int qmax, cudaStream_t st) {
  const long long rows = outer * C;
  const long long chunks = rows * ((inner + kThreads * 4 - 1) / (kThreads * 4));
  const long long cap = (long long)sm_count() * 8;
  const unsigned grid = (unsigned)(chunks < cap ? chunks : cap);
  const bool vec = (inner % 4 == 0) && aligned16(x) && aligned16(out) && (OP == 3 || aligned16(v)) &&
                   (OP != 2 || aligned16(gy));
  if (vec)
    adaround_kernel<OP, true><<<grid, kThreads, 0, st>>>(x, v, gy, scale, zp, out, rows, C, inner, (float)qmin, (float)qmax);
  else
    adaround_kernel<OP, false><<<grid, kThreads, 0, st>>>(x, v, gy, scale, zp, out, rows, C, inner, (float)qmin, (float)qmax);
  SB_LAUNCHED();
  return SB200_OK;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_adaround_fwd(const float* x, const float* v, const float* scale, const float* zero_point, float* out,
                       int64_t outer, int64_t channels, int64_t inner, int qmin, int qmax, int soft, void* stream) {
  SB_REQUIRE(x && v && scale && zero_point && out, "sb200_adaround_fwd: null pointer argument");
  SB_REQUIRE(outer > 0 && channels > 0 && inner > 0, "sb200_adaround_fwd: empty tensor");
  SB_REQUIRE(qmin <= qmax, "sb200_adaround_fwd: qmin > qmax");
  SB_REQUIRE(soft == 0 || soft == 1, "sb200_adaround_fwd: soft must be 0 (hard / eval) or 1 (soft / training)");
  if (soft) return launch<1>(x, v, nullptr, scale, zero_point, out, outer, channels, inner, qmin, qmax, (cudaStream_t)stream);
  return launch<0>(x, v, nullptr, scale, zero_point, out, outer, channels, inner, qmin, qmax, (cudaStream_t)stream);
}

int sb200_adaround_bwd(const float* x, const float* v, const float* scale, const float* zero_point,
                       const float* grad_y, float* grad_v, int64_t outer, int64_t channels, int64_t inner, int qmin,
                       int qmax, void* stream) {
  SB_REQUIRE(x && v && scale && zero_point && grad_y && grad_v, "sb200_adaround_bwd: null pointer argument");
  SB_REQUIRE(outer > 0 && channels > 0 && inner > 0, "sb200_adaround_bwd: empty tensor");
  SB_REQUIRE(qmin <= qmax, "sb200_adaround_bwd: qmin > qmax");
  return launch<2>(x, v, grad_y, scale, zero_point, grad_v, outer, channels, inner, qmin, qmax, (cudaStream_t)stream);
}

int sb200_adaround_init(const float* x, const float* scale, float* v, int64_t outer, int64_t channels, int64_t inner,
                        void* stream) {
  SB_REQUIRE(x && scale && v, "sb200_adaround_init: null pointer argument");
  SB_REQUIRE(outer > 0 && channels > 0 && inner > 0, "sb200_adaround_init: empty tensor");
  // zero_point is unused by the init; pass scale as a valid dummy pointer
  return launch<3>(x, nullptr, nullptr, scale, scale, v, outer, channels, inner, 0, 0, (cudaStream_t)stream);
}

}  // extern "C"
