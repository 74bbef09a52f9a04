// multi.cu -- ONE launch for the weight-side fake-quant work of a whole model.
//
// A QAT / PTQ forward runs `weight * w_mask` (sparse/modules/conv.py:39-43, linear.py:30-34) and the per-channel
// weight quantizer (quantizers/quant_tensor.py:181-184, torch_extensions/fake_quant_tensor.cu:170-224) once per
// QConv2d / QLinear: 54 small launches for ResNet-50, each far too short to fill the GPU (single layers are
// L2-resident and launch-bound: 0.25-0.29 of the HBM roofline when issued one by one).  Here the caller describes
// all tensors once (sb200_qdq_multi_plan -> a device-resident table) and every forward is a single launch
// (sb200_qdq_multi_run) whose work items are the (tensor, row) pairs of all tensors: a warp owns a row
// (= one channel slice of `inner` contiguous elements), so scale / zero-point / reciprocal are per-warp scalars
// and there is no per-element index arithmetic at all.  9 B/elem with a mask, 8 B/elem without.
#include "common.cuh"

namespace sb200 {

struct MultiEntry {        // device-side descriptor, 64 bytes
  const float* x;
  const uint8_t* mask;     // nullable
  const float* scale;      // `channels` floats
  const float* zero_point; // `channels` floats
  float* out;
  long long row_begin;     // first global row of this tensor
  int channels;
  int inner;
  float qmin, qmax;
};
static_assert(sizeof(MultiEntry) == 64, "MultiEntry layout");

constexpr int kMultiThreads = 256;

template <bool ALIGNED>
__device__ __forceinline__ void multi_row(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                          float* __restrict__ out, int inner, const QP& p, int lane) {
  if (ALIGNED) {  // row start 16-byte aligned for x / out and 4-byte aligned for the mask, inner % 4 == 0
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* o4 = reinterpret_cast<float4*>(out);
    const uchar4* m4 = reinterpret_cast<const uchar4*>(mask);
    const int nv = inner >> 2;
    for (int i = lane; i < nv; i += 64) {
      const bool two = i + 32 < nv;
      float4 a = ld_stream4(x4 + i), b = two ? ld_stream4(x4 + i + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (mask) {
        const uchar4 ma = __ldcs(m4 + i);
        a.x = __fmul_rn(a.x, ma.x ? 1.f : 0.f); a.y = __fmul_rn(a.y, ma.y ? 1.f : 0.f);
        a.z = __fmul_rn(a.z, ma.z ? 1.f : 0.f); a.w = __fmul_rn(a.w, ma.w ? 1.f : 0.f);
        if (two) {
          const uchar4 mb = __ldcs(m4 + i + 32);
          b.x = __fmul_rn(b.x, mb.x ? 1.f : 0.f); b.y = __fmul_rn(b.y, mb.y ? 1.f : 0.f);
          b.z = __fmul_rn(b.z, mb.z ? 1.f : 0.f); b.w = __fmul_rn(b.w, mb.w ? 1.f : 0.f);
        }
      }
      a.x = qdq1<0>(a.x, p, 0); a.y = qdq1<0>(a.y, p, 0); a.z = qdq1<0>(a.z, p, 0); a.w = qdq1<0>(a.w, p, 0);
      st_stream4(o4 + i, a);
      if (two) {
        b.x = qdq1<0>(b.x, p, 0); b.y = qdq1<0>(b.y, p, 0); b.z = qdq1<0>(b.z, p, 0); b.w = qdq1<0>(b.w, p, 0);
        st_stream4(o4 + i + 32, b);
      }
    }
  } else {
    for (int i = lane; i < inner; i += 32) {
      float v = ld_stream1(x + i);
      if (mask) v = __fmul_rn(v, __ldcs(mask + i) ? 1.f : 0.f);
      st_stream1(out + i, qdq1<0>(v, p, 0));
    }
  }
}

__global__ void __launch_bounds__(kMultiThreads) qdq_multi_kernel(const MultiEntry* __restrict__ table, int count,
                                                                  long long total_rows) {
  extern __shared__ long long s_begin[];  // row_begin of every tensor (+ total_rows): the search key
  for (int i = threadIdx.x; i <= count; i += blockDim.x) s_begin[i] = i < count ? table[i].row_begin : total_rows;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  int t = 0;
  for (long long row = warp0; row < total_rows; row += nwarps) {
    while (row >= s_begin[t + 1]) ++t;  // rows are visited in increasing order: the tensor index only moves forward
    const MultiEntry e = table[t];
    const long long r = row - e.row_begin;
    const int c = (int)(r % e.channels);
    QP p;
    p.set(__ldg(e.scale + c), __ldg(e.zero_point + c));
    p.qmin = e.qmin;
    p.qmax = e.qmax;
    const long long off = r * e.inner;
    const float* x = e.x + off;
    float* out = e.out + off;
    const uint8_t* mask = e.mask ? e.mask + off : nullptr;
    const bool aligned = ((e.inner & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) &&
                         (!mask || (reinterpret_cast<uintptr_t>(mask) & 3u) == 0);
    if (aligned) multi_row<true>(x, mask, out, e.inner, p, lane);
    else multi_row<false>(x, mask, out, e.inner, p, lane);
  }
}

// Structured (filter) pruning mask: mask[c, :] = (score[c] > thresh) ? 1 : 0 as a FLOAT mask shaped like the weight
// (the reference builds ones_like(x) and zeroes the pruned filters one index at a time through the CPU,
// sparse/sparsers/l1norm.py:27-40).  One warp per row, 128-bit stores.
__global__ void __launch_bounds__(256) mask_rows_gt_kernel(const float* __restrict__ score, const float* __restrict__ thresh,
                                                           float* __restrict__ mask, long long rows, long long inner) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float th = __ldg(thresh);
  for (long long r = warp0; r < rows; r += nwarps) {
    const float v = (__ldg(score + r) > th) ? 1.f : 0.f;
    float* row = mask + r * inner;
    if ((inner & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15u) == 0) {
      const float4 v4 = make_float4(v, v, v, v);
      for (long long i = lane; i < (inner >> 2); i += 32) reinterpret_cast<float4*>(row)[i] = v4;
    } else {
      for (long long i = lane; i < inner; i += 32) row[i] = v;
    }
  }
}

// MinMax observers of a whole model: running min/max states -> (min, max, scale, zero_point) in ONE launch
// (observers/minmax.py:14-25 + Observer.calc_qparams_with_minmax, observers/base.py:63-79, which the reference runs as
// ~15 tiny ATen ops per quantizer).  One CTA per quantizer; IEEE op order of the Python statement, NaN propagating like
// torch.minimum / torch.maximum.
struct QparamEntry {  // device-side descriptor, 64 bytes
  const uint32_t* state;  // {enc(min), enc(max)} per channel
  float* out_min;
  float* out_max;
  float* out_scale;
  float* out_zp;
  long long channels;
  float span;  // qmax - qmin
  int symmetric;
  long long pad;
};
static_assert(sizeof(QparamEntry) == 64, "QparamEntry layout");

__global__ void __launch_bounds__(128) minmax_qparams_kernel(const QparamEntry* __restrict__ table) {
  const QparamEntry e = table[blockIdx.x];
  for (long long c = threadIdx.x; c < e.channels; c += blockDim.x) {
    const float mn = dec_f32(e.state[2 * c]), mx = dec_f32(e.state[2 * c + 1]);
    const float min_neg = fmin_nan(mn, 0.f), max_pos = fmax_nan(mx, 0.f);
    float scale, zp;
    if (e.symmetric) {
      scale = fmax_nan(__fdiv_rn(__fmul_rn(fmax_nan(-min_neg, max_pos), 2.f), e.span), 1e-6f);
      zp = 0.f;
    } else {
      scale = fmax_nan(__fdiv_rn(__fsub_rn(max_pos, min_neg), e.span), 1e-6f);
      zp = rintf(__fdiv_rn(-min_neg, scale));
    }
    e.out_min[c] = mn;
    e.out_max[c] = mx;
    e.out_scale[c] = scale;
    e.out_zp[c] = zp;
  }
}

// PACT (quantizers/pact.py:43-46): backward of clamp(x, lo, hi) with learnable bounds, the only thing the reference's
// extra clamp pass in front of the fake-quant op contributes (the forward value is unchanged: the QDQ grid built from
// [lo, hi] clamps to the same points).   gx = gy * [lo <= x <= hi];  g_hi = sum gy * [x > hi];  g_lo = sum gy * [x < lo]
// (torch.clamp's gradient convention).  12 B/elem; per-CTA fp64 partials + a fixed-order finish => deterministic.
__global__ void __launch_bounds__(256) clamp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                        const float* __restrict__ lo_p, const float* __restrict__ hi_p,
                                                        float* __restrict__ gx, long long n, double2* __restrict__ partial) {
  __shared__ double s_red[2][8];
  const float lo = __ldg(lo_p), hi = __ldg(hi_p);
  double d_hi = 0.0, d_lo = 0.0;
  float a_hi = 0.f, a_lo = 0.f;
  int k = 0;
  auto one = [&](float xv, float g) -> float {
    a_hi += xv > hi ? g : 0.f;
    a_lo += xv < lo ? g : 0.f;
    return (xv >= lo && xv <= hi) ? g : 0.f;
  };
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx)) & 15u) == 0;
  const long long nv = vec ? (n >> 2) : 0;
  for (long long i = t0; i < nv; i += nt) {
    const float4 xv = ld_stream4(reinterpret_cast<const float4*>(x) + i), g = ld_stream4(reinterpret_cast<const float4*>(gy) + i);
    float4 o;
    o.x = one(xv.x, g.x); o.y = one(xv.y, g.y); o.z = one(xv.z, g.z); o.w = one(xv.w, g.w);
    st_stream4(reinterpret_cast<float4*>(gx) + i, o);
    if (++k == 64) { d_hi += a_hi; d_lo += a_lo; a_hi = a_lo = 0.f; k = 0; }  // bound the fp32 accumulation length
  }
  for (long long i = (nv << 2) + t0; i < n; i += nt) gx[i] = one(__ldcs(x + i), __ldcs(gy + i));
  d_hi += a_hi;
  d_lo += a_lo;
  d_hi = warp_sum(d_hi);
  d_lo = warp_sum(d_lo);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][wid] = d_hi; s_red[1][wid] = d_lo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double h = 0.0, l = 0.0;
    for (int w = 0; w < 8; ++w) { h += s_red[0][w]; l += s_red[1][w]; }
    partial[blockIdx.x] = make_double2(h, l);
  }
}
__global__ void clamp_bwd_finish_kernel(const double2* __restrict__ partial, int n, float* __restrict__ g_hi, float* __restrict__ g_lo) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double h = 0.0, l = 0.0;
    for (int i = 0; i < n; ++i) { h += partial[i].x; l += partial[i].y; }
    if (g_hi) g_hi[0] = (float)h;
    if (g_lo) g_lo[0] = (float)l;
  }
}
static inline int clamp_bwd_ctas(long long n) {
  long long c = (n / 4 + 255) / 256, cap = (long long)sm_count() * 4;
  return (int)(c < 1 ? 1 : (c > cap ? cap : c));
}

}  // namespace sb200

using namespace sb200;

extern "C" {

size_t sb200_qdq_multi_table_bytes(int count) { return count > 0 ? (size_t)count * sizeof(MultiEntry) : 0; }

int sb200_qdq_multi_plan(const sb200_qdq_tensor_desc* descs, int count, void* device_table, size_t table_bytes,
                         int64_t* total_rows, void* stream) {
  SB_REQUIRE(descs && device_table && total_rows, "sb200_qdq_multi_plan: null pointer argument");
  SB_REQUIRE(count > 0 && count <= 4096, "sb200_qdq_multi_plan: count must be in [1, 4096] (got %d)", count);
  SB_REQUIRE(table_bytes >= (size_t)count * sizeof(MultiEntry), "sb200_qdq_multi_plan: table too small");
  MultiEntry* host = new MultiEntry[count];
  long long rows = 0;
  for (int i = 0; i < count; ++i) {
    const sb200_qdq_tensor_desc& d = descs[i];
    if (!d.x || !d.scale || !d.zero_point || !d.out || d.outer <= 0 || d.channels <= 0 || d.inner <= 0 ||
        d.channels >= (1LL << 31) || d.inner >= (1LL << 31) || d.qmin > d.qmax) {
      delete[] host;
      set_error("sb200_qdq_multi_plan: bad descriptor %d (null pointer, empty tensor or qmin > qmax)", i);
      return SB200_E_INVALID;
    }
    host[i] = MultiEntry{d.x, d.mask, d.scale, d.zero_point, d.out, rows, (int)d.channels, (int)d.inner, (float)d.qmin,
                         (float)d.qmax};
    rows += d.outer * d.channels;
  }
  *total_rows = rows;
  // synchronous on purpose: the staging copy is freed right after, and a plan is built once per model
  cudaError_t e = cudaMemcpyAsync(device_table, host, (size_t)count * sizeof(MultiEntry), cudaMemcpyHostToDevice,
                                  (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  delete[] host;
  SB_CUDA(e);
  return SB200_OK;
}

int sb200_minmax_qparams_multi(const sb200_minmax_qparams_desc* descs, int count, void* device_table, size_t table_bytes,
                               void* stream) {
  SB_REQUIRE(descs && device_table, "sb200_minmax_qparams_multi: null pointer argument");
  SB_REQUIRE(count > 0 && count <= 65535, "sb200_minmax_qparams_multi: count must be in [1, 65535] (got %d)", count);
  SB_REQUIRE(table_bytes >= (size_t)count * sizeof(QparamEntry), "sb200_minmax_qparams_multi: table too small");
  QparamEntry* host = new QparamEntry[count];
  for (int i = 0; i < count; ++i) {
    const sb200_minmax_qparams_desc& d = descs[i];
    if (!d.state || !d.out_min || !d.out_max || !d.out_scale || !d.out_zero_point || d.channels <= 0 || d.qmin >= d.qmax) {
      delete[] host;
      set_error("sb200_minmax_qparams_multi: bad descriptor %d (null pointer, no channels or qmin >= qmax)", i);
      return SB200_E_INVALID;
    }
    host[i] = QparamEntry{d.state, d.out_min, d.out_max, d.out_scale, d.out_zero_point, (long long)d.channels,
                          (float)(d.qmax - d.qmin), d.symmetric ? 1 : 0, 0};
  }
  cudaError_t e = cudaMemcpyAsync(device_table, host, (size_t)count * sizeof(QparamEntry), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);  // the staging copy is freed right below
  delete[] host;
  SB_CUDA(e);
  minmax_qparams_kernel<<<(unsigned)count, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const QparamEntry*>(device_table));
  SB_LAUNCHED();
  return SB200_OK;
}

size_t sb200_clamp_bwd_workspace_bytes(int64_t n) { return n > 0 ? (size_t)clamp_bwd_ctas(n) * sizeof(double2) : 0; }

int sb200_clamp_bwd(const float* x, const float* grad_y, const float* lo, const float* hi, float* grad_x, float* grad_hi,
                    float* grad_lo, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  SB_REQUIRE(x && grad_y && lo && hi && grad_x, "sb200_clamp_bwd: null pointer argument");
  SB_REQUIRE(n > 0, "sb200_clamp_bwd: Kernel Failure, Tensor is empty: data");
  const int ctas = clamp_bwd_ctas(n);
  if (!workspace || workspace_bytes < (size_t)ctas * sizeof(double2)) {
    set_error("sb200_clamp_bwd: workspace too small");
    return SB200_E_WORKSPACE;
  }
  clamp_bwd_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(x, grad_y, lo, hi, grad_x, n, (double2*)workspace);
  SB_LAUNCHED();
  clamp_bwd_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const double2*)workspace, ctas, grad_hi, grad_lo);
  SB_LAUNCHED();
  return SB200_OK;
}

int sb200_mask_rows_gt(const float* score, const float* thresh, float* mask, int64_t rows, int64_t inner, void* stream) {
  SB_REQUIRE(score && thresh && mask, "sb200_mask_rows_gt: null pointer argument");
  SB_REQUIRE(rows > 0 && inner > 0, "sb200_mask_rows_gt: empty tensor");
  long long ctas = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (ctas > cap) ctas = cap;
  mask_rows_gt_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(score, thresh, mask, rows, inner);
  SB_LAUNCHED();
  return SB200_OK;
}

int sb200_qdq_multi_run(const void* device_table, int count, int64_t total_rows, void* stream) {
  SB_REQUIRE(device_table && count > 0 && total_rows > 0, "sb200_qdq_multi_run: bad arguments");
  long long ctas = (total_rows + (kMultiThreads / 32) - 1) / (kMultiThreads / 32);
  const long long cap = (long long)sm_count() * 8;
  if (ctas > cap) ctas = cap;
  qdq_multi_kernel<<<(unsigned)ctas, kMultiThreads, (size_t)(count + 1) * sizeof(long long), (cudaStream_t)stream>>>(
      reinterpret_cast<const MultiEntry*>(device_table), count, total_rows);
  SB_LAUNCHED();
  return SB200_OK;
}

}  // extern "C"
