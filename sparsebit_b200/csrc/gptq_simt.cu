// gptq_simt.cu -- GPTQ int4 group-wise dequant-matmul, HBM-bound SIMT path (decode / small M and
// every shape the tensor-core path cannot take).
//
// Replaces large_language_models/llama/quantization/cuda/cuda_kernel_4bit.cu:36-180
// (VecQuant4MatMulKernel) behind the same contract (cuda_kernel.cpp:10-23):
//   out[m, n] += sum_k (scales[n*G + k/gs] * nib(qweight[k/8, n], k%8) - zeros[n*G + k/gs]) * x[m, k]
//
// Layout facts taken from the reference: qweight int32 [ceil(K/8), N], nibble j of word (kw, n)
// is k = 8*kw + j (utils/quant.py:220-225); scales / zeros are [N, G] with zeros = zero*scale
// (utils/quant.py:188); a 128-wide K block never straddles a group because group_size % 128 == 0
// (cuda_kernel_4bit.cu:60, :109-112).
//
// Design: one thread per output column (coalesced 128 B weight rows per warp), a CTA owns a
// [S*128 x 128] slab of the packed matrix, all 16 word loads of a 128-K block are issued before
// the first is used, the activation block lives in shared memory (broadcast float4 reads) and
// the per-group affine is applied once per block on the integer dot product:
//   sum_k (s*q - z) * x  =  s * (sum_k q*x) - z * (sum_k x)
// so the inner loop is nibble -> float (3 ALU ops) + MT FMAs.  K is split across CTAs; partial
// results are combined with fp32 atomicAdd exactly like the reference (cuda_kernel_4bit.cu:158,179).
#include "common.cuh"

namespace sb200 {

constexpr int kCols = 128;   // threads per CTA == output columns per CTA
constexpr int kBlockK = 128; // K elements per block (16 packed rows)

__device__ __forceinline__ float nib_to_float(uint32_t w, int j) {
  // (w >> 4j) & 15 placed in the mantissa of 2^23, then 2^23 subtracted: exact, no I2F.
  return __uint_as_float(((w >> (4 * j)) & 0xFu) | 0x4B000000u) - 8388608.0f;
}

template <int MT>
__global__ void __launch_bounds__(kCols) gptq4_simt_kernel(const float* __restrict__ x, const uint32_t* __restrict__ qw,
                                                           float* __restrict__ out, const float* __restrict__ scales,
                                                           const float* __restrict__ zeros, int M, int K, int N,
                                                           int KW, int G, int group_size, int blocks_per_slice) {
  __shared__ __align__(16) float xs[MT][kBlockK];
  __shared__ float xsum[MT];
  __shared__ float xpart[MT][kCols / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = blockIdx.x * kCols + tid;
  const bool col_ok = n < N;
  const int nblk = (K + kBlockK - 1) / kBlockK;
  const int b0 = blockIdx.y * blocks_per_slice;
  const int b1 = min(b0 + blocks_per_slice, nblk);
  for (int m0 = 0; m0 < M; m0 += MT) {
    float acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = 0.f;
    for (int b = b0; b < b1; ++b) {
      const int k0 = b * kBlockK;
      // issue the 16 packed-word loads of this block first
      uint32_t w[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int row = (k0 >> 3) + r;
        w[r] = (col_ok && row < KW) ? __ldcs(qw + (size_t)row * N + n) : 0u;
      }
      // stage x[m0..m0+MT, k0..k0+127] (zero padded) and its row sums
      __syncthreads();
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const int k = k0 + tid;
        const float v = (m0 + m < M && k < K) ? __ldg(x + (size_t)(m0 + m) * K + k) : 0.f;
        xs[m][tid] = v;
        const float s = warp_sum(v);
        if (lane == 0) xpart[m][wid] = s;
      }
      __syncthreads();
      if (tid < MT) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kCols / 32; ++i) s += xpart[tid][i];
        xsum[tid] = s;
      }
      __syncthreads();
      const int g = k0 / group_size;
      const float sc = col_ok ? __ldg(scales + (size_t)n * G + g) : 0.f;
      const float zr = col_ok ? __ldg(zeros + (size_t)n * G + g) : 0.f;
      float dot[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) dot[m] = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] = nib_to_float(w[r], j);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 xa = *reinterpret_cast<const float4*>(&xs[m][r * 8]);
          const float4 xb = *reinterpret_cast<const float4*>(&xs[m][r * 8 + 4]);
          dot[m] = fmaf(q[0], xa.x, dot[m]);
          dot[m] = fmaf(q[1], xa.y, dot[m]);
          dot[m] = fmaf(q[2], xa.z, dot[m]);
          dot[m] = fmaf(q[3], xa.w, dot[m]);
          dot[m] = fmaf(q[4], xb.x, dot[m]);
          dot[m] = fmaf(q[5], xb.y, dot[m]);
          dot[m] = fmaf(q[6], xb.z, dot[m]);
          dot[m] = fmaf(q[7], xb.w, dot[m]);
        }
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) acc[m] += fmaf(sc, dot[m], -zr * xsum[m]);
    }
    if (col_ok) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
        if (m0 + m < M) atomicAdd(out + (size_t)(m0 + m) * N + n, acc[m]);
    }
  }
}

int gptq4_simt(const float* x, const int32_t* qweight, float* out, const float* scales, const float* zeros,
               long long M, long long K, long long N, long long KW, int group_size, cudaStream_t st) {
  const int G = (int)((K + group_size - 1) / group_size);
  const int nblk = (int)((K + kBlockK - 1) / kBlockK);
  const int colblocks = (int)((N + kCols - 1) / kCols);
  int want_slices = (sm_count() * 4 + colblocks - 1) / colblocks;
  if (want_slices < 1) want_slices = 1;
  if (want_slices > nblk) want_slices = nblk;
  const int S = (nblk + want_slices - 1) / want_slices;
  const int slices = (nblk + S - 1) / S;
  const dim3 grid((unsigned)colblocks, (unsigned)slices);
  const uint32_t* qw = reinterpret_cast<const uint32_t*>(qweight);
#define SB_GO(MT_) \
  gptq4_simt_kernel<MT_><<<grid, kCols, 0, st>>>(x, qw, out, scales, zeros, (int)M, (int)K, (int)N, (int)KW, G, group_size, S)
  if (M == 1) SB_GO(1);
  else if (M == 2) SB_GO(2);
  else if (M <= 4) SB_GO(4);
  else SB_GO(8);
#undef SB_GO
  SB_LAUNCHED();
  return SB200_OK;
}

}  // namespace sb200
