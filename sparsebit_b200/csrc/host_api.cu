// host_api.cu -- host-buffer ("end-to-end") entry points: the caller hands HOST pointers, the
// library pipelines H2D copy -> kernel -> D2H copy over a small ring of device chunks on its own
// streams.  This is the call a reference-side binding makes when its tensors live on the CPU
// (the reference's CalibrationRunner keeps every activation on the CPU between nodes,
// sparsebit/quantization/tools/calibration.py:38,157), and what bench.py times as `e2e`.
#include <mutex>

#include "common.cuh"

namespace sb200 {

constexpr int kRing = 4;
constexpr long long kChunkElems = 4LL << 20;  // 16 MiB of fp32 per chunk
constexpr int kMmSlots = 256;                 // MinMax states of calls in flight between two syncs

struct HostPipe {
  int device = -1;
  cudaStream_t streams[kRing] = {nullptr, nullptr, nullptr};
  cudaEvent_t ready = nullptr;
  float* d_in[kRing] = {nullptr, nullptr, nullptr};
  float* d_out[kRing] = {nullptr, nullptr, nullptr};
  long long cap = 0;       // elements per ring slot
  float* d_qp = nullptr;   // scale[C] | zp[C]
  long long qp_cap = 0;
  uint32_t* d_mm = nullptr;   // [kMmSlots][2]
  float* d_mmf = nullptr;     // [kMmSlots][2]
  cudaEvent_t done[kRing] = {nullptr, nullptr, nullptr, nullptr};
  int next_slot = 0;          // ring slot of the next chunk (kept across async calls)
  int mm_used = 0;            // MinMax slots handed out since the last sync
};
static HostPipe g_pipe;
static std::mutex g_pipe_mu;

static int pipe_prepare(HostPipe& p, long long slot_elems, long long qp_elems) {
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  if (p.device != dev) {  // (re)create for this device
    if (p.device >= 0) {
      for (int i = 0; i < kRing; ++i) {
        if (p.d_in[i]) cudaFree(p.d_in[i]);
        if (p.d_out[i]) cudaFree(p.d_out[i]);
        if (p.streams[i]) cudaStreamDestroy(p.streams[i]);
      }
      if (p.d_qp) cudaFree(p.d_qp);
      if (p.d_mm) cudaFree(p.d_mm);
      if (p.d_mmf) cudaFree(p.d_mmf);
      if (p.ready) cudaEventDestroy(p.ready);
      for (int i = 0; i < kRing; ++i)
        if (p.done[i]) cudaEventDestroy(p.done[i]);
      p = HostPipe();
    }
    for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamCreateWithFlags(&p.streams[i], cudaStreamNonBlocking));
    SB_CUDA(cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
    for (int i = 0; i < kRing; ++i) SB_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
    SB_CUDA(cudaMalloc(&p.d_mm, kMmSlots * 2 * sizeof(uint32_t)));
    SB_CUDA(cudaMalloc(&p.d_mmf, kMmSlots * 2 * sizeof(float)));
    p.device = dev;
  }
  if (slot_elems > p.cap) {
    for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamSynchronize(p.streams[i]));  // async calls may be in flight
    for (int i = 0; i < kRing; ++i) {
      if (p.d_in[i]) SB_CUDA(cudaFree(p.d_in[i]));
      if (p.d_out[i]) SB_CUDA(cudaFree(p.d_out[i]));
      p.d_in[i] = p.d_out[i] = nullptr;
    }
    p.cap = 0;
    for (int i = 0; i < kRing; ++i) {
      SB_CUDA(cudaMalloc(&p.d_in[i], (size_t)slot_elems * sizeof(float)));
      SB_CUDA(cudaMalloc(&p.d_out[i], (size_t)slot_elems * sizeof(float)));
    }
    p.cap = slot_elems;
  }
  if (qp_elems > p.qp_cap) {
    for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamSynchronize(p.streams[i]));
    if (p.d_qp) SB_CUDA(cudaFree(p.d_qp));
    p.d_qp = nullptr;
    SB_CUDA(cudaMalloc(&p.d_qp, (size_t)qp_elems * 2 * sizeof(float)));
    p.qp_cap = qp_elems;
  }
  return SB200_OK;
}

}  // namespace sb200

using namespace sb200;

extern "C" {

// Enqueue one per-tensor QDQ (+ optional MinMax) over host buffers: chunks go round-robin over the ring
// slots / streams, so copies of consecutive calls keep overlapping; nothing is waited for here.
static int pertensor_host_enqueue(HostPipe& p, const float* x_host, float scale, float zero_point, float* out_host,
                                  float* minmax_host, int64_t n, int qmin, int qmax, int rounding) {
  const long long chunk = n < kChunkElems ? n : kChunkElems;
  int rc = pipe_prepare(p, chunk, 4 * kMmSlots);
  if (rc) return rc;
  if (p.mm_used >= kMmSlots) {  // parameter / MinMax slots exhausted: drain before reusing them
    for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamSynchronize(p.streams[i]));
    p.mm_used = 0;
  }
  const int slot_id = p.mm_used++;
  float* d_qp = p.d_qp + 2 * slot_id;                 // {scale, zero_point} of THIS call
  uint32_t* d_mm = p.d_mm + 2 * slot_id;
  float* d_mmf = p.d_mmf + 2 * slot_id;
  const float qp[2] = {scale, zero_point};
  cudaStream_t s0 = p.streams[p.next_slot];
  SB_CUDA(cudaMemcpyAsync(d_qp, qp, sizeof(qp), cudaMemcpyHostToDevice, s0));  // pageable source: staged before return
  if (minmax_host) {
    rc = sb200_minmax_init(d_mm, 1, s0);
    if (rc) return rc;
  }
  SB_CUDA(cudaEventRecord(p.ready, s0));
  for (int i = 0; i < kRing; ++i)
    if (p.streams[i] != s0) SB_CUDA(cudaStreamWaitEvent(p.streams[i], p.ready, 0));
  bool used[kRing] = {false, false, false, false};
  for (long long off = 0; off < n; off += chunk) {
    const int slot = p.next_slot;
    p.next_slot = (p.next_slot + 1) % kRing;
    used[slot] = true;
    const long long len = (n - off) < chunk ? (n - off) : chunk;
    cudaStream_t st = p.streams[slot];
    SB_CUDA(cudaMemcpyAsync(p.d_in[slot], x_host + off, (size_t)len * 4, cudaMemcpyHostToDevice, st));
    if (minmax_host)
      rc = sb200_qdq_stats_pertensor_fwd(p.d_in[slot], d_qp, d_qp + 1, p.d_out[slot], d_mm, len, qmin, qmax, rounding, st);
    else
      rc = sb200_qdq_pertensor_fwd(p.d_in[slot], d_qp, d_qp + 1, p.d_out[slot], len, qmin, qmax, rounding, st);
    if (rc) return rc;
    SB_CUDA(cudaMemcpyAsync(out_host + off, p.d_out[slot], (size_t)len * 4, cudaMemcpyDeviceToHost, st));
  }
  if (minmax_host) {  // the statistics are complete once every stream that ran a chunk is done
    for (int i = 0; i < kRing; ++i)
      if (used[i] && p.streams[i] != s0) {
        SB_CUDA(cudaEventRecord(p.done[i], p.streams[i]));
        SB_CUDA(cudaStreamWaitEvent(s0, p.done[i], 0));
      }
    rc = sb200_minmax_read(d_mm, 1, d_mmf, d_mmf + 1, s0);
    if (rc) return rc;
    SB_CUDA(cudaMemcpyAsync(minmax_host, d_mmf, 2 * sizeof(float), cudaMemcpyDeviceToHost, s0));
  }
  return SB200_OK;
}

static int pipe_sync(HostPipe& p) {
  if (p.device < 0) return SB200_OK;
  for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamSynchronize(p.streams[i]));
  p.mm_used = 0;
  return SB200_OK;
}

int sb200_qdq_pertensor_fwd_host(const float* x_host, float scale, float zero_point, float* out_host,
                                 float* minmax_host, int64_t n, int qmin, int qmax, int rounding) {
  SB_REQUIRE(x_host && out_host, "sb200_qdq_pertensor_fwd_host: null pointer argument");
  SB_REQUIRE(n > 0, "sb200_qdq_pertensor_fwd_host: Kernel Failure, Tensor is empty: data");
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  int rc = pertensor_host_enqueue(g_pipe, x_host, scale, zero_point, out_host, minmax_host, n, qmin, qmax, rounding);
  if (rc) return rc;
  return pipe_sync(g_pipe);
}

int sb200_qdq_pertensor_fwd_host_async(const float* x_host, float scale, float zero_point, float* out_host,
                                       float* minmax_host, int64_t n, int qmin, int qmax, int rounding) {
  SB_REQUIRE(x_host && out_host, "sb200_qdq_pertensor_fwd_host_async: null pointer argument");
  SB_REQUIRE(n > 0, "sb200_qdq_pertensor_fwd_host_async: Kernel Failure, Tensor is empty: data");
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  return pertensor_host_enqueue(g_pipe, x_host, scale, zero_point, out_host, minmax_host, n, qmin, qmax, rounding);
}

int sb200_host_sync(void) {
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  return pipe_sync(g_pipe);
}

int sb200_qdq_perchannel_fwd_host(const float* x_host, const float* scale_host, const float* zero_point_host,
                                  float* out_host, int64_t outer, int64_t channels, int64_t inner, int qmin,
                                  int qmax, int rounding) {
  SB_REQUIRE(x_host && out_host && scale_host && zero_point_host, "sb200_qdq_perchannel_fwd_host: null pointer argument");
  SB_REQUIRE(outer > 0 && channels > 0 && inner > 0, "sb200_qdq_perchannel_fwd_host: Kernel Failure, Tensor is empty: data");
  std::lock_guard<std::mutex> lock(g_pipe_mu);
  HostPipe& p = g_pipe;
  const long long slab = channels * inner;  // one [C, inner] slab keeps the channel geometry intact
  long long slabs_per_chunk = kChunkElems / slab;
  if (slabs_per_chunk < 1) slabs_per_chunk = 1;
  if (slabs_per_chunk > outer) slabs_per_chunk = outer;
  int rc = pipe_prepare(p, slabs_per_chunk * slab, channels > 4 * kMmSlots ? channels : 4 * kMmSlots);
  if (rc) return rc;
  rc = pipe_sync(p);  // the parameter buffer is shared with in-flight async per-tensor calls
  if (rc) return rc;
  SB_CUDA(cudaMemcpyAsync(p.d_qp, scale_host, (size_t)channels * 4, cudaMemcpyHostToDevice, p.streams[0]));
  SB_CUDA(cudaMemcpyAsync(p.d_qp + p.qp_cap, zero_point_host, (size_t)channels * 4, cudaMemcpyHostToDevice, p.streams[0]));
  SB_CUDA(cudaEventRecord(p.ready, p.streams[0]));
  for (int i = 1; i < kRing; ++i) SB_CUDA(cudaStreamWaitEvent(p.streams[i], p.ready, 0));
  int slot = 0;
  for (long long o = 0; o < outer; o += slabs_per_chunk, slot = (slot + 1) % kRing) {
    const long long no = (outer - o) < slabs_per_chunk ? (outer - o) : slabs_per_chunk;
    const long long len = no * slab;
    cudaStream_t st = p.streams[slot];
    SB_CUDA(cudaMemcpyAsync(p.d_in[slot], x_host + o * slab, (size_t)len * 4, cudaMemcpyHostToDevice, st));
    rc = sb200_qdq_perchannel_fwd(p.d_in[slot], p.d_qp, p.d_qp + p.qp_cap, p.d_out[slot], no, channels, inner, qmin, qmax, rounding, st);
    if (rc) return rc;
    SB_CUDA(cudaMemcpyAsync(out_host + o * slab, p.d_out[slot], (size_t)len * 4, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < kRing; ++i) SB_CUDA(cudaStreamSynchronize(p.streams[i]));
  return SB200_OK;
}

}  // extern "C"
