// Device-side classification metrics (reference metrics/vision.py:27-58 PRF1 confusion matrix + mx.metric.Accuracy top-k,
// updated per batch at train.py:427-431 after a device->host copy of the logits).  Here the counters live on the device and a
// batch update is one kernel: no per-batch D2H synchronisation; the host reads the 64-bit counters once per epoch.
#include "tn_common.h"

namespace {

// One thread per sample.  pred = first maximal class (numpy argmax); the label's rank under a STABLE descending sort
// (np.argsort(-pred, kind='stable')): classes with a larger score, or an equal score and a lower index, come first.
__global__ void metrics_update_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int N, int C, int top_k,
                                      unsigned long long* __restrict__ conf, unsigned long long* __restrict__ hits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* row = logits + static_cast<size_t>(i) * C;
  const int y = labels[i];
  int best = 0;
  float bv = row[0];
  for (int c = 1; c < C; ++c) {
    const float v = row[c];
    if (v > bv) {
      bv = v;
      best = c;
    }
  }
  if (y >= 0 && y < C) {
    atomicAdd(&conf[static_cast<size_t>(y) * C + best], 1ull);
    const float yv = row[y];
    int rank = 0;
    for (int c = 0; c < C; ++c) {
      const float v = row[c];
      rank += (v > yv || (v == yv && c < y)) ? 1 : 0;
    }
    if (rank < 1) atomicAdd(&hits[0], 1ull);      // top-1
    if (rank < top_k) atomicAdd(&hits[1], 1ull);  // top-k
  }
  atomicAdd(&hits[2], 1ull);  // samples seen
}

}  // namespace

extern "C" {

// logits device fp32 (N,C), labels device int32 (N); conf device uint64 (C,C) [label][prediction], hits device uint64 [3] =
// {top-1 hits, top-k hits, samples}; counters are ACCUMULATED (zero them to reset).
int tn_metrics_update(const float* logits, const int32_t* labels, int N, int C, int top_k, unsigned long long* conf,
                      unsigned long long* hits, tn_stream_t stream) {
  if (N < 0 || C <= 0 || top_k < 1) return tn::set_error(TN_ERR_INVALID, "bad metrics_update arguments");
  if (N == 0) return TN_OK;
  if (!logits || !labels || !conf || !hits) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tn::ProfScope ps(tn::kProfOther, st);
  metrics_update_kernel<<<(N + 127) / 128, 128, 0, st>>>(logits, labels, N, C, top_k, conf, hits);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

}  // extern "C"
