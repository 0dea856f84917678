#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tn {

struct RnnScanParams {
  int gates;  // 3 = GRU [r,z,n], 4 = LSTM [i,f,g,o]
  int B, T, H;
  int ndir;          // 1 or 2
  int reverse_dir1;  // direction 1 scans right-to-left (bidirectional layer)
  // x*W_ih^T + b_ih for every (b,t): fp32 [B*T][ndir*gates*H]
  const float* gx;
  const float* WhhT;  // [ndir][H][gates*H]  (W_hh transposed)
  const float* bhh;   // [ndir][gates*H]
  const int* valid_len;  // nullable, [B]
  const float* h0;       // nullable, [ndir][B][H]
  const float* c0;       // nullable, [ndir][B][H]
  float* y;              // nullable, [B][T][ndir*H]  (pre-zeroed by the caller when valid_len is given)
  float* cseq;           // nullable, LSTM cell state per step [B][T][ndir*H] (saved for the backward pass)
  float* ymax;           // nullable, [B][ndir*H]   max over time
  float* h_final;        // nullable, [ndir][B][H]
  float* c_final;        // nullable, [ndir][B][H]
};

int rnn_scan_cluster_size(int gates, int H);
cudaError_t launch_rnn_scan(const RnnScanParams& p, cudaStream_t st);

}  // namespace tn
