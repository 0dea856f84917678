// C-ABI for Dense / temporal pooling / the fused (bi)RNN layer.
#include <string.h>

#include <memory>

#include "tn_common.h"
#include "tn_elementwise.h"
#include "tn_rnn.h"

struct tn_birnn {
  int device = 0, cell = 0, gates = 3, D = 0, H = 0, ndir = 1;
  tn::DeviceArena arena;
  tn::ConvDev proj;          // W_ih of all directions stacked: (ndir*G*H, D)
  tn::ConvDev proj3;         // the same weights split into bf16 hi/lo parts along K: [hi | hi | lo]  (ndir*G*H, 3D)
  int precise = 0;           // 1 -> input projection in split-bf16 (x = hi + lo): ~16 mantissa bits instead of 8
  const float* bih = nullptr;   // [ndir*G*H]
  const float* WhhT = nullptr;  // [ndir][H][G*H]
  const float* bhh = nullptr;   // [ndir*G*H]
};

namespace tn {
void birnn_dims(const tn_birnn* r, int* G, int* D, int* H, int* ndir) {
  *G = r->gates;
  *D = r->D;
  *H = r->H;
  *ndir = r->ndir;
}
const float* birnn_whhT(const tn_birnn* r) { return r->WhhT; }
const float* birnn_bhh(const tn_birnn* r) { return r->bhh; }
}  // namespace tn

static int birnn_forward_impl(tn_birnn_t* r, const void* x, int x_is_bf16, const int32_t* valid_len, int B, int T, float* y,
                              float* ymax, float* h_final, float* c_final, float* gx_out, float* cseq, void* workspace,
                              size_t workspace_bytes, tn_stream_t stream);

extern "C" {

int tn_dense_forward(const float* x, const float* weight, const float* bias, float* y, int rows, int in_dim,
                     int out_dim, tn_stream_t stream) {
  if (rows < 0 || in_dim <= 0 || out_dim <= 0) return tn::set_error(TN_ERR_INVALID, "bad dense shape");
  if (rows == 0) return TN_OK;
  if (!x || !weight || !y) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  TN_CUDA(tn::launch_dense(x, weight, bias, y, rows, in_dim, out_dim, static_cast<cudaStream_t>(stream)));
  return TN_OK;
}

int tn_temporal_pool(const float* x, float* y, int B, int T, int D, int pool, tn_stream_t stream) {
  if (B < 0 || T <= 0 || D <= 0) return tn::set_error(TN_ERR_INVALID, "bad pool shape");
  if (B == 0) return TN_OK;
  if (!x || !y) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  TN_CUDA(tn::launch_temporal_pool(x, y, B, T, D, pool == TN_POOL_MEAN ? 1 : 0, static_cast<cudaStream_t>(stream)));
  return TN_OK;
}

int tn_birnn_create(tn_birnn_t** out, int device, int cell, int D, int H, int ndir, const float* const* i2h_weight,
                    const float* const* h2h_weight, const float* const* i2h_bias, const float* const* h2h_bias) {
  if (!out) return tn::set_error(TN_ERR_INVALID, "null out");
  *out = nullptr;
  int rc = tn::check_arch(device);
  if (rc != TN_OK) return rc;
  if ((cell != TN_CELL_GRU && cell != TN_CELL_LSTM) || (ndir != 1 && ndir != 2) || D <= 0 || H <= 0)
    return tn::set_error(TN_ERR_INVALID, "bad rnn config");
  const int G = cell == TN_CELL_GRU ? 3 : 4;
  if (D % 16 != 0) return tn::set_error(TN_ERR_INVALID, "input width %d must be a multiple of 16", D);
  if (H % 32 != 0 || tn::rnn_scan_cluster_size(G, H) < 0) return tn::set_error(TN_ERR_INVALID, "unsupported hidden size %d", H);
  for (int d = 0; d < ndir; ++d)
    if (!i2h_weight[d] || !h2h_weight[d] || !i2h_bias[d] || !h2h_bias[d]) return tn::set_error(TN_ERR_INVALID, "null weight");
  TN_CUDA(cudaSetDevice(device));
  std::unique_ptr<tn_birnn> r(new tn_birnn);
  r->device = device;
  r->cell = cell;
  r->gates = G;
  r->D = D;
  r->H = H;
  r->ndir = ndir;
  const int GH = G * H;
  std::vector<float> wih(static_cast<size_t>(ndir) * GH * D), bih(static_cast<size_t>(ndir) * GH),
      bhh(static_cast<size_t>(ndir) * GH), whhT(static_cast<size_t>(ndir) * H * GH);
  for (int d = 0; d < ndir; ++d) {
    memcpy(&wih[static_cast<size_t>(d) * GH * D], i2h_weight[d], sizeof(float) * GH * D);
    memcpy(&bih[static_cast<size_t>(d) * GH], i2h_bias[d], sizeof(float) * GH);
    memcpy(&bhh[static_cast<size_t>(d) * GH], h2h_bias[d], sizeof(float) * GH);
    for (int j = 0; j < GH; ++j)
      for (int k = 0; k < H; ++k) whhT[(static_cast<size_t>(d) * H + k) * GH + j] = h2h_weight[d][static_cast<size_t>(j) * H + k];
  }
  if (!tn::make_conv(r->arena, wih.data(), ndir * GH, D, 1, 1, tn::kModeConv, &r->proj)) return TN_ERR_CUDA;
  {
    // x W^T ~= x_hi W_hi^T + x_lo W_hi^T + x_hi W_lo^T as ONE GEMM over K' = 3D: A' = [x_hi | x_lo | x_hi], W' = [W_hi | W_hi | W_lo]
    std::vector<float> w3(static_cast<size_t>(ndir) * GH * 3 * D);
    for (size_t n = 0; n < static_cast<size_t>(ndir) * GH; ++n)
      for (int k = 0; k < D; ++k) {
        const float w = wih[n * D + k];
        const float hi = __bfloat162float(__float2bfloat16(w));
        w3[n * 3 * D + k] = hi;
        w3[n * 3 * D + D + k] = hi;
        w3[n * 3 * D + 2 * D + k] = w - hi;
      }
    if (!tn::make_conv(r->arena, w3.data(), ndir * GH, 3 * D, 1, 1, tn::kModeConv, &r->proj3)) return TN_ERR_CUDA;
  }
  r->bih = static_cast<const float*>(r->arena.upload(bih.data(), bih.size() * sizeof(float)));
  r->bhh = static_cast<const float*>(r->arena.upload(bhh.data(), bhh.size() * sizeof(float)));
  r->WhhT = static_cast<const float*>(r->arena.upload(whhT.data(), whhT.size() * sizeof(float)));
  if (!r->bih || !r->bhh || !r->WhhT) return TN_ERR_CUDA;
  *out = r.release();
  return TN_OK;
}

void tn_birnn_destroy(tn_birnn_t* r) { delete r; }

int tn_birnn_set_precise(tn_birnn_t* r, int on) {
  if (!r) return tn::set_error(TN_ERR_INVALID, "null handle");
  r->precise = on ? 1 : 0;
  return TN_OK;
}


}  // extern "C"

namespace {
// Device-side re-pack of the layer's weights (training: the optimiser updates the fp32 parameters every step; rebuilding
// the handle through the host cost ~50 ms per step).  Same shared-memory image as tn::make_conv (K-major, 128-byte swizzle).
__device__ __forceinline__ size_t wpack_offset(int co, int ci, int BN, int nchunks) {
  const int t = co / BN, n = co - t * BN, c = ci >> 6, kk = ci & 63;
  return (static_cast<size_t>(t) * nchunks + c) * BN * 128 + static_cast<size_t>(n) * 128 + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
}
__global__ void birnn_repack_kernel(const float* __restrict__ w0, const float* __restrict__ w1, int GH, int D, int ndir,
                                    uint8_t* __restrict__ proj, int bn1, int nch1, uint8_t* __restrict__ proj3, int bn3, int nch3) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(ndir) * GH * D) return;
  const int co = static_cast<int>(idx / D), ci = static_cast<int>(idx - static_cast<size_t>(co) * D);
  const int d = co / GH;
  const float w = (d == 0 ? w0 : w1)[static_cast<size_t>(co - d * GH) * D + ci];
  const __nv_bfloat16 hi = __float2bfloat16(w);
  const __nv_bfloat16 lo = __float2bfloat16(w - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(proj + wpack_offset(co, ci, bn1, nch1)) = hi;
  *reinterpret_cast<__nv_bfloat16*>(proj3 + wpack_offset(co, ci, bn3, nch3)) = hi;
  *reinterpret_cast<__nv_bfloat16*>(proj3 + wpack_offset(co, D + ci, bn3, nch3)) = hi;
  *reinterpret_cast<__nv_bfloat16*>(proj3 + wpack_offset(co, 2 * D + ci, bn3, nch3)) = lo;
}
__global__ void birnn_small_update_kernel(const float* __restrict__ whh0, const float* __restrict__ whh1, const float* __restrict__ bi0,
                                          const float* __restrict__ bi1, const float* __restrict__ bh0, const float* __restrict__ bh1,
                                          int GH, int H, int ndir, float* __restrict__ WhhT, float* __restrict__ bih,
                                          float* __restrict__ bhh) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(ndir) * GH * H) return;
  const int d = static_cast<int>(idx / (static_cast<size_t>(GH) * H));
  const int rem = static_cast<int>(idx - static_cast<size_t>(d) * GH * H);
  const int j = rem / H, k = rem - j * H;
  WhhT[(static_cast<size_t>(d) * H + k) * GH + j] = (d == 0 ? whh0 : whh1)[rem];
  if (k == 0) {
    bih[d * GH + j] = (d == 0 ? bi0 : bi1)[j];
    bhh[d * GH + j] = (d == 0 ? bh0 : bh1)[j];
  }
}
}  // namespace

extern "C" {

int tn_birnn_update_weights(tn_birnn_t* r, const float* const* i2h_weight, const float* const* h2h_weight,
                            const float* const* i2h_bias, const float* const* h2h_bias, tn_stream_t stream) {
  if (!r || !i2h_weight || !h2h_weight || !i2h_bias || !h2h_bias) return tn::set_error(TN_ERR_INVALID, "null argument");
  for (int d = 0; d < r->ndir; ++d)
    if (!i2h_weight[d] || !h2h_weight[d] || !i2h_bias[d] || !h2h_bias[d]) return tn::set_error(TN_ERR_INVALID, "null device weight");
  const int GH = r->gates * r->H, D = r->D, H = r->H, nd = r->ndir;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n1 = static_cast<size_t>(nd) * GH * D, n2 = static_cast<size_t>(nd) * GH * H;
  birnn_repack_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, st>>>(
      i2h_weight[0], nd > 1 ? i2h_weight[1] : nullptr, GH, D, nd, const_cast<uint8_t*>(r->proj.wpack),
      tn::conv_gemm_pick_bn(r->proj.Cout), r->proj.num_chunks, const_cast<uint8_t*>(r->proj3.wpack),
      tn::conv_gemm_pick_bn(r->proj3.Cout), r->proj3.num_chunks);
  birnn_small_update_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, st>>>(
      h2h_weight[0], nd > 1 ? h2h_weight[1] : nullptr, i2h_bias[0], nd > 1 ? i2h_bias[1] : nullptr, h2h_bias[0],
      nd > 1 ? h2h_bias[1] : nullptr, GH, H, nd, const_cast<float*>(r->WhhT), const_cast<float*>(r->bih), const_cast<float*>(r->bhh));
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

size_t tn_birnn_workspace_bytes(const tn_birnn_t* r, int B, int T) {
  if (!r || B < 0 || T < 0) return 0;
  const size_t M = static_cast<size_t>(B) * T;
  return tn::align_up(M * 3 * r->D * sizeof(__nv_bfloat16), 1024) + tn::align_up(M * r->ndir * r->gates * r->H * sizeof(float), 1024) + 1024;
}

int tn_birnn_forward(tn_birnn_t* r, const void* x, int x_is_bf16, const int32_t* valid_len, int B, int T, float* y,
                     float* ymax, float* h_final, float* c_final, void* workspace, size_t workspace_bytes,
                     tn_stream_t stream) {
  return birnn_forward_impl(r, x, x_is_bf16, valid_len, B, T, y, ymax, h_final, c_final, nullptr, nullptr, workspace,
                            workspace_bytes, stream);
}

// Forward that also keeps what the backward pass needs: gx (B*T, ndir*G*H) input-projection pre-activations and, for
// LSTM, the cell state per step cseq (B,T,ndir*H).  y must be given.
int tn_birnn_forward_train(tn_birnn_t* r, const void* x, int x_is_bf16, int B, int T, float* y, float* ymax, float* gx,
                           float* cseq, void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!y || !gx) return tn::set_error(TN_ERR_INVALID, "y and gx are required for training");
  return birnn_forward_impl(r, x, x_is_bf16, nullptr, B, T, y, ymax, nullptr, nullptr, gx, cseq, workspace, workspace_bytes, stream);
}

}  // extern "C"

static int birnn_forward_impl(tn_birnn_t* r, const void* x, int x_is_bf16, const int32_t* valid_len, int B, int T, float* y,
                              float* ymax, float* h_final, float* c_final, float* gx_out, float* cseq, void* workspace,
                              size_t workspace_bytes, tn_stream_t stream) {
  if (!r || B < 0 || T < 0) return tn::set_error(TN_ERR_INVALID, "bad rnn handle / shape");
  if (B == 0 || T == 0) return TN_OK;
  if (!x || !workspace) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  if (workspace_bytes < tn_birnn_workspace_bytes(r, B, T))
    return tn::set_error(TN_ERR_WORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, tn_birnn_workspace_bytes(r, B, T));
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return tn::set_error(TN_ERR_INVALID, "workspace must be 1024-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t M = static_cast<size_t>(B) * T;
  const int N = r->ndir * r->gates * r->H;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws);
  float* gx = gx_out ? gx_out : reinterpret_cast<float*>(ws + tn::align_up(M * 3 * r->D * sizeof(__nv_bfloat16), 1024));
  const __nv_bfloat16* xin = static_cast<const __nv_bfloat16*>(x);
  const bool precise = r->precise && !x_is_bf16;
  const int Dk = precise ? 3 * r->D : r->D;
  const tn::ConvDev& proj = precise ? r->proj3 : r->proj;
  if (precise) {
    TN_CUDA(tn::launch_split3_bf16(static_cast<const float*>(x), xb, M, r->D, st));
    xin = xb;
  } else if (!x_is_bf16) {
    TN_CUDA(tn::launch_cast_bf16(static_cast<const float*>(x), xb, M * r->D, st));
    xin = xb;
  }
  // input projection for every (b,t) and both directions: one tensor-core GEMM, bias in the epilogue
  tn::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.in = xin;
  p.in_cstride = Dk;
  p.H = 1;
  p.W = 1;
  p.Cin = Dk;
  p.Ho = 1;
  p.Wo = 1;
  p.R = 1;
  p.S = 1;
  p.stride = 1;
  p.pad = 0;
  p.mode = tn::kModeConv;
  p.wpack = proj.wpack;
  p.num_chunks = proj.num_chunks;
  p.chunks_per_tap = proj.chunks_per_tap;
  p.out = gx;
  p.out_cstride = N;
  p.out_coff = 0;
  p.out_fp32 = 1;
  p.Cout = N;
  p.epi_shift = r->bih;
  p.M = static_cast<int>(M);
  TN_CUDA(tn::launch_conv_gemm(p, st));

  if (y && valid_len) TN_CUDA(cudaMemsetAsync(y, 0, M * r->ndir * r->H * sizeof(float), st));
  tn::RnnScanParams s;
  memset(&s, 0, sizeof(s));
  s.gates = r->gates;
  s.B = B;
  s.T = T;
  s.H = r->H;
  s.ndir = r->ndir;
  s.reverse_dir1 = 1;
  s.gx = gx;
  s.WhhT = r->WhhT;
  s.bhh = r->bhh;
  s.valid_len = valid_len;
  s.y = y;
  s.cseq = cseq;
  s.ymax = ymax;
  s.h_final = h_final;
  s.c_final = c_final;
  TN_CUDA(tn::launch_rnn_scan(s, st));
  return TN_OK;
}
