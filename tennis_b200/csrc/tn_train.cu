// Training step of the temporal head (the published "features -> BiGRU -> max -> Dense" setting with
// a frozen / pre-extracted backbone): softmax cross-entropy, Dense backward, max-over-time + BPTT through the fused
// (bi)GRU / (bi)LSTM layer, weight gradients, SGD-momentum and Adam updates (A.8).  All fp32.
//
//   rnn_bwd_kernel   reverse scan.  Per step: recompute hh = W_hh h_{t-1} + b_hh (coalesced reads of W_hh^T), gate
//                    derivatives, write d(gx) and d(hh) pre-activation gradients, dh_{t-1} = z*dh + W_hh^T-contract.
//   gemm_tn          dW = A^T B over the (b,t) rows (shared fp32 SGEMM): dW_ih = dgx^T X, dW_hh = dhh^T H_prev.
#include <math.h>
#include <string.h>

#include "tn_common.h"
#include "tn_rnn.h"

struct tn_birnn;  // defined in tn_birnn.cu
namespace tn {
// accessors implemented in tn_birnn.cu
void birnn_dims(const tn_birnn* r, int* G, int* D, int* H, int* ndir);
const float* birnn_whhT(const tn_birnn* r);
const float* birnn_bhh(const tn_birnn* r);
}  // namespace tn

namespace {

using namespace tn;

constexpr int kRB = 4;  // batch rows per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct RnnBwdParams {
  int B, T, H, ndir;
  const float* gx;     // [B*T][ndir*G*H]   x W_ih^T + b_ih (saved by the forward)
  const float* y;      // [B][T][ndir*H]    h_t
  const float* cseq;   // [B][T][ndir*H]    c_t (LSTM)
  const float* ymax;   // [B][ndir*H]       max_t h_t (nullable)
  const float* d_ymax; // [B][ndir*H]       gradient w.r.t. the max-pooled output (nullable)
  const float* dy;     // [B][T][ndir*H]    gradient w.r.t. every h_t (nullable)
  const float* WhhT;   // [ndir][H][G*H]
  const float* bhh;    // [ndir][G*H]
  float* dgx;          // [B*T][ndir*G*H]
  float* dhh;          // [B*T][ndir*G*H]   gradient w.r.t. (W_hh h_{t-1} + b_hh)
  float* hprev;        // [B*T][ndir*H]     h_{t-1} in scan order (input of dW_hh)
};

template <int G>
__global__ void __launch_bounds__(512) rnn_bwd_kernel(const RnnBwdParams p) {
  extern __shared__ float sm[];
  const int H = p.H, GH = G * H;
  float* s_hprev = sm;                 // [kRB][H]
  float* s_hh = s_hprev + kRB * H;     // [kRB][GH]
  float* s_dhh = s_hh + kRB * GH;      // [kRB][GH]
  float* s_dh = s_dhh + kRB * GH;      // [kRB][H]   dL/dh_t carried backwards
  float* s_dc = s_dh + kRB * H;        // [kRB][H]   dL/dc_t (LSTM)
  const int tid = threadIdx.x, nth = blockDim.x;
  const int dir = blockIdx.y, b0 = blockIdx.x * kRB;
  const float* WhhT = p.WhhT + static_cast<size_t>(dir) * H * GH;
  const float* bhh = p.bhh + static_cast<size_t>(dir) * GH;
  const int ystride = p.ndir * H, gstride = p.ndir * GH;
  for (int i = tid; i < kRB * H; i += nth) {
    s_dh[i] = 0.f;
    s_dc[i] = 0.f;
  }
  __syncthreads();
  for (int s = p.T - 1; s >= 0; --s) {
    const int pos = (dir == 0) ? s : p.T - 1 - s;             // time index of scan step s
    const int ppos = (dir == 0) ? pos - 1 : pos + 1;          // time index of the previous scan step
    // stage h_{t-1}; add this step's output gradients into dh
    for (int i = tid; i < kRB * H; i += nth) {
      const int b = i / H, u = i - b * H;
      float hp = 0.f, add = 0.f;
      if (b0 + b < p.B) {
        const size_t row = static_cast<size_t>(b0 + b) * p.T;
        if (s > 0) hp = p.y[(row + ppos) * ystride + dir * H + u];
        const float yv = p.y[(row + pos) * ystride + dir * H + u];
        if (p.dy) add += p.dy[(row + pos) * ystride + dir * H + u];
        if (p.d_ymax) {
          // F.max backward: the gradient goes to the arg-max position (first occurrence along time)
          const float mx = p.ymax[static_cast<size_t>(b0 + b) * ystride + dir * H + u];
          if (yv == mx) {
            bool first = true;
            for (int t2 = 0; t2 < pos && first; ++t2) first = p.y[(row + t2) * ystride + dir * H + u] != mx;
            if (first) add += p.d_ymax[static_cast<size_t>(b0 + b) * ystride + dir * H + u];
          }
        }
        p.hprev[(row + pos) * ystride + dir * H + u] = hp;
      }
      s_hprev[i] = hp;
      s_dh[i] += add;
    }
    __syncthreads();
    // hh[b][j] = sum_k WhhT[k][j] h_prev[b][k] + bhh[j]
    for (int j = tid; j < GH; j += nth) {
      float acc[kRB];
#pragma unroll
      for (int b = 0; b < kRB; ++b) acc[b] = 0.f;
      for (int k = 0; k < H; ++k) {
        const float w = __ldg(WhhT + static_cast<size_t>(k) * GH + j);
#pragma unroll
        for (int b = 0; b < kRB; ++b) acc[b] = fmaf(w, s_hprev[b * H + k], acc[b]);
      }
      const float bj = bhh[j];
#pragma unroll
      for (int b = 0; b < kRB; ++b) s_hh[b * GH + j] = acc[b] + bj;
    }
    __syncthreads();
    // gate derivatives
    for (int i = tid; i < kRB * H; i += nth) {
      const int b = i / H, u = i - b * H;
      if (b0 + b >= p.B) {
#pragma unroll
        for (int g = 0; g < G; ++g) s_dhh[b * GH + g * H + u] = 0.f;
        continue;
      }
      const size_t grow = (static_cast<size_t>(b0 + b) * p.T + pos) * gstride + dir * GH;
      const float dh = s_dh[i];
      float dgx[G], dhh[G], dh_direct;
      if (G == 3) {
        const float hh_r = s_hh[b * GH + u], hh_z = s_hh[b * GH + H + u], hh_n = s_hh[b * GH + 2 * H + u];
        const float r = sigmoidf_(p.gx[grow + u] + hh_r);
        const float z = sigmoidf_(p.gx[grow + H + u] + hh_z);
        const float n = tanhf(p.gx[grow + 2 * H + u] + r * hh_n);
        const float hp = s_hprev[i];
        const float dn_pre = dh * (1.f - z) * (1.f - n * n);
        const float dz_pre = dh * (hp - n) * z * (1.f - z);
        const float dr_pre = dn_pre * hh_n * r * (1.f - r);
        dgx[0] = dr_pre; dgx[1] = dz_pre; dgx[2] = dn_pre;
        dhh[0] = dr_pre; dhh[1] = dz_pre; dhh[2] = dn_pre * r;
        dh_direct = dh * z;
      } else {
        const size_t crow = (static_cast<size_t>(b0 + b) * p.T + pos) * ystride + dir * H + u;
        const float ig = sigmoidf_(p.gx[grow + u] + s_hh[b * GH + u]);
        const float fg = sigmoidf_(p.gx[grow + H + u] + s_hh[b * GH + H + u]);
        const float gg = tanhf(p.gx[grow + 2 * H + u] + s_hh[b * GH + 2 * H + u]);
        const float og = sigmoidf_(p.gx[grow + (G - 1) * H + u] + s_hh[b * GH + (G - 1) * H + u]);
        const float ct = p.cseq[crow];
        const float cp = (s > 0) ? p.cseq[(static_cast<size_t>(b0 + b) * p.T + ppos) * ystride + dir * H + u] : 0.f;
        const float tc = tanhf(ct);
        const float dc = s_dc[i] + dh * og * (1.f - tc * tc);
        dgx[0] = dc * gg * ig * (1.f - ig);
        dgx[1] = dc * cp * fg * (1.f - fg);
        dgx[2] = dc * ig * (1.f - gg * gg);
        dgx[G - 1] = dh * tc * og * (1.f - og);
#pragma unroll
        for (int g = 0; g < G; ++g) dhh[g] = dgx[g];
        s_dc[i] = dc * fg;
        dh_direct = 0.f;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        p.dgx[grow + g * H + u] = dgx[g];
        p.dhh[grow + g * H + u] = dhh[g];
        s_dhh[b * GH + g * H + u] = dhh[g];
      }
      s_dh[i] = dh_direct;  // recurrent part added below
    }
    __syncthreads();
    // dh_{t-1}[b][k] += sum_j dhh[b][j] W_hh[j][k] = sum_j dhh[b][j] WhhT[k][j]
    for (int i = tid; i < kRB * H; i += nth) {
      const int b = i / H, k = i - b * H;
      const float* w = WhhT + static_cast<size_t>(k) * GH;
      const float* d = s_dhh + b * GH;
      float acc = 0.f;
      for (int j = 0; j < GH; ++j) acc = fmaf(d[j], __ldg(w + j), acc);
      s_dh[i] += acc;
    }
    __syncthreads();
  }
}

// out[j] = sum_m A[m*lda + off + j]: one block per 32 columns, 8 row lanes, shared-memory reduction
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int lda, int off, float* __restrict__ out, int M, int N) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (j < N)
    for (int m = ty; m < M; m += 8) acc += A[static_cast<size_t>(m) * lda + off + j];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < N) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][tx];
    out[j] = s;
  }
}

__global__ void cast_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}

// per-sample softmax cross-entropy and its gradient w.r.t. the logits (head gradient = 1 per sample, A.8)
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int* __restrict__ labels, float* __restrict__ loss,
                                  float* __restrict__ dlogits, int Bn, int C) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bn) return;
  const float* x = logits + static_cast<size_t>(b) * C;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, x[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(x[c] - m);
  const float lse = m + logf(s);
  const int y = labels[b];
  if (loss) loss[b] = lse - x[y];
  if (dlogits)
    for (int c = 0; c < C; ++c) dlogits[static_cast<size_t>(b) * C + c] = expf(x[c] - lse) - (c == y ? 1.f : 0.f);
}

// gluonnlp MaskedSoftmaxCELoss (A.8): per-token CE x SequenceMask(valid_len) weights, MEAN over the padded length T -> (B,)
__global__ void __launch_bounds__(128) masked_ce_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                        const float* __restrict__ vlen, float* __restrict__ loss, int T, int V) {
  __shared__ float red[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = min(max(static_cast<int>(vlen[b]), 0), T);
  float total = 0.f;
  for (int t = warp; t < len; t += 4) {  // one warp per valid token
    const float* x = pred + (static_cast<size_t>(b) * T + t) * V;
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, x[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(x[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int y = min(max(static_cast<int>(lrintf(label[static_cast<size_t>(b) * T + t])), 0), V - 1);
    total += (m + logf(s)) - x[y];
  }
  if (lane == 0) red[warp] = total;
  __syncthreads();
  if (tid == 0) loss[b] = (red[0] + red[1] + red[2] + red[3]) / static_cast<float>(T);
}

// Dense backward: dW[j][k] = sum_r dy[r][j] x[r][k]; db[j] = sum_r dy[r][j]; dx[r][k] = sum_j dy[r][j] W[j][k]
__global__ void dense_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dW,
                                   float* __restrict__ db, int R, int in_dim, int out_dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out_dim * in_dim) return;
  const int j = idx / in_dim, k = idx - j * in_dim;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc = fmaf(dy[static_cast<size_t>(r) * out_dim + j], x[static_cast<size_t>(r) * in_dim + k], acc);
  dW[idx] = acc;
  if (k == 0 && db) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += dy[static_cast<size_t>(r) * out_dim + j];
    db[j] = s;
  }
}
__global__ void dense_bwd_x_kernel(const float* __restrict__ W, const float* __restrict__ dy, float* __restrict__ dx, int R,
                                   int in_dim, int out_dim) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(R) * in_dim) return;
  const size_t r = idx / in_dim;
  const int k = static_cast<int>(idx - r * in_dim);
  float acc = 0.f;
  for (int j = 0; j < out_dim; ++j) acc = fmaf(dy[r * out_dim + j], W[static_cast<size_t>(j) * in_dim + k], acc);
  dx[idx] = acc;
}

// gluon Trainer.step(n) with 'sgd': g' = rescale*g + wd*w ; m = mu*m - lr*g' ; w += m       (A.8)
__global__ void sgd_mom_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ mom, size_t n, float lr,
                               float mu, float wd, float rescale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gp = rescale * g[i] + wd * w[i];
  const float m = mu * mom[i] - lr * gp;
  mom[i] = m;
  w[i] += m;
}
// 'adam': g' = rescale*g + wd*w; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; w -= lr_t m / (sqrt(v) + eps), lr_t bias-corrected
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, float b1, float b2, float eps, float wd, float rescale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gp = rescale * g[i] + wd * w[i];
  const float mi = b1 * m[i] + (1.f - b1) * gp;
  const float vi = b2 * v[i] + (1.f - b2) * gp * gp;
  m[i] = mi;
  v[i] = vi;
  w[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// C (N1 x N2) = A[:, a_off : a_off+N1]^T  B[:, b_off : b_off+N2]  over M rows: the shared fp32 SGEMM (tn_seq_train.cu)
int gemm_tn(const float* A, int lda, int a_off, const float* Bm, int ldb, int b_off, float* C, int ldc, int M, int N1, int N2,
            cudaStream_t st) {
  return tn_sgemm(1, 0, N1, N2, M, 1.0f, A + a_off, lda, Bm + b_off, ldb, 0.0f, C, ldc, st);
}

}  // namespace

extern "C" {

int tn_softmax_ce(const float* logits, const int32_t* labels, float* loss, float* dlogits, int B, int C, tn_stream_t stream) {
  if (B < 0 || C <= 0) return set_error(TN_ERR_INVALID, "bad shape");
  if (B == 0) return TN_OK;
  if (!logits || !labels) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(kProfOther, st);
  softmax_ce_kernel<<<(B + 127) / 128, 128, 0, st>>>(logits, labels, loss, dlogits, B, C);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_masked_softmax_ce(const float* pred, const float* label, const float* valid_len, float* loss, int B, int T, int V,
                         tn_stream_t stream) {
  if (B < 0 || T <= 0 || V <= 0) return set_error(TN_ERR_INVALID, "bad shape");
  if (B == 0) return TN_OK;
  if (!pred || !label || !valid_len || !loss) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(kProfOther, st);
  masked_ce_kernel<<<B, 128, 0, st>>>(pred, label, valid_len, loss, T, V);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_dense_backward(const float* x, const float* weight, const float* dy, float* dx, float* dweight, float* dbias, int rows,
                      int in_dim, int out_dim, tn_stream_t stream) {
  if (rows < 0 || in_dim <= 0 || out_dim <= 0) return set_error(TN_ERR_INVALID, "bad shape");
  if (!x || !dy || !dweight) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope ps(kProfOther, st);
    dense_bwd_w_kernel<<<(out_dim * in_dim + 255) / 256, 256, 0, st>>>(x, dy, dweight, dbias, rows, in_dim, out_dim);
  }
  if (dx && rows > 0) {
    if (!weight) return set_error(TN_ERR_INVALID, "weight needed for dx");
    ProfScope ps(kProfOther, st);
    dense_bwd_x_kernel<<<static_cast<unsigned>((static_cast<size_t>(rows) * in_dim + 255) / 256), 256, 0, st>>>(weight, dy, dx, rows,
                                                                                                             in_dim, out_dim);
  }
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

size_t tn_birnn_backward_workspace_bytes(const tn_birnn_t* r, int B, int T) {
  if (!r) return 0;
  int G, D, H, ndir;
  birnn_dims(r, &G, &D, &H, &ndir);
  const size_t M = static_cast<size_t>(B) * T;
  return (2 * M * ndir * G * H + M * ndir * H + M * D) * sizeof(float) + 4096;
}

// Backward of tn_birnn_forward for zero initial state and no valid_len (the CNNRNN head, definitions.py:103-110).
//   x (B,T,D) as given to the forward; gx / y / cseq saved by tn_birnn_forward_train; d_ymax and/or dy: output gradients.
//   gradients, concatenated over directions like the forward's stacking: dW_ih (ndir*G*H, D), dW_hh (ndir*G*H, H),
//   db_ih, db_hh (ndir*G*H).
int tn_birnn_backward(tn_birnn_t* r, const void* x, int x_is_bf16, int B, int T, const float* gx, const float* y, const float* cseq,
                      const float* ymax, const float* d_ymax, const float* dy, float* dW_ih, float* dW_hh, float* db_ih,
                      float* db_hh, void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!r || B <= 0 || T <= 0) return set_error(TN_ERR_INVALID, "bad arguments");
  int G, D, H, ndir;
  birnn_dims(r, &G, &D, &H, &ndir);
  if (!x || !gx || !y || !dW_ih || !dW_hh || !db_ih || !db_hh || !workspace || (G == 4 && !cseq) || (!d_ymax && !dy) || (d_ymax && !ymax))
    return set_error(TN_ERR_INVALID, "null pointer");
  if (workspace_bytes < tn_birnn_backward_workspace_bytes(r, B, T)) return set_error(TN_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t M = static_cast<size_t>(B) * T;
  const int GH = G * H;
  float* dgx = static_cast<float*>(workspace);
  float* dhh = dgx + M * ndir * GH;
  float* hprev = dhh + M * ndir * GH;
  float* xf = hprev + M * ndir * H;
  const float* xin = static_cast<const float*>(x);
  if (x_is_bf16) {
    cast_bf16_to_f32_kernel<<<static_cast<unsigned>((M * D + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), xf, M * D);
    xin = xf;
  }
  RnnBwdParams p;
  p.B = B; p.T = T; p.H = H; p.ndir = ndir;
  p.gx = gx; p.y = y; p.cseq = cseq; p.ymax = ymax; p.d_ymax = d_ymax; p.dy = dy;
  p.WhhT = birnn_whhT(r); p.bhh = birnn_bhh(r);
  p.dgx = dgx; p.dhh = dhh; p.hprev = hprev;
  const size_t smem = (static_cast<size_t>(kRB) * (3 * H + 2 * GH)) * sizeof(float);
  dim3 grid((B + kRB - 1) / kRB, ndir);
  {
    ProfScope ps(kProfOther, st);
    if (G == 3) rnn_bwd_kernel<3><<<grid, 512, smem, st>>>(p);
    else rnn_bwd_kernel<4><<<grid, 512, smem, st>>>(p);
  }
  TN_CUDA(cudaGetLastError());
  for (int d = 0; d < ndir; ++d) {
    int rc = gemm_tn(dgx, ndir * GH, d * GH, xin, D, 0, dW_ih + static_cast<size_t>(d) * GH * D, D, static_cast<int>(M), GH, D, st);
    if (rc != TN_OK) return rc;
    rc = gemm_tn(dhh, ndir * GH, d * GH, hprev, ndir * H, d * H, dW_hh + static_cast<size_t>(d) * GH * H, H, static_cast<int>(M), GH, H, st);
    if (rc != TN_OK) return rc;
  }
  colsum_kernel<<<(ndir * GH + 31) / 32, 256, 0, st>>>(dgx, ndir * GH, 0, db_ih, static_cast<int>(M), ndir * GH);
  colsum_kernel<<<(ndir * GH + 31) / 32, 256, 0, st>>>(dhh, ndir * GH, 0, db_hh, static_cast<int>(M), ndir * GH);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_sgd_mom_update(float* weight, const float* grad, float* mom, size_t n, float lr, float momentum, float wd, float rescale_grad,
                      tn_stream_t stream) {
  if (n == 0) return TN_OK;
  if (!weight || !grad || !mom) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(kProfOther, st);
  sgd_mom_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(weight, grad, mom, n, lr, momentum, wd, rescale_grad);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_adam_update(float* weight, const float* grad, float* mean, float* var, size_t n, float lr, float beta1, float beta2, float eps,
                   float wd, float rescale_grad, int t, tn_stream_t stream) {
  if (n == 0) return TN_OK;
  if (!weight || !grad || !mean || !var || t < 1) return set_error(TN_ERR_INVALID, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double c1 = 1.0 - pow(static_cast<double>(beta1), t), c2 = 1.0 - pow(static_cast<double>(beta2), t);
  const float lr_t = static_cast<float>(lr * sqrt(c2) / c1);
  ProfScope ps(kProfOther, st);
  adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(weight, grad, mean, var, n, lr_t, beta1, beta2, eps, wd, rescale_grad);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

}  // extern "C"
