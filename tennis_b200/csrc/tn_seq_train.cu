// Training kernels of the GNMT captioner (SURVEY.md §8a G1-G5 forward with saved activations, G9 backward):
// everything `train_gnmt.py:330-337` differentiates through -- GRU/LSTM cells unrolled with valid_length
// (gnmt.py:143-145, Appendix A.3/A.4), scaled-Luong attention (A.5), target embedding, MaskedSoftmaxCELoss (A.8) and
// output dropout (gnmt.py:152,389).  All fp32: the captioner is tiny (H = 128, B = 128) and its gradients are compared with
// autograd of the CPU oracle at 1e-4, so these are plain SIMT kernels; the Python side (models/captioning/train_graph.py)
// sequences them step by step.  Matrices are row-major with explicit row strides so that a time step of a (B, T, C) tensor
// is addressed in place (row stride T*C), never copied.
#include <math.h>

#include "tn_common.h"

namespace {

using namespace tn;

// ---------------------------------------------------------------------------------------------------- SGEMM
// C[M,N] = alpha * op(A) * op(B) + beta * C.  op(A) is M x K: A[m*lda + k] (TA = 0) or A[k*lda + m] (TA = 1);
// op(B) is K x N: B[k*ldb + n] (TB = 0) or B[n*ldb + k] (TB = 1).  64 x 64 tile, K step 16, 4 x 4 outputs per thread.
template <int TA, int TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, float beta, float* __restrict__ C,
                                                    int ldc) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  // The next K tile is fetched into registers while the current one is multiplied: the per-step GEMMs of the recurrences run
  // on a handful of CTAs, where an unprefetched loop pays the full global-load latency every 16 columns of K.
  // split-K (gridDim.z > 1, beta == 1 only): this CTA reduces K range [kbeg, kend) and adds its partial sum atomically
  const int kchunk = ((K + static_cast<int>(gridDim.z) - 1) / static_cast<int>(gridDim.z) + 15) / 16 * 16;
  const int kbeg = blockIdx.z * kchunk;
  const int kend = min(K, kbeg + kchunk);
  if (kbeg >= kend) return;
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * 256;
      int k, m;
      if (TA) { k = i >> 6; m = i & 63; } else { m = i >> 4; k = i & 15; }
      const int gm = m0 + m;
      int gk = k0 + k;
      ra[q] = (gm < M && gk < kend) ? (TA ? A[static_cast<size_t>(gk) * lda + gm] : A[static_cast<size_t>(gm) * lda + gk]) : 0.f;
      int n;
      if (TB) { n = i >> 4; k = i & 15; } else { k = i >> 6; n = i & 63; }
      const int gn = n0 + n;
      gk = k0 + k;
      rb[q] = (gn < N && gk < kend) ? (TB ? B[static_cast<size_t>(gn) * ldb + gk] : B[static_cast<size_t>(gk) * ldb + gn]) : 0.f;
    }
  };
  fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * 256;
      if (TA) As[i >> 6][i & 63] = ra[q]; else As[i & 15][i >> 4] = ra[q];
      if (TB) Bs[i & 15][i >> 4] = rb[q]; else Bs[i >> 6][i & 63] = rb[q];
    }
    __syncthreads();
    if (k0 + 16 < kend) fetch(k0 + 16);
    // two-level accumulation: each 16-deep K tile is summed on its own and then added to the running total, so the rounding
    // error grows with K/16 + 16 instead of K (the CNN training path reduces over up to 4608 products per output)
    float part[4][4] = {};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* c = C + static_cast<size_t>(gm) * ldc + gn;
      if (gridDim.z > 1) atomicAdd(c, alpha * acc[i][j]);
      else *c = (beta == 0.f) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *c;
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------- RNN cell, one step
// gi = x W_i2h^T, gh = h_prev W_h2h^T (no biases), rows strided.  save: LSTM [i f g o] (4H), GRU [r z n gh_n+b_hn] (4H).
struct CellFwd {
  int cell, B, H;
  const float* gi; long long gi_stride;
  const float* gh; long long gh_stride;
  const float* bi; const float* bh;
  const float* h_prev; long long hp_stride;   // null -> zeros
  const float* c_prev; long long cp_stride;   // LSTM; null -> zeros
  float* h_out; long long ho_stride;
  float* h_out2; long long ho2_stride;        // optional second copy of h (e.g. into the next layer's concat input)
  float* c_out; long long co_stride;
  float* save; long long sv_stride;
};

__global__ void cell_fwd_kernel(const CellFwd p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * p.H) return;
  const int b = idx / p.H, u = idx - b * p.H;
  const int H = p.H;
  const float* gi = p.gi + b * p.gi_stride;
  const float* gh = p.gh + b * p.gh_stride;
  const float hp = p.h_prev ? p.h_prev[b * p.hp_stride + u] : 0.f;
  float* sv = p.save + b * p.sv_stride;
  float h;
  if (p.cell == 0) {  // GRU [r, z, n]
    const float r = sigmoidf_(gi[u] + p.bi[u] + gh[u] + p.bh[u]);
    const float z = sigmoidf_(gi[H + u] + p.bi[H + u] + gh[H + u] + p.bh[H + u]);
    const float ghn = gh[2 * H + u] + p.bh[2 * H + u];
    const float n = tanhf(gi[2 * H + u] + p.bi[2 * H + u] + r * ghn);
    h = (1.f - z) * n + z * hp;
    sv[u] = r; sv[H + u] = z; sv[2 * H + u] = n; sv[3 * H + u] = ghn;
  } else {  // LSTM [i, f, g, o]
    const float cp = p.c_prev ? p.c_prev[b * p.cp_stride + u] : 0.f;
    const float i = sigmoidf_(gi[u] + p.bi[u] + gh[u] + p.bh[u]);
    const float f = sigmoidf_(gi[H + u] + p.bi[H + u] + gh[H + u] + p.bh[H + u]);
    const float g = tanhf(gi[2 * H + u] + p.bi[2 * H + u] + gh[2 * H + u] + p.bh[2 * H + u]);
    const float o = sigmoidf_(gi[3 * H + u] + p.bi[3 * H + u] + gh[3 * H + u] + p.bh[3 * H + u]);
    const float c = f * cp + i * g;
    h = o * tanhf(c);
    p.c_out[b * p.co_stride + u] = c;
    sv[u] = i; sv[H + u] = f; sv[2 * H + u] = g; sv[3 * H + u] = o;
  }
  p.h_out[b * p.ho_stride + u] = h;
  if (p.h_out2) p.h_out2[b * p.ho2_stride + u] = h;
}

// Backward of one step.  dh / dc: running gradients w.r.t. this step's (h, c) coming from step t+1 (in/out: overwritten with
// the DIRECT part of the gradient w.r.t. h_{t-1} / c_{t-1}; the caller adds dgh W_h2h).  dy: gradient w.r.t. this step's
// output (nullable).  valid_len/t: rows with t >= len contribute nothing (A.4: states are taken at len-1, outputs masked);
// at t == len-1 the final-state gradients dh_last / dc_last are added.  dgi / dgh: gate pre-activation gradients (LSTM: equal).
struct CellBwd {
  int cell, B, H, t;
  const int* valid_len;
  const float* save; long long sv_stride;
  const float* h_prev; long long hp_stride;
  const float* c_prev; long long cp_stride;
  const float* c_cur; long long cc_stride;
  const float* dy; long long dy_stride;
  const float* dy2; long long dy2_stride;     // second output-gradient source (h_out2's consumer), nullable
  const float* dh_last; const float* dc_last; // (B,H) contiguous, nullable
  float* dh; float* dc;                        // (B,H) contiguous
  float* dgi; long long dgi_stride;
  float* dgh; long long dgh_stride;            // GRU only (LSTM: null)
};

__global__ void cell_bwd_kernel(const CellBwd p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * p.H) return;
  const int b = idx / p.H, u = idx - b * p.H;
  const int H = p.H;
  const int len = p.valid_len ? p.valid_len[b] : (1 << 30);
  float* dgi = p.dgi + b * p.dgi_stride;
  float* dgh = p.dgh ? p.dgh + b * p.dgh_stride : nullptr;
  const int G = p.cell == 0 ? 3 : 4;
  if (p.t >= len) {
    for (int g = 0; g < G; ++g) {
      dgi[g * H + u] = 0.f;
      if (dgh) dgh[g * H + u] = 0.f;
    }
    p.dh[idx] = 0.f;
    if (p.dc) p.dc[idx] = 0.f;
    return;
  }
  float dh = p.dh[idx];
  float dc = p.dc ? p.dc[idx] : 0.f;
  if (p.valid_len ? (p.t == len - 1) : false) {
    // anything flowing in from t+1 belongs to padded steps: drop it, take the final-state gradient instead
    dh = p.dh_last ? p.dh_last[idx] : 0.f;
    dc = p.dc_last ? p.dc_last[idx] : 0.f;
  }
  if (p.dy) dh += p.dy[b * p.dy_stride + u];
  if (p.dy2) dh += p.dy2[b * p.dy2_stride + u];
  const float* sv = p.save + b * p.sv_stride;
  const float hp = p.h_prev ? p.h_prev[b * p.hp_stride + u] : 0.f;
  if (p.cell == 0) {
    const float r = sv[u], z = sv[H + u], n = sv[2 * H + u], ghn = sv[3 * H + u];
    const float dn = dh * (1.f - z);
    const float dz = dh * (hp - n);
    const float dn_pre = dn * (1.f - n * n);
    const float dr_pre = dn_pre * ghn * r * (1.f - r);
    const float dz_pre = dz * z * (1.f - z);
    dgi[u] = dr_pre; dgi[H + u] = dz_pre; dgi[2 * H + u] = dn_pre;
    dgh[u] = dr_pre; dgh[H + u] = dz_pre; dgh[2 * H + u] = dn_pre * r;
    p.dh[idx] = dh * z;
  } else {
    const float i = sv[u], f = sv[H + u], g = sv[2 * H + u], o = sv[3 * H + u];
    const float cp = p.c_prev ? p.c_prev[b * p.cp_stride + u] : 0.f;
    const float tc = tanhf(p.c_cur[b * p.cc_stride + u]);
    const float dct = dc + dh * o * (1.f - tc * tc);
    dgi[u] = dct * g * i * (1.f - i);
    dgi[H + u] = dct * cp * f * (1.f - f);
    dgi[2 * H + u] = dct * i * (1.f - g * g);
    dgi[3 * H + u] = dh * tc * o * (1.f - o);
    p.dh[idx] = 0.f;
    p.dc[idx] = dct * f;
  }
}

// ---------------------------------------------------------------------------------------------------- attention (A.5)
// q (B,H) = projected query (unscaled); scores_t = (q/sqrt(H)) . mem[b,t]; masked softmax over t < len; ctx = sum w_t mem_t.
// One block per row.  ctx is written to up to two strided destinations (the concat inputs of the neighbouring cells).
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ q, long long q_stride, const float* __restrict__ mem,
                                                       const int* __restrict__ src_len, int T, int H, float* __restrict__ w,
                                                       float* __restrict__ ctx1, long long c1_stride, float* __restrict__ ctx2,
                                                       long long c2_stride) {
  extern __shared__ float sm[];  // q (H) | scores (T) | red (32)
  float* sq = sm;
  float* sc = sm + H;
  float* red = sc + T;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int len = src_len ? min(src_len[b], T) : T;
  const float scale = rsqrtf(static_cast<float>(H));
  for (int i = tid; i < H; i += 128) sq[i] = q[b * q_stride + i] * scale;
  __syncthreads();
  const float* mb = mem + static_cast<size_t>(b) * T * H;
  for (int t = wid; t < T; t += 4) {  // one warp per source position
    float s = 0.f;
    if (t < len) {
      for (int i = lane; i < H; i += 32) s = fmaf(sq[i], mb[static_cast<size_t>(t) * H + i], s);
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) sc[t] = (t < len) ? s : -1e18f;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < T; t += 128) m = fmaxf(m, sc[t]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int t = tid; t < T; t += 128) {
    const float e = expf(sc[t] - m);
    sc[t] = e;
    sum += e;
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  for (int t = tid; t < T; t += 128) {
    const float wt = (t < len) ? sc[t] * inv : 0.f;  // softmax * mask
    sc[t] = wt;
    w[static_cast<size_t>(b) * T + t] = wt;
  }
  __syncthreads();
  for (int i = tid; i < H; i += 128) {
    float a = 0.f;
    for (int t = 0; t < len; ++t) a = fmaf(sc[t], mb[static_cast<size_t>(t) * H + i], a);
    ctx1[b * c1_stride + i] = a;
    if (ctx2) ctx2[b * c2_stride + i] = a;
  }
}

// dctx = dctx1 + dctx2 (either nullable).  dw_t = dctx . mem_t; ds = w (dw - sum w dw); dq = (sum_t ds_t mem_t)/sqrt(H);
// dmem[b,t] += w_t dctx + ds_t q/sqrt(H).
__global__ void __launch_bounds__(128) attn_bwd_kernel(const float* __restrict__ q, long long q_stride, const float* __restrict__ mem,
                                                       const int* __restrict__ src_len, int T, int H, const float* __restrict__ w,
                                                       const float* __restrict__ d1, long long d1_stride,
                                                       const float* __restrict__ d2, long long d2_stride, float* __restrict__ dq,
                                                       long long dq_stride, float* __restrict__ dmem) {
  extern __shared__ float sm[];  // q (H) | dctx (H) | ds (T) | red (32)
  float* sq = sm;
  float* sd = sm + H;
  float* ds = sd + H;
  float* red = ds + T;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int len = src_len ? min(src_len[b], T) : T;
  const float scale = rsqrtf(static_cast<float>(H));
  for (int i = tid; i < H; i += 128) {
    sq[i] = q[b * q_stride + i] * scale;
    float d = 0.f;
    if (d1) d += d1[b * d1_stride + i];
    if (d2) d += d2[b * d2_stride + i];
    sd[i] = d;
  }
  __syncthreads();
  const float* mb = mem + static_cast<size_t>(b) * T * H;
  const float* wb = w + static_cast<size_t>(b) * T;
  float part = 0.f;
  for (int t = wid; t < len; t += 4) {
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s = fmaf(sd[i], mb[static_cast<size_t>(t) * H + i], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      ds[t] = s;  // dw_t for now
      part += wb[t] * s;
    }
  }
  if (lane == 0) red[wid] = part;
  __syncthreads();
  const float dot = red[0] + red[1] + red[2] + red[3];
  for (int t = tid; t < len; t += 128) ds[t] = wb[t] * (ds[t] - dot);
  __syncthreads();
  float* dmb = dmem + static_cast<size_t>(b) * T * H;
  for (int i = tid; i < H; i += 128) {
    float a = 0.f;
    const float di = sd[i], qi = sq[i];
    for (int t = 0; t < len; ++t) {
      const size_t o = static_cast<size_t>(t) * H + i;
      a = fmaf(ds[t], mb[o], a);
      dmb[o] += wb[t] * di + ds[t] * qi;
    }
    dq[b * dq_stride + i] = a * scale;
  }
}

// ---------------------------------------------------------------------------------------------------- embedding, loss, dropout
__global__ void embedding_bwd_kernel(const float* __restrict__ ids, const float* __restrict__ dy, long long dy_stride,
                                     float* __restrict__ dW, int N, int E, int V) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * E) return;
  const int r = idx / E, e = idx - r * E;
  const int v = static_cast<int>(ids[r]);
  if (v < 0 || v >= V) return;
  atomicAdd(dW + static_cast<size_t>(v) * E + e, dy[r * dy_stride + e]);
}

// MaskedSoftmaxCELoss forward + gradient: loss[b] = sum_{t<len} CE_bt / T; dpred[b,t,:] = g[b]/T (softmax - onehot) (t < len).
__global__ void __launch_bounds__(128) masked_ce_grad_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                             const float* __restrict__ valid_len, const float* __restrict__ g,
                                                             float* __restrict__ loss_tok, float* __restrict__ dpred, int T, int V) {
  __shared__ float red[4];
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row - b * T;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* x = pred + static_cast<size_t>(row) * V;
  float* d = dpred ? dpred + static_cast<size_t>(row) * V : nullptr;
  const bool valid = static_cast<float>(t) < valid_len[b];
  if (!valid) {
    if (d) for (int v = tid; v < V; v += 128) d[v] = 0.f;
    if (tid == 0) loss_tok[row] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int v = tid; v < V; v += 128) m = fmaxf(m, x[v]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int v = tid; v < V; v += 128) s += expf(x[v] - m);
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  s = red[0] + red[1] + red[2] + red[3];
  const float lse = m + logf(s);
  const int lab = static_cast<int>(label[row]);
  if (tid == 0) loss_tok[row] = (lse - x[lab]) / static_cast<float>(T);
  if (d) {
    const float gs = g[b] / static_cast<float>(T);
    for (int v = tid; v < V; v += 128) d[v] = gs * (expf(x[v] - lse) - (v == lab ? 1.f : 0.f));
  }
}

__global__ void rowsum_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = 0; c < cols; ++c) s += x[static_cast<size_t>(r) * cols + c];
  y[r] = s;
}

// Inverted dropout with a counter-based generator (one 64-bit mix per element): mask in {0, 1/(1-p)}.
__device__ __forceinline__ uint32_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 32);
}
__global__ void dropout_mask_kernel(float* __restrict__ mask, size_t n, float p, unsigned long long seed) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = (mix64(seed * 0x100000001B3ull + i) >> 8) * (1.0f / 16777216.0f);
  mask[i] = (u < p) ? 0.f : 1.f / (1.f - p);
}
// y = x * mask (elementwise); with seq_len: rows (b,t) with t >= len[b] become zero (SequenceMask, gnmt.py:157-159,298-301).
__global__ void mul_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask, const int* __restrict__ seq_len,
                                float* __restrict__ y, int B, int T, int C) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * T * C) return;
  const int row = static_cast<int>(i / C);
  const int b = row / T, t = row - b * T;
  float v = x[i];
  if (mask) v *= mask[i];
  if (seq_len && t >= seq_len[b]) v = 0.f;
  y[i] = v;
}
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

inline unsigned blocks_for(size_t n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

}  // namespace

extern "C" {

int tn_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
             float beta, float* C, int ldc, tn_stream_t stream) {
  if (M < 0 || N < 0 || K < 0) return set_error(TN_ERR_INVALID, "negative GEMM size");
  if (M == 0 || N == 0) return TN_OK;
  if (!A || !B || !C) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  // accumulating GEMMs with few output tiles and a long K (dh += dgh W_h2h, dW_h2h += dgh^T h, every step of a BPTT loop):
  // split K across CTAs so that more than a handful of SMs work on them
  if (beta == 1.f && grid.x * grid.y < 32 && K >= 256) {
    int splits = 64 / static_cast<int>(grid.x * grid.y);
    if (splits > K / 64) splits = K / 64;
    if (splits > 1) grid.z = splits;
  }
  ProfScope prof_scope(kProfOther, st);
  if (!transA && !transB) sgemm_kernel<0, 0><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!transA && transB) sgemm_kernel<0, 1><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (transA && !transB) sgemm_kernel<1, 0><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else sgemm_kernel<1, 1><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_rnn_cell_forward(int cell, int B, int H, const float* gi, long long gi_stride, const float* gh, long long gh_stride,
                        const float* bi, const float* bh, const float* h_prev, long long hp_stride, const float* c_prev,
                        long long cp_stride, float* h_out, long long ho_stride, float* h_out2, long long ho2_stride, float* c_out,
                        long long co_stride, float* save, long long sv_stride, tn_stream_t stream) {
  if (cell != TN_CELL_GRU && cell != TN_CELL_LSTM) return set_error(TN_ERR_INVALID, "unknown cell %d", cell);
  if (B <= 0 || H <= 0) return TN_OK;
  if (!gi || !gh || !bi || !bh || !h_out || !save || (cell == TN_CELL_LSTM && !c_out)) return set_error(TN_ERR_INVALID, "null device pointer");
  CellFwd p{cell, B, H, gi, gi_stride, gh, gh_stride, bi, bh, h_prev, hp_stride, c_prev, cp_stride, h_out, ho_stride,
            h_out2, ho2_stride, c_out, co_stride, save, sv_stride};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  cell_fwd_kernel<<<blocks_for(static_cast<size_t>(B) * H, 256), 256, 0, st>>>(p);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_rnn_cell_backward(int cell, int B, int H, int t, const int32_t* valid_len, const float* save, long long sv_stride,
                         const float* h_prev, long long hp_stride, const float* c_prev, long long cp_stride, const float* c_cur,
                         long long cc_stride, const float* dy, long long dy_stride, const float* dy2, long long dy2_stride,
                         const float* dh_last, const float* dc_last, float* dh, float* dc, float* dgi, long long dgi_stride,
                         float* dgh, long long dgh_stride, tn_stream_t stream) {
  if (cell != TN_CELL_GRU && cell != TN_CELL_LSTM) return set_error(TN_ERR_INVALID, "unknown cell %d", cell);
  if (B <= 0 || H <= 0) return TN_OK;
  if (!save || !dh || !dgi || (cell == TN_CELL_GRU && !dgh) || (cell == TN_CELL_LSTM && (!dc || !c_cur)))
    return set_error(TN_ERR_INVALID, "null device pointer");
  CellBwd p{cell, B, H, t, valid_len, save, sv_stride, h_prev, hp_stride, c_prev, cp_stride, c_cur, cc_stride, dy, dy_stride,
            dy2, dy2_stride, dh_last, dc_last, dh, cell == TN_CELL_LSTM ? dc : nullptr, dgi, dgi_stride,
            cell == TN_CELL_GRU ? dgh : nullptr, dgh_stride};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  cell_bwd_kernel<<<blocks_for(static_cast<size_t>(B) * H, 256), 256, 0, st>>>(p);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_attention_forward(const float* q, long long q_stride, const float* mem, const int32_t* src_len, int B, int T, int H,
                         float* w, float* ctx1, long long c1_stride, float* ctx2, long long c2_stride, tn_stream_t stream) {
  if (B <= 0) return TN_OK;
  if (!q || !mem || !w || !ctx1 || T <= 0 || H <= 0) return set_error(TN_ERR_INVALID, "bad attention arguments");
  const size_t smem = (static_cast<size_t>(H) + T + 32) * sizeof(float);
  if (smem > 48 * 1024) return set_error(TN_ERR_INVALID, "attention row (H=%d, T=%d) exceeds 48 KB of shared memory", H, T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  attn_fwd_kernel<<<B, 128, smem, st>>>(q, q_stride, mem, src_len, T, H, w, ctx1, c1_stride, ctx2, c2_stride);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_attention_backward(const float* q, long long q_stride, const float* mem, const int32_t* src_len, int B, int T, int H,
                          const float* w, const float* dctx1, long long d1_stride, const float* dctx2, long long d2_stride,
                          float* dq, long long dq_stride, float* dmem, tn_stream_t stream) {
  if (B <= 0) return TN_OK;
  if (!q || !mem || !w || !dq || !dmem || (!dctx1 && !dctx2) || T <= 0 || H <= 0) return set_error(TN_ERR_INVALID, "bad attention arguments");
  const size_t smem = (2 * static_cast<size_t>(H) + T + 32) * sizeof(float);
  if (smem > 48 * 1024) return set_error(TN_ERR_INVALID, "attention row (H=%d, T=%d) exceeds 48 KB of shared memory", H, T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  attn_bwd_kernel<<<B, 128, smem, st>>>(q, q_stride, mem, src_len, T, H, w, dctx1, d1_stride, dctx2, d2_stride, dq, dq_stride, dmem);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_embedding_backward(const float* ids, const float* dy, long long dy_stride, float* dweight, int N, int E, int V,
                          tn_stream_t stream) {
  if (N <= 0 || E <= 0) return TN_OK;
  if (!ids || !dy || !dweight) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  embedding_bwd_kernel<<<blocks_for(static_cast<size_t>(N) * E, 256), 256, 0, st>>>(ids, dy, dy_stride, dweight, N, E, V);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_masked_softmax_ce_grad(const float* pred, const float* label, const float* valid_len, const float* head_grad,
                              float* loss, float* dpred, float* workspace_bt, int B, int T, int V, tn_stream_t stream) {
  if (B <= 0 || T <= 0) return TN_OK;
  if (!pred || !label || !valid_len || !loss || !workspace_bt || (dpred && !head_grad)) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  masked_ce_grad_kernel<<<B * T, 128, 0, st>>>(pred, label, valid_len, head_grad, workspace_bt, dpred, T, V);
  rowsum_kernel<<<blocks_for(B, 128), 128, 0, st>>>(workspace_bt, loss, B, T);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_dropout_mask(float* mask, size_t n, float p, unsigned long long seed, tn_stream_t stream) {
  if (n == 0) return TN_OK;
  if (!mask || !(p >= 0.f && p < 1.f)) return set_error(TN_ERR_INVALID, "bad dropout arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dropout_mask_kernel<<<blocks_for(n, 256), 256, 0, st>>>(mask, n, p, seed);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_mul_mask(const float* x, const float* mask, const int32_t* seq_len, float* y, int B, int T, int C, tn_stream_t stream) {
  const size_t n = static_cast<size_t>(B) * T * C;
  if (n == 0) return TN_OK;
  if (!x || !y) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mul_mask_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, mask, seq_len, y, B, T, C);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_axpy(float* y, const float* x, float a, size_t n, tn_stream_t stream) {
  if (n == 0) return TN_OK;
  if (!x || !y) return set_error(TN_ERR_INVALID, "null device pointer");
  axpy_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(y, x, a, n);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

/* Whole-sequence drivers of the two calls above: the per-step loop of cell.unroll (gnmt.py:143-145) runs here instead of in
 * Python, one h2h GEMM + one cell kernel per step on `stream`.  Layouts: GI (B,T,G*H) = X W_i2h^T; Hb / Cb (B,T+1,H) with
 * slot 0 = initial state and slot t+1 = state after step t; S (B,T,4H) saved gate values; scratch (B,G*H). */
int tn_rnn_unroll_forward(int cell, int B, int T, int H, const float* GI, const float* Wh, const float* bi, const float* bh,
                          float* Hb, float* Cb, float* S, float* scratch, tn_stream_t stream) {
  if (cell != TN_CELL_GRU && cell != TN_CELL_LSTM) return set_error(TN_ERR_INVALID, "unknown cell %d", cell);
  if (B <= 0 || T <= 0 || H <= 0) return TN_OK;
  if (!GI || !Wh || !bi || !bh || !Hb || !S || !scratch || (cell == TN_CELL_LSTM && !Cb)) return set_error(TN_ERR_INVALID, "null device pointer");
  const int G = cell == TN_CELL_GRU ? 3 : 4;
  const long long GH = static_cast<long long>(G) * H, hs = static_cast<long long>(T + 1) * H;
  for (int t = 0; t < T; ++t) {
    const float* hp = Hb + static_cast<size_t>(t) * H;
    int rc = tn_sgemm(0, 1, B, static_cast<int>(GH), H, 1.f, hp, static_cast<int>(hs), Wh, H, 0.f, scratch, static_cast<int>(GH), stream);
    if (rc != TN_OK) return rc;
    rc = tn_rnn_cell_forward(cell, B, H, GI + static_cast<size_t>(t) * GH, T * GH, scratch, GH, bi, bh, hp, hs,
                             Cb ? Cb + static_cast<size_t>(t) * H : nullptr, hs, Hb + static_cast<size_t>(t + 1) * H, hs, nullptr, 0,
                             Cb ? Cb + static_cast<size_t>(t + 1) * H : nullptr, hs, S + static_cast<size_t>(t) * 4 * H,
                             static_cast<long long>(T) * 4 * H, stream);
    if (rc != TN_OK) return rc;
  }
  return TN_OK;
}

/* Reverse-time loop: dY (B,T,H) output gradients; dh / dc (B,H) zero-initialised running state gradients, on return the
 * gradients w.r.t. the initial state; DGI (and DGH for GRU; pass DGI for LSTM) (B,T,G*H) gate gradients of every step;
 * dWh (G*H,H) zero-initialised, accumulated. */
int tn_rnn_unroll_backward(int cell, int B, int T, int H, const int32_t* valid_len, const float* S, const float* Hb, const float* Cb,
                           const float* Wh, const float* dY, const float* dh_last, const float* dc_last, float* dh, float* dc,
                           float* DGI, float* DGH, float* dWh, tn_stream_t stream) {
  if (cell != TN_CELL_GRU && cell != TN_CELL_LSTM) return set_error(TN_ERR_INVALID, "unknown cell %d", cell);
  if (B <= 0 || T <= 0 || H <= 0) return TN_OK;
  if (!S || !Hb || !Wh || !dY || !dh || !DGI || !DGH || !dWh || (cell == TN_CELL_LSTM && (!Cb || !dc))) return set_error(TN_ERR_INVALID, "null device pointer");
  const int G = cell == TN_CELL_GRU ? 3 : 4;
  const long long GH = static_cast<long long>(G) * H, hs = static_cast<long long>(T + 1) * H;
  for (int t = T - 1; t >= 0; --t) {
    float* dgh = DGH + static_cast<size_t>(t) * GH;
    int rc = tn_rnn_cell_backward(cell, B, H, t, valid_len, S + static_cast<size_t>(t) * 4 * H, static_cast<long long>(T) * 4 * H,
                                  Hb + static_cast<size_t>(t) * H, hs, Cb ? Cb + static_cast<size_t>(t) * H : nullptr, hs,
                                  Cb ? Cb + static_cast<size_t>(t + 1) * H : nullptr, hs, dY + static_cast<size_t>(t) * H,
                                  static_cast<long long>(T) * H, nullptr, 0, dh_last, dc_last, dh, dc,
                                  DGI + static_cast<size_t>(t) * GH, T * GH, cell == TN_CELL_GRU ? dgh : nullptr, T * GH, stream);
    if (rc != TN_OK) return rc;
    rc = tn_sgemm(0, 0, B, H, static_cast<int>(GH), 1.f, dgh, static_cast<int>(T * GH), Wh, H, 1.f, dh, H, stream);  // dh += dgh W_h2h
    if (rc != TN_OK) return rc;
    rc = tn_sgemm(1, 0, static_cast<int>(GH), H, B, 1.f, dgh, static_cast<int>(T * GH), Hb + static_cast<size_t>(t) * H,
                  static_cast<int>(hs), 1.f, dWh, H, stream);  // dW_h2h += dgh^T h_{t-1}
    if (rc != TN_OK) return rc;
  }
  return TN_OK;
}

}  // extern "C"
