// Training-mode building blocks of the per-frame CNN (SURVEY.md §8a V7 with a trainable backbone: `with ag.record(): out =
// net(x); ...; ag.backward(losses)` at train.py:415-421 through gluoncv's DenseNet-121 / ResNet-18 v2, Appendix A.2):
// BatchNorm with batch statistics (biased variance, running = 0.9 running + 0.1 batch) fused with ReLU, its backward, im2col /
// col2im around the shared fp32 SGEMM for the convolutions, max / average pooling with their backwards.
//
// fp32 NHWC activations (row = pixel, row stride = channel count of the buffer, so DenseNet's concat stays a channel offset).  The
// contractions run on the tensor cores (tn_gemm_tc.cu); this file holds the memory-bound rest: BatchNorm statistics / apply / backward
// (row-split reductions with a deterministic second stage), im2col / col2im for the strided and 7x7 convolutions, pooling.  The
// inference path (tn_backbone.cu: bf16, tcgen05) is untouched.  Parity bar: gradients vs torch.autograd of the fp64 oracle.
#include <math.h>

#include "tn_common.h"

namespace {

using namespace tn;

inline unsigned nblk(size_t n, int bs) { return static_cast<unsigned>((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------------- im2col / col2im (NHWC)
// col[(n,oy,ox)][(r*S+s)*C + c] = x[(n, oy*stride-pad+r, ox*stride-pad+s)][c]  (0 outside the image)
__global__ void im2col_kernel(const float* __restrict__ x, long long ldx, int N, int H, int W, int C, int R, int S, int stride,
                              int pad, int Ho, int Wo, float* __restrict__ col) {
  const size_t K = static_cast<size_t>(R) * S * C;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * K;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / K;
    const int k = static_cast<int>(i - row * K);
    const int c = k % C, rs = k / C, s = rs % S, r = rs / S;
    const int ox = static_cast<int>(row % Wo), oy = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const int iy = oy * stride - pad + r, ix = ox * stride - pad + s;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(static_cast<size_t>(n) * H * W + static_cast<size_t>(iy) * W + ix) * ldx + c];
    col[i] = v;
  }
}
// dx[(n,iy,ix)][c] += dcol[...] (scatter with atomics: windows overlap)
__global__ void col2im_kernel(const float* __restrict__ dcol, int N, int H, int W, int C, int R, int S, int stride, int pad, int Ho,
                              int Wo, float* __restrict__ dx, long long lddx) {
  const size_t K = static_cast<size_t>(R) * S * C;
  const size_t total = static_cast<size_t>(N) * Ho * Wo * K;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / K;
    const int k = static_cast<int>(i - row * K);
    const int c = k % C, rs = k / C, s = rs % S, r = rs / S;
    const int ox = static_cast<int>(row % Wo), oy = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    const int iy = oy * stride - pad + r, ix = ox * stride - pad + s;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      atomicAdd(dx + (static_cast<size_t>(n) * H * W + static_cast<size_t>(iy) * W + ix) * lddx + c, dcol[i]);
  }
}

// ---------------------------------------------------------------------------------------------------- BatchNorm (training)
// Per-channel reductions over M = frames x pixels rows: the rows are split over gridDim.y blocks (one block = 32 channels x 8 row
// lanes over a contiguous row range) so that the whole machine streams the activation once; the per-block partial sums go to a
// stream-ordered scratch buffer and a second, tiny kernel adds them in a fixed order (deterministic) in double precision.
// Variance by the shifted-data formula with the channel's first row as the shift: one pass over x, no cancellation between two
// large sums (var = (S2 - S1^2/M)/M with S1 = sum(x - k), S2 = sum((x - k)^2)).
constexpr int kBnMaxSplits = 512;

struct BnSplit {
  int splits;
  long long rows_per_block;
};
inline BnSplit bn_split(long long M, int C) {
  const int cgroups = (C + 31) / 32;
  long long s = (148LL * 8 + cgroups - 1) / cgroups;
  const long long by_rows = (M + 63) / 64;
  if (s > by_rows) s = by_rows;
  if (s > kBnMaxSplits) s = kBnMaxSplits;
  if (s < 1) s = 1;
  long long rpb = (M + s - 1) / s;
  rpb = (rpb + 7) / 8 * 8;
  BnSplit r;
  r.rows_per_block = rpb;
  r.splits = static_cast<int>((M + rpb - 1) / rpb);
  return r;
}

// part[(split * 2 + {0,1}) * C + c] = sum (x - x[0][c]), sum (x - x[0][c])^2 over the block's rows
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(const float* __restrict__ x, long long ldx, long long M, int C,
                                                               long long rows_per_block, float* __restrict__ part) {
  __shared__ float r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long m0 = blockIdx.y * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
  if (c < C) {
    const float k = x[c];
    const float* xc = x + c;
    long long m = m0 + ty;
    for (; m + 24 < m1; m += 32) {  // four independent loads in flight per thread
      const float d0 = xc[m * ldx] - k, d1 = xc[(m + 8) * ldx] - k, d2 = xc[(m + 16) * ldx] - k, d3 = xc[(m + 24) * ldx] - k;
      a1 += d0; a2 = fmaf(d0, d0, a2);
      b1 += d1; b2 = fmaf(d1, d1, b2);
      a1 += d2; a2 = fmaf(d2, d2, a2);
      b1 += d3; b2 = fmaf(d3, d3, b2);
    }
    for (; m < m1; m += 8) {
      const float d = xc[m * ldx] - k;
      a1 += d; a2 = fmaf(d, d, a2);
    }
  }
  r1[ty][tx] = a1 + b1;
  r2[ty][tx] = a2 + b2;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < 8; ++r) {
      s1 += r1[r][tx];
      s2 += r2[r][tx];
    }
    part[(static_cast<size_t>(blockIdx.y) * 2 + 0) * C + c] = s1;
    part[(static_cast<size_t>(blockIdx.y) * 2 + 1) * C + c] = s2;
  }
}
// one block = 32 channels x 8 split lanes: lane r adds splits r, r+8, ... in double, the eight lane sums are added in order
__global__ void __launch_bounds__(256) bn_stats_finalize_kernel(const float* __restrict__ x, const float* __restrict__ part, int splits,
                                                                long long M, int C, float* __restrict__ mean, float* __restrict__ var,
                                                                float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                float momentum) {
  __shared__ double r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double s1 = 0.0, s2 = 0.0;
  if (c < C)
    for (int s = ty; s < splits; s += 8) {
      s1 += static_cast<double>(part[(static_cast<size_t>(s) * 2 + 0) * C + c]);
      s2 += static_cast<double>(part[(static_cast<size_t>(s) * 2 + 1) * C + c]);
    }
  r1[ty][tx] = s1;
  r2[ty][tx] = s2;
  __syncthreads();
  if (ty != 0 || c >= C) return;
  s1 = s2 = 0.0;
  for (int r = 0; r < 8; ++r) {
    s1 += r1[r][tx];
    s2 += r2[r][tx];
  }
  const double inv = 1.0 / static_cast<double>(M);
  const float mu = static_cast<float>(static_cast<double>(x[c]) + s1 * inv);
  double vv = (s2 - s1 * s1 * inv) * inv;  // biased, also for the running estimate (MXNet)
  if (vv < 0.0) vv = 0.0;
  const float v = static_cast<float>(vv);
  mean[c] = mu;
  var[c] = v;
  if (running_mean) running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mu;
  if (running_var) running_var[c] = momentum * running_var[c] + (1.f - momentum) * v;
}

// y = relu?( (x - mean) * invstd * gamma + beta ); VEC = 4: rows and channel counts allow 16-byte accesses
template <int VEC>
__global__ void bn_apply_kernel(const float* __restrict__ x, long long ldx, long long M, int C, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int relu, float* __restrict__ y, long long ldy) {
  const unsigned Cv = static_cast<unsigned>(C / VEC);
  const size_t total = static_cast<size_t>(M) * Cv;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t m;
    int c;
    if (total < 0xffffffffull) {
      const unsigned iu = static_cast<unsigned>(i), mu = iu / Cv;
      m = mu;
      c = static_cast<int>(iu - mu * Cv) * VEC;
    } else {
      m = i / Cv;
      c = static_cast<int>(i - m * Cv) * VEC;
    }
    float xv[VEC], o[VEC];
    if (VEC == 4) *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(x + m * ldx + c);
    else xv[0] = x[m * ldx + c];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float g = gamma ? gamma[c + q] : 1.f, b = beta ? beta[c + q] : 0.f;
      float v = (xv[q] - mean[c + q]) * rsqrtf(var[c + q] + eps) * g + b;
      if (relu) v = fmaxf(v, 0.f);
      o[q] = v;
    }
    if (VEC == 4) *reinterpret_cast<float4*>(y + m * ldy + c) = *reinterpret_cast<const float4*>(o);
    else y[m * ldy + c] = o[0];
  }
}

// dgamma[c] = sum dy' xhat, dbeta[c] = sum dy'  with dy' = dy * (y > 0) when relu; same row split as the statistics
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y,
                                                             long long ldy, const float* __restrict__ dy, long long lddy, long long M,
                                                             int C, const float* __restrict__ mean, const float* __restrict__ var,
                                                             float eps, int relu, long long rows_per_block, float* __restrict__ part) {
  __shared__ float rg[8][33], rb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const long long m0 = blockIdx.y * rows_per_block;
  long long m1 = m0 + rows_per_block;
  if (m1 > M) m1 = M;
  float ag = 0.f, ab = 0.f, bg = 0.f, bb = 0.f;
  if (c < C) {
    const float mu = mean[c], is = rsqrtf(var[c] + eps);
    long long m = m0 + ty;
    for (; m + 8 < m1; m += 16) {
      float d0 = dy[m * lddy + c], d1 = dy[(m + 8) * lddy + c];
      const float x0 = x[m * ldx + c], x1 = x[(m + 8) * ldx + c];
      if (relu) {
        if (!(y[m * ldy + c] > 0.f)) d0 = 0.f;
        if (!(y[(m + 8) * ldy + c] > 0.f)) d1 = 0.f;
      }
      ag = fmaf(d0, (x0 - mu) * is, ag);
      ab += d0;
      bg = fmaf(d1, (x1 - mu) * is, bg);
      bb += d1;
    }
    for (; m < m1; m += 8) {
      float d = dy[m * lddy + c];
      if (relu && !(y[m * ldy + c] > 0.f)) d = 0.f;
      ag = fmaf(d, (x[m * ldx + c] - mu) * is, ag);
      ab += d;
    }
  }
  rg[ty][tx] = ag + bg;
  rb[ty][tx] = ab + bb;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sg = 0.f, sb = 0.f;
    for (int r = 0; r < 8; ++r) {
      sg += rg[r][tx];
      sb += rb[r][tx];
    }
    part[(static_cast<size_t>(blockIdx.y) * 2 + 0) * C + c] = sg;
    part[(static_cast<size_t>(blockIdx.y) * 2 + 1) * C + c] = sb;
  }
}
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ part, int splits, int C,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double r1[8][33], r2[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double sg = 0.0, sb = 0.0;
  if (c < C)
    for (int s = ty; s < splits; s += 8) {
      sg += static_cast<double>(part[(static_cast<size_t>(s) * 2 + 0) * C + c]);
      sb += static_cast<double>(part[(static_cast<size_t>(s) * 2 + 1) * C + c]);
    }
  r1[ty][tx] = sg;
  r2[ty][tx] = sb;
  __syncthreads();
  if (ty != 0 || c >= C) return;
  sg = sb = 0.0;
  for (int r = 0; r < 8; ++r) {
    sg += r1[r][tx];
    sb += r2[r][tx];
  }
  dgamma[c] = static_cast<float>(sg);
  dbeta[c] = static_cast<float>(sb);
}

// dx (+)= gamma * invstd / M * (M dy' - dbeta - xhat dgamma)
template <int VEC>
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y, long long ldy,
                                    const float* __restrict__ dy, long long lddy, long long M, int C, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ gamma, float eps, int relu,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx,
                                    long long lddx, int accumulate) {
  const unsigned Cv = static_cast<unsigned>(C / VEC);
  const size_t total = static_cast<size_t>(M) * Cv;
  const float invM = 1.f / static_cast<float>(M);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t m;
    int c;
    if (total < 0xffffffffull) {
      const unsigned iu = static_cast<unsigned>(i), mu = iu / Cv;
      m = mu;
      c = static_cast<int>(iu - mu * Cv) * VEC;
    } else {
      m = i / Cv;
      c = static_cast<int>(i - m * Cv) * VEC;
    }
    float xv[VEC], dv[VEC], yv[VEC], o[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(x + m * ldx + c);
      *reinterpret_cast<float4*>(dv) = *reinterpret_cast<const float4*>(dy + m * lddy + c);
      if (relu) *reinterpret_cast<float4*>(yv) = *reinterpret_cast<const float4*>(y + m * ldy + c);
      if (accumulate) *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(dx + m * lddx + c);
    } else {
      xv[0] = x[m * ldx + c];
      dv[0] = dy[m * lddy + c];
      if (relu) yv[0] = y[m * ldy + c];
      if (accumulate) o[0] = dx[m * lddx + c];
    }
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const float is = rsqrtf(var[c + q] + eps);
      float d = dv[q];
      if (relu && !(yv[q] > 0.f)) d = 0.f;
      const float xh = (xv[q] - mean[c + q]) * is;
      const float g = gamma ? gamma[c + q] : 1.f;
      const float v = g * is * (d - invM * (dbeta[c + q] + xh * dgamma[c + q]));
      o[q] = accumulate ? o[q] + v : v;
    }
    if (VEC == 4) *reinterpret_cast<float4*>(dx + m * lddx + c) = *reinterpret_cast<const float4*>(o);
    else dx[m * lddx + c] = o[0];
  }
}

// ---------------------------------------------------------------------------------------------------- pooling (NHWC)
// max-pool k/stride/pad with -inf padding; idx = flat input pixel of the winner (first maximum in window order, like torch)
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, long long ldx, int N, int H, int W, int C, int k, int stride, int pad,
                                   int Ho, int Wo, float* __restrict__ y, long long ldy, int* __restrict__ idx) {
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t row = i / C;
    const int ox = static_cast<int>(row % Wo), oy = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    float best = -INFINITY;
    int bi = -1;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        const int iy = oy * stride - pad + r, ix = ox * stride - pad + s;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const int pix = (n * H + iy) * W + ix;
        const float v = x[static_cast<size_t>(pix) * ldx + c];
        if (v > best) {
          best = v;
          bi = pix;
        }
      }
    y[row * ldy + c] = best;
    idx[i] = bi;
  }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, long long lddy, const int* __restrict__ idx, size_t rows, int C,
                                   float* __restrict__ dx, long long lddx) {
  const size_t total = rows * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t row = i / C;
    const int pix = idx[i];
    if (pix >= 0) atomicAdd(dx + static_cast<size_t>(pix) * lddx + c, dy[row * lddy + c]);
  }
}
// average pool with window k and stride k (floor, no padding): transitions (2), DenseNet tail (7), global (k = H)
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, long long ldx, int N, int H, int W, int C, int kh, int kw, int Ho,
                                   int Wo, float* __restrict__ y, long long ldy) {
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  const float inv = 1.f / static_cast<float>(kh * kw);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t row = i / C;
    const int ox = static_cast<int>(row % Wo), oy = static_cast<int>((row / Wo) % Ho), n = static_cast<int>(row / (static_cast<size_t>(Wo) * Ho));
    float s = 0.f;
    for (int r = 0; r < kh; ++r)
      for (int q = 0; q < kw; ++q) s += x[(static_cast<size_t>(n) * H * W + static_cast<size_t>(oy * kh + r) * W + ox * kw + q) * ldx + c];
    y[row * ldy + c] = s * inv;
  }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, long long lddy, int N, int H, int W, int C, int kh, int kw, int Ho,
                                   int Wo, float* __restrict__ dx, long long lddx, int accumulate) {
  const size_t total = static_cast<size_t>(N) * H * W * C;
  const float inv = 1.f / static_cast<float>(kh * kw);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const int ix = static_cast<int>(pix % W), iy = static_cast<int>((pix / W) % H), n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    const int oy = iy / kh, ox = ix / kw;
    float v = 0.f;
    if (oy < Ho && ox < Wo) v = dy[(static_cast<size_t>(n) * Ho * Wo + static_cast<size_t>(oy) * Wo + ox) * lddy + c] * inv;
    float* o = dx + pix * lddx + c;
    *o = accumulate ? *o + v : v;
  }
}

}  // namespace

extern "C" {

int tn_im2col_nhwc(const float* x, long long ldx, int N, int H, int W, int C, int R, int S, int stride, int pad, float* col,
                   tn_stream_t stream) {
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return set_error(TN_ERR_INVALID, "empty im2col");
  if (!x || !col) return set_error(TN_ERR_INVALID, "null device pointer");
  const size_t total = static_cast<size_t>(N) * Ho * Wo * R * S * C;
  ProfScope ps(kProfOther, static_cast<cudaStream_t>(stream));
  im2col_kernel<<<min(nblk(total, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, N, H, W, C, R, S, stride, pad,
                                                                                               Ho, Wo, col);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_col2im_nhwc(const float* dcol, int N, int H, int W, int C, int R, int S, int stride, int pad, float* dx, long long lddx,
                   tn_stream_t stream) {
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return set_error(TN_ERR_INVALID, "empty col2im");
  if (!dcol || !dx) return set_error(TN_ERR_INVALID, "null device pointer");
  const size_t total = static_cast<size_t>(N) * Ho * Wo * R * S * C;
  ProfScope ps(kProfOther, static_cast<cudaStream_t>(stream));
  col2im_kernel<<<min(nblk(total, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(dcol, N, H, W, C, R, S, stride, pad, Ho,
                                                                                               Wo, dx, lddx);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_bn_train_forward(const float* x, long long ldx, long long M, int C, const float* gamma, const float* beta, float eps,
                        float momentum, float* running_mean, float* running_var, int relu, float* mean, float* var, float* y,
                        long long ldy, tn_stream_t stream) {
  if (M <= 0 || C <= 0) return TN_OK;
  if (!x || !mean || !var || !y) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(kProfOther, st);
  const BnSplit sp = bn_split(M, C);
  float* part = nullptr;
  TN_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), static_cast<size_t>(sp.splits) * 2 * C * sizeof(float), st));
  bn_stats_partial_kernel<<<dim3((C + 31) / 32, sp.splits), 256, 0, st>>>(x, ldx, M, C, sp.rows_per_block, part);
  bn_stats_finalize_kernel<<<(C + 31) / 32, 256, 0, st>>>(x, part, sp.splits, M, C, mean, var, running_mean, running_var, momentum);
  TN_CUDA(cudaFreeAsync(part, st));
  const bool v4 = C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 == 0;
  if (v4)
    bn_apply_kernel<4><<<min(nblk(static_cast<size_t>(M) * (C / 4), 256), 148u * 32u), 256, 0, st>>>(x, ldx, M, C, mean, var, gamma, beta,
                                                                                                 eps, relu, y, ldy);
  else
    bn_apply_kernel<1><<<min(nblk(static_cast<size_t>(M) * C, 256), 148u * 32u), 256, 0, st>>>(x, ldx, M, C, mean, var, gamma, beta, eps,
                                                                                            relu, y, ldy);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_bn_train_backward(const float* x, long long ldx, const float* y, long long ldy, const float* dy, long long lddy, long long M,
                         int C, const float* mean, const float* var, const float* gamma, float eps, int relu, float* dgamma,
                         float* dbeta, float* dx, long long lddx, int accumulate, tn_stream_t stream) {
  if (M <= 0 || C <= 0) return TN_OK;
  if (!x || !dy || !mean || !var || !dgamma || !dbeta || !dx || (relu && !y)) return set_error(TN_ERR_INVALID, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(kProfOther, st);
  const BnSplit sp = bn_split(M, C);
  float* part = nullptr;
  TN_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&part), static_cast<size_t>(sp.splits) * 2 * C * sizeof(float), st));
  bn_bwd_partial_kernel<<<dim3((C + 31) / 32, sp.splits), 256, 0, st>>>(x, ldx, y, ldy, dy, lddy, M, C, mean, var, eps, relu,
                                                                         sp.rows_per_block, part);
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 256, 0, st>>>(part, sp.splits, C, dgamma, dbeta);
  TN_CUDA(cudaFreeAsync(part, st));
  const bool v4 = C % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && (!relu || ldy % 4 == 0) &&
                  (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
                   (relu ? reinterpret_cast<uintptr_t>(y) : 0)) % 16 == 0;
  if (v4)
    bn_bwd_apply_kernel<4><<<min(nblk(static_cast<size_t>(M) * (C / 4), 256), 148u * 32u), 256, 0, st>>>(
        x, ldx, y, ldy, dy, lddy, M, C, mean, var, gamma, eps, relu, dgamma, dbeta, dx, lddx, accumulate);
  else
    bn_bwd_apply_kernel<1><<<min(nblk(static_cast<size_t>(M) * C, 256), 148u * 32u), 256, 0, st>>>(
        x, ldx, y, ldy, dy, lddy, M, C, mean, var, gamma, eps, relu, dgamma, dbeta, dx, lddx, accumulate);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_maxpool_nhwc_forward(const float* x, long long ldx, int N, int H, int W, int C, int k, int stride, int pad, float* y,
                            long long ldy, int32_t* idx, tn_stream_t stream) {
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return set_error(TN_ERR_INVALID, "empty pool");
  if (!x || !y || !idx) return set_error(TN_ERR_INVALID, "null device pointer");
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  maxpool_fwd_kernel<<<min(nblk(total, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, N, H, W, C, k, stride, pad,
                                                                                                    Ho, Wo, y, ldy, idx);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_maxpool_nhwc_backward(const float* dy, long long lddy, const int32_t* idx, long long rows, int C, float* dx, long long lddx,
                             tn_stream_t stream) {
  if (rows <= 0 || C <= 0) return TN_OK;
  if (!dy || !idx || !dx) return set_error(TN_ERR_INVALID, "null device pointer");
  maxpool_bwd_kernel<<<min(nblk(static_cast<size_t>(rows) * C, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, lddy, idx, static_cast<size_t>(rows), C, dx, lddx);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_avgpool_nhwc_forward(const float* x, long long ldx, int N, int H, int W, int C, int kh, int kw, float* y, long long ldy,
                            tn_stream_t stream) {
  const int Ho = H / kh, Wo = W / kw;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return set_error(TN_ERR_INVALID, "empty pool");
  if (!x || !y) return set_error(TN_ERR_INVALID, "null device pointer");
  const size_t total = static_cast<size_t>(N) * Ho * Wo * C;
  avgpool_fwd_kernel<<<min(nblk(total, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, N, H, W, C, kh, kw, Ho, Wo, y,
                                                                                                    ldy);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_avgpool_nhwc_backward(const float* dy, long long lddy, int N, int H, int W, int C, int kh, int kw, float* dx, long long lddx,
                             int accumulate, tn_stream_t stream) {
  const int Ho = H / kh, Wo = W / kw;
  if (N <= 0 || Ho <= 0 || Wo <= 0) return set_error(TN_ERR_INVALID, "empty pool");
  if (!dy || !dx) return set_error(TN_ERR_INVALID, "null device pointer");
  const size_t total = static_cast<size_t>(N) * H * W * C;
  avgpool_bwd_kernel<<<min(nblk(total, 256), 148u * 32u), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, lddy, N, H, W, C, kh, kw, Ho, Wo,
                                                                                                    dx, lddx, accumulate);
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

}  // extern "C"
