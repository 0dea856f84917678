// Training-mode building blocks of the per-frame CNN (SURVEY.md §8a V7 with a trainable backbone: `with ag.record(): out =
// net(x); ...; ag.backward(losses)` at train.py:415-421 through gluoncv's DenseNet-121 / ResNet-18 v2, Appendix A.2):
// BatchNorm with batch statistics (biased variance, running = 0.9 running + 0.1 batch) fused with ReLU, its backward, im2col /
// col2im around the shared fp32 SGEMM for the convolutions, max / average pooling with their backwards.
//
// This is the FIRST CORRECT path of the CNN backward, not the fast one: fp32 NHWC activations (row = pixel, row stride = channel
// count of the buffer, so DenseNet's concat stays a channel offset), SIMT kernels, convolutions as explicit im2col GEMMs.  The
// inference path (tn_backbone.cu: bf16, tcgen05) is untouched.  Parity bar: gradients vs torch.autograd of the fp32 oracle.
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
// One block per 32 channels, 8 row lanes; fp32 two-pass statistics (mean, then centred second moment) for accuracy.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long ldx, long long M, int C,
                                                       float* __restrict__ mean, float* __restrict__ var,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                                       float momentum) {
  __shared__ float red[8][33];
  __shared__ float smean[32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < C)
    for (long long m = ty; m < M; m += 8) acc += x[m * ldx + c];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
    float s = 0.f;
    for (int r = 0; r < 8; ++r) s += red[r][tx];
    smean[tx] = s / static_cast<float>(M);
  }
  __syncthreads();
  const float mu = smean[tx];
  acc = 0.f;
  if (c < C)
    for (long long m = ty; m < M; m += 8) {
      const float d = x[m * ldx + c] - mu;
      acc = fmaf(d, d, acc);
    }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
    for (int r = 0; r < 8; ++r) s += red[r][tx];
    const float v = s / static_cast<float>(M);  // biased, also for the running estimate (MXNet)
    mean[c] = mu;
    var[c] = v;
    if (running_mean) running_mean[c] = momentum * running_mean[c] + (1.f - momentum) * mu;
    if (running_var) running_var[c] = momentum * running_var[c] + (1.f - momentum) * v;
  }
}

// y = relu?( (x - mean) * invstd * gamma + beta )
__global__ void bn_apply_kernel(const float* __restrict__ x, long long ldx, long long M, int C, const float* __restrict__ mean,
                                const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int relu, float* __restrict__ y, long long ldy) {
  const size_t total = static_cast<size_t>(M) * C;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t m = i / C;
    const int c = static_cast<int>(i - m * C);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float v = (x[m * ldx + c] - mean[c]) * rsqrtf(var[c] + eps) * g + b;
    if (relu) v = fmaxf(v, 0.f);
    y[m * ldy + c] = v;
  }
}

// dgamma[c] = sum dy' xhat, dbeta[c] = sum dy'  with dy' = dy * (y > 0) when relu
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y,
                                                            long long ldy, const float* __restrict__ dy, long long lddy, long long M,
                                                            int C, const float* __restrict__ mean, const float* __restrict__ var,
                                                            float eps, int relu, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float rg[8][33], rb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float ag = 0.f, ab = 0.f;
  if (c < C) {
    const float mu = mean[c], is = rsqrtf(var[c] + eps);
    for (long long m = ty; m < M; m += 8) {
      float d = dy[m * lddy + c];
      if (relu && !(y[m * ldy + c] > 0.f)) d = 0.f;
      ag = fmaf(d, (x[m * ldx + c] - mu) * is, ag);
      ab += d;
    }
  }
  rg[ty][tx] = ag;
  rb[ty][tx] = ab;
  __syncthreads();
  if (ty == 0 && c < C) {
    float sg = 0.f, sb = 0.f;
    for (int r = 0; r < 8; ++r) {
      sg += rg[r][tx];
      sb += rb[r][tx];
    }
    dgamma[c] = sg;
    dbeta[c] = sb;
  }
}

// dx (+)= gamma * invstd / M * (M dy' - dbeta - xhat dgamma)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y, long long ldy,
                                    const float* __restrict__ dy, long long lddy, long long M, int C, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ gamma, float eps, int relu,
                                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, float* __restrict__ dx,
                                    long long lddx, int accumulate) {
  const size_t total = static_cast<size_t>(M) * C;
  const float invM = 1.f / static_cast<float>(M);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t m = i / C;
    const int c = static_cast<int>(i - m * C);
    const float is = rsqrtf(var[c] + eps);
    float d = dy[m * lddy + c];
    if (relu && !(y[m * ldy + c] > 0.f)) d = 0.f;
    const float xh = (x[m * ldx + c] - mean[c]) * is;
    const float g = gamma ? gamma[c] : 1.f;
    const float v = g * is * (d - invM * (dbeta[c] + xh * dgamma[c]));
    float* o = dx + m * lddx + c;
    *o = accumulate ? *o + v : v;
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
  bn_stats_kernel<<<(C + 31) / 32, 256, 0, st>>>(x, ldx, M, C, mean, var, running_mean, running_var, momentum);
  bn_apply_kernel<<<min(nblk(static_cast<size_t>(M) * C, 256), 148u * 32u), 256, 0, st>>>(x, ldx, M, C, mean, var, gamma, beta, eps, relu,
                                                                                       y, ldy);
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
  bn_bwd_reduce_kernel<<<(C + 31) / 32, 256, 0, st>>>(x, ldx, y, ldy, dy, lddy, M, C, mean, var, eps, relu, dgamma, dbeta);
  bn_bwd_apply_kernel<<<min(nblk(static_cast<size_t>(M) * C, 256), 148u * 32u), 256, 0, st>>>(x, ldx, y, ldy, dy, lddy, M, C, mean, var, gamma,
                                                                                           eps, relu, dgamma, dbeta, dx, lddx, accumulate);
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
