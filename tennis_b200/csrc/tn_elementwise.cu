// HBM-bound helpers around the conv GEMMs: layout conversion, max-pool, the BN+ReLU+avg-pool tail,
// small dense layers and temporal pooling.  All are 16-byte-vectorised, coalesced along channels.
#include "tn_elementwise.h"
#include "tn_common.h"
#include "tn_ptx.cuh"

namespace tn {

// ---------------------------------------------------------------- input conversion
// (n,3,H,W) fp32 NCHW  ->  (n,H,W,4) bf16 NHWC with a zero 4th channel; optional per-channel affine
// (ResNet-v2's leading BatchNorm(scale=False, center=False) is folded here).
__global__ void convert_nchw_f32_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t npix_total,
                                        int hw, float s0, float s1, float s2, float b0, float b1, float b2) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix_total) return;
  size_t f = i / hw;
  size_t r = i - f * hw;
  const float* base = in + f * 3 * hw + r;
  float x0 = __ldg(base) * s0 + b0;
  float x1 = __ldg(base + hw) * s1 + b1;
  float x2 = __ldg(base + 2 * static_cast<size_t>(hw)) * s2 + b2;
  uint2 v = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, 0.f));
  reinterpret_cast<uint2*>(out)[i] = v;
}

// (n,H,W,3) uint8 NHWC (what a decoder produces) -> ToTensor + Normalize(mean,std) (train.py:142-147)
// + optional affine -> (n,H,W,4) bf16.  Moves 4x fewer bytes over PCIe than the fp32 path (K13).
__global__ void convert_nhwc_u8_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                       size_t npix_total, float s0, float s1, float s2, float b0, float b1, float b2) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= npix_total) return;
  const uint8_t* p = in + i * 3;
  float x0 = static_cast<float>(p[0]) * s0 + b0;
  float x1 = static_cast<float>(p[1]) * s1 + b1;
  float x2 = static_cast<float>(p[2]) * s2 + b2;
  reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, 0.f));
}

cudaError_t launch_convert_nchw_f32(const float* in, __nv_bfloat16* out, int n, int h, int w, const float* scale3,
                                    const float* shift3, cudaStream_t st) {
  size_t total = static_cast<size_t>(n) * h * w;
  if (total == 0) return cudaSuccess;
  float s[3] = {1, 1, 1}, b[3] = {0, 0, 0};
  if (scale3) { s[0] = scale3[0]; s[1] = scale3[1]; s[2] = scale3[2]; }
  if (shift3) { b[0] = shift3[0]; b[1] = shift3[1]; b[2] = shift3[2]; }
  ProfScope prof_scope(kProfOther, st);
  convert_nchw_f32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, total, h * w, s[0], s[1],
                                                                                      s[2], b[0], b[1], b[2]);
  return cudaGetLastError();
}

cudaError_t launch_convert_nhwc_u8(const uint8_t* in, __nv_bfloat16* out, int n, int h, int w, const float* scale3,
                                   const float* shift3, cudaStream_t st) {
  size_t total = static_cast<size_t>(n) * h * w;
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  convert_nhwc_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      in, out, total, scale3[0], scale3[1], scale3[2], shift3[0], shift3[1], shift3[2]);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- max-pool 3x3 / stride 2 / pad 1
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__global__ void maxpool3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int H,
                                  int W, int C, int Ho, int Wo, int out_cstride, int out_coff) {
  const int cg = C / 8;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t total = static_cast<size_t>(n) * Ho * Wo * cg;
  if (i >= total) return;
  int c8 = static_cast<int>(i % cg);
  size_t pix = i / cg;
  int ox = static_cast<int>(pix % Wo);
  size_t t = pix / Wo;
  int oy = static_cast<int>(t % Ho);
  size_t f = t / Ho;
  const uint32_t NEG = 0xFF80FF80u;  // (-inf, -inf) bf16
  uint4 m = make_uint4(NEG, NEG, NEG, NEG);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= W) continue;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((f * H + iy) * W + ix) * C + c8 * 8));
      m.x = bf16x2_max(m.x, v.x);
      m.y = bf16x2_max(m.y, v.y);
      m.z = bf16x2_max(m.z, v.z);
      m.w = bf16x2_max(m.w, v.w);
    }
  }
  *reinterpret_cast<uint4*>(out + pix * out_cstride + out_coff + c8 * 8) = m;
}

cudaError_t launch_maxpool3s2(const __nv_bfloat16* in, __nv_bfloat16* out, int n, int H, int W, int C, int Ho, int Wo,
                              int out_cstride, int out_coff, cudaStream_t st) {
  size_t total = static_cast<size_t>(n) * Ho * Wo * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  maxpool3s2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, out, n, H, W, C, Ho, Wo, out_cstride,
                                                                               out_coff);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- tail: BN -> ReLU -> AvgPool(k, stride k, valid) -> Flatten
// feats[f, c*(ph*pw) + py*pw + px]  (Gluon Flatten of (N,C,ph,pw) is channel-major)
__global__ void tail_pool_kernel(const __nv_bfloat16* __restrict__ in, int n, int H, int W, int C, int cstride, int kh,
                                 int kw, int ph, int pw, const float* __restrict__ scale, const float* __restrict__ shift,
                                 float* __restrict__ feats, __nv_bfloat16* __restrict__ feats_bf16) {
  const int cg = C / 8;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t total = static_cast<size_t>(n) * ph * pw * cg;
  if (i >= total) return;
  int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  int px = static_cast<int>(t % pw);
  t /= pw;
  int py = static_cast<int>(t % ph);
  size_t f = t / ph;
  float sc[8], sh[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + c8 * 8 + j);
    sh[j] = __ldg(shift + c8 * 8 + j);
    acc[j] = 0.f;
  }
  for (int dy = 0; dy < kh; ++dy) {
    for (int dx = 0; dx < kw; ++dx) {
      int iy = py * kh + dy, ix = px * kw + dx;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((f * H + iy) * W + ix) * cstride + c8 * 8));
      uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 x = unpack_bf16x2(w4[j]);
        acc[2 * j] += fmaxf(fmaf(x.x, sc[2 * j], sh[2 * j]), 0.f);
        acc[2 * j + 1] += fmaxf(fmaf(x.y, sc[2 * j + 1], sh[2 * j + 1]), 0.f);
      }
    }
  }
  const float inv = 1.f / static_cast<float>(kh * kw);
  const int pp = ph * pw;
  const size_t D = static_cast<size_t>(C) * pp;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = acc[j] * inv;
    size_t o = f * D + static_cast<size_t>(c8 * 8 + j) * pp + py * pw + px;
    feats[o] = v;
    if (feats_bf16) feats_bf16[o] = __float2bfloat16(v);
  }
}

cudaError_t launch_tail_pool(const __nv_bfloat16* in, int n, int H, int W, int C, int cstride, int kh, int kw, int ph,
                             int pw, const float* scale, const float* shift, float* feats, __nv_bfloat16* feats_bf16,
                             cudaStream_t st) {
  size_t total = static_cast<size_t>(n) * ph * pw * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  tail_pool_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, st>>>(in, n, H, W, C, cstride, kh, kw, ph, pw,
                                                                              scale, shift, feats, feats_bf16);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- transition pre-pass: BN -> ReLU -> AvgPool 2x2 (bf16 out)
// DenseNet transition = BN -> ReLU -> conv1x1 -> avgpool2; the pool commutes with the 1x1 conv, so the activated input is
// pooled first (this kernel, a pure streaming pass) and the GEMM then runs on 4x fewer rows through the TMA-fed path.
__global__ void bn_relu_pool2_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int C, int cstride, int Ho, int Wo,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     __nv_bfloat16* __restrict__ out, size_t total) {
  const int cg = C / 8;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  const int ox = static_cast<int>(t % Wo);
  t /= Wo;
  const int oy = static_cast<int>(t % Ho);
  const size_t f = t / Ho;
  const __nv_bfloat16* base = in + ((f * H + 2 * oy) * W + 2 * ox) * cstride + c8 * 8;
  uint4 v[4];
  v[0] = __ldg(reinterpret_cast<const uint4*>(base));
  v[1] = __ldg(reinterpret_cast<const uint4*>(base + cstride));
  v[2] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(W) * cstride));
  v[3] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(W + 1) * cstride));
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c8 * 8)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c8 * 8) + 1);
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c8 * 8)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c8 * 8) + 1);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t w4[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = unpack_bf16x2(w4[j]);
      acc[2 * j] += fmaxf(fmaf(x.x, sc[2 * j], sh[2 * j]), 0.f);
      acc[2 * j + 1] += fmaxf(fmaf(x.y, sc[2 * j + 1], sh[2 * j + 1]), 0.f);
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0] * 0.25f, acc[1] * 0.25f);
  o.y = pack_bf16x2(acc[2] * 0.25f, acc[3] * 0.25f);
  o.z = pack_bf16x2(acc[4] * 0.25f, acc[5] * 0.25f);
  o.w = pack_bf16x2(acc[6] * 0.25f, acc[7] * 0.25f);
  *reinterpret_cast<uint4*>(out + i * 8) = o;
}

cudaError_t launch_bn_relu_pool2(const __nv_bfloat16* in, int n, int H, int W, int C, int cstride, const float* scale,
                                 const float* shift, __nv_bfloat16* out, cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(n) * Ho * Wo * (C / 8);
  if (total == 0) return cudaSuccess;
  if ((C % 8) != 0 || (cstride % 8) != 0) return cudaErrorInvalidValue;
  ProfScope prof_scope(kProfOther, st);
  bn_relu_pool2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, H, W, C, cstride, Ho, Wo, scale, shift,
                                                                                  out, total);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- small dense layer (fp32), one warp per output
// y[r, j] = b[j] + sum_k x[r, k] * W[j, k]     (Gluon Dense: weight (out, in))
__global__ void dense_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                             float* __restrict__ y, int rows, int in_dim, int out_dim) {
  const int warps_per_block = blockDim.x >> 5;
  size_t wid = static_cast<size_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<size_t>(rows) * out_dim) return;
  const int j = static_cast<int>(wid % out_dim);
  const size_t r = wid / out_dim;
  const float* xr = x + r * in_dim;
  const float* wj = W + static_cast<size_t>(j) * in_dim;
  float acc = 0.f;
  if ((in_dim & 3) == 0) {
    for (int k = lane * 4; k < in_dim; k += 128) {
      float4 a = __ldg(reinterpret_cast<const float4*>(xr + k));
      float4 w = __ldg(reinterpret_cast<const float4*>(wj + k));
      acc = fmaf(a.x, w.x, acc);
      acc = fmaf(a.y, w.y, acc);
      acc = fmaf(a.z, w.z, acc);
      acc = fmaf(a.w, w.w, acc);
    }
  } else {
    for (int k = lane; k < in_dim; k += 32) acc = fmaf(__ldg(xr + k), __ldg(wj + k), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[r * out_dim + j] = acc + (b ? __ldg(b + j) : 0.f);
}

cudaError_t launch_dense(const float* x, const float* W, const float* b, float* y, int rows, int in_dim, int out_dim,
                         cudaStream_t st) {
  size_t total = static_cast<size_t>(rows) * out_dim;
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  dense_kernel<<<static_cast<unsigned>((total + 7) / 8), 256, 0, st>>>(x, W, b, y, rows, in_dim, out_dim);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- temporal pooling over axis 1 of (B,T,D)
__global__ void temporal_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int D, int mean) {
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * D) return;
  size_t b = i / D;
  int d = static_cast<int>(i - b * D);
  const float* p = x + b * T * D + d;
  float acc = mean ? 0.f : -INFINITY;
  for (int t = 0; t < T; ++t) {
    float v = __ldg(p + static_cast<size_t>(t) * D);
    acc = mean ? acc + v : fmaxf(acc, v);
  }
  y[i] = mean ? acc / static_cast<float>(T) : acc;
}

cudaError_t launch_temporal_pool(const float* x, float* y, int B, int T, int D, int mean, cudaStream_t st) {
  size_t total = static_cast<size_t>(B) * D;
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  temporal_pool_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, y, B, T, D, mean);
  return cudaGetLastError();
}

// fp32 -> bf16 cast (feature matrices fed to the tensor-core input projection)
__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
    reinterpret_cast<uint2*>(y + i)[0] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (; i < n; ++i) y[i] = __float2bfloat16(x[i]);
  }
}
cudaError_t launch_cast_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  size_t threads = (n + 3) / 4;
  ProfScope prof_scope(kProfOther, st);
  cast_bf16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(x, y, n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- space-to-depth input for the 7x7/2 stem
// A 7x7 / stride 2 / pad 3 conv equals a 4x4 / stride 1 conv on the 2x2 space-to-depth image (filter padded to 8x8):
//   out(oy,ox) = sum_{a,b<4} sum_{py,px<2,c<3} W8[2a+py][2b+px][c] * Z[oy-2+a][ox-2+b][(py,px,c)],  W8[u][v] = W[u-1][v-1].
// Z is stored zero-padded (2 px top/left, 1 px bottom/right) with 16 channels per pixel (12 used), so that one filter row
// (4 px x 16 ch) is 128 contiguous bytes == one K-chunk of the implicit GEMM, fetched with plain 16-byte loads.
template <typename SrcT, bool NHWC>
__global__ void s2d_convert_kernel(const SrcT* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int H, int W, int Hz,
                                   int Wz, float s0, float s1, float s2, float b0, float b1, float b2) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n) * Hz * Wz;
  if (i >= total) return;
  const int Xp = static_cast<int>(i % Wz);
  const size_t t = i / Wz;
  const int Yp = static_cast<int>(t % Hz);
  const size_t f = t / Hz;
  const float sc[3] = {s0, s1, s2}, sh[3] = {b0, b1, b2};
  float v[12];
#pragma unroll
  for (int py = 0; py < 2; ++py) {
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int y = 2 * (Yp - 2) + py, x = 2 * (Xp - 2) + px;
      const bool ok = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float val = 0.f;
        if (ok) {
          const size_t src = NHWC ? ((f * H + y) * W + x) * 3 + c : ((f * 3 + c) * H + y) * W + x;
          val = static_cast<float>(in[src]) * sc[c] + sh[c];
        }
        v[(py * 2 + px) * 3 + c] = val;
      }
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
  dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), 0u, 0u);
}

cudaError_t launch_s2d_convert(const void* in, int is_u8_nhwc, __nv_bfloat16* out, int n, int h, int w, int Hz, int Wz,
                               const float* scale3, const float* shift3, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * Hz * Wz;
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (is_u8_nhwc)
    s2d_convert_kernel<uint8_t, true><<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(in), out, n, h, w, Hz, Wz, scale3[0],
                                                            scale3[1], scale3[2], shift3[0], shift3[1], shift3[2]);
  else
    s2d_convert_kernel<float, false><<<grid, 256, 0, st>>>(static_cast<const float*>(in), out, n, h, w, Hz, Wz, scale3[0],
                                                           scale3[1], scale3[2], shift3[0], shift3[1], shift3[2]);
  return cudaGetLastError();
}

// split-bf16 operand for the "precise" input projection: row m -> [hi(x) | lo(x) | hi(x)]
__global__ void split3_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t M, int D) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const size_t m = i / D;
  const int k = static_cast<int>(i - m * D);
  const float v = x[i];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  __nv_bfloat16* row = y + m * 3 * D;
  row[k] = hi;
  row[D + k] = lo;
  row[2 * D + k] = hi;
}
cudaError_t launch_split3_bf16(const float* x, __nv_bfloat16* y, size_t M, int D, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  split3_bf16_kernel<<<static_cast<unsigned>((M * D + 255) / 256), 256, 0, st>>>(x, y, M, D);
  return cudaGetLastError();
}

}  // namespace tn
