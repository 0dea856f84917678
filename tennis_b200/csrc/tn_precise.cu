// Elementwise stages of the fp32-grade ("precise") CNN forward.
//
// The reference network computes in fp32 end to end (train.py:204, evaluate.py:125).  The speed path stores activations and
// operands as single bf16 values (8 mantissa bits): logits land within ~1e-2 of the fp32 oracle.  The precise path keeps every
// tensor as a PAIR of bf16 planes, x = hi + lo (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits), and every contraction as
// three tensor-core products accumulated in fp32:  x*w ~= hi*Wh + hi*Wl + lo*Wh.  The products run through the SAME tcgen05
// implicit-GEMM kernel as the speed path (tn_conv_gemm.cu): the three terms are listed as extra "row taps" of the K loop -- the
// lo plane lies a fixed number of rows behind the hi plane, and the weight chunks of a tap are the Wh or Wl image.  Everything
// that is not a contraction lives here, in fp32 on (hi + lo):
//   s2d_convert_x2      frames -> zero-padded space-to-depth image planes
//   maxpool_f32_split   stem output (fp32, padded s2d grid) -> max-pool 3/2/1 -> planes of dense block 1
//   bn_relu_split       concat planes -> relu(bn(x)) -> activated operand planes (the 1x1 conv's A matrix)
//   bn_relu_pool2_x2    transition: relu(bn(x)) averaged over 2x2 -> operand planes
//   split_store         fp32 GEMM output -> planes at a channel offset / into the zero-padded bottleneck layout
//   tail_pool_x2        relu(bn(x)) -> average pool -> fp32 features (channel-major flatten)
#include "tn_precise.h"

#include "tn_common.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

__device__ __forceinline__ void split2(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(v);
  lo = __float2bfloat16(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join2(__nv_bfloat16 hi, __nv_bfloat16 lo) { return __bfloat162float(hi) + __bfloat162float(lo); }

// 8 channels of a plane pair <-> 8 floats
__device__ __forceinline__ void load8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&x)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo));
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 h = unpack_bf16x2(aw[j]), l = unpack_bf16x2(bw[j]);
    x[2 * j] = h.x + l.x;
    x[2 * j + 1] = h.y + l.y;
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&x)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h0, l0, h1, l1;
    split2(x[2 * j], h0, l0);
    split2(x[2 * j + 1], h1, l1);
    hw[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    lw[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

template <typename SrcT, bool NHWC>
__global__ void s2d_convert_x2_kernel(const SrcT* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t plane, int n, int H,
                                      int W, int Hz, int Wz, float s0, float s1, float s2, float b0, float b1, float b2) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(n) * Hz * Wz;
  if (i >= total) return;
  const int Xp = static_cast<int>(i % Wz);
  const size_t t = i / Wz;
  const int Yp = static_cast<int>(t % Hz);
  const size_t f = t / Hz;
  const float sc[3] = {s0, s1, s2}, sh[3] = {b0, b1, b2};
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = 0.f;
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
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    a[k] = v[k];
    b[k] = v[8 + k];
  }
  store8(out + i * 16, out + plane + i * 16, a);
  store8(out + i * 16 + 8, out + plane + i * 16 + 8, b);
}

// in: fp32 (n, Hz, Wz, C) of which (Hs, Ws) is the valid stem output (already +shift, ReLU) -> max 3x3 / stride 2 / pad 1
__global__ void maxpool_f32_split_kernel(const float* __restrict__ in, int Hz, int Wz, int Hs, int Ws, int C, int Hp, int Wp,
                                         __nv_bfloat16* __restrict__ out, size_t plane, int out_cstride, size_t total) {
  const int cg = C / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  const int ox = static_cast<int>(t % Wp);
  t /= Wp;
  const int oy = static_cast<int>(t % Hp);
  const size_t f = t / Hp;
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int dy = 0; dy < 3; ++dy) {
    const int iy = oy * 2 - 1 + dy;
    if (iy < 0 || iy >= Hs) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int ix = ox * 2 - 1 + dx;
      if (ix < 0 || ix >= Ws) continue;
      const float4* p = reinterpret_cast<const float4*>(in + ((f * Hz + iy) * Wz + ix) * C + c8 * 8);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
      m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
    }
  }
  __nv_bfloat16* dst = out + ((f * Hp + oy) * Wp + ox) * out_cstride + c8 * 8;
  store8(dst, dst + plane, m);
}

// relu(x * scale + shift) of channels [0, C) of a plane pair with channel stride `cstride` -> operand planes with row pitch `opitch`
__global__ void bn_relu_split_kernel(const __nv_bfloat16* __restrict__ in, size_t in_plane, int C, int cstride,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     __nv_bfloat16* __restrict__ out, size_t out_plane, int opitch, size_t total) {
  const int cg = C / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % cg);
  const size_t pix = i / cg;
  float x[8];
  const __nv_bfloat16* src = in + pix * cstride + c8 * 8;
  load8(src, src + in_plane, x);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = fmaxf(fmaf(x[j], __ldg(scale + c8 * 8 + j), __ldg(shift + c8 * 8 + j)), 0.f);
  __nv_bfloat16* dst = out + pix * opitch + c8 * 8;
  store8(dst, dst + out_plane, x);
}

__global__ void bn_relu_pool2_x2_kernel(const __nv_bfloat16* __restrict__ in, size_t in_plane, int H, int W, int C, int cstride,
                                        int Ho, int Wo, const float* __restrict__ scale, const float* __restrict__ shift,
                                        __nv_bfloat16* __restrict__ out, size_t out_plane, int opitch, size_t total) {
  const int cg = C / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  const size_t opix = t;
  const int ox = static_cast<int>(t % Wo);
  t /= Wo;
  const int oy = static_cast<int>(t % Ho);
  const size_t f = t / Ho;
  float sc[8], sh[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(scale + c8 * 8 + j);
    sh[j] = __ldg(shift + c8 * 8 + j);
    acc[j] = 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat16* src = in + ((f * H + 2 * oy + (q >> 1)) * W + 2 * ox + (q & 1)) * cstride + c8 * 8;
    float x[8];
    load8(src, src + in_plane, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += fmaxf(fmaf(x[j], sc[j], sh[j]), 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
  __nv_bfloat16* dst = out + opix * opitch + c8 * 8;
  store8(dst, dst + out_plane, acc);
}

// fp32 GEMM rows enumerate a (Hg, Wg) grid per frame; rows with (y, x) in [y0, y0+Ho) x [x0, x0+Wo) are written to the pixel
// (f, y-y0, x-x0) of a destination with `pad` zero-border pixels on every side: channels [coff, coff + C), plane pair.
__global__ void split_store_kernel(const float* __restrict__ in, int Hg, int Wg, int y0, int x0, int Ho, int Wo, int C, int pad,
                                   __nv_bfloat16* __restrict__ out, size_t plane, int cstride, int coff, size_t total) {
  const int cg = C / 8;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  const size_t row = t;
  const int x = static_cast<int>(t % Wg);
  t /= Wg;
  const int y = static_cast<int>(t % Hg);
  const size_t f = t / Hg;
  if (y < y0 || y >= y0 + Ho || x < x0 || x >= x0 + Wo) return;
  const float4* p = reinterpret_cast<const float4*>(in + row * C + c8 * 8);
  const float4 a = __ldg(p), b = __ldg(p + 1);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const size_t opix = (f * (Ho + 2 * pad) + (y - y0) + pad) * (Wo + 2 * pad) + (x - x0) + pad;
  __nv_bfloat16* dst = out + opix * cstride + coff + c8 * 8;
  store8(dst, dst + plane, v);
}

__global__ void tail_pool_x2_kernel(const __nv_bfloat16* __restrict__ in, size_t plane, int n, int H, int W, int C, int cstride,
                                    int kh, int kw, int ph, int pw, const float* __restrict__ scale,
                                    const float* __restrict__ shift, float* __restrict__ feats,
                                    __nv_bfloat16* __restrict__ feats_bf16) {
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
  for (int dy = 0; dy < kh; ++dy)
    for (int dx = 0; dx < kw; ++dx) {
      const __nv_bfloat16* src = in + ((f * H + py * kh + dy) * W + px * kw + dx) * cstride + c8 * 8;
      float x[8];
      load8(src, src + plane, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += fmaxf(fmaf(x[j], sc[j], sh[j]), 0.f);
    }
  const float inv = 1.f / static_cast<float>(kh * kw);
  const int pp = ph * pw;
  const size_t D = static_cast<size_t>(C) * pp;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v = acc[j] * inv;
    const size_t o = f * D + static_cast<size_t>(c8 * 8 + j) * pp + py * pw + px;
    feats[o] = v;
    if (feats_bf16) feats_bf16[o] = __float2bfloat16(v);
  }
}

inline unsigned blocks(size_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace

cudaError_t launch_s2d_convert_x2(const void* in, int is_u8_nhwc, __nv_bfloat16* out, size_t plane, int n, int h, int w, int Hz,
                                  int Wz, const float* scale3, const float* shift3, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * Hz * Wz;
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  if (is_u8_nhwc)
    s2d_convert_x2_kernel<uint8_t, true><<<blocks(total, 256), 256, 0, st>>>(static_cast<const uint8_t*>(in), out, plane, n, h, w, Hz,
                                                                            Wz, scale3[0], scale3[1], scale3[2], shift3[0],
                                                                            shift3[1], shift3[2]);
  else
    s2d_convert_x2_kernel<float, false><<<blocks(total, 256), 256, 0, st>>>(static_cast<const float*>(in), out, plane, n, h, w, Hz, Wz,
                                                                           scale3[0], scale3[1], scale3[2], shift3[0], shift3[1],
                                                                           shift3[2]);
  return cudaGetLastError();
}

cudaError_t launch_maxpool_f32_split(const float* in, int n, int Hz, int Wz, int Hs, int Ws, int C, int Hp, int Wp,
                                     __nv_bfloat16* out, size_t plane, int out_cstride, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * Hp * Wp * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  maxpool_f32_split_kernel<<<blocks(total, 256), 256, 0, st>>>(in, Hz, Wz, Hs, Ws, C, Hp, Wp, out, plane, out_cstride, total);
  return cudaGetLastError();
}

cudaError_t launch_bn_relu_split(const __nv_bfloat16* in, size_t in_plane, size_t npix, int C, int cstride, const float* scale,
                                 const float* shift, __nv_bfloat16* out, size_t out_plane, int opitch, cudaStream_t st) {
  const size_t total = npix * (C / 8);
  if (total == 0) return cudaSuccess;
  if ((C % 8) != 0 || (cstride % 8) != 0 || (opitch % 8) != 0) return cudaErrorInvalidValue;
  ProfScope prof_scope(kProfOther, st);
  bn_relu_split_kernel<<<blocks(total, 256), 256, 0, st>>>(in, in_plane, C, cstride, scale, shift, out, out_plane, opitch, total);
  return cudaGetLastError();
}

cudaError_t launch_bn_relu_pool2_x2(const __nv_bfloat16* in, size_t in_plane, int n, int H, int W, int C, int cstride,
                                    const float* scale, const float* shift, __nv_bfloat16* out, size_t out_plane, int opitch,
                                    cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(n) * Ho * Wo * (C / 8);
  if (total == 0) return cudaSuccess;
  if ((C % 8) != 0 || (cstride % 8) != 0 || (opitch % 8) != 0) return cudaErrorInvalidValue;
  ProfScope prof_scope(kProfOther, st);
  bn_relu_pool2_x2_kernel<<<blocks(total, 256), 256, 0, st>>>(in, in_plane, H, W, C, cstride, Ho, Wo, scale, shift, out, out_plane,
                                                             opitch, total);
  return cudaGetLastError();
}

cudaError_t launch_split_store(const float* in, int n, int Hg, int Wg, int y0, int x0, int Ho, int Wo, int C, int pad,
                               __nv_bfloat16* out, size_t plane, int cstride, int coff, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * Hg * Wg * (C / 8);
  if (total == 0) return cudaSuccess;
  if ((C % 8) != 0 || (cstride % 8) != 0 || (coff % 8) != 0) return cudaErrorInvalidValue;
  ProfScope prof_scope(kProfOther, st);
  split_store_kernel<<<blocks(total, 256), 256, 0, st>>>(in, Hg, Wg, y0, x0, Ho, Wo, C, pad, out, plane, cstride, coff, total);
  return cudaGetLastError();
}

cudaError_t launch_tail_pool_x2(const __nv_bfloat16* in, size_t plane, int n, int H, int W, int C, int cstride, int kh, int kw,
                                int ph, int pw, const float* scale, const float* shift, float* feats, __nv_bfloat16* feats_bf16,
                                cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * ph * pw * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  tail_pool_x2_kernel<<<blocks(total, 128), 128, 0, st>>>(in, plane, n, H, W, C, cstride, kh, kw, ph, pw, scale, shift, feats,
                                                         feats_bf16);
  return cudaGetLastError();
}

}  // namespace tn
