// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One kernel covers every contraction on the CNN path (SURVEY.md §2.2 K1/K2 and the GRU input
// projection of K4):  D[M = pixels, N = C_out] = A[M, K = taps*C_in] * W[N, K]^T
//   * A is never materialised: 128 producer threads gather the NHWC bf16 activation rows for one
//     (tap, 64-channel) K-chunk straight from HBM/L2 with 16-byte loads, apply the *consumer's*
//     pre-activation BatchNorm+ReLU in registers (DenseNet/ResNet-v2 put BN+ReLU in front of every
//     conv, so it cannot be folded into the producer layer), and write the chunk into shared memory in
//     the UMMA 128-byte-swizzled K-major layout.
//   * W is pre-packed on the host into exactly that shared-memory image, one contiguous blob per
//     K-chunk, and brought in with a single bulk async copy (TMA engine, UBLKCP) that signals an
//     mbarrier with its byte count.
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x (chunk/16) and commits to the
//     stage's "empty" mbarrier; the ring is NSTAGE deep.
//   * epilogue: tcgen05.ld the fp32 accumulators, apply the folded BN2 scale/shift (+ReLU) or bias,
//     optional residual add, and store bf16/fp32 NHWC at a channel offset -> DenseNet's concat is
//     "write your 32 channels into the block buffer in place".
//
// Modes: CONV (generic RxS, stride, zero padding), POOL2 (transition: 2x2 average of the activated
// input feeds a 1x1 conv -- avg-pool and a 1x1 conv commute, so the GEMM runs on 4x fewer rows),
// STEM (7x7/2 on a channel-padded NHWC4 image; K-chunk = two filter rows of 8 px * 4 ch).
#include "tn_conv_gemm.h"
#include "tn_common.h"
#include "tn_ptx.cuh"

#include <stdio.h>

namespace tn {

namespace {

constexpr int kThreads = 128;
constexpr int kBM = 128;
constexpr int kABytes = kBM * 128;  // 128 rows x 64 bf16

template <int BN>
struct Cfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kNStage = (BN >= 256) ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr int kSmem = kNStage * kStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg64(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// y = relu?(x*scale+shift) on 8 packed bf16
__device__ __forceinline__ uint4 bn_act8(uint4 x, const float (&sc)[8], const float (&sh)[8], bool relu) {
  uint32_t in[4] = {x.x, x.y, x.z, x.w};
  uint32_t out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = unpack_bf16x2(in[j]);
    float a = fmaf(f.x, sc[2 * j], sh[2 * j]);
    float b = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
    if (relu) {
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    }
    out[j] = pack_bf16x2(a, b);
  }
  return make_uint4(out[0], out[1], out[2], out[3]);
}
__device__ __forceinline__ void bn_act8_accum(uint4 x, const float (&sc)[8], const float (&sh)[8], bool relu,
                                              float (&acc)[8]) {
  uint32_t in[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = unpack_bf16x2(in[j]);
    float a = fmaf(f.x, sc[2 * j], sh[2 * j]);
    float b = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
    if (relu) {
      a = fmaxf(a, 0.f);
      b = fmaxf(b, 0.f);
    }
    acc[2 * j] += a;
    acc[2 * j + 1] += b;
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads) conv_gemm_kernel(const ConvGemmParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kNStage * C::kStage);
  uint64_t* empty_bar = full_bar + C::kNStage;
  uint64_t* accum_bar = empty_bar + C::kNStage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < C::kNStage; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m0 = blockIdx.x * kBM;
  const int n_tile = blockIdx.y;
  const uint8_t* wtile = p.wpack + static_cast<size_t>(n_tile) * p.num_chunks * C::kBBytes;

  // ---- per-thread producer geometry: rows (tid>>3)+16*i, 16-byte K-group g = tid&7
  const int g = tid & 7;
  const int rbase = tid >> 3;
  const uint32_t a_off = static_cast<uint32_t>(rbase * 128 + ((g ^ (rbase & 7)) << 4));  // + i*2048
  int pix[8];  // input pixel index of the (0,0) tap
  int pos[8];  // (iy0 << 16) | (ix0 & 0xffff)
  {
    const int hw = p.Ho * p.Wo;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + rbase + 16 * i;
      if (m < p.M) {
        const int f = m / hw;
        const int rem = m - f * hw;
        const int oy = rem / p.Wo;
        const int ox = rem - oy * p.Wo;
        const int iy0 = oy * p.stride - p.pad;
        const int ix0 = ox * p.stride - p.pad;
        pix[i] = (f * p.H + iy0) * p.W + ix0;
        pos[i] = (iy0 << 16) | (ix0 & 0xffff);
      } else {
        pix[i] = 0;
        pos[i] = (-20000 << 16);  // every tap out of bounds -> zero row
      }
    }
  }

  const uint32_t idesc = umma_idesc_bf16_m128(BN);
  const bool has_pro = p.pro_scale != nullptr;
  const bool pro_relu = p.pro_relu != 0;

  uint4 regs[8];
  uint32_t okmask = 0;

  // Issue the global loads of chunk c into registers (MODE CONV / STEM); consumed one iteration later.
  auto load_chunk = [&](int c) {
    okmask = 0;
    if (MODE == kModeConv) {
      const int tap = c / p.chunks_per_tap;
      const int kc = (c - tap * p.chunks_per_tap) * 64;
      const int r = tap / p.S;
      const int s = tap - r * p.S;
      const int ch = kc + g * 8;
      const bool ch_ok = ch < p.Cin;
      const int dpix = r * p.W + s;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int iy = (pos[i] >> 16) + r;
        const int ix = static_cast<int>(static_cast<short>(pos[i] & 0xffff)) + s;
        const bool ok = ch_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        if (ok) {
          regs[i] = ldg128(p.in + static_cast<size_t>(pix[i] + dpix) * p.in_cstride + ch);
          okmask |= 1u << i;
        } else {
          regs[i] = make_uint4(0, 0, 0, 0);
        }
      }
    } else if (MODE == kModeStem) {
      // chunk c = filter rows 2c, 2c+1 ; group g -> row 2c+(g>>2), pixels ix0+2*(g&3), +1 ; 4 ch (8 B) per pixel
      const int r = 2 * c + (g >> 2);
      const int s = 2 * (g & 3);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int iy = (pos[i] >> 16) + r;
        const int ix = static_cast<int>(static_cast<short>(pos[i] & 0xffff)) + s;
        const bool row_ok = r < p.R && iy >= 0 && iy < p.H;
        uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
        const __nv_bfloat16* src = p.in + static_cast<size_t>(pix[i] + r * p.W + s) * 4;
        if (row_ok && ix >= 0 && ix < p.W) a = ldg64(src);
        if (row_ok && ix + 1 >= 0 && ix + 1 < p.W) b = ldg64(src + 4);
        regs[i] = make_uint4(a.x, a.y, b.x, b.y);
      }
    }
  };

  auto store_chunk = [&](int c, uint32_t a_stage) {
    if (MODE == kModeConv) {
      float sc[8], sh[8];
      if (has_pro) {
        const int kc = (c % p.chunks_per_tap) * 64;
        const int ch = kc + g * 8;
        if (ch < p.Cin) {
          const float4* s4 = reinterpret_cast<const float4*>(p.pro_scale + ch);
          const float4* h4 = reinterpret_cast<const float4*>(p.pro_shift + ch);
          float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1), h0 = __ldg(h4), h1 = __ldg(h4 + 1);
          sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
          sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { sc[j] = 0.f; sh[j] = 0.f; }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 v = regs[i];
        if (has_pro && ((okmask >> i) & 1u)) v = bn_act8(v, sc, sh, pro_relu);
        sts128(a_stage + a_off + i * 2048, v);
      }
    } else if (MODE == kModeStem) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sts128(a_stage + a_off + i * 2048, regs[i]);
    } else {  // kModePool2: 1x1 conv on the 2x2 average of the activated input (R=S=1, stride 2 geometry)
      const int kc = c * 64;
      const int ch = kc + g * 8;
      const bool ch_ok = ch < p.Cin;
      float sc[8], sh[8];
      if (ch_ok) {
        const float4* s4 = reinterpret_cast<const float4*>(p.pro_scale + ch);
        const float4* h4 = reinterpret_cast<const float4*>(p.pro_shift + ch);
        float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1), h0 = __ldg(h4), h1 = __ldg(h4 + 1);
        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = 0.f; sh[j] = 0.f; }
      }
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        uint4 out = make_uint4(0, 0, 0, 0);
        if (ch_ok && (pos[i] >> 16) >= 0) {
          const __nv_bfloat16* src = p.in + static_cast<size_t>(pix[i]) * p.in_cstride + ch;
          const size_t rowstep = static_cast<size_t>(p.W) * p.in_cstride;
          uint4 v00 = ldg128(src), v01 = ldg128(src + p.in_cstride);
          uint4 v10 = ldg128(src + rowstep), v11 = ldg128(src + rowstep + p.in_cstride);
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          bn_act8_accum(v00, sc, sh, pro_relu, acc);
          bn_act8_accum(v01, sc, sh, pro_relu, acc);
          bn_act8_accum(v10, sc, sh, pro_relu, acc);
          bn_act8_accum(v11, sc, sh, pro_relu, acc);
          out.x = pack_bf16x2(acc[0] * 0.25f, acc[1] * 0.25f);
          out.y = pack_bf16x2(acc[2] * 0.25f, acc[3] * 0.25f);
          out.z = pack_bf16x2(acc[4] * 0.25f, acc[5] * 0.25f);
          out.w = pack_bf16x2(acc[6] * 0.25f, acc[7] * 0.25f);
        }
        sts128(a_stage + a_off + i * 2048, out);
      }
    }
  };

  // ---------------------------------------------------------------- main loop over K-chunks
  const int nchunks = p.num_chunks;
  if (MODE != kModePool2) load_chunk(0);
  for (int c = 0; c < nchunks; ++c) {
    const int stage = c % C::kNStage;
    const int round = c / C::kNStage;
    const uint32_t a_stage = smem_base + stage * C::kStage;
    uint8_t* b_stage = smem + stage * C::kStage + kABytes;
    if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1);  // UMMAs that read this stage are done
    if (tid == 0) {
      mbar_arrive_expect_tx(&full_bar[stage], C::kBBytes);
      bulk_g2s(b_stage, wtile + static_cast<size_t>(c) * C::kBBytes, C::kBBytes, &full_bar[stage]);
    }
    store_chunk(c, a_stage);
    if (MODE != kModePool2 && c + 1 < nchunks) load_chunk(c + 1);  // in flight across the sync + MMA issue
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&full_bar[stage], round & 1);  // weights landed
      tc_fence_after();
      int kv;
      if (MODE == kModeStem) {
        kv = (2 * c + 1 < p.R) ? 64 : 32;
      } else {
        const int kc = (c % p.chunks_per_tap) * 64;
        kv = min(64, p.Cin - kc);
      }
      const uint64_t da = umma_desc_sw128(a_stage);
      const uint64_t db = umma_desc_sw128(a_stage + kABytes);
      for (int k = 0; k < kv / 16; ++k) {
        umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);
      if (c == nchunks - 1) umma_commit(accum_bar);
    }
  }

  // ---------------------------------------------------------------- epilogue
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  {
    const int m = m0 + warp * 32 + lane;
    const bool row_ok = m < p.M;
    size_t orow = static_cast<size_t>(m);  // output row (pixel) index
    if (p.out_pad && row_ok) {
      const int hw = p.Ho * p.Wo;
      const int f = m / hw;
      const int rem = m - f * hw;
      const int oy = rem / p.Wo;
      const int ox = rem - oy * p.Wo;
      orow = (static_cast<size_t>(f) * (p.Ho + 2) + oy + 1) * (p.Wo + 2) + ox + 1;
    }
    const int ncol_tile = min(BN, p.Cout - n_tile * BN);
    for (int cb = 0; cb * 32 < ncol_tile; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      const int n0 = n_tile * BN + cb * 32;
      if (row_ok) {
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.epi_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= __ldg(p.epi_scale + n0 + j);
        }
        if (p.epi_shift != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(p.epi_shift + n0 + j);
        }
        if (p.res != nullptr) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.res + static_cast<size_t>(m) * p.res_cstride + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 rv = __ldg(r4 + q);
            uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 t = unpack_bf16x2(w[j]);
              f[q * 8 + 2 * j] += t.x;
              f[q * 8 + 2 * j + 1] += t.y;
            }
          }
        }
        if (p.epi_relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.out_fp32) {
          float4* o = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.out_cstride +
                                                p.out_coff + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
        } else {
          uint4* o = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.out_cstride + p.out_coff + n0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q] = make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                              pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<BN>(tmem_base);
}

template <int BN, int MODE>
cudaError_t launch_t(const ConvGemmParams& p, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool configured = false;  // per instantiation; handles are single-device (see tennis_b200.h)
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv_gemm_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((p.M + kBM - 1) / kBM, (p.Cout + BN - 1) / BN);
  conv_gemm_kernel<BN, MODE><<<grid, kThreads, C::kSmem, stream>>>(p);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_bn(const ConvGemmParams& p, cudaStream_t stream) {
  switch (p.mode) {
    case kModeConv: return launch_t<BN, kModeConv>(p, stream);
    case kModePool2: return launch_t<BN, kModePool2>(p, stream);
    case kModeStem: return launch_t<BN, kModeStem>(p, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace

int conv_gemm_pick_bn(int cout) {
  if (cout <= 32) return 32;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  return 256;
}

size_t conv_gemm_wpack_bytes(int cout, int num_chunks) {
  const int bn = conv_gemm_pick_bn(cout);
  const int ntiles = (cout + bn - 1) / bn;
  return static_cast<size_t>(ntiles) * num_chunks * bn * 128;
}

cudaError_t launch_conv_gemm(const ConvGemmParams& p, cudaStream_t stream) {
  ProfScope prof_scope(kProfConvGemm, stream);
  switch (conv_gemm_pick_bn(p.Cout)) {
    case 32: return launch_bn<32>(p, stream);
    case 64: return launch_bn<64>(p, stream);
    case 128: return launch_bn<128>(p, stream);
    default: return launch_bn<256>(p, stream);
  }
}

}  // namespace tn
