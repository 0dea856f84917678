// 3x3 / stride 1 / pad 1 convolution with C_in = C_out = 64 or 128: the stride-1 convolutions of ResNet-18 v2's first two stages
// (gluoncv BasicBlockV2 at 56 x 56 / 28 x 28, reference train.py:32 `--backbone resnet18_v2`, the scripts' default backbone): 49 %
// of that network's forward time while they ran through the generic gather-mode GEMM (profiles/r2_summary.md section 8).
//
// The input is a PRE-ACTIVATED, ZERO-PADDED bf16 tensor (F, H+2, W+2, C) = a 2-D matrix of rows, so a filter tap (dy, dx) is a
// row shift by dy*(W+2) + dx.  One CTA TMA-loads one halo tile (128 + 2(W+2) + 2 rows, one box per 64-channel half) per 128
// outputs and issues 9 taps x NH halves x 4 (K = 16) UMMAs of M128 x N64 into ONE accumulator: the A operand of every tap is the
// same shared-memory tile through a descriptor advanced by dy*(W+2) + dx rows (SWIZZLE_128B is a function of the absolute
// shared-memory address, tools/umma_shift_probe.cu, so any row shift is legal).  The first build stacked the three dx taps along N
// (M128 x N192, a third of the MMAs) and applied the dx shift to accumulator rows in the epilogue like the DenseNet growth conv; with
// 64 output channels that epilogue (64 shuffles + selects per thread and tile) was instruction-issue bound at 2.3 us per tile, three
// different warp organisations of it ran at the same speed (profiles/r2_summary.md section 8).  N = 64 taps cost 50 % more tensor
// time and leave an epilogue of one TMEM load, the per-channel math and the stores.
// Weights (9 taps x NH x 64 x 64 bf16 = 72 KB per 64 input channels, the FOLLOWING BatchNorm's scale folded in for conv1) stay
// resident; for C = 128 the output channels are split over CTA parity.  Up to four halo buffers, accumulators double-buffered in
// TMEM, two epilogue warp groups (one per accumulator buffer).
//
// Epilogue variants (pre-activation ResNet: bn1 -> relu -> conv1 -> bn2 -> relu -> conv2 -> + x):
//   conv1:  t = relu(acc + shift2)                      -> padded, activated input of conv2 (border rows written as zeros)
//   conv2:  y = acc + x (residual, unpadded raw)        -> the block output (unpadded, raw)
//           a = relu(scale1' * bf16(y) + shift1')       -> optionally, the padded activated input of the NEXT block's conv1
// so that no convolution of the stage needs a prologue and no activation makes an extra trip through HBM.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tn_common.h"
#include "tn_conv3x3_c64.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

constexpr int kThreads = 576;          // TMA warp, MMA warp, 2 groups of 8 epilogue warps (one group per accumulator buffer)
constexpr int kTileRows = 128;         // outputs per tile (every accumulator row is an output)
constexpr int kC = 64;
constexpr int kN = kC;                 // UMMA N: the 64 output channels of this CTA
constexpr int kWBlob = kN * 128;       // one (tap, c_in half) blob: 64 rows x 64 bf16, swizzled (8 KB)
constexpr int kWBytes = 9 * kWBlob;    // 72 KB per 64 input channels
constexpr int kTmemCols = 128;         // 2 x 64
constexpr int kMaxBuf = 4;
constexpr int kMaxSmem = 227 * 1024;

struct C64Params {
  int Wp, HpWp, H, W, NR;
  int RH, a_half_bytes, a_bytes, nbuf, num_tiles;
  int cout;                       // 64, or 128: two N halves of 64 handled by the even / odd CTAs
  const uint8_t* wpack;
  const float* shift;             // per output channel, added before relu / residual (nullable)
  int relu;
  const __nv_bfloat16* res;       // unpadded residual (F, H, W, res_cs) or null
  int res_cs;
  __nv_bfloat16* out;             // unpadded output (F, H, W, out_cs) or null
  int out_cs;
  __nv_bfloat16* out_pad;         // padded output (F, H+2, W+2, cout) or null
  const float* act_scale;         // when both outputs are written: out_pad = relu(act_scale * bf16(value) + act_shift)
  const float* act_shift;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// NH = 64-channel halves of C_in (1: 64 -> 64, 2: 128 -> 128 with the output channels split over CTA parity)
template <int NH>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmap, const C64Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                          // NH x 72 KB
  uint8_t* sA = smem + NH * kWBytes;           // nbuf x a_bytes
  const int nsplit = p.cout / kC;              // CTAs per tile (one per 64 output channels)
  const int n_half = nsplit == 2 ? (blockIdx.x & 1) : 0;
  const int tile0 = nsplit == 2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int tstep = nsplit == 2 ? (gridDim.x >> 1) : gridDim.x;
  uint8_t* tail = sA + p.nbuf * p.a_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);  // [kMaxBuf]
  uint64_t* a_empty = a_full + kMaxBuf;                  // [kMaxBuf]
  uint64_t* acc_full = a_empty + kMaxBuf;                // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* w_full = acc_empty + 2;                      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x;
  griddep_launch_dependents();
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kMaxBuf; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);  // one arrive per epilogue warp
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // weights are constants: fetch them before waiting on the producer grid
    mbar_arrive_expect_tx(w_full, NH * kWBytes);
    const uint8_t* wsrc = p.wpack + static_cast<size_t>(n_half) * NH * kWBytes;  // [n half][dy][c_in half]
    for (int b = 0; b < 9 * NH; ++b) bulk_g2s(sW + b * kWBlob, wsrc + b * kWBlob, kWBlob, w_full);
  }
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = tile0; t < p.num_tiles; t += tstep, ++it) {
        const int buf = it % p.nbuf;
        const int use = it / p.nbuf;
        mbar_wait(&a_empty[buf], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[buf], static_cast<uint32_t>(NH * p.RH * 128));
        // one box of RH rows x 64 channels per half (tile row 0 = padded row q0 - (W+2) - 1); rows outside the tensor are
        // zero-filled by the TMA engine.  Splitting the box or using fewer buffers changes nothing (measured).
        for (int half = 0; half < NH; ++half)
          tma_load_2d(smem_u32(sA + buf * p.a_bytes + half * p.a_half_bytes), &tmap, half * 64, t * kTileRows - 1 - p.Wp,
                      &a_full[buf]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(kN);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int t = tile0; t < p.num_tiles; t += tstep, ++it) {
        const int buf = it % p.nbuf;
        const int ab = it & 1;
        mbar_wait(&a_full[buf], (it / p.nbuf) & 1);
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kN;
        const uint32_t a_base = smem_u32(sA + buf * p.a_bytes);
        const uint32_t w_base = smem_u32(sW);
        uint32_t acc = 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
            for (int half = 0; half < NH; ++half) {
              // tap (dy, dx) of output row m reads tile row m + dy*(W+2) + dx
              const uint64_t da = umma_desc_sw128(a_base + half * p.a_half_bytes + (dy * p.Wp + dx) * 128);
              const uint64_t db = umma_desc_sw128(w_base + ((dy * 3 + dx) * NH + half) * kWBlob);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
                acc = 1;
              }
            }
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[ab]);
      }
    }
  } else {
    // epilogue: warps 2..17 = two groups of eight; group g drains accumulator buffer g (tiles it = g, g+2, ...).  Within a group:
    // lane quarter qw = warp & 3, output channels [32 hf, 32 hf + 32) of this CTA's 64.  One TMEM load, the per-channel math and
    // lane-per-row 16-byte global accesses.  (A shared-memory transposed variant with four lanes per row segment, and three warp
    // organisations of the earlier dx-stacked epilogue, all ran at the same speed: the tile period is set by the shared-memory
    // data pipe that the tensor core's operand reads, the TMA writes and the LSU share -- profiles/r2_summary.md section 8.)
    const int grp = (warp - 2) >> 3;
    const int qw = warp & 3;
    const int hf = ((warp - 2) >> 2) & 1;
    const int r = qw * 32 + lane;
    const int cb32 = n_half * kC + hf * 32;
    int it = 0;
    for (int t = tile0; t < p.num_tiles; t += tstep, ++it) {
      const int ab = it & 1;
      if (ab != grp) continue;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is warp-collective; the store branch below diverges
      tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + ab * kN + hf * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);  // accumulator drained: the MMA warp may overwrite it
      const int q = t * kTileRows + r;
      if (q >= p.NR) continue;
      const int f = q / p.HpWp;
      const int rem = q - f * p.HpWp;
      const int yp = rem / p.Wp;
      const int xp = rem - yp * p.Wp;
      const bool interior = yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
      uint32_t packed[16];
      if (interior) {
        const size_t urow = static_cast<size_t>(f * p.H + yp - 1) * p.W + (xp - 1);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + (p.shift ? __ldg(p.shift + cb32 + j) : 0.f);
        if (p.res) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.res + urow * p.res_cs + cb32);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 rv = __ldg(r4 + c);
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f2 = unpack_bf16x2(w[e]);
              o[8 * c + 2 * e] += f2.x;
              o[8 * c + 2 * e + 1] += f2.y;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = fmaxf(o[j], 0.f);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(o[2 * j], o[2 * j + 1]);
        if (p.out) {  // raw block output, unpadded rows
          uint4* dst = reinterpret_cast<uint4*>(p.out + urow * p.out_cs + cb32);
#pragma unroll
          for (int c = 0; c < 4; ++c) dst[c] = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        }
        if (p.out_pad && p.act_scale) {  // the next block's activation, from the ROUNDED block output
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 f2 = unpack_bf16x2(packed[j]);
            const float a = fmaxf(fmaf(f2.x, __ldg(p.act_scale + cb32 + 2 * j), __ldg(p.act_shift + cb32 + 2 * j)), 0.f);
            const float b = fmaxf(fmaf(f2.y, __ldg(p.act_scale + cb32 + 2 * j + 1), __ldg(p.act_shift + cb32 + 2 * j + 1)), 0.f);
            packed[j] = pack_bf16x2(a, b);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = 0u;  // border rows of the padded output stay zero
      }
      if (p.out_pad) {
        uint4* dst = reinterpret_cast<uint4*>(p.out_pad + static_cast<size_t>(q) * p.cout + cb32);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// out_pad[(f, y+1, x+1)][c] = relu(scale[c] * x[(f, y, x)][c] + shift[c]), zeros on the border; 8 channels per thread
__global__ void bn_relu_pad_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int F, int H, int W, int C,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   __nv_bfloat16* __restrict__ out) {
  const int cg = C / 8;
  const int Hp = H + 2, Wp = W + 2;
  const size_t total = static_cast<size_t>(F) * Hp * Wp * cg;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c8 = static_cast<int>(i % cg);
    const size_t row = i / cg;
    const int xp = static_cast<int>(row % Wp);
    const size_t t = row / Wp;
    const int yp = static_cast<int>(t % Hp);
    const size_t f = t / Hp;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (xp >= 1 && xp <= W && yp >= 1 && yp <= H) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + ((f * H + yp - 1) * W + (xp - 1)) * x_cs + c8 * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f2 = unpack_bf16x2(w[e]);
        const int c = c8 * 8 + 2 * e;
        pk[e] = pack_bf16x2(fmaxf(fmaf(f2.x, scale[c], shift[c]), 0.f), fmaxf(fmaf(f2.y, scale[c + 1], shift[c + 1]), 0.f));
      }
      o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    *reinterpret_cast<uint4*>(out + row * C + c8 * 8) = o;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

namespace {
constexpr int kFixedSmem = 256 + 1024;  // barriers, alignment slack
}
bool conv3x3_c64_supported(int H, int W, int C) {
  if (C != 64 && C != 128) return false;
  const int RH = 128 + 2 * (W + 2) + 2;
  const int NH = C / 64;
  const size_t a_bytes = NH * align_up(static_cast<size_t>(RH) * 128, 1024);
  // one TMA box (<= 256 rows) per halo half; resident weights + at least one halo buffer must fit
  return H >= 1 && W >= 1 && RH <= 256 && NH * kWBytes + kFixedSmem + a_bytes <= static_cast<size_t>(kMaxSmem);
}

bool make_conv3x3_c64(DeviceArena& arena, const float* w /* (C,C,3,3) OIHW */, int C, const float* fold_scale, Conv3x3C64Dev* out) {
  if (C != 64 && C != 128) return false;
  const int NH = C / 64, NS = C / 64;  // input halves, output halves
  std::vector<uint8_t> blob(static_cast<size_t>(NS) * NH * kWBytes, 0);
  for (int ns = 0; ns < NS; ++ns)
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx)
        for (int half = 0; half < NH; ++half)
          for (int n = 0; n < kN; ++n) {
            const int c = ns * kC + n;
            for (int kk = 0; kk < 64; ++kk) {
              const int ci = half * 64 + kk;
              float v = w[((static_cast<size_t>(c) * C + ci) * 3 + dy) * 3 + dx];
              if (fold_scale) v *= fold_scale[c];
              __nv_bfloat16 b = __float2bfloat16(v);
              const size_t off = ((static_cast<size_t>(ns) * 9 + dy * 3 + dx) * NH + half) * kWBlob + n * 128 +
                                 (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
              memcpy(&blob[off], &b, 2);
            }
          }
  out->wpack = static_cast<const uint8_t*>(arena.upload(blob.data(), blob.size()));
  out->C = C;
  return out->wpack != nullptr;
}

cudaError_t launch_bn_relu_pad(const __nv_bfloat16* x, int x_cs, int F, int H, int W, int C, const float* scale, const float* shift,
                               __nv_bfloat16* out_pad, cudaStream_t st) {
  if (C % 8) return cudaErrorInvalidValue;
  const size_t total = static_cast<size_t>(F) * (H + 2) * (W + 2) * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  bn_relu_pad_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(x, x_cs, F, H, W, C, scale, shift, out_pad);
  return cudaGetLastError();
}

cudaError_t launch_conv3x3_c64(const Conv3x3C64Dev& cv, const __nv_bfloat16* in_padded, int F, int H, int W, const float* shift,
                               int relu, const __nv_bfloat16* res, int res_cs, __nv_bfloat16* out, int out_cs,
                               __nv_bfloat16* out_pad, const float* act_scale, const float* act_shift, int num_sms,
                               cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  const int C = cv.C, NH = C / 64;
  if (!conv3x3_c64_supported(H, W, C) || (!out && !out_pad)) return cudaErrorInvalidValue;
  C64Params p;
  memset(&p, 0, sizeof(p));
  p.cout = C;
  p.Wp = W + 2;
  p.HpWp = (H + 2) * (W + 2);
  p.H = H;
  p.W = W;
  const long long NR = static_cast<long long>(F) * p.HpWp;
  if (NR >= (1LL << 31) - 256) return cudaErrorInvalidValue;
  p.NR = static_cast<int>(NR);
  p.RH = 128 + 2 * p.Wp + 2;
  p.a_half_bytes = static_cast<int>(align_up(static_cast<size_t>(p.RH) * 128, 1024));
  p.a_bytes = NH * p.a_half_bytes;
  p.num_tiles = static_cast<int>((NR + kTileRows - 1) / kTileRows);
  const int fixed = NH * kWBytes + kFixedSmem;
  int nbuf = (kMaxSmem - fixed) / p.a_bytes;
  if (nbuf > kMaxBuf) nbuf = kMaxBuf;
  if (nbuf < 1) return cudaErrorInvalidValue;
  p.nbuf = nbuf;
  p.wpack = cv.wpack;
  p.shift = shift;
  p.relu = relu;
  p.res = res;
  p.res_cs = res_cs;
  p.out = out;
  p.out_cs = out_cs;
  p.out_pad = out_pad;
  p.act_scale = (out && out_pad) ? act_scale : nullptr;
  p.act_shift = (out && out_pad) ? act_shift : nullptr;
  if (out && out_pad && (!act_scale || !act_shift)) return cudaErrorInvalidValue;
  const int smem = fixed + p.nbuf * p.a_bytes;

  CUtensorMap tmap;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(p.NR)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(C) * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.RH)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(in_padded), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  static int configured_smem[2] = {0, 0};
  if (smem > configured_smem[NH - 1]) {
    cudaError_t e = NH == 1 ? cudaFuncSetAttribute(conv3x3_c64_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                            : cudaFuncSetAttribute(conv3x3_c64_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_smem[NH - 1] = smem;
  }
  ProfScope prof_scope(kProfConvGemm, st);
  if (NH == 1) {
    const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
    return launch_pdl(conv3x3_c64_kernel<1>, dim3(grid), dim3(kThreads), smem, st, tmap, p);
  }
  int pairs = num_sms / 2;  // even CTAs: output channels [0, 64), odd CTAs: [64, 128)
  if (pairs > p.num_tiles) pairs = p.num_tiles;
  return launch_pdl(conv3x3_c64_kernel<2>, dim3(2 * pairs), dim3(kThreads), smem, st, tmap, p);
}

}  // namespace tn
