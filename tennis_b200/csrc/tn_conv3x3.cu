// 3x3 / stride 1 / pad 1 convolution, C_in = 128 -> C_out = 32 (the DenseNet growth conv, 44 % of the network's
// FLOPs), as a persistent, warp-specialised, TMA-fed tcgen05 kernel.
//
// Layout trick: the 128-channel bottleneck activation is stored ZERO-PADDED in HBM, (F, H+2, W+2, 128) bf16, i.e.
// as a 2-D matrix [NR = F*(H+2)*(W+2) rows][128 ch].  In padded-flattened row space every filter tap is a pure row
// shift: tap (dy,dx) of output row q reads row q + dy*(W+2) + dx, and image borders need no masks (they read the
// stored zeros).  A CTA therefore
//   * TMA-loads ONE halo tile (rows q0-(W+2) .. q0+128+(W+2), two 64-channel halves, 128B-swizzled) per 126 outputs
//     -- the im2col matrix is never materialised and each activation byte crosses L2->SMEM ~2x instead of 9x;
//   * issues, per tile, 3 (dy) x 2 (halves) x 4 (K=16) UMMAs with M=128, N=96: the A operand of tap-row dy is the
//     SAME shared-memory tile addressed through a descriptor whose start address is advanced by (1+dy)*(W+2) rows
//     (verified on hardware: SWIZZLE_128B is a function of the absolute smem address, tools/umma_shift_probe.cu);
//     the three dx taps are stacked along N (3 x 32), which keeps the SMEM operand traffic per MMA at ~7 KB/48 clk
//     instead of 5 KB/16 clk for an N=32 MMA;
//   * epilogue: out[q] = D[q-1][0:32] + D[q][32:64] + D[q+1][64:96]  (the dx shift is applied on accumulator rows
//     with warp shuffles + a 1 KB smem exchange between the four epilogue warps), cast to bf16 and written into the
//     dense block's concat buffer at its channel offset.  Rows 0 and 127 of a tile have no neighbour: tiles advance
//     by 126 rows.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..9 = epilogue (two per TMEM lane
// quarter, 16 of the 32 output channels each).  TMEM accumulators
// are double-buffered (2 x 96 columns) so the epilogue of tile i overlaps the MMAs of tile i+1; the halo tile is
// double-buffered when it fits.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "tn_common.h"
#include "tn_conv3x3.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

constexpr int kThreads = 320;            // TMA warp, MMA warp, 8 epilogue warps
constexpr int kTileRows = 126;         // valid outputs per tile
constexpr int kN = 96;                 // 3 dx taps x 32 output channels
constexpr int kWBlob = kN * 128;       // one (dy, half) weight blob: 96 rows x 64 bf16, swizzled
constexpr int kWBytes = 6 * kWBlob;    // 72 KB
constexpr int kTmemCols = 256;         // 2 x 96 used
constexpr int kMaxBuf = 4;             // halo buffers: as many as fit beside the weights (2 at 56 x 56, 3 at 28 x 28, 4 below)

struct Conv3x3Params {
  int Wp, HpWp, H, W, NR;     // padded width, padded rows per frame, unpadded dims, total padded rows
  int RH;                     // halo tile rows = 128 + 2*Wp
  int nbox, box_rows;         // TMA boxes per 64-channel half and their height (multiple of 8 when nbox == 2)
  int a_half_bytes;           // nbox*box_rows*128 rounded up to 1024
  int nbuf;                   // 1..kMaxBuf halo buffers
  int l2_ahead;               // TMA L2 prefetch distance in tiles (0 = off)
  int num_tiles;
  const uint8_t* wpack;       // 6 blobs [dy][half]
  __nv_bfloat16* out;
  int out_cstride, out_coff;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                  const Conv3x3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                   // 72 KB
  uint8_t* sA = smem + kWBytes;                         // nbuf x 2 halves x a_half_bytes
  const int a_buf_bytes = 2 * p.a_half_bytes;
  uint8_t* tail = sA + p.nbuf * a_buf_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);  // [kMaxBuf]
  uint64_t* a_empty = a_full + kMaxBuf;                  // [kMaxBuf]
  uint64_t* acc_full = a_empty + kMaxBuf;                // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* w_full = acc_empty + 2;                      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* xch = reinterpret_cast<float*>(tail + 256);     // [2 parity][2 halves][4 quarters][2][16]

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
    mbar_arrive_expect_tx(w_full, kWBytes);
    for (int b = 0; b < 6; ++b) bulk_g2s(sW + b * kWBlob, p.wpack + b * kWBlob, kWBlob, w_full);
  }
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // The two halo buffers (all that fits beside the 72 KB of weights) give one tile of look-ahead.  Optional experiment: an L2
      // prefetch of the tiles p.l2_ahead iterations ahead (TN_3X3_L2_AHEAD) costs no shared memory and turns the later TMA load
      // into an L2 hit -- measured SLOWER at every distance (the kernel is not latency-bound on DRAM), so it is off by default.
      auto prefetch_tile = [&](int t) {
        if (t >= p.num_tiles) return;
        const int row0 = t * kTileRows - 1 - p.Wp;
        for (int half = 0; half < 2; ++half)
          for (int b = 0; b < p.nbox; ++b)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap), "r"(half * 64),
                         "r"(row0 + b * p.box_rows)
                         : "memory");
      };
      for (int a = 1; a < p.l2_ahead; ++a) prefetch_tile(blockIdx.x + a * gridDim.x);
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it % p.nbuf;
        const int use = it / p.nbuf;
        if (p.l2_ahead > 0) prefetch_tile(t + p.l2_ahead * gridDim.x);
        mbar_wait(&a_empty[buf], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[buf], static_cast<uint32_t>(2 * p.nbox * p.box_rows * 128));
        const int row0 = t * kTileRows - 1 - p.Wp;
        for (int half = 0; half < 2; ++half) {
          const uint32_t dst = smem_u32(sA + buf * a_buf_bytes + half * p.a_half_bytes);
          // box = (64 ch, box_rows rows); rows outside the tensor are zero-filled by the TMA engine
          for (int b = 0; b < p.nbox; ++b)
            tma_load_2d(dst + b * p.box_rows * 128, &tmap, half * 64, row0 + b * p.box_rows, &a_full[buf]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(kN);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it % p.nbuf;
        const int ab = it & 1;
        mbar_wait(&a_full[buf], (it / p.nbuf) & 1);
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kN;
        const uint32_t a_base = smem_u32(sA + buf * a_buf_bytes);
        const uint32_t w_base = smem_u32(sW);
        uint32_t acc = 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t da = umma_desc_sw128(a_base + half * p.a_half_bytes + dy * p.Wp * 128);
            const uint64_t db = umma_desc_sw128(w_base + (dy * 2 + half) * kWBlob);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, acc);
              acc = 1;
            }
          }
        }
        umma_commit(&a_empty[buf]);   // halo buffer may be refilled once these MMAs retire
        umma_commit(&acc_full[ab]);   // accumulator ready for the epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9: lane quarter warp&3, channel half)
    const int qw = warp & 3;              // lanes 32*qw .. 32*qw+31
    const int hf = (warp - 2) >> 2;       // output channels [16*hf, 16*hf+16)
    const int r = qw * 32 + lane;         // accumulator row of this thread
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int ab = it & 1;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v0[16], v1[16], v2[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + ab * kN + hf * 16;
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + 32, v1);
      tmem_ld16(taddr + 64, v2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);  // accumulator drained: MMA warp may overwrite it

      // dx-tap exchange across the four lane quarters: quarter qw publishes its last row's dx=-1 block (lane 31) and its first
      // row's dx=+1 block (lane 0) with 4 x STS.128 by two active lanes, and reads its neighbours' rows the same way -- 8 + 4
      // shared-memory wavefronts per warp and tile.  (Round 1 moved these 16 floats with scalar, lane-predicated LDS/STS inside the
      // channel loop: 64 wavefronts per warp, 512 per tile = 22 % of the kernel's L1TEX data-pipe time, profiles/r2_smem_pipe.md.)
      float* x = xch + (it & 1) * 256 + hf * 128;
      if (lane == 0 || lane == 31) {
        uint4* dst = reinterpret_cast<uint4*>(x + (qw * 2 + (lane == 0 ? 1 : 0)) * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint4 w;
          w.x = (lane == 0) ? v2[4 * j4 + 0] : v0[4 * j4 + 0];
          w.y = (lane == 0) ? v2[4 * j4 + 1] : v0[4 * j4 + 1];
          w.z = (lane == 0) ? v2[4 * j4 + 2] : v0[4 * j4 + 2];
          w.w = (lane == 0) ? v2[4 * j4 + 3] : v0[4 * j4 + 3];
          dst[j4] = w;
        }
      }
      if (hf == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      float nb[16];  // lane 0: row above my quarter (dx=-1 block of quarter qw-1's last row); lane 31: row below (dx=+1 block)
      const bool need_up = lane == 0 && qw > 0, need_dn = lane == 31 && qw < 3;
      if (need_up || need_dn) {
        const uint4* src = reinterpret_cast<const uint4*>(x + (need_up ? ((qw - 1) * 2 + 0) : ((qw + 1) * 2 + 1)) * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const uint4 w = src[j4];
          nb[4 * j4 + 0] = __uint_as_float(w.x);
          nb[4 * j4 + 1] = __uint_as_float(w.y);
          nb[4 * j4 + 2] = __uint_as_float(w.z);
          nb[4 * j4 + 3] = __uint_as_float(w.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) nb[j] = 0.f;
      }
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);      // D[r-1][j]
        float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);    // D[r+1][64+j]
        if (need_up) up = nb[j];
        if (need_dn) dn = nb[j];
        o[j] = up + __uint_as_float(v1[j]) + dn;
      }
      const int q = t * kTileRows - 1 + r;
      if (r >= 1 && r <= kTileRows && q < p.NR) {
        const int f = q / p.HpWp;
        const int rem = q - f * p.HpWp;
        const int yp = rem / p.Wp;
        const int xp = rem - yp * p.Wp;
        if (yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(f * p.H + yp - 1) * p.W + (xp - 1)) * p.out_cstride +
                                                p.out_coff + hf * 16);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            dst[c] = make_uint4(pack_bf16x2(o[8 * c], o[8 * c + 1]), pack_bf16x2(o[8 * c + 2], o[8 * c + 3]),
                                pack_bf16x2(o[8 * c + 4], o[8 * c + 5]), pack_bf16x2(o[8 * c + 6], o[8 * c + 7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

// zero the 1-pixel border of a padded (F, Hp, Wp, C) bf16 buffer (interior is overwritten by the producer conv)
__global__ void zero_border_kernel(__nv_bfloat16* buf, int F, int Hp, int Wp, int C) {
  const int per_frame = 2 * Wp + 2 * (Hp - 2);
  const int cg = C / 8;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(F) * per_frame * cg) return;
  const int c8 = static_cast<int>(i % cg);
  size_t t = i / cg;
  const int b = static_cast<int>(t % per_frame);
  const size_t f = t / per_frame;
  int y, x;
  if (b < Wp) { y = 0; x = b; }
  else if (b < 2 * Wp) { y = Hp - 1; x = b - Wp; }
  else { const int k = b - 2 * Wp; y = 1 + (k >> 1); x = (k & 1) ? Wp - 1 : 0; }
  *reinterpret_cast<uint4*>(buf + ((f * Hp + y) * Wp + x) * C + c8 * 8) = make_uint4(0, 0, 0, 0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

constexpr int kMaxSmem = 227 * 1024;

}  // namespace

namespace {
void halo_geometry(int W, int* RH, int* nbox, int* box_rows, int* a_half) {
  *RH = 128 + 2 * (W + 2);
  *nbox = (*RH <= 256) ? 1 : 2;
  *box_rows = (*nbox == 1) ? *RH : static_cast<int>(align_up(static_cast<size_t>((*RH + 1) / 2), 8));
  *a_half = static_cast<int>(align_up(static_cast<size_t>(*nbox) * *box_rows * 128, 1024));
}
}  // namespace

bool conv3x3_halo_supported(int H, int W) {
  int RH, nbox, box_rows, a_half;
  halo_geometry(W, &RH, &nbox, &box_rows, &a_half);
  return box_rows <= 256 && kWBytes + 2 * a_half + 4096 <= kMaxSmem && H >= 1 && W >= 1;
}

bool make_conv3x3(DeviceArena& arena, const float* w /* (32,128,3,3) OIHW */, Conv3x3Dev* out) {
  std::vector<uint8_t> blob(kWBytes, 0);
  for (int dy = 0; dy < 3; ++dy)
    for (int half = 0; half < 2; ++half)
      for (int n = 0; n < kN; ++n) {
        const int dx = n / 32, c = n % 32;
        for (int kk = 0; kk < 64; ++kk) {
          const int ci = half * 64 + kk;
          const float v = w[((static_cast<size_t>(c) * 128 + ci) * 3 + dy) * 3 + dx];
          __nv_bfloat16 b = __float2bfloat16(v);
          const size_t off = static_cast<size_t>(dy * 2 + half) * kWBlob + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
          memcpy(&blob[off], &b, 2);
        }
      }
  out->wpack = static_cast<const uint8_t*>(arena.upload(blob.data(), blob.size()));
  return out->wpack != nullptr;
}

cudaError_t launch_zero_border(__nv_bfloat16* buf, int F, int Hp, int Wp, int C, cudaStream_t st) {
  const size_t total = static_cast<size_t>(F) * (2 * Wp + 2 * (Hp - 2)) * (C / 8);
  if (total == 0) return cudaSuccess;
  ProfScope prof_scope(kProfOther, st);
  zero_border_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(buf, F, Hp, Wp, C);
  return cudaGetLastError();
}

// in_padded: (F, H+2, W+2, 128) bf16 with zero borders; out: (F, H, W, out_cstride) bf16, 32 channels at out_coff.
cudaError_t launch_conv3x3_halo(const Conv3x3Dev& cv, const __nv_bfloat16* in_padded, int F, int H, int W,
                                __nv_bfloat16* out, int out_cstride, int out_coff, int num_sms, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  Conv3x3Params p;
  p.Wp = W + 2;
  p.HpWp = (H + 2) * (W + 2);
  p.H = H;
  p.W = W;
  const long long NR = static_cast<long long>(F) * p.HpWp;
  if (NR >= (1ll << 31) - 1024) return cudaErrorInvalidValue;
  p.NR = static_cast<int>(NR);
  halo_geometry(W, &p.RH, &p.nbox, &p.box_rows, &p.a_half_bytes);
  const int fixed = kWBytes + 4096;
  p.nbuf = (kMaxSmem - fixed) / (2 * p.a_half_bytes);
  if (p.nbuf > kMaxBuf) p.nbuf = kMaxBuf;
  if (const char* e = getenv("TN_3X3_NBUF")) p.nbuf = std::max(1, std::min(p.nbuf, atoi(e)));  // A/B: cap the look-ahead
  if (p.nbuf < 1) return cudaErrorInvalidValue;
  p.num_tiles = (p.NR + kTileRows - 1) / kTileRows;
  {
    const char* e = getenv("TN_3X3_L2_AHEAD");  // read per call: tools/ab_bench.py sweeps it in-process
    p.l2_ahead = e ? atoi(e) : 0;  // off: measured +0.4 ms per 2048-frame step at 3 tiles ahead (profiles/r2_ab_3x3_l2_prefetch.log)
    if (p.l2_ahead < 0) p.l2_ahead = 0;
  }
  p.wpack = cv.wpack;
  p.out = out;
  p.out_cstride = out_cstride;
  p.out_coff = out_coff;
  const int smem = fixed + p.nbuf * 2 * p.a_half_bytes;

  CUtensorMap tmap;
  cuuint64_t gdim[2] = {128, static_cast<cuuint64_t>(p.NR)};
  cuuint64_t gstride[1] = {128 * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(in_padded), gdim, gstride, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;

  static int configured_smem = 0;
  if (smem > configured_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured_smem = smem;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  ProfScope prof_scope(kProfConvGemm, st);
  return launch_pdl(conv3x3_halo_kernel, dim3(grid), dim3(kThreads), smem, st, tmap, p);
}

}  // namespace tn
