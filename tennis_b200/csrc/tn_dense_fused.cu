// Fused DenseNet layer:  BN1+ReLU -> conv1x1 (C_in -> 128) -> BN2+ReLU -> conv3x3 (128 -> 32) -> in-place concat,
// ONE persistent kernel per layer; the 128-channel bottleneck never leaves the SM.
//
// Unfused, every layer writes the bottleneck (256 B/pixel) to HBM and reads it back (~1.9x with halos); in dense blocks
// 1-2 those two transfers are ~60 % of the layer's HBM traffic and the 1x1 GEMMs run at 11-15 % tensor utilisation because
// they wait on memory.  Here a CTA owns TR whole image rows (TR*(W+2) <= 128 padded positions):
//   1. TMA (4-D tensor map c,x,y,f; box 64 ch x (W+2) x (TR+2) rows, out-of-image pixels zero-filled) streams the raw
//      concat-buffer rows of the tile PLUS its one-row halo, K-chunk by K-chunk, with the matching 1x1 weight chunk;
//   2. 8 transformer warps apply BN1+ReLU in place (fp32 math);
//   3. the MMA warp accumulates the 1x1 conv for all (TR+2)*(W+2) <= 256 halo rows in TMEM (two M=128 tiles, N=128);
//      the halo rows are recomputed by the neighbouring tile -- free, the tensor pipe was idle;
//   4. epilogue warps read TMEM, add the folded BN2 shift, ReLU, round to bf16 and write the bottleneck tile straight into
//      shared memory in the UMMA 128B-swizzled layout (image-border positions are written as zeros = the 3x3's padding);
//   5. the MMA warp runs the 3x3 conv from that tile exactly like tn_conv3x3.cu (row-shifted descriptors for dy, N = 96 =
//      3 dx taps stacked), and the epilogue combines the dx taps and stores the 32 new channels at their channel offset.
// Wide maps (W = 56) are cut into 14-column patches: a 16 x 10 halo tile yields 14 x 8 outputs (1.43x recompute instead of
// 2.07x for two full rows).  Shared memory (W=56): A/B1 ring 3 x 36 KB, bottleneck tile 40 KB, 3x3 weights 72 KB = 223 KB.
// TMEM: 2x128 + 96 columns.
#include <cuda.h>
#include <string.h>

#include "tn_common.h"
#include "tn_dense_fused.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

constexpr int kTransWarps = 8;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (1 + kTransWarps + 1 + kEpiWarps) * 32;  // 576: TMA, 8 transformers, MMA, 8 epilogue
constexpr int kMmaWarp = 1 + kTransWarps;
constexpr int kEpiWarp0 = kMmaWarp + 1;
constexpr int kN1 = 128, kN2 = 96;
constexpr int kB1Bytes = kN1 * 128;     // one 64-wide K-chunk of the 1x1 weights
constexpr int kW2Blob = kN2 * 128;
constexpr int kW2Bytes = 6 * kW2Blob;   // 72 KB
constexpr int kMaxNS = 4;               // ring stages: as many as fit (FusedParams::ns)
constexpr int kTmemCols = 512;          // acc1: 2 x 128, acc2: 96

struct FusedParams {
  int ns;                        // ring stages in use (2..kMaxNS)
  int F, H, W, Wp, TR, NY;       // tile = TR image rows x PW image columns; Wp = PW + 2 (row pitch of the halo tile); NY = TR + 2
  int PW, tiles_x;               // PW == W (row bands) or a divisor-sized column patch (W = 56: 14 -> 16 x 10 halo tile)
  int halo_rows;                 // NY * Wp  (<= 256)
  int a_bytes;                   // halo_rows * 128 rounded up to 1024
  int tiles_per_frame, num_tiles;
  int Cin, nchunks;
  const float* bn1_scale;        // [Cin]
  const float* bn1_shift;
  const float* bn2_shift;        // [128] (scale folded into w1)
  const uint8_t* w1pack;         // nchunks x [128 rows x 128 B] swizzled
  const uint8_t* w2pack;         // 6 blobs [dy][half], 96 rows x 128 B swizzled
  __nv_bfloat16* out;            // concat buffer (F,H,W,out_cstride), 32 channels written at out_coff
  int out_cstride, out_coff;
};

__device__ __forceinline__ uint32_t cvt_pack(float a, float b, bool relu) {
  uint32_t r;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) dense_layer_fused_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = p.a_bytes + kB1Bytes;
  uint8_t* sRing = smem;                                   // ns x (A | B1)
  uint8_t* sHalo = sRing + p.ns * stage_bytes;              // 2 halves x a_bytes (bottleneck tile, bf16, swizzled)
  uint8_t* sW2 = sHalo + 2 * p.a_bytes;                    // 72 KB
  uint8_t* tail = sW2 + kW2Bytes;
  float* sShift2 = reinterpret_cast<float*>(tail);         // [128]
  float* xch = sShift2 + 128;                              // [2 halves][4 quarters][2][16]
  uint64_t* tma_full = reinterpret_cast<uint64_t*>(xch + 256);  // [kMaxNS]
  uint64_t* a_ready = tma_full + kMaxNS;                   // [kMaxNS]  transformers done
  uint64_t* empty_bar = a_ready + kMaxNS;                  // [kMaxNS]  MMA1 done with the stage
  uint64_t* acc1_full = empty_bar + kMaxNS;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* halo_full = acc1_empty + 1;
  uint64_t* halo_empty = halo_full + 1;
  uint64_t* acc2_full = halo_empty + 1;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* w2_full = acc2_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w2_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kMaxNS; ++s) {
      mbar_init(&tma_full[s], 1);
      mbar_init(&a_ready[s], kTransWarps);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, kEpiWarps);
    mbar_init(halo_full, kEpiWarps);
    mbar_init(halo_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kEpiWarps);
    mbar_init(w2_full, 1);
    mbar_fence_init();
  }
  if (tid < 128) sShift2[tid] = p.bn2_shift[tid];
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc1 = tmem_base;          // columns [0,256): two M tiles
  const uint32_t tm_acc2 = tmem_base + 256;    // columns [256,352)

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(w2_full, kW2Bytes);
      for (int b = 0; b < 6; ++b) bulk_g2s(sW2 + b * kW2Blob, p.w2pack + b * kW2Blob, kW2Blob, w2_full);
      int stage = 0;
      uint32_t phase = 1;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int f = t / p.tiles_per_frame;
        const int tt = t - f * p.tiles_per_frame;
        const int yt = tt / p.tiles_x, xt = tt - yt * p.tiles_x;
        const int y0 = yt * p.TR - 1;  // unpadded image row / column of the first halo row / column
        const int x0 = xt * p.PW - 1;
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&empty_bar[stage], phase);
          uint8_t* st = sRing + stage * stage_bytes;
          mbar_arrive_expect_tx(&tma_full[stage], static_cast<uint32_t>(p.halo_rows * 128 + kB1Bytes));
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(st)), "l"(&tmap), "r"(c * 64), "r"(x0), "r"(y0), "r"(f), "r"(smem_u32(&tma_full[stage]))
              : "memory");
          bulk_g2s(st + p.a_bytes, p.w1pack + static_cast<size_t>(c) * kB1Bytes, kB1Bytes, &tma_full[stage]);
          if (++stage == p.ns) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 1 && warp <= kTransWarps) {
    // ================================================================ transformers: BN1 + ReLU in place
    const int ttid = tid - 32;  // 0..255
    const int j = ttid & 7;     // 16-byte slot within the 128-byte row
    const int r0 = ttid >> 3;   // first row (0..31), then +32
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      for (int c = 0; c < p.nchunks; ++c) {
        mbar_wait(&tma_full[stage], phase);
        const uint32_t a_stage = smem_u32(sRing + stage * stage_bytes);
        {
          // rows r0 + 32*i share (r & 7), so this thread always handles channel group g of the chunk: load BN1 once
          const int g = j ^ (r0 & 7);
          const int ch = c * 64 + g * 8;
          if (ch < p.Cin) {
            const float4* s4 = reinterpret_cast<const float4*>(p.bn1_scale + ch);
            const float4* h4 = reinterpret_cast<const float4*>(p.bn1_shift + ch);
            const float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1), h0 = __ldg(h4), h1 = __ldg(h4 + 1);
#pragma unroll 4
            for (int r = r0; r < p.halo_rows; r += 32) {
              const uint32_t addr = a_stage + r * 128 + (j << 4);
              const uint4 x = lds128(addr);
              uint4 o;
              o.x = cvt_pack(fmaf(bf_lo(x.x), s0.x, h0.x), fmaf(bf_hi(x.x), s0.y, h0.y), true);
              o.y = cvt_pack(fmaf(bf_lo(x.y), s0.z, h0.z), fmaf(bf_hi(x.y), s0.w, h0.w), true);
              o.z = cvt_pack(fmaf(bf_lo(x.z), s1.x, h1.x), fmaf(bf_hi(x.z), s1.y, h1.y), true);
              o.w = cvt_pack(fmaf(bf_lo(x.w), s1.z, h1.z), fmaf(bf_hi(x.w), s1.w, h1.w), true);
              sts128(addr, o);
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[stage]);
        if (++stage == p.ns) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_bf16_m128(kN1);
      const uint32_t idesc2 = umma_idesc_bf16_m128(kN2);
      mbar_wait(w2_full, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        // ---- 1x1 conv over the halo rows (two M tiles)
        mbar_wait(acc1_empty, (it & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.nchunks; ++c) {
          mbar_wait(&a_ready[stage], phase);
          tc_fence_after();
          const int kv = min(64, p.Cin - c * 64);
          const uint32_t a_addr = smem_u32(sRing + stage * stage_bytes);
          const uint64_t db = umma_desc_sw128(a_addr + p.a_bytes);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (mt * 128 < p.halo_rows) {
              const uint64_t da = umma_desc_sw128(a_addr + mt * 128 * 128);
              for (int k = 0; k < kv / 16; ++k) umma_bf16_ss(tm_acc1 + mt * kN1, da + 2 * k, db + 2 * k, idesc1, (c > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.ns) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(acc1_full);
        // ---- 3x3 conv from the bottleneck tile the epilogue warps put in shared memory
        mbar_wait(halo_full, it & 1);
        mbar_wait(acc2_empty, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t h_base = smem_u32(sHalo);
        const uint32_t w_base = smem_u32(sW2);
        uint32_t acc = 0;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t da = umma_desc_sw128(h_base + half * p.a_bytes + dy * p.Wp * 128);
            const uint64_t db = umma_desc_sw128(w_base + (dy * 2 + half) * kW2Blob);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(tm_acc2, da + 2 * k, db + 2 * k, idesc2, acc);
              acc = 1;
            }
          }
        }
        umma_commit(halo_empty);
        umma_commit(acc2_full);
      }
    }
  } else {
    // ================================================================ epilogue warps
    const int ew = warp - kEpiWarp0;  // 0..7
    const int qw = warp & 3;          // TMEM lane quarter
    // epilogue 1 role: M tile = ew>>2 ... but the quarter is fixed by warp%4, so pair (qw, mt) with mt = ew >> 2
    const int mt = ew >> 2;
    // epilogue 2 role: quarter qw, output-channel half hf = ew >> 2
    const int hf = ew >> 2;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int f = t / p.tiles_per_frame;
      const int tt = t - f * p.tiles_per_frame;
      const int yt = tt / p.tiles_x, xt = tt - yt * p.tiles_x;
      const int yp0 = yt * p.TR;  // padded row / column index of the first HALO row / column (padded index 0 = border)
      const int xp0 = xt * p.PW;
      // ------------------------------------------------ epilogue 1: TMEM -> (+shift2, ReLU, bf16) -> bottleneck tile in smem
      mbar_wait(acc1_full, it & 1);
      tc_fence_after();
      mbar_wait(halo_empty, (it & 1) ^ 1);  // previous tile's 3x3 MMAs no longer read the tile
      {
        const int r = mt * 128 + qw * 32 + lane;  // halo row of this thread
        const bool in_tile = r < p.halo_rows;
        const int yy = r / p.Wp, xx = r - yy * p.Wp;
        const int yp = yp0 + yy;
        const int xp = xp0 + xx;
        const bool interior = in_tile && xp >= 1 && xp <= p.W && yp >= 1 && yp <= p.H;
        const uint32_t row_addr = smem_u32(sHalo) + r * 128;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tm_acc1 + (static_cast<uint32_t>(qw * 32) << 16) + mt * kN1 + cb * 32, v);
          tmem_ld_wait();
          if (in_tile) {
            const uint32_t half_addr = row_addr + (cb >> 1) * p.a_bytes;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o = make_uint4(0, 0, 0, 0);
              if (interior) {
                const float4 sa = *reinterpret_cast<const float4*>(sShift2 + cb * 32 + q * 8);
                const float4 sb = *reinterpret_cast<const float4*>(sShift2 + cb * 32 + q * 8 + 4);
                o.x = cvt_pack(__uint_as_float(v[q * 8 + 0]) + sa.x, __uint_as_float(v[q * 8 + 1]) + sa.y, true);
                o.y = cvt_pack(__uint_as_float(v[q * 8 + 2]) + sa.z, __uint_as_float(v[q * 8 + 3]) + sa.w, true);
                o.z = cvt_pack(__uint_as_float(v[q * 8 + 4]) + sb.x, __uint_as_float(v[q * 8 + 5]) + sb.y, true);
                o.w = cvt_pack(__uint_as_float(v[q * 8 + 6]) + sb.z, __uint_as_float(v[q * 8 + 7]) + sb.w, true);
              }
              const int chunk = (cb & 1) * 4 + q;  // 16-byte chunk within the 64-channel half
              sts128(half_addr + ((chunk ^ (r & 7)) << 4), o);
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(acc1_empty);
          mbar_arrive(halo_full);
        }
      }
      // ------------------------------------------------ epilogue 2: 3x3 accumulators -> combine dx taps -> concat buffer
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
      {
        const int r = qw * 32 + lane;
        uint32_t v0[16], v1[16], v2[16];
        const uint32_t taddr = tm_acc2 + (static_cast<uint32_t>(qw * 32) << 16) + hf * 16;
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 32, v1);
        tmem_ld16(taddr + 64, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc2_empty);
        float* x = xch + (it & 1) * 0 + hf * 128;  // single-buffered: protected by the two named barriers below
        if (lane == 31) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) x[(qw * 2 + 0) * 16 + jj] = __uint_as_float(v0[jj]);
        }
        if (lane == 0) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) x[(qw * 2 + 1) * 16 + jj] = __uint_as_float(v2[jj]);
        }
        if (hf == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        float o[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[jj]), 1);
          float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[jj]), 1);
          if (lane == 0 && qw > 0) up = x[((qw - 1) * 2 + 0) * 16 + jj];
          if (lane == 31 && qw < 3) dn = x[((qw + 1) * 2 + 1) * 16 + jj];
          o[jj] = up + __uint_as_float(v1[jj]) + dn;
        }
        if (hf == 0) asm volatile("bar.sync 1, 128;" ::: "memory");  // exchange buffer may be rewritten by the next tile
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        const int yy = r / p.Wp, xx = r - yy * p.Wp;
        const int y = yp0 + yy;  // output image row (unpadded): padded row yp0+1+yy -> image row yp0+yy
        const int xo = xp0 + xx - 1;  // output image column
        if (yy < p.TR && xx >= 1 && xx <= p.PW && xo < p.W && y < p.H) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(f * p.H + y) * p.W + xo) * p.out_cstride + p.out_coff + hf * 16);
          dst[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          dst[1] = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

constexpr int kSmemLimit = 227 * 1024;

// Ring depth: as many stages as fit beside the bottleneck tile and the 3x3 weights (2..4).
int ring_stages(int a_bytes) {
  const int fixed = 1024 + 2 * a_bytes + kW2Bytes + 512 + 1024 + 256;
  int ns = (kSmemLimit - fixed) / (a_bytes + kB1Bytes);
  return ns > kMaxNS ? kMaxNS : ns;
}

bool geometry(int H, int W, int* TR, int* PW, int* halo_rows, int* a_bytes, int* smem) {
  // Wide maps are cut into column patches so that the one-pixel halo costs less: W = 56 -> 14 x 8 outputs from a 16 x 10
  // halo tile (1.43x recompute of the 1x1 conv instead of 2.07x for two full rows), and the smaller tile buys a deeper ring.
  *PW = W;
  if (W + 2 > 32 && W % 14 == 0) *PW = 14;
  const int Wp = *PW + 2;
  if (Wp > 128) return false;
  *TR = 128 / Wp;
  if (*TR > H) *TR = H;
  if (*TR < 1) return false;
  *halo_rows = (*TR + 2) * Wp;
  if (*halo_rows > 256) return false;
  *a_bytes = static_cast<int>(align_up(static_cast<size_t>(*halo_rows) * 128, 1024));
  const int ns = ring_stages(*a_bytes);
  if (ns < 2) return false;
  *smem = 1024 + ns * (*a_bytes + kB1Bytes) + 2 * *a_bytes + kW2Bytes + 512 + 1024 + 256;
  return *smem <= kSmemLimit;
}

}  // namespace

bool dense_fused_supported(int H, int W) {
  int TR, PW, hr, ab, sm;
  return geometry(H, W, &TR, &PW, &hr, &ab, &sm) && (PW + 2) <= 256 && (TR + 2) <= 256;
}

cudaError_t launch_dense_layer_fused(const __nv_bfloat16* blk, int blk_cstride, int F, int H, int W, int Cin, const float* bn1_scale,
                                     const float* bn1_shift, const uint8_t* w1pack, int w1_chunks, const float* bn2_shift,
                                     const uint8_t* w2pack, int num_sms, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  FusedParams p;
  int smem;
  if (!geometry(H, W, &p.TR, &p.PW, &p.halo_rows, &p.a_bytes, &smem)) return cudaErrorInvalidValue;
  p.F = F;
  p.H = H;
  p.W = W;
  p.Wp = p.PW + 2;
  p.NY = p.TR + 2;
  p.ns = ring_stages(p.a_bytes);
  p.tiles_x = (W + p.PW - 1) / p.PW;
  p.tiles_per_frame = ((H + p.TR - 1) / p.TR) * p.tiles_x;
  p.num_tiles = F * p.tiles_per_frame;
  p.Cin = Cin;
  p.nchunks = w1_chunks;
  p.bn1_scale = bn1_scale;
  p.bn1_shift = bn1_shift;
  p.bn2_shift = bn2_shift;
  p.w1pack = w1pack;
  p.w2pack = w2pack;
  p.out = const_cast<__nv_bfloat16*>(blk);
  p.out_cstride = blk_cstride;
  p.out_coff = Cin;
  CUtensorMap tmap;
  cuuint64_t gdim[4] = {static_cast<cuuint64_t>(blk_cstride), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(F)};
  cuuint64_t gstride[3] = {static_cast<cuuint64_t>(blk_cstride) * 2, static_cast<cuuint64_t>(W) * blk_cstride * 2,
                           static_cast<cuuint64_t>(H) * W * blk_cstride * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.Wp), static_cast<cuuint32_t>(p.NY), 1};  // halo tile: (PW+2) x (TR+2) pixels
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(blk), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(dense_layer_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  ProfScope prof_scope(kProfConvGemm, st);
  dense_layer_fused_kernel<<<grid, kThreads, smem, st>>>(tmap, p);
  return cudaGetLastError();
}

}  // namespace tn
