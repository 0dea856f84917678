// Fused DenseNet layer:  BN1+ReLU -> conv1x1 (C_in -> 128) -> BN2+ReLU -> conv3x3 (128 -> 32) -> in-place concat,
// ONE persistent kernel per layer; the 128-channel bottleneck never leaves the SM.
//
// Why (profiles/r2_k1_launches.md): after the tensor-memory 1x1 kernel the two-kernel schedule of dense blocks 1-2 is DRAM-bound
// (1x1 convs 5.0-5.9 TB/s of the 6.55 TB/s copy peak) and 36 of its 92 GB per 2048 frames are the bottleneck's round trip (written
// by the 1x1, read back ~1.4x by the 3x3).  Round 1's fused kernel removed those bytes but ran ONE tile per SM through six
// serial phases with a 256-row halo tile and was slower.  This version pipelines two tiles:
//   * tile = 14 x 6 output pixels, halo 16 x 8 = 128 rows = ONE M=128 UMMA tile for the 1x1 (no second, mostly empty M tile);
//   * the 1x1's A operand goes through TENSOR MEMORY as in tn_conv1x1_ts.cu: TMA (4-D box with zero-filled out-of-image pixels)
//     -> 8 row-owning transformer warps (LDS.128, packed fp32 FFMA2, BN1 parameters from a shared-memory table) -> tcgen05.st
//     into four TMEM slots -> tcgen05.mma with A from TMEM; the raw stage is released as soon as it has been read;
//   * the 1x1 accumulator is double-buffered in TMEM and the single MMA thread is a small SCHEDULER: it polls two in-order queues
//     (1x1 K-chunks whose operand slot is full; 3x3 halves whose bottleneck half-tile is written) with non-blocking mbarrier
//     tests and issues whatever is ready, so the 1x1 of tiles t+1, t+2 fills the tensor pipe while epilogue 1 turns tile t's
//     accumulator into the bf16 bottleneck tile in shared memory (+BN2 shift, ReLU, zeros at image-border positions = the 3x3's
//     padding) -- round 2's first version issued in a fixed order and measured 1.3x SLOWER than the two-kernel path because
//     the 3x3 of tile t waited behind the transformers of tile t+1 (profiles/r2_fused_v1_launches.md);
//   * the bottleneck tile is handed over per 64-channel HALF (own full/empty barriers): the 3x3 runs half 0 (12 UMMAs) then half 1,
//     so epilogue 1 of the next tile rewrites half 0 while the tensor core still reads half 1;
//   * separate warps for the two epilogues (8 for TMEM -> bottleneck tile, 4 for the 3x3 accumulator -> dx-tap combine ->
//     256-bit stores into the concat buffer), so neither waits for the other;
//   * 3x3 as in tn_conv3x3.cu: row-shifted descriptors for dy, the three dx taps stacked along N (96).
// 1x1 weights stay resident in shared memory for K <= 256 and are streamed with the activation chunks above that (block 2).
// Shared memory: 3x3 weights 72 KB + bottleneck tile 32 KB + [W1 <= 64 KB] + ring.
// TMEM (512 columns): 1x1 accumulators 2 x 128, 3x3 accumulator 96, four 32-column operand slots.  22 warps.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tn_common.h"
#include "tn_dense_fused.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

constexpr int kXfWarps = 8;     // warps 0-7   transformers (TMEM lane quarter = warp & 3, channel half = warp >> 2)
constexpr int kEpi1Warp0 = 8;   // warps 8-15  epilogue 1 (quarter = warp & 3, 64-channel half = (warp - 8) >> 2)
constexpr int kEpi1Warps = 8;
constexpr int kEpi2Warp0 = 16;  // warps 16-19 epilogue 2 (quarter = warp & 3)
constexpr int kEpi2Warps = 4;
constexpr int kMmaWarp = 20;
constexpr int kTmaWarp = 21;
constexpr int kThreads = 22 * 32;
constexpr int kN1 = 128, kN2 = 96;
constexpr int kABytes = 128 * 128;      // raw A stage: 128 halo rows x 64 bf16
constexpr int kB1Bytes = kN1 * 128;     // one 64-wide K-chunk of the 1x1 weights
constexpr int kW2Blob = kN2 * 128;
constexpr int kW2Bytes = 6 * kW2Blob;   // 72 KB
constexpr int kHaloHalf = 128 * 128;    // bottleneck tile: 128 rows x 64 ch per half (3x3 reads past row 127 only feed unused rows)
constexpr int kMaxNS = 6;
constexpr int kMaxResident = 4;         // K <= 256 resident
constexpr int kMaxChunks = 8;           // K <= 512
constexpr int kTmemCols = 512;
constexpr int kAcc1Col = 0, kAcc2Col = 256, kASlotCol = 352;  // acc1: 2 x 128, acc2: 96, A slots: 4 x 32
constexpr int kASlots = 4;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kTailBytes = kMaxChunks * 512 /*BN1 table*/ + 512 /*BN2 shift*/ + 2048 /*dx exchange*/ + 512 /*barriers*/;

struct FusedParams {
  int ns, resident;
  int F, H, W, Wp, TR, NY, PW;   // tile = TR image rows x PW image columns; halo tile Wp x NY = (PW+2) x (TR+2) <= 128 rows
  int tiles_x, tiles_y, tiles_per_frame, num_tiles;
  int halo_rows;
  int Cin, nchunks;
  const float* bn1_scale;        // [Cin]
  const float* bn1_shift;
  const float* bn2_shift;        // [128] (scale folded into w1)
  const uint8_t* w1pack;         // nchunks x [128 rows x 128 B] swizzled
  const uint8_t* w2pack;         // 6 blobs [dy][half], 96 rows x 128 B swizzled
  __nv_bfloat16* out;            // concat buffer (F,H,W,out_cstride), 32 channels written at out_coff
  int out_cstride, out_coff;
};

__device__ __forceinline__ uint32_t cvt_pack_relu(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// relu(x * s + b) on a bf16 pair, fp32 math (one packed FFMA2), single rounding back to bf16
__device__ __forceinline__ uint32_t bn_relu2(uint32_t x, const float4 sb) {
  float y0 = __uint_as_float(x << 16), y1 = __uint_as_float(x & 0xffff0000u);
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %5};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(y0), "+f"(y1)
      : "f"(sb.x), "f"(sb.y), "f"(sb.z), "f"(sb.w));
  return cvt_pack_relu(y0, y1);
}

__global__ void __launch_bounds__(kThreads, 1) dense_layer_fused_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nchunks = p.nchunks;
  const bool resident = p.resident != 0;
  const int stage_bytes = kABytes + (resident ? 0 : kB1Bytes);
  uint8_t* sW2 = smem;                                        // 72 KB
  uint8_t* sHalo = sW2 + kW2Bytes;                            // 2 halves x 16 KB (bottleneck tile, bf16, 128B-swizzled)
  uint8_t* sW1 = sHalo + 2 * kHaloHalf;                       // resident 1x1 weights (resident mode)
  uint8_t* sRing = sW1 + (resident ? nchunks * kB1Bytes : 0);  // ns x (A [| B1])
  uint8_t* tail = sRing + p.ns * stage_bytes;
  float4* sPar = reinterpret_cast<float4*>(tail);             // [nchunks][32 pairs] (s0, s1, b0, b1)
  float* sShift2 = reinterpret_cast<float*>(tail + kMaxChunks * 512);  // [128]
  float* xch = sShift2 + 128;                                 // [2 parity][2 halves][4 quarters][2][16]
  uint64_t* a_full = reinterpret_cast<uint64_t*>(xch + 512);  // [kMaxNS]
  uint64_t* a_empty = a_full + kMaxNS;                        // [kMaxNS]
  uint64_t* t_full = a_empty + kMaxNS;                        // [kASlots]
  uint64_t* t_empty = t_full + kASlots;                       // [kASlots]
  uint64_t* acc1_full = t_empty + kASlots;                    // [2]
  uint64_t* acc1_empty = acc1_full + 2;                       // [2]
  uint64_t* halo_full = acc1_empty + 2;                       // [2] per 64-channel half
  uint64_t* halo_empty = halo_full + 2;                       // [2]
  uint64_t* acc2_full = halo_empty + 2;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* w_full = acc2_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kMaxNS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], kXfWarps + (resident ? 0 : 1));
    }
    for (int i = 0; i < kASlots; ++i) {
      mbar_init(&t_full[i], kXfWarps);
      mbar_init(&t_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], kEpi1Warps);
      mbar_init(&halo_full[i], kEpi1Warps / 2);
      mbar_init(&halo_empty[i], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kEpi2Warps);
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (tid < 128) sShift2[tid] = p.bn2_shift[tid];
  for (int i = tid; i < nchunks * 32; i += kThreads) {  // BN1 table, one float4 per channel pair
    const int ch = 2 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < p.Cin) {
      v.x = p.bn1_scale[ch];
      v.z = p.bn1_shift[ch];
    }
    if (ch + 1 < p.Cin) {
      v.y = p.bn1_scale[ch + 1];
      v.w = p.bn1_shift[ch + 1];
    }
    sPar[i] = v;
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // weights are constants: fetch them before waiting on the producer grid
    mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(kW2Bytes + (resident ? nchunks * kB1Bytes : 0)));
    for (int b = 0; b < 6; ++b) bulk_g2s(sW2 + b * kW2Blob, p.w2pack + b * kW2Blob, kW2Blob, w_full);
    if (resident)
      for (int c = 0; c < nchunks; ++c) bulk_g2s(sW1 + c * kB1Bytes, p.w1pack + static_cast<size_t>(c) * kB1Bytes, kB1Bytes, w_full);
  }
  griddep_wait();

  if (warp == kTmaWarp) {
    // ================================================================ TMA producer: raw halo tiles (+ streamed 1x1 weights)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 1;
      const uint32_t tx = static_cast<uint32_t>(p.halo_rows * 128 + (resident ? 0 : kB1Bytes));
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int f = t / p.tiles_per_frame;
        const int tt = t - f * p.tiles_per_frame;
        const int yt = tt / p.tiles_x, xt = tt - yt * p.tiles_x;
        const int y0 = yt * p.TR - 1;  // image row / column of the first halo row / column (-1 = zero padding)
        const int x0 = xt * p.PW - 1;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&a_empty[stage], phase);
          uint8_t* st = sRing + stage * stage_bytes;
          mbar_arrive_expect_tx(&a_full[stage], tx);
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(smem_u32(st)), "l"(&tmap), "r"(c * 64), "r"(x0), "r"(y0), "r"(f), "r"(smem_u32(&a_full[stage]))
              : "memory");
          if (!resident) bulk_g2s(st + kABytes, p.w1pack + static_cast<size_t>(c) * kB1Bytes, kB1Bytes, &a_full[stage]);
          if (++stage == p.ns) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp < kXfWarps) {
    // ================================================================ transformers: smem (raw) -> BN1+ReLU -> TMEM
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int hh = warp >> 2;  // channels [32hh, 32hh+32) of every 64-channel chunk
    const int r = q * 32 + lane;  // halo row == TMEM lane
    const uint32_t row_off = static_cast<uint32_t>(r * 128);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kASlotCol + hh * 16;
    int stage = 0, slot = 0;
    uint32_t sphase = 0, tphase = 1;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&a_full[stage], sphase);
        const uint32_t a_row = smem_u32(sRing + stage * stage_bytes) + row_off;
        const float4* par = sPar + c * 32 + hh * 16;
        uint4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = lds128(a_row + (((4 * hh + i) ^ sw) << 4));
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[4 * i + 0] = bn_relu2(x[i].x, par[4 * i + 0]);
          o[4 * i + 1] = bn_relu2(x[i].y, par[4 * i + 1]);
          o[4 * i + 2] = bn_relu2(x[i].z, par[4 * i + 2]);
          o[4 * i + 3] = bn_relu2(x[i].w, par[4 * i + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[stage]);  // raw stage read: the TMA warp may refill it (streamed: once the MMAs retire too)
        mbar_wait(&t_empty[slot], tphase);            // UMMAs that read this TMEM slot have retired
        tc_fence_after();
        tmem_st16(t_lane + slot * 32, o);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_full[slot]);
        if (++stage == p.ns) {
          stage = 0;
          sphase ^= 1u;
        }
        if (++slot == kASlots) {
          slot = 0;
          tphase ^= 1u;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer = scheduler over two in-order queues
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_bf16_m128(kN1);
      const uint32_t idesc2 = umma_idesc_bf16_m128(kN2);
      mbar_wait(w_full, 0);
      const int n_my = (static_cast<int>(blockIdx.x) < p.num_tiles) ? (p.num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      // queue 1: 1x1 K-chunks (tile j1, chunk c1);  queue 2: 3x3 halves (tile j2, half h2)
      int j1 = 0, c1 = 0, slot = 0, bstage = 0, j2 = 0, h2 = 0;
      uint32_t tphase = 0, bphase = 0;
      const uint32_t h_base = smem_u32(sHalo);
      const uint32_t w_base = smem_u32(sW2);
      uint32_t idle = 0;
      while (j2 < n_my) {
        bool did = false;
        // ---- 3x3 half (priority: it releases the bottleneck half-tile and the epilogues)
        if (mbar_test(&halo_full[h2], j2 & 1) && (h2 == 1 || mbar_test(acc2_empty, (j2 & 1) ^ 1))) {
          tc_fence_after();
          const uint32_t d2 = tmem_base + kAcc2Col;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t da = umma_desc_sw128(h_base + h2 * kHaloHalf + dy * p.Wp * 128);
            const uint64_t db = umma_desc_sw128(w_base + (dy * 2 + h2) * kW2Blob);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(d2, da + 2 * k, db + 2 * k, idesc2, (h2 | dy | k) ? 1u : 0u);
          }
          umma_commit(&halo_empty[h2]);
          if (h2 == 1) {
            umma_commit(acc2_full);
            ++j2;
          }
          h2 ^= 1;
          did = true;
        } else if (j1 < n_my && (c1 > 0 || mbar_test(&acc1_empty[j1 & 1], ((j1 >> 1) & 1) ^ 1)) && mbar_test(&t_full[slot], tphase) &&
                   (resident || mbar_test(&a_full[bstage], bphase))) {
          // ---- one K-chunk of the 1x1 conv of tile j1
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + kAcc1Col + (j1 & 1) * kN1;
          const int kv = min(64, p.Cin - c1 * 64);
          const uint32_t a_tmem = tmem_base + kASlotCol + slot * 32;
          const uint32_t b_addr = resident ? smem_u32(sW1 + c1 * kB1Bytes) : smem_u32(sRing + bstage * stage_bytes + kABytes);
          const uint64_t db = umma_desc_sw128(b_addr);
          for (int k = 0; k < kv / 16; ++k) umma_bf16_ts(d_tmem, a_tmem + 8 * k, db + 2 * k, idesc1, (c1 > 0 || k > 0) ? 1u : 0u);
          umma_commit(&t_empty[slot]);
          if (!resident) {
            umma_commit(&a_empty[bstage]);
            if (++bstage == p.ns) {
              bstage = 0;
              bphase ^= 1u;
            }
          }
          if (++slot == kASlots) {
            slot = 0;
            tphase ^= 1u;
          }
          if (++c1 == nchunks) {
            umma_commit(&acc1_full[j1 & 1]);
            c1 = 0;
            ++j1;
          }
          did = true;
        }
        if (did) {
          idle = 0;
        } else if (__nanosleep(40), ++idle > (1u << 22)) {  // back off: every failed poll is a shared-memory access on the data pipe
          printf("tn: fused dense layer: MMA scheduler stalled (block %d, j1=%d c1=%d j2=%d h2=%d)\n", blockIdx.x, j1, c1, j2, h2);
          __trap();
        }
      }
    }
  } else if (warp >= kEpi1Warp0 && warp < kEpi1Warp0 + kEpi1Warps) {
    // ================================================================ epilogue 1: TMEM -> (+shift2, ReLU, bf16) -> bottleneck tile
    const int qw = warp & 3;
    const int half = (warp - kEpi1Warp0) >> 2;  // channels [64*half, 64*half + 64) == one swizzled half of the tile
    const int r = qw * 32 + lane;               // halo row of this thread
    const int yy = r / p.Wp, xx = r - yy * p.Wp;
    const uint32_t row_addr = smem_u32(sHalo) + half * kHaloHalf + r * 128;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int f = t / p.tiles_per_frame;
      const int tt = t - f * p.tiles_per_frame;
      const int yt = tt / p.tiles_x, xt = tt - yt * p.tiles_x;
      const int yp = yt * p.TR + yy;  // padded coordinates (0 = border)
      const int xp = xt * p.PW + xx;
      const bool interior = r < p.halo_rows && xp >= 1 && xp <= p.W && yp >= 1 && yp <= p.H;
      const int ab = it & 1;
      mbar_wait(&acc1_full[ab], (it >> 1) & 1);
      mbar_wait(&halo_empty[half], (it & 1) ^ 1);  // the previous tile's 3x3 MMAs no longer read this half of the bottleneck tile
      tc_fence_after();
#pragma unroll
      for (int cbl = 0; cbl < 2; ++cbl) {
        const int cb = half * 2 + cbl;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + kAcc1Col + ab * kN1 + cb * 32, v);
        tmem_ld_wait();
        if (cbl == 1) {  // accumulator drained by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc1_empty[ab]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (interior) {
            const float4 sa = *reinterpret_cast<const float4*>(sShift2 + cb * 32 + q * 8);
            const float4 sb = *reinterpret_cast<const float4*>(sShift2 + cb * 32 + q * 8 + 4);
            o.x = cvt_pack_relu(__uint_as_float(v[q * 8 + 0]) + sa.x, __uint_as_float(v[q * 8 + 1]) + sa.y);
            o.y = cvt_pack_relu(__uint_as_float(v[q * 8 + 2]) + sa.z, __uint_as_float(v[q * 8 + 3]) + sa.w);
            o.z = cvt_pack_relu(__uint_as_float(v[q * 8 + 4]) + sb.x, __uint_as_float(v[q * 8 + 5]) + sb.y);
            o.w = cvt_pack_relu(__uint_as_float(v[q * 8 + 6]) + sb.z, __uint_as_float(v[q * 8 + 7]) + sb.w);
          }
          const int chunk = cbl * 4 + q;  // 16-byte chunk within the 64-channel half
          sts128(row_addr + ((chunk ^ (r & 7)) << 4), o);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&halo_full[half]);
    }
  } else if (warp >= kEpi2Warp0 && warp < kEpi2Warp0 + kEpi2Warps) {
    // ================================================================ epilogue 2: 3x3 accumulator -> combine dx taps -> concat buffer
    const int qw = warp & 3;
    const int r = qw * 32 + lane;  // accumulator row: output pixel at halo position (yy + 1, xx)
    const int yy = r / p.Wp, xx = r - yy * p.Wp;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int f = t / p.tiles_per_frame;
      const int tt = t - f * p.tiles_per_frame;
      const int yt = tt / p.tiles_x, xt = tt - yt * p.tiles_x;
      const int y = yt * p.TR + yy;       // output image row
      const int xo = xt * p.PW + xx - 1;  // output image column
      const bool valid = yy < p.TR && xx >= 1 && xx <= p.PW && xo < p.W && y < p.H;
      __nv_bfloat16* dst = p.out + (static_cast<size_t>(f * p.H + y) * p.W + xo) * p.out_cstride + p.out_coff;
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v0[16], v1[16], v2[16];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + kAcc2Col + hf * 16;
        tmem_ld16(taddr, v0);
        tmem_ld16(taddr + 32, v1);
        tmem_ld16(taddr + 64, v2);
        tmem_ld_wait();
        if (hf == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);
        }
        // dx-tap exchange across the lane quarters with vector stores/loads by the two boundary lanes (see tn_conv3x3.cu)
        float* x = xch + (it & 1) * 256 + hf * 128;
        if (lane == 0 || lane == 31) {
          uint4* xd = reinterpret_cast<uint4*>(x + (qw * 2 + (lane == 0 ? 1 : 0)) * 16);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint4 w;
            w.x = (lane == 0) ? v2[4 * j4 + 0] : v0[4 * j4 + 0];
            w.y = (lane == 0) ? v2[4 * j4 + 1] : v0[4 * j4 + 1];
            w.z = (lane == 0) ? v2[4 * j4 + 2] : v0[4 * j4 + 2];
            w.w = (lane == 0) ? v2[4 * j4 + 3] : v0[4 * j4 + 3];
            xd[j4] = w;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float nb[16];
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
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j + u]), 1);    // D[r-1][dx=-1 block]
            float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j + u]), 1);  // D[r+1][dx=+1 block]
            if (need_up) up = nb[j + u];
            if (need_dn) dn = nb[j + u];
            e[u] = up + __uint_as_float(v1[j + u]) + dn;
          }
          o[j >> 1] = pack_bf16x2(e[0], e[1]);
        }
        if (valid) stg256(dst + hf * 16, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
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

// Tile geometry: PW output columns x TR output rows from a (PW+2) x (TR+2) halo tile of at most 128 rows (one M=128 UMMA).
bool geometry(int H, int W, int* TR, int* PW) {
  if (H < 1 || W < 1) return false;
  int pw = W;
  if (W > 14) pw = 14;            // 16-wide halo rows: 16 x 8 = 128 -> 14 x 6 outputs
  const int wp = pw + 2;
  int tr = 128 / wp - 2;
  if (tr > H) tr = H;
  if (tr < 1) return false;
  // the last valid accumulator row (TR-1)*Wp + PW must stay below 128 - 1 (its dx=+1 neighbour row is read)
  if ((tr - 1) * wp + pw + 1 > 127) return false;
  *TR = tr;
  *PW = pw;
  return true;
}

}  // namespace

bool dense_fused_supported(int H, int W) {
  int TR, PW;
  return geometry(H, W, &TR, &PW);
}

cudaError_t launch_dense_layer_fused(const __nv_bfloat16* blk, int blk_cstride, int F, int H, int W, int Cin, const float* bn1_scale,
                                     const float* bn1_shift, const uint8_t* w1pack, int w1_chunks, const float* bn2_shift,
                                     const uint8_t* w2pack, int num_sms, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  FusedParams p;
  memset(&p, 0, sizeof(p));
  if (!geometry(H, W, &p.TR, &p.PW)) return cudaErrorInvalidValue;
  if (w1_chunks < 1 || w1_chunks > kMaxChunks || (Cin % 16) != 0 || (blk_cstride % 16) != 0 || (Cin % 16) != 0 ||
      (reinterpret_cast<uintptr_t>(blk) & 31) != 0)
    return cudaErrorInvalidValue;
  p.F = F;
  p.H = H;
  p.W = W;
  p.Wp = p.PW + 2;
  p.NY = p.TR + 2;
  p.halo_rows = p.Wp * p.NY;
  p.tiles_x = (W + p.PW - 1) / p.PW;
  p.tiles_y = (H + p.TR - 1) / p.TR;
  p.tiles_per_frame = p.tiles_x * p.tiles_y;
  const long long nt = static_cast<long long>(F) * p.tiles_per_frame;
  if (nt <= 0) return cudaSuccess;
  if (nt >= (1ll << 31)) return cudaErrorInvalidValue;
  p.num_tiles = static_cast<int>(nt);
  p.Cin = Cin;
  p.nchunks = w1_chunks;
  static const int force_stream = getenv("TN_FUSED_STREAM_W1") ? atoi(getenv("TN_FUSED_STREAM_W1")) : 0;
  p.resident = (w1_chunks <= kMaxResident && !force_stream) ? 1 : 0;
  const int fixed = 1024 + kW2Bytes + 2 * kHaloHalf + kTailBytes + (p.resident ? w1_chunks * kB1Bytes : 0);
  const int stage_bytes = kABytes + (p.resident ? 0 : kB1Bytes);
  p.ns = (kSmemLimit - fixed) / stage_bytes;
  if (p.ns > kMaxNS) p.ns = kMaxNS;
  if (p.ns < 2) return cudaErrorInvalidValue;
  const int smem = fixed + p.ns * stage_bytes;
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
  return launch_pdl(dense_layer_fused_kernel, dim3(grid), dim3(kThreads), smem, st, tmap, p);
}

}  // namespace tn
