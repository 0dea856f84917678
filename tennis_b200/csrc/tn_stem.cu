// Stem convolution 7x7 / stride 2 / pad 3 (3 -> 64 channels) + folded BN + ReLU, as a persistent TMA + tcgen05 kernel on
// the 2x2 space-to-depth image (see s2d_convert_kernel in tn_elementwise.cu):
//     out(oy,ox) = sum_{a,b<4} Wz[a][b] . Z[oy+a][ox+b][0:16]          (Z zero-padded, 16 channels per pixel, 32 B)
// GEMM rows enumerate the padded (Hz, Wz) grid, q = (f*Hz + y)*Wz + x, so tap (a,b) of row q reads pixel q + a*Wz + b:
//   * per tile of 128 rows the TMA engine loads FOUR raw pixel strips (one per filter row a): 136 px x 32 B, SWIZZLE_32B;
//     each activation byte crosses L2->SMEM ~4x (the Toeplitz/im2col formulation moved it 16x and was L2-read bound);
//   * a K=16 UMMA consumes exactly one pixel (16 ch = 32 B = one SWIZZLE_32B row), so the 4 horizontal taps are the SAME
//     strip addressed through descriptors whose start address is advanced by b rows (verified: tools/umma_sw32_probe.cu);
//     16 UMMAs (M128 x N64 x K16) per tile, weights (16 taps x 64 x 16) resident in shared memory;
//   * epilogue (8 warps): tcgen05.ld -> +shift -> ReLU -> bf16 -> swizzled per-warp staging -> coalesced 16-byte stores of the
//     valid rows (y < Ho, x < Wo) into the compact (n, Ho, Wo, 64) NHWC output.
#include <cuda.h>
#include <string.h>

#include "tn_common.h"
#include "tn_ptx.cuh"
#include "tn_stem.h"

namespace tn {

namespace {

constexpr int kThreads = 320;          // TMA warp, MMA warp, 8 epilogue warps
constexpr int kN = 64;
constexpr int kStripRows = 136;        // 128 + 3 halo pixels, rounded up to the 8-row swizzle atom
constexpr int kStripBytes = 5120;      // 136*32 = 4352 rounded up to 1024
constexpr int kABuf = 4 * kStripBytes; // four filter rows
constexpr int kTapBytes = kN * 32;     // one tap's weights: 64 rows x 16 bf16
constexpr int kWBytes = 16 * kTapBytes;
constexpr int kStageBytes = 8 * 32 * 64;  // per-warp epilogue staging: 32 rows x 64 B
constexpr int kSmem = 1024 + kWBytes + 2 * kABuf + kStageBytes + 256 /*shift*/ + 256 /*barriers*/;

struct StemParams {
  int Hz, Wz, Ho, Wo;
  int M;          // n * Hz * Wz
  int num_tiles;
  const uint8_t* wpack;
  const float* shift;   // [64]
  __nv_bfloat16* out;   // (n, Ho, Wo, 64)
};

__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;  // SBO: 8 rows x 32 B
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;         // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) stem_s2d_kernel(const __grid_constant__ CUtensorMap tmap, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = sW + kWBytes;
  uint8_t* sStage = sA + 2 * kABuf;
  float* sShift = reinterpret_cast<float*>(sStage + kStageBytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sShift + 64);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (tid < 64) sShift[tid] = p.shift[tid];
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // weights are constants: fetch them before waiting on the producer grid
    mbar_arrive_expect_tx(w_full, kWBytes);
    bulk_g2s(sW, p.wpack, kWBytes, w_full);
  }
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[buf], 4 * kStripRows * 32);
        for (int a = 0; a < 4; ++a) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
              ::"r"(smem_u32(sA + buf * kABuf + a * kStripBytes)), "l"(&tmap), "r"(0), "r"(t * 128 + a * p.Wz),
                "r"(smem_u32(&a_full[buf]))
              : "memory");
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(kN);
      mbar_wait(w_full, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1, ab = it & 1;
        mbar_wait(&a_full[buf], (it >> 1) & 1);
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kN;
        const uint32_t a_base = smem_u32(sA + buf * kABuf);
        const uint32_t w_base = smem_u32(sW);
        uint32_t acc = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            umma_bf16_ss(d_tmem, desc_sw32(a_base + a * kStripBytes + b * 32), desc_sw32(w_base + (a * 4 + b) * kTapBytes), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[ab]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int qw = warp & 3;
    const int hf = ew >> 2;  // channels [32*hf, 32*hf+32)
    const uint32_t stg = smem_u32(sStage + ew * 32 * 64);
    const int ghw = p.Hz * p.Wz;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int ab = it & 1;
      const int q = t * 128 + qw * 32 + lane;
      int orow = -1;
      if (q < p.M) {
        const int f = q / ghw;
        const int rem = q - f * ghw;
        const int y = rem / p.Wz;
        const int x = rem - y * p.Wz;
        if (y < p.Ho && x < p.Wo) orow = (f * p.Ho + y) * p.Wo + x;
      }
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + ab * kN + hf * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      // +shift, ReLU, bf16 -> staging row `lane` (64 B, 16-byte chunks XOR-swizzled by (row>>1)&3: conflict-free both ways)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float f8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f8[j] = fmaxf(__uint_as_float(v[8 * c + j]) + sShift[hf * 32 + 8 * c + j], 0.f);
        const uint4 pk = make_uint4(pack_bf16x2(f8[0], f8[1]), pack_bf16x2(f8[2], f8[3]), pack_bf16x2(f8[4], f8[5]),
                                    pack_bf16x2(f8[6], f8[7]));
        const uint32_t addr = stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
      }
      __syncwarp();
      // coalesced: 4 lanes x 16 B per row, 8 rows per instruction
      const int c = lane & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = (lane >> 2) + 8 * j;
        const int orr = __shfl_sync(0xffffffffu, orow, rr);
        if (orr >= 0) {
          uint4 val;
          const uint32_t addr = stg + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w) : "r"(addr) : "memory");
          *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(orr) * 64 + hf * 32 + c * 8) = val;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// Stem conv + BN + ReLU + MaxPool(3, stride 2, pad 1) in one kernel: the (n,112,112,64) stem activation (3.3 GB for 2048
// frames, written once and read once by a separate pool kernel) never exists.  One tile = one POOLED row (f, py): the
// three stem rows 2py-1, 2py, 2py+1 are three accumulators (3 x 16 UMMAs over six pixel strips); the epilogue takes the
// vertical max across accumulators (same TMEM lane), the horizontal max with warp shuffles (+ one value per lane quarter
// through shared memory), adds the folded BN shift, ReLU (both commute with max), and even columns write the pooled
// pixel into channels [0,64) of the first dense block's concat buffer.
struct StemPoolParams {
  int Hz, Wz, Hs, Ws, Hp, Wp;   // padded s2d grid, stem output size, pooled output size
  int num_tiles;                // n * Hp
  const uint8_t* wpack;
  const float* shift;
  __nv_bfloat16* out;           // (n, Hp, Wp, out_cstride), channels [0,64)
  int out_cstride;
};
constexpr int kPoolABuf = 6 * kStripBytes;
// Weights of the fused-pool kernel, stacked along N: pixel strip s (0..5 = s2d rows 2py-1+s) feeds stem row j (0..2) with filter
// row a = s - j, so for a fixed (s, b) ONE UMMA with B = [W[s-j][b] for every valid j] (N = 64, 128 or 192) updates all the
// accumulators the strip contributes to.  24 UMMAs per tile instead of 48 of N = 64: the A operand (4 KB per UMMA) is read from
// shared memory half as often -- 192 KB instead of 288 KB of operand reads per tile, at the same tensor time.
//   s:        0    1     2      3      4    5
//   j range:  0   0-1   0-2    0-2    1-2   2
__host__ __device__ constexpr int pool_jmin(int s) { return s <= 3 ? 0 : s - 3; }
__host__ __device__ constexpr int pool_jmax(int s) { return s <= 2 ? s : 2; }
__host__ __device__ constexpr int pool_blob_off(int s) {  // byte offset of strip s's four (b) blobs
  int rows = 0;
  for (int i = 0; i < s; ++i) rows += 4 * 64 * (pool_jmax(i) - pool_jmin(i) + 1);
  return rows * 32;
}
constexpr int kPoolWBytes = pool_blob_off(6);  // 96 KB
constexpr int kPoolSmem = 1024 + kPoolWBytes + 2 * kPoolABuf + 4096 /*exchange*/ + 256 + 256;

__global__ void __launch_bounds__(kThreads, 1) stem_pool_kernel(const __grid_constant__ CUtensorMap tmap, const StemPoolParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = sW + kPoolWBytes;
  float* xch = reinterpret_cast<float*>(sA + 2 * kPoolABuf);  // [2 parity][2 halves][4 quarters][32]
  float* sShift = xch + 1024;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sShift + 64);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 8);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (tid < 64) sShift[tid] = p.shift[tid];
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // weights are constants: fetch them before waiting on the producer grid
    mbar_arrive_expect_tx(w_full, kPoolWBytes);
    for (int off = 0; off < kPoolWBytes; off += 32768) bulk_g2s(sW + off, p.wpack + off, 32768, w_full);
  }
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const int f = t / p.Hp, py = t - f * p.Hp;
        mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[buf], 6 * kStripRows * 32);
        for (int s = 0; s < 6; ++s) {  // Zp rows 2py-1 .. 2py+4 of frame f, from column 0
          const int row = (f * p.Hz + 2 * py - 1 + s) * p.Wz;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
              ::"r"(smem_u32(sA + buf * kPoolABuf + s * kStripBytes)), "l"(&tmap), "r"(0), "r"(row), "r"(smem_u32(&a_full[buf]))
              : "memory");
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(w_full, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1, ab = it & 1;
        mbar_wait(&a_full[buf], (it >> 1) & 1);
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + buf * kPoolABuf);
        const uint32_t w_base = smem_u32(sW);
        const uint32_t d_base = tmem_base + ab * 192;
        // strip 2 first: its N = 192 UMMA touches all three accumulators, so it can initialise them (accumulate = 0)
        constexpr int kOrder[6] = {2, 3, 1, 4, 0, 5};
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int s_ = kOrder[i];
          const int nj = pool_jmax(s_) - pool_jmin(s_) + 1;
          const uint32_t idesc = umma_idesc_bf16_m128(64 * nj);
          const uint32_t d_tmem = d_base + pool_jmin(s_) * kN;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            umma_bf16_ss(d_tmem, desc_sw32(a_base + s_ * kStripBytes + b * 32), desc_sw32(w_base + pool_blob_off(s_) + b * nj * kTapBytes),
                         idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&acc_full[ab]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int qw = warp & 3;
    const int hf = ew >> 2;
    const int x = qw * 32 + lane;  // stem column of this thread
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int ab = it & 1;
      const int f = t / p.Hp, py = t - f * p.Hp;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      float m[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) m[c] = -INFINITY;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + ab * 192 + j * kN + hf * 32, v);
        tmem_ld_wait();
        const int y = 2 * py - 1 + j;
        if (y >= 0 && y < p.Hs) {
#pragma unroll
          for (int c = 0; c < 32; ++c) m[c] = fmaxf(m[c], __uint_as_float(v[c]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      if (x >= p.Ws) {
#pragma unroll
        for (int c = 0; c < 32; ++c) m[c] = -INFINITY;  // columns past the image edge never win the max
      }
      // column 32*qw-1 (lane 31 of the previous lane quarter) through shared memory: 8 x STS.128 by one lane, 8 x LDS.128 by
      // lane 0 (scalar, lane-predicated accesses inside the channel loop cost 64 wavefronts per warp and tile)
      float* xq = xch + (it & 1) * 256 + hf * 128;
      if (lane == 31) {
        float4* dstq = reinterpret_cast<float4*>(xq + qw * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) dstq[c4] = make_float4(m[4 * c4], m[4 * c4 + 1], m[4 * c4 + 2], m[4 * c4 + 3]);
      }
      if (hf == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      float lq[32];
      if (lane == 0 && qw > 0) {
        const float4* srcq = reinterpret_cast<const float4*>(xq + (qw - 1) * 32);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 w4 = srcq[c4];
          lq[4 * c4] = w4.x; lq[4 * c4 + 1] = w4.y; lq[4 * c4 + 2] = w4.z; lq[4 * c4 + 3] = w4.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) lq[c] = -INFINITY;
      }
      float o[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float left = __shfl_up_sync(0xffffffffu, m[c], 1);
        const float right = __shfl_down_sync(0xffffffffu, m[c], 1);
        if (lane == 0) left = lq[c];
        o[c] = fmaxf(fmaxf(left, m[c]), right);  // lane 31 is an odd column: never a pooling centre
      }
      if ((x & 1) == 0 && x < p.Ws && (x >> 1) < p.Wp) {
        uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(f * p.Hp + py) * p.Wp + (x >> 1)) * p.out_cstride + hf * 32);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          float f8[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) f8[jj] = fmaxf(o[8 * c4 + jj] + sShift[hf * 32 + 8 * c4 + jj], 0.f);
          dst[c4] = make_uint4(pack_bf16x2(f8[0], f8[1]), pack_bf16x2(f8[2], f8[3]), pack_bf16x2(f8[4], f8[5]), pack_bf16x2(f8[6], f8[7]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
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

}  // namespace

bool make_stem(DeviceArena& arena, const float* w /* (64,3,7,7) */, const float* fold_scale, StemDev* out) {
  std::vector<uint8_t> blob(kWBytes, 0);
  for (int n = 0; n < 64; ++n)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) {
            const int r = 2 * a + py - 1, s = 2 * b + px - 1;  // W8[u][v] = W[u-1][v-1]
            if (r < 0 || s < 0 || r > 6 || s > 6) continue;
            for (int c = 0; c < 3; ++c) {
              const int k = (py * 2 + px) * 3 + c;  // channel of the space-to-depth pixel
              float v = w[((static_cast<size_t>(n) * 3 + c) * 7 + r) * 7 + s];
              if (fold_scale) v *= fold_scale[n];
              const __nv_bfloat16 bv = __float2bfloat16(v);
              const size_t off = static_cast<size_t>(a * 4 + b) * kTapBytes + n * 32 + ((((k >> 3) ^ ((n >> 2) & 1))) << 4) + (k & 7) * 2;
              memcpy(&blob[off], &bv, 2);
            }
          }
  out->wpack = static_cast<const uint8_t*>(arena.upload(blob.data(), blob.size()));
  // stacked image for the fused-pool kernel: blob (s, b) = rows [(j - jmin) * 64 + n] = tap (a = s - j, b) of output channel n
  std::vector<uint8_t> pool(kPoolWBytes, 0);
  for (int s_ = 0; s_ < 6; ++s_) {
    const int jmin = pool_jmin(s_), nj = pool_jmax(s_) - jmin + 1;
    for (int b = 0; b < 4; ++b)
      for (int j = jmin; j < jmin + nj; ++j) {
        const int a = s_ - j;
        memcpy(&pool[static_cast<size_t>(pool_blob_off(s_)) + static_cast<size_t>(b) * nj * kTapBytes + static_cast<size_t>(j - jmin) * kTapBytes],
               &blob[static_cast<size_t>(a * 4 + b) * kTapBytes], kTapBytes);
      }
  }
  out->wpack_pool = static_cast<const uint8_t*>(arena.upload(pool.data(), pool.size()));
  return out->wpack != nullptr && out->wpack_pool != nullptr;
}

cudaError_t launch_stem_s2d(const StemDev& sd, const __nv_bfloat16* z, int n, int Hz, int Wz, int Ho, int Wo, const float* shift,
                            __nv_bfloat16* out, int num_sms, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  const long long total = static_cast<long long>(n) * Hz * Wz;
  if (total <= 0) return cudaSuccess;
  if (total >= (1ll << 31) - 4096) return cudaErrorInvalidValue;
  StemParams p;
  p.Hz = Hz;
  p.Wz = Wz;
  p.Ho = Ho;
  p.Wo = Wo;
  p.M = static_cast<int>(total);
  p.num_tiles = (p.M + 127) / 128;
  p.wpack = sd.wpack;
  p.shift = shift;
  p.out = out;
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {16, static_cast<cuuint64_t>(total)};
  cuuint64_t gstride[1] = {32};
  cuuint32_t box[2] = {16, kStripRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(z), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  ProfScope prof_scope(kProfConvGemm, st);
  return launch_pdl(stem_s2d_kernel, dim3(grid), dim3(kThreads), kSmem, st, tmap, p);
}

cudaError_t launch_stem_pool(const StemDev& sd, const __nv_bfloat16* z, int n, int Hz, int Wz, int Hs, int Ws, int Hp, int Wp,
                             const float* shift, __nv_bfloat16* out, int out_cstride, int num_sms, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  const long long total = static_cast<long long>(n) * Hz * Wz;
  if (total <= 0) return cudaSuccess;
  if (total >= (1ll << 31) - 4096 || Ws > 128) return cudaErrorInvalidValue;
  StemPoolParams p;
  p.Hz = Hz; p.Wz = Wz; p.Hs = Hs; p.Ws = Ws; p.Hp = Hp; p.Wp = Wp;
  p.num_tiles = n * Hp;
  p.wpack = sd.wpack_pool;
  p.shift = shift;
  p.out = out;
  p.out_cstride = out_cstride;
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {16, static_cast<cuuint64_t>(total)};
  cuuint64_t gstride[1] = {32};
  cuuint32_t box[2] = {16, kStripRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(z), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoolSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  ProfScope prof_scope(kProfConvGemm, st);
  return launch_pdl(stem_pool_kernel, dim3(grid), dim3(kThreads), kPoolSmem, st, tmap, p);
}

}  // namespace tn
