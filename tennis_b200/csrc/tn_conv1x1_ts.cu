// 1x1 / stride-1 convolution with a BN+ReLU pre-activation (every DenseNet bottleneck conv, K <= 512) whose A operand goes
// through TENSOR MEMORY: tcgen05.mma in the "TS" form (A from TMEM, B from shared memory).
//
// Why: the ncu captures of round 1 (profiles/r1_ncu_summary.md, re-read in profiles/r2_smem_pipe.md) show the TMA-fed 1x1
// kernel with the in-place shared-memory transform keeps the L1TEX data pipe ~80 % busy with three clients -- LSU (the
// lds -> bn/relu -> sts rewrite of every A tile plus a 2 x 32 KB epilogue staging round trip), the tensor core (A and B operand
// reads) and the TMA writes -- while the tensor pipe idles at 11-25 %.  Here
//   * the raw 128 x 64 bf16 tile still lands in shared memory by TMA (deep, register-free prefetch), but a transformer thread
//     owns one ROW: it reads its 64 B with conflict-free LDS.128 (the 128B swizzle spreads 8 consecutive rows over all banks),
//     applies relu(x*s + b) in fp32 (packed FFMA2, parameters broadcast from a shared-memory table), and writes the bf16 result
//     with tcgen05.st into a ring of TMEM slots: the activated operand never returns to shared memory and the tensor core
//     never reads A from it (-32 KB of data-pipe traffic per 128 x 64 chunk);
//   * the raw stage is released as soon as the transformers have read it (not when the MMAs retire);
//   * weights stay resident in shared memory for K <= 512 (8 chunks x 16 KB) -- possible because
//   * the epilogue stores straight from registers: a thread holds 32 consecutive channels of one output row, i.e. whole
//     32-byte sectors, written with 256-bit stores (STG.256); no staging buffer, no extra shared-memory round trip.
// Warp roles (18 warps): 0-7 transformers (lane quarter = warp % 4, channel half = warp / 4), 8 MMA issuer (+TMEM owner),
// 9-16 epilogue, 17 TMA producer.  TMEM: columns [0,256) two accumulators, [256,512) eight A slots of 32 columns.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "tn_common.h"
#include "tn_conv1x1_ts.h"
#include "tn_ptx.cuh"

namespace tn {

namespace {

constexpr int kXfWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kEpiWarp0 = 9;
constexpr int kEpiWarps = 8;
constexpr int kTmaWarp = 17;
constexpr int kThreads = 18 * 32;
constexpr int kBM = 128, kBN = 128;
constexpr int kABytes = kBM * 128;   // raw A stage: 128 rows x 64 bf16
constexpr int kBBytes = kBN * 128;   // one K-chunk of the weights
constexpr int kMaxChunks = 8;        // K <= 512 resident
constexpr int kMaxStages = 8;
constexpr int kASlots = 8;           // TMEM A slots (32 columns = 128 rows x 64 bf16 each)
constexpr int kAccCols = 2 * kBN;
constexpr int kTmemCols = 512;
constexpr int kOnesBytes = 1024;
constexpr int kParBytesPerChunk = 32 * 16;  // 32 channel pairs x (s0, s1, b0, b1)
constexpr int kSmemLimit = 227 * 1024;
constexpr int kFixedBytes = 1024 /*align*/ + kOnesBytes + kBBytes /*bias operand*/ + 512 /*barriers*/;

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// relu?(x * s + b) on a bf16 pair, fp32 math (one packed FFMA2), single rounding back to bf16
__device__ __forceinline__ uint32_t bn_act2(uint32_t x, const float4 sb, bool relu) {
  float y0 = __uint_as_float(x << 16), y1 = __uint_as_float(x & 0xffff0000u);
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tmov.b64 rc, {%4, %5};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(y0), "+f"(y1)
      : "f"(sb.x), "f"(sb.y), "f"(sb.z), "f"(sb.w));
  uint32_t r;
  if (relu)
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y1), "f"(y0));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y1), "f"(y0));
  return r;
}
__device__ __forceinline__ uint32_t cvt_pack(float a, float b, bool relu) {
  uint32_t r;
  if (relu)
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

__global__ void __launch_bounds__(kThreads, 1) conv1x1_ts_kernel(const ConvGemmParams p, const __grid_constant__ CUtensorMap tmap,
                                                                const int ns) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nchunks = p.num_chunks;
  uint8_t* sB = smem;                                // nchunks x 16 KB resident weights
  uint8_t* sA = sB + nchunks * kBBytes;              // ns x 16 KB raw activation stages
  uint8_t* sOnes = sA + ns * kABytes;                // A operand of the bias MMA
  uint8_t* sBias = sOnes + kOnesBytes;               // B operand of the bias MMA
  float4* sPar = reinterpret_cast<float4*>(sBias + kBBytes);  // [nchunks][32 pairs] (s0, s1, b0, b1)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sPar) + nchunks * kParBytesPerChunk);
  uint64_t* a_empty = a_full + kMaxStages;
  uint64_t* t_full = a_empty + kMaxStages;
  uint64_t* t_empty = t_full + kASlots;
  uint64_t* acc_full = t_empty + kASlots;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int num_m_tiles = (p.M + kBM - 1) / kBM;

  griddep_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], kXfWarps);
    }
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&t_full[s], kXfWarps);
      mbar_init(&t_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
  // folded BN shift of the OUTPUT side is added by the tensor core: one K=16 UMMA with A = [1 1 1 0...] and
  // B[n] = [hi mid lo 0...] of shift[n] (same trick as tn_conv_gemm.cu)
  const bool has_shift = p.epi_shift != nullptr;
  if (has_shift) {
    for (int i = tid; i < kOnesBytes / 16; i += kThreads) {
      const int r = i >> 3, slot = i & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (slot == (r & 7)) {
        v.x = 0x3F803F80u;
        v.y = 0x00003F80u;
      }
      *reinterpret_cast<uint4*>(sOnes + i * 16) = v;
    }
    for (int i = tid; i < kBN * 8; i += kThreads) {
      const int r = i >> 3, slot = i & 7;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (slot == (r & 7) && r < p.Cout) {
        const float sh = p.epi_shift[r];
        const __nv_bfloat16 hi = __float2bfloat16(sh);
        const float r1 = sh - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16(r1);
        const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
        v.x = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(mid)) << 16);
        v.y = static_cast<uint32_t>(__bfloat16_as_ushort(lo));
      }
      *reinterpret_cast<uint4*>(sBias + i * 16) = v;
    }
    fence_proxy_async_smem();
  }
  // pre-activation parameters, one float4 per channel pair; channels past Cin contribute nothing (never fed to the MMA)
  for (int i = tid; i < nchunks * 32; i += kThreads) {
    const int ch = 2 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < p.Cin) {
      v.x = p.pro_scale[ch];
      v.z = p.pro_shift[ch];
    }
    if (ch + 1 < p.Cin) {
      v.y = p.pro_scale[ch + 1];
      v.w = p.pro_shift[ch + 1];
    }
    sPar[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {  // weights are constants: fetch them before waiting on the producer grid
    mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(nchunks * kBBytes));
    for (int c = 0; c < nchunks; ++c) bulk_g2s(sB + c * kBBytes, p.wpack + static_cast<size_t>(c) * kBBytes, kBBytes, w_full);
  }
  griddep_wait();

  if (warp == kTmaWarp) {
    // ================================================================ TMA producer: raw activation tiles
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 1;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&a_empty[stage], phase);
          mbar_arrive_expect_tx(&a_full[stage], kABytes);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
              ::"r"(smem_u32(sA + stage * kABytes)), "l"(&tmap), "r"(c * 64), "r"(tile * kBM), "r"(smem_u32(&a_full[stage]))
              : "memory");
          if (++stage == ns) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp < kXfWarps) {
    // ================================================================ transformers: smem (raw) -> BN+ReLU -> TMEM
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int h = warp >> 2;  // channels [32h, 32h+32) of every 64-channel chunk
    const int r = q * 32 + lane;
    const uint32_t row_off = static_cast<uint32_t>(r * 128);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    const bool relu = p.pro_relu != 0;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kAccCols + h * 16;
    int stage = 0, slot = 0;
    uint32_t sphase = 0, tphase = 1;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&a_full[stage], sphase);
        const uint32_t a_row = smem_u32(sA + stage * kABytes) + row_off;
        uint4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = lds128(a_row + (((4 * h + i) ^ sw) << 4));
        const float4* par = sPar + c * 32 + h * 16;
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[4 * i + 0] = bn_act2(x[i].x, par[4 * i + 0], relu);
          o[4 * i + 1] = bn_act2(x[i].y, par[4 * i + 1], relu);
          o[4 * i + 2] = bn_act2(x[i].z, par[4 * i + 2], relu);
          o[4 * i + 3] = bn_act2(x[i].w, par[4 * i + 3], relu);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_empty[stage]);  // raw stage consumed: the TMA warp may refill it
        mbar_wait(&t_empty[slot], tphase);            // UMMAs that read this TMEM slot have retired
        tc_fence_after();
        tmem_st16(t_lane + slot * 32, o);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_full[slot]);
        if (++stage == ns) {
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
    // ================================================================ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(kBN);
      mbar_wait(w_full, 0);
      int slot = 0, tcount = 0;
      uint32_t tphase = 0;
      for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++tcount) {
        const int ab = tcount & 1;
        mbar_wait(&acc_empty[ab], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * kBN;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&t_full[slot], tphase);
          tc_fence_after();
          const int kv = min(64, p.Cin - c * 64);
          const uint32_t a_tmem = tmem_base + kAccCols + slot * 32;
          const uint64_t db = umma_desc_sw128(smem_u32(sB + c * kBBytes));
          for (int k = 0; k < kv / 16; ++k) umma_bf16_ts(d_tmem, a_tmem + 8 * k, db + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
          umma_commit(&t_empty[slot]);
          if (++slot == kASlots) {
            slot = 0;
            tphase ^= 1u;
          }
        }
        if (has_shift) {
          // SBO = 0: all sixteen 8-row groups of the A operand alias the same 8 "ones" rows
          const uint64_t d_ones = umma_desc_sw128(smem_u32(sOnes)) & ~(static_cast<uint64_t>(0x3FFF) << 32);
          umma_bf16_ss(d_tmem, d_ones, umma_desc_sw128(smem_u32(sBias)), idesc, 1u);
        }
        umma_commit(&acc_full[ab]);
      }
    }
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ================================================================ epilogue: TMEM -> registers -> 256-bit global stores
    const int ew = warp - kEpiWarp0;
    const int qw = warp & 3;
    const int half = ew >> 2;  // warps (e, e+4) share a lane quarter and split the 32-column blocks
    const int ncb = (min(kBN, p.Cout) + 31) / 32;
    const int ncb_half = (ncb + 1) / 2;
    const int cb_begin = half * ncb_half;
    const int cb_end = min(ncb, cb_begin + ncb_half);
    const bool relu = p.epi_relu != 0;
    const int esz = p.out_fp32 ? 4 : 2;
    const int hw = p.Ho * p.Wo;
    uint8_t* out_base = static_cast<uint8_t*>(p.out) + static_cast<size_t>(p.out_coff) * esz;
    const size_t out_row_bytes = static_cast<size_t>(p.out_cstride) * esz;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++tcount) {
      const int ab = tcount & 1;
      mbar_wait(&acc_full[ab], (tcount >> 1) & 1);
      tc_fence_after();
      if (cb_begin >= cb_end) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
        continue;
      }
      const int m = tile * kBM + qw * 32 + lane;
      long long orow = -1;
      if (m < p.M) {
        orow = m;
        if (p.out_pad) {  // zero-padded (F, Ho+2, Wo+2, C) destination of the halo 3x3 kernel
          const int f = m / hw;
          const int rem = m - f * hw;
          const int oy = rem / p.Wo;
          const int ox = rem - oy * p.Wo;
          orow = (static_cast<long long>(f) * (p.Ho + 2) + oy + 1) * (p.Wo + 2) + ox + 1;
        }
      }
      for (int cb = cb_begin; cb < cb_end; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(qw * 32) << 16) + ab * kBN + cb * 32, v);
        tmem_ld_wait();
        if (cb + 1 == cb_end) {  // last read of this accumulator by this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        if (orow < 0) continue;
        uint8_t* dst = out_base + static_cast<size_t>(orow) * out_row_bytes + static_cast<size_t>(cb) * 32 * esz;
        if (p.out_fp32) {
          if (relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            stg256(dst + g * 32, v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3], v[8 * g + 4], v[8 * g + 5], v[8 * g + 6],
                   v[8 * g + 7]);
        } else {
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = cvt_pack(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), relu);
          stg256(dst, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
          stg256(dst + 32, o[8], o[9], o[10], o[11], o[12], o[13], o[14], o[15]);
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

int ring_stages(int nchunks) {
  const int avail = kSmemLimit - kFixedBytes - nchunks * (kBBytes + kParBytesPerChunk);
  int ns = avail / kABytes;
  return ns > kMaxStages ? kMaxStages : ns;
}

}  // namespace

bool conv1x1_ts_eligible(const ConvGemmParams& p) {
  if (getenv("TN_NO_TS") != nullptr) return false;  // read per call: tools/ab_bench.py flips it in-process
  const int esz = p.out_fp32 ? 4 : 2;
  return p.tma_taps == 0 && p.mode == kModeConv && p.R == 1 && p.S == 1 && p.stride == 1 && p.pad == 0 && p.H == p.Ho &&
         p.W == p.Wo && (p.in_cstride % 64) == 0 && (reinterpret_cast<uintptr_t>(p.in) & 15) == 0 &&
         p.num_chunks * 64 <= p.in_cstride && p.pro_scale != nullptr && p.pro_shift != nullptr && p.pro_clamp == nullptr &&
         p.res == nullptr && p.Cout > 64 && p.Cout <= kBN && (p.Cout % 32) == 0 && p.num_chunks >= 1 &&
         p.num_chunks <= kMaxChunks && (p.Cin % 16) == 0 && p.out_pad != 2 &&
         (reinterpret_cast<uintptr_t>(p.out) & 31) == 0 && ((static_cast<size_t>(p.out_cstride) * esz) % 32) == 0 &&
         ((static_cast<size_t>(p.out_coff) * esz) % 32) == 0 && ring_stages(p.num_chunks) >= 3;
}

cudaError_t launch_conv1x1_ts(const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  EncodeTiledFn encode = get_encode();
  if (!encode) return cudaErrorNotSupported;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(p.in_cstride), static_cast<cuuint64_t>(p.M)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(p.in_cstride) * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {64, kBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(p.in), gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  int ns = ring_stages(p.num_chunks);
  if (p.stage_cap >= 2 && p.stage_cap < ns) ns = p.stage_cap;
  const int smem = kFixedBytes + p.num_chunks * (kBBytes + kParBytesPerChunk) + ns * kABytes;
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int m_tiles = (p.M + kBM - 1) / kBM;
  const int grid = m_tiles < num_sms ? m_tiles : num_sms;
  return launch_pdl(conv1x1_ts_kernel, dim3(grid), dim3(kThreads), smem, stream, p, tmap, ns);
}

}  // namespace tn
