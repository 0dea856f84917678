// Tensor-core GEMM for the TRAINING path of the per-frame CNN (SURVEY.md §8a V7 with a trainable backbone: train.py:415-424,
// `ag.backward` through gluoncv's DenseNet-121 / ResNet-18 v2): forward, data-gradient and weight-gradient contractions of the
// convolutions on tcgen05 instead of the fp32 SIMT SGEMM of tn_seq_train.cu.
//
// Arithmetic: the reference is fp32.  Operands are fp32 in HBM (activations, gradients, weights); `tn_split_bf16` rewrites an
// operand as two bf16 planes  x = hi + lo  (hi = rn_bf16(x), lo = rn_bf16(x - hi): 16 mantissa bits together), K-major, optionally
// transposed and/or re-indexed into the zero-padded (H+2)x(W+2) row space of a 3x3 convolution.  The GEMM accumulates
//   hi*hi + hi*lo + lo*hi   (passes = 3; the dropped lo*lo term is 2^-16 relative)  or  hi*hi only (passes = 1, plain bf16)
// in fp32 in tensor memory.
//
//   D[m, n] = sum_taps sum_k A[m + a_row_off[t], a_k_off[t] + k] * B[n + b_row_off[t], b_k_off[t] + k]
//
// * forward 1x1 / im2col:  A = activations (pixels x K), B = weights (Cout x K), one tap
// * forward 3x3 stride 1:  A = padded activations, a_row_off = dy*(W+2)+dx, B = weights (Cout x 9*Cin), b_k_off = t*Cin
//                          (the im2col matrix is never formed: a tap is a row shift in padded-flattened row space)
// * data gradient:         A = dY (pixels x Cout) [padded, a_row_off = -(dy*(W+2)+dx) for a 3x3], B = W^T (Cin x taps*Cout)
// * weight gradient:       A = X^T (Cin x pixels), B = dY^T (Cout x pixels), contraction over the pixels, split across CTAs
//                          (split-K into a workspace, deterministic reduction); a 3x3 tap (dy, dx) is a shift of the contraction
//                          index: TMA needs 16-byte aligned inner coordinates, so the padded row pitch of these planes is a
//                          multiple of 8 pixels (dy shifts stay aligned) and dY^T exists in three copies shifted by dx
//
// Kernel: one CTA per (128-row M tile, BN-column N tile, K split); warp 0 = TMA producer (SWIZZLE_128B boxes of 64 k x 128 / BN
// rows through a 3-stage mbarrier ring), warp 1 = TMEM allocation + the single MMA-issuing thread (tcgen05.mma kind::f16,
// M128 x BN x K16), warps 2-5 = epilogue (tcgen05.ld -> alpha/beta -> fp32 stores, or raw partial sums).  96 KB of shared memory
// and <= 128 TMEM columns per CTA: two CTAs share an SM so one tile's epilogue overlaps the other's main loop.
#include <cuda.h>

#include "tn_common.h"
#include "tn_ptx.cuh"

namespace {

using namespace tn;

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kMaxTaps = 9;
constexpr uint32_t kABytes = kBM * kBK * 2;  // 16 KB

struct TcParams {
  int M, N, K;       // tile space: rows of A, rows of B, contraction length per tap
  int BN;            // 32 / 64 / 96 / 128
  int ntaps, passes;
  int kchunks;       // ceil(K / 64)
  int iters_total;   // ntaps * kchunks
  int splits, iters_per_split;
  int a_row_off[kMaxTaps], a_k_off[kMaxTaps], b_row_off[kMaxTaps], b_k_off[kMaxTaps];
  float alpha, beta;
  float* C;          // direct output (splits == 1 and unit column stride) ...
  long long ldc;
  float* partial;    // ... or [splits][M][Npad] raw sums
  int Npad;
  int unpad_h, unpad_w;  // > 0: m indexes the padded (h+2)x(w+2) row space; border rows are dropped, the rest re-indexed
  int vec4;              // direct output rows are 16-byte aligned
  int tile_taps, ntn;    // tile_taps: the taps are independent outputs (blockIdx.y = tap * ntn + n tile), not accumulated
  long long c_tap_stride;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// padded row index -> unpadded row index, or -1 for a border / out-of-range row
__device__ __forceinline__ long long unpad_row(long long m, int h, int w) {
  const int wp = w + 2, hp = h + 2;
  const int xp = static_cast<int>(m % wp);
  const long long t = m / wp;
  const int yp = static_cast<int>(t % hp);
  const long long n = t / hp;
  if (xp < 1 || xp > w || yp < 1 || yp > h) return -1;
  return (n * h + (yp - 1)) * w + (xp - 1);
}

__global__ void __launch_bounds__(kThreads, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int BN = p.BN;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * kBK * 2;
  const uint32_t stage_bytes = kABytes + ((b_bytes + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tsel = p.tile_taps ? static_cast<int>(blockIdx.y) / p.ntn : -1;  // this CTA's tap when taps are output tiles
  const int m0 = blockIdx.x * kBM, n0 = (p.tile_taps ? static_cast<int>(blockIdx.y) % p.ntn : static_cast<int>(blockIdx.y)) * BN;
  const int split = blockIdx.z;
  const int it0 = split * p.iters_per_split;
  int it1 = it0 + p.iters_per_split;
  if (it1 > p.iters_total) it1 = p.iters_total;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (BN <= 32) tmem_alloc<32>(&tmem_slot);
    else if (BN <= 64) tmem_alloc<64>(&tmem_slot);
    else tmem_alloc<128>(&tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int slot = 0;
      for (int it = it0; it < it1; ++it) {
        const int tap_it = it / p.kchunks, chunk = it - tap_it * p.kchunks;
        const int tap = tsel >= 0 ? tsel : tap_it;
        const int k0 = chunk * kBK;
        for (int ps = 0; ps < p.passes; ++ps, ++slot) {
          const int s = slot % kStages;
          const uint32_t par = static_cast<uint32_t>((slot / kStages) & 1);
          mbar_wait(&empty_bar[s], par ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], kABytes + b_bytes);
          const uint32_t a_dst = base + s * stage_bytes, b_dst = a_dst + kABytes;
          tma_load_2d(a_dst, ps == 2 ? &tmAl : &tmAh, p.a_k_off[tap] + k0, m0 + p.a_row_off[tap], &full_bar[s]);
          tma_load_2d(b_dst, ps == 1 ? &tmBl : &tmBh, p.b_k_off[tap] + k0, n0 + p.b_row_off[tap], &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128(BN);
      int slot = 0;
      for (int it = it0; it < it1; ++it) {
        const int tap = it / p.kchunks, chunk = it - tap * p.kchunks;
        int ksteps = (p.K - chunk * kBK + 15) >> 4;  // the tail chunk issues only the K16 steps that hold data
        if (ksteps > 4) ksteps = 4;
        for (int ps = 0; ps < p.passes; ++ps, ++slot) {
          const int s = slot % kStages;
          const uint32_t par = static_cast<uint32_t>((slot / kStages) & 1);
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          const uint32_t a_addr = base + s * stage_bytes, b_addr = a_addr + kABytes;
          for (int ks = 0; ks < ksteps; ++ks)
            umma_bf16_ss(tmem, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc,
                         (slot > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
        }
      }
      umma_commit(&acc_bar);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const long long m = static_cast<long long>(m0) + q * 32 + lane;
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    const bool direct = p.partial == nullptr;
    long long orow = m;
    bool row_ok = m < p.M;
    if (direct && p.unpad_h > 0 && row_ok) {
      orow = unpad_row(m, p.unpad_h, p.unpad_w);
      row_ok = orow >= 0;
    }
    const int tcol = tsel >= 0 ? tsel : 0, ntc = p.tile_taps ? p.ntaps : 1;
    float* dst = direct ? p.C + orow * p.ldc + tcol * p.c_tap_stride + n0
                        : p.partial + ((static_cast<long long>(split) * p.M + m) * ntc + tcol) * p.Npad + n0;
    const float alpha = direct ? p.alpha : 1.f, beta = direct ? p.beta : 0.f;
    const bool vec = direct ? (p.vec4 != 0) : true;
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      tmem_ld_wait();
      if (!row_ok) continue;
      const int nrem = p.N - (n0 + c);
      if (vec && nrem >= 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o = make_float4(alpha * __uint_as_float(v[j]), alpha * __uint_as_float(v[j + 1]), alpha * __uint_as_float(v[j + 2]),
                                 alpha * __uint_as_float(v[j + 3]));
          float4* d4 = reinterpret_cast<float4*>(dst + c + j);
          if (beta != 0.f) {
            const float4 old = *d4;
            o.x = fmaf(beta, old.x, o.x); o.y = fmaf(beta, old.y, o.y); o.z = fmaf(beta, old.z, o.z); o.w = fmaf(beta, old.w, o.w);
          }
          *d4 = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nrem) {
            float o = alpha * __uint_as_float(v[j]);
            if (beta != 0.f) o = fmaf(beta, dst[c + j], o);
            dst[c + j] = o;
          }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (BN <= 32) tmem_dealloc<32>(tmem);
    else if (BN <= 64) tmem_dealloc<64>(tmem);
    else tmem_dealloc<128>(tmem);
  }
}

// C[orow(m) * rs + n * cs + t * ts] = alpha * sum_s partial[s][m][t][n] + beta * C[...]
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, int Npad, int T, float alpha,
                                     float beta, float* __restrict__ C, long long rs, long long cs, long long ts, int unpad_h,
                                     int unpad_w) {
  const long long total = static_cast<long long>(M) * N * T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % N);
    const long long mt = i / N;
    const int t = static_cast<int>(mt % T);
    const long long m = mt / T;
    long long orow = m;
    if (unpad_h > 0) {
      orow = unpad_row(m, unpad_h, unpad_w);
      if (orow < 0) continue;
    }
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[((static_cast<long long>(k) * M + m) * T + t) * Npad + n];
    float* d = C + orow * rs + n * cs + t * ts;
    *d = beta != 0.f ? fmaf(beta, *d, alpha * s) : alpha * s;
  }
}

// ---------------------------------------------------------------------------------------------------- operand preparation
// Both kernels walk the OUTPUT index, so the zero border of a padded plane is written by the same pass (no memset).
// padded index j -> source row (n, y, x), or -1 on the border / outside the grid (32-bit arithmetic: the callers check the range)
__device__ __forceinline__ long long src_row_of(long long j, long long padded_total, int h, int w, int pitch) {
  if (h <= 0) return j;
  if (j < 0 || j >= padded_total) return -1;
  const unsigned ju = static_cast<unsigned>(j);
  const unsigned t = ju / static_cast<unsigned>(pitch);
  const int xp = static_cast<int>(ju - t * static_cast<unsigned>(pitch));
  const unsigned n = t / static_cast<unsigned>(h + 2);
  const int yp = static_cast<int>(t - n * static_cast<unsigned>(h + 2));
  if (xp < 1 || xp > w || yp < 1 || yp > h) return -1;
  return (static_cast<long long>(n) * h + (yp - 1)) * w + (xp - 1);
}
__device__ __forceinline__ void split1(float v, __nv_bfloat16* h, __nv_bfloat16* l) {
  const __nv_bfloat16 hh = __float2bfloat16_rn(v);
  *h = hh;
  *l = __float2bfloat16_rn(v - __bfloat162float(hh));
}

// out[j][c] = split(src[row(j) * ld + c]) (zeros on the border); four columns per thread
__global__ void split_rows_kernel(const float* __restrict__ src, long long ld, long long out_rows, int cols, int pad_h, int pad_w,
                                  int pitch, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long out_ld, int vec_in) {
  const int c4n = (cols + 3) >> 2;
  const long long total = out_rows * c4n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long j;
    int c;
    if (total < 0x7fffffffLL) {
      const unsigned iu = static_cast<unsigned>(i), ju = iu / static_cast<unsigned>(c4n);
      j = ju;
      c = static_cast<int>(iu - ju * static_cast<unsigned>(c4n)) * 4;
    } else {
      j = i / c4n;
      c = static_cast<int>(i - j * c4n) * 4;
    }
    const long long r = src_row_of(j, out_rows, pad_h, pad_w, pitch);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r >= 0) {
      const float* s = src + r * ld + c;
      if (vec_in && c + 4 <= cols) {
        const float4 t = *reinterpret_cast<const float4*>(s);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        for (int q = 0; q < 4; ++q)
          if (c + q < cols) v[q] = s[q];
      }
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split1(v[q], &h[q], &l[q]);
    __nv_bfloat16* ph = hi + j * out_ld + c;
    if (c + 4 <= cols) {  // out_ld is a multiple of 8 and c of 4: 8-byte aligned
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<const uint2*>(h);
      if (lo) *reinterpret_cast<uint2*>(lo + j * out_ld + c) = *reinterpret_cast<const uint2*>(l);
    } else {
      for (int q = 0; q < 4; ++q)
        if (c + q < cols) {
          ph[q] = h[q];
          if (lo) lo[j * out_ld + c + q] = l[q];
        }
    }
  }
}

// out[c][j] = split(src[row(j - shift) * ld + c]): 64 (j) x 64 (c) tiles through shared memory, two j per thread on the way out
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, long long ld, long long out_k, int cols,
                                                              int pad_h, int pad_w, int pitch, int shift,
                                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                              long long out_ld) {
  __shared__ float tile[64][65];
  __shared__ long long srow[64];
  const long long j0 = static_cast<long long>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  if (threadIdx.x < 64) {
    const long long j = j0 + threadIdx.x;
    srow[threadIdx.x] = j < out_k ? src_row_of(j - shift, out_k, pad_h, pad_w, pitch) : -1;
  }
  __syncthreads();
  {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 columns x 4 rows per pass
    const int c = c0 + tx;
    float v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {  // sixteen independent loads in flight
      const long long r = srow[ty + 4 * q];
      v[q] = (r >= 0 && c < cols) ? src[r * ld + c] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) tile[ty + 4 * q][tx] = v[q];
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 j pairs x 8 columns per pass
  const long long j = j0 + 2 * tx;
  if (j >= out_k) return;
  const bool pair = j + 1 < out_k;  // out_ld is even and j0 a multiple of 64: the pair is 4-byte aligned
  for (int cc = ty; cc < 64; cc += 8) {
    const int c = c0 + cc;
    if (c >= cols) break;
    __nv_bfloat16 h[2], l[2];
    split1(tile[2 * tx][cc], &h[0], &l[0]);
    split1(tile[2 * tx + 1][cc], &h[1], &l[1]);
    const long long o = static_cast<long long>(c) * out_ld + j;
    if (pair) {
      *reinterpret_cast<uint32_t*>(hi + o) = *reinterpret_cast<const uint32_t*>(h);
      if (lo) *reinterpret_cast<uint32_t*>(lo + o) = *reinterpret_cast<const uint32_t*>(l);
    } else {
      hi[o] = h[0];
      if (lo) lo[o] = l[0];
    }
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

bool make_map(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, long long kdim, long long rows, long long ld, int box_rows) {
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kdim), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {kBK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_bn(int N) { return N <= 32 ? 32 : N <= 64 ? 64 : N <= 96 ? 96 : 128; }

int num_sms_cached() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// split-K plan: enough CTAs for two per SM when the tile grid alone cannot fill the machine
int plan_splits(int M, int N, int T, int iters_total, int BN, bool force_partial, size_t ws_bytes, int Npad) {
  const long long tiles = static_cast<long long>((M + kBM - 1) / kBM) * ((N + BN - 1) / BN) * T;
  const int target = 2 * num_sms_cached();
  int splits = 1;
  if (tiles < target && iters_total >= 8) {
    splits = static_cast<int>((target + tiles - 1) / tiles);
    if (splits > iters_total / 4) splits = iters_total / 4;
    if (splits > 128) splits = 128;
    if (splits < 1) splits = 1;
  }
  const size_t per = static_cast<size_t>(M) * Npad * T * sizeof(float);
  if (splits > 1 || force_partial) {
    const size_t fit = per ? ws_bytes / per : 0;
    if (fit < 1) return force_partial ? -1 : 1;
    if (static_cast<size_t>(splits) > fit) splits = static_cast<int>(fit);
  }
  return splits;
}

}  // namespace

extern "C" {

long long tn_gemm_tc_workspace_bytes(int M, int N, int tile_taps) {
  if (M <= 0 || N <= 0) return 0;
  const long long T = tile_taps > 1 ? tile_taps : 1;
  const long long Npad = (N + 3) / 4 * 4;
  const int BN = pick_bn(N);
  const long long tiles = static_cast<long long>((M + kBM - 1) / kBM) * ((N + BN - 1) / BN) * T;
  const int target = 2 * num_sms_cached();
  long long splits = tiles < target ? (target + tiles - 1) / tiles : 1;
  if (splits > 128) splits = 128;
  return splits * M * Npad * T * static_cast<long long>(sizeof(float));
}

int tn_split_bf16(const float* src, long long ld, long long rows, int cols, int transpose, int pad_h, int pad_w, int pad_pitch,
                  int shift, void* hi, void* lo, long long out_ld, tn_stream_t stream) {
  if (rows < 0 || cols < 0) return set_error(TN_ERR_INVALID, "negative size");
  if (rows == 0 || cols == 0) return TN_OK;
  if (!src || !hi) return set_error(TN_ERR_INVALID, "null device pointer");
  if (out_ld % 8 != 0) return set_error(TN_ERR_INVALID, "plane row stride %lld is not a multiple of 8 elements (TMA: 16 bytes)", out_ld);
  if ((pad_h > 0) != (pad_w > 0)) return set_error(TN_ERR_INVALID, "pad_h and pad_w must both be set");
  if (pad_h > 0 && rows % (static_cast<long long>(pad_h) * pad_w) != 0)
    return set_error(TN_ERR_INVALID, "rows %lld is not a whole number of %dx%d frames", rows, pad_h, pad_w);
  if (pad_h > 0 && pad_pitch == 0) pad_pitch = pad_w + 2;
  if (pad_h > 0 && pad_pitch < pad_w + 2) return set_error(TN_ERR_INVALID, "pad_pitch %d < pad_w + 2", pad_pitch);
  if (shift != 0 && !(transpose && pad_h > 0 && shift >= -1 && shift <= 1))
    return set_error(TN_ERR_INVALID, "a shift of -1/+1 needs a transposed, padded plane");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof_scope(kProfOther, st);
  __nv_bfloat16* h = static_cast<__nv_bfloat16*>(hi);
  __nv_bfloat16* l = static_cast<__nv_bfloat16*>(lo);
  // extent of the pixel index on the output side (the whole padded grid, borders included)
  const long long out_n = pad_h > 0 ? rows / (static_cast<long long>(pad_h) * pad_w) * (pad_h + 2) * pad_pitch : rows;
  if (pad_h > 0 && out_n >= 0x7fffffffLL) return set_error(TN_ERR_INVALID, "padded pixel count %lld exceeds 2^31", out_n);
  if (!transpose) {
    const int vec_in = (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0);
    const long long total = out_n * ((cols + 3) / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    split_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, ld, out_n, cols, pad_h, pad_w, pad_pitch, h, l, out_ld, vec_in);
  } else {
    dim3 grid(static_cast<unsigned>((out_n + 63) / 64), static_cast<unsigned>((cols + 63) / 64));
    split_transpose_kernel<<<grid, 256, 0, st>>>(src, ld, out_n, cols, pad_h, pad_w, pad_pitch, shift, h, l, out_ld);
  }
  TN_CUDA(cudaGetLastError());
  return TN_OK;
}

int tn_gemm_tc(int M, int N, int K, int ntaps, const int* taps, int passes, const void* a_hi, const void* a_lo, long long a_rows,
               long long a_kdim, long long a_ld, const void* b_hi, const void* b_lo, long long b_rows, long long b_kdim,
               long long b_ld, float alpha, float beta, float* C, long long c_row_stride, long long c_col_stride, int tile_taps,
               long long c_tap_stride, int unpad_h, int unpad_w, void* workspace, long long workspace_bytes, tn_stream_t stream) {
  if (M < 0 || N < 0 || K < 0) return set_error(TN_ERR_INVALID, "negative GEMM size");
  if (M == 0 || N == 0) return TN_OK;
  if (passes != 1 && passes != 3) return set_error(TN_ERR_INVALID, "passes must be 1 (bf16) or 3 (split bf16), got %d", passes);
  if (ntaps < 1 || ntaps > kMaxTaps) return set_error(TN_ERR_INVALID, "ntaps %d outside [1, %d]", ntaps, kMaxTaps);
  if (!a_hi || !b_hi || !C || (passes == 3 && (!a_lo || !b_lo))) return set_error(TN_ERR_INVALID, "null device pointer");
  if (a_ld % 8 || b_ld % 8) return set_error(TN_ERR_INVALID, "plane row strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a_hi) | reinterpret_cast<uintptr_t>(b_hi) | reinterpret_cast<uintptr_t>(a_lo) |
       reinterpret_cast<uintptr_t>(b_lo)) % 16)
    return set_error(TN_ERR_INVALID, "operand planes must be 16-byte aligned");
  if ((unpad_h > 0) != (unpad_w > 0)) return set_error(TN_ERR_INVALID, "unpad_h and unpad_w must both be set");
  int dev = 0;
  TN_CUDA(cudaGetDevice(&dev));
  if (int rc = check_arch(dev)) return rc;
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(TN_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  TcParams p = {};
  p.M = M; p.N = N; p.K = K;
  p.BN = pick_bn(N);
  p.ntaps = ntaps;
  p.passes = passes;
  p.kchunks = (K + kBK - 1) / kBK;
  if (p.kchunks < 1) p.kchunks = 1;
  p.tile_taps = tile_taps ? 1 : 0;
  p.ntn = (N + p.BN - 1) / p.BN;
  p.c_tap_stride = c_tap_stride;
  const int T = p.tile_taps ? ntaps : 1;
  p.iters_total = (p.tile_taps ? 1 : ntaps) * p.kchunks;
  for (int t = 0; t < ntaps; ++t) {
    p.a_row_off[t] = taps ? taps[4 * t + 0] : 0;
    p.a_k_off[t] = taps ? taps[4 * t + 1] : 0;
    p.b_row_off[t] = taps ? taps[4 * t + 2] : 0;
    p.b_k_off[t] = taps ? taps[4 * t + 3] : 0;
  }
  for (int t = 0; t < ntaps; ++t)
    if (p.a_k_off[t] % 8 || p.b_k_off[t] % 8)
      return set_error(TN_ERR_INVALID, "tap %d: contraction offsets (%d, %d) must be multiples of 8 elements (TMA: 16-byte inner "
                       "coordinate)", t, p.a_k_off[t], p.b_k_off[t]);
  p.alpha = alpha;
  p.beta = beta;
  p.Npad = (N + 3) / 4 * 4;
  p.unpad_h = unpad_h;
  p.unpad_w = unpad_w;
  const bool force_partial = c_col_stride != 1;
  const size_t ws = workspace ? static_cast<size_t>(workspace_bytes < 0 ? 0 : workspace_bytes) : 0;
  int splits = plan_splits(M, N, T, p.iters_total, p.BN, force_partial, ws, p.Npad);
  if (splits < 0) return set_error(TN_ERR_INVALID, "a strided output needs a workspace of at least %lld bytes",
                                   static_cast<long long>(M) * p.Npad * T * 4);
  p.iters_per_split = (p.iters_total + splits - 1) / splits;
  splits = (p.iters_total + p.iters_per_split - 1) / p.iters_per_split;  // no empty split
  p.splits = splits;
  const bool use_partial = splits > 1 || force_partial;
  p.partial = use_partial ? static_cast<float*>(workspace) : nullptr;
  p.C = C;
  p.ldc = c_row_stride;
  p.vec4 = (c_row_stride % 4 == 0) && (reinterpret_cast<uintptr_t>(C) % 16 == 0) && (p.BN % 4 == 0);

  CUtensorMap mAh, mAl, mBh, mBl;
  if (!make_map(enc, &mAh, a_hi, a_kdim, a_rows, a_ld, kBM) || !make_map(enc, &mAl, a_lo ? a_lo : a_hi, a_kdim, a_rows, a_ld, kBM) ||
      !make_map(enc, &mBh, b_hi, b_kdim, b_rows, b_ld, p.BN) || !make_map(enc, &mBl, b_lo ? b_lo : b_hi, b_kdim, b_rows, b_ld, p.BN))
    return set_error(TN_ERR_CUDA, "cuTensorMapEncodeTiled failed (A %lldx%lld ld %lld, B %lldx%lld ld %lld)", a_rows, a_kdim, a_ld,
                     b_rows, b_kdim, b_ld);

  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * kBK * 2;
  const int smem = kStages * static_cast<int>(kABytes + ((b_bytes + 1023u) & ~1023u)) + 1024;
  static int configured = 0;
  if (smem > configured) {
    TN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid((M + kBM - 1) / kBM, p.ntn * T, splits);
  {
    ProfScope prof_scope(kProfConvGemm, st);
    gemm_tc_kernel<<<grid, kThreads, smem, st>>>(mAh, mAl, mBh, mBl, p);
    TN_CUDA(cudaGetLastError());
  }
  if (use_partial) {
    ProfScope prof_scope(kProfOther, st);
    const long long total = static_cast<long long>(M) * N * T;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    splitk_reduce_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(p.partial, splits, M, N, p.Npad, T, alpha, beta, C, c_row_stride,
                                                                      c_col_stride, c_tap_stride, unpad_h, unpad_w);
    TN_CUDA(cudaGetLastError());
  }
  return TN_OK;
}

}  // extern "C"
