// Hardware probe (development aid): can a K-major SWIZZLE_128B UMMA operand start at an arbitrary ROW of a
// larger shared-memory tile (start address not 1024-B aligned)?  The 3x3 convolution wants this: one halo tile in
// smem, nine tap-shifted A operands that differ only in their start row.  Also checks that a 2D TMA tile load with
// SWIZZLE_128B (incl. out-of-bounds rows -> zero fill) produces the same smem image as the software swizzle.
//   nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe umma_shift_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../tennis_b200/csrc/tn_ptx.cuh"

using namespace tn;

constexpr int ROWS = 192;  // smem A tile rows
constexpr int N = 32;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: software-swizzled fill, mode 1: TMA fill (row0 = first global row to load, may be negative)
__global__ void probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int use_base_off, int mode,
                      const __grid_constant__ CUtensorMap tmap, int row0, uint8_t* smem_dump) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                  // ROWS*128
  uint8_t* sB = smem + ROWS * 128;     // N*128   (ROWS*128 is a multiple of 1024)
  uint64_t* bar = (uint64_t*)(sB + N * 128);
  uint64_t* tbar = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(tbar, 1);
    mbar_fence_init();
  }
  if (tid < 32) tmem_alloc<32>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm = *slot;
  if (mode == 0) {
    for (int u = tid; u < ROWS * 8; u += blockDim.x) {
      int r = u >> 3, g = u & 7;
      uint4 v = *(const uint4*)(A + (size_t)r * 64 + g * 8);
      *(uint4*)(sA + r * 128 + ((g ^ (r & 7)) << 4)) = v;
    }
  } else {
    if (tid == 0) {
      mbar_arrive_expect_tx(tbar, ROWS * 128);
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
          ::"r"(smem_u32(sA)), "l"(&tmap), "r"(0), "r"(row0), "r"(smem_u32(tbar))
          : "memory");
    }
    mbar_wait(tbar, 0);
  }
  for (int u = tid; u < N * 8; u += blockDim.x) {
    int r = u >> 3, g = u & 7;
    uint4 v = *(const uint4*)(B + (size_t)r * 64 + g * 8);
    *(uint4*)(sB + r * 128 + ((g ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (smem_dump) {
    for (int i = tid; i < ROWS * 128; i += blockDim.x) smem_dump[i] = sA[i];
  }
  if (tid == 0) {
    uint32_t a_addr = smem_u32(sA) + shift * 128;
    uint32_t bo = use_base_off ? ((a_addr >> 7) & 7) : 0;
    uint32_t idesc = umma_idesc_bf16_m128(N);
    for (int k = 0; k < 4; ++k) {
      umma_bf16_ss(tm, desc_sw128(a_addr, bo) + 2 * k, desc_sw128(smem_u32(sB), 0) + 2 * k, idesc, k > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  int warp = tid >> 5, lane = tid & 31;
  uint32_t v[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc<32>(tm);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int GROWS = 400;  // global A rows
  std::vector<__nv_bfloat16> hA(GROWS * 64), hB(N * 64);
  std::vector<float> fA(GROWS * 64), fB(N * 64);
  srand(1);
  for (int i = 0; i < GROWS * 64; ++i) {
    float v = (rand() % 2001 - 1000) / 1000.f;
    hA[i] = __float2bfloat16(v);
    fA[i] = __bfloat162float(hA[i]);
  }
  for (int i = 0; i < N * 64; ++i) {
    float v = (rand() % 2001 - 1000) / 1000.f;
    hB[i] = __float2bfloat16(v);
    fB[i] = __bfloat162float(hB[i]);
  }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  uint8_t* dDump;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dD, 128 * N * 4);
  cudaMalloc(&dDump, ROWS * 128);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (!encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {64, (cuuint64_t)GROWS};
  cuuint64_t gstride[1] = {64 * 2};
  cuuint32_t box[2] = {64, ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)cr);

  size_t smem = ROWS * 128 + N * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> hD(128 * N);
  std::vector<uint8_t> dump0(ROWS * 128), dump1(ROWS * 128);
  for (int mode = 0; mode < 2; ++mode) {
    for (int ubo = 0; ubo < 2; ++ubo) {
      for (int shift : {0, 1, 2, 3, 5, 8, 9, 30, 59, 64}) {
        int row0 = (mode == 1) ? -7 : 0;  // TMA: start 7 rows before the tensor -> first 7 smem rows must be zero
        probe<<<1, 128, smem>>>(dA, dB, dD, shift, ubo, mode, tmap, row0, mode == 0 ? dDump : dDump);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d ubo %d shift %d: CUDA error %s\n", mode, ubo, shift, cudaGetErrorString(e));
          return 2;
        }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            int gr = m + shift + row0;
            if (gr >= 0 && gr < GROWS)
              for (int k = 0; k < 64; ++k) ref += (double)fA[gr * 64 + k] * fB[n * 64 + k];
            maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
          }
        printf("fill=%s base_off=%d shift=%2d  max err %.5f  %s\n", mode ? "tma" : "sw ", ubo, shift, maxerr,
               maxerr < 1e-2 ? "OK" : "MISMATCH");
      }
    }
  }
  return 0;
}
