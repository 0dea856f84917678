// Hardware probe (development aid): K-major SWIZZLE_32B UMMA operands (32-byte rows = 16 bf16 = one K=16 step) with a
// row-shifted start address, filled by TMA (box 16 elems x R rows, SWIZZLE_32B).  Needed by the space-to-depth stem:
// one raw pixel row per smem row, the 4 horizontal taps are row shifts of the same tile.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <vector>
#include "../tennis_b200/csrc/tn_ptx.cuh"
using namespace tn;
constexpr int ROWS = 144, N = 64;
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
__global__ void probe(const __nv_bfloat16* B, float* D, int shift, const __grid_constant__ CUtensorMap tmap, int row0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;               // ROWS*32
  uint8_t* sB = smem + 8192;        // N*32, software swizzle
  uint64_t* bar = (uint64_t*)(smem + 16384);
  uint64_t* tbar = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(tbar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc<64>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = *slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(tbar, ROWS * 32);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(sA)), "l"(&tmap), "r"(0), "r"(row0), "r"(smem_u32(tbar)) : "memory");
  }
  mbar_wait(tbar, 0);
  for (int u = tid; u < N * 2; u += blockDim.x) {  // B row n: two 16-byte chunks, chunk' = chunk ^ ((n>>2)&1)
    int n = u >> 1, c = u & 1;
    uint4 v = *(const uint4*)(B + (size_t)n * 16 + c * 8);
    *(uint4*)(sB + n * 32 + ((c ^ ((n >> 2) & 1)) << 4)) = v;
  }
  fence_proxy_async_smem(); __syncthreads();
  if (tid == 0) {
    umma_bf16_ss(tm, desc_sw32(smem_u32(sA) + shift * 32, 256), desc_sw32(smem_u32(sB), 256), umma_idesc_bf16_m128(N), 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0); tc_fence_after();
  int warp = tid >> 5, lane = tid & 31;
  uint32_t v[32];
  for (int cb = 0; cb < 2; ++cb) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + cb * 32, v); tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + cb * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (tid < 32) tmem_dealloc<64>(tm);
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int GROWS = 400;
  std::vector<__nv_bfloat16> hA(GROWS * 16), hB(N * 16);
  std::vector<float> fA(GROWS * 16), fB(N * 16);
  srand(2);
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {16, (cuuint64_t)GROWS}; cuuint64_t gstride[1] = {32};
  cuuint32_t box[2] = {16, ROWS}; cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)cr);
  if (cr) return 0;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  std::vector<float> hD(128 * N);
  for (int shift : {0, 1, 2, 3, 5, 8, 13}) {
    int row0 = -2;
    probe<<<1, 128, 20480>>>(dB, dD, shift, tmap, row0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0; int gr = m + shift + row0;
        if (gr >= 0 && gr < GROWS) for (int k = 0; k < 16; ++k) ref += (double)fA[gr * 16 + k] * fB[n * 16 + k];
        maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
      }
    printf("SW32 shift=%2d max err %.5f %s\n", shift, maxerr, maxerr < 1e-2 ? "OK" : "MISMATCH");
  }
  return 0;
}
