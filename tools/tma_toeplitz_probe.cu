// Hardware probe (development aid): does a 2-D TMA tensor map with OVERLAPPING rows (row stride 32 B, row length 128 B)
// encode and load correctly?  The space-to-depth stem wants it: output pixel j reads the 4-pixel window starting at
// input pixel j, i.e. the im2col matrix of one filter row is a Toeplitz view of the image.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
#include "../tennis_b200/csrc/tn_ptx.cuh"
using namespace tn;
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int row0, uint8_t* dump) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint64_t* bar = (uint64_t*)(smem + 128 * 128);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 128 * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"(&tmap), "r"(0), "r"(row0), "r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) dump[i] = smem[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int NPIX = 1000;  // pixels of 16 bf16 (32 B)
  std::vector<__nv_bfloat16> h(NPIX * 16);
  for (int i = 0; i < NPIX * 16; ++i) h[i] = __float2bfloat16((float)(i % 4096));
  __nv_bfloat16* d; uint8_t* dd;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&dd, 128 * 128);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {64, (cuuint64_t)(NPIX - 3)};
  cuuint64_t gstride[1] = {32};  // overlapping rows
  cuuint32_t box[2] = {64, 128}; cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (row stride 32B < row bytes 128B) rc=%d\n", (int)cr);
  if (cr != CUDA_SUCCESS) return 0;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 + 2048);
  for (int row0 : {0, 7, 900}) {
    probe<<<1, 128, 128 * 128 + 2048>>>(tmap, row0, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> out(128 * 128);
    cudaMemcpy(out.data(), dd, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 64; ++k) {
        int g = k >> 3, e2 = k & 7;
        __nv_bfloat16 v; memcpy(&v, &out[r * 128 + ((g ^ (r & 7)) << 4) + e2 * 2], 2);
        int j = row0 + r;
        float ref = (j < NPIX - 3) ? (float)((j * 16 + k) % 4096) : 0.f;
        if (__bfloat162float(v) != __bfloat162float(__float2bfloat16(ref))) ++bad;
      }
    printf("row0=%d mismatches=%d\n", row0, bad);
  }
  return 0;
}
