// Persistent TMA + tcgen05 kernel for the DenseNet growth conv (3x3, 128 -> 32) on a zero-padded input layout.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tn_common.h"

namespace tn {

struct Conv3x3Dev {
  const uint8_t* wpack = nullptr;  // 6 blobs [dy][64-ch half], each 96 rows (dx*32 + c_out) x 64 bf16, 128B-swizzled
};

bool conv3x3_halo_supported(int H, int W);
bool make_conv3x3(DeviceArena& arena, const float* w_oihw_32x128x3x3, Conv3x3Dev* out);
cudaError_t launch_zero_border(__nv_bfloat16* buf, int F, int Hp, int Wp, int C, cudaStream_t st);
cudaError_t launch_conv3x3_halo(const Conv3x3Dev& cv, const __nv_bfloat16* in_padded, int F, int H, int W,
                                __nv_bfloat16* out, int out_cstride, int out_coff, int num_sms, cudaStream_t st);

}  // namespace tn
