// Halo-tile TMA + tcgen05 kernel for the 3x3 / stride 1 convolutions with 64 -> 64 and 128 -> 128 channels (ResNet-18 v2 stages 1-2)
// on a pre-activated, zero-padded input layout (see tn_conv3x3_c64.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tn_common.h"

namespace tn {

struct Conv3x3C64Dev {
  const uint8_t* wpack = nullptr;  // blobs [c_out half][dy][c_in half], each 192 rows (dx*64 + c_out) x 64 bf16, 128B-swizzled
  int C = 0;                       // channels in = out: 64 or 128
};

bool conv3x3_c64_supported(int H, int W, int C);
// w (C, C, 3, 3) OIHW, C = 64 or 128; fold_scale (host, C floats or null): per-output-channel scale folded into the weights
bool make_conv3x3_c64(DeviceArena& arena, const float* w_oihw, int C, const float* fold_scale, Conv3x3C64Dev* out);
// out_pad (F, H+2, W+2, C) = relu(scale * x + shift) on the interior, zeros on the border
cudaError_t launch_bn_relu_pad(const __nv_bfloat16* x, int x_cs, int F, int H, int W, int C, const float* scale, const float* shift,
                               __nv_bfloat16* out_pad, cudaStream_t st);
// value = conv(in_padded) + shift (+ res) (relu); out (unpadded) and / or out_pad (padded; when both are given out_pad holds
// relu(act_scale * bf16(value) + act_shift), the pre-activated input of the next convolution)
cudaError_t launch_conv3x3_c64(const Conv3x3C64Dev& cv, const __nv_bfloat16* in_padded, int F, int H, int W, const float* shift,
                               int relu, const __nv_bfloat16* res, int res_cs, __nv_bfloat16* out, int out_cs,
                               __nv_bfloat16* out_pad, const float* act_scale, const float* act_shift, int num_sms,
                               cudaStream_t st);

}  // namespace tn
