// Fused DenseNet layer kernel (1x1 bottleneck conv + 3x3 growth conv, bottleneck kept in shared memory).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tn {

bool dense_fused_supported(int H, int W);
// blk: dense-block concat buffer (F,H,W,blk_cstride) bf16; reads channels [0,Cin), writes 32 channels at offset Cin.
// w1pack: 1x1 weights (BN2 scale folded), w1_chunks x [128 rows x 128 B]; w2pack: 3x3 weights in the halo-kernel layout.
cudaError_t launch_dense_layer_fused(const __nv_bfloat16* blk, int blk_cstride, int F, int H, int W, int Cin, const float* bn1_scale,
                                     const float* bn1_shift, const uint8_t* w1pack, int w1_chunks, const float* bn2_shift,
                                     const uint8_t* w2pack, int num_sms, cudaStream_t st);

}  // namespace tn
