// Persistent TMA + tcgen05 stem kernel (7x7/2 conv as a 4x4/1 conv on the space-to-depth image).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tn_common.h"

namespace tn {

struct StemDev {
  const uint8_t* wpack = nullptr;  // 16 taps x (64 rows x 16 bf16), SWIZZLE_32B image, BN scale folded in
  const uint8_t* wpack_pool = nullptr;  // the same taps stacked along N per pixel strip for the fused conv+pool kernel (96 KB)
};

bool make_stem(DeviceArena& arena, const float* w_64x3x7x7, const float* fold_scale, StemDev* out);
// z: zero-padded space-to-depth image (n, Hz, Wz, 16) bf16; out: (n, Ho, Wo, 64) bf16 = relu(conv + shift)
cudaError_t launch_stem_s2d(const StemDev& sd, const __nv_bfloat16* z, int n, int Hz, int Wz, int Ho, int Wo, const float* shift,
                            __nv_bfloat16* out, int num_sms, cudaStream_t st);

// conv + shift + ReLU + MaxPool(3,2,1) fused: out (n, Hp, Wp, out_cstride) channels [0,64); requires Ws <= 128
cudaError_t launch_stem_pool(const StemDev& sd, const __nv_bfloat16* z, int n, int Hz, int Wz, int Hs, int Ws, int Hp, int Wp,
                             const float* shift, __nv_bfloat16* out, int out_cstride, int num_sms, cudaStream_t st);

}  // namespace tn
