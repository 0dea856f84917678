#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tn {

cudaError_t launch_convert_nchw_f32(const float* in, __nv_bfloat16* out, int n, int h, int w, const float* scale3,
                                    const float* shift3, cudaStream_t st);
cudaError_t launch_convert_nhwc_u8(const uint8_t* in, __nv_bfloat16* out, int n, int h, int w, const float* scale3,
                                   const float* shift3, cudaStream_t st);
// frames (fp32 NCHW or uint8 NHWC) -> zero-padded 2x2 space-to-depth image (n, Hz, Wz, 16) bf16 for the stem (see .cu)
cudaError_t launch_s2d_convert(const void* in, int is_u8_nhwc, __nv_bfloat16* out, int n, int h, int w, int Hz, int Wz,
                               const float* scale3, const float* shift3, cudaStream_t st);
cudaError_t launch_maxpool3s2(const __nv_bfloat16* in, __nv_bfloat16* out, int n, int H, int W, int C, int Ho, int Wo,
                              int out_cstride, int out_coff, cudaStream_t st);
// Transition pre-pass: out[n, H/2, W/2, C] = avgpool2x2(relu(in * scale + shift)), bf16, dense channel stride C.
cudaError_t launch_bn_relu_pool2(const __nv_bfloat16* in, int n, int H, int W, int C, int cstride, const float* scale,
                                 const float* shift, __nv_bfloat16* out, cudaStream_t st);
cudaError_t launch_tail_pool(const __nv_bfloat16* in, int n, int H, int W, int C, int cstride, int kh, int kw, int ph,
                             int pw, const float* scale, const float* shift, float* feats, __nv_bfloat16* feats_bf16,
                             cudaStream_t st);
cudaError_t launch_dense(const float* x, const float* W, const float* b, float* y, int rows, int in_dim, int out_dim,
                         cudaStream_t st);
cudaError_t launch_temporal_pool(const float* x, float* y, int B, int T, int D, int mean, cudaStream_t st);
cudaError_t launch_cast_bf16(const float* x, __nv_bfloat16* y, size_t n, cudaStream_t st);
// (M,D) fp32 -> (M,3D) bf16 rows [hi | lo | hi] with hi = bf16(x), lo = bf16(x - hi)
cudaError_t launch_split3_bf16(const float* x, __nv_bfloat16* y, size_t M, int D, cudaStream_t st);

}  // namespace tn
