// Elementwise stages of the split-bf16 ("precise", fp32-grade) CNN forward; see tn_precise.cu.
// A tensor is a PAIR of bf16 planes of identical layout, the lo plane `plane` elements behind the hi plane.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace tn {

cudaError_t launch_s2d_convert_x2(const void* in, int is_u8_nhwc, __nv_bfloat16* out, size_t plane, int n, int h, int w, int Hz,
                                  int Wz, const float* scale3, const float* shift3, cudaStream_t st);
cudaError_t launch_maxpool_f32_split(const float* in, int n, int Hz, int Wz, int Hs, int Ws, int C, int Hp, int Wp,
                                     __nv_bfloat16* out, size_t plane, int out_cstride, cudaStream_t st);
cudaError_t launch_bn_relu_split(const __nv_bfloat16* in, size_t in_plane, size_t npix, int C, int cstride, const float* scale,
                                 const float* shift, __nv_bfloat16* out, size_t out_plane, int opitch, cudaStream_t st);
cudaError_t launch_bn_relu_pool2_x2(const __nv_bfloat16* in, size_t in_plane, int n, int H, int W, int C, int cstride,
                                    const float* scale, const float* shift, __nv_bfloat16* out, size_t out_plane, int opitch,
                                    cudaStream_t st);
cudaError_t launch_split_store(const float* in, int n, int Hg, int Wg, int y0, int x0, int Ho, int Wo, int C, int pad,
                               __nv_bfloat16* out, size_t plane, int cstride, int coff, cudaStream_t st);
cudaError_t launch_tail_pool_x2(const __nv_bfloat16* in, size_t plane, int n, int H, int W, int C, int cstride, int kh, int kw,
                                int ph, int pw, const float* scale, const float* shift, float* feats, __nv_bfloat16* feats_bf16,
                                cudaStream_t st);

}  // namespace tn
