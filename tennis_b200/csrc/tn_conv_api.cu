// C-ABI for a single fused convolution (building block + test hook of the conv GEMM kernel).
#include <string.h>

#include <memory>

#include "tn_common.h"
#include "tn_elementwise.h"

struct tn_conv {
  int device = 0;
  tn::DeviceArena arena;
  tn::ConvDev cv;
  const float *pro_scale = nullptr, *pro_shift = nullptr, *epi_scale = nullptr, *epi_shift = nullptr;
};

extern "C" {

int tn_conv_create(tn_conv_t** out, int device, const float* weight, int Cout, int Cin, int R, int S, int mode,
                   const float* pro_scale, const float* pro_shift, const float* epi_scale, const float* epi_shift) {
  if (!out || !weight) return tn::set_error(TN_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = tn::check_arch(device);
  if (rc != TN_OK) return rc;
  if (Cout <= 0 || Cout % 32 != 0) return tn::set_error(TN_ERR_INVALID, "Cout=%d must be a positive multiple of 32", Cout);
  if (mode == tn::kModeStem) {
    if (Cin > 4 || S > 8) return tn::set_error(TN_ERR_INVALID, "stem mode needs Cin<=4 and S<=8");
  } else if (mode == tn::kModeConv || mode == tn::kModePool2) {
    if (Cin % 16 != 0) return tn::set_error(TN_ERR_INVALID, "Cin=%d must be a multiple of 16", Cin);
    if (mode == tn::kModePool2 && (R != 1 || S != 1 || !pro_scale)) return tn::set_error(TN_ERR_INVALID, "pool2 mode is a 1x1 conv with a prologue");
  } else {
    return tn::set_error(TN_ERR_INVALID, "unknown conv mode %d", mode);
  }
  std::vector<float> zero_shift;
  if (epi_scale && !epi_shift) {
    zero_shift.assign(Cout, 0.f);
    epi_shift = zero_shift.data();
  }
  if ((pro_scale == nullptr) != (pro_shift == nullptr)) return tn::set_error(TN_ERR_INVALID, "prologue scale/shift must come together");
  TN_CUDA(cudaSetDevice(device));
  std::unique_ptr<tn_conv> c(new tn_conv);
  c->device = device;
  if (!tn::make_conv(c->arena, weight, Cout, Cin, R, S, mode, &c->cv, epi_scale)) return TN_ERR_CUDA;
  auto up = [&](const float* h, int n) -> const float* {
    return h ? static_cast<const float*>(c->arena.upload(h, n * sizeof(float))) : nullptr;
  };
  c->pro_scale = up(pro_scale, Cin);
  c->pro_shift = up(pro_shift, Cin);
  c->epi_scale = nullptr;  // folded into the packed weights
  c->epi_shift = up(epi_shift, Cout);
  *out = c.release();
  return TN_OK;
}

void tn_conv_destroy(tn_conv_t* c) { delete c; }

int tn_conv_forward(tn_conv_t* c, const void* x, int in_cstride, int n, int H, int W, int stride, int pad, int pro_relu,
                    int epi_relu, void* out, int out_cstride, int out_coff, int out_fp32, const void* residual,
                    int res_cstride, tn_stream_t stream) {
  if (!c || n < 0) return tn::set_error(TN_ERR_INVALID, "bad conv handle");
  if (n == 0) return TN_OK;
  if (!x || !out) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  const tn::ConvDev& cv = c->cv;
  int Ho, Wo;
  if (cv.mode == tn::kModePool2) {
    Ho = H / 2;
    Wo = W / 2;
    stride = 2;
    pad = 0;
  } else {
    if (stride <= 0 || pad < 0) return tn::set_error(TN_ERR_INVALID, "bad stride/pad");
    Ho = (H + 2 * pad - cv.R) / stride + 1;
    Wo = (W + 2 * pad - cv.S) / stride + 1;
  }
  if (Ho <= 0 || Wo <= 0) return tn::set_error(TN_ERR_INVALID, "empty output");
  if (H > 16000 || W > 16000) return tn::set_error(TN_ERR_INVALID, "image too large");
  const int esz_in = 8, esz_out = out_fp32 ? 4 : 8;
  if (cv.mode != tn::kModeStem && (in_cstride % esz_in != 0 || in_cstride < cv.Cin))
    return tn::set_error(TN_ERR_INVALID, "in_cstride=%d must be a multiple of 8 and >= Cin", in_cstride);
  if (out_cstride % esz_out != 0 || out_coff % esz_out != 0 || out_coff + cv.Cout > out_cstride)
    return tn::set_error(TN_ERR_INVALID, "bad output channel stride/offset");
  if (residual && res_cstride % 8 != 0) return tn::set_error(TN_ERR_INVALID, "bad residual stride");
  tn::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.in = static_cast<const __nv_bfloat16*>(x);
  p.in_cstride = cv.mode == tn::kModeStem ? 4 : in_cstride;
  p.H = H;
  p.W = W;
  p.Cin = cv.Cin;
  p.Ho = Ho;
  p.Wo = Wo;
  p.R = cv.R;
  p.S = cv.S;
  p.stride = stride;
  p.pad = pad;
  p.mode = cv.mode;
  p.pro_scale = c->pro_scale;
  p.pro_shift = c->pro_shift;
  p.pro_relu = pro_relu;
  p.wpack = cv.wpack;
  p.num_chunks = cv.num_chunks;
  p.chunks_per_tap = cv.chunks_per_tap;
  p.out = out;
  p.out_cstride = out_cstride;
  p.out_coff = out_coff;
  p.out_fp32 = out_fp32;
  p.Cout = cv.Cout;
  p.epi_scale = c->epi_scale;
  p.epi_shift = c->epi_shift;
  p.epi_relu = epi_relu;
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.res_cstride = res_cstride;
  p.M = n * Ho * Wo;
  TN_CUDA(tn::launch_conv_gemm(p, static_cast<cudaStream_t>(stream)));
  return TN_OK;
}

int tn_frames_to_nhwc4(const float* frames, void* out_bf16, int n, int h, int w, tn_stream_t stream) {
  if (n < 0 || h <= 0 || w <= 0) return tn::set_error(TN_ERR_INVALID, "bad frame shape");
  if (n == 0) return TN_OK;
  if (!frames || !out_bf16) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  TN_CUDA(tn::launch_convert_nchw_f32(frames, static_cast<__nv_bfloat16*>(out_bf16), n, h, w, nullptr, nullptr,
                                      static_cast<cudaStream_t>(stream)));
  return TN_OK;
}

}  // extern "C"
