// Shared host-side helpers for the C-ABI translation units: error text, CUDA checks, device arena,
// weight packing into the UMMA shared-memory image.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/tennis_b200.h"
#include "tn_conv_gemm.h"

namespace tn {

std::string& last_error();
int set_error(int code, const char* fmt, ...);

#define TN_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::tn::set_error(TN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                 \
  } while (0)

int check_arch(int device);  // TN_OK iff compute capability 10.x

// Owns device allocations of a handle.
struct DeviceArena {
  std::vector<void*> ptrs;
  ~DeviceArena();
  // returns nullptr on failure (error text set)
  void* upload(const void* host, size_t bytes);
  void* alloc(size_t bytes);
};

struct BnDev {
  const float* scale = nullptr;  // gamma / sqrt(var + eps)
  const float* shift = nullptr;  // beta - mean * scale
  int C = 0;
};
struct ConvDev {
  const uint8_t* wpack = nullptr;
  int Cin = 0, Cout = 0, R = 1, S = 1;
  int num_chunks = 0, chunks_per_tap = 0;
  int mode = kModeConv;
};

// Fold inference BatchNorm (eps 1e-5, running statistics) into scale/shift.  p -> {gamma,beta,mean,var}, each C.
bool make_bn(DeviceArena& arena, const float* gamma, const float* beta, const float* mean, const float* var, int C,
             BnDev* out, std::vector<float>* host_scale = nullptr, std::vector<float>* host_shift = nullptr);
// Pack OIHW fp32 weights (Cout,Cin,R,S) into per-(n_tile, K-chunk) swizzled bf16 blobs.
// `fold_scale` (optional, host, Cout): per-output-channel scale multiplied into the weights (folded BatchNorm gamma/sigma).
bool make_conv(DeviceArena& arena, const float* w, int Cout, int Cin, int R, int S, int mode, ConvDev* out,
               const float* fold_scale = nullptr);

// 1x1 conv whose BN+ReLU pre-activation runs in the "clamp" form (ConvGemmParams::pro_clamp): packs w (Cout,Cin) with the input
// BatchNorm scale (and the optional output-side scale) folded in, uploads the per-channel clamp bounds and the corrected shift.
bool make_conv1x1_clamp(DeviceArena& arena, const float* w, int Cout, int Cin, const float* in_scale, const float* in_shift,
                        const float* out_scale, const float* out_shift, ConvDev* out, const uint4** clamp_dev,
                        const float** shift_dev);

// ---- optional per-launch device timing (bench.py roofline): CUDA events on the launch stream around each kernel.
enum ProfKind : int { kProfConvGemm = 0, kProfOther = 1 };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
  int kind;
  cudaStream_t st;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) { prof_begin(kind, st); }
  ~ProfScope() { prof_end(kind, st); }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Programmatic dependent launch: the kernel may start (prologue: barrier init, TMEM alloc, weight fetch) while the
// previous kernel on the stream drains; it must execute griddep_wait() (tn_ptx.cuh) before touching activations.
// TN_NO_PDL=1 falls back to plain stream order.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace tn
