// Parameter block + launcher of the tcgen05 implicit-GEMM convolution (see tn_conv_gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tn {

enum ConvMode : int {
  kModeConv = 0,   // generic RxS / stride / zero-pad conv (also plain GEMM: R=S=1, H=1, W=M)
  kModePool2 = 1,  // 1x1 conv applied to the 2x2/stride-2 average of the BN+ReLU-activated input
  kModeStem = 2,   // 7x7/2 conv on a channel-padded (C=4) NHWC image, no prologue
};

struct ConvGemmParams {
  // input activation: NHWC bf16, channel stride in_cstride (>= Cin), channels [0, Cin) are read
  const __nv_bfloat16* in;
  int in_cstride;
  int H, W, Cin;
  // geometry
  int Ho, Wo, R, S, stride, pad;
  int mode;
  // pre-activation applied while gathering A (null -> identity): y = relu?(x * scale[c] + shift[c])
  const float* pro_scale;
  const float* pro_shift;
  int pro_relu;
  // "clamp" form of the same pre-activation (TMA-fed 1x1 convs only; replaces pro_scale/pro_shift when non-null):
  //   relu(s*x + b) = s * (clamp(x, lo, hi) - t),  t = -b/s,  (lo, hi) = (t, +inf) for s > 0, (-inf, t) for s < 0,
  // with s folded into the packed weights and -sum_c W'[n][c]*t[c] folded into epi_shift by make_conv1x1_clamp()
  // (tn_common.cu).  The in-place transform is then two packed bf16 min/max per channel pair and is EXACT (no rounding of
  // the activated operand).  Layout: per group of 8 channels one uint4 of lo (8 bf16) followed by one uint4 of hi.
  const uint4* pro_clamp;
  // packed weights: [n_tile][chunk][BN rows x 128 B, 128B-swizzled K-major image]
  const uint8_t* wpack;
  int num_chunks;      // K-chunks of 64 per output tile
  int chunks_per_tap;  // ceil(Cin / 64) (CONV / POOL2)
  // output: NHWC, written at channel offset out_coff with channel stride out_cstride
  void* out;
  int out_cstride;
  int out_coff;
  int out_fp32;  // 0 -> bf16, 1 -> fp32
  int out_pad;   // 1 -> output rows address a zero-padded (F, Ho+2, Wo+2, C) buffer (interior pixels only)
                 // 2 -> GEMM rows enumerate the (H, W) INPUT grid; only rows with y < Ho and x < Wo are written, compacted
  // "row-tap" TMA mode (space-to-depth stem): the A tile of K-chunk c is the 2-D box at row (tile*128 + (c / chunks_per_tap)
  // * tma_tap_rows) of a tensor map with `tma_rows` rows of `tma_row_bytes` stride (rows may overlap: Toeplitz view).
  int tma_taps;        // 0 = not used (1x1 convs are auto-detected), >0 = number of row taps
  int tma_tap_rows;
  // tma_use_off != 0: tap t starts at row tma_tap_off[t] instead of t * tma_tap_rows (row offsets may be negative or repeat: the
  // split-bf16 "precise" convolutions list every spatial tap three times -- hi*Wh, hi*Wl, lo*Wh -- with the lo plane a fixed
  // number of rows behind the hi plane; rows outside the tensor are zero-filled by the TMA engine)
  int tma_use_off;
  int tma_tap_off[28];
  long long tma_rows;
  int tma_row_bytes;
  int Cout;      // multiple of 32
  // epilogue: y = relu?(acc + shift[n] + residual); a per-channel scale must be folded into the packed weights
  const float* epi_scale;  // must be null
  const float* epi_shift;
  int epi_relu;
  const __nv_bfloat16* res;  // optional residual, NHWC bf16 at the output resolution
  int res_cstride;
  int M;  // F * Ho * Wo
  int stage_cap;    // > 0: use at most this many ring stages (latency experiments)
  int l2_prefetch;  // TMA modes: K-chunks prefetched into L2 ahead of the stage ring (0 = library default, < 0 = off)
};

int conv_gemm_pick_bn(int cout);
size_t conv_gemm_wpack_bytes(int cout, int num_chunks);
cudaError_t launch_conv_gemm(const ConvGemmParams& p, cudaStream_t stream);

}  // namespace tn
