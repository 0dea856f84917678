// Per-frame CNN feature extractor: layer plans for DenseNet-121 and ResNet-18 v2 `features`
//  expressed as launches of the tcgen05 conv GEMM + the HBM-bound
// helpers.  Activations live in HBM as NHWC bf16; a dense block is ONE buffer of its final channel count
// and every dense layer writes its 32 new channels in place (concat == channel offset).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <memory>

#include "tn_common.h"
#include "tn_conv3x3.h"
#include "tn_conv3x3_c64.h"
#include "tn_dense_fused.h"
#include "tn_elementwise.h"
#include "tn_precise.h"
#include "tn_stem.h"

namespace {

using namespace tn;

struct DenseLayer {
  BnDev bn1, bn2;
  ConvDev conv1, conv2;
  Conv3x3Dev conv2h;  // same weights, packed for the halo kernel
  // conv1 with BN1+ReLU in clamp form (bn1 scale folded into the weights, threshold term folded into the shift)
  ConvDev conv1c;
  const uint4* clamp1 = nullptr;
  const float* shift1c = nullptr;
  // split-bf16 ("precise") images: every tap three times (hi*Wh, hi*Wl, lo*Wh), see make_conv_x3
  ConvDev conv1x, conv2x;
  int cin;
};
struct Transition {
  BnDev bn;
  ConvDev conv;
  ConvDev convx;
};
struct ResBlock {
  BnDev bn1, bn2;
  ConvDev conv1, conv2, ds;
  Conv3x3C64Dev c1h, c2h;  // stride-1 3x3 convs with 64 / 128 channels: the same weights packed for the halo kernel
  bool has_ds;
  bool h1 = false, h2 = false;  // conv1 / conv2 can run on the halo kernel (tn_conv3x3_c64.cu)
  int cin, c, stride;
};

const int kDenseCfg[4] = {6, 12, 24, 16};
const int kGrowth = 32, kBott = 128;

}  // namespace

struct tn_backbone {
  int arch = 0, device = 0;
  tn::DeviceArena arena;
  // stem (both archs)
  tn::ConvDev stem;
  tn::ConvDev stem_x3;  // precise mode: 4 filter rows x 3 products, K = 64 (4 px x 16 ch) per tap
  tn::StemDev stem_s2d;
  int precise = 0;      // 0 = bf16 speed path, 1 = split-bf16 fp32-grade path (DenseNet only)
  tn::BnDev bn0;
  float in_scale[3] = {1, 1, 1}, in_shift[3] = {0, 0, 0};  // ResNet-v2 bn_data folded into the input conversion
  // DenseNet
  std::vector<DenseLayer> layers[4];
  Transition trans[3];
  // ResNet
  std::vector<ResBlock> rblocks;
  tn::BnDev bn_final;
  int feat_channels = 0;
  int num_sms = 148;
};

namespace {

size_t densenet_param_count() {
  size_t n = 64 * 3 * 49 + 4 * 64;
  int c = 64;
  for (int b = 0; b < 4; ++b) {
    for (int l = 0; l < kDenseCfg[b]; ++l) {
      n += 4 * c + static_cast<size_t>(kBott) * c + 4 * kBott + static_cast<size_t>(kGrowth) * kBott * 9;
      c += kGrowth;
    }
    if (b < 3) {
      n += 4 * c + static_cast<size_t>(c / 2) * c;
      c /= 2;
    }
  }
  n += 4 * c;
  return n;
}

size_t resnet18_param_count() {
  size_t n = 4 * 3 + 64 * 3 * 49 + 4 * 64;
  int cin = 64;
  const int ch[4] = {64, 128, 256, 512};
  for (int s = 0; s < 4; ++s) {
    for (int b = 0; b < 2; ++b) {
      const int c = ch[s];
      n += 4 * cin + static_cast<size_t>(c) * cin * 9 + 4 * c + static_cast<size_t>(c) * c * 9;
      if (b == 0 && cin != c) n += static_cast<size_t>(c) * cin;
      cin = c;
    }
  }
  n += 4 * 512;
  return n;
}

struct Cursor {
  const float* p;
  size_t left;
  const float* take(size_t n) {
    if (n > left) return nullptr;
    const float* r = p;
    p += n;
    left -= n;
    return r;
  }
};

bool take_bn(Cursor& cur, DeviceArena& arena, int C, BnDev* bn, std::vector<float>* hs = nullptr,
             std::vector<float>* hb = nullptr) {
  const float* g = cur.take(C);
  const float* b = cur.take(C);
  const float* m = cur.take(C);
  const float* v = cur.take(C);
  if (!v) return false;
  return make_bn(arena, g, b, m, v, C, bn, hs, hb);
}
bool take_conv(Cursor& cur, DeviceArena& arena, int Cout, int Cin, int R, int S, int mode, ConvDev* cv) {
  const float* w = cur.take(static_cast<size_t>(Cout) * Cin * R * S);
  if (!w) return false;
  return make_conv(arena, w, Cout, Cin, R, S, mode, cv);
}
// Split-bf16 image of a convolution for the precise path: w (Cout,Cin,R,S) [x fold_scale] = Wh + Wl with Wh = bf16(w),
// Wl = bf16(w - Wh); packed as a conv with 3*R*S row taps [spatial tap][Wh, Wl, Wh] -- the activation side supplies
// [hi, hi, lo] for the three taps (tn_precise.cu), so the K loop accumulates hi*Wh + hi*Wl + lo*Wh in fp32.
bool make_conv_x3(DeviceArena& arena, const float* w, int Cout, int Cin, int R, int S, const float* fold_scale, ConvDev* cv) {
  const int sp = R * S, taps = 3 * sp;
  std::vector<float> wx(static_cast<size_t>(Cout) * Cin * taps);
  for (int n = 0; n < Cout; ++n)
    for (int c = 0; c < Cin; ++c)
      for (int t = 0; t < sp; ++t) {
        float v = w[(static_cast<size_t>(n) * Cin + c) * sp + t];
        if (fold_scale) v *= fold_scale[n];
        const float hi = __bfloat162float(__float2bfloat16(v));
        const float lo = __bfloat162float(__float2bfloat16(v - hi));
        float* dst = &wx[(static_cast<size_t>(n) * Cin + c) * taps + 3 * t];
        dst[0] = hi;
        dst[1] = lo;
        dst[2] = hi;
      }
  return make_conv(arena, wx.data(), Cout, Cin, taps, 1, kModeConv, cv, nullptr);
}
// 7x7/2 stem as a 4x4/1 conv on the space-to-depth image: weights (64,3,7,7) -> (64, 64 = 4 px x 16 ch, 4 rows, 1)
bool take_stem_bn(Cursor& cur, DeviceArena& arena, ConvDev* cv, BnDev* bn, StemDev* sd, ConvDev* cvx = nullptr) {
  const float* w = cur.take(static_cast<size_t>(64) * 3 * 49);
  if (!w) return false;
  std::vector<float> hs, hb;
  if (!take_bn(cur, arena, 64, bn, &hs, &hb)) return false;
  if (!make_stem(arena, w, hs.data(), sd)) return false;
  std::vector<float> ws(static_cast<size_t>(64) * 64 * 4, 0.f);  // [n][ci = b*16 + (py*2+px)*3 + c][a][0]
  for (int n = 0; n < 64; ++n)
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) {
            const int r = 2 * a + py - 1, s = 2 * b + px - 1;  // W8[u][v] = W[u-1][v-1]
            if (r < 0 || s < 0 || r > 6 || s > 6) continue;
            for (int c = 0; c < 3; ++c) {
              const int ci = b * 16 + (py * 2 + px) * 3 + c;
              ws[(static_cast<size_t>(n) * 64 + ci) * 4 + a] = w[((static_cast<size_t>(n) * 3 + c) * 7 + r) * 7 + s];
            }
          }
  if (cvx && !make_conv_x3(arena, ws.data(), 64, 64, 4, 1, hs.data(), cvx)) return false;
  return make_conv(arena, ws.data(), 64, 64, 4, 1, kModeConv, cv, hs.data());
}
// conv followed (in parameter order) by the BatchNorm whose scale is folded into its weights; the shift stays in `bn`
bool take_conv_bn(Cursor& cur, DeviceArena& arena, int Cout, int Cin, int R, int S, int mode, ConvDev* cv, BnDev* bn,
                  std::vector<float>* hs_out = nullptr, std::vector<float>* hb_out = nullptr, ConvDev* cvx = nullptr) {
  const float* w = cur.take(static_cast<size_t>(Cout) * Cin * R * S);
  if (!w) return false;
  std::vector<float> hs, hb;
  if (!take_bn(cur, arena, Cout, bn, &hs, &hb)) return false;
  if (hs_out) *hs_out = hs;
  if (hb_out) *hb_out = hb;
  if (cvx && !make_conv_x3(arena, w, Cout, Cin, R, S, hs.data(), cvx)) return false;
  return make_conv(arena, w, Cout, Cin, R, S, mode, cv, hs.data());
}

struct Dims {
  int H0, W0, Hs, Ws, Hp, Wp;
};
Dims stem_dims(int h, int w) {
  Dims d;
  d.H0 = h;
  d.W0 = w;
  d.Hs = (h + 6 - 7) / 2 + 1;
  d.Ws = (w + 6 - 7) / 2 + 1;
  d.Hp = (d.Hs + 2 - 3) / 2 + 1;
  d.Wp = (d.Ws + 2 - 3) / 2 + 1;
  return d;
}

// Bump allocator over the caller's workspace (dry-run when base == nullptr).
struct Bump {
  uint8_t* base;
  size_t off = 0;
  template <typename T>
  T* get(size_t count) {
    off = align_up(off, 1024);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

ConvGemmParams conv_params(const ConvDev& cv, const __nv_bfloat16* in, int in_cstride, int n, int H, int W, int Ho,
                           int Wo, int stride, int pad, const BnDev* pro, void* out, int out_cstride, int out_coff,
                           const BnDev* epi, bool epi_relu) {
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.in = in;
  p.in_cstride = in_cstride;
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
  if (pro) {
    p.pro_scale = pro->scale;
    p.pro_shift = pro->shift;
    p.pro_relu = 1;
  }
  p.wpack = cv.wpack;
  p.num_chunks = cv.num_chunks;
  p.chunks_per_tap = cv.chunks_per_tap;
  p.out = out;
  p.out_cstride = out_cstride;
  p.out_coff = out_coff;
  p.out_fp32 = 0;
  p.Cout = cv.Cout;
  if (epi) p.epi_shift = epi->shift;  // epi->scale is already folded into the packed weights
  p.epi_relu = epi_relu ? 1 : 0;
  p.M = n * Ho * Wo;
  return p;
}

struct DensePlan {
  Dims d;
  int Hb[4], Wb[4], ctot[4], cin0[4];
  int ph, pw;
};
bool densenet_plan(int h, int w, DensePlan* pl) {
  pl->d = stem_dims(h, w);
  int H = pl->d.Hp, W = pl->d.Wp, c = 64;
  for (int b = 0; b < 4; ++b) {
    pl->Hb[b] = H;
    pl->Wb[b] = W;
    pl->cin0[b] = c;
    c += kDenseCfg[b] * kGrowth;
    pl->ctot[b] = c;
    if (b < 3) {
      c /= 2;
      H /= 2;
      W /= 2;
    }
  }
  if (H < 7 || W < 7) return false;
  pl->ph = (H - 7) / 7 + 1;
  pl->pw = (W - 7) / 7 + 1;
  return true;
}

// Frames per pass of one dense block (see densenet_forward).  Defaults are set from measurements on B200.
int chunk_frames(int b, int n) {
  static const int dflt[4] = {0, 0, 0, 0};
  static const char* names[4] = {"TN_CHUNK_B1", "TN_CHUNK_B2", "TN_CHUNK_B3", "TN_CHUNK_B4"};
  const char* e = getenv(names[b]);  // read per call: the tuning sweep (tools/chunk_sweep.py) changes it in-process
  int c = e ? atoi(e) : dflt[b];
  return (c <= 0 || c > n) ? n : c;
}

int densenet_forward(tn_backbone* bb, const __nv_bfloat16* in4, int n, int h, int w, float* feats, void* feats_bf16,
                     Bump& ws, bool dry, cudaStream_t st) {
  DensePlan pl;
  if (!densenet_plan(h, w, &pl)) return set_error(TN_ERR_INVALID, "input %dx%d too small for DenseNet-121", h, w);
  const Dims& d = pl.d;
  __nv_bfloat16* stem = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * d.Hs * d.Ws * 64);
  // bottleneck buffer: zero-padded (n, H+2, W+2, 128) layout for the halo 3x3 kernel
  __nv_bfloat16* bott = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * (pl.Hb[0] + 2) * (pl.Wb[0] + 2) * kBott);
  __nv_bfloat16* blk[4];
  for (int b = 0; b < 4; ++b) blk[b] = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * pl.Hb[b] * pl.Wb[b] * pl.ctot[b]);
  // pooled, activated transition input (largest at transition 1)
  __nv_bfloat16* pooled = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * pl.Hb[1] * pl.Wb[1] * pl.ctot[0]);
  if (dry) return TN_OK;

  // stem: conv7x7/2 -> BN -> ReLU (epilogue) ; max-pool 3/2/1 into channels [0,64) of block 1
  {
    // 7x7/2 stem == 4x4/1 conv on the zero-padded space-to-depth image (tn_stem.cu), BN folded, ReLU
    if (d.Ws <= 128 && !getenv("TN_NO_STEM_POOL_FUSION")) {
      // stem conv + BN + ReLU + max-pool in one kernel, written straight into channels [0,64) of block 1
      TN_CUDA(launch_stem_pool(bb->stem_s2d, in4, n, d.Hs + 3, d.Ws + 3, d.Hs, d.Ws, d.Hp, d.Wp, bb->bn0.shift, blk[0], pl.ctot[0],
                               bb->num_sms, st));
    } else {
      TN_CUDA(launch_stem_s2d(bb->stem_s2d, in4, n, d.Hs + 3, d.Ws + 3, d.Hs, d.Ws, bb->bn0.shift, stem, bb->num_sms, st));
      TN_CUDA(launch_maxpool3s2(stem, blk[0], n, d.Hs, d.Ws, 64, d.Hp, d.Wp, pl.ctot[0], 0, st));
    }
  }
  for (int b = 0; b < 4; ++b) {
    const int H = pl.Hb[b], W = pl.Wb[b], ct = pl.ctot[b];
    const bool halo = conv3x3_halo_supported(H, W) && static_cast<long long>(n) * (H + 2) * (W + 2) < (1ll << 31) - 4096;
    // Frame-chunked schedule: the dense layers of a block run chunk by chunk so that the bottleneck tile (and, for small
    // enough chunks, the concat buffer) of a chunk stays resident in the 126 MB L2 between the 1x1 conv that writes it and the
    // 3x3 conv that reads it.  The bottleneck buffer is reused by every chunk (same addresses -> dirty lines are overwritten
    // in L2 and never reach HBM).  TN_CHUNK_B<1..4>=<frames> overrides; 0 = whole batch in one pass.
    const int cs = chunk_frames(b, n);
    if (halo) TN_CUDA(launch_zero_border(bott, cs, H + 2, W + 2, kBott, st));
    // OPT-IN (TN_DENSE_FUSED_MIN_W=<min map width>, e.g. 28 = dense blocks 1-2): ONE fused kernel per dense layer, the bottleneck
    // stays in shared memory (tn_dense_fused.cu).  It removes a third of the step's DRAM bytes (92 -> 62 GB per 2048 frames) but
    // measures slower than the two-kernel schedule: with 14 x 6-pixel tiles the 3x3's tensor-core operand reads (84 useful of 128
    // accumulator rows) saturate the L1TEX data pipe (profiles/r2_fused_dense_layer.md).  The parity test keeps it honest.
    const char* fused_env = getenv("TN_DENSE_FUSED_MIN_W");
    const int fused_min_w = fused_env ? atoi(fused_env) : (1 << 30);
    const bool fused_block = dense_fused_supported(H, W) && W >= fused_min_w;
    // BN1+ReLU as a bf16 clamp (ConvGemmParams::pro_clamp) is opt-in: measured on B200 it is no faster than the fp32
    // scale/shift transform (the 1x1 kernels are bound by the memory system, not by the transformer warps) and its
    // systematic threshold rounding raises the mean feature error by ~16 % (profiles/r1_ab_clamp_prologue.log).
    const char* clamp_env = getenv("TN_CLAMP_PROLOGUE");
    const bool clamp = clamp_env != nullptr && clamp_env[0] == '1';
    for (int f0 = 0; f0 < n; f0 += cs) {
      const int nf = (n - f0 < cs) ? (n - f0) : cs;
      __nv_bfloat16* xb = blk[b] + static_cast<size_t>(f0) * H * W * ct;
      for (const DenseLayer& L : bb->layers[b]) {
        if (fused_block && L.conv1.num_chunks <= 8) {
          TN_CUDA(launch_dense_layer_fused(xb, ct, nf, H, W, L.cin, L.bn1.scale, L.bn1.shift, L.conv1.wpack, L.conv1.num_chunks,
                                           L.bn2.shift, L.conv2h.wpack, bb->num_sms, st));
          continue;
        }
        // BN1+ReLU (prologue) -> 1x1 conv -> BN2+ReLU (epilogue) -> bottleneck
        ConvGemmParams p1 = conv_params(clamp ? L.conv1c : L.conv1, xb, ct, nf, H, W, H, W, 1, 0, clamp ? nullptr : &L.bn1, bott,
                                        kBott, 0, &L.bn2, true);
        if (clamp) {  // BN1+ReLU as an exact bf16 clamp; scale in the weights, threshold term in the shift
          p1.pro_clamp = L.clamp1;
          p1.epi_shift = L.shift1c;
        }
        p1.out_pad = halo ? 1 : 0;
        TN_CUDA(launch_conv_gemm(p1, st));
        // 3x3 conv, 32 new channels written in place at channel offset cin
        if (halo) {
          TN_CUDA(launch_conv3x3_halo(L.conv2h, bott, nf, H, W, xb, ct, L.cin, bb->num_sms, st));
        } else {
          ConvGemmParams p2 = conv_params(L.conv2, bott, kBott, nf, H, W, H, W, 1, 1, nullptr, xb, ct, L.cin, nullptr, false);
          TN_CUDA(launch_conv_gemm(p2, st));
        }
      }
    }
    if (b < 3) {
      // transition: BN+ReLU -> 1x1 conv -> avgpool 2x2, computed as conv1x1(avgpool(relu(bn(x))))
      const int Ho = pl.Hb[b + 1], Wo = pl.Wb[b + 1];
      if ((H % 2) == 0 && (W % 2) == 0 && (ct % 64) == 0 && !getenv("TN_TRANS_GATHER_POOL")) {
        // streaming pre-pass (BN+ReLU+pool, 1/4 of the bytes out) + a plain TMA-fed 1x1 GEMM on the pooled rows
        TN_CUDA(launch_bn_relu_pool2(blk[b], n, H, W, ct, ct, bb->trans[b].bn.scale, bb->trans[b].bn.shift, pooled, st));
        ConvGemmParams p = conv_params(bb->trans[b].conv, pooled, ct, n, Ho, Wo, Ho, Wo, 1, 0, nullptr, blk[b + 1],
                                       pl.ctot[b + 1], 0, nullptr, false);
        p.mode = kModeConv;
        TN_CUDA(launch_conv_gemm(p, st));
      } else {  // in-GEMM gather of the four pooled pixels (odd maps)
        ConvGemmParams p = conv_params(bb->trans[b].conv, blk[b], ct, n, H, W, Ho, Wo, 2, 0, &bb->trans[b].bn, blk[b + 1],
                                       pl.ctot[b + 1], 0, nullptr, false);
        TN_CUDA(launch_conv_gemm(p, st));
      }
    }
  }
  TN_CUDA(launch_tail_pool(blk[3], n, pl.Hb[3], pl.Wb[3], pl.ctot[3], pl.ctot[3], 7, 7, pl.ph, pl.pw, bb->bn_final.scale,
                           bb->bn_final.shift, feats, static_cast<__nv_bfloat16*>(feats_bf16), st));
  return TN_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// fp32-grade ("precise") DenseNet-121 forward: every tensor is a pair of bf16 planes (hi, lo), every contraction three
// tensor-core products (hi*Wh + hi*Wl + lo*Wh) accumulated in fp32 by the SAME tcgen05 GEMM kernel as the speed path, with the
// three terms listed as row taps of the K loop (ConvGemmParams::tma_tap_off; the lo plane lies `plane_rows` rows behind the hi
// plane).  BN/ReLU/pooling run in fp32 on hi+lo (tn_precise.cu).  Reference arithmetic is fp32 (train.py:204); this path meets the
// 1e-3 logit target of BASELINE.json, the bf16 speed path does not.
struct PlanePair {
  __nv_bfloat16* hi = nullptr;
  size_t plane = 0;  // elements from the hi to the lo plane
};
template <typename B>
PlanePair get_planes(B& ws, size_t elems_per_plane, size_t slack = 0) {
  PlanePair pp;
  pp.plane = elems_per_plane;
  pp.hi = ws.template get<__nv_bfloat16>(2 * elems_per_plane + slack);
  return pp;
}

// GEMM over row taps: rows of `a` (pitch `pitch` elements, `cin` channels per tap) x packed split weights -> fp32 out (M, Cout)
cudaError_t gemm_x3(const ConvDev& cv, const PlanePair& a, long long rows_per_plane, int pitch_bytes, int cin, int M,
                    const int* sp_off, int nsp, float* out, const float* shift, bool relu, cudaStream_t st) {
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.in = a.hi;
  p.in_cstride = cv.chunks_per_tap * 64;
  p.H = 1;
  p.W = M;
  p.Ho = 1;
  p.Wo = M;
  p.Cin = cin;
  p.R = 1;
  p.S = 1;
  p.stride = 1;
  p.mode = kModeConv;
  p.wpack = cv.wpack;
  p.num_chunks = cv.num_chunks;
  p.chunks_per_tap = cv.chunks_per_tap;
  p.out = out;
  p.out_cstride = cv.Cout;
  p.out_fp32 = 1;
  p.Cout = cv.Cout;
  p.epi_shift = shift;
  p.epi_relu = relu ? 1 : 0;
  p.M = M;
  p.tma_taps = 3 * nsp;
  p.tma_use_off = 1;
  p.tma_rows = 2 * rows_per_plane;
  p.tma_row_bytes = pitch_bytes;
  if (3 * nsp > 28 || cv.num_chunks != 3 * nsp * cv.chunks_per_tap) return cudaErrorInvalidValue;
  for (int t = 0; t < nsp; ++t) {
    p.tma_tap_off[3 * t + 0] = sp_off[t];                                       // hi * Wh
    p.tma_tap_off[3 * t + 1] = sp_off[t];                                       // hi * Wl
    p.tma_tap_off[3 * t + 2] = sp_off[t] + static_cast<int>(rows_per_plane);    // lo * Wh
  }
  return launch_conv_gemm(p, st);
}

int densenet_forward_precise(tn_backbone* bb, const void* frames, int dtype, int n, int h, int w, float* feats, void* feats_bf16,
                             Bump& ws, bool dry, cudaStream_t st) {
  DensePlan pl;
  if (!densenet_plan(h, w, &pl)) return set_error(TN_ERR_INVALID, "input %dx%d too small for DenseNet-121", h, w);
  const Dims& d = pl.d;
  const int Hz = d.Hs + 3, Wz = d.Ws + 3;
  const size_t zpix = static_cast<size_t>(n) * Hz * Wz;
  if (zpix * 2 >= (1ull << 31) - 65536) return set_error(TN_ERR_INVALID, "precise path: frame chunk too large");
  PlanePair z = get_planes(ws, zpix * 16, 256);             // space-to-depth image, row = pixel (32 B); 4-pixel Toeplitz rows
  float* stem_f = ws.get<float>(zpix * 64);                 // stem output on the padded s2d grid
  PlanePair blk[4];
  for (int b = 0; b < 4; ++b) blk[b] = get_planes(ws, static_cast<size_t>(n) * pl.Hb[b] * pl.Wb[b] * pl.ctot[b]);
  const size_t nr0 = static_cast<size_t>(n) * (pl.Hb[0] + 2) * (pl.Wb[0] + 2);
  PlanePair bott = get_planes(ws, nr0 * kBott, 1024);       // zero-padded bottleneck, rows = padded positions
  size_t act_elems = 0, out_floats = 0;
  for (int b = 0; b < 4; ++b) {
    const size_t npix = static_cast<size_t>(n) * pl.Hb[b] * pl.Wb[b];
    const size_t nr = static_cast<size_t>(n) * (pl.Hb[b] + 2) * (pl.Wb[b] + 2);
    const size_t kmax = static_cast<size_t>((pl.ctot[b] + 63) / 64) * 64;
    act_elems = std::max(act_elems, npix * kmax);
    out_floats = std::max(out_floats, std::max(npix * kBott, nr * kGrowth));
    if (b < 3) out_floats = std::max(out_floats, (npix / 4) * static_cast<size_t>(pl.ctot[b] / 2));
  }
  PlanePair act = get_planes(ws, act_elems, 1024);          // activated operand of the 1x1 convs / transitions
  float* gout = ws.get<float>(out_floats);                  // fp32 GEMM output
  if (dry) return TN_OK;

  // ---- input -> s2d planes; stem conv (4 filter rows x 3 products, K = 4 px x 16 ch) + BN + ReLU -> fp32; max-pool -> block 1
  {
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    float s[3] = {bb->in_scale[0], bb->in_scale[1], bb->in_scale[2]}, bsh[3] = {bb->in_shift[0], bb->in_shift[1], bb->in_shift[2]};
    if (dtype == TN_FRAMES_U8_NHWC) {
      for (int c = 0; c < 3; ++c) {
        const float s1 = 1.f / (255.f * stdv[c]), b1 = -mean[c] / stdv[c];
        s[c] = s1 * bb->in_scale[c];
        bsh[c] = b1 * bb->in_scale[c] + bb->in_shift[c];
      }
    } else if (dtype != TN_FRAMES_F32_NCHW) {
      return set_error(TN_ERR_INVALID, "unknown frames dtype %d", dtype);
    }
    TN_CUDA(cudaMemsetAsync(z.hi + 2 * z.plane, 0, 256 * sizeof(__nv_bfloat16), st));  // slack read by the last Toeplitz rows
    TN_CUDA(launch_s2d_convert_x2(frames, dtype == TN_FRAMES_U8_NHWC, z.hi, z.plane, n, h, w, Hz, Wz, s, bsh, st));
    int off[4];
    for (int a = 0; a < 4; ++a) off[a] = a * Wz;
    TN_CUDA(gemm_x3(bb->stem_x3, z, static_cast<long long>(zpix), 32, 64, static_cast<int>(zpix), off, 4, stem_f, bb->bn0.shift, true,
                    st));
    TN_CUDA(launch_maxpool_f32_split(stem_f, n, Hz, Wz, d.Hs, d.Ws, 64, d.Hp, d.Wp, blk[0].hi, blk[0].plane, pl.ctot[0], st));
  }
  for (int b = 0; b < 4; ++b) {
    const int H = pl.Hb[b], W = pl.Wb[b], ct = pl.ctot[b], Wp = W + 2;
    const size_t npix = static_cast<size_t>(n) * H * W;
    const size_t nr = static_cast<size_t>(n) * (H + 2) * Wp;
    PlanePair bp = bott;
    bp.plane = nr * kBott;  // this block's bottleneck planes: (n, H+2, W+2, 128) each, lo plane right behind the hi plane
    TN_CUDA(launch_zero_border(bp.hi, n, H + 2, Wp, kBott, st));
    TN_CUDA(launch_zero_border(bp.hi + bp.plane, n, H + 2, Wp, kBott, st));
    int off9[9];
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) off9[dy * 3 + dx] = (dy - 1) * Wp + (dx - 1);
    const int zero = 0;
    for (const DenseLayer& L : bb->layers[b]) {
      const int pitch = L.conv1x.chunks_per_tap * 64;
      PlanePair a = act;
      a.plane = npix * pitch;
      // relu(bn1(x)) -> operand planes; 1x1 conv (3 products) + BN2 shift + ReLU -> fp32 -> padded bottleneck planes
      TN_CUDA(launch_bn_relu_split(blk[b].hi, blk[b].plane, npix, L.cin, ct, L.bn1.scale, L.bn1.shift, a.hi, a.plane, pitch, st));
      TN_CUDA(gemm_x3(L.conv1x, a, static_cast<long long>(npix), pitch * 2, L.cin, static_cast<int>(npix), &zero, 1, gout,
                      L.bn2.shift, true, st));
      TN_CUDA(launch_split_store(gout, n, H, W, 0, 0, H, W, kBott, 1, bp.hi, bp.plane, kBott, 0, st));
      // 3x3 conv: 9 row-shift taps x 3 products on the padded-flattened bottleneck -> fp32 on the padded grid -> 32 new channels
      TN_CUDA(gemm_x3(L.conv2x, bp, static_cast<long long>(nr), kBott * 2, kBott, static_cast<int>(nr), off9, 9, gout, nullptr,
                      false, st));
      TN_CUDA(launch_split_store(gout, n, H + 2, Wp, 1, 1, H, W, kGrowth, 0, blk[b].hi, blk[b].plane, ct, L.cin, st));
    }
    if (b < 3) {
      if ((H % 2) != 0 || (W % 2) != 0) return set_error(TN_ERR_INVALID, "precise path needs even feature maps (got %dx%d)", H, W);
      const int Ho = pl.Hb[b + 1], Wo = pl.Wb[b + 1];
      const size_t opix = static_cast<size_t>(n) * Ho * Wo;
      const int pitch = bb->trans[b].convx.chunks_per_tap * 64;
      PlanePair a = act;
      a.plane = opix * pitch;
      TN_CUDA(launch_bn_relu_pool2_x2(blk[b].hi, blk[b].plane, n, H, W, ct, ct, bb->trans[b].bn.scale, bb->trans[b].bn.shift, a.hi,
                                      a.plane, pitch, st));
      TN_CUDA(gemm_x3(bb->trans[b].convx, a, static_cast<long long>(opix), pitch * 2, ct, static_cast<int>(opix), &zero, 1, gout,
                      nullptr, false, st));
      TN_CUDA(launch_split_store(gout, n, Ho, Wo, 0, 0, Ho, Wo, ct / 2, 0, blk[b + 1].hi, blk[b + 1].plane, pl.ctot[b + 1], 0, st));
    }
  }
  TN_CUDA(launch_tail_pool_x2(blk[3].hi, blk[3].plane, n, pl.Hb[3], pl.Wb[3], pl.ctot[3], pl.ctot[3], 7, 7, pl.ph, pl.pw,
                              bb->bn_final.scale, bb->bn_final.shift, feats, static_cast<__nv_bfloat16*>(feats_bf16), st));
  return TN_OK;
}

// frames per pass of the precise path (bounds its ~15 MB/frame workspace: 7.7 GB at 512); measured on 2048 frames: 128 -> 195 ms,
// 256 -> 176 ms, 512 -> 160 ms (fewer, fuller launches); TN_PRECISE_CHUNK overrides (read once)
int precise_chunk() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("TN_PRECISE_CHUNK");
    v = e ? atoi(e) : 512;
    if (v < 1) v = 512;
  }
  return v;
}

int resnet_forward(tn_backbone* bb, const __nv_bfloat16* in4, int n, int h, int w, float* feats, void* feats_bf16,
                   Bump& ws, bool dry, cudaStream_t st) {
  const Dims d = stem_dims(h, w);
  if (d.Hp < 8 || d.Wp < 8) return set_error(TN_ERR_INVALID, "input %dx%d too small for ResNet-18", h, w);
  __nv_bfloat16* stem = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * d.Hs * d.Ws * 64);
  const size_t act = static_cast<size_t>(n) * d.Hp * d.Wp * 64;  // largest post-pool activation
  __nv_bfloat16* xa = ws.get<__nv_bfloat16>(act);
  __nv_bfloat16* xb = ws.get<__nv_bfloat16>(act);
  __nv_bfloat16* tb = ws.get<__nv_bfloat16>(act);
  __nv_bfloat16* rb = ws.get<__nv_bfloat16>(act);
  // stride-1 3x3 convs with 64 / 128 channels (stages 1-2) run the halo kernel on pre-activated, zero-padded tensors
  // (TN_RESNET_NO_HALO=1: everything through the gather GEMM)
  const bool halo_on = !getenv("TN_RESNET_NO_HALO");
  const size_t padded = static_cast<size_t>(n) * (d.Hp + 2) * (d.Wp + 2) * 64;  // stage 1 is the largest padded activation
  __nv_bfloat16* apad = halo_on ? ws.get<__nv_bfloat16>(padded) : nullptr;
  __nv_bfloat16* tpad = halo_on ? ws.get<__nv_bfloat16>(padded) : nullptr;
  if (dry) return TN_OK;
  {
    // 7x7/2 stem == 4x4/1 conv on the zero-padded space-to-depth image (tn_stem.cu), BN folded, ReLU
    if (d.Ws <= 128 && !getenv("TN_NO_STEM_POOL_FUSION")) {
      TN_CUDA(launch_stem_pool(bb->stem_s2d, in4, n, d.Hs + 3, d.Ws + 3, d.Hs, d.Ws, d.Hp, d.Wp, bb->bn0.shift, xa, 64, bb->num_sms, st));
    } else {
      TN_CUDA(launch_stem_s2d(bb->stem_s2d, in4, n, d.Hs + 3, d.Ws + 3, d.Hs, d.Ws, bb->bn0.shift, stem, bb->num_sms, st));
      TN_CUDA(launch_maxpool3s2(stem, xa, n, d.Hs, d.Ws, 64, d.Hp, d.Wp, 64, 0, st));
    }
  }
  int H = d.Hp, W = d.Wp;
  __nv_bfloat16* x = xa;
  __nv_bfloat16* y = xb;
  bool apad_ready = false;  // apad already holds relu(bn1(x)) of the current block (written by the previous block's conv2)
  for (size_t bi = 0; bi < bb->rblocks.size(); ++bi) {
    const ResBlock& B = bb->rblocks[bi];
    const int Ho = (H + 2 - 3) / B.stride + 1, Wo = (W + 2 - 3) / B.stride + 1;
    const bool use_h2 = halo_on && B.h2 && conv3x3_c64_supported(Ho, Wo, B.c);
    const bool use_h1 = use_h2 && B.h1 && conv3x3_c64_supported(H, W, B.c);
    if (use_h2) {
      const __nv_bfloat16* res = x;
      int res_cs = B.cin;
      if (B.has_ds) {  // downsample acts on relu(bn1(x)) (BasicBlockV2)
        ConvGemmParams pd = conv_params(B.ds, x, B.cin, n, H, W, Ho, Wo, B.stride, 0, &B.bn1, rb, B.c, 0, nullptr, false);
        TN_CUDA(launch_conv_gemm(pd, st));
        res = rb;
        res_cs = B.c;
      }
      if (use_h1) {  // conv1: relu(bn2(conv(a))) -> padded, activated
        if (!apad_ready) TN_CUDA(launch_bn_relu_pad(x, B.cin, n, H, W, B.cin, B.bn1.scale, B.bn1.shift, apad, st));
        TN_CUDA(launch_conv3x3_c64(B.c1h, apad, n, H, W, B.bn2.shift, 1, nullptr, 0, nullptr, 0, tpad, nullptr, nullptr, bb->num_sms, st));
      } else {       // strided conv1 through the gather GEMM, written into the padded layout (borders zeroed first)
        TN_CUDA(launch_zero_border(tpad, n, Ho + 2, Wo + 2, B.c, st));
        ConvGemmParams p1 = conv_params(B.conv1, x, B.cin, n, H, W, Ho, Wo, B.stride, 1, &B.bn1, tpad, B.c, 0, &B.bn2, true);
        p1.out_pad = 1;
        TN_CUDA(launch_conv_gemm(p1, st));
      }
      // conv2: + residual -> raw block output (+ the next block's activated, padded input when that block starts with a halo conv)
      const ResBlock* nx = bi + 1 < bb->rblocks.size() ? &bb->rblocks[bi + 1] : nullptr;
      const bool next_h1 = nx && nx->h1 && nx->cin == B.c && conv3x3_c64_supported(Ho, Wo, nx->c);
      TN_CUDA(launch_conv3x3_c64(B.c2h, tpad, n, Ho, Wo, nullptr, 0, res, res_cs, y, B.c, next_h1 ? apad : nullptr,
                                 next_h1 ? nx->bn1.scale : nullptr, next_h1 ? nx->bn1.shift : nullptr, bb->num_sms, st));
      apad_ready = next_h1;
      __nv_bfloat16* t = x;
      x = y;
      y = t;
      H = Ho;
      W = Wo;
      continue;
    }
    apad_ready = false;
    const __nv_bfloat16* res = x;
    int res_cs = B.cin;
    if (B.has_ds) {  // downsample acts on relu(bn1(x)) (BasicBlockV2)
      ConvGemmParams pd = conv_params(B.ds, x, B.cin, n, H, W, Ho, Wo, B.stride, 0, &B.bn1, rb, B.c, 0, nullptr, false);
      TN_CUDA(launch_conv_gemm(pd, st));
      res = rb;
      res_cs = B.c;
    }
    ConvGemmParams p1 = conv_params(B.conv1, x, B.cin, n, H, W, Ho, Wo, B.stride, 1, &B.bn1, tb, B.c, 0, &B.bn2, true);
    TN_CUDA(launch_conv_gemm(p1, st));
    ConvGemmParams p2 = conv_params(B.conv2, tb, B.c, n, Ho, Wo, Ho, Wo, 1, 1, nullptr, y, B.c, 0, nullptr, false);
    p2.res = res;
    p2.res_cstride = res_cs;
    TN_CUDA(launch_conv_gemm(p2, st));
    __nv_bfloat16* t = x;
    x = y;
    y = t;
    H = Ho;
    W = Wo;
  }
  // BN -> ReLU -> GlobalAvgPool -> Flatten
  TN_CUDA(launch_tail_pool(x, n, H, W, 512, 512, H, W, 1, 1, bb->bn_final.scale, bb->bn_final.shift, feats,
                           static_cast<__nv_bfloat16*>(feats_bf16), st));
  return TN_OK;
}

int backbone_run(tn_backbone* bb, const void* frames, int dtype, int n, int h, int w, float* feats, void* feats_bf16,
                 void* workspace, bool dry, size_t* need, cudaStream_t st) {
  if (bb->precise) {
    if (bb->arch != TN_ARCH_DENSENET121) return set_error(TN_ERR_INVALID, "the precise path is implemented for DenseNet-121 only");
    const int D = tn_backbone_feature_dim(bb->arch, h, w);
    if (D <= 0) return set_error(TN_ERR_INVALID, "input %dx%d too small", h, w);
    const size_t frame_bytes = dtype == TN_FRAMES_U8_NHWC ? static_cast<size_t>(h) * w * 3 : static_cast<size_t>(h) * w * 3 * sizeof(float);
    size_t need_max = 0;
    const int chunk = precise_chunk();
    for (int f0 = 0; f0 < n || f0 == 0; f0 += chunk) {
      const int nf = (n - f0 < chunk) ? (n - f0) : chunk;
      Bump wsp{static_cast<uint8_t*>(workspace)};
      int rc = densenet_forward_precise(bb, dry ? nullptr : static_cast<const uint8_t*>(frames) + static_cast<size_t>(f0) * frame_bytes,
                                        dtype, nf, h, w, dry ? nullptr : feats + static_cast<size_t>(f0) * D,
                                        (dry || !feats_bf16) ? nullptr : static_cast<__nv_bfloat16*>(feats_bf16) + static_cast<size_t>(f0) * D,
                                        wsp, dry, st);
      if (rc != TN_OK) return rc;
      need_max = std::max(need_max, align_up(wsp.off, 1024));
      if (dry) break;  // the first chunk is the largest
    }
    if (need) *need = need_max;
    return TN_OK;
  }
  Bump ws{static_cast<uint8_t*>(workspace)};
  const Dims sd = stem_dims(h, w);
  const int Hz = sd.Hs + 3, Wz = sd.Ws + 3;  // zero-padded space-to-depth image, 16 ch per pixel
  __nv_bfloat16* in4 = ws.get<__nv_bfloat16>(static_cast<size_t>(n) * Hz * Wz * 16);
  if (!dry) {
    if (dtype == TN_FRAMES_F32_NCHW) {
      TN_CUDA(launch_s2d_convert(frames, 0, in4, n, h, w, Hz, Wz, bb->in_scale, bb->in_shift, st));
    } else if (dtype == TN_FRAMES_U8_NHWC) {
      // ToTensor (/255) + Normalize(mean,std) (train.py:142-147), then the arch's input affine
      const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
      float s[3], b[3];
      for (int c = 0; c < 3; ++c) {
        const float s1 = 1.f / (255.f * stdv[c]), b1 = -mean[c] / stdv[c];
        s[c] = s1 * bb->in_scale[c];
        b[c] = b1 * bb->in_scale[c] + bb->in_shift[c];
      }
      TN_CUDA(launch_s2d_convert(frames, 1, in4, n, h, w, Hz, Wz, s, b, st));
    } else {
      return set_error(TN_ERR_INVALID, "unknown frames dtype %d", dtype);
    }
  }
  int rc = (bb->arch == TN_ARCH_DENSENET121) ? densenet_forward(bb, in4, n, h, w, feats, feats_bf16, ws, dry, st)
                                             : resnet_forward(bb, in4, n, h, w, feats, feats_bf16, ws, dry, st);
  if (need) *need = align_up(ws.off, 1024);
  return rc;
}

}  // namespace

extern "C" {

size_t tn_backbone_param_count(int arch) {
  if (arch == TN_ARCH_DENSENET121) return densenet_param_count();
  if (arch == TN_ARCH_RESNET18_V2) return resnet18_param_count();
  return 0;
}

int tn_backbone_feature_dim(int arch, int h, int w) {
  if (arch == TN_ARCH_RESNET18_V2) return 512;
  if (arch == TN_ARCH_DENSENET121) {
    DensePlan pl;
    if (!densenet_plan(h, w, &pl)) return -1;
    return 1024 * pl.ph * pl.pw;
  }
  return -1;
}

int tn_backbone_create(tn_backbone_t** out, int arch, int device, const float* params, size_t n_params) {
  if (!out || !params) return tn::set_error(TN_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = tn::check_arch(device);
  if (rc != TN_OK) return rc;
  if (arch != TN_ARCH_DENSENET121 && arch != TN_ARCH_RESNET18_V2) return tn::set_error(TN_ERR_INVALID, "unknown arch %d", arch);
  if (n_params != tn_backbone_param_count(arch))
    return tn::set_error(TN_ERR_INVALID, "expected %zu parameters for arch %d, got %zu", tn_backbone_param_count(arch), arch,
                         n_params);
  TN_CUDA(cudaSetDevice(device));
  std::unique_ptr<tn_backbone> bb(new tn_backbone);
  bb->arch = arch;
  bb->device = device;
  TN_CUDA(cudaDeviceGetAttribute(&bb->num_sms, cudaDevAttrMultiProcessorCount, device));
  Cursor cur{params, n_params};
  bool ok = true;
  if (arch == TN_ARCH_DENSENET121) {
    ok = ok && take_stem_bn(cur, bb->arena, &bb->stem, &bb->bn0, &bb->stem_s2d, &bb->stem_x3);
    int c = 64;
    for (int b = 0; b < 4 && ok; ++b) {
      for (int l = 0; l < kDenseCfg[b] && ok; ++l) {
        DenseLayer L;
        L.cin = c;
        std::vector<float> s1, b1, s2, b2;
        ok = ok && take_bn(cur, bb->arena, c, &L.bn1, &s1, &b1);
        const float* w1 = cur.p;
        ok = ok && take_conv_bn(cur, bb->arena, kBott, c, 1, 1, tn::kModeConv, &L.conv1, &L.bn2, &s2, &b2, &L.conv1x);
        ok = ok && tn::make_conv1x1_clamp(bb->arena, w1, kBott, c, s1.data(), b1.data(), s2.data(), b2.data(), &L.conv1c,
                                          &L.clamp1, &L.shift1c);
        const float* w2 = cur.p;
        ok = ok && take_conv(cur, bb->arena, kGrowth, kBott, 3, 3, tn::kModeConv, &L.conv2);
        ok = ok && make_conv3x3(bb->arena, w2, &L.conv2h);
        ok = ok && make_conv_x3(bb->arena, w2, kGrowth, kBott, 3, 3, nullptr, &L.conv2x);
        bb->layers[b].push_back(L);
        c += kGrowth;
      }
      if (b < 3 && ok) {
        ok = ok && take_bn(cur, bb->arena, c, &bb->trans[b].bn);
        const float* wt = cur.p;
        ok = ok && take_conv(cur, bb->arena, c / 2, c, 1, 1, tn::kModePool2, &bb->trans[b].conv);
        ok = ok && make_conv_x3(bb->arena, wt, c / 2, c, 1, 1, nullptr, &bb->trans[b].convx);
        c /= 2;
      }
    }
    ok = ok && take_bn(cur, bb->arena, c, &bb->bn_final);
    bb->feat_channels = c;
  } else {
    tn::BnDev bnd;
    std::vector<float> hs, hb;
    ok = ok && take_bn(cur, bb->arena, 3, &bnd, &hs, &hb);
    if (ok) {
      for (int i = 0; i < 3; ++i) {
        bb->in_scale[i] = hs[i];
        bb->in_shift[i] = hb[i];
      }
    }
    ok = ok && take_stem_bn(cur, bb->arena, &bb->stem, &bb->bn0, &bb->stem_s2d);
    int cin = 64;
    const int ch[4] = {64, 128, 256, 512};
    for (int s = 0; s < 4 && ok; ++s) {
      for (int b = 0; b < 2 && ok; ++b) {
        ResBlock B;
        B.cin = cin;
        B.c = ch[s];
        B.stride = (b == 0 && s > 0) ? 2 : 1;
        B.has_ds = (b == 0 && cin != B.c);
        ok = ok && take_bn(cur, bb->arena, cin, &B.bn1);
        const float* w1 = cur.p;
        std::vector<float> hs2;
        ok = ok && take_conv_bn(cur, bb->arena, B.c, cin, 3, 3, tn::kModeConv, &B.conv1, &B.bn2, &hs2);
        const float* w2 = cur.p;
        ok = ok && take_conv(cur, bb->arena, B.c, B.c, 3, 3, tn::kModeConv, &B.conv2);
        B.h2 = ok && (B.c == 64 || B.c == 128);
        B.h1 = B.h2 && cin == B.c && B.stride == 1;
        if (B.h1) ok = ok && tn::make_conv3x3_c64(bb->arena, w1, B.c, hs2.data(), &B.c1h);
        if (B.h2) ok = ok && tn::make_conv3x3_c64(bb->arena, w2, B.c, nullptr, &B.c2h);
        if (B.has_ds) ok = ok && take_conv(cur, bb->arena, B.c, cin, 1, 1, tn::kModeConv, &B.ds);
        bb->rblocks.push_back(B);
        cin = B.c;
      }
    }
    ok = ok && take_bn(cur, bb->arena, 512, &bb->bn_final);
    bb->feat_channels = 512;
  }
  if (!ok) {
    if (tn::last_error().empty()) tn::set_error(TN_ERR_INVALID, "parameter blob exhausted");
    return TN_ERR_CUDA;
  }
  *out = bb.release();
  return TN_OK;
}

void tn_backbone_destroy(tn_backbone_t* bb) { delete bb; }

int tn_backbone_set_precision(tn_backbone_t* bb, int mode) {
  if (!bb || (mode != TN_PRECISION_BF16 && mode != TN_PRECISION_SPLIT_BF16)) return tn::set_error(TN_ERR_INVALID, "bad precision mode");
  if (mode == TN_PRECISION_SPLIT_BF16 && bb->arch != TN_ARCH_DENSENET121)
    return tn::set_error(TN_ERR_INVALID, "TN_PRECISION_SPLIT_BF16 is implemented for DenseNet-121 only");
  bb->precise = mode == TN_PRECISION_SPLIT_BF16 ? 1 : 0;
  return TN_OK;
}

size_t tn_backbone_workspace_bytes(const tn_backbone_t* bb, int n_frames, int h, int w) {
  if (!bb || n_frames < 0) return 0;
  size_t need = 0;
  int rc = backbone_run(const_cast<tn_backbone*>(bb), nullptr, 0, n_frames, h, w, nullptr, nullptr, nullptr, true, &need, 0);
  return rc == TN_OK ? need : 0;
}

int tn_backbone_forward(tn_backbone_t* bb, const void* frames, int frames_dtype, int n_frames, int h, int w,
                        float* feats, void* feats_bf16, void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!bb || n_frames < 0) return tn::set_error(TN_ERR_INVALID, "bad backbone handle / frame count");
  if (n_frames == 0) return TN_OK;
  if (!frames || !feats || !workspace) return tn::set_error(TN_ERR_INVALID, "null device pointer");
  const size_t need = tn_backbone_workspace_bytes(bb, n_frames, h, w);
  if (need == 0) return TN_ERR_INVALID;
  if (workspace_bytes < need) return tn::set_error(TN_ERR_WORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, need);
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return tn::set_error(TN_ERR_INVALID, "workspace must be 1024-byte aligned");
  return backbone_run(bb, frames, frames_dtype, n_frames, h, w, feats, feats_bf16, workspace, false, nullptr,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
