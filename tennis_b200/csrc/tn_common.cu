#include "tn_common.h"

#include <stdlib.h>

#include <math.h>
#include <string.h>

namespace tn {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int check_arch(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count)
    return set_error(TN_ERR_ARCH, "no CUDA device %d (%s); tennis_b200 has no CPU fallback", device,
                     e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
  int major = 0;
  TN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) return set_error(TN_ERR_ARCH, "device %d is sm_%d0, kernels are built for sm_100a only", device, major);
  return TN_OK;
}

DeviceArena::~DeviceArena() {
  for (void* p : ptrs) cudaFree(p);
}
void* DeviceArena::alloc(size_t bytes) {
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error(TN_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  ptrs.push_back(d);
  return d;
}
void* DeviceArena::upload(const void* host, size_t bytes) {
  void* d = alloc(bytes);
  if (!d) return nullptr;
  cudaError_t e = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error(TN_ERR_CUDA, "cudaMemcpy H2D(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return d;
}

bool make_bn(DeviceArena& arena, const float* gamma, const float* beta, const float* mean, const float* var, int C,
             BnDev* out, std::vector<float>* host_scale, std::vector<float>* host_shift) {
  std::vector<float> sc(C), sh(C);
  for (int c = 0; c < C; ++c) {
    // fp32 like the reference runtime: (x - mean) / sqrt(var + eps) * gamma + beta
    float inv = 1.0f / sqrtf(var[c] + 1e-5f);
    sc[c] = gamma[c] * inv;
    sh[c] = beta[c] - mean[c] * sc[c];
  }
  out->C = C;
  out->scale = static_cast<const float*>(arena.upload(sc.data(), C * sizeof(float)));
  out->shift = static_cast<const float*>(arena.upload(sh.data(), C * sizeof(float)));
  if (host_scale) *host_scale = sc;
  if (host_shift) *host_shift = sh;
  return out->scale && out->shift;
}

bool make_conv(DeviceArena& arena, const float* w, int Cout, int Cin, int R, int S, int mode, ConvDev* out,
               const float* fold_scale) {
  const int BN = conv_gemm_pick_bn(Cout);
  const int ntiles = (Cout + BN - 1) / BN;
  int cpt, nchunks;
  if (mode == kModeStem) {
    cpt = 1;
    nchunks = (R + 1) / 2;
  } else {
    cpt = (Cin + 63) / 64;
    nchunks = R * S * cpt;
  }
  const size_t bytes = static_cast<size_t>(ntiles) * nchunks * BN * 128;
  std::vector<uint8_t> blob(bytes, 0);
  auto put = [&](int t, int c, int n, int kk, float v) {
    size_t off = (static_cast<size_t>(t) * nchunks + c) * BN * 128 + static_cast<size_t>(n) * 128 +
                 ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
    __nv_bfloat16 b = __float2bfloat16(v);
    memcpy(&blob[off], &b, 2);
  };
  for (int t = 0; t < ntiles; ++t) {
    for (int c = 0; c < nchunks; ++c) {
      for (int n = 0; n < BN; ++n) {
        const int co = t * BN + n;
        if (co >= Cout) continue;
        for (int kk = 0; kk < 64; ++kk) {
          float v = 0.f;
          if (mode == kModeStem) {
            const int r = 2 * c + (kk >> 5), q = kk & 31, s = q >> 2, ch = q & 3;
            if (r < R && s < S && ch < Cin) v = w[((static_cast<size_t>(co) * Cin + ch) * R + r) * S + s];
          } else {
            const int tap = c / cpt, ci = (c % cpt) * 64 + kk;
            const int r = tap / S, s = tap % S;
            if (ci < Cin) v = w[((static_cast<size_t>(co) * Cin + ci) * R + r) * S + s];
          }
          if (fold_scale) v *= fold_scale[co];
          if (v != 0.f) put(t, c, n, kk, v);
        }
      }
    }
  }
  out->Cin = Cin;
  out->Cout = Cout;
  out->R = R;
  out->S = S;
  out->num_chunks = nchunks;
  out->chunks_per_tap = cpt;
  out->mode = mode;
  out->wpack = static_cast<const uint8_t*>(arena.upload(blob.data(), bytes));
  return out->wpack != nullptr;
}

bool make_conv1x1_clamp(DeviceArena& arena, const float* w, int Cout, int Cin, const float* in_scale, const float* in_shift,
                        const float* out_scale, const float* out_shift, ConvDev* out, const uint4** clamp_dev,
                        const float** shift_dev) {
  // relu(s*x + b) = s * (clamp(x, lo, hi) - t) with t = -b/s (see ConvGemmParams::pro_clamp).  t is rounded to bf16 (the
  // comparison against bf16 x is then exact); the weights carry s (and the following BatchNorm's scale) and are rounded to
  // bf16 once; the constant -sum_c Wq[n][c]*t[c] is accumulated in double from the ROUNDED weights so that a channel that is
  // clamped everywhere cancels exactly.  Channels whose scale is zero (or so small that t leaves the bf16 range) are
  // constants relu(b): their weight is dropped and W*relu(b) goes into the shift.
  const uint16_t kPosInf = 0x7F80, kNegInf = 0xFF80;
  std::vector<float> wf(static_cast<size_t>(Cout) * Cin);
  std::vector<float> tq(Cin, 0.f);
  std::vector<uint16_t> clamp(static_cast<size_t>(2) * ((Cin + 7) / 8) * 8, 0);
  std::vector<double> bias(Cout, 0.0);
  for (int n = 0; n < Cout; ++n) bias[n] = out_shift ? out_shift[n] : 0.0;
  for (int c = 0; c < Cin; ++c) {
    const float s = in_scale[c], b = in_shift[c];
    const float t = (s != 0.f) ? -b / s : 0.f;
    const bool constant = (s == 0.f) || !(fabsf(t) < 1e30f);
    uint16_t lo = kNegInf, hi = kPosInf;
    if (!constant) {
      const __nv_bfloat16 tb = __float2bfloat16(t);
      uint16_t bits;
      memcpy(&bits, &tb, 2);
      tq[c] = __bfloat162float(tb);
      if (s > 0.f) lo = bits;
      else hi = bits;
    }
    const size_t slot = static_cast<size_t>(c / 8) * 16 + (c % 8);
    clamp[slot] = lo;
    clamp[slot + 8] = hi;
    for (int n = 0; n < Cout; ++n) {
      const float os = out_scale ? out_scale[n] : 1.f;
      const float wv = w[static_cast<size_t>(n) * Cin + c];
      if (constant) {
        wf[static_cast<size_t>(n) * Cin + c] = 0.f;
        bias[n] += static_cast<double>(wv) * os * fmaxf(b, 0.f);
      } else {
        const float folded = wv * s * os;
        wf[static_cast<size_t>(n) * Cin + c] = folded;
        bias[n] -= static_cast<double>(__bfloat162float(__float2bfloat16(folded))) * tq[c];
      }
    }
  }
  if (!make_conv(arena, wf.data(), Cout, Cin, 1, 1, kModeConv, out, nullptr)) return false;
  std::vector<float> sh(Cout);
  for (int n = 0; n < Cout; ++n) sh[n] = static_cast<float>(bias[n]);
  *clamp_dev = static_cast<const uint4*>(arena.upload(clamp.data(), clamp.size() * sizeof(uint16_t)));
  *shift_dev = static_cast<const float*>(arena.upload(sh.data(), Cout * sizeof(float)));
  return *clamp_dev && *shift_dev;
}

// ---------------------------------------------------------------- launch counters / event timing
namespace {
struct ProfState {
  bool timing = false;
  long long launches[2] = {0, 0};
  std::vector<cudaEvent_t> begin[2], end[2];
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};
ProfState& prof() {
  static ProfState s;
  return s;
}
}  // namespace

void prof_begin(int kind, cudaStream_t st) {
  ProfState& s = prof();
  s.launches[kind]++;
  if (!s.timing) return;
  cudaEvent_t e = s.get();
  cudaEventRecord(e, st);
  s.begin[kind].push_back(e);
}
void prof_end(int kind, cudaStream_t st) {
  ProfState& s = prof();
  if (!s.timing) return;
  cudaEvent_t e = s.get();
  cudaEventRecord(e, st);
  s.end[kind].push_back(e);
}

bool pdl_enabled() {
  static const bool on = getenv("TN_NO_PDL") == nullptr;
  return on;
}

}  // namespace tn

extern "C" {

int tn_profile_enable(int timing_on) {
  tn::prof().timing = timing_on != 0;
  return TN_OK;
}

int tn_profile_read(double* conv_gemm_ms, long long* conv_gemm_launches, double* other_ms, long long* other_launches,
                    int reset) {
  tn::ProfState& s = tn::prof();
  double ms[2] = {0, 0};
  for (int k = 0; k < 2; ++k) {
    const size_t n = s.begin[k].size() < s.end[k].size() ? s.begin[k].size() : s.end[k].size();
    for (size_t i = 0; i < n; ++i) {
      cudaError_t e = cudaEventSynchronize(s.end[k][i]);
      if (e != cudaSuccess) return tn::set_error(TN_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
      float t = 0.f;
      cudaEventElapsedTime(&t, s.begin[k][i], s.end[k][i]);
      ms[k] += t;
    }
  }
  if (conv_gemm_ms) *conv_gemm_ms = ms[0];
  if (other_ms) *other_ms = ms[1];
  if (conv_gemm_launches) *conv_gemm_launches = s.launches[0];
  if (other_launches) *other_launches = s.launches[1];
  if (reset) {
    for (int k = 0; k < 2; ++k) {
      for (cudaEvent_t e : s.begin[k]) s.pool.push_back(e);
      for (cudaEvent_t e : s.end[k]) s.pool.push_back(e);
      s.begin[k].clear();
      s.end[k].clear();
      s.launches[k] = 0;
    }
  }
  return TN_OK;
}


int tn_version(void) { return 100; }
const char* tn_last_error(void) { return tn::last_error().c_str(); }
int tn_device_check(int device) { return tn::check_arch(device); }

}  // extern "C"
