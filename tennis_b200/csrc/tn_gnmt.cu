// GNMT decoder: one-step kernels, teacher-forced decode_seq and the whole beam search on the device
// (SURVEY.md §8a G3-G8, Appendix A.5-A.7).
//
// Per decode step the reference launches ~40 tiny library ops and re-gathers every state tensor (including the
// beam-tiled encoder memory) through `take`; here a step is 2L+3 kernels and nothing is gathered: each kernel reads its
// recurrent inputs THROUGH the beam back-pointer of the previous step (`src_row`), and the encoder memory is indexed by
// row / beam so it exists once per source sentence.
//   cell_kernel      h' = cell([xa ; xb], h)  for a 16-row x 16-unit tile, weights pre-transposed for coalesced reads
//   attn_kernel      q = W_q h / sqrt(H); masked softmax(q . mem_t) ; ctx = sum_t w_t mem_t      (scaled Luong, A.5)
//   proj_kernel      logits = W_p out + b  [-> log_softmax]
//   beam_kernel      length-penalised scores, top-`beam` with lowest-index tie-break over beam*V+beam candidates,
//                    back-pointers, sample history, alive / valid_length bookkeeping (A.7) -- one block per sentence
#include <math.h>
#include <string.h>

#include <memory>

#include <stdlib.h>

#include "tn_common.h"

namespace {

using namespace tn;

constexpr float kNeg = -1e18f;
constexpr int kMaxBeam = 16;
constexpr int kMaxLayers = 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------------
// One recurrent cell step for a tile of 16 rows x 16 hidden units.
//   x = [xa_row ; xb_row] with xa_row = xa[(xa_index ? xa_index[r] : r)], xb_row = xb[(src_row ? src_row[r] : r)]
//   h/c read at row (src_row ? src_row[r] : r)  (beam back-pointer), written at row r.
template <int G>
__global__ void __launch_bounds__(256) cell_kernel(const float* __restrict__ xa, int Da, const int* __restrict__ xa_index,
                                                   const float* __restrict__ xb, int Db, const int* __restrict__ xb_src,
                                                   const float* __restrict__ h_in, const float* __restrict__ c_in,
                                                   const int* __restrict__ src_row, const float* __restrict__ WihT,
                                                   const float* __restrict__ WhhT, const float* __restrict__ bih,
                                                   const float* __restrict__ bhh, float* __restrict__ h_out,
                                                   float* __restrict__ c_out, float* __restrict__ out_plus_res,
                                                   int R, int H) {
  extern __shared__ float sm[];
  const int K = Da + Db;
  float* sx = sm;            // [16][K]
  float* sh = sm + 16 * K;   // [16][H]
  const int r0 = blockIdx.x * 16, u0 = blockIdx.y * 16;
  const int tid = threadIdx.x;
  for (int i = tid; i < 16 * K; i += 256) {
    const int rr = i / K, k = i - rr * K;
    const int r = r0 + rr;
    float v = 0.f;
    if (r < R) {
      if (k < Da) {
        const int ra = xa_index ? xa_index[r] : r;
        v = xa[static_cast<size_t>(ra) * Da + k];
      } else {
        const int rb = xb_src ? xb_src[r] : r;
        v = xb[static_cast<size_t>(rb) * Db + (k - Da)];
      }
    }
    sx[i] = v;
  }
  for (int i = tid; i < 16 * H; i += 256) {
    const int rr = i / H, k = i - rr * H;
    const int r = r0 + rr;
    float v = 0.f;
    if (r < R) v = h_in[static_cast<size_t>(src_row ? src_row[r] : r) * H + k];
    sh[i] = v;
  }
  __syncthreads();
  const int rr = tid >> 4, uu = tid & 15;
  const int r = r0 + rr, u = u0 + uu;
  if (r >= R || u >= H) return;
  const int GH = G * H;
  float ai[G], ah[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ai[g] = bih[g * H + u];
    ah[g] = bhh[g * H + u];
  }
  const float* x = sx + rr * K;
  for (int k = 0; k < K; ++k) {
    const float xv = x[k];
    const float* w = WihT + static_cast<size_t>(k) * GH + u;
#pragma unroll
    for (int g = 0; g < G; ++g) ai[g] = fmaf(xv, __ldg(w + g * H), ai[g]);
  }
  const float* hp = sh + rr * H;
  for (int k = 0; k < H; ++k) {
    const float hv = hp[k];
    const float* w = WhhT + static_cast<size_t>(k) * GH + u;
#pragma unroll
    for (int g = 0; g < G; ++g) ah[g] = fmaf(hv, __ldg(w + g * H), ah[g]);
  }
  const float hprev = hp[u];
  float hnew;
  if (G == 3) {  // GRU [r,z,n]
    const float rg = sigmoidf_(ai[0] + ah[0]);
    const float zg = sigmoidf_(ai[1] + ah[1]);
    const float ng = tanhf(ai[2] + rg * ah[2]);
    hnew = (1.f - zg) * ng + zg * hprev;
  } else {  // LSTM [i,f,g,o]
    const float cprev = c_in[static_cast<size_t>(src_row ? src_row[r] : r) * H + u];
    const float ig = sigmoidf_(ai[0] + ah[0]);
    const float fg = sigmoidf_(ai[1] + ah[1]);
    const float gg = tanhf(ai[2] + ah[2]);
    const float og = sigmoidf_(ai[G - 1] + ah[G - 1]);
    const float cn = fg * cprev + ig * gg;
    c_out[static_cast<size_t>(r) * H + u] = cn;
    hnew = og * tanhf(cn);
  }
  h_out[static_cast<size_t>(r) * H + u] = hnew;
  // layer output handed to the next layer / projection; with use_residual: out = h' + layer input (gnmt.py:395-396)
  if (out_plus_res) out_plus_res[static_cast<size_t>(r) * H + u] = hnew + x[u];
}

// ---------------------------------------------------------------------------------------------------------------------
// Same cell step, weights staged in shared memory (default).  The kernel above walks K + H = 356..384 dependent global loads
// per thread (one weight row per iteration): ncu shows 82 warps stalled on long_scoreboard per issue and 2.3 % issue utilisation,
// ~400 us per launch -- 82 % of a decode step (profiles/r2_gnmt_decode.md).  Here a CTA owns 32 rows x 8 hidden units: it loads
// its (K + H) x 8 x G weight slice ONCE (48 KB, coalesced 32-byte runs) and the 32 input/state rows, then every thread reads
// weights as one LDS.128 per k (all gates of its unit) and its row as one LDS.128 per four k.
template <int G>
__global__ void __launch_bounds__(256) cell_kernel_v2(const float* __restrict__ xa, int Da, const int* __restrict__ xa_index,
                                                      const float* __restrict__ xb, int Db, const int* __restrict__ xb_src,
                                                      const float* __restrict__ h_in, const float* __restrict__ c_in,
                                                      const int* __restrict__ src_row, const float* __restrict__ WihT,
                                                      const float* __restrict__ WhhT, const float* __restrict__ bih,
                                                      const float* __restrict__ bhh, float* __restrict__ h_out,
                                                      float* __restrict__ c_out, float* __restrict__ out_plus_res,
                                                      int R, int H) {
  extern __shared__ __align__(16) float sm[];
  const int K = Da + Db, KT = K + H, KTp = KT + 4;  // +4: the four rows of a warp start in different bank groups
  float* sW = sm;              // [KT][8 units][4] (gate slot 3 unused for the GRU)
  float* sx = sm + KT * 32;    // [32][KTp]: row = [xa_row ; xb_row ; h_row]
  const int r0 = blockIdx.x * 32, u0 = blockIdx.y * 8;
  const int tid = threadIdx.x;
  const int GH = G * H;
  // staging: 16-byte loads, six in flight per thread (a one-load-per-iteration loop exposes the full L2 latency 48 times per
  // thread and was 90 % of this kernel's time); requires Da, Db, H multiples of 4 (checked by the launcher)
  constexpr int U = 6;
  {
    const int n4 = KT * G * 2;  // (k, gate, half of the 8 units) -> one float4
    for (int base = tid; base < n4; base += 256 * U) {
      float4 v[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = base + q * 256;
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n4) {
          const int hf = i & 1, g = (i >> 1) % G, k = i / (2 * G);
          const float* W = k < K ? WihT + static_cast<size_t>(k) * GH : WhhT + static_cast<size_t>(k - K) * GH;
          v[q] = __ldg(reinterpret_cast<const float4*>(W + g * H + u0 + 4 * hf));
        }
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = base + q * 256;
        if (i < n4) {
          const int hf = i & 1, g = (i >> 1) % G, k = i / (2 * G);
          float* d = sW + (k * 8 + 4 * hf) * 4 + g;
          d[0] = v[q].x;
          d[4] = v[q].y;
          d[8] = v[q].z;
          d[12] = v[q].w;
        }
      }
    }
    const int r4 = KT / 4, m4 = 32 * r4;  // rows as float4: [xa | xb | h]
    for (int base = tid; base < m4; base += 256 * U) {
      float4 v[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = base + q * 256;
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < m4) {
          const int rr = i / r4, k = (i - rr * r4) * 4;
          const int r = r0 + rr;
          if (r < R) {
            const float* src;
            if (k < Da) src = xa + static_cast<size_t>(xa_index ? xa_index[r] : r) * Da + k;
            else if (k < K) src = xb + static_cast<size_t>(xb_src ? xb_src[r] : r) * Db + (k - Da);
            else src = h_in + static_cast<size_t>(src_row ? src_row[r] : r) * H + (k - K);
            v[q] = __ldg(reinterpret_cast<const float4*>(src));
          }
        }
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int i = base + q * 256;
        if (i < m4) *reinterpret_cast<float4*>(sx + (i / r4) * KTp + (i % r4) * 4) = v[q];
      }
    }
  }
  __syncthreads();
  const int rr = tid >> 3, uu = tid & 7;
  const int r = r0 + rr, u = u0 + uu;
  if (r >= R || u >= H) return;
  float ai[G], ah[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ai[g] = bih[g * H + u];
    ah[g] = bhh[g * H + u];
  }
  const float* x = sx + rr * KTp;
  const float4* w4 = reinterpret_cast<const float4*>(sW) + uu;  // [k * 8]
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    const float4 xv = *reinterpret_cast<const float4*>(x + k);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = w4[(k + kk) * 8];
      const float wg[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int g = 0; g < G; ++g) ai[g] = fmaf(xs[kk], wg[g], ai[g]);
    }
  }
  for (; k < K; ++k) {
    const float4 w = w4[k * 8];
    const float wg[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int g = 0; g < G; ++g) ai[g] = fmaf(x[k], wg[g], ai[g]);
  }
  const float* hp = x + K;
  for (int kh = 0; kh < H; ++kh) {  // (K is not always a multiple of 4: scalar row reads keep this loop alignment-free)
    const float4 w = w4[(K + kh) * 8];
    const float wg[4] = {w.x, w.y, w.z, w.w};
    const float hv = hp[kh];
#pragma unroll
    for (int g = 0; g < G; ++g) ah[g] = fmaf(hv, wg[g], ah[g]);
  }
  const float hprev = hp[u];
  float hnew;
  if (G == 3) {  // GRU [r,z,n]
    const float rg = sigmoidf_(ai[0] + ah[0]);
    const float zg = sigmoidf_(ai[1] + ah[1]);
    const float ng = tanhf(ai[2] + rg * ah[2]);
    hnew = (1.f - zg) * ng + zg * hprev;
  } else {  // LSTM [i,f,g,o]
    const float cprev = c_in[static_cast<size_t>(src_row ? src_row[r] : r) * H + u];
    const float ig = sigmoidf_(ai[0] + ah[0]);
    const float fg = sigmoidf_(ai[1] + ah[1]);
    const float gg = tanhf(ai[2] + ah[2]);
    const float og = sigmoidf_(ai[G - 1] + ah[G - 1]);
    const float cn = fg * cprev + ig * gg;
    c_out[static_cast<size_t>(r) * H + u] = cn;
    hnew = og * tanhf(cn);
  }
  h_out[static_cast<size_t>(r) * H + u] = hnew;
  if (out_plus_res) out_plus_res[static_cast<size_t>(r) * H + u] = hnew + x[u];
}

// ---------------------------------------------------------------------------------------------------------------------
// Scaled-Luong attention for one decoder row per block (128 threads).  mem: (B,T,H); row r attends to sentence r / beam.
__global__ void __launch_bounds__(128) attn_kernel(const float* __restrict__ query, const float* __restrict__ WqT,
                                                   const float* __restrict__ mem, const int* __restrict__ src_len,
                                                   int beam, int T, int H, float* __restrict__ ctx, int use_mask) {
  extern __shared__ float sm[];
  float* sq = sm;          // [H] projected, scaled query
  float* sx = sm + H;      // [H] raw query
  float* ss = sm + 2 * H;  // [T] scores -> weights
  __shared__ float red[4];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = r / beam;
  const int len = use_mask ? min(max(src_len[b], 0), T) : T;
  for (int k = tid; k < H; k += 128) sx[k] = query[static_cast<size_t>(r) * H + k];
  __syncthreads();
  const float inv = rsqrtf(static_cast<float>(H));
  for (int j = tid; j < H; j += 128) {
    float a = 0.f;
    for (int k = 0; k < H; ++k) a = fmaf(sx[k], __ldg(WqT + static_cast<size_t>(k) * H + j), a);
    sq[j] = a * inv;  // contrib.div_sqrt_dim on the projected query
  }
  __syncthreads();
  const float* mb = mem + static_cast<size_t>(b) * T * H;
  // scores: one warp per source position
  for (int t = warp; t < T; t += 4) {
    float a = 0.f;
    for (int k = lane; k < H; k += 32) a = fmaf(sq[k], __ldg(mb + static_cast<size_t>(t) * H + k), a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ss[t] = (t < len) ? a : kNeg;  // where(mask, score, -1e18)
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < T; t += 128) m = fmaxf(m, ss[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int t = tid; t < T; t += 128) {
    const float e = expf(ss[t] - m);
    ss[t] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = red[0] + red[1] + red[2] + red[3];
  const float invs = 1.f / s;
  // context = sum_t softmax_t * mask_t * mem_t
  for (int j = tid; j < H; j += 128) {
    float a = 0.f;
    for (int t = 0; t < len; ++t) a = fmaf(ss[t] * invs, __ldg(mb + static_cast<size_t>(t) * H + j), a);
    ctx[static_cast<size_t>(r) * H + j] = a;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// logits[r] = W_p out[r] + b ; optional log_softmax ; optional zeroing of out[r] (decode_seq's SequenceMask on outputs)
__global__ void __launch_bounds__(128) proj_kernel(const float* __restrict__ out, const float* __restrict__ WpT,
                                                   const float* __restrict__ bp, float* __restrict__ logits, int H, int V,
                                                   int log_softmax, const int* __restrict__ row_valid, size_t row_stride) {
  extern __shared__ float sm[];
  float* sx = sm;      // [H]
  float* sl = sm + H;  // [V]
  __shared__ float red[4];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool valid = row_valid ? row_valid[r] != 0 : true;
  for (int k = tid; k < H; k += 128) sx[k] = valid ? out[static_cast<size_t>(r) * H + k] : 0.f;
  __syncthreads();
  for (int v = tid; v < V; v += 128) {
    float a = bp[v];
    for (int k = 0; k < H; ++k) a = fmaf(sx[k], __ldg(WpT + static_cast<size_t>(k) * V + v), a);
    sl[v] = a;
  }
  __syncthreads();
  float* dst = logits + static_cast<size_t>(r) * row_stride;
  if (!log_softmax) {
    for (int v = tid; v < V; v += 128) dst[v] = sl[v];
    return;
  }
  float m = -INFINITY;
  for (int v = tid; v < V; v += 128) m = fmaxf(m, sl[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float s = 0.f;
  for (int v = tid; v < V; v += 128) s += expf(sl[v] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  const float lse = m + logf(red[0] + red[1] + red[2] + red[3]);
  for (int v = tid; v < V; v += 128) dst[v] = sl[v] - lse;
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention, one block per SENTENCE (default when the memory fits in shared memory): the `beam` decoder rows of a sentence share
// its encoder memory, so the (len x H) tile is staged in shared memory ONCE per block and serves the scores and the context of all
// rows; the round-1 kernel above ran one block per row and walked the tile twice through dependent global loads (ncu: 89 us per
// launch, long_scoreboard 11 warps per issue).  Same arithmetic: q = (x Wq^T) / sqrt(H), score_t = q . m_t (masked -> -1e18),
// softmax over T, ctx = sum_t (w_t / sum) m_t.
constexpr int kAttnMaxBeam = 8;
__global__ void __launch_bounds__(256) attn_kernel_v2(const float* __restrict__ query, const float* __restrict__ WqT,
                                                      const float* __restrict__ mem, const int* __restrict__ src_len, int beam,
                                                      int T, int H, float* __restrict__ ctx, int use_mask) {
  extern __shared__ __align__(16) float sm[];
  const int Tp = T | 1;                    // odd row pitch: conflict-free along t (scores) and along h (context)
  float* smT = sm;                         // [H][Tp]  encoder memory of this sentence, TRANSPOSED
  float* sx = smT + static_cast<size_t>(H) * Tp;  // [beam][H] raw queries
  float* sqp = sx + beam * H;              // [2][beam][H] partial projections (two halves of k)
  float* sq = sqp + 2 * beam * H;          // [beam][H] projected, scaled queries
  float* ss = sq + beam * H;               // [beam][T] scores -> weights
  float* sinv = ss + beam * T;             // [beam] 1 / sum
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = use_mask ? min(max(src_len[b], 0), T) : T;
  const float* mb = mem + static_cast<size_t>(b) * T * H;
  {  // stage + transpose: 16-byte global loads (4 in flight per thread), scalar conflict-free shared stores
    const int n4 = len * H / 4, H4 = H / 4;
    for (int base = tid; base < n4; base += 256 * 7) {
      float4 v[7];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int i = base + q * 256;
        v[q] = i < n4 ? __ldg(reinterpret_cast<const float4*>(mb) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int i = base + q * 256;
        if (i < n4) {
          const int t = i / H4, k = (i - t * H4) * 4;
          smT[(k + 0) * Tp + t] = v[q].x;
          smT[(k + 1) * Tp + t] = v[q].y;
          smT[(k + 2) * Tp + t] = v[q].z;
          smT[(k + 3) * Tp + t] = v[q].w;
        }
      }
    }
  }
  for (int i = tid; i < beam * H; i += 256) sx[i] = query[static_cast<size_t>(b) * beam * H + i];
  __syncthreads();
  {  // query projection: thread (half of k, column j) for all rows of the sentence; 16 weight loads in flight
    const int kh = tid >> 7, j = tid & 127;
    for (int j0 = j; j0 < H; j0 += 128) {
      float a[kAttnMaxBeam];
#pragma unroll
      for (int i = 0; i < kAttnMaxBeam; ++i) a[i] = 0.f;
      const int k0 = kh * (H / 2), k1 = k0 + H / 2;
      for (int k = k0; k < k1; k += 16) {
        float w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) w[q] = (k + q < k1) ? __ldg(WqT + static_cast<size_t>(k + q) * H + j0) : 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (k + q < k1) {
#pragma unroll
            for (int i = 0; i < kAttnMaxBeam; ++i)
              if (i < beam) a[i] = fmaf(sx[i * H + k + q], w[q], a[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kAttnMaxBeam; ++i)
        if (i < beam) sqp[(kh * beam + i) * H + j0] = a[i];
    }
  }
  __syncthreads();
  const float inv = rsqrtf(static_cast<float>(H));
  for (int i = tid; i < beam * H; i += 256) sq[i] = (sqp[i] + sqp[beam * H + i]) * inv;
  __syncthreads();
  // scores: one thread per source position, every beam row reuses the loaded memory element
  for (int t = tid; t < T; t += 256) {
    float a[kAttnMaxBeam];
#pragma unroll
    for (int i = 0; i < kAttnMaxBeam; ++i) a[i] = 0.f;
    if (t < len) {
      for (int k = 0; k < H; ++k) {
        const float m = smT[k * Tp + t];
#pragma unroll
        for (int i = 0; i < kAttnMaxBeam; ++i)
          if (i < beam) a[i] = fmaf(sq[i * H + k], m, a[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kAttnMaxBeam; ++i)
      if (i < beam) ss[i * T + t] = (t < len) ? a[i] : kNeg;  // where(mask, score, -1e18)
  }
  __syncthreads();
  // softmax: warp i owns row i
  for (int i = warp; i < beam; i += 8) {
    float m = -INFINITY;
    for (int t = lane; t < T; t += 32) m = fmaxf(m, ss[i * T + t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float e = expf(ss[i * T + t] - m);
      ss[i * T + t] = e;
      s += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sinv[i] = 1.f / s;
  }
  __syncthreads();
  // context = sum_t softmax_t * mask_t * mem_t
  for (int idx = tid; idx < beam * H; idx += 256) {
    const int i = idx / H, j = idx - i * H;
    const float invs = sinv[i];
    const float* mj = smT + j * Tp;
    const float* wi = ss + i * T;
    float a = 0.f;
    for (int t = 0; t < len; ++t) a = fmaf(wi[t] * invs, mj[t], a);
    ctx[(static_cast<size_t>(b) * beam + i) * H + j] = a;
  }
}

// logits for a tile of 16 rows per block with the projection matrix staged in shared memory (default when H x V fits): the kernel
// below runs one block per row and walks W^T through H dependent global loads per thread (ncu: 42 us per launch).
__global__ void __launch_bounds__(256) proj_kernel_v2(const float* __restrict__ out, const float* __restrict__ WpT,
                                                      const float* __restrict__ bp, float* __restrict__ logits, int R, int H, int V,
                                                      int Vp, int log_softmax, const int* __restrict__ row_valid, size_t row_stride) {
  extern __shared__ __align__(16) float sm[];
  float* sW = sm;                                   // [H][Vp]
  float* sx = sW + static_cast<size_t>(H) * Vp;     // [16][H]
  float* sl = sx + 16 * H;                          // [16][Vp]
  const int r0 = blockIdx.x * 16, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int base = tid; base < H * Vp; base += 256 * 16) {  // sixteen loads in flight per thread
    float w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int i = base + q * 256;
      const int k = i / Vp, v = i - k * Vp;
      w[q] = (i < H * Vp && v < V) ? __ldg(WpT + static_cast<size_t>(k) * V + v) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (base + q * 256 < H * Vp) sW[base + q * 256] = w[q];
  }
  for (int i = tid; i < 16 * H; i += 256) {
    const int rr = i / H, r = r0 + rr;
    const bool valid = r < R && (row_valid ? row_valid[r] != 0 : true);
    sx[i] = valid ? out[static_cast<size_t>(r) * H + (i - rr * H)] : 0.f;
  }
  __syncthreads();
  for (int v = tid; v < Vp; v += 256) {
    float acc[16];
    const float bias = v < V ? bp[v] : 0.f;
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) acc[rr] = bias;
    for (int k = 0; k < H; k += 4) {
      const float w0 = sW[(k + 0) * Vp + v], w1 = sW[(k + 1) * Vp + v], w2 = sW[(k + 2) * Vp + v], w3 = sW[(k + 3) * Vp + v];
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const float4 xv = *reinterpret_cast<const float4*>(sx + rr * H + k);
        acc[rr] = fmaf(xv.x, w0, acc[rr]);
        acc[rr] = fmaf(xv.y, w1, acc[rr]);
        acc[rr] = fmaf(xv.z, w2, acc[rr]);
        acc[rr] = fmaf(xv.w, w3, acc[rr]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) sl[rr * Vp + v] = acc[rr];
  }
  __syncthreads();
  for (int rr = warp; rr < 16; rr += 8) {
    const int r = r0 + rr;
    if (r >= R) continue;
    float* dst = logits + static_cast<size_t>(r) * row_stride;
    const float* row = sl + rr * Vp;
    if (!log_softmax) {
      for (int v = lane; v < V; v += 32) dst[v] = row[v];
      continue;
    }
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(row[v] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float lse = m + logf(s);
    for (int v = lane; v < V; v += 32) dst[v] = row[v] - lse;
  }
}

// host-side dispatch: staged kernels when their tiles fit in shared memory, the per-row kernels otherwise (TN_GNMT_ATTN_V1 /
// TN_GNMT_PROJ_V1 force the latter for A/B)
void launch_attn(const float* query, const float* WqT, const float* mem, const int* src_len, int beam, int T, int H, float* ctx,
                 int use_mask, int R, cudaStream_t st) {
  static const bool v1 = getenv("TN_GNMT_ATTN_V1") != nullptr;
  const size_t smem2 = (static_cast<size_t>(T | 1) * H + 4 * static_cast<size_t>(beam) * H + static_cast<size_t>(beam) * T + beam + 8) * sizeof(float);
  if (!v1 && beam <= kAttnMaxBeam && (H % 4) == 0 && (H % 2) == 0 && R % beam == 0 && smem2 <= 200 * 1024) {
    static size_t cfg = 0;
    if (smem2 > cfg) {
      cudaFuncSetAttribute(attn_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
      cfg = smem2;
    }
    attn_kernel_v2<<<R / beam, 256, smem2, st>>>(query, WqT, mem, src_len, beam, T, H, ctx, use_mask);
    return;
  }
  attn_kernel<<<R, 128, (2 * H + T) * sizeof(float), st>>>(query, WqT, mem, src_len, beam, T, H, ctx, use_mask);
}

void launch_proj(const float* out, const float* WpT, const float* bp, float* logits, int R, int H, int V, int log_softmax,
                 const int* row_valid, size_t row_stride, cudaStream_t st) {
  static const bool v1 = getenv("TN_GNMT_PROJ_V1") != nullptr;
  const int Vp = (V + 31) / 32 * 32;
  const size_t smem2 = (static_cast<size_t>(H) * Vp + 16 * static_cast<size_t>(H) + 16 * static_cast<size_t>(Vp)) * sizeof(float);
  if (!v1 && (H % 4) == 0 && smem2 <= 200 * 1024) {
    static size_t cfg = 0;
    if (smem2 > cfg) {
      cudaFuncSetAttribute(proj_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
      cfg = smem2;
    }
    proj_kernel_v2<<<(R + 15) / 16, 256, smem2, st>>>(out, WpT, bp, logits, R, H, V, Vp, log_softmax, row_valid, row_stride);
    return;
  }
  proj_kernel<<<R, 128, (H + V) * sizeof(float), st>>>(out, WpT, bp, logits, H, V, log_softmax, row_valid, row_stride);
}

// ---------------------------------------------------------------------------------------------------------------------
// One beam-search update for sentence b = blockIdx.x (gluonnlp _BeamSearchStepUpdate, A.7).
struct BeamState {
  float* scores;      // [B][beam]
  float* alive;       // [B][beam]
  int* vlen;          // [B][beam]
  int* samples_in;    // [B][beam][Lmax]
  int* samples_out;   // [B][beam][Lmax]
  int* src_row;       // [B*beam]  back-pointer (global row) for the next step's state reads
  int* next_ids;      // [B*beam]  relu(word)
  int* alive_total;   // [max_length] number of alive beams after each step (host polls for early exit)
};
__global__ void __launch_bounds__(128) beam_kernel(const float* __restrict__ logp, BeamState st, int beam, int V, int Lmax,
                                                   int step /*1-based*/, int eos, float prev_lp, float lp) {
  extern __shared__ float cand[];  // [beam*V + beam]
  __shared__ float s_val[4];
  __shared__ int s_idx[4];
  __shared__ int chosen[kMaxBeam];
  __shared__ float chosen_val[kMaxBeam];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NC = beam * V + beam;
  for (int i = tid; i < beam * V; i += 128) {
    const int k = i / V;
    const float sc = st.scores[b * beam + k];
    const float al = st.alive[b * beam + k];
    const float c = (logp[static_cast<size_t>(b * beam + k) * V + (i - k * V)] + sc * prev_lp) / lp;
    cand[i] = al * c + (1.f - al) * kNeg;  // same arithmetic as the reference (keeps -1e18 exact for dead beams)
  }
  for (int k = tid; k < beam; k += 128) cand[beam * V + k] = (st.alive[b * beam + k] > 0.f) ? kNeg : st.scores[b * beam + k];
  __syncthreads();
  // top-`beam`: repeated arg-max, ties -> lowest index
  for (int sel = 0; sel < beam; ++sel) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < NC; i += 128) {
      const float v = cand[i];
      if (v > bv || (v == bv && i < bi)) {
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_val[warp] = bv;
      s_idx[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 4; ++w)
        if (s_val[w] > s_val[0] || (s_val[w] == s_val[0] && s_idx[w] < s_idx[0])) {
          s_val[0] = s_val[w];
          s_idx[0] = s_idx[w];
        }
      chosen[sel] = s_idx[0];
      chosen_val[sel] = s_val[0];
      cand[s_idx[0]] = -INFINITY;  // remove from further selection
    }
    __syncthreads();
  }
  // bookkeeping (all reads of the old state happen before any write: stage in registers/shared first)
  __shared__ float n_scores[kMaxBeam], n_alive[kMaxBeam];
  __shared__ int n_vlen[kMaxBeam], n_src[kMaxBeam], n_word[kMaxBeam];
  if (tid < beam) {
    const int idx = chosen[tid];
    const bool use_prev = idx >= beam * V;
    const int word = use_prev ? -1 : idx % V;
    const int src = use_prev ? idx - beam * V : idx / V;
    n_word[tid] = word;
    n_src[tid] = src;
    n_scores[tid] = chosen_val[tid];
    n_vlen[tid] = st.vlen[b * beam + src] + 1 - (use_prev ? 1 : 0);
    n_alive[tid] = st.alive[b * beam + src] * ((word != eos) ? 1.f : 0.f);
  }
  __syncthreads();
  // sample history: out[k][0..step-1] = in[src_k][0..step-1]; out[k][step] = word
  for (int i = tid; i < beam * (step + 1); i += 128) {
    const int k = i / (step + 1), j = i - k * (step + 1);
    const int v = (j < step) ? st.samples_in[static_cast<size_t>(b * beam + n_src[k]) * Lmax + j] : n_word[k];
    st.samples_out[static_cast<size_t>(b * beam + k) * Lmax + j] = v;
  }
  if (tid < beam) {
    st.scores[b * beam + tid] = n_scores[tid];
    st.alive[b * beam + tid] = n_alive[tid];
    st.vlen[b * beam + tid] = n_vlen[tid];
    st.src_row[b * beam + tid] = b * beam + n_src[tid];
    st.next_ids[b * beam + tid] = max(n_word[tid], 0);
    if (n_alive[tid] > 0.f) atomicAdd(&st.alive_total[step - 1], 1);
  }
}

// after max_length steps: append EOS to beams still alive, -1 otherwise; valid_length += alive
__global__ void beam_finish_kernel(BeamState st, int rows, int Lmax, int col, int eos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const bool al = st.alive[r] > 0.f;
  st.samples_in[static_cast<size_t>(r) * Lmax + col] = al ? eos : -1;
  if (al) st.vlen[r] += 1;
}

__global__ void beam_init_kernel(BeamState st, int rows, int beam, int Lmax, int bos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  st.scores[r] = (r % beam == 0) ? 0.f : kNeg;
  st.alive[r] = 1.f;
  st.vlen[r] = 1;
  st.samples_in[static_cast<size_t>(r) * Lmax] = bos;
  st.next_ids[r] = bos;
}

// state tiling for beam search: dst[(b*beam+k)] = src[b]
__global__ void tile_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int beam, int H) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(B) * beam * H) return;
  const size_t row = i / H;
  dst[i] = src[(row / beam) * H + (i - row * H)];
}

__global__ void float_ids_to_int_kernel(const float* __restrict__ f, int* __restrict__ ids, int n, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = min(max(static_cast<int>(lrintf(f[i])), 0), V - 1);
}
__global__ void valid_rows_kernel(const int* __restrict__ vl, int* __restrict__ valid, int B, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) valid[i] = t < vl[i] ? 1 : 0;
}

}  // namespace

struct tn_gnmt {
  int device = 0, cell = 0, G = 4, H = 0, E = 0, V = 0, L = 0, use_residual = 0;
  tn::DeviceArena arena;
  const float *WihT[kMaxLayers] = {}, *WhhT[kMaxLayers] = {}, *bih[kMaxLayers] = {}, *bhh[kMaxLayers] = {};
  const float *WqT = nullptr, *embed = nullptr, *WpT = nullptr, *bp = nullptr;
};

namespace {

const float* upload_transposed(tn::DeviceArena& arena, const float* w, int rows, int cols) {
  std::vector<float> t(static_cast<size_t>(rows) * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[static_cast<size_t>(c) * rows + r] = w[static_cast<size_t>(r) * cols + c];
  return static_cast<const float*>(arena.upload(t.data(), t.size() * sizeof(float)));
}

// device-side scratch for one decode step over R rows
struct StepBufs {
  float *h[2][kMaxLayers], *c[2][kMaxLayers];  // ping-pong recurrent state
  float *att[2];                                // attention vector ping-pong
  float *ctx, *lay[kMaxLayers];                 // attention context, per-layer outputs (residual variant)
};

size_t step_scratch_floats(const tn_gnmt* g, int R) {
  return static_cast<size_t>(R) * g->H * (4 * g->L + 2 + 1 + g->L) + 64;
}
void carve_step(const tn_gnmt* g, int R, float* base, StepBufs* sb) {
  const size_t n = static_cast<size_t>(R) * g->H;
  float* p = base;
  for (int s = 0; s < 2; ++s)
    for (int l = 0; l < g->L; ++l) {
      sb->h[s][l] = p; p += n;
      sb->c[s][l] = p; p += n;
    }
  sb->att[0] = p; p += n;
  sb->att[1] = p; p += n;
  sb->ctx = p; p += n;
  for (int l = 0; l < g->L; ++l) { sb->lay[l] = p; p += n; }
}

// One decoder step (gnmt.py:345-404) for R rows: reads state set `cur` through src_row, writes state set 1-cur.
// Returns the device pointer of the last layer's output (R,H).
const float* run_step(const tn_gnmt* g, const StepBufs& sb, int cur, const int* ids, const int* src_row, const float* mem,
                      const int* src_len, int use_mask, int beam, int T, int R, cudaStream_t st, cudaError_t* err,
                      const float* step_emb = nullptr) {
  const int H = g->H, nxt = 1 - cur;
  dim3 grid((R + 15) / 16, (H + 15) / 16);
  static const bool cell_v1 = getenv("TN_GNMT_CELL_V1") != nullptr;  // A/B: the round-1 global-memory cell kernel
  auto launch_cell = [&](int l, const float* xa, int Da, const int* xa_index, const float* xb, int Db, const int* xb_src,
                         float* out_res) {
    ProfScope ps(kProfOther, st);
    if (!cell_v1 && (H % 8) == 0 && (Da % 4) == 0 && (Db % 4) == 0) {
      const int KT = Da + Db + H;
      const size_t smem2 = (static_cast<size_t>(KT) * 32 + 32 * static_cast<size_t>(KT + 4)) * sizeof(float);
      if (smem2 <= 200 * 1024) {
        dim3 grid2((R + 31) / 32, H / 8);
        if (g->G == 3) {
          static size_t cfg3 = 0;
          if (smem2 > cfg3) {
            cudaFuncSetAttribute(cell_kernel_v2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
            cfg3 = smem2;
          }
          cell_kernel_v2<3><<<grid2, 256, smem2, st>>>(xa, Da, xa_index, xb, Db, xb_src, sb.h[cur][l], sb.c[cur][l], src_row,
                                                       g->WihT[l], g->WhhT[l], g->bih[l], g->bhh[l], sb.h[nxt][l], sb.c[nxt][l],
                                                       out_res, R, H);
        } else {
          static size_t cfg4 = 0;
          if (smem2 > cfg4) {
            cudaFuncSetAttribute(cell_kernel_v2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2));
            cfg4 = smem2;
          }
          cell_kernel_v2<4><<<grid2, 256, smem2, st>>>(xa, Da, xa_index, xb, Db, xb_src, sb.h[cur][l], sb.c[cur][l], src_row,
                                                       g->WihT[l], g->WhhT[l], g->bih[l], g->bhh[l], sb.h[nxt][l], sb.c[nxt][l],
                                                       out_res, R, H);
        }
        return;
      }
    }
    const size_t smem = (16 * static_cast<size_t>(Da + Db) + 16 * H) * sizeof(float);
    if (g->G == 3)
      cell_kernel<3><<<grid, 256, smem, st>>>(xa, Da, xa_index, xb, Db, xb_src, sb.h[cur][l], sb.c[cur][l], src_row, g->WihT[l],
                                              g->WhhT[l], g->bih[l], g->bhh[l], sb.h[nxt][l], sb.c[nxt][l], out_res, R, H);
    else
      cell_kernel<4><<<grid, 256, smem, st>>>(xa, Da, xa_index, xb, Db, xb_src, sb.h[cur][l], sb.c[cur][l], src_row, g->WihT[l],
                                              g->WhhT[l], g->bih[l], g->bhh[l], sb.h[nxt][l], sb.c[nxt][l], out_res, R, H);
  };
  // layer 0: [embedding(step_input) ; previous attention vector]
  // (step_emb: rows are the already embedded inputs -- the decoder BLOCK's own call, gnmt.py:306-404)
  if (step_emb) launch_cell(0, step_emb, g->E, nullptr, sb.att[cur], H, src_row, nullptr);
  else launch_cell(0, g->embed, g->E, ids, sb.att[cur], H, src_row, nullptr);
  {
    ProfScope ps(kProfOther, st);
    launch_attn(sb.h[nxt][0], g->WqT, mem, src_len, beam, T, H, sb.att[nxt], use_mask, R, st);
  }
  const float* prev = sb.h[nxt][0];
  for (int l = 1; l < g->L; ++l) {
    float* out_res = g->use_residual ? sb.lay[l] : nullptr;
    launch_cell(l, prev, H, nullptr, sb.att[nxt], H, nullptr, out_res);
    prev = g->use_residual ? sb.lay[l] : sb.h[nxt][l];
  }
  *err = cudaGetLastError();
  return prev;
}

}  // namespace

extern "C" {

int tn_gnmt_create(tn_gnmt_t** out, int device, int cell, int H, int E, int V, int num_layers, int use_residual,
                   const float* const* i2h_weight, const float* const* h2h_weight, const float* const* i2h_bias,
                   const float* const* h2h_bias, const float* query_weight, const float* embed_weight,
                   const float* proj_weight, const float* proj_bias) {
  if (!out) return set_error(TN_ERR_INVALID, "null out");
  *out = nullptr;
  int rc = check_arch(device);
  if (rc != TN_OK) return rc;
  if ((cell != TN_CELL_GRU && cell != TN_CELL_LSTM) || H <= 0 || E <= 0 || V <= 0 || num_layers < 1 || num_layers > kMaxLayers)
    return set_error(TN_ERR_INVALID, "bad GNMT decoder config");
  TN_CUDA(cudaSetDevice(device));
  std::unique_ptr<tn_gnmt> g(new tn_gnmt);
  g->device = device;
  g->cell = cell;
  g->G = cell == TN_CELL_GRU ? 3 : 4;
  g->H = H;
  g->E = E;
  g->V = V;
  g->L = num_layers;
  g->use_residual = use_residual;
  const int GH = g->G * H;
  for (int l = 0; l < num_layers; ++l) {
    const int in = l == 0 ? E + H : 2 * H;
    if (!i2h_weight[l] || !h2h_weight[l] || !i2h_bias[l] || !h2h_bias[l]) return set_error(TN_ERR_INVALID, "null weight");
    g->WihT[l] = upload_transposed(g->arena, i2h_weight[l], GH, in);
    g->WhhT[l] = upload_transposed(g->arena, h2h_weight[l], GH, H);
    g->bih[l] = static_cast<const float*>(g->arena.upload(i2h_bias[l], GH * sizeof(float)));
    g->bhh[l] = static_cast<const float*>(g->arena.upload(h2h_bias[l], GH * sizeof(float)));
    if (!g->WihT[l] || !g->WhhT[l] || !g->bih[l] || !g->bhh[l]) return TN_ERR_CUDA;
  }
  g->WqT = upload_transposed(g->arena, query_weight, H, H);
  g->embed = static_cast<const float*>(g->arena.upload(embed_weight, static_cast<size_t>(V) * E * sizeof(float)));
  g->WpT = upload_transposed(g->arena, proj_weight, V, H);
  g->bp = static_cast<const float*>(g->arena.upload(proj_bias, V * sizeof(float)));
  if (!g->WqT || !g->embed || !g->WpT || !g->bp) return TN_ERR_CUDA;
  *out = g.release();
  return TN_OK;
}

void tn_gnmt_destroy(tn_gnmt_t* g) { delete g; }

size_t tn_gnmt_workspace_bytes(const tn_gnmt_t* g, int rows, int max_len) {
  if (!g || rows < 0) return 0;
  const size_t f = step_scratch_floats(g, rows) + static_cast<size_t>(rows) * g->V;
  const size_t ints = static_cast<size_t>(rows) * (6 + 2 * (static_cast<size_t>(max_len) + 2)) + max_len + 64;
  return tn::align_up(f * sizeof(float), 256) + tn::align_up(ints * sizeof(int), 256) + 4096;
}

// NMTModel.decode_step / GNMTDecoder.__call__ (gnmt.py:306-343, utils/translation.py:51-53): one step for R rows.
//   step_ids: device float (R) token ids (the reference feeds float ids) ; h_in/c_in: device (L,R,H) ; att_in (R,H)
//   mem (Bm,T,H) with row r attending to sentence r / rows_per_mem ; src_len device int32 (Bm) or NULL (no mask)
//   outputs: logits (R,V) raw, h_out/c_out (L,R,H), att_out (R,H)
int tn_gnmt_decode_step(tn_gnmt_t* g, const float* step_ids, const float* h_in, const float* c_in, const float* att_in,
                        const float* mem, const int32_t* src_len, int rows_per_mem, int R, int T, float* logits,
                        float* h_out, float* c_out, float* att_out, void* workspace, size_t workspace_bytes,
                        tn_stream_t stream) {
  if (!g || R < 0 || T <= 0 || rows_per_mem <= 0) return set_error(TN_ERR_INVALID, "bad decode_step arguments");
  if (R == 0) return TN_OK;
  if (!step_ids || !h_in || !att_in || !mem || !logits || !h_out || !att_out || !workspace || (g->G == 4 && (!c_in || !c_out)))
    return set_error(TN_ERR_INVALID, "null device pointer");
  if (workspace_bytes < tn_gnmt_workspace_bytes(g, R, 1)) return set_error(TN_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(R) * g->H;
  StepBufs sb;
  carve_step(g, R, static_cast<float*>(workspace), &sb);
  int* ids = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + tn::align_up((step_scratch_floats(g, R) + static_cast<size_t>(R) * g->V) * sizeof(float), 256));
  for (int l = 0; l < g->L; ++l) {
    TN_CUDA(cudaMemcpyAsync(sb.h[0][l], h_in + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g->G == 4) TN_CUDA(cudaMemcpyAsync(sb.c[0][l], c_in + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  TN_CUDA(cudaMemcpyAsync(sb.att[0], att_in, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  float_ids_to_int_kernel<<<(R + 127) / 128, 128, 0, st>>>(step_ids, ids, R, g->V);
  cudaError_t e;
  const float* outp = run_step(g, sb, 0, ids, nullptr, mem, src_len, src_len != nullptr, rows_per_mem, T, R, st, &e);
  TN_CUDA(e);
  {
    tn::ProfScope ps(tn::kProfOther, st);
    launch_proj(outp, g->WpT, g->bp, logits, R, g->H, g->V, 0, nullptr, g->V, st);
  }
  TN_CUDA(cudaGetLastError());
  for (int l = 0; l < g->L; ++l) {
    TN_CUDA(cudaMemcpyAsync(h_out + l * n, sb.h[1][l], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g->G == 4) TN_CUDA(cudaMemcpyAsync(c_out + l * n, sb.c[1][l], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  TN_CUDA(cudaMemcpyAsync(att_out, sb.att[1], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return TN_OK;
}

// GNMTDecoder.__call__(step_input, states) (gnmt.py:306-404): the decoder BLOCK's step -- inputs are already embedded
// (R,E) rows, the output is the last layer's rnn_out (R,H) (no target projection).  States as in tn_gnmt_decode_step.
int tn_gnmt_decoder_step(tn_gnmt_t* g, const float* step_emb, const float* h_in, const float* c_in, const float* att_in,
                         const float* mem, const int32_t* src_len, int rows_per_mem, int R, int T, float* out, float* h_out,
                         float* c_out, float* att_out, void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!g || R < 0 || T <= 0 || rows_per_mem <= 0) return set_error(TN_ERR_INVALID, "bad decoder_step arguments");
  if (R == 0) return TN_OK;
  if (!step_emb || !h_in || !att_in || !mem || !out || !h_out || !att_out || !workspace || (g->G == 4 && (!c_in || !c_out)))
    return set_error(TN_ERR_INVALID, "null device pointer");
  if (workspace_bytes < tn_gnmt_workspace_bytes(g, R, 1)) return set_error(TN_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(R) * g->H;
  StepBufs sb;
  carve_step(g, R, static_cast<float*>(workspace), &sb);
  for (int l = 0; l < g->L; ++l) {
    TN_CUDA(cudaMemcpyAsync(sb.h[0][l], h_in + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g->G == 4) TN_CUDA(cudaMemcpyAsync(sb.c[0][l], c_in + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  TN_CUDA(cudaMemcpyAsync(sb.att[0], att_in, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  cudaError_t e;
  const float* outp = run_step(g, sb, 0, nullptr, nullptr, mem, src_len, src_len != nullptr, rows_per_mem, T, R, st, &e, step_emb);
  TN_CUDA(e);
  TN_CUDA(cudaMemcpyAsync(out, outp, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < g->L; ++l) {
    TN_CUDA(cudaMemcpyAsync(h_out + l * n, sb.h[1][l], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g->G == 4) TN_CUDA(cudaMemcpyAsync(c_out + l * n, sb.c[1][l], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  TN_CUDA(cudaMemcpyAsync(att_out, sb.att[1], n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return TN_OK;
}

// NMTModel.decode_seq (gnmt.py:254-304): teacher forcing over T_tgt steps; logits (B,T_tgt,V); decoder outputs past
// tgt_valid_len are zeroed before the projection (SequenceMask on outputs -> logits == bias there).
int tn_gnmt_decode_seq(tn_gnmt_t* g, const float* tgt_ids /*(B,T_tgt) float*/, const int32_t* tgt_valid_len, const float* h0,
                       const float* c0, const float* mem, const int32_t* src_len, int B, int T_src, int T_tgt, float* logits,
                       void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!g || B < 0 || T_src <= 0 || T_tgt < 0) return set_error(TN_ERR_INVALID, "bad decode_seq arguments");
  if (B == 0 || T_tgt == 0) return TN_OK;
  if (!tgt_ids || !h0 || !mem || !logits || !workspace || (g->G == 4 && !c0)) return set_error(TN_ERR_INVALID, "null device pointer");
  if (workspace_bytes < tn_gnmt_workspace_bytes(g, B, T_tgt)) return set_error(TN_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(B) * g->H;
  StepBufs sb;
  carve_step(g, B, static_cast<float*>(workspace), &sb);
  int* ints = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + tn::align_up((step_scratch_floats(g, B) + static_cast<size_t>(B) * g->V) * sizeof(float), 256));
  int* ids_all = ints;                 // [B*T_tgt]
  int* valid = ints + B * T_tgt;       // [B]
  int* ids_t = valid + B;              // [B]
  for (int l = 0; l < g->L; ++l) {
    TN_CUDA(cudaMemcpyAsync(sb.h[0][l], h0 + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g->G == 4) TN_CUDA(cudaMemcpyAsync(sb.c[0][l], c0 + l * n, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  TN_CUDA(cudaMemsetAsync(sb.att[0], 0, n * sizeof(float), st));
  float_ids_to_int_kernel<<<(B * T_tgt + 127) / 128, 128, 0, st>>>(tgt_ids, ids_all, B * T_tgt, g->V);
  int cur = 0;
  for (int t = 0; t < T_tgt; ++t) {
    // ids of this step: column t of (B,T_tgt)
    TN_CUDA(cudaMemcpy2DAsync(ids_t, sizeof(int), ids_all + t, T_tgt * sizeof(int), sizeof(int), B, cudaMemcpyDeviceToDevice, st));
    cudaError_t e;
    const float* outp = run_step(g, sb, cur, ids_t, nullptr, mem, src_len, src_len != nullptr, 1, T_src, B, st, &e);
    TN_CUDA(e);
    const int* rv = nullptr;
    if (tgt_valid_len) {
      valid_rows_kernel<<<(B + 127) / 128, 128, 0, st>>>(tgt_valid_len, valid, B, t);
      rv = valid;
    }
    {
      tn::ProfScope ps(tn::kProfOther, st);
      launch_proj(outp, g->WpT, g->bp, logits + static_cast<size_t>(t) * g->V, B, g->H, g->V, 0, rv, static_cast<size_t>(T_tgt) * g->V, st);
    }
    TN_CUDA(cudaGetLastError());
    cur = 1 - cur;
  }
  return TN_OK;
}

// BeamSearchTranslator.translate minus the encoder (utils/translation.py:77-82 + gluonnlp BeamSearchSampler, A.7).
//   mem (B,T,H), src_len int32 (B) or NULL, h0/c0 (L,B,H) decoder initial states from the encoder
//   outputs: samples (B,beam,max_len+2) int32 [only the first *out_len columns are meaningful], scores (B,beam),
//            valid_len (B,beam) int32; out_len (host int) = length dimension L of the reference's result.
int tn_gnmt_beam_search(tn_gnmt_t* g, const float* mem, const int32_t* src_len, const float* h0, const float* c0, int B, int T,
                        int beam, int max_len, float alpha, float K, int bos, int eos, int32_t* samples, float* scores,
                        int32_t* valid_len, int* out_len, void* workspace, size_t workspace_bytes, tn_stream_t stream) {
  if (!g || B < 0 || T <= 0 || beam < 1 || beam > kMaxBeam || max_len < 1) return set_error(TN_ERR_INVALID, "bad beam search arguments");
  if (out_len) *out_len = 0;
  if (B == 0) return TN_OK;
  if (!mem || !h0 || !samples || !scores || !valid_len || !out_len || !workspace || (g->G == 4 && !c0))
    return set_error(TN_ERR_INVALID, "null pointer");
  const int R = B * beam, Lmax = max_len + 2;
  if (workspace_bytes < tn_gnmt_workspace_bytes(g, R, max_len)) return set_error(TN_ERR_WORKSPACE, "workspace too small");
  if (static_cast<size_t>(beam) * g->V + beam > 48 * 1024 / sizeof(float) - 64) return set_error(TN_ERR_INVALID, "beam*V too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  StepBufs sb;
  carve_step(g, R, static_cast<float*>(workspace), &sb);
  float* logp = static_cast<float*>(workspace) + step_scratch_floats(g, R);
  int* ints = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + tn::align_up((step_scratch_floats(g, R) + static_cast<size_t>(R) * g->V) * sizeof(float), 256));
  BeamState bs;
  bs.scores = scores;
  bs.vlen = valid_len;
  bs.alive = reinterpret_cast<float*>(ints);  // [R]
  bs.src_row = ints + R;                      // [R]
  bs.next_ids = ints + 2 * R;                 // [R]
  bs.alive_total = ints + 3 * R;              // [max_len]
  int* smp[2] = {ints + 3 * R + max_len, ints + 3 * R + max_len + static_cast<size_t>(R) * Lmax};
  const size_t n = static_cast<size_t>(R) * g->H;
  // tile the encoder states x beam (row = b*beam + k); attention vector starts at zero (gnmt.py:244)
  for (int l = 0; l < g->L; ++l) {
    tile_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(h0 + static_cast<size_t>(l) * B * g->H, sb.h[0][l], B, beam, g->H);
    if (g->G == 4) tile_rows_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(c0 + static_cast<size_t>(l) * B * g->H, sb.c[0][l], B, beam, g->H);
  }
  TN_CUDA(cudaMemsetAsync(sb.att[0], 0, n * sizeof(float), st));
  TN_CUDA(cudaMemsetAsync(bs.alive_total, 0, max_len * sizeof(int), st));
  TN_CUDA(cudaMemsetAsync(smp[0], 0xff, 2 * static_cast<size_t>(R) * Lmax * sizeof(int), st));  // -1 everywhere
  bs.samples_in = smp[0];
  bs.samples_out = smp[1];
  beam_init_kernel<<<(R + 127) / 128, 128, 0, st>>>(bs, R, beam, Lmax, bos);
  TN_CUDA(cudaGetLastError());

  std::vector<int> alive_host(max_len, -1);
  int cur = 0, done_step = -1, polled = 0;
  const int kPoll = 8;
  for (int i = 0; i < max_len; ++i) {
    cudaError_t e;
    const float* outp = run_step(g, sb, cur, bs.next_ids, i == 0 ? nullptr : bs.src_row, mem, src_len, src_len != nullptr, beam, T, R, st, &e);
    TN_CUDA(e);
    {
      tn::ProfScope ps(tn::kProfOther, st);
      launch_proj(outp, g->WpT, g->bp, logp, R, g->H, g->V, 1, nullptr, g->V, st);
    }
    {
      tn::ProfScope ps(tn::kProfOther, st);
      // BeamSearchScorer (A.7): the length penalties are Python floats (double) cast to fp32 scalars
      const int step = i + 1;
      const double den = pow(static_cast<double>(K) + 1.0, static_cast<double>(alpha));
      const float prev_lp = step != 1 ? static_cast<float>(pow(static_cast<double>(K) + step - 1.0, static_cast<double>(alpha)) / den) : 1.f;
      const float lp = static_cast<float>(pow(static_cast<double>(K) + step, static_cast<double>(alpha)) / den);
      beam_kernel<<<B, 128, (static_cast<size_t>(beam) * g->V + beam) * sizeof(float), st>>>(logp, bs, beam, g->V, Lmax, step, eos, prev_lp, lp);
    }
    TN_CUDA(cudaGetLastError());
    int* t = bs.samples_in;
    bs.samples_in = bs.samples_out;
    bs.samples_out = t;
    cur = 1 - cur;
    if ((i + 1) % kPoll == 0 || i + 1 == max_len) {  // "if no beam is alive: stop" -- polled every kPoll steps; extra steps
      TN_CUDA(cudaMemcpyAsync(alive_host.data() + polled, bs.alive_total + polled, (i + 1 - polled) * sizeof(int),   // only append -1
                              cudaMemcpyDeviceToHost, st));
      TN_CUDA(cudaStreamSynchronize(st));
      for (int j = polled; j <= i && done_step < 0; ++j)
        if (alive_host[j] == 0) done_step = j;
      polled = i + 1;
      if (done_step >= 0) break;
    }
  }
  int L;
  if (done_step >= 0) {
    L = done_step + 2;  // BOS + (done_step+1) words
  } else {
    beam_finish_kernel<<<(R + 127) / 128, 128, 0, st>>>(bs, R, Lmax, max_len + 1, eos);
    TN_CUDA(cudaGetLastError());
    L = max_len + 2;
  }
  TN_CUDA(cudaMemcpyAsync(samples, bs.samples_in, static_cast<size_t>(R) * Lmax * sizeof(int), cudaMemcpyDeviceToDevice, st));
  *out_len = L;
  return TN_OK;
}

}  // extern "C"
