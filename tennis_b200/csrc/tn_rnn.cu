// Recurrent scan for the fused (bi)GRU / (bi)LSTM layer (Appendix A.3/A.4).
//
// The input projection x*W_ih^T + b_ih for all timesteps is one tensor-core GEMM (tn_conv_gemm.cu);
// this kernel runs the serial part.  A thread-block cluster of CL CTAs owns a tile of RB batch rows of
// one direction; CTA q keeps the W_hh^T slice of its H/CL hidden units (all gates) resident in shared
// memory in fp32 for the whole scan, computes their gates each step, and publishes the new h slice to
// every CTA of the cluster through distributed shared memory, followed by one cluster barrier.
// max-over-time (CNNRNN, definitions.py:107) is accumulated in registers so (B,T,2H) need not be
// written at all when only the pooled vector is wanted.
//
// valid_length semantics follow MXNet's unroll(valid_length=...) (A.4): row b runs len[b] steps; the
// reverse direction starts at position len[b]-1; outputs past len[b] stay zero (caller pre-zeroes y);
// the returned state is the one after step len[b]-1.
#include "tn_rnn.h"

#include "tn_common.h"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace tn {

namespace {

constexpr int kRB = 4;  // batch rows per cluster

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Packed fp32x2 FMA (sm_100 FFMA2): d.{x,y} += a.{x,y} * b.{x,y}, each lane a fused rn fma.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d.x), "+f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

constexpr int kHReg = 128;  // hidden size of the register-resident variant

// WREG: the thread's W_hh^T column (kHReg weights) lives in registers for the whole scan, so a step reads
// only the 2 KB hidden state from shared memory (broadcast LDS.128) and issues packed FFMA2s; h is laid
// out [row][k] so that consecutive k pair up.  Otherwise the slice is staged in shared memory ([H][NJ])
// and h is laid out [k][row].
template <int G, int CL, bool WREG>
__global__ void __launch_bounds__(WREG ? G * kHReg / CL : 1024, 1) rnn_scan_kernel(const RnnScanParams p) {
  extern __shared__ float smem_f[];
  const int H = WREG ? kHReg : p.H;
  const int HS = H / CL;   // hidden units owned by this CTA
  const int NJ = G * HS;   // gate columns owned by this CTA == blockDim.x
  float* Wt = smem_f;                       // [H][NJ] (absent when WREG)
  float* hbuf = Wt + (WREG ? 0 : static_cast<size_t>(H) * NJ);  // [2][H][kRB]  (WREG: [2][kRB][H])
  float* hh = hbuf + 2 * H * kRB;           // [kRB][NJ]

  const int tid = threadIdx.x;
  const int q = (CL > 1) ? static_cast<int>(cg::this_cluster().block_rank()) : 0;
  const int tile = blockIdx.x / CL;
  const int dir = blockIdx.y;
  const int b0 = tile * kRB;
  const int GH = G * H;

  // ---- stage the W_hh^T slice: column j=(g,u) <- row (g*H + q*HS + u) of W_hh ; global WhhT is [dir][H(k)][G*H]
  float wreg[WREG ? kHReg : 2];
  if (WREG) {
    const int g = tid / HS, u = tid - g * HS;
    const float* src = p.WhhT + static_cast<size_t>(dir) * H * GH + g * H + q * HS + u;
#pragma unroll
    for (int k = 0; k < kHReg; ++k) wreg[k] = __ldg(src + static_cast<size_t>(k) * GH);
  } else {
    const float* src = p.WhhT + static_cast<size_t>(dir) * H * GH;
    for (int idx = tid; idx < H * NJ; idx += blockDim.x) {
      const int k = idx / NJ;
      const int j = idx - k * NJ;
      const int g = j / HS, u = j - g * HS;
      Wt[idx] = __ldg(src + static_cast<size_t>(k) * GH + g * H + q * HS + u);
    }
  }
  // initial hidden state (zeros unless h0 given): hbuf[0][k][b]
  for (int idx = tid; idx < H * kRB; idx += blockDim.x) {
    const int k = WREG ? idx % H : idx / kRB;
    const int b = WREG ? idx / H : idx - k * kRB;
    float v = 0.f;
    if (p.h0 && b0 + b < p.B) v = p.h0[(static_cast<size_t>(dir) * p.B + b0 + b) * H + k];
    hbuf[idx] = v;
    hbuf[H * kRB + idx] = v;
  }
  const int j_gate = tid / HS;        // gate of my column
  const int j_unit = tid - j_gate * HS;
  const float bhh = __ldg(p.bhh + static_cast<size_t>(dir) * GH + j_gate * H + q * HS + j_unit);

  // ---- phase-2 items: (b,u) pairs, item = tid + n*blockDim  (kRB*HS items, blockDim = G*HS threads)
  constexpr int IPT = (kRB + G - 1) / G;
  int it_b[IPT], it_u[IPT], it_len[IPT];
  float it_c[IPT], it_h[IPT], it_max[IPT];
  int maxlen = 0;
#pragma unroll
  for (int n = 0; n < IPT; ++n) {
    const int item = tid + n * blockDim.x;
    it_b[n] = -1;
    it_u[n] = 0;
    it_len[n] = 0;
    it_c[n] = 0.f;
    it_h[n] = 0.f;
    it_max[n] = -INFINITY;
    if (item < kRB * HS) {
      const int b = item / HS, u = item - b * HS;
      if (b0 + b < p.B) {
        it_b[n] = b;
        it_u[n] = u;
        it_len[n] = p.valid_len ? min(max(p.valid_len[b0 + b], 0), p.T) : p.T;
        if (p.c0) it_c[n] = p.c0[(static_cast<size_t>(dir) * p.B + b0 + b) * H + q * HS + u];
        if (p.h0) it_h[n] = p.h0[(static_cast<size_t>(dir) * p.B + b0 + b) * H + q * HS + u];
      }
    }
  }
  for (int b = 0; b < kRB; ++b) {
    if (b0 + b < p.B) maxlen = max(maxlen, p.valid_len ? min(max(p.valid_len[b0 + b], 0), p.T) : p.T);
  }
  if (CL > 1) cg::this_cluster().sync(); else __syncthreads();

  const int gx_row = p.ndir * GH;  // floats per (b,t) row of gx
  // input-projection gates of step s are loaded ONE STEP AHEAD (during step s-1): the ~1 us L2/HBM latency of these loads was
  // on the critical path of every step when they were issued at the top of the step that consumes them
  float gxv[IPT][G];
  auto load_gx = [&](int s, float (&dst)[IPT][G]) {
#pragma unroll
    for (int n = 0; n < IPT; ++n) {
#pragma unroll
      for (int g = 0; g < G; ++g) dst[n][g] = 0.f;
      if (it_b[n] >= 0 && s < it_len[n]) {
        const int pos = (dir == 0 || p.reverse_dir1 == 0) ? s : it_len[n] - 1 - s;
        const float* gp = p.gx + (static_cast<size_t>(b0 + it_b[n]) * p.T + pos) * gx_row + dir * GH + q * HS + it_u[n];
#pragma unroll
        for (int g = 0; g < G; ++g) dst[n][g] = __ldg(gp + g * H);
      }
    }
  };
  load_gx(0, gxv);
  for (int s = 0; s < maxlen; ++s) {
    const float* hcur = hbuf + (s & 1) * H * kRB;
    float* hnext_local = hbuf + ((s + 1) & 1) * H * kRB;
    float gxn[IPT][G];
    load_gx(s + 1, gxn);  // s + 1 >= it_len -> zeros, no load

    // phase 1: hh[b][j] = sum_k Wt[k][j] * h[k][b] + bhh[j]
    float acc[kRB];
#pragma unroll
    for (int b = 0; b < kRB; ++b) acc[b] = 0.f;
    if (WREG) {
      // Two rows per pass, eight k per block: the four LDS.128 of a block are issued together and feed eight packed FFMA2 on
      // four independent accumulator chains.  (With all four rows live next to the 128 weight registers ptxas had no room to
      // hoist the loads: every FFMA2 waited ~30 cycles on its own LDS through one reused register -- 256 serial load->fma
      // round trips = the whole 3.8 us step, profiles/r2_gru_head.md.)
#pragma unroll
      for (int bp = 0; bp < kRB; bp += 2) {
        float2 a0 = make_float2(0.f, 0.f), a0b = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f), a1b = make_float2(0.f, 0.f);
        const float* h0 = hcur + bp * kHReg;
        const float* h1 = h0 + kHReg;
#pragma unroll
        for (int k = 0; k < kHReg; k += 8) {
          const float4 p0 = *reinterpret_cast<const float4*>(h0 + k);
          const float4 p1 = *reinterpret_cast<const float4*>(h0 + k + 4);
          const float4 q0 = *reinterpret_cast<const float4*>(h1 + k);
          const float4 q1 = *reinterpret_cast<const float4*>(h1 + k + 4);
          const float2 w01 = make_float2(wreg[k], wreg[k + 1]), w23 = make_float2(wreg[k + 2], wreg[k + 3]);
          const float2 w45 = make_float2(wreg[k + 4], wreg[k + 5]), w67 = make_float2(wreg[k + 6], wreg[k + 7]);
          ffma2(a0, w01, make_float2(p0.x, p0.y));
          ffma2(a1, w01, make_float2(q0.x, q0.y));
          ffma2(a0b, w23, make_float2(p0.z, p0.w));
          ffma2(a1b, w23, make_float2(q0.z, q0.w));
          ffma2(a0, w45, make_float2(p1.x, p1.y));
          ffma2(a1, w45, make_float2(q1.x, q1.y));
          ffma2(a0b, w67, make_float2(p1.z, p1.w));
          ffma2(a1b, w67, make_float2(q1.z, q1.w));
        }
        acc[bp] = (a0.x + a0.y) + (a0b.x + a0b.y);
        acc[bp + 1] = (a1.x + a1.y) + (a1b.x + a1b.y);
      }
    } else {
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const float w = Wt[k * NJ + tid];
        const float4 hv = *reinterpret_cast<const float4*>(hcur + k * kRB);
        acc[0] = fmaf(w, hv.x, acc[0]);
        acc[1] = fmaf(w, hv.y, acc[1]);
        acc[2] = fmaf(w, hv.z, acc[2]);
        acc[3] = fmaf(w, hv.w, acc[3]);
      }
    }
#pragma unroll
    for (int b = 0; b < kRB; ++b) hh[b * NJ + tid] = acc[b] + bhh;
    __syncthreads();

    // phase 2: gate math for my (b,u) items
#pragma unroll
    for (int n = 0; n < IPT; ++n) {
      if (it_b[n] < 0) continue;
      const int b = it_b[n], u = it_u[n];
      float hnew = it_h[n];
      if (s < it_len[n]) {
        if (G == 3) {  // GRU, gate order [r, z, n]
          const float r = sigmoidf_(gxv[n][0] + hh[b * NJ + 0 * HS + u]);
          const float z = sigmoidf_(gxv[n][1] + hh[b * NJ + 1 * HS + u]);
          const float nn = tanhf(gxv[n][2] + r * hh[b * NJ + 2 * HS + u]);
          hnew = (1.f - z) * nn + z * it_h[n];
        } else {  // LSTM, gate order [i, f, g, o]
          const float ig = sigmoidf_(gxv[n][0] + hh[b * NJ + 0 * HS + u]);
          const float fg = sigmoidf_(gxv[n][1] + hh[b * NJ + 1 * HS + u]);
          const float gg = tanhf(gxv[n][2] + hh[b * NJ + 2 * HS + u]);
          const float og = sigmoidf_(gxv[n][3] + hh[b * NJ + 3 * HS + u]);
          it_c[n] = fg * it_c[n] + ig * gg;
          hnew = og * tanhf(it_c[n]);
        }
        it_h[n] = hnew;
        it_max[n] = fmaxf(it_max[n], hnew);
        if (p.y) {
          const int pos = (dir == 0 || p.reverse_dir1 == 0) ? s : it_len[n] - 1 - s;
          p.y[(static_cast<size_t>(b0 + b) * p.T + pos) * (p.ndir * H) + dir * H + q * HS + u] = hnew;
        }
        if (G == 4 && p.cseq) {
          const int pos = (dir == 0 || p.reverse_dir1 == 0) ? s : it_len[n] - 1 - s;
          p.cseq[(static_cast<size_t>(b0 + b) * p.T + pos) * (p.ndir * H) + dir * H + q * HS + u] = it_c[n];
        }
      }
      // publish (frozen rows re-publish their last state so both buffers stay coherent)
      const int hidx = WREG ? b * H + q * HS + u : (q * HS + u) * kRB + b;
      if (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int r = 0; r < CL; ++r) cluster.map_shared_rank(hnext_local, r)[hidx] = hnew;
      } else {
        hnext_local[hidx] = hnew;
      }
    }
#pragma unroll
    for (int n = 0; n < IPT; ++n) {
#pragma unroll
      for (int g = 0; g < G; ++g) gxv[n][g] = gxn[n][g];
    }
    if (CL > 1) cg::this_cluster().sync(); else __syncthreads();
  }

  // ---- outputs
#pragma unroll
  for (int n = 0; n < IPT; ++n) {
    if (it_b[n] < 0) continue;
    const size_t o = (static_cast<size_t>(b0 + it_b[n])) * (p.ndir * H) + dir * H + q * HS + it_u[n];
    if (p.ymax) p.ymax[o] = it_max[n];
    const size_t so = (static_cast<size_t>(dir) * p.B + b0 + it_b[n]) * H + q * HS + it_u[n];
    if (p.h_final) p.h_final[so] = it_h[n];
    if (G == 4 && p.c_final) p.c_final[so] = it_c[n];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// GRU, H = 128 (the event detector's head, definitions.py:94-96): K-split scan.
// The 384-thread register-resident kernel above keeps 128 recurrent weights per thread; with 168 registers per thread ptxas has no
// room left to software-pipeline the hidden-state loads, every FFMA2 waits ~30 cycles on its own LDS and a step takes ~7400
// cycles at 25 % issue utilisation (ncu: stall_short_scoreboard 5.9 warps per issue, profiles/r2_gru_head.md).  Here TWO threads
// share one gate column: 768 threads, 64 weights each; threads 0-383 take k in [0,64), threads 384-767 k in [64,128) (warp-
// uniform, so every hidden-state LDS.128 stays a single-address broadcast) and the two partial dot products meet in shared memory.
// Twice the warps hide the shared-memory latency.
constexpr int kKS = 64;  // weights per thread
__global__ void __launch_bounds__(768, 1) gru_scan_h128_kernel(const RnnScanParams p) {
  constexpr int H = 128, G = 3, NJ = G * H, GH = G * H;
  __shared__ __align__(16) float hbuf[2][kRB][H];
  __shared__ __align__(16) float hh[2][kRB][NJ];  // partial sums of the two k halves
  __shared__ float gxs[2][G][kRB * H];            // projected gates of the current / next step, one item per thread
  const int tid = threadIdx.x;
  const int kh = tid >= NJ ? 1 : 0;  // which half of k (warp-uniform: 384 = 12 warps)
  const int j = tid - kh * NJ;       // gate column (g, u)
  const int tile = blockIdx.x, dir = blockIdx.y;
  const int b0 = tile * kRB;

  float wreg[kKS];
  {
    const float* src = p.WhhT + static_cast<size_t>(dir) * H * GH + j + static_cast<size_t>(kh * kKS) * GH;
#pragma unroll
    for (int k = 0; k < kKS; ++k) wreg[k] = __ldg(src + static_cast<size_t>(k) * GH);
  }
  const float bhh = kh == 0 ? __ldg(p.bhh + static_cast<size_t>(dir) * GH + j) : 0.f;
  for (int idx = tid; idx < kRB * H; idx += 768) {
    const int b = idx / H, k = idx - b * H;
    float v = 0.f;
    if (p.h0 && b0 + b < p.B) v = p.h0[(static_cast<size_t>(dir) * p.B + b0 + b) * H + k];
    hbuf[0][b][k] = v;
    hbuf[1][b][k] = v;
  }
  // phase-2 item of this thread: (row b, unit u) for tid < kRB * H
  const bool has_item = tid < kRB * H && b0 + tid / H < p.B;
  const int ib = tid / H, iu = tid - ib * H;
  int ilen = 0;
  float ih = 0.f, imax = -INFINITY;
  if (has_item) {
    ilen = p.valid_len ? min(max(p.valid_len[b0 + ib], 0), p.T) : p.T;
    if (p.h0) ih = p.h0[(static_cast<size_t>(dir) * p.B + b0 + ib) * H + iu];
  }
  int maxlen = 0;
  for (int b = 0; b < kRB; ++b)
    if (b0 + b < p.B) maxlen = max(maxlen, p.valid_len ? min(max(p.valid_len[b0 + b], 0), p.T) : p.T);
  const int gx_row = p.ndir * GH;
  // The projected gates of step s+1 are fetched with cp.async straight into shared memory while step s computes: no registers
  // are held across the step (ptxas spilled the prefetched values of the register version to local memory, which made every step
  // wait for the global load it was meant to hide -- STL at 7 % of the samples in profiles/r2_summary.md section 5).
  auto prefetch_gx = [&](int s) {
    if (has_item && s < ilen) {
      const int pos = (dir == 0 || p.reverse_dir1 == 0) ? s : ilen - 1 - s;
      const float* gp = p.gx + (static_cast<size_t>(b0 + ib) * p.T + pos) * gx_row + dir * GH + iu;
#pragma unroll
      for (int g = 0; g < G; ++g)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&gxs[s & 1][g][tid]))), "l"(gp + g * H) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch_gx(0);
  __syncthreads();

  for (int s = 0; s < maxlen; ++s) {
    const float* hcur = &hbuf[s & 1][0][0] + kh * kKS;
    prefetch_gx(s + 1);  // next step's gates: their L2/HBM latency hides behind this step
    float part[kRB];
#pragma unroll
    for (int bp = 0; bp < kRB; bp += 2) {
      float2 a0 = make_float2(0.f, 0.f), a0b = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f), a1b = make_float2(0.f, 0.f);
      const float* h0 = hcur + bp * H;
      const float* h1 = h0 + H;
#pragma unroll
      for (int k = 0; k < kKS; k += 8) {
        const float4 p0 = *reinterpret_cast<const float4*>(h0 + k);
        const float4 p1 = *reinterpret_cast<const float4*>(h0 + k + 4);
        const float4 q0 = *reinterpret_cast<const float4*>(h1 + k);
        const float4 q1 = *reinterpret_cast<const float4*>(h1 + k + 4);
        const float2 w01 = make_float2(wreg[k], wreg[k + 1]), w23 = make_float2(wreg[k + 2], wreg[k + 3]);
        const float2 w45 = make_float2(wreg[k + 4], wreg[k + 5]), w67 = make_float2(wreg[k + 6], wreg[k + 7]);
        ffma2(a0, w01, make_float2(p0.x, p0.y));
        ffma2(a1, w01, make_float2(q0.x, q0.y));
        ffma2(a0b, w23, make_float2(p0.z, p0.w));
        ffma2(a1b, w23, make_float2(q0.z, q0.w));
        ffma2(a0, w45, make_float2(p1.x, p1.y));
        ffma2(a1, w45, make_float2(q1.x, q1.y));
        ffma2(a0b, w67, make_float2(p1.z, p1.w));
        ffma2(a1b, w67, make_float2(q1.z, q1.w));
      }
      part[bp] = (a0.x + a0.y) + (a0b.x + a0b.y);
      part[bp + 1] = (a1.x + a1.y) + (a1b.x + a1b.y);
    }
#pragma unroll
    for (int b = 0; b < kRB; ++b) hh[kh][b][j] = part[b] + bhh;
    __syncthreads();

    asm volatile("cp.async.wait_group 1;" ::: "memory");  // this step's gates have landed (only the group of step s+1 may be pending)
    float hnew = ih;
    if (has_item && s < ilen) {
      const float gx0 = gxs[s & 1][0][tid], gx1 = gxs[s & 1][1][tid], gx2 = gxs[s & 1][2][tid];
      const float r = sigmoidf_(gx0 + (hh[0][ib][iu] + hh[1][ib][iu]));
      const float z = sigmoidf_(gx1 + (hh[0][ib][H + iu] + hh[1][ib][H + iu]));
      const float nn = tanhf(gx2 + r * (hh[0][ib][2 * H + iu] + hh[1][ib][2 * H + iu]));
      hnew = (1.f - z) * nn + z * ih;
      ih = hnew;
      imax = fmaxf(imax, hnew);
      if (p.y) {
        const int pos = (dir == 0 || p.reverse_dir1 == 0) ? s : ilen - 1 - s;
        p.y[(static_cast<size_t>(b0 + ib) * p.T + pos) * (p.ndir * H) + dir * H + iu] = hnew;
      }
    }
    if (tid < kRB * H) hbuf[(s + 1) & 1][ib][iu] = hnew;  // frozen rows re-publish their last state
    __syncthreads();
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (has_item) {
    const size_t o = static_cast<size_t>(b0 + ib) * (p.ndir * H) + dir * H + iu;
    if (p.ymax) p.ymax[o] = imax;
    if (p.h_final) p.h_final[(static_cast<size_t>(dir) * p.B + b0 + ib) * H + iu] = ih;
  }
}

template <int G, int CL, bool WREG = false>
cudaError_t launch_t(const RnnScanParams& p, cudaStream_t st) {
  const int HS = p.H / CL;
  const int NJ = G * HS;
  const size_t smem = ((WREG ? 0 : static_cast<size_t>(p.H) * NJ) + 2 * p.H * kRB + kRB * NJ) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(rnn_scan_kernel<G, CL, WREG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  const int tiles = (p.B + kRB - 1) / kRB;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * CL, p.ndir, 1);
  cfg.blockDim = dim3(NJ, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CL > 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rnn_scan_kernel<G, CL, WREG>, p);
}

}  // namespace

int rnn_scan_cluster_size(int G, int H) {
  // smallest power-of-two cluster whose per-CTA W_hh^T slice (+ state) fits in 224 KB of shared memory
  for (int cl = 1; cl <= 8; cl *= 2) {
    const size_t nj = static_cast<size_t>(G) * (H / cl);
    const size_t bytes = (static_cast<size_t>(H) * nj + 2 * H * kRB + kRB * nj) * sizeof(float);
    if (H % cl == 0 && bytes <= 224 * 1024 && nj <= 1024) return cl;
  }
  return -1;
}

cudaError_t launch_rnn_scan(const RnnScanParams& p, cudaStream_t st) {
  if (p.B == 0 || p.T == 0) return cudaSuccess;
  const int G = p.gates;
  const int cl = rnn_scan_cluster_size(G, p.H);
  if (cl < 0 || (G != 3 && G != 4)) return cudaErrorInvalidValue;
  ProfScope prof_scope(kProfOther, st);
  // flagship hidden size: weights in registers.  GRU: K-split pairs (768 threads x 64 weights); TN_RNN_NO_KSPLIT=1 selects the
  // round-1 kernel (384 threads x 128 weights) for A/B.  LSTM: a 2-CTA cluster of 256 threads x 128 weights.
  if (p.H == kHReg && G == 3 && !getenv("TN_RNN_NO_WREG") && !getenv("TN_RNN_NO_KSPLIT")) {
    const int tiles = (p.B + kRB - 1) / kRB;
    gru_scan_h128_kernel<<<dim3(tiles, p.ndir), 768, 0, st>>>(p);
    return cudaGetLastError();
  }
  if (p.H == kHReg && !getenv("TN_RNN_NO_WREG")) return G == 3 ? launch_t<3, 1, true>(p, st) : launch_t<4, 2, true>(p, st);
  if (G == 3) {
    switch (cl) {
      case 1: return launch_t<3, 1>(p, st);
      case 2: return launch_t<3, 2>(p, st);
      case 4: return launch_t<3, 4>(p, st);
      default: return launch_t<3, 8>(p, st);
    }
  } else {
    switch (cl) {
      case 1: return launch_t<4, 1>(p, st);
      case 2: return launch_t<4, 2>(p, st);
      case 4: return launch_t<4, 4>(p, st);
      default: return launch_t<4, 8>(p, st);
    }
  }
}

}  // namespace tn
