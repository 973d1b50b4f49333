// HBM-bound kernels of the Wav2Vec2 path (sm_100a, SIMT with packed fp32x2 math, 128-bit accesses):
//
//   wave_stats / conv0_fold / conv0 :  extractor layer 0 = Conv1D(k=10, s=5, Cin=1) + GroupNorm(groups == C)
//        + erf-GELU (reference: feature_extractor.py:40-47,54-59; tensorflow_addons.py:207-231).
//        GroupNorm needs per-(b, c) statistics over all T0 frames BEFORE any output can be written.  Because
//        Cin = 1, y[t,c] = sum_j w[j,c] x[5t+j] is linear in the 10-sample window, so
//            sum_t y   = w_c . s         s[j]   = sum_t x[5t+j]
//            sum_t y^2 = w_c^T G w_c     G[i,j] = sum_t x[5t+i] x[5t+j]
//        and the statistics come from the waveform alone (65 numbers per utterance, fp64).  The normalisation
//        is then folded into per-(b, c) weights/bias and the 1.6 GB activation is written exactly once.
//   ln_rows : LayerNorm over the channel axis (+ optional GELU) with fp32 statistics, one warp per row
//        (reference: encoder.py:96-108,116-132,232-234,267-275; feature_extractor.py:50,86-88,93).
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int C0_K = 10;        // taps of extractor layer 0
constexpr int C0_S = 5;         // stride
constexpr int C0_NSTAT = 65;    // 10 sums + 55 upper-triangular Gram entries
constexpr int STATS_WIN_PER_BLOCK = 2048;

// ------------------------------------------------------------------------------------ wave_stats
// Bandwidth-bound: a CTA stages its contiguous slice of the waveform (2048 windows = 10245 samples, 41 KB) in smem with
// coalesced 128-bit loads - every sample is read from HBM exactly once, ~20 independent loads in flight per thread -
// then each thread walks 16 windows out of smem (stride 5 floats: conflict-free) with the 65 running sums in fp32
// registers.  The 128 per-thread partials of each statistic are summed through the (re-used) smem slice: 4 serial adds +
// one 5-step shuffle tree per statistic and warp instead of a 65 x 5 fp64 shuffle butterfly in every warp.  Only the
// cross-CTA accumulation (25 CTAs per utterance at L = 246000) is fp64.
constexpr int STATS_THREADS = 128;
constexpr int STATS_SEG = STATS_WIN_PER_BLOCK * C0_S + (C0_K - C0_S);   // samples a CTA touches

__global__ void __launch_bounds__(STATS_THREADS) wave_stats_kernel(const float* __restrict__ wave, int L, int T0,
                                                                   double* __restrict__ stats) {
  extern __shared__ __align__(16) float seg[];   // max(STATS_SEG, 65 * 128) floats
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int w_begin = blockIdx.x * STATS_WIN_PER_BLOCK;
  const int nwin = min(T0 - w_begin, STATS_WIN_PER_BLOCK);
  const float* x = wave + (size_t)b * L + (size_t)w_begin * C0_S;
  const int n = nwin * C0_S + (C0_K - C0_S);      // valid samples of this slice (all inside the utterance)
  const int tid = threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    const int n4 = n >> 2;
    for (int i = tid; i < n4; i += STATS_THREADS) reinterpret_cast<float4*>(seg)[i] = __ldg(reinterpret_cast<const float4*>(x) + i);
    for (int i = 4 * n4 + tid; i < n; i += STATS_THREADS) seg[i] = __ldg(x + i);
  } else {
    for (int i = tid; i < n; i += STATS_THREADS) seg[i] = __ldg(x + i);
  }
  __syncthreads();
  float acc[C0_NSTAT];
#pragma unroll
  for (int i = 0; i < C0_NSTAT; ++i) acc[i] = 0.0f;
  for (int w = tid; w < nwin; w += STATS_THREADS) {
    float v[C0_K];
#pragma unroll
    for (int j = 0; j < C0_K; ++j) v[j] = seg[w * C0_S + j];
    int q = C0_K;
#pragma unroll
    for (int i = 0; i < C0_K; ++i) {
      acc[i] += v[i];
#pragma unroll
      for (int j = i; j < C0_K; ++j) {
        acc[q] = fmaf(v[i], v[j], acc[q]);
        ++q;
      }
    }
  }
  __syncthreads();                                  // the slice is dead: reuse it as red[65][128]
#pragma unroll
  for (int i = 0; i < C0_NSTAT; ++i) seg[i * STATS_THREADS + tid] = acc[i];
  __syncthreads();
  const int lane = lane_id(), warp = tid >> 5;
  for (int i = warp; i < C0_NSTAT; i += STATS_THREADS / 32) {
    const float4 p = *reinterpret_cast<const float4*>(seg + i * STATS_THREADS + 4 * lane);
    const float s = warp_sum((p.x + p.y) + (p.z + p.w));
    if (lane == 0) atomicAdd(stats + (size_t)b * C0_NSTAT + i, (double)s);
  }
}

// ------------------------------------------------------------------------------------ conv0_fold
// folded[b][j][c] = w[j][c] * gamma[c] * rstd[b,c];  fbias[b][c] = beta[c] - mean[b,c] * gamma[c] * rstd[b,c]
__global__ void conv0_fold_kernel(const float* __restrict__ kernel /*[10][C]*/, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, const double* __restrict__ stats, int C, int T0,
                                  float eps, float* __restrict__ folded /*[B][10][C] or null*/,
                                  float* __restrict__ fbias /*[B][C]*/, float* __restrict__ fscale /*[B][C] or null*/) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* st = stats + (size_t)b * C0_NSTAT;
  double w[C0_K];
#pragma unroll
  for (int j = 0; j < C0_K; ++j) w[j] = (double)kernel[j * C + c];
  double sum = 0.0, sq = 0.0;
  int q = C0_K;
#pragma unroll
  for (int i = 0; i < C0_K; ++i) {
    sum += w[i] * st[i];
#pragma unroll
    for (int j = i; j < C0_K; ++j) {
      const double g = st[q++];
      sq += (i == j ? 1.0 : 2.0) * w[i] * w[j] * g;
    }
  }
  const double mean = sum / (double)T0;
  double var = sq / (double)T0 - mean * mean;  // biased variance (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  const double scale = (double)gamma[c] / sqrt(var + (double)eps);
  if (folded != nullptr) {
#pragma unroll
    for (int j = 0; j < C0_K; ++j) folded[((size_t)b * C0_K + j) * C + c] = (float)(w[j] * scale);
  }
  fbias[(size_t)b * C + c] = (float)((double)beta[c] - mean * scale);
  if (fscale != nullptr) fscale[(size_t)b * C + c] = (float)scale;
}

// ------------------------------------------------------------------------------------ conv0 im2col
// a[b][t][0:64] = bf16(x[b][5t + j]) for j < 10, zero above: one 128-byte row per frame (and per plane).
__global__ void __launch_bounds__(256) conv0_im2col_kernel(const float* __restrict__ wave, int L, int T0,
                                                           __nv_bfloat16* __restrict__ a_hi,
                                                           __nv_bfloat16* __restrict__ a_lo) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T0) return;
  const float* x = wave + (size_t)b * L + (size_t)t * C0_S;
  float v[12];
#pragma unroll
  for (int j = 0; j < C0_K; ++j) v[j] = __ldg(x + j);
  v[10] = v[11] = 0.0f;
  uint32_t hi[6], lo[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) hi[j] = split_bf16x2(v[2 * j], v[2 * j + 1], lo[j]);
  const size_t row = ((size_t)b * T0 + t) * 64;
  uint4* ph = reinterpret_cast<uint4*>(a_hi + row);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  ph[1] = make_uint4(hi[4], hi[5], 0u, 0u);
#pragma unroll
  for (int j = 2; j < 8; ++j) ph[j] = z;
  if (a_lo != nullptr) {
    uint4* pl = reinterpret_cast<uint4*>(a_lo + row);
    pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    pl[1] = make_uint4(lo[4], lo[5], 0u, 0u);
#pragma unroll
    for (int j = 2; j < 8; ++j) pl[j] = z;
  }
}

// ------------------------------------------------------------------------------------ conv0 main
// One CTA = 256 frames x 512 channels of one utterance.  Thread (fl, cg): frames [64 fl, 64 fl + 64),
// channels [8 cg, 8 cg + 8).  The waveform segment sits in smem as duplicated pairs (x, x) so that one
// LDS.64 feeds an FFMA2 whose other operand is a (w[c], w[c+1]) register pair; consecutive frames share
// 5 of their 10 samples, so the window slides in registers (5 new LDS.64 per frame).
constexpr int C0_FRAMES_PER_CTA = 256;
constexpr int C0_THREADS = 256;

template <bool GELU, bool OUT_F32>
__global__ void __launch_bounds__(C0_THREADS)
conv0_kernel(const float* __restrict__ wave, int L, int T0, int C, const float* __restrict__ wts, int wts_batch_stride,
             const float* __restrict__ bias, int bias_batch_stride, float* __restrict__ out_f32,
             __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  __shared__ __align__(16) float2 xs[C0_FRAMES_PER_CTA * C0_S + C0_K];
  const int b = blockIdx.y;
  const int t_base = blockIdx.x * C0_FRAMES_PER_CTA;
  const float* x = wave + (size_t)b * L;
  const int first = t_base * C0_S;
  for (int i = threadIdx.x; i < C0_FRAMES_PER_CTA * C0_S + C0_K; i += C0_THREADS) {
    const int g = first + i;
    const float v = (g < L) ? __ldg(x + g) : 0.0f;
    xs[i] = make_float2(v, v);
  }
  const int cg = threadIdx.x & 63;   // channel group (8 channels)
  const int fl = threadIdx.x >> 6;   // frame lane (0..3): 64 consecutive frames each
  const int c0 = cg * 8;
  // per-thread weights: w2[j][p] = (w[j][c0+2p], w[j][c0+2p+1])
  uint64_t w2[C0_K][4];
  uint64_t b2[4];
  {
    const float* wp = wts + (size_t)b * wts_batch_stride + c0;
#pragma unroll
    for (int j = 0; j < C0_K; ++j) {
      const float4 lo = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * C));
      const float4 hi = __ldg(reinterpret_cast<const float4*>(wp + (size_t)j * C) + 1);
      w2[j][0] = pack2(lo.x, lo.y);
      w2[j][1] = pack2(lo.z, lo.w);
      w2[j][2] = pack2(hi.x, hi.y);
      w2[j][3] = pack2(hi.z, hi.w);
    }
    if (bias != nullptr) {
      const float* bp = bias + (size_t)b * bias_batch_stride + c0;
      const float4 lo = __ldg(reinterpret_cast<const float4*>(bp));
      const float4 hi = __ldg(reinterpret_cast<const float4*>(bp) + 1);
      b2[0] = pack2(lo.x, lo.y);
      b2[1] = pack2(lo.z, lo.w);
      b2[2] = pack2(hi.x, hi.y);
      b2[3] = pack2(hi.z, hi.w);
    } else {
      b2[0] = b2[1] = b2[2] = b2[3] = pack2(0.0f, 0.0f);
    }
  }
  __syncthreads();

  const int f_begin = fl * 64;
  const uint64_t* xp = reinterpret_cast<const uint64_t*>(xs) + f_begin * C0_S;
  uint64_t win[C0_K];
#pragma unroll
  for (int j = 0; j < C0_S; ++j) win[C0_S + j] = xp[j];  // becomes taps 0..4 of the first frame
  for (int f = 0; f < 64; ++f) {
    const int t = t_base + f_begin + f;
    if (t >= T0) break;
#pragma unroll
    for (int j = 0; j < C0_S; ++j) {
      win[j] = win[C0_S + j];
      win[C0_S + j] = xp[(f + 1) * C0_S + j];
    }
    uint64_t acc[4] = {b2[0], b2[1], b2[2], b2[3]};
#pragma unroll
    for (int j = 0; j < C0_K; ++j) {
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[p] = fma2(win[j], w2[j][p], acc[p]);
    }
    float v[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) unpack2(acc[p], v[2 * p], v[2 * p + 1]);
    if (GELU) {
#pragma unroll
      for (int p = 0; p < 4; ++p) gelu_erf_x2(v[2 * p], v[2 * p + 1]);
    }
    const size_t o = ((size_t)b * T0 + t) * C + c0;
    if (OUT_F32) {
      float4* op = reinterpret_cast<float4*>(out_f32 + o);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) hi[p] = split_bf16x2(v[2 * p], v[2 * p + 1], lo[p]);
      *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (out_lo != nullptr) *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ------------------------------------------------------------------------------------ ln_rows
// y = (x - mean) * rsqrt(var + eps) * gamma + beta over the last axis (biased variance, two-pass in
// registers), optional GELU, outputs fp32 and/or bf16 hi(/lo).  One warp per row, float4 accesses.
// EXACT: d == 128 * MAXV (every lane owns exactly MAXV float4: 512 / 768 / 1024 channels) - no bounds predicates.
template <int MAXV, bool EXACT = false>
__global__ void __launch_bounds__(256, (MAXV <= 6) ? 3 : (MAXV <= 8) ? 2 : 1)
ln_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
               int rows, int d, int gelu, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
               __nv_bfloat16* __restrict__ out_lo, float* __restrict__ stats, int out_format) {
  pdl_trigger();
  pdl_wait();
  // Persistent over rows: warp w of the grid takes rows w, w + W, ... and fetches its NEXT row before it reduces the current
  // one, so every warp keeps two rows (2 x 4 d bytes) in flight - a grid of one-row warps drains and refills the SM instead.
  const int lane = lane_id();
  const int nvec = d >> 2;
  const int wstride = gridDim.x * 8;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 nx[MAXV];
  {
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)row * d);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) nx[i] = (EXACT || lane + 32 * i < nvec) ? __ldg(xp + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (; row < rows; row += wstride) {
  float4 v[MAXV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    v[i] = nx[i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  if (row + wstride < rows) {
    const float4* xn = reinterpret_cast<const float4*>(x + (size_t)(row + wstride) * d);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) nx[i] = (EXACT || lane + 32 * i < nvec) ? __ldg(xn + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (EXACT || lane + 32 * i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  if (stats != nullptr && lane == 0) *reinterpret_cast<float2*>(stats + 2 * (size_t)row) = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (EXACT || idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      float y0 = fmaf((v[i].x - mean) * rstd, g.x, bt.x);
      float y1 = fmaf((v[i].y - mean) * rstd, g.y, bt.y);
      float y2 = fmaf((v[i].z - mean) * rstd, g.z, bt.z);
      float y3 = fmaf((v[i].w - mean) * rstd, g.w, bt.w);
      if (gelu == 2) {          // tf.nn.gelu(approximate=True): config.py:14
        y0 = gelu_tanh_tf(y0);
        y1 = gelu_tanh_tf(y1);
        y2 = gelu_tanh_tf(y2);
        y3 = gelu_tanh_tf(y3);
      } else if (gelu) {
        gelu_erf_x2(y0, y1);
        gelu_erf_x2(y2, y3);
      }
      const size_t o = (size_t)row * d + 4 * (size_t)idx;
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(y0, y1, y2, y3);
      if (out_hi != nullptr && out_format == 0) {
        uint32_t l0, l1;
        const uint32_t h0 = split_bf16x2(y0, y1, l0), h1 = split_bf16x2(y2, y3, l1);
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(h0, h1);
        if (out_lo != nullptr) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(l0, l1);
      } else if (out_hi != nullptr && out_format == 1) {     // fp16 planes of y * 2^4
        uint32_t l0, l1;
        const uint32_t h0 = split_f16x2(y0 * ACT_SCALE, y1 * ACT_SCALE, l0), h1 = split_f16x2(y2 * ACT_SCALE, y3 * ACT_SCALE, l1);
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(h0, h1);
        if (out_lo != nullptr) *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(l0, l1);
      } else if (out_hi != nullptr) {                        // fp16 hi + e4m3 pair plane [rows][2 d] (d % 64 == 0)
        uint16_t la, lb, ha, hb;
        const uint32_t h0 = split_f16_f8x2(y0 * ACT_SCALE, y1 * ACT_SCALE, la, ha);
        const uint32_t h1 = split_f16_f8x2(y2 * ACT_SCALE, y3 * ACT_SCALE, lb, hb);
        *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(h0, h1);
        const int c = 4 * idx;
        uint8_t* p8 = reinterpret_cast<uint8_t*>(out_lo) + (size_t)row * d * 2 + (size_t)(c >> 6) * 128 + (c & 63);
        *reinterpret_cast<uint32_t*>(p8) = la | ((uint32_t)lb << 16);
        *reinterpret_cast<uint32_t*>(p8 + 64) = ha | ((uint32_t)hb << 16);
      }
    }
  }
  }  // rows of this warp
}

// ------------------------------------------------------------------------------------ LayerNorm statistics from partial sums
// stats[r] = (mean, rstd) of row r from `parts` partial (sum, sum of squares) pairs laid out [parts][rows][2] - what the residual
// GEMMs write through w2v2_gemm_args.row_stats_out (one pair per 64-column group).  Added in index order: deterministic.
__global__ void __launch_bounds__(256)
row_stats_finalize_kernel(const float2* __restrict__ parts, int nparts, int rows, float inv_dim, float eps, float2* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  float s1 = 0.0f, s2 = 0.0f;
  for (int i = 0; i < nparts; ++i) {
    const float2 v = __ldg(parts + (size_t)i * rows + r);
    s1 += v.x;
    s2 += v.y;
  }
  const float mean = s1 * inv_dim;
  const float var = fmaxf(fmaf(-mean, mean, s2 * inv_dim), 0.0f);
  stats[r] = make_float2(mean, rsqrtf(var + eps));
}

// ------------------------------------------------------------------------------------ utterance normalisation
// Wav2Vec2Processor._normalize (processor.py:101-106): (x - mean) / sqrt(var + 1e-5), biased variance, per utterance over
// its `len` real samples ("before padding", data_utils.py:233); samples past `len` are written as the padding value 0.
__global__ void __launch_bounds__(1024)
normalize_utterance_kernel(const float* __restrict__ x, const int* __restrict__ lengths, int L, float eps,
                           float* __restrict__ out) {
  __shared__ double red[2][32];
  __shared__ float s_mean, s_rstd;
  const int b = blockIdx.x;
  const int n = lengths ? min(max(lengths[b], 0), L) : L;
  const float* xp = x + (size_t)b * L;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) {
    const double v = (double)__ldg(xp + i);
    s += v;
    q += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane_id() == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < 32; ++w) {
      ts += red[0][w];
      tq += red[1][w];
    }
    const double mean = n > 0 ? ts / n : 0.0;
    double var = n > 0 ? tq / n - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    s_mean = (float)mean;
    s_rstd = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float mean = s_mean, rstd = s_rstd;
  float* op = out + (size_t)b * L;
  for (int i = threadIdx.x; i < L; i += 1024) op[i] = (i < n) ? (__ldg(xp + i) - mean) * rstd : 0.0f;
}

// fp32 -> bf16 hi(/lo) planes (weight packing, staging test inputs)
__global__ void split_bf16_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo != nullptr) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_wave_stats(const float* wave, int batch, int num_samples, double* stats, void* stream) {
  W2V2_CHECK_ARG(wave && stats, "null pointer");
  W2V2_CHECK_ARG(batch > 0 && num_samples >= C0_K, "need batch > 0 and at least 10 samples");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - C0_K) / C0_S;
  W2V2_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * C0_NSTAT * batch, s));
  dim3 grid((T0 + STATS_WIN_PER_BLOCK - 1) / STATS_WIN_PER_BLOCK, batch);
  constexpr size_t smem = sizeof(float) * (STATS_SEG > C0_NSTAT * STATS_THREADS ? STATS_SEG : C0_NSTAT * STATS_THREADS);
  W2V2_CUDA(launch_pdl(wave_stats_kernel, grid, dim3(STATS_THREADS), smem, s, 0, wave, num_samples, T0, stats));
  return 0;
}

extern "C" int w2v2_conv0_im2col(const float* wave, int batch, int num_samples, void* a_hi, void* a_lo, void* stream) {
  W2V2_CHECK_ARG(wave && a_hi, "null pointer");
  W2V2_CHECK_ARG(batch > 0 && num_samples >= C0_K, "need batch > 0 and at least 10 samples");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - C0_K) / C0_S;
  dim3 grid((T0 + 255) / 256, batch);
  W2V2_CUDA(launch_pdl(conv0_im2col_kernel, grid, dim3(256), 0, s, 0, wave, num_samples, T0,
                       reinterpret_cast<__nv_bfloat16*>(a_hi), reinterpret_cast<__nv_bfloat16*>(a_lo)));
  return 0;
}

extern "C" int w2v2_conv0_fold(const float* kernel, const float* gamma, const float* beta, const double* stats,
                               int batch, int num_samples, int channels, float eps, float* folded_w,
                               float* folded_b, float* scale, void* stream) {
  W2V2_CHECK_ARG(kernel && gamma && beta && stats && folded_b && (folded_w || scale), "null pointer");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - C0_K) / C0_S;
  dim3 grid((channels + 127) / 128, batch);
  W2V2_CUDA(launch_pdl(conv0_fold_kernel, grid, dim3(128), 0, s, 0, kernel, gamma, beta, stats, channels, T0, eps, folded_w,
                       folded_b, scale));
  return 0;
}

extern "C" int w2v2_conv0(const float* wave, int batch, int num_samples, int channels, const float* weights,
                          int weights_batch_stride, const float* bias, int bias_batch_stride, int gelu,
                          float* out_f32, void* out_hi, void* out_lo, void* stream) {
  W2V2_CHECK_ARG(wave && weights, "null pointer");
  W2V2_CHECK_ARG(channels == 512, "extractor layer 0 is built for 512 output channels");
  W2V2_CHECK_ARG((out_f32 != nullptr) != (out_hi != nullptr), "exactly one of out_f32 / out_hi");
  W2V2_CHECK_ARG(batch > 0 && num_samples >= C0_K, "need batch > 0 and at least 10 samples");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - C0_K) / C0_S;
  dim3 grid((T0 + C0_FRAMES_PER_CTA - 1) / C0_FRAMES_PER_CTA, batch);
  auto* hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  auto* lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  if (out_f32 != nullptr) {
    if (gelu)
      conv0_kernel<true, true><<<grid, C0_THREADS, 0, s>>>(wave, num_samples, T0, channels, weights, weights_batch_stride,
                                                           bias, bias_batch_stride, out_f32, hi, lo);
    else
      conv0_kernel<false, true><<<grid, C0_THREADS, 0, s>>>(wave, num_samples, T0, channels, weights,
                                                            weights_batch_stride, bias, bias_batch_stride, out_f32, hi, lo);
  } else {
    if (gelu)
      conv0_kernel<true, false><<<grid, C0_THREADS, 0, s>>>(wave, num_samples, T0, channels, weights,
                                                            weights_batch_stride, bias, bias_batch_stride, out_f32, hi, lo);
    else
      conv0_kernel<false, false><<<grid, C0_THREADS, 0, s>>>(wave, num_samples, T0, channels, weights,
                                                             weights_batch_stride, bias, bias_batch_stride, out_f32, hi, lo);
  }
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_ln_rows(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d,
                            int gelu, float* out_f32, void* out_hi, void* out_lo, void* stream) {
  return w2v2_ln_rows_stats(x, gamma, beta, eps, rows, d, gelu, out_f32, out_hi, out_lo, nullptr, stream);
}

extern "C" int w2v2_ln_rows_stats(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d,
                                  int gelu, float* out_f32, void* out_hi, void* out_lo, float* stats, void* stream) {
  return w2v2_ln_rows_ex(x, gamma, beta, eps, rows, d, gelu, out_f32, out_hi, out_lo, stats, 0, stream);
}

extern "C" int w2v2_ln_rows_ex(const float* x, const float* gamma, const float* beta, float eps, int64_t rows, int d,
                               int gelu, float* out_f32, void* out_hi, void* out_lo, float* stats, int out_format,
                               void* stream) {
  W2V2_CHECK_ARG(x && gamma && beta, "null pointer");
  W2V2_CHECK_ARG(out_format >= 0 && out_format <= 2, "out_format must be 0 (bf16), 1 (fp16) or 2 (fp16 + e4m3 pairs)");
  W2V2_CHECK_ARG(out_format != 2 || (out_hi && out_lo && d % 64 == 0), "out_format 2 writes both planes and needs d % 64 == 0");
  W2V2_CHECK_ARG(d > 0 && d % 4 == 0 && d <= 2048, "d must be a multiple of 4, at most 2048");
  W2V2_CHECK_ARG(out_f32 || out_hi || stats, "at least one output");
  W2V2_CHECK_ARG(out_lo == nullptr || out_hi != nullptr, "out_lo requires out_hi");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto* hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  auto* lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
#define LN_LAUNCH(MAXV, EXACT)                                                                                          \
  do {                                                                                                                  \
    auto kern = ln_rows_kernel<MAXV, EXACT>;                                                                            \
    static int resident[64] = {0}; /* persistent grid = resident CTAs of this instantiation, per device */              \
    int dev = 0;                                                                                                        \
    W2V2_CUDA(cudaGetDevice(&dev));                                                                                     \
    if (resident[dev & 63] == 0) {                                                                                      \
      int sms = 0, per_sm = 0;                                                                                          \
      W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));                                     \
      W2V2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));                                  \
      const char* e = getenv("W2V2_LN_CTAS_PER_SM"); /* A/B switch; 999 = one row per warp */                           \
      if (e && atoi(e) > 0) per_sm = atoi(e);                                                                           \
      resident[dev & 63] = sms * (per_sm > 0 ? per_sm : 1);                                                             \
    }                                                                                                                   \
    const long long want = (rows + 7) / 8;                                                                              \
    const unsigned grid = (unsigned)(want < resident[dev & 63] ? want : resident[dev & 63]);                            \
    W2V2_CUDA(launch_pdl(kern, dim3(grid), dim3(256), 0, s, 0, x, gamma, beta, eps, (int)rows, d, gelu, out_f32, hi, lo, \
                         stats, out_format));                                                                           \
  } while (0)
  if (d == 512) LN_LAUNCH(4, true);
  else if (d == 768) LN_LAUNCH(6, true);
  else if (d == 1024) LN_LAUNCH(8, true);
  else if (d <= 1024) LN_LAUNCH(8, false);
  else LN_LAUNCH(16, false);
#undef LN_LAUNCH
  return 0;
}

extern "C" int w2v2_row_stats_finalize(const float* parts, int nparts, int64_t rows, int dim, float eps, float* stats,
                                       void* stream) {
  W2V2_CHECK_ARG(parts && stats && nparts > 0 && dim > 0, "null pointer or empty reduction");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  W2V2_CUDA(launch_pdl(row_stats_finalize_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, s, 0,
                       reinterpret_cast<const float2*>(parts), nparts, (int)rows, 1.0f / (float)dim, eps, reinterpret_cast<float2*>(stats)));
  return 0;
}

extern "C" int w2v2_split_bf16(const float* x, int64_t n, void* hi, void* lo, void* stream) {
  W2V2_CHECK_ARG(x && hi, "null pointer");
  if (n <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  split_bf16_kernel<<<grid, 256, 0, s>>>(x, (size_t)n, reinterpret_cast<__nv_bfloat16*>(hi),
                                         reinterpret_cast<__nv_bfloat16*>(lo));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_normalize_utterances(const float* wave, const int32_t* lengths, int batch, int num_samples, float eps,
                                         float* out, void* stream) {
  W2V2_CHECK_ARG(wave && out, "null pointer");
  W2V2_CHECK_ARG(batch > 0 && num_samples > 0, "batch and num_samples must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  normalize_utterance_kernel<<<batch, 1024, 0, s>>>(wave, lengths, num_samples, eps, out);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
