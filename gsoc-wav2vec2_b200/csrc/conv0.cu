// Extractor layer 0 in ONE kernel:  out[b,t,c] = gelu(scale[b,c] * sum_j w[j,c] x[b,5t+j] + shift[b,c])  -> bf16.
// (reference: Conv1D feature_extractor.py:31-37, GroupNorm tensorflow_addons.py:207-231 folded to scale/shift by
//  w2v2_wave_stats + w2v2_conv0_fold, GELU feature_extractor.py:58.)
//
// This is the HBM-bound kernel of the path: 4 L bytes in, 2 * 512 * T0 bytes out per utterance, 10 MACs + one GELU per
// output.  Design:
//  * no im2col in HBM: a CTA stages its 256-frame slice of the waveform (1296 samples) in shared memory as bf16, twice
//    (the second copy shifted by one sample), so that the 10-tap window of ANY frame starts on a 4-byte boundary in one of
//    the two copies and an A fragment of the warp-level MMA is four plain LDS.32;
//  * the 10 MACs run on the tensor cores (mma.sync m16n8k16, taps padded 10 -> 16 with zero weights): 1/2 instruction
//    per output pair instead of 5 FFMA2, which leaves the issue slots to the GELU;
//  * the weight (B) fragments, scale and shift of a warp's 64 channels live in registers for the whole CTA; the
//    n-tile -> channel map is permuted so that the 8 accumulator pairs of a lane are 2 x 8 CONSECUTIVE channels:
//    outputs leave as 16-byte stores, a quad writes 64 contiguous bytes, no shuffles, no smem transpose;
//  * PASSES == 3 (parity mode): x and w as bf16 hi+lo planes, 3 MMAs, erf-exact GELU, hi+lo outputs;
//    PASSES == 1 (throughput mode): single MMA, tanh-form GELU (|err| < 5e-4, below bf16 resolution), hi output.
//  * LN (the "layer"-norm extractor of the robust / large checkpoints, feature_extractor.py:48-50): the normalisation runs over
//    the 512 CHANNELS of a frame, and a CTA holds all of them for its 256 frames - the row statistics are two block reductions
//    per 16-frame m-tile (quad shuffle, 8 warp partials through smem; two-pass: mean, then centred squares) and the fp32
//    conv output (1.6 GB written and read back by w2v2_ln_rows at 16 x 246000) never exists.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int CM_TT = 256;                 // frames per CTA
constexpr int CM_NS = CM_TT * 5 + 32;      // staged samples (window of the last frame + fragment over-read, zero filled)
constexpr int CM_C = 512;
#ifndef CONV0_STREAM_STORES
#define CONV0_STREAM_STORES 1
#endif

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// TF_APPROX: tf.nn.gelu(approximate=True), the reference's is_gelu_approx switch.  OUT_FMT (w2v2.h W2V2_OUT_*): 0 = bf16 hi
// (+ lo when PASSES == 3), 1 = fp16 plane of value * 2^4, 2 = fp16 plane + e4m3 pair plane (fp16f8 mode); formats 1 / 2 always
// run the 3-MMA window product (the kernel is HBM-bound, the extra MMAs are free) and the erf-exact GELU.
// LN: scale = gamma [512], shift = beta [512] (no batch axis), cbias = conv bias [512] or null, eps = LayerNorm epsilon
template <int PASSES, bool TF_APPROX = false, int OUT_FMT = 0, bool LN = false>
__global__ void __launch_bounds__(256, 2)
conv0_mma_kernel(const float* __restrict__ wave, int L, int T0, const float* __restrict__ kernel /*[10][512]*/,
                 const float* __restrict__ scale /*[B][512]*/, const float* __restrict__ shift /*[B][512]*/,
                 __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                 const float* __restrict__ cbias = nullptr, float eps = 0.0f) {
  constexpr int NPLANES = (PASSES == 3) ? 2 : 1;
  // xs[plane][copy][i]: copy 0 = samples, copy 1 = samples shifted by one (xs[.][1][i] = x[i + 1])
  __shared__ __align__(16) __nv_bfloat16 xs[NPLANES][2][CM_NS];
  __shared__ float red[LN ? 2 : 1][8][16];     // LN: [pass][warp][row of the m-tile] partial sums
  const int b = blockIdx.y;
  const int t_base = blockIdx.x * CM_TT;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int wbase = warp * 64;

  // ---- per-lane constants (weights of the module: safe to read before the predecessor kernel has finished)
  // B fragment of n-tile i, column n = g  <->  channel wbase + 32 (i >> 2) + 8 (g >> 1) + 2 (i & 3) + (g & 1)
  uint32_t bh[8][2], bl[(PASSES == 3) ? 8 : 1][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ch = wbase + 32 * (i >> 2) + 8 * (g >> 1) + 2 * (i & 3) + (g & 1);
    const float w0 = __ldg(kernel + (2 * q) * CM_C + ch), w1 = __ldg(kernel + (2 * q + 1) * CM_C + ch);
    const float w8 = (q == 0) ? __ldg(kernel + 8 * CM_C + ch) : 0.0f, w9 = (q == 0) ? __ldg(kernel + 9 * CM_C + ch) : 0.0f;
    if (PASSES == 3) {
      bh[i][0] = split_bf16x2(w0, w1, bl[i][0]);
      bh[i][1] = split_bf16x2(w8, w9, bl[i][1]);
    } else {
      bh[i][0] = pack_bf16x2(w0, w1);
      bh[i][1] = pack_bf16x2(w8, w9);
    }
  }
  pdl_trigger();
  pdl_wait();

  // ---- stage the waveform slice
  {
    const float* x = wave + (size_t)b * L;
    const int first = t_base * 5;
    for (int i = threadIdx.x; i < CM_NS; i += 256) {
      const int gi = first + i;
      const float v = (gi < L) ? __ldg(x + gi) : 0.0f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      xs[0][0][i] = h;
      if (i > 0) xs[0][1][i - 1] = h;
      if (PASSES == 3) {
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        xs[NPLANES - 1][0][i] = l;
        if (i > 0) xs[NPLANES - 1][1][i - 1] = l;
      }
    }
    if (threadIdx.x == 0) {
      xs[0][1][CM_NS - 1] = __float2bfloat16_rn(0.0f);
      if (PASSES == 3) xs[NPLANES - 1][1][CM_NS - 1] = __float2bfloat16_rn(0.0f);
    }
  }
  // scale / shift of this lane's output channels: n-tile i, columns (2q, 2q+1) <-> channels co(i), co(i) + 1
  uint64_t sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = wbase + 32 * (i >> 2) + 8 * q + 2 * (i & 3);
    const size_t bo = LN ? 0 : (size_t)b * CM_C;
    const float2 s2 = __ldg(reinterpret_cast<const float2*>(scale + bo + co));
    const float2 h2 = __ldg(reinterpret_cast<const float2*>(shift + bo + co));
    sc[i] = pack2(s2.x, s2.y);
    sh[i] = pack2(h2.x, h2.y);
  }
  __syncthreads();

  // window of row r starts at sample 5 r (+ 2 q for this lane's k pair): even offsets read copy 0, odd offsets copy 1
  const int par = g & 1;
  const uint32_t* xw_hi = reinterpret_cast<const uint32_t*>(xs[0][par]) + ((5 * g + 2 * q - par) >> 1);
  const uint32_t* xw_lo = reinterpret_cast<const uint32_t*>(xs[NPLANES - 1][par]) + ((5 * g + 2 * q - par) >> 1);
  const size_t out_base = (size_t)b * T0 * CM_C + wbase + 8 * q;

#pragma unroll 1
  for (int mt = 0; mt < CM_TT / 16; ++mt) {
    const int t0 = t_base + 16 * mt;
    if (t0 >= T0) break;
    uint32_t ah[4], al[4];
    ah[0] = xw_hi[40 * mt];
    ah[1] = xw_hi[40 * mt + 20];   // row g + 8: 40 samples further
    ah[2] = xw_hi[40 * mt + 4];    // taps 2q + 8, 2q + 9
    ah[3] = xw_hi[40 * mt + 24];
    if (PASSES == 3) {
      al[0] = xw_lo[40 * mt];
      al[1] = xw_lo[40 * mt + 20];
      al[2] = xw_lo[40 * mt + 4];
      al[3] = xw_lo[40 * mt + 24];
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
      if (LN && cbias != nullptr) {     // conv bias (config.conv_bias): the accumulators start from it
        const float2 cb = __ldg(reinterpret_cast<const float2*>(cbias + wbase + 32 * (i >> 2) + 8 * q + 2 * (i & 3)));
        acc[i][0] = acc[i][2] = cb.x;
        acc[i][1] = acc[i][3] = cb.y;
      }
      if (PASSES == 3) {
        mma_bf16_16816(acc[i], al, bh[i][0], bh[i][1]);
        mma_bf16_16816(acc[i], ah, bl[i][0], bl[i][1]);
      }
      mma_bf16_16816(acc[i], ah, bh[i][0], bh[i][1]);
    }
    if constexpr (LN) {
      // LayerNorm over the 512 channels of rows g and g + 8 (biased variance, two-pass like w2v2_ln_rows)
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s0 += acc[i][0] + acc[i][1];
        s1 += acc[i][2] + acc[i][3];
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (q == 0) {
        red[0][warp][g] = s0;
        red[0][warp][g + 8] = s1;
      }
      __syncthreads();
      float m0 = 0.0f, m1 = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        m0 += red[0][w][g];
        m1 += red[0][w][g + 8];
      }
      m0 *= 1.0f / CM_C;
      m1 *= 1.0f / CM_C;
      float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] -= m0;
        acc[i][1] -= m0;
        acc[i][2] -= m1;
        acc[i][3] -= m1;
        q0 += acc[i][0] * acc[i][0] + acc[i][1] * acc[i][1];
        q1 += acc[i][2] * acc[i][2] + acc[i][3] * acc[i][3];
      }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
      q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      if (q == 0) {
        red[LN ? 1 : 0][warp][g] = q0;
        red[LN ? 1 : 0][warp][g + 8] = q1;
      }
      __syncthreads();   // (the next m-tile's first partials go to red[0], read by everybody before this barrier)
      float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        v0 += red[LN ? 1 : 0][w][g];
        v1 += red[LN ? 1 : 0][w][g + 8];
      }
      const float r0 = rsqrtf(v0 * (1.0f / CM_C) + eps), r1 = rsqrtf(v1 * (1.0f / CM_C) + eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] *= r0;
        acc[i][1] *= r0;
        acc[i][2] *= r1;
        acc[i][3] *= r1;
      }
    }
    uint32_t oh[2][8], ol[(PASSES == 3) ? 2 : 1][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        uint64_t v = fma2(pack2(acc[i][2 * r], acc[i][2 * r + 1]), sc[i], sh[i]);
        float v0, v1;
        if (OUT_FMT != 0) {
          unpack2(v, v0, v1);
          if (TF_APPROX) {
            v0 = gelu_tanh_tf(v0);
            v1 = gelu_tanh_tf(v1);
          } else {
            gelu_erf_x2(v0, v1);
          }
          if (OUT_FMT == 1) {
            oh[r][i] = pack_f16x2(v0 * ACT_SCALE, v1 * ACT_SCALE);
          } else {
            uint16_t l8, h8;
            oh[r][i] = split_f16_f8x2(v0 * ACT_SCALE, v1 * ACT_SCALE, l8, h8);
            ol[r][i] = l8 | ((uint32_t)h8 << 16);
          }
        } else if (TF_APPROX) {
          unpack2(v, v0, v1);
          v0 = gelu_tanh_tf(v0);
          v1 = gelu_tanh_tf(v1);
          if (PASSES == 3) oh[r][i] = split_bf16x2(v0, v1, ol[r][i]);
          else oh[r][i] = pack_bf16x2(v0, v1);
        } else if (PASSES == 3) {
          unpack2(v, v0, v1);
          gelu_erf_x2(v0, v1);
          oh[r][i] = split_bf16x2(v0, v1, ol[r][i]);
        } else {
          v = gelu_tanh_p2(v);
          unpack2(v, v0, v1);
          oh[r][i] = pack_bf16x2(v0, v1);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int t = t0 + g + 8 * r;
      if (t < T0) {
        __nv_bfloat16* p = out_hi + out_base + (size_t)t * CM_C;
        if (CONV0_STREAM_STORES) {
          __stcs(reinterpret_cast<uint4*>(p), make_uint4(oh[r][0], oh[r][1], oh[r][2], oh[r][3]));
          __stcs(reinterpret_cast<uint4*>(p + 32), make_uint4(oh[r][4], oh[r][5], oh[r][6], oh[r][7]));
        } else {
        *reinterpret_cast<uint4*>(p) = make_uint4(oh[r][0], oh[r][1], oh[r][2], oh[r][3]);
        *reinterpret_cast<uint4*>(p + 32) = make_uint4(oh[r][4], oh[r][5], oh[r][6], oh[r][7]);
        }
        if (OUT_FMT == 2) {
          // e4m3 pair plane [B*T0][2 * 512] bytes: the warp's 64 channels are one 128-byte group; this lane holds channels
          // 8q..8q+7 (n-tiles 0-3) and 32+8q.. (n-tiles 4-7): 8 lo bytes each, the hi bytes 64 further
          uint8_t* p8 = reinterpret_cast<uint8_t*>(out_lo) + ((size_t)b * T0 + t) * (2 * CM_C) + warp * 128 + 8 * q;
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const uint32_t* o4 = &ol[r][4 * hb];
            __stcs(reinterpret_cast<uint2*>(p8 + 32 * hb), make_uint2((o4[0] & 0xFFFFu) | (o4[1] << 16), (o4[2] & 0xFFFFu) | (o4[3] << 16)));
            __stcs(reinterpret_cast<uint2*>(p8 + 32 * hb + 64), make_uint2((o4[0] >> 16) | (o4[1] & 0xFFFF0000u), (o4[2] >> 16) | (o4[3] & 0xFFFF0000u)));
          }
        } else if (PASSES == 3 && OUT_FMT == 0) {
          __nv_bfloat16* pl = out_lo + out_base + (size_t)t * CM_C;
          __stcs(reinterpret_cast<uint4*>(pl), make_uint4(ol[r][0], ol[r][1], ol[r][2], ol[r][3]));
          __stcs(reinterpret_cast<uint4*>(pl + 32), make_uint4(ol[r][4], ol[r][5], ol[r][6], ol[r][7]));
        }
      }
    }
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_conv0_gn_gelu(const float* wave, int batch, int num_samples, int channels, const float* kernel,
                                  const float* scale, const float* shift, void* out_hi, void* out_lo, int passes,
                                  int gelu_approx, void* stream) {
  W2V2_CHECK_ARG(wave && kernel && scale && shift && out_hi, "null pointer");
  W2V2_CHECK_ARG(channels == CM_C, "extractor layer 0 is built for 512 output channels");
  W2V2_CHECK_ARG(batch > 0 && num_samples >= 10, "need batch > 0 and at least 10 samples");
  W2V2_CHECK_ARG(passes == 1 || passes == 3 || passes == 17 || passes == 25, "passes must be 1, 3, 17 (fp16) or 25 (fp16f8)");
  W2V2_CHECK_ARG((passes == 3 || passes == 25) == (out_lo != nullptr), "out_lo is written exactly in the two-plane modes (3, 25)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - 10) / 5;
  dim3 grid((T0 + CM_TT - 1) / CM_TT, batch);
  auto* hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  auto* lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  if (passes == 17 || passes == 25) {
    auto kern = passes == 17 ? (gelu_approx ? conv0_mma_kernel<3, true, 1> : conv0_mma_kernel<3, false, 1>)
                             : (gelu_approx ? conv0_mma_kernel<3, true, 2> : conv0_mma_kernel<3, false, 2>);
    W2V2_CUDA(launch_pdl(kern, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, scale, shift, hi, lo, (const float*)nullptr, 0.0f));
  } else if (gelu_approx) {
    if (passes == 1)
      W2V2_CUDA(launch_pdl(conv0_mma_kernel<1, true>, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, scale, shift, hi, lo, (const float*)nullptr, 0.0f));
    else
      W2V2_CUDA(launch_pdl(conv0_mma_kernel<3, true>, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, scale, shift, hi, lo, (const float*)nullptr, 0.0f));
  } else if (passes == 1)
    W2V2_CUDA(launch_pdl(conv0_mma_kernel<1>, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, scale, shift, hi, lo, (const float*)nullptr, 0.0f));
  else
    W2V2_CUDA(launch_pdl(conv0_mma_kernel<3>, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, scale, shift, hi, lo, (const float*)nullptr, 0.0f));
  return 0;
}

extern "C" int w2v2_conv0_ln_gelu(const float* wave, int batch, int num_samples, int channels, const float* kernel,
                                  const float* conv_bias, const float* gamma, const float* beta, float eps, void* out_hi,
                                  void* out_lo, int passes, int gelu_approx, void* stream) {
  W2V2_CHECK_ARG(wave && kernel && gamma && beta && out_hi, "null pointer");
  W2V2_CHECK_ARG(channels == CM_C, "extractor layer 0 is built for 512 output channels");
  W2V2_CHECK_ARG(batch > 0 && num_samples >= 10, "need batch > 0 and at least 10 samples");
  W2V2_CHECK_ARG(passes == 1 || passes == 3 || passes == 17 || passes == 25, "passes must be 1, 3, 17 (fp16) or 25 (fp16f8)");
  W2V2_CHECK_ARG((passes == 3 || passes == 25) == (out_lo != nullptr), "out_lo is written exactly in the two-plane modes (3, 25)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int T0 = 1 + (num_samples - 10) / 5;
  dim3 grid((T0 + CM_TT - 1) / CM_TT, batch);
  auto* hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  auto* lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  // the LayerNorm output feeds a GELU of unit-variance arguments: every mode takes the erf-exact (or tf-approximate) form,
  // the bf16 mode keeps its single MMA
#define C0LN(P, A, F) W2V2_CUDA(launch_pdl(conv0_mma_kernel<P, A, F, true>, grid, dim3(256), 0, s, 0, wave, num_samples, T0, kernel, \
                                           gamma, beta, hi, lo, conv_bias, eps))
  if (passes == 17) { if (gelu_approx) C0LN(3, true, 1); else C0LN(3, false, 1); }
  else if (passes == 25) { if (gelu_approx) C0LN(3, true, 2); else C0LN(3, false, 2); }
  else if (passes == 1) { if (gelu_approx) C0LN(1, true, 0); else C0LN(1, false, 0); }
  else { if (gelu_approx) C0LN(3, true, 0); else C0LN(3, false, 0); }
#undef C0LN
  return 0;
}
