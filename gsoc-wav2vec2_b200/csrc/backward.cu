// Row / element kernels of the stage-2 fine-tune step (reference: src/main.py:234-250 - Keras `fit` differentiates the
// whole encoder, the conv extractor stays frozen, main.py:236-237).  They are the backward counterparts of the kernels in
// rowops.cu and of the GEMM epilogues; the matrix products of the backward pass reuse the tcgen05 GEMM (w2v2_gemm_bf16):
//     dgrad  dX = dY . W^T          A = dY [M, out],   weight operand = the TF kernel itself ([in, out] == W^T, K-major)
//     wgrad  dW = X^T . dY          A = X^T [in, M],   weight operand = dY^T [out, M]   (w2v2_transpose_bf16 makes both)
//
//   ln_bwd         backward of LayerNormalization (encoder.py:96-108,232-234; feature_extractor.py:86-88)
//   gelu_rows      forward GELU on a saved fp32 pre-activation (training keeps the pre-activation for the backward)
//   dact_colsum    dPre = dAct * gelu'(pre) (+ column sums = bias gradient); without `pre` a plain column sum
//   transpose_bf16 [M, N] -> [N, Mpad] (zero padded) so that a gradient product over the row index becomes a K-major GEMM
//   lm_head_dgrad  dHidden = dLogits . kernel^T (backward of the Dense at modeling.py:231,254)
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

// ------------------------------------------------------------------------------------ LayerNorm backward
// y = xh * gamma + beta, xh = (x - mean) * rstd.  With g = gamma * dy:  dx = rstd * (g - mean(g) - xh * mean(g * xh)),
// dgamma = sum_rows dy * xh, dbeta = sum_rows dy.  One warp per row (statistics recomputed from x, two-pass in registers);
// a CTA walks many rows and keeps its column partial sums in registers, then merges them through shared memory and issues
// one atomicAdd per column.  `colsum` (optional) receives sum_rows dx: the bias gradient of the Dense that produced x.
template <int MAXV, bool EXACT = false>   // EXACT: d == 128 * MAXV, no bounds predicates (512 / 768 / 1024 channels)
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy, float eps,
              int rows, int d, float* __restrict__ dx_f32, __nv_bfloat16* __restrict__ dx_hi,
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ colsum) {
  extern __shared__ float s_part[];  // [3][d]
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int nvec = d >> 2;
  for (int i = threadIdx.x; i < 3 * d; i += 256) s_part[i] = 0.0f;
  __syncthreads();
  float4 ag[MAXV], ab[MAXV], ac[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) ag[i] = ab[i] = ac[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gm[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    gm[i] = (EXACT || lane + 32 * i < nvec) ? __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);

  // one CTA per SM (see the launcher): a warp keeps its NEXT row (x and dy, 8 d bytes) in flight while it reduces the current one
  const int rstride = gridDim.x * 8;
  int row = blockIdx.x * 8 + warp;
  float4 nv[MAXV], ng[MAXV];
  auto fetch = [&](int r) {
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)r * d);
    const float4* dp = reinterpret_cast<const float4*>(dy + (size_t)r * d);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (EXACT || idx < nvec) {
        nv[i] = __ldg(xp + idx);
        ng[i] = __ldg(dp + idx);
      } else {
        nv[i] = ng[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  if (row < rows) fetch(row);
  for (; row < rows; row += rstride) {
    float4 v[MAXV], g[MAXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      v[i] = nv[i];
      g[i] = ng[i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    if (row + rstride < rows) fetch(row + rstride);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (EXACT || lane + 32 * i < nvec) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
    float sa = 0.0f, sb = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      // v <- xh;  accumulate dgamma / dbeta;  g <- gamma * dy
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;
      ag[i].x = fmaf(g[i].x, v[i].x, ag[i].x); ag[i].y = fmaf(g[i].y, v[i].y, ag[i].y);
      ag[i].z = fmaf(g[i].z, v[i].z, ag[i].z); ag[i].w = fmaf(g[i].w, v[i].w, ag[i].w);
      ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
      g[i].x *= gm[i].x; g[i].y *= gm[i].y; g[i].z *= gm[i].z; g[i].w *= gm[i].w;
      sa += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sb += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float ma = warp_sum(sa) / (float)d, mb = warp_sum(sb) / (float)d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (EXACT || idx < nvec) {
        float4 o;
        o.x = rstd * (g[i].x - ma - v[i].x * mb);
        o.y = rstd * (g[i].y - ma - v[i].y * mb);
        o.z = rstd * (g[i].z - ma - v[i].z * mb);
        o.w = rstd * (g[i].w - ma - v[i].w * mb);
        ac[i].x += o.x; ac[i].y += o.y; ac[i].z += o.z; ac[i].w += o.w;
        const size_t off = (size_t)row * d + 4 * (size_t)idx;
        if (dx_f32 != nullptr) *reinterpret_cast<float4*>(dx_f32 + off) = o;
        if (dx_hi != nullptr)
          *reinterpret_cast<uint2*>(dx_hi + off) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    }
  }
  // merge the 8 warps' partial column sums, then one atomic per column and CTA
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (EXACT || idx < nvec) {
      const int c = 4 * idx;
      atomicAdd(&s_part[c + 0], ag[i].x); atomicAdd(&s_part[c + 1], ag[i].y);
      atomicAdd(&s_part[c + 2], ag[i].z); atomicAdd(&s_part[c + 3], ag[i].w);
      atomicAdd(&s_part[d + c + 0], ab[i].x); atomicAdd(&s_part[d + c + 1], ab[i].y);
      atomicAdd(&s_part[d + c + 2], ab[i].z); atomicAdd(&s_part[d + c + 3], ab[i].w);
      if (colsum != nullptr) {
        atomicAdd(&s_part[2 * d + c + 0], ac[i].x); atomicAdd(&s_part[2 * d + c + 1], ac[i].y);
        atomicAdd(&s_part[2 * d + c + 2], ac[i].z); atomicAdd(&s_part[2 * d + c + 3], ac[i].w);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += 256) {
    if (dgamma != nullptr) atomicAdd(dgamma + c, s_part[c]);
    if (dbeta != nullptr) atomicAdd(dbeta + c, s_part[d + c]);
    if (colsum != nullptr) atomicAdd(colsum + c, s_part[2 * d + c]);
  }
}

// ------------------------------------------------------------------------------------ GELU forward on saved pre-activations
template <bool FAST>
__global__ void __launch_bounds__(256)
gelu_rows_kernel(const float* __restrict__ pre, size_t n4, __nv_bfloat16* __restrict__ out_hi,
                 __nv_bfloat16* __restrict__ out_lo, DropSpec dr) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 v = __ldg(reinterpret_cast<const float4*>(pre) + i);
    gelu_x2<FAST>(v.x, v.y);
    gelu_x2<FAST>(v.z, v.w);
    if (dr.thr16) {   // encoder.py:128: dropout on the activated intermediate
      const uint64_t bits = drop_bits4(dr, i);
      v.x = drop_keep(bits, 0, dr.thr16) ? v.x * dr.scale : 0.0f;
      v.y = drop_keep(bits, 1, dr.thr16) ? v.y * dr.scale : 0.0f;
      v.z = drop_keep(bits, 2, dr.thr16) ? v.z * dr.scale : 0.0f;
      v.w = drop_keep(bits, 3, dr.thr16) ? v.w * dr.scale : 0.0f;
    }
    uint32_t l0, l1;
    const uint32_t h0 = split_bf16x2(v.x, v.y, l0), h1 = split_bf16x2(v.z, v.w, l1);
    reinterpret_cast<uint2*>(out_hi)[i] = make_uint2(h0, h1);
    if (out_lo != nullptr) reinterpret_cast<uint2*>(out_lo)[i] = make_uint2(l0, l1);
  }
}

// ------------------------------------------------------------------------------------ activation backward + column sums
// gelu'(x) = Phi(x) + x phi(x)  (derivative of the exact erf form, config.py:14).
__device__ __forceinline__ float gelu_grad(float x) {
  // Phi through the erfc-as-exp2 polynomial of the forward (w2v2_common.cuh: erfc(a / sqrt2) = exp2(a R(a)), |err| 2.6e-7)
  // and phi as one more exp2: two MUFU.EX2 + ~12 FMAs instead of erff + expf (the kernel was instruction-bound)
  const float ax = fabsf(x);
  const float a = fminf(ax, 6.0f);
  float r = 3.2121541153173894e-05f;
  r = fmaf(r, a, -0.0007558754878118634f);
  r = fmaf(r, a, 0.008020005188882351f);
  r = fmaf(r, a, -0.053288985043764114f);
  r = fmaf(r, a, -0.45888903737068176f);
  r = fmaf(r, a, -1.1511517763137817f);
  const float half_erfc = 0.5f * ex2_approx(r * a);                 // Phi(-|x|)
  const float cdf = (x >= 0.0f) ? 1.0f - half_erfc : half_erfc;
  const float pdf = 0.3989422804014327f * ex2_approx(-0.7213475204444817f * x * x);
  return fmaf(x, pdf, cdf);
}
// thread = 4 consecutive columns, walks the rows of its chunk (blockIdx.y); dy bf16, pre fp32 (or null), out bf16 (or null)
__global__ void __launch_bounds__(256)
dact_colsum_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ pre, int rows, int cols,
                   __nv_bfloat16* __restrict__ out_hi, float* __restrict__ colsum, DropSpec dr) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // U rows per batch: all their loads are issued before the first use (a one-row-at-a-time loop is a chain of dependent HBM
  // latencies: 25 us for a 9 MB column sum)
  constexpr int U = 8;
  for (int rb = r0; rb < r1; rb += U) {
    uint2 raw[U];
    float4 pv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = min(rb + u, r1 - 1);                 // clamped: the tail re-reads the last row, its result is discarded
      const size_t off = (size_t)r * cols + c;
      raw[u] = __ldg(reinterpret_cast<const uint2*>(dy + off));
      if (pre != nullptr) pv[u] = __ldg(reinterpret_cast<const float4*>(pre + off));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = rb + u;
      if (r >= r1) break;
      const size_t off = (size_t)r * cols + c;
      float4 g = make_float4(bf16_lo_to_f32(raw[u].x), bf16_hi_to_f32(raw[u].x), bf16_lo_to_f32(raw[u].y), bf16_hi_to_f32(raw[u].y));
      if (dr.thr16) {   // gradient of a dropout that sat AFTER the activation (or after a Dense when pre == null)
        const uint64_t bits = drop_bits4(dr, off >> 2);
        g.x = drop_keep(bits, 0, dr.thr16) ? g.x * dr.scale : 0.0f;
        g.y = drop_keep(bits, 1, dr.thr16) ? g.y * dr.scale : 0.0f;
        g.z = drop_keep(bits, 2, dr.thr16) ? g.z * dr.scale : 0.0f;
        g.w = drop_keep(bits, 3, dr.thr16) ? g.w * dr.scale : 0.0f;
      }
      if (pre != nullptr) {
        const float4 p = pv[u];
        g.x *= gelu_grad(p.x); g.y *= gelu_grad(p.y); g.z *= gelu_grad(p.z); g.w *= gelu_grad(p.w);
      }
      if (out_hi != nullptr) {
        const uint2 o = make_uint2(pack_bf16x2(g.x, g.y), pack_bf16x2(g.z, g.w));
        *reinterpret_cast<uint2*>(out_hi + off) = o;
        // the bias gradient is the column sum of what the wgrad GEMM will see (the rounded values)
        g = make_float4(bf16_lo_to_f32(o.x), bf16_hi_to_f32(o.x), bf16_lo_to_f32(o.y), bf16_hi_to_f32(o.y));
      }
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
  }
  if (colsum != nullptr) {
    atomicAdd(colsum + c + 0, acc.x); atomicAdd(colsum + c + 1, acc.y);
    atomicAdd(colsum + c + 2, acc.z); atomicAdd(colsum + c + 3, acc.w);
  }
}

// ------------------------------------------------------------------------------------ dropout (forward and backward)
// out = (resid ? resid : 0) + keep(x) * x / (1 - p)  on fp32, optional bf16 hi copy; in place (out == x) is allowed.
// Used for tf.keras.layers.Dropout at feature_extractor.py:95, encoder.py:118,270, modeling.py:253 and for their gradients.
__global__ void __launch_bounds__(256)
dropout_rows_kernel(const float* __restrict__ x, const float* __restrict__ resid, size_t n4, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_hi, DropSpec dr) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    const uint64_t bits = drop_bits4(dr, i);
    v.x = drop_keep(bits, 0, dr.thr16) ? v.x * dr.scale : 0.0f;
    v.y = drop_keep(bits, 1, dr.thr16) ? v.y * dr.scale : 0.0f;
    v.z = drop_keep(bits, 2, dr.thr16) ? v.z * dr.scale : 0.0f;
    v.w = drop_keep(bits, 3, dr.thr16) ? v.w * dr.scale : 0.0f;
    if (resid != nullptr) {
      const float4 r = reinterpret_cast<const float4*>(resid)[i];
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32)[i] = v;
    if (out_hi != nullptr) reinterpret_cast<uint2*>(out_hi)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}
// the keep mask itself (1 = kept), for tests: elementwise sites use index = flat element index
__global__ void dropout_mask_kernel(size_t n, uint8_t* __restrict__ out, DropSpec dr) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = drop_keep(drop_bits4(dr, i >> 2), (int)(i & 3), dr.thr16) ? 1 : 0;
}

// keep mask of the attention-probability dropout, out[bh][q][k]
__global__ void attn_dropout_mask_kernel(int BH, int T, uint8_t* __restrict__ out, DropSpec dr) {
  const size_t n = (size_t)BH * T * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % T);
    const size_t row = i / T;
    const int q = (int)(row % T), bh = (int)(row / T);
    out[i] = drop_keep(drop_bits4(dr, attn_row_group(bh, q, T) + (k >> 2)), k & 3, dr.thr16) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------ bf16 transpose with zero padding
// out[n][m] = in[m][n] for m < rows, 0 for rows <= m < out_ld.  64 x 64 tiles through shared memory.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int rows, int cols, __nv_bfloat16* __restrict__ out, int out_ld) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty + 8 * i;
    const int n = n0 + 2 * tx;
    uint32_t v = 0u;
    if (m < rows && n < cols) v = __ldg(reinterpret_cast<const uint32_t*>(in + (size_t)m * cols + n));   // cols is even
    *reinterpret_cast<uint32_t*>(&tile[ty + 8 * i][2 * tx]) = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty + 8 * i;
    const int m = m0 + 2 * tx;
    if (n < cols && m < out_ld) {
      __nv_bfloat162 o;
      o.x = tile[2 * tx][ty + 8 * i];
      o.y = tile[2 * tx + 1][ty + 8 * i];
      *reinterpret_cast<__nv_bfloat162*>(out + (size_t)n * out_ld + m) = o;
    }
  }
}

// ------------------------------------------------------------------------------------ lm_head dgrad (fp32, exact)
// out[m][n] = sum_v g[m][v] * kernel[n][v];  kernel = TF Dense kernel [hidden][vocab], vocab <= 64.
__global__ void __launch_bounds__(256)
lm_head_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ kernel, int rows, int d, int V,
                     float* __restrict__ out) {
  __shared__ float sg[8][64];
  const int m0 = blockIdx.x * 8;
  for (int i = threadIdx.x; i < 8 * V; i += 256) {
    const int r = i / V, v = i - r * V;
    sg[r][v] = (m0 + r < rows) ? __ldg(g + (size_t)(m0 + r) * V + v) : 0.0f;
  }
  __syncthreads();
  for (int n = threadIdx.x; n < d; n += 256) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.0f;
    for (int v = 0; v < V; ++v) {
      const float w = __ldg(kernel + (size_t)n * V + v);
#pragma unroll
      for (int r = 0; r < 8; ++r) acc[r] = fmaf(sg[r][v], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (m0 + r < rows) out[(size_t)(m0 + r) * d + n] = acc[r];
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_ln_bwd(const float* x, const float* gamma, const float* dy, float eps, int64_t rows, int d,
                           float* dx_f32, void* dx_hi, float* dgamma, float* dbeta, float* colsum, void* stream) {
  W2V2_CHECK_ARG(x && gamma && dy, "null pointer");
  W2V2_CHECK_ARG(d > 0 && d % 4 == 0 && d <= 1024, "d must be a multiple of 4, at most 1024");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // every CTA ends in up to 3 d global atomics and same-address atomics serialise in L2: no more CTAs than SMs
  // (W2V2_LNBWD_CTAS overrides, A/B switch)
  static const int max_ctas = [] { const char* e = getenv("W2V2_LNBWD_CTAS"); const int v = e ? atoi(e) : 148; return v > 0 ? v : 148; }();
  int grid = (int)((rows + 7) / 8);
  if (grid > max_ctas) grid = max_ctas;
  auto* hi = reinterpret_cast<__nv_bfloat16*>(dx_hi);
  const size_t sm = 3 * d * sizeof(float);
  if (d == 512) ln_bwd_kernel<4, true><<<grid, 256, sm, s>>>(x, gamma, dy, eps, (int)rows, d, dx_f32, hi, dgamma, dbeta, colsum);
  else if (d == 768) ln_bwd_kernel<6, true><<<grid, 256, sm, s>>>(x, gamma, dy, eps, (int)rows, d, dx_f32, hi, dgamma, dbeta, colsum);
  else if (d == 1024) ln_bwd_kernel<8, true><<<grid, 256, sm, s>>>(x, gamma, dy, eps, (int)rows, d, dx_f32, hi, dgamma, dbeta, colsum);
  else ln_bwd_kernel<8><<<grid, 256, sm, s>>>(x, gamma, dy, eps, (int)rows, d, dx_f32, hi, dgamma, dbeta, colsum);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_gelu_rows(const float* pre, int64_t n, int fast, void* out_hi, void* out_lo, float drop_p,
                              uint64_t seed, uint32_t site, void* stream) {
  W2V2_CHECK_ARG(pre && out_hi, "null pointer");
  W2V2_CHECK_ARG(n % 4 == 0, "element count must be a multiple of 4");
  if (n <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t n4 = (size_t)n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  auto* hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  auto* lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  const DropSpec dr = make_drop(drop_p, seed, site);
  if (fast) gelu_rows_kernel<true><<<grid, 256, 0, s>>>(pre, n4, hi, lo, dr);
  else gelu_rows_kernel<false><<<grid, 256, 0, s>>>(pre, n4, hi, lo, dr);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

// Same operation for cols % 8 == 0.  The one-row-per-iteration kernel above ends in `cols` fp32 atomics per 16-row CTA: 384 adds
// onto every address at 6144 rows, and same-address L2 atomics serialise (~20 of the 25 us of a 9 MB column sum).  Here a CTA is
// 32 column threads (8 columns = one 16-byte bf16 load each: 256 columns) x 8 row lanes, walks `rows_per_cta` rows with U rows
// per lane in flight, reduces its 8 lanes through shared memory and issues ONE atomic per column.
__global__ void __launch_bounds__(256)
dact_colsum8_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ pre, int rows, int cols, int rows_per_cta,
                    __nv_bfloat16* __restrict__ out_hi, float* __restrict__ colsum, DropSpec dr) {
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 8;
  const bool col_ok = c < cols;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  constexpr int U = 4;
  if (col_ok) {
    for (int rb = r0 + ty; rb < r1; rb += 8 * U) {
      uint4 raw[U];
      float4 p0[U], p1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t off = (size_t)min(rb + 8 * u, r1 - 1) * cols + c;    // clamped: a tail lane re-reads the last row, result discarded
        raw[u] = __ldg(reinterpret_cast<const uint4*>(dy + off));
        if (pre != nullptr) {
          p0[u] = __ldg(reinterpret_cast<const float4*>(pre + off));
          p1[u] = __ldg(reinterpret_cast<const float4*>(pre + off + 4));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = rb + 8 * u;
        if (r >= r1) break;
        const size_t off = (size_t)r * cols + c;
        const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
        float g[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          g[2 * i] = bf16_lo_to_f32(w[i]);
          g[2 * i + 1] = bf16_hi_to_f32(w[i]);
        }
        if (dr.thr16) {   // gradient of a dropout that sat AFTER the activation (or after a Dense when pre == null)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t bits = drop_bits4(dr, (off >> 2) + h);
#pragma unroll
            for (int e = 0; e < 4; ++e) g[4 * h + e] = drop_keep(bits, e, dr.thr16) ? g[4 * h + e] * dr.scale : 0.0f;
          }
        }
        if (pre != nullptr) {
          g[0] *= gelu_grad(p0[u].x); g[1] *= gelu_grad(p0[u].y); g[2] *= gelu_grad(p0[u].z); g[3] *= gelu_grad(p0[u].w);
          g[4] *= gelu_grad(p1[u].x); g[5] *= gelu_grad(p1[u].y); g[6] *= gelu_grad(p1[u].z); g[7] *= gelu_grad(p1[u].w);
        }
        if (out_hi != nullptr) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i] = pack_bf16x2(g[2 * i], g[2 * i + 1]);
            // the bias gradient is the column sum of what the wgrad GEMM will see (the rounded values)
            g[2 * i] = bf16_lo_to_f32(o[i]);
            g[2 * i + 1] = bf16_hi_to_f32(o[i]);
          }
          *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(o[0], o[1], o[2], o[3]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += g[i];
      }
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][8 * tx + i] = acc[i];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float s_ = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s_ += red[l][threadIdx.x];
    atomicAdd(colsum + cc, s_);
  }
}

extern "C" int w2v2_dact_colsum(const void* dy_hi, const float* pre, int64_t rows, int cols, void* out_hi, float* colsum,
                                float drop_p, uint64_t seed, uint32_t site, void* stream) {
  W2V2_CHECK_ARG(dy_hi && (out_hi || colsum), "null pointer");
  W2V2_CHECK_ARG(cols > 0 && cols % 4 == 0, "cols must be a multiple of 4");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cols % 8 == 0) {
    // about two CTAs per SM, as few CTA rows as that allows: every CTA row adds once onto each column's address
    const int gx8 = (cols + 255) / 256;
    int gy8 = (296 + gx8 - 1) / gx8;
    const int max_gy = (int)((rows + 31) / 32);
    if (gy8 > max_gy) gy8 = max_gy;
    const int rpc = (int)((rows + gy8 - 1) / gy8);
    gy8 = (int)((rows + rpc - 1) / rpc);
    dact_colsum8_kernel<<<dim3(gx8, gy8), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy_hi), pre, (int)rows, cols, rpc,
                                                       reinterpret_cast<__nv_bfloat16*>(out_hi), colsum, make_drop(drop_p, seed, site));
    W2V2_CUDA(cudaGetLastError());
    return 0;
  }
  const int gx = (cols / 4 + 255) / 256;
  int gy = (int)((rows + 15) / 16);   // short dependent-load chains: many row chunks, 4 atomics per thread at the end
  if (gy > 1024) gy = 1024;
  dact_colsum_kernel<<<dim3(gx, gy), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(dy_hi), pre, (int)rows, cols,
                                                  reinterpret_cast<__nv_bfloat16*>(out_hi), colsum, make_drop(drop_p, seed, site));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_transpose_bf16(const void* in, int64_t rows, int cols, void* out, int64_t out_ld, void* stream) {
  W2V2_CHECK_ARG(in && out, "null pointer");
  W2V2_CHECK_ARG(cols > 0 && cols % 2 == 0 && out_ld >= rows && out_ld % 2 == 0, "cols and out_ld must be even, out_ld >= rows");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((out_ld + 63) / 64), (unsigned)((cols + 63) / 64));
  transpose_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in), (int)rows, cols,
                                             reinterpret_cast<__nv_bfloat16*>(out), (int)out_ld);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_lm_head_dgrad(const float* grad_logits, const float* kernel, int64_t rows, int hidden_size, int vocab,
                                  float* out, void* stream) {
  W2V2_CHECK_ARG(grad_logits && kernel && out, "null pointer");
  W2V2_CHECK_ARG(vocab > 0 && vocab <= 64 && hidden_size > 0, "vocab must be in 1..64");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  lm_head_dgrad_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(grad_logits, kernel, (int)rows, hidden_size, vocab, out);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_dropout_rows(const float* x, const float* resid, int64_t n, float drop_p, uint64_t seed, uint32_t site,
                                 float* out_f32, void* out_hi, void* stream) {
  W2V2_CHECK_ARG(x && (out_f32 || out_hi), "null pointer");
  W2V2_CHECK_ARG(n % 4 == 0, "element count must be a multiple of 4");
  W2V2_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "drop_p must be in [0, 1)");
  if (n <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t n4 = (size_t)n / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  dropout_rows_kernel<<<grid, 256, 0, s>>>(x, resid, n4, out_f32, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                           make_drop(drop_p, seed, site));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_attn_dropout_mask(int batch_heads, int frames, float drop_p, uint64_t seed, uint32_t site, uint8_t* out,
                                      void* stream) {
  W2V2_CHECK_ARG(out != nullptr && batch_heads > 0 && frames > 0, "bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  attn_dropout_mask_kernel<<<148 * 8, 256, 0, s>>>(batch_heads, frames, out, make_drop(drop_p, seed, site));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_dropout_mask(int64_t n, float drop_p, uint64_t seed, uint32_t site, uint8_t* out, void* stream) {
  W2V2_CHECK_ARG(out != nullptr, "null pointer");
  if (n <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  dropout_mask_kernel<<<grid, 256, 0, s>>>((size_t)n, out, make_drop(drop_p, seed, site));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
