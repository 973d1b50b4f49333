// Kernels of the stage-1 fine-tune step (reference: src/main.py:210-223 - the Wav2Vec2 body is frozen, only the CTC
// head trains): weight gradient of lm_head and a fused Keras-Adam update on a flat fp32 buffer.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

// dW[k][n] += sum_m h[m][k] * g[m][n],  db[n] += sum_m g[m][n]   (Dense kernel layout [in, out], modeling.py:231)
// grid = (ceil(d / 64), row_splits); block = 256 = 64 k-lanes x 4 n-groups of V/4 columns (V <= 128).
template <int NPT>
__global__ void __launch_bounds__(256)
lm_head_wgrad_kernel(const float* __restrict__ h, const float* __restrict__ g, int M, int d, int V,
                     float* __restrict__ dW, float* __restrict__ db) {
  const int k = blockIdx.x * 64 + (threadIdx.x & 63);
  const int ng = threadIdx.x >> 6;           // 0..3
  const int n_begin = ng * NPT;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc[NPT];
#pragma unroll
  for (int j = 0; j < NPT; ++j) acc[j] = 0.0f;
  float bacc = 0.0f;  // bias gradient: lanes with k == first k of block 0 accumulate column (n_begin + lane%NPT)
  for (int m = m0; m < m1; ++m) {
    const float hv = (k < d) ? __ldg(h + (size_t)m * d + k) : 0.0f;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = n_begin + j;
      const float gv = (n < V) ? __ldg(g + (size_t)m * V + n) : 0.0f;   // warp-uniform address: broadcast
      acc[j] = fmaf(hv, gv, acc[j]);
    }
    if (blockIdx.x == 0 && (threadIdx.x & 63) < NPT) {
      const int n = n_begin + (threadIdx.x & 63);
      if (n < V) bacc += __ldg(g + (size_t)m * V + n);
    }
  }
  if (k < d) {
#pragma unroll
    for (int j = 0; j < NPT; ++j)
      if (n_begin + j < V) atomicAdd(dW + (size_t)k * V + n_begin + j, acc[j]);
  }
  if (blockIdx.x == 0 && (threadIdx.x & 63) < NPT) {
    const int n = n_begin + (threadIdx.x & 63);
    if (n < V) atomicAdd(db + n, bacc);
  }
}

// Keras Adam (non-amsgrad): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t * m / (sqrt(v) + eps),
// lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t) is computed by the host.
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2, float eps) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_lm_head_wgrad(const float* hidden, const float* grad_logits, int64_t rows, int hidden_size,
                                  int vocab, float* grad_kernel, float* grad_bias, void* stream) {
  W2V2_CHECK_ARG(hidden && grad_logits && grad_kernel && grad_bias, "null pointer");
  W2V2_CHECK_ARG(rows > 0 && hidden_size > 0 && vocab > 0 && vocab <= 128, "need rows, hidden_size > 0 and 0 < vocab <= 128");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  W2V2_CUDA(cudaMemsetAsync(grad_kernel, 0, sizeof(float) * (size_t)hidden_size * vocab, s));
  W2V2_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * vocab, s));
  const int splits = (int)((rows + 255) / 256 < 64 ? (rows + 255) / 256 : 64);
  dim3 grid((hidden_size + 63) / 64, splits);
  const int npt = (vocab + 3) / 4;
  if (npt <= 8) lm_head_wgrad_kernel<8><<<grid, 256, 0, s>>>(hidden, grad_logits, (int)rows, hidden_size, vocab, grad_kernel, grad_bias);
  else if (npt <= 16) lm_head_wgrad_kernel<16><<<grid, 256, 0, s>>>(hidden, grad_logits, (int)rows, hidden_size, vocab, grad_kernel, grad_bias);
  else lm_head_wgrad_kernel<32><<<grid, 256, 0, s>>>(hidden, grad_logits, (int)rows, hidden_size, vocab, grad_kernel, grad_bias);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_adam(float* weights, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1,
                         float beta2, float eps, void* stream) {
  W2V2_CHECK_ARG(weights && grads && m && v, "null pointer");
  if (n <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  adam_kernel<<<grid, 256, 0, s>>>(weights, grads, m, v, (size_t)n, lr_t, beta1, beta2, eps);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ batched weight re-packing
// After every optimizer step the fp32 master weights (one flat buffer) must be re-expressed in the kernels' operand
// layouts: forward GEMMs want W[out][in] bf16 (the transpose of the TF Dense kernel, q rows pre-scaled by dh^-1/2,
// q/k/v fused), the dgrad GEMMs want the TF kernel itself as bf16 (q/k/v side by side).  Done tensor by tensor from the
// host this is ~270 tiny launches (1.5 ms per step); here it is ONE launch over a job table that never changes
// (w2v2_pack_weights).  A job copies a [rows x cols] fp32 matrix to a bf16 (or fp32) destination with its own leading
// dimension, optionally transposed, optionally scaled; tiles of 64 x 64 go through shared memory.
namespace w2v2 {

struct PackTile {
  int job, r0, c0;
};

__global__ void __launch_bounds__(256)
pack_weights_kernel(const w2v2_pack_job* __restrict__ jobs, const PackTile* __restrict__ tiles) {
  __shared__ float tile[64][65];
  const PackTile t = tiles[blockIdx.x];
  const w2v2_pack_job j = jobs[t.job];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  const float* src = reinterpret_cast<const float*>(j.src);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = t.r0 + ty + 4 * i, c = t.c0 + tx;
    tile[ty + 4 * i][tx] = (r < j.rows && c < j.cols) ? src[(size_t)r * j.src_ld + c] * j.scale : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // destination element (dr, dc): transposed jobs write dst[c][r], plain jobs dst[r][c]; tx runs along the dst row
    int sr, sc, dr, dc;
    if (j.transpose) {
      sc = ty + 4 * i; sr = tx;                // dst row = source column
      dr = t.c0 + sc; dc = t.r0 + sr;
      if (dr >= j.cols || dc >= j.rows) continue;
    } else {
      sr = ty + 4 * i; sc = tx;
      dr = t.r0 + sr; dc = t.c0 + sc;
      if (dr >= j.rows || dc >= j.cols) continue;
    }
    const float v = tile[sr][sc];
    const size_t o = (size_t)dr * j.dst_ld + dc;
    if (j.dst_f32) reinterpret_cast<float*>(j.dst)[o] = v;
    else reinterpret_cast<__nv_bfloat16*>(j.dst)[o] = __float2bfloat16_rn(v);
  }
}

}  // namespace w2v2

extern "C" int w2v2_pack_weights(const w2v2_pack_job* jobs_dev, const int32_t* tiles_dev, int num_tiles, void* stream) {
  W2V2_CHECK_ARG(jobs_dev && tiles_dev, "null pointer");
  if (num_tiles <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  pack_weights_kernel<<<num_tiles, 256, 0, s>>>(jobs_dev, reinterpret_cast<const PackTile*>(tiles_dev));
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
