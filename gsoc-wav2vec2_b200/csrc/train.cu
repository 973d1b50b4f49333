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
