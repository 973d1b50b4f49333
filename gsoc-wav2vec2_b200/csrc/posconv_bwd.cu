// Weight gradient of the grouped positional convolution (forward: posconv.cu; reference PositionalConvEmbedding.call
// encoder.py:177-181, differentiated by Keras `fit` in the stage-2 fine-tune, main.py:234-250):
//     dW[j][ci][g*cpg + co] = sum_{b,t} x[b, t + j - k/2, g*cpg + ci] * dpre[b, t, g*cpg + co]
// in the layout of the (weight-normalised) TF kernel [k, cin/groups, cout].  The input gradient reuses the forward kernel
// (w2v2_posconv with linear = 1, shift = 1 and flipped / transposed taps); the weight-norm chain rule
// (tensorflow_addons.py:16-21) is parameter-sized algebra done by the host.
//
// One CTA = 4 taps of one group (one tap per warp); it walks all (b, t) in 64-frame chunks: the x window
// (64 + 3 rows) and the dpre chunk sit in shared memory, tap j's A operand is the window shifted by j rows, and a
// cpg x cpg fp32 accumulator per warp lives in registers (warp-level mma.sync m16n8k16, A and B both loaded with
// ldmatrix.trans because the reduction index t is the ROW index of both tiles).  No atomics: every (tap, group) slice of
// dW has exactly one owner.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int PW_TAPS = 4;
constexpr int PW_CHUNK = 64;

template <int CPG>
__global__ void __launch_bounds__(128)
posconv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dpre, int B, int T, int d,
                     int ktaps, float* __restrict__ dW) {
  constexpr int LD = CPG + 8;
  constexpr int MT = CPG / 16, NT = CPG / 8;
  constexpr int WROWS = PW_CHUNK + PW_TAPS;   // window rows (taps j0 .. j0 + 3)
  __shared__ __align__(16) __nv_bfloat16 sx[WROWS * LD], sg[PW_CHUNK * LD];
  const int j0 = blockIdx.x * PW_TAPS, g = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, q = lane & 3;
  float acc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.0f;
  constexpr int PIECES = CPG / 8;   // 16-byte pieces per row
  for (int b = 0; b < B; ++b) {
    const __nv_bfloat16* xb = x + (size_t)b * T * d + g * CPG;
    const __nv_bfloat16* gb = dpre + (size_t)b * T * d + g * CPG;
    for (int t0 = 0; t0 < T; t0 += PW_CHUNK) {
      __syncthreads();
      for (int i = tid; i < WROWS * PIECES; i += 128) {
        const int r = i / PIECES, p = i - r * PIECES;
        const int t = t0 + j0 - ktaps / 2 + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t >= 0 && t < T) v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)t * d) + p);
        *reinterpret_cast<uint4*>(sx + r * LD + 8 * p) = v;
      }
      for (int i = tid; i < PW_CHUNK * PIECES; i += 128) {
        const int r = i / PIECES, p = i - r * PIECES;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (t0 + r < T) v = __ldg(reinterpret_cast<const uint4*>(gb + (size_t)(t0 + r) * d) + p);
        *reinterpret_cast<uint4*>(sg + r * LD + 8 * p) = v;
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < PW_CHUNK / 16; ++ks) {
        uint32_t bf[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint32_t addr = smem_u32(sg + (16 * ks + (lane & 15)) * LD + 8 * n);
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(bf[n][0]), "=r"(bf[n][1]) : "r"(addr));
        }
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          // A[ci][tt] = window[tt + warp][ci]: matrices (mi, ki) = (0,0) (1,0) (0,1) (1,1) for lanes 0-7, 8-15, 16-23, 24-31
          uint32_t a[4];
          const int mat = lane >> 3, r = lane & 7;
          const uint32_t addr = smem_u32(sx + (16 * ks + 8 * (mat >> 1) + r + warp) * LD + 16 * m + 8 * (mat & 1));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
#pragma unroll
          for (int n = 0; n < NT; ++n) {
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+f"(acc[m][n][0]), "+f"(acc[m][n][1]), "+f"(acc[m][n][2]), "+f"(acc[m][n][3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bf[n][0]), "r"(bf[n][1]));
          }
        }
      }
    }
  }
  const int j = j0 + warp;
  if (j < ktaps) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int ci = 16 * m + gq + 8 * r, co = 8 * n + 2 * q;
          *reinterpret_cast<float2*>(dW + ((size_t)j * CPG + ci) * d + g * CPG + co) =
              make_float2(acc[m][n][2 * r], acc[m][n][2 * r + 1]);
        }
  }
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int w2v2_posconv_wgrad(const void* x_hi, const void* dpre_hi, int batch, int frames, int hidden, int groups,
                                  int ktaps, float* grad_kernel, void* stream) {
  W2V2_CHECK_ARG(x_hi && dpre_hi && grad_kernel, "null pointer");
  W2V2_CHECK_ARG(groups > 0 && hidden % groups == 0, "hidden must be divisible by groups");
  const int cpg = hidden / groups;
  W2V2_CHECK_ARG(cpg == 48 || cpg == 64, "channels per group must be 48 or 64");
  W2V2_CHECK_ARG(ktaps > 0 && ktaps % PW_TAPS == 0, "ktaps must be a multiple of 4");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ktaps / PW_TAPS, groups);
  auto* xp = reinterpret_cast<const __nv_bfloat16*>(x_hi);
  auto* gp = reinterpret_cast<const __nv_bfloat16*>(dpre_hi);
  if (cpg == 48) posconv_wgrad_kernel<48><<<grid, 128, 0, s>>>(xp, gp, batch, frames, hidden, ktaps, grad_kernel);
  else posconv_wgrad_kernel<64><<<grid, 128, 0, s>>>(xp, gp, batch, frames, hidden, ktaps, grad_kernel);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
