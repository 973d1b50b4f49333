// CTC loss (forward + gradient w.r.t. the unnormalised logits) and greedy frame argmax, sm_100a.
//
// Reference: CTCLoss.call, losses.py:14-45 -> tf.nn.ctc_loss(labels, logits, label_length, logit_length,
// logits_time_major=False, blank_index=pad_id), Keras SUM reduction (losses.py:6), / division_factor
// (losses.py:45).  Conventions kept: blank = pad_id; label_length = number of non-pad labels
// (losses.py:32-33, the first label_length dense entries are used); logit_length = T for every sample
// (losses.py:29-30 - padded frames are NOT excluded).
//
// One CTA per utterance.  The two time recursions are the only sequential part (latency-bound: one block barrier per
// frame), so they run CONCURRENTLY: threads 0-511 sweep alpha forwards, threads 512-1023 sweep beta backwards, each half with
// its own named barrier and one thread per extended-label state s (S = 2*len+1); both tables go to a caller-provided
// workspace.  The gradient  d loss / d logit[t,k] = softmax[t,k] - occupancy[t,k]  is then embarrassingly parallel: one warp
// per frame combines alpha_t(s) + beta_t(s).  (First version: alpha sweep, then a beta sweep with three barriers per frame
// and the gradient inside the loop - 1.4 ms at T = 768; this one 0.4 ms.)
#include <math.h>

#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int CTC_THREADS = 1024;  // two groups of 512: alpha sweep / beta sweep
constexpr int CTC_GROUP = 512;

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(fminf(a, b) - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_loss_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int T, int V, int Lmax, int blank,
                float scale, float* __restrict__ ws, float* __restrict__ loss_out, float* __restrict__ grad,
                int lp_in_smem) {
  extern __shared__ float dyn[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int Smax = 2 * Lmax + 1;
  // carve: ext[Smax] (int), 4 x buf[Smax] (alpha / beta ping-pong), occ[32 warps][V], lse[T], then lp[T*V] when it fits
  int* ext = reinterpret_cast<int*>(dyn);
  float* bufs = dyn + Smax;
  float* occ_all = bufs + 4 * Smax;
  float* lse = occ_all + 32 * V;
  float* lp = lse + T;
  __shared__ int s_len;
  __shared__ float s_logp;

  const float* lg = logits + (size_t)b * T * V;
  const int* lab = labels + (size_t)b * Lmax;
  if (tid == 0) s_len = 0;
  __syncthreads();
  int cnt = 0;
  for (int i = tid; i < Lmax; i += blockDim.x) cnt += (lab[i] != blank);
  if (cnt) atomicAdd(&s_len, cnt);
  __syncthreads();
  const int len = s_len;
  const int S = 2 * len + 1;
  for (int s = tid; s < S; s += blockDim.x) ext[s] = (s & 1) ? lab[s >> 1] : blank;
  // log-softmax normaliser per frame (one warp per frame)
  for (int t = tid >> 5; t < T; t += blockDim.x >> 5) {
    float m = -INFINITY;
    for (int k = lane_id(); k < V; k += 32) m = fmaxf(m, lg[(size_t)t * V + k]);
    m = warp_max(m);
    float e = 0.0f;
    for (int k = lane_id(); k < V; k += 32) e += expf(lg[(size_t)t * V + k] - m);
    e = warp_sum(e);
    if (lane_id() == 0) lse[t] = m + logf(e);
  }
  __syncthreads();
  if (lp_in_smem)
    for (int i = tid; i < T * V; i += blockDim.x) lp[i] = lg[i] - lse[i / V];
  __syncthreads();
  auto LP = [&](int t, int k) -> float { return lp_in_smem ? lp[t * V + k] : lg[(size_t)t * V + k] - lse[t]; };

  float* aw = ws + (size_t)b * 2 * T * Smax;       // alpha_t(s)
  float* bw = aw + (size_t)T * Smax;               // beta_t(s), including the emission at t (Graves' convention)
  const int group = tid / CTC_GROUP, gt = tid - group * CTC_GROUP;
  const bool want_beta = grad != nullptr;
  if (group == 0) {
    // ---- alpha sweep (threads 0..511, named barrier 1)
    float* prev = bufs;
    float* cur = bufs + Smax;
    for (int s = gt; s < S; s += CTC_GROUP) {
      const float a = (s < 2) ? LP(0, ext[s]) : -INFINITY;
      prev[s] = a;
      aw[s] = a;
    }
    asm volatile("bar.sync 1, 512;" ::: "memory");
    for (int t = 1; t < T; ++t) {
      for (int s = gt; s < S; s += CTC_GROUP) {
        const int e = ext[s];
        const float a0 = prev[s];
        const float a1 = (s >= 1) ? prev[s - 1] : -INFINITY;
        const float a2 = (s >= 2 && e != blank && e != ext[s - 2]) ? prev[s - 2] : -INFINITY;
        const float a = lse3(a0, a1, a2) + LP(t, e);
        cur[s] = a;
        aw[(size_t)t * Smax + s] = a;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      float* tmp = prev;
      prev = cur;
      cur = tmp;
    }
    if (gt == 0) {
      const float logp = (S > 1) ? lse2(prev[S - 1], prev[S - 2]) : prev[S - 1];
      s_logp = logp;
      loss_out[b] = -logp * scale;
    }
  } else if (want_beta) {
    // ---- beta sweep (threads 512..1023, named barrier 2), concurrent with the alpha sweep
    float* prev = bufs + 2 * Smax;   // beta at t
    float* cur = bufs + 3 * Smax;    // beta at t - 1
    for (int s = gt; s < S; s += CTC_GROUP) {
      const float v = (s >= S - 2) ? LP(T - 1, ext[s]) : -INFINITY;
      prev[s] = v;
      bw[(size_t)(T - 1) * Smax + s] = v;
    }
    asm volatile("bar.sync 2, 512;" ::: "memory");
    for (int t = T - 1; t > 0; --t) {
      for (int s = gt; s < S; s += CTC_GROUP) {
        const int e = ext[s];
        const float b0 = prev[s];
        const float b1 = (s + 1 < S) ? prev[s + 1] : -INFINITY;
        const float b2 = (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != e) ? prev[s + 2] : -INFINITY;
        const float v = lse3(b0, b1, b2) + LP(t - 1, e);
        cur[s] = v;
        bw[(size_t)(t - 1) * Smax + s] = v;
      }
      asm volatile("bar.sync 2, 512;" ::: "memory");
      float* tmp = prev;
      prev = cur;
      cur = tmp;
    }
  }
  __syncthreads();                  // both tables complete (block-scope visibility of the global writes), s_logp set
  if (!want_beta) return;
  const float logp = s_logp;
  float* gr = grad + (size_t)b * T * V;
  // ---- gradient: one warp per frame, occupancy[k] = sum over states with label k of exp(alpha + beta - lp - logp)
  const int warp = tid >> 5, lane = lane_id();
  float* occ = occ_all + warp * V;
  for (int t = warp; t < T; t += CTC_THREADS / 32) {
    for (int k = lane; k < V; k += 32) occ[k] = 0.0f;
    __syncwarp();
    float wb = 0.0f;                // blank states reduce in registers first
    for (int s = lane; s < S; s += 32) {
      const int e = ext[s];
      const float x = aw[(size_t)t * Smax + s] + bw[(size_t)t * Smax + s] - LP(t, e) - logp;
      const float w = (x == -INFINITY) ? 0.0f : expf(x);
      if (s & 1) {
        if (w != 0.0f) atomicAdd(&occ[e], w);
      } else {
        wb += w;
      }
    }
    wb = warp_sum(wb);
    if (lane == 0 && wb != 0.0f) atomicAdd(&occ[blank], wb);
    __syncwarp();
    for (int k = lane; k < V; k += 32) gr[(size_t)t * V + k] = (expf(LP(t, k)) - occ[k]) * scale;
    __syncwarp();
  }
}

__global__ void frame_argmax_kernel(const float* __restrict__ logits, long rows, int V, int* __restrict__ ids) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = logits + r * V;
  float best = p[0];
  int arg = 0;
  for (int k = 1; k < V; ++k) {
    const float v = p[k];
    if (v > best) {  // first maximum wins, like np/tf argmax
      best = v;
      arg = k;
    }
  }
  ids[r] = arg;
}

}  // namespace w2v2

using namespace w2v2;

extern "C" int64_t w2v2_ctc_workspace_bytes(int batch, int frames, int max_label_len) {
  return (int64_t)2 * batch * frames * (2 * (int64_t)max_label_len + 1) * (int64_t)sizeof(float);   // alpha and beta tables
}

extern "C" int w2v2_ctc_loss(const float* logits, const int32_t* labels, int batch, int frames, int vocab,
                             int max_label_len, int blank, float scale, void* workspace, int64_t* reserved,
                             float* loss_per_sample, float* grad_logits, void* stream) {
  (void)reserved;
  W2V2_CHECK_ARG(logits && labels && workspace && loss_per_sample, "null pointer");
  W2V2_CHECK_ARG(batch > 0 && frames > 0 && vocab > 1 && max_label_len > 0, "sizes must be positive");
  W2V2_CHECK_ARG(blank >= 0 && blank < vocab, "blank index out of range");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int Smax = 2 * max_label_len + 1;
  size_t fixed = (size_t)(5 * Smax + 32 * vocab + frames) * sizeof(float);
  size_t with_lp = fixed + (size_t)frames * vocab * sizeof(float);
  const int lp_in_smem = with_lp <= 200 * 1024 ? 1 : 0;
  const size_t smem = lp_in_smem ? with_lp : fixed;
  W2V2_CHECK_ARG(smem <= 200 * 1024, "sequence too long for the shared-memory plan");
  W2V2_CUDA(cudaFuncSetAttribute(ctc_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_loss_kernel<<<batch, CTC_THREADS, smem, s>>>(logits, labels, frames, vocab, max_label_len, blank, scale,
                                                    reinterpret_cast<float*>(workspace), loss_per_sample, grad_logits,
                                                    lp_in_smem);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int w2v2_frame_argmax(const float* logits, int64_t rows, int vocab, int32_t* ids, void* stream) {
  W2V2_CHECK_ARG(logits && ids && vocab > 0, "bad arguments");
  if (rows <= 0) return 0;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  frame_argmax_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(logits, rows, vocab, ids);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
