// Backward of the attention core (reference forward: TransformerAttention.get_context, encoder.py:34-54; Keras `fit`
// differentiates it in the stage-2 fine-tune, main.py:234-250).  Inputs are the tensors the forward already keeps:
// the packed projection qkv[b][t][0:3d] = [q | k | v] (q pre-scaled by dh^-1/2), the context O = softmax(q k^T) v and its
// gradient dO, both [b][t][d] bf16.  Output dqkv[b][t][0:3d] in the same packed layout (the q part multiplied by `q_scale`,
// so that every downstream product uses the UNSCALED projection kernels).
//
//     P = softmax(S), S = Q K^T;  dV = P^T dO;  dP = dO V^T;  D_i = sum_j P_ij dP_ij;  dS = P o (dP - D);
//     dQ = dS K;  dK = dS^T Q
// D is NOT taken from the stored context as sum_c dO_ic O_ic (the FlashAttention shortcut): O was rounded to bf16 by the forward
// and formed from bf16-rounded probabilities, so that D misses sum_j P_ij dP_ij of THIS pass by ~2^-9 |dO||O|, a per-row error
// eps_i that leaks into dQ_i as eps_i * (P-weighted mean key) and into dK.  With near-uniform attention over 768 keys the true dS
// is a small difference of large terms and that leak was 5 - 27 x the true dq / dk weight gradients of the upper layers
// (tests/test_full_size_gpu.py).  Sweep 1 therefore also forms dP and accumulates sum_j P_ij dP_ij online: the row sum of dS is
// then zero to fp32 rounding, like the exact softmax Jacobian.
//
// Two kernels, no atomics, deterministic:
//   attn_bwd_dq   : CTA = 64 queries of one (b, h).  Sweep 1 over the key blocks rebuilds the softmax statistics
//                   (running max / sum), sweep 2 forms P, dP, dS and accumulates dQ; it also stores lse2 (log2-domain
//                   log-sum-exp) and D for the second kernel.
//   attn_bwd_dkv  : CTA = 64 keys of one (b, h); loops over the query blocks with the TRANSPOSED products
//                   (S^T = K Q^T, dP^T = V dO^T) so that P^T / dS^T come out of the MMA already in the layout the next
//                   MMA needs as its A operand.
// Tensor cores through warp-level mma.sync m16n8k16 (bf16 in, fp32 accumulate): this kernel is ~2 % of a train step's
// FLOPs budget at 768 frames; a tcgen05 version only matters once the GEMMs around it are at peak.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int AB_BLK = 64;      // queries / keys per tile
constexpr int AB_DH = 64;
constexpr int AB_LD = 72;       // padded row stride (bf16 elements): 144 B, keeps ldmatrix / LDS.32 conflict free
constexpr float AB_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment (16 x 16, row-major tile[m][k])
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __nv_bfloat16* tile, int m0, int k0, int g, int q) {
  a[0] = *reinterpret_cast<const uint32_t*>(tile + (m0 + g) * AB_LD + k0 + 2 * q);
  a[1] = *reinterpret_cast<const uint32_t*>(tile + (m0 + g + 8) * AB_LD + k0 + 2 * q);
  a[2] = *reinterpret_cast<const uint32_t*>(tile + (m0 + g) * AB_LD + k0 + 2 * q + 8);
  a[3] = *reinterpret_cast<const uint32_t*>(tile + (m0 + g + 8) * AB_LD + k0 + 2 * q + 8);
}
// B fragment (k16 x n8) with B[k][n] = tile[n0 + n][k0 + k]   ("tile rows are the n index": Q K^T style products)
__device__ __forceinline__ void load_b_nk(uint32_t& b0, uint32_t& b1, const __nv_bfloat16* tile, int n0, int k0, int g, int q) {
  b0 = *reinterpret_cast<const uint32_t*>(tile + (n0 + g) * AB_LD + k0 + 2 * q);
  b1 = *reinterpret_cast<const uint32_t*>(tile + (n0 + g) * AB_LD + k0 + 2 * q + 8);
}
// B fragment (k16 x n8) with B[k][n] = tile[k0 + k][n0 + n]   ("tile rows are the k index": P V style products)
__device__ __forceinline__ void load_b_kn(uint32_t& b0, uint32_t& b1, const __nv_bfloat16* tile, int k0, int n0, int lane) {
  const uint32_t addr = smem_u32(tile + (k0 + (lane & 15)) * AB_LD + n0);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(addr));
}
// the same two fragment kinds for TWO adjacent n-tiles with one ldmatrix.x4: b[0], b[1] = n-tile n0, b[2], b[3] = n-tile n0 + 8
__device__ __forceinline__ void load_b_nk_x4(uint32_t (&b)[4], const __nv_bfloat16* tile, int n0, int k0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  const uint32_t addr = smem_u32(tile + (n0 + 8 * (mat >> 1) + r) * AB_LD + k0 + 8 * (mat & 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(addr));
}
__device__ __forceinline__ void load_b_kn_x4(uint32_t (&b)[4], const __nv_bfloat16* tile, int k0, int n0, int lane) {
  const int mat = lane >> 3, r = lane & 7;
  const uint32_t addr = smem_u32(tile + (k0 + 8 * (mat & 1) + r) * AB_LD + n0 + 8 * (mat >> 1));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(addr));
}
// copy a [64 x 64] bf16 tile (rows t0.., row stride ld_g elements) into padded smem, zero beyond `t_end`
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int t0, int t_end, size_t ld_g,
                                          int tid, int nthreads) {
  for (int i = tid; i < AB_BLK * 8; i += nthreads) {   // 8 x 16-byte pieces per row
    const int r = i >> 3, p = i & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t0 + r < t_end) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(t0 + r) * ld_g) + p);
    *reinterpret_cast<uint4*>(dst + r * AB_LD + 8 * p) = v;
  }
}

// asynchronous (cp.async, 16 B per request) version: rows >= t_end are zero-filled through src-size 0
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* dst, const __nv_bfloat16* src, int t0, int t_end, size_t ld_g,
                                                int tid, int nthreads) {
  for (int i = tid; i < AB_BLK * 8; i += nthreads) {
    const int r = i >> 3, p = i & 7;
    const bool ok = t0 + r < t_end;
    const __nv_bfloat16* g = src + (ok ? (size_t)(t0 + r) * ld_g + 8 * p : 0);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + r * AB_LD + 8 * p)), "l"(g),
                 "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_f32(float* dst, const float* src, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Dropout keep bits of one [64 queries x 64 keys] tile, computed cooperatively: byte [q][kg] holds the 4 keep bits of keys
// 4 kg .. 4 kg + 3 of query row q (one hash per byte, 8 per thread) - the MMA fragments own 1 or 2 keys of a 4-key group, so
// hashing per fragment element would repeat every draw 2 to 4 times.
__device__ __forceinline__ void fill_keep_tile(uint8_t (*keep)[16], const DropSpec& dr, int bh, int q_first, int k_first, int T,
                                               int tid) {
  for (int i = tid; i < AB_BLK * 16; i += 128) {
    const int ql = i >> 4, kg = i & 15;
    const uint64_t bits = drop_bits4(dr, attn_row_group(bh, q_first + ql, T) + ((k_first >> 2) + kg));
    uint32_t m = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) m |= drop_keep(bits, e, dr.thr16) ? (1u << e) : 0u;
    keep[ql][kg] = (uint8_t)m;
  }
}

// ------------------------------------------------------------------------------------ dQ (+ lse2, D)
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o,
                   const __nv_bfloat16* __restrict__ d_o, const int* __restrict__ kv_len, int T, int H, float q_scale,
                   __nv_bfloat16* __restrict__ dqkv, float* __restrict__ lse2_out, float* __restrict__ dsum_out, DropSpec dr) {
  __shared__ __align__(16) __nv_bfloat16 sQ[AB_BLK * AB_LD], sdO[AB_BLK * AB_LD], sK[AB_BLK * AB_LD], sV[AB_BLK * AB_LD];
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int d = H * AB_DH;
  const int t0 = blockIdx.x * AB_BLK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int klen = kv_len ? min(kv_len[b], T) : T;
  const size_t ld3 = 3 * (size_t)d;
  const __nv_bfloat16* qp = qkv + (size_t)b * T * ld3 + h * AB_DH;
  const __nv_bfloat16* kp = qp + d;
  const __nv_bfloat16* vp = qp + 2 * d;
  const __nv_bfloat16* op = o + (size_t)b * T * d + h * AB_DH;
  const __nv_bfloat16* dop = d_o + (size_t)b * T * d + h * AB_DH;

  load_tile(sQ, qp, t0, T, ld3, tid, 128);
  load_tile(sdO, dop, t0, T, d, tid, 128);
  __syncthreads();
  const int m0 = warp * 16;
  uint32_t aq[4][4], ado[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    load_a(aq[ks], sQ, m0, 16 * ks, g, q);
    load_a(ado[ks], sdO, m0, 16 * ks, g, q);
  }
  (void)op;   // the stored context is not needed: D comes from this pass's own P and dP (see the header)
  float dsum[2] = {0.0f, 0.0f};   // running sum_j exp2(s_ij - mx_i) dP_ij, normalised after sweep 1

  const int nkb = (klen + AB_BLK - 1) / AB_BLK;
  // K / V tiles stream through a 2-deep cp.async pipeline over the 2 * nkb stages of both sweeps (sweep 1 needs K only);
  // the Q / dO tiles are dead once their fragments sit in registers, so they serve as the second pair of buffers
  __syncthreads();
  __nv_bfloat16* const bufK[2] = {sK, sQ};
  __nv_bfloat16* const bufV[2] = {sV, sdO};
  auto issue = [&](int stage) {
    const int kb2 = (stage >= nkb) ? stage - nkb : stage;
    load_tile_async(bufK[stage & 1], kp, kb2 * AB_BLK, klen, ld3, tid, 128);
    load_tile_async(bufV[stage & 1], vp, kb2 * AB_BLK, klen, ld3, tid, 128);
    cp_async_commit();
  };
  issue(0);
  float mx[2] = {-INFINITY, -INFINITY}, sum[2] = {0.0f, 0.0f};
  __shared__ uint8_t s_keep[2][AB_BLK][16];    // dropout keep bits of the current / next tile
  if (dr.thr16) {
    fill_keep_tile(s_keep[0], dr, bh, t0, 0, T, tid);
    __syncthreads();
  }
  // ---- sweep 1: softmax statistics (log2 domain) and D_i = sum_j P_ij dP_ij
  for (int kb = 0; kb < nkb; ++kb) {
    issue(kb + 1);                 // stage nkb (first of sweep 2) always exists
    cp_async_wait<1>();
    __syncthreads();
    const __nv_bfloat16* sK = bufK[kb & 1];
    const __nv_bfloat16* sV = bufV[kb & 1];
    if (dr.thr16) fill_keep_tile(s_keep[(kb + 1) & 1], dr, bh, t0, ((kb + 1 < nkb) ? kb + 1 : 0) * AB_BLK, T, tid);   // next tile (or sweep 2's first)
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = s[j + 1][e] = dp[j][e] = dp[j + 1][e] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        load_b_nk_x4(bb, sK, 8 * j, 16 * ks, lane);
        mma16816(s[j], aq[ks], bb[0], bb[1]);
        mma16816(s[j + 1], aq[ks], bb[2], bb[3]);
        load_b_nk_x4(bb, sV, 8 * j, 16 * ks, lane);
        mma16816(dp[j], ado[ks], bb[0], bb[1]);
        mma16816(dp[j + 1], ado[ks], bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kb * AB_BLK + 8 * j + 2 * q + e;
          float v = s[j][2 * r + e] * AB_LOG2E;
          if (key >= klen) v = -INFINITY;
          s[j][2 * r + e] = v;
          bm = fmaxf(bm, v);
        }
      }
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
      const float nm = fmaxf(mx[r], bm);
      float ps = 0.0f, pd = 0.0f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pe = ex2_approx(s[j][2 * r + e] - nm);     // 0 for masked keys (s = -inf)
          float dpv = dp[j][2 * r + e];
          if (dr.thr16) {   // the same dropped-and-rescaled dP as sweep 2 uses
            const uint32_t kb4 = s_keep[kb & 1][m0 + g + 8 * r][2 * j + (q >> 1)];
            dpv = ((kb4 >> (2 * (q & 1) + e)) & 1u) ? dpv * dr.scale : 0.0f;
          }
          ps += pe;
          pd = fmaf(pe, dpv, pd);
        }
      }
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      pd += __shfl_xor_sync(0xffffffffu, pd, 1);
      pd += __shfl_xor_sync(0xffffffffu, pd, 2);
      const float resc = exp2f(mx[r] - nm);
      sum[r] = sum[r] * resc + ps;
      dsum[r] = dsum[r] * resc + pd;
      mx[r] = nm;
    }
    __syncthreads();               // every warp is done with this buffer before the stage after next refills it
  }
  float lse2[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lse2[r] = mx[r] + log2f(sum[r]);
    dsum[r] = dsum[r] / sum[r];
    const int t = t0 + m0 + g + 8 * r;
    if (q == 0 && t < T) {
      lse2_out[(size_t)bh * T + t] = lse2[r];
      dsum_out[(size_t)bh * T + t] = dsum[r];
    }
  }

  // ---- sweep 2: dQ += (P o (dO V^T - D)) K      (keep bits of its first tile were filled by sweep 1's last iteration)
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.0f;
  for (int kb = 0; kb < nkb; ++kb) {
    const int stage = nkb + kb;
    if (kb + 1 < nkb) {
      issue(stage + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* sK = bufK[stage & 1];
    const __nv_bfloat16* sV = bufV[stage & 1];
    if (dr.thr16 && kb + 1 < nkb) fill_keep_tile(s_keep[(stage + 1) & 1], dr, bh, t0, (kb + 1) * AB_BLK, T, tid);   // visible after the loop's closing barrier
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = s[j + 1][e] = dp[j][e] = dp[j + 1][e] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        load_b_nk_x4(bb, sK, 8 * j, 16 * ks, lane);
        mma16816(s[j], aq[ks], bb[0], bb[1]);
        mma16816(s[j + 1], aq[ks], bb[2], bb[3]);
        load_b_nk_x4(bb, sV, 8 * j, 16 * ks, lane);
        mma16816(dp[j], ado[ks], bb[0], bb[1]);
        mma16816(dp[j + 1], ado[ks], bb[2], bb[3]);
      }
    }
    // dS (bf16) as A fragments: key pairs (2 jj, 2 jj + 1) of n-tiles form one 16-wide k step
    uint32_t ads[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ds[4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int key = kb * AB_BLK + 8 * j + 2 * q + e;
          const float p = (key < klen) ? ex2_approx(s[j][2 * r + e] * AB_LOG2E - lse2[r]) : 0.0f;
          float dpv = dp[j][2 * r + e];
          if (dr.thr16) {   // dP = dP_dropped o mask / (1 - p_drop)
            const uint32_t kb4 = s_keep[stage & 1][m0 + g + 8 * r][2 * j + (q >> 1)];
            dpv = ((kb4 >> (2 * (q & 1) + e)) & 1u) ? dpv * dr.scale : 0.0f;
          }
          ds[2 * r + e] = p * (dpv - dsum[r]);
        }
      }
      ads[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);   // row g
      ads[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);   // row g + 8
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // k = 16 keys
#pragma unroll
      for (int j = 0; j < 8; j += 2) {    // n = 2 x 8 head channels
        uint32_t bb[4];
        load_b_kn_x4(bb, sK, 16 * ks, 8 * j, lane);
        mma16816(dq[j], ads[ks], bb[0], bb[1]);
        mma16816(dq[j + 1], ads[ks], bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = t0 + m0 + g + 8 * r;
    if (t < T) {
      __nv_bfloat16* dst = dqkv + ((size_t)b * T + t) * ld3 + h * AB_DH;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(dst + 8 * j + 2 * q) = pack_bf16x2(dq[j][2 * r] * q_scale, dq[j][2 * r + 1] * q_scale);
    }
  }
}

// ------------------------------------------------------------------------------------ dK, dV
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ d_o,
                    const int* __restrict__ kv_len, int T, int H, const float* __restrict__ lse2_in,
                    const float* __restrict__ dsum_in, __nv_bfloat16* __restrict__ dqkv, DropSpec dr) {
  __shared__ __align__(16) __nv_bfloat16 sQ[AB_BLK * AB_LD], sdO[AB_BLK * AB_LD], sK[AB_BLK * AB_LD], sV[AB_BLK * AB_LD];
  __shared__ float s_lse2[2][AB_BLK], s_ds2[2][AB_BLK];
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int d = H * AB_DH;
  const int k0 = blockIdx.x * AB_BLK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int klen = kv_len ? min(kv_len[b], T) : T;
  const size_t ld3 = 3 * (size_t)d;
  const __nv_bfloat16* qp = qkv + (size_t)b * T * ld3 + h * AB_DH;
  const __nv_bfloat16* kp = qp + d;
  const __nv_bfloat16* vp = qp + 2 * d;
  const __nv_bfloat16* dop = d_o + (size_t)b * T * d + h * AB_DH;

  load_tile(sK, kp, k0, klen, ld3, tid, 128);
  load_tile(sV, vp, k0, klen, ld3, tid, 128);
  __syncthreads();
  const int m0 = warp * 16;   // this warp's 16 keys
  uint32_t ak[4][4], av[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    load_a(ak[ks], sK, m0, 16 * ks, g, q);
    load_a(av[ks], sV, m0, 16 * ks, g, q);
  }
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.0f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.0f;
  }
  const bool key_ok[2] = {k0 + m0 + g < klen, k0 + m0 + g + 8 < klen};
  const int nqb = (T + AB_BLK - 1) / AB_BLK;
  // Q / dO tiles (+ lse2, D) stream through a 2-deep cp.async pipeline; the K / V tiles are dead once their fragments sit
  // in registers, so they serve as the second pair of buffers
  __syncthreads();
  __nv_bfloat16* const bufQ[2] = {sQ, sK};
  __nv_bfloat16* const bufdO[2] = {sdO, sV};
  auto issue = [&](int qb2) {
    load_tile_async(bufQ[qb2 & 1], qp, qb2 * AB_BLK, T, ld3, tid, 128);
    load_tile_async(bufdO[qb2 & 1], dop, qb2 * AB_BLK, T, d, tid, 128);
    const int t = qb2 * AB_BLK + (tid & 63);
    const float* src = (tid < 64 ? lse2_in : dsum_in) + (size_t)bh * T + (t < T ? t : 0);
    cp_async_f32(tid < 64 ? &s_lse2[qb2 & 1][tid] : &s_ds2[qb2 & 1][tid - 64], src, t < T);
    cp_async_commit();
  };
  __shared__ uint8_t s_keep[2][AB_BLK][16];    // dropout keep bits [query][key group] of the current / next tile
  if (dr.thr16) {
    fill_keep_tile(s_keep[0], dr, bh, 0, k0, T, tid);
    __syncthreads();
  }
  issue(0);
  for (int qb = 0; qb < nqb; ++qb) {
    if (qb + 1 < nqb) {
      issue(qb + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* sQ = bufQ[qb & 1];
    const __nv_bfloat16* sdO = bufdO[qb & 1];
    const float* s_lse = s_lse2[qb & 1];
    const float* s_ds = s_ds2[qb & 1];
    if (dr.thr16 && qb + 1 < nqb) fill_keep_tile(s_keep[(qb + 1) & 1], dr, bh, (qb + 1) * AB_BLK, k0, T, tid);
    // S^T = K Q^T and dP^T = V dO^T : [16 keys] x [64 queries]
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) st[j][e] = st[j + 1][e] = dpt[j][e] = dpt[j + 1][e] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bb[4];
        load_b_nk_x4(bb, sQ, 8 * j, 16 * ks, lane);
        mma16816(st[j], ak[ks], bb[0], bb[1]);
        mma16816(st[j + 1], ak[ks], bb[2], bb[3]);
        load_b_nk_x4(bb, sdO, 8 * j, 16 * ks, lane);
        mma16816(dpt[j], av[ks], bb[0], bb[1]);
        mma16816(dpt[j + 1], av[ks], bb[2], bb[3]);
      }
    }
    uint32_t ap[4][4], ads[4][4];   // P^T and dS^T as A fragments (k = query index)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[4], ds[4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int qi = 8 * j + 2 * q + e;
          const bool ok = key_ok[r] && (qb * AB_BLK + qi < T);
          const float pv = ok ? ex2_approx(st[j][2 * r + e] * AB_LOG2E - s_lse[qi]) : 0.0f;
          float keep = 1.0f;
          if (dr.thr16) {
            const int kl = m0 + g + 8 * r;           // key inside this CTA's 64-key block
            keep = ((s_keep[qb & 1][qi][kl >> 2] >> (kl & 3)) & 1u) ? dr.scale : 0.0f;
          }
          p[2 * r + e] = pv * keep;                                   // dV uses the dropped probabilities
          ds[2 * r + e] = pv * (dpt[j][2 * r + e] * keep - s_ds[qi]);
        }
      }
      ap[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p[0], p[1]);
      ap[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
      ads[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
      ads[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // k = 16 queries
#pragma unroll
      for (int j = 0; j < 8; j += 2) {    // n = 2 x 8 head channels
        uint32_t bb[4];
        load_b_kn_x4(bb, sdO, 16 * ks, 8 * j, lane);
        mma16816(dv[j], ap[ks], bb[0], bb[1]);
        mma16816(dv[j + 1], ap[ks], bb[2], bb[3]);
        load_b_kn_x4(bb, sQ, 16 * ks, 8 * j, lane);
        mma16816(dk[j], ads[ks], bb[0], bb[1]);
        mma16816(dk[j + 1], ads[ks], bb[2], bb[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int t = k0 + m0 + g + 8 * r;
    if (t < T) {
      __nv_bfloat16* dst = dqkv + ((size_t)b * T + t) * ld3 + h * AB_DH;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(dst + d + 8 * j + 2 * q) = pack_bf16x2(dk[j][2 * r], dk[j][2 * r + 1]);
        *reinterpret_cast<uint32_t*>(dst + 2 * d + 8 * j + 2 * q) = pack_bf16x2(dv[j][2 * r], dv[j][2 * r + 1]);
      }
    }
  }
}

int launch_attn_bwd_tc(const void* qkv_hi, const void* dctx_hi, int B, int T, int H, const int* kv_len, float q_scale,
                       float* lse2, float* dsum, void* dqkv_hi, DropSpec dr, cudaStream_t stream);   // attn_bwd_tc.cu

}  // namespace w2v2

using namespace w2v2;

extern "C" int64_t w2v2_attn_bwd_workspace_bytes(int batch, int frames, int num_heads) {
  return (int64_t)2 * batch * num_heads * frames * (int64_t)sizeof(float);
}

extern "C" int w2v2_attn_bwd(const void* qkv_hi, const void* ctx_hi, const void* dctx_hi, int batch, int frames,
                             int num_heads, int head_size, const int32_t* kv_len, float q_scale, void* workspace,
                             void* dqkv_hi, float drop_p, uint64_t seed, uint32_t site, void* stream) {
  W2V2_CHECK_ARG(qkv_hi && ctx_hi && dctx_hi && workspace && dqkv_hi, "null pointer");
  W2V2_CHECK_ARG(head_size == AB_DH, "built for head_size 64");
  W2V2_CHECK_ARG(batch > 0 && frames > 0 && num_heads > 0, "batch, frames, num_heads must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* lse2 = reinterpret_cast<float*>(workspace);
  float* dsum = lse2 + (size_t)batch * num_heads * frames;
  W2V2_CHECK_ARG(drop_p >= 0.0f && drop_p < 1.0f, "drop_p must be in [0, 1)");
  const DropSpec dr = make_drop(drop_p, seed, site);
  // default: the tcgen05 / TMEM kernels of attn_bwd_tc.cu; W2V2_ATTN_BWD=mma keeps the mma.sync pair below (the reference
  // implementation the new kernels were validated against, and an A/B switch)
  static const bool use_mma = [] { const char* e = getenv("W2V2_ATTN_BWD"); return e && strcmp(e, "mma") == 0; }();
  if (!use_mma)
    return launch_attn_bwd_tc(qkv_hi, dctx_hi, batch, frames, num_heads, kv_len, q_scale, lse2, dsum, dqkv_hi, dr, s);
  dim3 grid((frames + AB_BLK - 1) / AB_BLK, batch * num_heads);
  attn_bwd_dq_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv_hi),
                                          reinterpret_cast<const __nv_bfloat16*>(ctx_hi),
                                          reinterpret_cast<const __nv_bfloat16*>(dctx_hi), kv_len, frames, num_heads, q_scale,
                                          reinterpret_cast<__nv_bfloat16*>(dqkv_hi), lse2, dsum, dr);
  W2V2_CUDA(cudaGetLastError());
  attn_bwd_dkv_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(qkv_hi),
                                           reinterpret_cast<const __nv_bfloat16*>(dctx_hi), kv_len, frames, num_heads, lse2,
                                           dsum, reinterpret_cast<__nv_bfloat16*>(dqkv_hi), dr);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}
