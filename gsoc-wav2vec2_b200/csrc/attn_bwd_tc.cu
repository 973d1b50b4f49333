// Backward of the attention core on tcgen05 / TMEM (sm_100a): the same math as attn_bwd.cu (which remains as the mma.sync
// reference implementation, W2V2_ATTN_BWD=mma), restructured like the forward kernel in attn.cu.
//
//     P = softmax(S), S = Q K^T;  dV = P^T dO;  dP = dO V^T;  D_i = sum_j P_ij dP_ij;  dS = P o (dP - D);
//     dQ = dS K;  dK = dS^T Q                  (reference forward: TransformerAttention.get_context, encoder.py:34-54)
//
// Two kernels, no atomics, deterministic; one CTA per SM, all 512 TMEM columns:
//   attn_bwd_dq_tc  : CTA = 128 queries of one (b, h), two sweeps over the 128-key chunks.
//        sweep 1:  S = Q K^T and dP = dO V^T on the tensor cores -> the softmax threads accumulate the online statistics
//                  (running max / sum) AND sum_j P_ij dP_ij, i.e. lse and the CONSISTENT D of this pass (see attn_bwd.cu for why D
//                  must not come from the stored context);
//        sweep 2:  S, dP again -> dS = P o (dP - D) as bf16 into TMEM -> dQ += dS K with dS as the TMEM A operand and the K tile as
//                  an MN-major B operand (exactly the forward's O += P V).
//   attn_bwd_dkv_tc : CTA = 128 keys of one (b, h), one sweep over the 128-query chunks with the TRANSPOSED products
//                  S^T = K Q^T, dP^T = V dO^T (rows = keys): P^T and dS^T are written to TMEM as A operands of
//                  dV += P^T dO and dK += dS^T Q (dO / Q chunks as MN-major B operands); lse and D come from the first kernel.
// Eight softmax warps per CTA: warp w owns TMEM lane quadrant w & 3 (32 rows) and column half w >> 2 (64 of the 128 columns of
// a chunk), so a thread holds 64 S and 64 dP values; the second sweep needs no cross-thread exchange, the first merges the two
// halves' statistics once at its end.  Dropout on the probabilities (encoder.py:41-43) is regenerated from the forward's
// stateless stream: per 4-key group in the row-major kernel, through a cooperative keep-bit tile in the transposed one.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int BT_BM = 128, BT_BN = 128, BT_DH = 64;
constexpr int BT_THREADS = 384;              // warps 0-7: softmax; warp 8: TMA + TMEM alloc; warp 9: MMA issuer; 10-11 idle
constexpr int BT_TILE = BT_BM * BT_DH * 2;   // 16 KB
constexpr int BT_REGS_SOFTMAX = 208, BT_REGS_CONTROL = 40;
constexpr float BT_LOG2E = 1.4426950408889634f;

struct BwdParams {
  int T, d, H;
  const int* kv_len;
  float q_scale;
  float* lse2;          // [B*H][T] log2-domain log-sum-exp (written by dq, read by dkv)
  float* dsum;          // [B*H][T] D
  __nv_bfloat16* dqkv;  // [B][T][3d]
  DropSpec drop;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// 64 consecutive fp32 TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld64(uint32_t addr, uint32_t (&a)[32], uint32_t (&b)[32]) {
  tmem_ld_32x32b_x32(addr, a);
  tmem_ld_32x32b_x32(addr + 32, b);
}

// ------------------------------------------------------------------------------------ dQ (+ lse2, D)
// smem: Q 16 KB | dO 16 KB | 2 x (K 16 KB, V 16 KB) | barriers | merge scratch
struct DqSmem {
  static constexpr int Q_OFF = 0, DO_OFF = BT_TILE, KV_OFF = 2 * BT_TILE, KV_STAGE = 2 * BT_TILE, STAGES = 2;
  static constexpr int BAR_OFF = KV_OFF + STAGES * KV_STAGE;
  static constexpr int SCR_OFF = BAR_OFF + 128;                 // [128 rows][4] floats: (m, l, acc) of half 1, then (lse, D)
  static constexpr int TOTAL = SCR_OFF + BT_BM * 16;
};

__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const BwdParams p) {
  using S = DqSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* qd_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* sd_full = bars + 5;
  uint64_t* sd_empty = bars + 6;
  uint64_t* ds_full = bars + 7;
  uint64_t* ds_empty = bars + 8;
  uint64_t* dq_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float4* scratch = reinterpret_cast<float4*>(smem + S::SCR_OFF);

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int q0 = blockIdx.x * BT_BM, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.H + h;
  const int kv_raw = (p.kv_len != nullptr) ? min(p.kv_len[b], p.T) : p.T;
  const int klen = (kv_raw <= 0) ? p.T : kv_raw;
  const int nchunks = (klen + BT_BN - 1) / BT_BN;
  const int nsteps = 2 * nchunks;

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 9 && elect_one()) {
    mbar_init(qd_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(sd_full, 1);
    mbar_init(sd_empty, 8);
    mbar_init(ds_full, 8);
    mbar_init(ds_empty, 1);
    mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_ds = tmem_base + 256, tmem_dq = tmem_base + 320;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BT_REGS_CONTROL));
    if (warp == 8) {
      // ---------------------------------------------------------------- TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(qd_full, 2 * BT_TILE);
        tma_load_3d(smem + S::Q_OFF, &tm_qkv, qd_full, h * BT_DH, q0, b);
        tma_load_3d(smem + S::DO_OFF, &tm_do, qd_full, h * BT_DH, q0, b);
        for (int s = 0; s < nsteps; ++s) {
          const int stage = s & 1, j = (s >= nchunks) ? s - nchunks : s;
          mbar_wait(&kv_empty[stage], ((s >> 1) & 1) ^ 1);
          uint8_t* kbuf = smem + S::KV_OFF + stage * S::KV_STAGE;
          mbar_arrive_expect_tx(&kv_full[stage], S::KV_STAGE);
          tma_load_3d(kbuf, &tm_qkv, &kv_full[stage], p.d + h * BT_DH, j * BT_BN, b);
          tma_load_3d(kbuf + BT_TILE, &tm_qkv, &kv_full[stage], 2 * p.d + h * BT_DH, j * BT_BN, b);
        }
      }
    } else if (warp == 9) {
      // ---------------------------------------------------------------- MMA issuer
      if (elect_one()) {
        constexpr uint32_t idesc_qk = idesc_bf16(BT_BM, BT_BN, 0, 0);   // S = Q K^T, dP = dO V^T: both operands K-major
        constexpr uint32_t idesc_dq = idesc_bf16(BT_BM, BT_DH, 0, 1);   // dQ += dS K: A = dS (TMEM), B = K MN-major
        const uint64_t dq_ = desc_kmajor_sw128(smem_u32(smem + S::Q_OFF));
        const uint64_t ddo = desc_kmajor_sw128(smem_u32(smem + S::DO_OFF));
        mbar_wait(qd_full, 0);
        for (int s = 0; s < nsteps; ++s) {
          const int stage = s & 1;
          const uint32_t k_addr = smem_u32(smem + S::KV_OFF + stage * S::KV_STAGE);
          mbar_wait(&kv_full[stage], (s >> 1) & 1);
          if (s > 0) mbar_wait(sd_empty, (s - 1) & 1);   // the softmax warps have pulled S / dP of the previous step out of TMEM
          tc_fence_after();
          const uint64_t dk = desc_kmajor_sw128(k_addr), dv = desc_kmajor_sw128(k_addr + BT_TILE);
#pragma unroll
          for (int k = 0; k < BT_DH / 16; ++k) umma_f16(tmem_s, dq_ + 2 * k, dk + 2 * k, idesc_qk, k != 0);
#pragma unroll
          for (int k = 0; k < BT_DH / 16; ++k) umma_f16(tmem_dp, ddo + 2 * k, dv + 2 * k, idesc_qk, k != 0);
          umma_commit(sd_full);
          if (s < nchunks) {
            umma_commit(&kv_empty[stage]);               // sweep 1 only reads K / V through these two products
          } else {
            const int u = s - nchunks;
            mbar_wait(ds_full, u & 1);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < BT_BN / 16; ++ks)
              umma_f16_tmem_a(tmem_dq, tmem_ds + ks * 8, desc_mnmajor_sw128(k_addr + ks * 2048, 1024, 1024), idesc_dq, (u | ks) != 0);
            umma_commit(ds_empty);
            umma_commit(&kv_empty[stage]);
            if (s == nsteps - 1) umma_commit(dq_done);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BT_REGS_SOFTMAX));
    // ---------------------------------------------------------------- softmax warps: row = 32 (w & 3) + lane, column half = w >> 2
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t cs = (uint32_t)(half * 64);
    float m_run = -INFINITY, l_run = 0.0f, a_run = 0.0f;
    float lse = 0.0f, dsum = 0.0f;
    const uint64_t row_group = attn_row_group(bh, q0 + r, p.T);
    for (int s = 0; s < nsteps; ++s) {
      const bool sweep2 = s >= nchunks;
      const int j = sweep2 ? s - nchunks : s;
      const int key0 = j * BT_BN + (int)cs;          // first key of this thread's 64 columns
      if (s == nchunks) {
        // ---- end of sweep 1: merge the two column halves' (m, l, acc) -> lse, D
        if (half == 1) scratch[r] = make_float4(m_run, l_run, a_run, 0.0f);
        named_bar_sync(1, 256);
        if (half == 0) {
          const float4 o = scratch[r];
          const float M = fmaxf(m_run, o.x);
          const float w0 = ex2_approx(m_run - M), w1 = ex2_approx(o.x - M);
          const float L = l_run * w0 + o.y * w1, A = a_run * w0 + o.z * w1;
          lse = M + log2f(L);
          dsum = A / L;
          const int t = q0 + r;
          if (t < p.T) {
            p.lse2[(size_t)bh * p.T + t] = lse;
            p.dsum[(size_t)bh * p.T + t] = dsum;
          }
        }
        named_bar_sync(1, 256);                       // half 1 has read its own entry before half 0 overwrites it
        if (half == 0) scratch[r] = make_float4(lse, dsum, 0.0f, 0.0f);
        named_bar_sync(1, 256);
        if (half == 1) {
          const float4 o = scratch[r];
          lse = o.x;
          dsum = o.y;
        }
      }
      uint32_t sS[2][32], sD[2][32];
      mbar_wait(sd_full, s & 1);
      tc_fence_after();
      tmem_ld64(tmem_s + lane_sel + cs, sS[0], sS[1]);
      tmem_ld64(tmem_dp + lane_sel + cs, sD[0], sD[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sd_empty);
      // dropout: dP_eff = keep ? dP / (1 - p) : 0 (the forward dropped P AFTER the softmax)
      if (p.drop.thr16) {
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const uint64_t bits = drop_bits4(p.drop, row_group + (uint64_t)((key0 >> 2) + pc * 8 + i4));
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v = __uint_as_float(sD[pc][4 * i4 + e]);
              sD[pc][4 * i4 + e] = drop_keep(bits, e, p.drop.thr16) ? __float_as_uint(v * p.drop.scale) : 0u;
            }
          }
        }
      }
      const bool ragged = key0 + 64 > klen;          // only the last chunk can hold masked keys
      if (ragged) {
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (key0 + pc * 32 + i >= klen) sS[pc][i] = 0xff800000u;   // -inf: probability exactly 0 in both sweeps
        }
      }
      if (!sweep2) {
        // ---- online statistics of this thread's 64 keys (log2 domain); packed fp32x2 math, 4 independent max chains
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              mx[c] = fmaxf(fmaxf(mx[c], __uint_as_float(sS[pc][i + 2 * c])), __uint_as_float(sS[pc][i + 2 * c + 1]));
          }
        }
        const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * BT_LOG2E;
        const float m_new = fmaxf(m_run, cmax);
        if (m_new > -INFINITY) {
          const uint64_t l2 = pack2(BT_LOG2E, BT_LOG2E), mn2 = pack2(-m_new, -m_new);
          uint64_t ps2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)}, pd2[2] = {pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll
          for (int pc = 0; pc < 2; ++pc) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0, a1;
              unpack2(fma2(pack2(__uint_as_float(sS[pc][i]), __uint_as_float(sS[pc][i + 1])), l2, mn2), a0, a1);
              const uint64_t e2 = pack2(ex2_approx(a0), ex2_approx(a1));
              ps2[(i >> 1) & 1] = add2(ps2[(i >> 1) & 1], e2);
              pd2[(i >> 1) & 1] = fma2(e2, pack2(__uint_as_float(sD[pc][i]), __uint_as_float(sD[pc][i + 1])), pd2[(i >> 1) & 1]);
            }
          }
          float p0, p1, p2, p3, d0, d1, d2, d3;
          unpack2(ps2[0], p0, p1);
          unpack2(ps2[1], p2, p3);
          unpack2(pd2[0], d0, d1);
          unpack2(pd2[1], d2, d3);
          const float resc = ex2_approx(m_run - m_new);   // 0 when nothing had been accumulated
          l_run = l_run * resc + ((p0 + p1) + (p2 + p3));
          a_run = a_run * resc + ((d0 + d1) + (d2 + d3));
          m_run = m_new;
        }
      } else {
        // ---- dS = P o (dP - D) -> bf16 pairs -> TMEM (A operand of dQ += dS K)
        uint32_t pk[32];
        const uint64_t l2 = pack2(BT_LOG2E, BT_LOG2E), nl2 = pack2(-lse, -lse), nd2 = pack2(-dsum, -dsum);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int i = 2 * c;
          float a0, a1, r0, r1;
          unpack2(fma2(pack2(__uint_as_float(sS[i >> 5][i & 31]), __uint_as_float(sS[i >> 5][(i & 31) + 1])), l2, nl2), a0, a1);
          const uint64_t p2 = pack2(ex2_approx(a0), ex2_approx(a1));   // masked keys: ex2(-inf) = 0
          unpack2(mul2(p2, add2(pack2(__uint_as_float(sD[i >> 5][i & 31]), __uint_as_float(sD[i >> 5][(i & 31) + 1])), nd2)), r0, r1);
          pk[c] = pack_bf16x2(r0, r1);
        }
        const int u = s - nchunks;
        if (u > 0) mbar_wait(ds_empty, (u - 1) & 1);      // the previous dQ MMAs have read the dS columns
        tc_fence_after();
        tmem_st_32x32b_x32(tmem_ds + lane_sel + half * 32, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full);
      }
    }
    // ---- epilogue: dQ * q_scale -> dqkv[:, 0:d]; each half stores 32 of the 64 head channels
    mbar_wait(dq_done, 0);
    tc_fence_after();
    uint32_t rr[32];
    tmem_ld_32x32b_x32(tmem_dq + lane_sel + half * 32, rr);
    tmem_ld_wait();
    const int t = q0 + r;
    if (t < p.T) {
      __nv_bfloat16* dst = p.dqkv + ((size_t)b * p.T + t) * (3 * (size_t)p.d) + h * BT_DH + half * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = pack_bf16x2(__uint_as_float(rr[8 * q + 2 * e]) * p.q_scale, __uint_as_float(rr[8 * q + 2 * e + 1]) * p.q_scale);
        *reinterpret_cast<uint4*>(dst + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------ dK, dV
// smem: K 16 KB | V 16 KB | 2 x (Q 16 KB, dO 16 KB) | barriers | 2 x (lse[128], D[128]) | 2 x keep tile [128 q][32] bytes
struct DkvSmem {
  static constexpr int K_OFF = 0, V_OFF = BT_TILE, QD_OFF = 2 * BT_TILE, QD_STAGE = 2 * BT_TILE, STAGES = 2;
  static constexpr int BAR_OFF = QD_OFF + STAGES * QD_STAGE;
  static constexpr int VEC_OFF = BAR_OFF + 128;                 // [2][2][128] floats
  static constexpr int KEEP_OFF = VEC_OFF + 2 * 2 * 128 * 4;    // [2][128][32] bytes
  static constexpr int TOTAL = KEEP_OFF + 2 * 128 * 32;
};

__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do, const BwdParams p) {
  using S = DkvSmem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* kv_full = bars + 0;
  uint64_t* qd_full = bars + 1;    // [2]
  uint64_t* qd_empty = bars + 3;   // [2]
  uint64_t* sd_full = bars + 5;
  uint64_t* sd_empty = bars + 6;
  uint64_t* pds_full = bars + 7;
  uint64_t* pds_empty = bars + 8;
  uint64_t* done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* vec = reinterpret_cast<float*>(smem + S::VEC_OFF);
  uint8_t* keep = smem + S::KEEP_OFF;

  const int warp = threadIdx.x >> 5, lane = lane_id();
  const int k0 = blockIdx.x * BT_BM, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.H + h;
  const int kv_raw = (p.kv_len != nullptr) ? min(p.kv_len[b], p.T) : p.T;
  const int klen = (kv_raw <= 0) ? p.T : kv_raw;
  const int nchunks = (p.T + BT_BN - 1) / BT_BN;     // query chunks

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 9 && elect_one()) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(sd_full, 1);
    mbar_init(sd_empty, 8);
    mbar_init(pds_full, 8);
    mbar_init(pds_empty, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 128, tmem_pt = tmem_base + 256, tmem_dst = tmem_base + 320;
  const uint32_t tmem_dv = tmem_base + 384, tmem_dk = tmem_base + 448;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BT_REGS_CONTROL));
    if (warp == 8) {
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * BT_TILE);
        tma_load_3d(smem + S::K_OFF, &tm_qkv, kv_full, p.d + h * BT_DH, k0, b);
        tma_load_3d(smem + S::V_OFF, &tm_qkv, kv_full, 2 * p.d + h * BT_DH, k0, b);
        for (int c = 0; c < nchunks; ++c) {
          const int stage = c & 1;
          mbar_wait(&qd_empty[stage], ((c >> 1) & 1) ^ 1);
          uint8_t* qbuf = smem + S::QD_OFF + stage * S::QD_STAGE;
          mbar_arrive_expect_tx(&qd_full[stage], S::QD_STAGE);
          tma_load_3d(qbuf, &tm_qkv, &qd_full[stage], h * BT_DH, c * BT_BN, b);
          tma_load_3d(qbuf + BT_TILE, &tm_do, &qd_full[stage], h * BT_DH, c * BT_BN, b);
        }
      }
    } else if (warp == 9) {
      if (elect_one()) {
        constexpr uint32_t idesc_t = idesc_bf16(BT_BM, BT_BN, 0, 0);    // S^T = K Q^T, dP^T = V dO^T
        constexpr uint32_t idesc_o = idesc_bf16(BT_BM, BT_DH, 0, 1);    // dV += P^T dO, dK += dS^T Q: A in TMEM, B MN-major
        const uint64_t dk = desc_kmajor_sw128(smem_u32(smem + S::K_OFF));
        const uint64_t dv = desc_kmajor_sw128(smem_u32(smem + S::V_OFF));
        mbar_wait(kv_full, 0);
        for (int c = 0; c < nchunks; ++c) {
          const int stage = c & 1;
          const uint32_t q_addr = smem_u32(smem + S::QD_OFF + stage * S::QD_STAGE);
          mbar_wait(&qd_full[stage], (c >> 1) & 1);
          if (c > 0) mbar_wait(sd_empty, (c - 1) & 1);
          tc_fence_after();
          const uint64_t dq_ = desc_kmajor_sw128(q_addr), ddo = desc_kmajor_sw128(q_addr + BT_TILE);
#pragma unroll
          for (int k = 0; k < BT_DH / 16; ++k) umma_f16(tmem_st, dk + 2 * k, dq_ + 2 * k, idesc_t, k != 0);
#pragma unroll
          for (int k = 0; k < BT_DH / 16; ++k) umma_f16(tmem_dpt, dv + 2 * k, ddo + 2 * k, idesc_t, k != 0);
          umma_commit(sd_full);
          mbar_wait(pds_full, c & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < BT_BN / 16; ++ks) {
            umma_f16_tmem_a(tmem_dv, tmem_pt + ks * 8, desc_mnmajor_sw128(q_addr + BT_TILE + ks * 2048, 1024, 1024), idesc_o, (c | ks) != 0);
            umma_f16_tmem_a(tmem_dk, tmem_dst + ks * 8, desc_mnmajor_sw128(q_addr + ks * 2048, 1024, 1024), idesc_o, (c | ks) != 0);
          }
          umma_commit(pds_empty);
          umma_commit(&qd_empty[stage]);
          if (c == nchunks - 1) umma_commit(done);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BT_REGS_SOFTMAX));
    // ---------------------------------------------------------------- row = key 32 (w & 3) + lane, column (query) half = w >> 2
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t cs = (uint32_t)(half * 64);
    const bool key_ok = k0 + r < klen;
    const int tid = threadIdx.x;                     // 0..255
    for (int c = 0; c < nchunks; ++c) {
      const int q_first = c * BT_BN;
      float* lsev = vec + (c & 1) * 256;
      float* dv_ = lsev + 128;
      uint8_t* kp = keep + (c & 1) * (128 * 32);
      // ---- this chunk's lse / D vectors (+ dropout keep bits [query][4-key group]) -> smem
      {
        const int q = q_first + (tid & 127);
        const float* src = (tid < 128 ? p.lse2 : p.dsum) + (size_t)bh * p.T;
        // queries past T: lse = +inf makes every probability of that column exactly 0
        (tid < 128 ? lsev : dv_)[tid & 127] = (q < p.T) ? __ldg(src + q) : (tid < 128 ? INFINITY : 0.0f);
        if (p.drop.thr16) {
          for (int i = tid; i < 128 * 32; i += 256) {
            const int ql = i >> 5, kg = i & 31;
            const uint64_t bits = drop_bits4(p.drop, attn_row_group(bh, q_first + ql, p.T) + (uint64_t)((k0 >> 2) + kg));
            uint32_t m = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) m |= drop_keep(bits, e, p.drop.thr16) ? (1u << e) : 0u;
            kp[i] = (uint8_t)m;
          }
        }
      }
      named_bar_sync(1, 256);
      uint32_t sS[2][32], sD[2][32];
      mbar_wait(sd_full, c & 1);
      tc_fence_after();
      tmem_ld64(tmem_st + lane_sel + cs, sS[0], sS[1]);
      tmem_ld64(tmem_dpt + lane_sel + cs, sD[0], sD[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sd_empty);
      uint32_t pp[32], pd[32];     // P^T (dropped) and dS^T as bf16 pairs over this thread's 64 queries
      const uint64_t l2 = pack2(BT_LOG2E, BT_LOG2E);
      const float rowmask = key_ok ? 0.0f : -INFINITY;     // a masked key row: every probability exactly 0
#pragma unroll
      for (int g4 = 0; g4 < 16; ++g4) {                    // four queries per step: one 16-byte smem read of lse and of D
        const float4 l4 = *reinterpret_cast<const float4*>(lsev + cs + 4 * g4);
        const float4 d4 = *reinterpret_cast<const float4*>(dv_ + cs + 4 * g4);
        const float lq[4] = {l4.x, l4.y, l4.z, l4.w}, dq4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int i = 4 * g4 + 2 * hp;
          float a0, a1;
          unpack2(fma2(pack2(__uint_as_float(sS[i >> 5][i & 31]), __uint_as_float(sS[i >> 5][(i & 31) + 1])), l2,
                       pack2(rowmask - lq[2 * hp], rowmask - lq[2 * hp + 1])), a0, a1);
          const uint64_t p2 = pack2(ex2_approx(a0), ex2_approx(a1));
          uint64_t dp2 = pack2(__uint_as_float(sD[i >> 5][i & 31]), __uint_as_float(sD[i >> 5][(i & 31) + 1]));
          uint64_t pdrop2 = p2;
          if (p.drop.thr16) {
            const int ql = (int)cs + i;
            const float k0f = ((kp[ql * 32 + (r >> 2)] >> (r & 3)) & 1u) ? p.drop.scale : 0.0f;
            const float k1f = ((kp[(ql + 1) * 32 + (r >> 2)] >> (r & 3)) & 1u) ? p.drop.scale : 0.0f;
            const uint64_t kf2 = pack2(k0f, k1f);
            dp2 = mul2(dp2, kf2);
            pdrop2 = mul2(p2, kf2);                                      // dV uses the dropped probabilities
          }
          float q0v, q1v, s0v, s1v;
          unpack2(pdrop2, q0v, q1v);
          unpack2(mul2(p2, add2(dp2, pack2(-dq4[2 * hp], -dq4[2 * hp + 1]))), s0v, s1v);
          pp[i >> 1] = pack_bf16x2(q0v, q1v);
          pd[i >> 1] = pack_bf16x2(s0v, s1v);
        }
      }
      if (c > 0) mbar_wait(pds_empty, (c - 1) & 1);
      tc_fence_after();
      tmem_st_32x32b_x32(tmem_pt + lane_sel + half * 32, pp);
      tmem_st_32x32b_x32(tmem_dst + lane_sel + half * 32, pd);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // ---- epilogue: half 0 stores dK, half 1 stores dV (64 head channels each)
    mbar_wait(done, 0);
    tc_fence_after();
    const int t = k0 + r;
    __nv_bfloat16* dst = p.dqkv + ((size_t)b * p.T + (t < p.T ? t : 0)) * (3 * (size_t)p.d) + (half == 0 ? p.d : 2 * p.d) + h * BT_DH;
#pragma unroll
    for (int piece = 0; piece < 2; ++piece) {
      uint32_t rr[32];
      tmem_ld_32x32b_x32((half == 0 ? tmem_dk : tmem_dv) + lane_sel + piece * 32, rr);
      tmem_ld_wait();
      if (t < p.T) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2(__uint_as_float(rr[8 * q + 2 * e]), __uint_as_float(rr[8 * q + 2 * e + 1]));
          *reinterpret_cast<uint4*>(dst + piece * 32 + 8 * q) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc<512>(tmem_base);
}

int launch_attn_bwd_tc(const void* qkv_hi, const void* dctx_hi, int B, int T, int H, const int* kv_len, float q_scale,
                       float* lse2, float* dsum, void* dqkv_hi, DropSpec dr, cudaStream_t stream) {
  const int d = H * BT_DH;
  CUtensorMap tm_qkv, tm_do;
  const uint64_t qdims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
  const uint64_t qstr[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
  const uint64_t odims[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)B};
  const uint64_t ostr[2] = {(uint64_t)d * 2, (uint64_t)T * d * 2};
  const uint32_t box[3] = {BT_DH, BT_BM, 1};
  int rc = make_tmap(&tm_qkv, qkv_hi, 3, qdims, qstr, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if ((rc = make_tmap(&tm_do, dctx_hi, 3, odims, ostr, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  BwdParams p;
  p.T = T;
  p.d = d;
  p.H = H;
  p.kv_len = kv_len;
  p.q_scale = q_scale;
  p.lse2 = lse2;
  p.dsum = dsum;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv_hi);
  p.drop = dr;
  static unsigned long long done_a = 0, done_b = 0;
  W2V2_CUDA(ensure_dyn_smem(attn_bwd_dq_tc_kernel, DqSmem::TOTAL, done_a));
  W2V2_CUDA(ensure_dyn_smem(attn_bwd_dkv_tc_kernel, DkvSmem::TOTAL, done_b));
  dim3 grid((T + BT_BM - 1) / BT_BM, H, B);
  attn_bwd_dq_tc_kernel<<<grid, BT_THREADS, DqSmem::TOTAL, stream>>>(tm_qkv, tm_do, p);
  W2V2_CUDA(cudaGetLastError());
  attn_bwd_dkv_tc_kernel<<<grid, BT_THREADS, DkvSmem::TOTAL, stream>>>(tm_qkv, tm_do, p);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace w2v2
