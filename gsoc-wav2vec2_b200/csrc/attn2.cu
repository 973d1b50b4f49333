// Attention forward, single-pass modes, in the FlashAttention-4 shape: ONE persistent CTA per SM works on TWO query tiles
// (q tiles 2i and 2i + 1 of one (utterance, head)) that share every K / V chunk.
//     warpgroup 0 / 1 (warps 0-3 / 4-7): softmax of tile A / tile B, one score row per thread (208 registers)
//     warpgroup 2 (warps 8-11): output warps - O / l -> context of a finished tile, while the softmax warps are in the next one
//     warpgroup 3 (warps 12-15): warp 12 = TMEM alloc + TMA producer, warps 13 / 14 = MMA issuers of tile A / tile B (one elected
//         thread each; ONE issuer walking both tiles in a fixed order makes the faster softmax warpgroup wait for the slower one
//         at every chunk - ncu: 20 % of the softmax warps' samples on the S barrier)
// TMEM (all 512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// Shared memory: Q_A, Q_B (32 KB) + a 6-stage K / V ring (192 KB): both tiles read the same K / V tiles - half the L2 traffic of
// two independent CTAs.  Semantics, operand layouts and the lazy rescaling are those of attn.cu (reference: encoder.py:34-54,
// mask :256-263); the 3-pass (parity) modes stay in attn.cu.  Why this shape: profiles/r2_attn_fwd.md - with two CTAs per SM the
// register file cannot host an output warpgroup next to a 128-value score row per softmax thread; here it can.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int A2_BM = 128, A2_BN = 128, A2_DH = 64;
constexpr int A2_THREADS = 512;
constexpr int A2_REGS_SOFTMAX = 208, A2_REGS_OUTPUT = 48, A2_REGS_CONTROL = 48;   // 128 x (2 x 208 + 48 + 48) = 65536
constexpr int A2_TILE = A2_BM * A2_DH * 2;       // 16 KB: one [128][64] 16-bit tile
constexpr int A2_STAGES = 6;
// Groups of 4 keys (one bit per group of a 32-key block) whose exponentials run as a polynomial on the FMA pipe.  Measured per layer
// at B = 32, T = 768: none 87.1 us, 1 of 8 86.7, 2 of 8 82.6 - 82.9 (groups 2 and 5; 84.4 for groups 3 and 7), 3 of 8 83.1 - 84.5,
// 4 of 8 88.4.
#ifndef A2_POLY_MASK
#define A2_POLY_MASK 0x24
#endif
#ifndef A2_PINGPONG
#define A2_PINGPONG 0
#endif
constexpr int A2_Q_OFF = 0;
constexpr int A2_KV_OFF = 2 * A2_TILE;
constexpr int A2_STAGE_BYTES = 2 * A2_TILE;      // K and V
constexpr int A2_L_OFF = A2_KV_OFF + A2_STAGES * A2_STAGE_BYTES;   // row sums: [2 tiles][128] floats
constexpr int A2_BAR_OFF = A2_L_OFF + 1024;
constexpr int A2_SMEM = A2_BAR_OFF + 512;

// 2^x for x <= ~9 on the FMA / ALU pipes (FA4's MUFU off-load): x = n + f with n = rint(x) from the 1.5 * 2^23 trick, |f| <= 0.5,
// 2^f as a degree-3 minimax polynomial (1.0e-4 relative: below the 2^-9 / 2^-11 rounding of the bf16 / fp16 probabilities), 2^n by
// adding n to the exponent field.  -inf (masked keys) clamps to 2^-125.
__device__ __forceinline__ void ex2_poly_x2(float& a0, float& a1) {
  const uint64_t x = pack2(fmaxf(a0, -125.0f), fmaxf(a1, -125.0f));
  const uint64_t t = add2(x, pack2(12582912.0f, 12582912.0f));
  const uint64_t nf = add2(t, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(nf, pack2(-1.0f, -1.0f), x);
  uint64_t q = fma2(f, pack2(0.05500892922282219f, 0.05500892922282219f), pack2(0.24221095442771912f, 0.24221095442771912f));
  q = fma2(q, f, pack2(0.6932829022407532f, 0.6932829022407532f));
  q = fma2(q, f, pack2(1.0f, 1.0f));
  float p0, p1, t0, t1;
  unpack2(q, p0, p1);
  unpack2(t, t0, t1);
  a0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  a1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

struct Attn2Params {
  int T, d, H;
  const int* kv_len;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  DropSpec drop;
  int out_format;
  int q_pairs;        // ceil(ceil(T / 128) / 2)
  int n_items;        // q_pairs * H * B
};

template <bool FP16, bool DROP>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm, const Attn2Params p) {
  constexpr float LOG2E = FP16 ? 1.4426950408889634f / (ACT_SCALE * ACT_SCALE) : 1.4426950408889634f;   // of the SCALED score
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;               // [6]
  uint64_t* kv_empty = bars + 8;              // [6]
  // per tile w = 0 (A), 1 (B): index 14 + 8 w + k
  auto tb = [&](int w, int k) { return bars + 14 + 8 * w + k; };
  enum { S_FULL = 0, S_EMPTY, P_FULL, P_EMPTY, PV_DONE, O_FULL, L_FULL, EPI_DONE };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  float* lbuf = reinterpret_cast<float*>(smem + A2_L_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  auto item_coords = [&](int it, int& q0, int& h, int& b) {
    q0 = (it % p.q_pairs) * (2 * A2_BM);
    const int bh = it / p.q_pairs;
    h = bh % p.H;
    b = bh / p.H;
  };
  auto item_kv_len = [&](int b) {       // see attn.cu: an utterance without a valid key attends over all keys
    const int kv_raw = (p.kv_len != nullptr) ? min(p.kv_len[b], p.T) : p.T;
    return (kv_raw <= 0) ? p.T : kv_raw;
  };

  if (warp == 12 && elect_one()) tma_prefetch_desc(&tm);
  if (warp == 13 && elect_one()) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 2);                 // both tiles' issuers release Q and every K / V stage
    for (int i = 0; i < A2_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(tb(w, S_FULL), 1);
      mbar_init(tb(w, S_EMPTY), 4);
      mbar_init(tb(w, P_FULL), 4);
      mbar_init(tb(w, P_EMPTY), 1);
      mbar_init(tb(w, PV_DONE), 1);
      mbar_init(tb(w, O_FULL), 1);
      mbar_init(tb(w, L_FULL), 4);
      mbar_init(tb(w, EPI_DONE), 4);
    }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_CONTROL));
    if (warp == 12) {
      // ---------------------------------------------------------------- TMA producer
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0, it = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
          int q0, h, b;
          item_coords(item, q0, h, b);
          const int nchunks = (item_kv_len(b) + A2_BN - 1) / A2_BN;
          mbar_wait(q_empty, (it & 1) ^ 1);
          mbar_arrive_expect_tx(q_full, 2 * A2_TILE);
          tma_load_3d(smem + A2_Q_OFF, &tm, q_full, h * A2_DH, q0, b);
          tma_load_3d(smem + A2_Q_OFF + A2_TILE, &tm, q_full, h * A2_DH, q0 + A2_BM, b);   // rows >= T: zero filled
          for (int j = 0; j < nchunks; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            uint8_t* kbuf = smem + A2_KV_OFF + stage * A2_STAGE_BYTES;
            mbar_arrive_expect_tx(&kv_full[stage], A2_STAGE_BYTES);
            tma_load_3d(kbuf, &tm, &kv_full[stage], p.d + h * A2_DH, j * A2_BN, b);
            tma_load_3d(kbuf + A2_TILE, &tm, &kv_full[stage], 2 * p.d + h * A2_DH, j * A2_BN, b);
            if (++stage == A2_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (warp <= 14) {
      // ---------------------------------------------------------------- MMA issuer of tile w
      const int w = warp - 13;
      if (elect_one()) {
        constexpr uint32_t idesc_s = idesc_16bit(FP16, A2_BM, A2_BN, 0, 0);
        constexpr uint32_t idesc_pv = idesc_16bit(FP16, A2_BM, A2_DH, 0, 1);
        const uint32_t q_addr = smem_u32(smem + A2_Q_OFF);
        auto chunks_of = [&](int item) {
          int q0, h, b;
          item_coords(item, q0, h, b);
          return (item_kv_len(b) + A2_BN - 1) / A2_BN;
        };
        // Q K^T of one chunk (S_w <- Q_w K^T)
        auto issue_qk = [&](int stage) {
          const uint64_t dq = desc_kmajor_sw128(q_addr + w * A2_TILE);
          const uint64_t dk = desc_kmajor_sw128(smem_u32(smem + A2_KV_OFF + stage * A2_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < A2_DH / 16; ++k) umma_f16(tmem_base + 128 * w, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
          umma_commit(tb(w, S_FULL));
        };
        int stage = 0;
        uint32_t kv_phase = 0, g = 0, it = 0;
        int item = blockIdx.x;
        int nchunks = (item < p.n_items) ? chunks_of(item) : 0;
        if (item < p.n_items) {
          mbar_wait(q_full, 0);
          mbar_wait(&kv_full[0], 0);
          tc_fence_after();
          issue_qk(0);
          if (nchunks == 1) umma_commit(q_empty);
        }
        for (; item < p.n_items; item += gridDim.x, ++it) {
          const int item_next = item + gridDim.x;
          const int nchunks_next = (item_next < p.n_items) ? chunks_of(item_next) : 0;
          for (int j = 0; j < nchunks; ++j, ++g) {
            const uint32_t par = g & 1;
            const uint32_t v_addr = smem_u32(smem + A2_KV_OFF + stage * A2_STAGE_BYTES) + A2_TILE;
            int nstage = stage + 1;
            uint32_t nphase = kv_phase;
            if (nstage == A2_STAGES) {
              nstage = 0;
              nphase ^= 1;
            }
            const bool same_item = j + 1 < nchunks;
            if (same_item || nchunks_next > 0) {
              if (!same_item) mbar_wait(q_full, (it + 1) & 1);
              mbar_wait(&kv_full[nstage], nphase);
              mbar_wait(tb(w, S_EMPTY), par);       // S_w(g) is in registers
              tc_fence_after();
              issue_qk(nstage);
              if (same_item ? (j + 2 == nchunks) : (nchunks_next == 1)) umma_commit(q_empty);
            }
            mbar_wait(tb(w, P_FULL), par);
            if (j == 0 && it > 0) mbar_wait(tb(w, EPI_DONE), (it - 1) & 1);   // the output warps hold the previous O_w
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < A2_BN / 16; ++ks) {
              const uint64_t dv = desc_mnmajor_sw128(v_addr + ks * 2048, 1024, 1024);
              umma_f16_tmem_a(tmem_base + 256 + 64 * w, tmem_base + 384 + 64 * w + ks * 8, dv, idesc_pv, (j | ks) != 0);
            }
            umma_commit(tb(w, PV_DONE));
            umma_commit(tb(w, P_EMPTY));
            if (j + 1 == nchunks) umma_commit(tb(w, O_FULL));
            umma_commit(&kv_empty[stage]);
            stage = nstage;
            kv_phase = nphase;
          }
          nchunks = nchunks_next;
        }
      }
    }
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- output warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A2_REGS_OUTPUT));
    const int ow = warp - 8;
    const int r = ow * 32 + lane;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      int q0, h, b;
      item_coords(item, q0, h, b);
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        mbar_wait(tb(w, L_FULL), it & 1);
        const float l_run = lbuf[w * 128 + r];
        mbar_wait(tb(w, O_FULL), it & 1);
        tc_fence_after();
        const int t = q0 + w * A2_BM + r;
        const float inv = (1.0f / l_run) * ((FP16 ? 1.0f / ACT_SCALE : 1.0f) * (p.out_format != 0 ? ACT_SCALE : 1.0f));
        const size_t off = ((size_t)b * p.T + t) * p.d + (size_t)h * A2_DH;
        const uint32_t o_addr = tmem_base + 256 + 64 * w + ((uint32_t)(ow * 32) << 16);
#pragma unroll
        for (int c16 = 0; c16 < 4; ++c16) {
          uint32_t rr[16];
          tmem_ld_32x32b_x16(o_addr + 16 * c16, rr);
          tmem_ld_wait();
          if (c16 == 3) {       // O_w and its row sums are in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tb(w, EPI_DONE));
          }
          if (t < p.T) {
            if (p.out_format == 2) {
              uint8_t* p8 = reinterpret_cast<uint8_t*>(p.out_lo) + ((size_t)b * p.T + t) * p.d * 2 + (size_t)h * 128 + 16 * c16;
              uint32_t hi[8];
              uint16_t l8[8], h8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                hi[e] = split_f16_f8x2(__uint_as_float(rr[2 * e]) * inv, __uint_as_float(rr[2 * e + 1]) * inv, l8[e], h8[e]);
              *reinterpret_cast<uint4*>(p.out_hi + off + 16 * c16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(p.out_hi + off + 16 * c16 + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(p8) = make_uint4(l8[0] | ((uint32_t)l8[1] << 16), l8[2] | ((uint32_t)l8[3] << 16),
                                                         l8[4] | ((uint32_t)l8[5] << 16), l8[6] | ((uint32_t)l8[7] << 16));
              *reinterpret_cast<uint4*>(p8 + 64) = make_uint4(h8[0] | ((uint32_t)h8[1] << 16), h8[2] | ((uint32_t)h8[3] << 16),
                                                              h8[4] | ((uint32_t)h8[5] << 16), h8[6] | ((uint32_t)h8[7] << 16));
            } else {
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float v0 = __uint_as_float(rr[8 * q + 2 * e]) * inv, v1 = __uint_as_float(rr[8 * q + 2 * e + 1]) * inv;
                  hi[e] = (p.out_format == 0) ? split_bf16x2(v0, v1, lo[e]) : split_f16x2(v0, v1, lo[e]);
                }
                *reinterpret_cast<uint4*>(p.out_hi + off + 16 * c16 + 8 * q) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              }
            }
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax of tile w = warpgroup index (one row per thread)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(A2_REGS_SOFTMAX));
    const int w = warp >> 2, sw = warp & 3;
    const int r = sw * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(sw * 32) << 16;
    const uint32_t s_addr = tmem_base + 128 * w + lane_sel;
    const uint32_t o_addr = tmem_base + 256 + 64 * w + lane_sel;
    const uint32_t pt_addr = tmem_base + 384 + 64 * w + lane_sel;
    uint64_t* s_full = tb(w, S_FULL);
    uint64_t* s_empty = tb(w, S_EMPTY);
    uint64_t* p_full = tb(w, P_FULL);
    uint64_t* p_empty = tb(w, P_EMPTY);
    uint64_t* pv_done = tb(w, PV_DONE);
    uint32_t g = 0, it = 0;
    // MUFU ping-pong (A2_PINGPONG, OFF): named barriers 1 (A may go) and 2 (B may go) make the exponential phases of the two
    // softmax warpgroups strictly alternate, FA3-style.  Measured SLOWER here (87.9 -> 96.5 us per layer): one warp per scheduler
    // does not saturate the MUFU on its own (its FFMA2 / pack / sum instructions interleave with the exp2s), so two overlapping
    // exponential phases use the pipe better than two serialised ones.  Kept as a compile-time switch with that finding.
    if (A2_PINGPONG && w == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      int q0, h, b;
      item_coords(item, q0, h, b);
      q0 += w * A2_BM;
      const int kv_len = item_kv_len(b);
      const int nchunks = (kv_len + A2_BN - 1) / A2_BN;
      float m_used = -INFINITY;
      float l_run = 0.0f;
      for (int j = 0; j < nchunks; ++j, ++g) {
        const uint32_t par = g & 1;
        const int key0 = j * A2_BN;
        const bool partial = key0 + A2_BN > kv_len;
        uint32_t sr[4][32];
        mbar_wait(s_full, par);
        tc_fence_after();
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) tmem_ld_32x32b_x32(s_addr + pc * 32, sr[pc]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty);
        if (partial) {
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (key0 + pc * 32 + i >= kv_len) sr[pc][i] = 0xff800000u;
          }
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              mx[c] = fmaxf(fmaxf(mx[c], __uint_as_float(sr[pc][i + 2 * c])), __uint_as_float(sr[pc][i + 2 * c + 1]));
          }
        }
        const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        if (j == 0) {
          m_used = cmax;
        } else {
          const bool grow = (cmax - m_used) * LOG2E > 8.0f;
          if (__any_sync(0xffffffffu, grow)) {
            // rare: rescale O_w (TMEM) and l by 2^(m_used - m_new)
            const float m_new = fmaxf(m_used, cmax);
            const float alpha = ex2_approx((m_used - m_new) * LOG2E);
            mbar_wait(pv_done, par ^ 1);
            tc_fence_after();
#pragma unroll
            for (int piece = 0; piece < 2; ++piece) {
              uint32_t ob[32];
              tmem_ld_32x32b_x32(o_addr + piece * 32, ob);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ob[i] = __float_as_uint(__uint_as_float(ob[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + piece * 32, ob);
            }
            tmem_st_wait();
            tc_fence_before();
            l_run *= alpha;
            m_used = m_new;
          }
        }
        const float mneg = -m_used * LOG2E;
        const uint64_t l2e2 = pack2(LOG2E, LOG2E), mneg2 = pack2(mneg, mneg);
        uint64_t sum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
        const bool drop_on = DROP && p.drop.thr16 != 0;
        const uint64_t rg = drop_on ? attn_row_group(b * p.H + h, q0 + r, p.T) + (uint64_t)(key0 >> 2) : 0;
        if (A2_PINGPONG) {
          if (w == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        // probabilities, packed to 16 bits IN PLACE as they are produced: keys [0, 64) -> sr[0][0..31], keys [64, 128) -> sr[2][0..31]
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float a[4];
            unpack2(fma2(pack2(__uint_as_float(sr[pc][i]), __uint_as_float(sr[pc][i + 1])), l2e2, mneg2), a[0], a[1]);
            unpack2(fma2(pack2(__uint_as_float(sr[pc][i + 2]), __uint_as_float(sr[pc][i + 3])), l2e2, mneg2), a[2], a[3]);
            if (!DROP && ((A2_POLY_MASK >> (i >> 2)) & 1)) {     // (the training forward keeps the backward kernel's exp2)
              ex2_poly_x2(a[0], a[1]);     // the groups of 4 keys selected by A2_POLY_MASK (one bit per group of a 32-key block) leave the MUFU alone
              ex2_poly_x2(a[2], a[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] = ex2_approx(a[e]);
            }
            sum2[(i >> 1) & 3] = add2(sum2[(i >> 1) & 3], pack2(a[0], a[1]));
            sum2[((i >> 1) + 1) & 3] = add2(sum2[((i >> 1) + 1) & 3], pack2(a[2], a[3]));
            if (drop_on) {
              const uint64_t bits = drop_bits4(p.drop, rg + pc * 8 + (i >> 2));
#pragma unroll
              for (int e = 0; e < 4; ++e) a[e] = drop_keep(bits, e, p.drop.thr16) ? a[e] * p.drop.scale : 0.0f;
            }
            const int c = (pc & 1) * 16 + (i >> 1);
            sr[pc & 2][c] = FP16 ? pack_f16x2(a[0], a[1]) : pack_bf16x2(a[0], a[1]);
            sr[pc & 2][c + 1] = FP16 ? pack_f16x2(a[2], a[3]) : pack_bf16x2(a[2], a[3]);
          }
        }
        if (A2_PINGPONG) {
          // hand the MUFU to the other warpgroup (B's very last hand-over would have no taker)
          const bool last_chunk = (j + 1 == nchunks) && (item + (int)gridDim.x >= p.n_items);
          if (w == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
          else if (!last_chunk) asm volatile("bar.arrive 1, 256;" ::: "memory");
        }
        {
          float s0, s1, s2, s3, s4, s5, s6, s7;
          unpack2(sum2[0], s0, s1);
          unpack2(sum2[1], s2, s3);
          unpack2(sum2[2], s4, s5);
          unpack2(sum2[3], s6, s7);
          l_run += ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        }
        mbar_wait(p_empty, par ^ 1);
        tc_fence_after();
        tmem_st_32x32b_x32(pt_addr, sr[0]);
        tmem_st_32x32b_x32(pt_addr + 32, sr[2]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // hand the row sums to the output warps and go on
      if (it > 0) mbar_wait(tb(w, EPI_DONE), (it - 1) & 1);
      lbuf[w * 128 + r] = l_run;
      __syncwarp();
      if (lane == 0) mbar_arrive(tb(w, L_FULL));
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 12) tmem_dealloc<512>(tmem_base);
}

template <bool FP16, bool DROP>
static int launch_attn2(const void* qkv_hi, int B, int T, int H, const int* kv_len, void* out_hi, void* out_lo, int out_format,
                        DropSpec drop, cudaStream_t stream) {
  const int d = H * A2_DH;
  CUtensorMap tm;
  const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
  const uint32_t box[3] = {A2_DH, A2_BM, 1};
  int rc = make_tmap(&tm, qkv_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  Attn2Params p;
  p.T = T;
  p.d = d;
  p.H = H;
  p.kv_len = kv_len;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  p.drop = drop;
  p.out_format = out_format;
  p.q_pairs = ((T + A2_BM - 1) / A2_BM + 1) / 2;
  p.n_items = p.q_pairs * H * B;
  auto kern = attn_fwd2_kernel<FP16, DROP>;
  static unsigned long long smem_attr_done = 0;
  W2V2_CUDA(ensure_dyn_smem(kern, A2_SMEM, smem_attr_done));
  int dev = 0, sms = 0;
  W2V2_CUDA(cudaGetDevice(&dev));
  W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  dim3 grid(p.n_items < sms ? p.n_items : sms);
  W2V2_CUDA(launch_pdl(kern, grid, dim3(A2_THREADS), (size_t)A2_SMEM, stream, 0, tm, p));
  return 0;
}

// entry used by attn.cu's dispatcher for the single-pass modes (passes 1 / 17)
int attn_fwd2(const void* qkv_hi, int B, int T, int H, const int* kv_len, void* out_hi, void* out_lo, bool fp16, int out_format,
              DropSpec drop, cudaStream_t stream) {
  if (fp16) return drop.thr16 ? launch_attn2<true, true>(qkv_hi, B, T, H, kv_len, out_hi, out_lo, out_format, drop, stream)
                              : launch_attn2<true, false>(qkv_hi, B, T, H, kv_len, out_hi, out_lo, out_format, drop, stream);
  return drop.thr16 ? launch_attn2<false, true>(qkv_hi, B, T, H, kv_len, out_hi, out_lo, out_format, drop, stream)
                    : launch_attn2<false, false>(qkv_hi, B, T, H, kv_len, out_hi, out_lo, out_format, drop, stream);
}

}  // namespace w2v2
