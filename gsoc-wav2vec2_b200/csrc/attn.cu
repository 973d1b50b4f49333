// Non-causal multi-head attention forward for sm_100a (head size 64):
//     ctx[b, t, h*64:(h+1)*64] = softmax_k( q[b,t,h,:] . k[b,k,h,:]  (+ key mask) ) . v[b,k,h,:]
// Reference: TransformerAttention.get_context, encoder.py:34-47 (scores, additive mask, softmax, PV, head
// merge) with the head split of encoder.py:49-54 and the key mask of encoder.py:256-263 (-10000 on padded
// keys, which underflows to an exact 0 probability in fp32 -> implemented as exclusion via kv_len).
// The q scaling of encoder.py:28 is folded into the q projection weights by the host.
//
// Layout: one packed activation  qkv[b, t, 0:3d] = [q | k | v]  (bf16, written by the fused QKV GEMM).
// Head h of q/k/v is a 64-column slice, so every operand tile is ONE 3-D TMA box {64, 128, 1} of the
// same tensor map - no head-split / transpose kernels, and the context is written straight into its
// merged [b, t, d] position.
//
// One CTA = 128 queries of one (b, h).  Loop over 128-key chunks:
//     S = Q K^T        tcgen05.mma 128x128x16 (x4), both operands K-major SW128         -> TMEM S
//     softmax warps:   the whole 128-column S row is pulled into registers with four tcgen05.ld in flight and ONE
//                      wait (tcgen05.ld latency is long while the tensor pipe is busy), S is released at once so
//                      the next chunk's Q K^T overlaps the exp2 / sum / bf16 packing; P -> smem (SW128 image)
//     O += P V         tcgen05.mma 128x64x16 (x8), A = P (K-major), B = V (MN-major SW128) -> TMEM O (accumulating)
//     The running max used for the exponent is only advanced when a row's true max outgrows it by more than
//     2^8 ("lazy rescaling"): then, and only then, O (TMEM) and the row sum are rescaled - rare after chunk 0.
//     Final: O / l -> bf16 hi(/lo).
// Two CTAs are resident per SM (<= 113 KB smem, 256 TMEM columns each) so one CTA's MMAs overlap the
// other's softmax.  PASSES = 3 (parity mode) adds the hi/lo cross terms for both contractions.
//
// The kernel is PERSISTENT: grid = resident CTAs (2 x SMs, 1 x SMs in the 3-pass modes), every CTA walks the tile list
// (q tile fastest, then head, then utterance) with stride gridDim.x.  The K/V ring, the S / P / O hand-shakes and their
// mbarrier phases run straight across tile boundaries: while the softmax warps finish the last chunk and write the
// context of tile i, the producer has already fetched Q and the first K/V stages of tile i+1 and the MMA warp has
// issued its Q K(0)^T - the 2.2 us prologue (barrier init, TMEM alloc, first loads) and most of the 1.6 us epilogue
// that a one-tile CTA exposes 2304 times (30 % of its lifetime, measured with %globaltimer stamps: profiles/r2_attn_fwd.md)
// are paid once per CTA.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int AT_BM = 128;      // queries per CTA
constexpr int AT_BN = 128;      // keys per chunk
constexpr int AT_DH = 64;       // head size
constexpr int AT_THREADS = 256; // warpgroup 0 (warps 0-3): softmax; warpgroup 1: warp 4 TMEM alloc + TMA, warp 5 MMA, 6-7 idle
constexpr int AT_REGS_SOFTMAX = 208;  // setmaxnreg split (2 CTAs/SM): 128 x 208 + 128 x 40 registers per CTA
constexpr int AT_REGS_CONTROL = 40;
constexpr int AT_TILE = AT_BM * AT_DH * 2;  // 16 KB: one [128][64] bf16 tile
#ifndef AT_P_IN_TMEM
#define AT_P_IN_TMEM 1   // single-pass mode: keep the probabilities in TMEM (A operand of the PV MMAs) instead of smem
#endif

template <int PASSES>
struct AttnSmem {
  static constexpr int NPL = (PASSES == 3) ? 2 : 1;         // planes (hi, lo)
  static constexpr bool P_TMEM = (PASSES == 1) && AT_P_IN_TMEM;       // no P buffer in smem: a third K/V stage instead
  static constexpr int KV_STAGES = P_TMEM ? 3 : 2;
  static constexpr int Q_OFF = 0;
  static constexpr int KV_OFF = Q_OFF + NPL * AT_TILE;
  static constexpr int KV_STAGE_BYTES = 2 * NPL * AT_TILE;  // K and V, each NPL planes
  static constexpr int P_OFF = KV_OFF + KV_STAGES * KV_STAGE_BYTES;
  static constexpr int P_BYTES = P_TMEM ? 0 : NPL * 2 * AT_TILE;   // [128][128] bf16 = two [128][64] halves per plane
  static constexpr int BAR_OFF = P_OFF + P_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128;   // 2 x (TOTAL + 1 KB reserved) must fit in 228 KB
};

struct AttnParams {
  int T;               // frames per utterance
  int d;               // hidden size (H * 64)
  const int* kv_len;   // [B] or null
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  DropSpec drop;       // attention-probability dropout of the training forward (encoder.py:42); thr16 == 0: off
  int H;
  int out_format;      // W2V2_OUT_*: 0 bf16 hi(/lo), 1 fp16 hi(/lo) of value * 2^4, 2 fp16 hi + e4m3 pair plane [rows][2 d]
  int q_tiles;         // ceil(T / 128)
  int n_tiles;         // q_tiles * H * B
};

// Optional timeline instrumentation (-DAT_STAMPS): %globaltimer stamps of thread 0 / the MMA thread per tile, read back with
// w2v2_attn_debug_stamps (tools/attn_stamps.py).  Compiled out of release builds.
#ifdef AT_STAMPS
__device__ unsigned long long g_attn_stamps[4096 * 32];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STAMPX(tile, i) do { if ((tile) < 4096) g_attn_stamps[(tile) * 32 + (i)] = gtime(); } while (0)
#define STAMP(tile, i) do { if (threadIdx.x == 0) STAMPX(tile, i); } while (0)
#else
#define STAMPX(tile, i) do { } while (0)
#define STAMP(tile, i) do { } while (0)
#endif

// FP16: q / k / v are fp16 planes of value * 2^4 (modes 17 / 19): S comes out at 2^8, O at 2^4; P is packed as fp16.
// (Round 2 tried, on this kernel: pulling S out of TMEM piece by piece with the next piece's load in flight during the current
//  piece's exponentials + per-piece lazy rescaling, and a degree-3 exp2 polynomial on the FMA pipe for every 4th pair.  Every
//  combination measured 1.65 - 1.70 ms per step against 1.46 - 1.50 for this one-wait / one-max form: the per-piece bookkeeping costs
//  more issue slots and dependent latency than the hidden TMEM load saves, and the kernel is not MUFU-throughput bound.  A third
//  variant - 64-key chunks, P written over the S columns, 128 TMEM columns and 65 KB of smem per CTA so that THREE CTAs share an
//  SM - measured 1.60 - 1.64 ms: Q K(j+1)^T can then only be issued after P V(j) has read P, and that serial chain per CTA costs
//  more than the third resident CTA hides.  ncu of this kernel: MUFU pipe 55 %, tensor pipe 27 %, issue slots 36 %, ~660 warp
//  instructions per softmax warp and chunk in ~4000 cycles - a dependent-latency problem of two warps per scheduler.)
template <int PASSES, bool FP16>
__global__ void __launch_bounds__(AT_THREADS, (PASSES == 1) ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                const AttnParams p) {
  using S = AttnSmem<PASSES>;
  constexpr int KV_STAGES = S::KV_STAGES;
  constexpr int TMEM_COLS = 256;  // S: columns [0,128), O: columns [128,192), P (single-pass mode): columns [192,256)
  constexpr bool P_IN_TMEM = (PASSES == 1) && AT_P_IN_TMEM;
  constexpr float LOG2E = FP16 ? 1.4426950408889634f / (ACT_SCALE * ACT_SCALE) : 1.4426950408889634f;   // of the SCALED score

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SW128 tiles need 1024-byte alignment (no slack is reserved)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;             // [KV_STAGES <= 3]
  uint64_t* kv_empty = bars + 4;            // [KV_STAGES <= 3]
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;
  uint64_t* pv_done = bars + 11;
  uint64_t* q_empty = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  // tile t -> (q tile, head, utterance), q tile fastest: CTAs that run together share the K / V of a head in L2
  auto tile_coords = [&](int t, int& q0, int& h, int& b) {
    q0 = (t % p.q_tiles) * AT_BM;
    const int bh = t / p.q_tiles;
    h = bh % p.H;
    b = bh / p.H;
  };
  // An utterance with NO valid key: the reference adds the same -10000 to every score (encoder.py:256-263), which the
  // softmax cancels - i.e. it attends over all T keys.  Reproduce that instead of producing exp(-inf - -inf) = NaN.
  // (kv_len is only read after pdl_wait(): the first use is inside the role loops)
  auto tile_kv_len = [&](int b) {
    const int kv_raw = (p.kv_len != nullptr) ? min(p.kv_len[b], p.T) : p.T;
    return (kv_raw <= 0) ? p.T : kv_raw;
  };

  if (warp == 4 && elect_one()) {
    tma_prefetch_desc(&tm_hi);
    if (PASSES == 3) tma_prefetch_desc(&tm_lo);
  }
  if (warp == 5 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(p_empty, 1);
    mbar_init(pv_done, 1);
    mbar_init(q_empty, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_o = tmem_base + 128;
  const uint32_t tmem_p = tmem_base + 192;

  if (warp >= 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AT_REGS_CONTROL));
  if (warp == 4) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
        int q0, h, b;
        tile_coords(t, q0, h, b);
        const int nchunks = (tile_kv_len(b) + AT_BN - 1) / AT_BN;
        mbar_wait(q_empty, (it & 1) ^ 1);      // the last Q K^T of the previous tile has read the Q buffer
        mbar_arrive_expect_tx(q_full, S::NPL * AT_TILE);
        tma_load_3d(smem + S::Q_OFF, &tm_hi, q_full, h * AT_DH, q0, b);
        if (PASSES == 3) tma_load_3d(smem + S::Q_OFF + AT_TILE, &tm_lo, q_full, h * AT_DH, q0, b);
        for (int j = 0; j < nchunks; ++j) {
          mbar_wait(&kv_empty[stage], phase ^ 1);
          uint8_t* kbuf = smem + S::KV_OFF + stage * S::KV_STAGE_BYTES;
          uint8_t* vbuf = kbuf + S::NPL * AT_TILE;
          mbar_arrive_expect_tx(&kv_full[stage], S::KV_STAGE_BYTES);
          tma_load_3d(kbuf, &tm_hi, &kv_full[stage], p.d + h * AT_DH, j * AT_BN, b);
          tma_load_3d(vbuf, &tm_hi, &kv_full[stage], 2 * p.d + h * AT_DH, j * AT_BN, b);
          if (PASSES == 3) {
            tma_load_3d(kbuf + AT_TILE, &tm_lo, &kv_full[stage], p.d + h * AT_DH, j * AT_BN, b);
            tma_load_3d(vbuf + AT_TILE, &tm_lo, &kv_full[stage], 2 * p.d + h * AT_DH, j * AT_BN, b);
          }
          if (++stage == KV_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = idesc_16bit(FP16, AT_BM, AT_BN, 0, 0);   // S = Q K^T : A, B K-major
      constexpr uint32_t idesc_pv = idesc_16bit(FP16, AT_BM, AT_DH, 0, 1);  // PV: A = P K-major, B = V MN-major
      const uint32_t q_addr = smem_u32(smem + S::Q_OFF);
      const uint32_t p_addr = smem_u32(smem + S::P_OFF);
      auto chunks_of = [&](int t) {
        int q0, h, b;
        tile_coords(t, q0, h, b);
        return (tile_kv_len(b) + AT_BN - 1) / AT_BN;
      };
      // Q K^T of one chunk; `last_of_tile`: the Q buffer may be refilled once these MMAs have retired
      auto issue_qk = [&](int stage, bool last_of_tile) {
        const uint32_t k_addr = smem_u32(smem + S::KV_OFF + stage * S::KV_STAGE_BYTES);
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
          const uint32_t qa = q_addr + ((pass == 1) ? AT_TILE : 0);
          const uint32_t ka = k_addr + ((pass == 2) ? AT_TILE : 0);
          const uint64_t dq = desc_kmajor_sw128(qa), dk = desc_kmajor_sw128(ka);
#pragma unroll
          for (int k = 0; k < AT_DH / 16; ++k) umma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, (pass | k) != 0);
        }
        umma_commit(s_full);
        if (last_of_tile) umma_commit(q_empty);
      };
      // The chunks of all tiles of this CTA form ONE sequence g = 0, 1, ...: chunk g lives in K/V stage g % KV_STAGES, its
      // S / P hand-shakes complete phase g & 1.  Iteration g issues Q K^T of chunk g + 1 (possibly the first chunk of the
      // NEXT tile, with that tile's Q) and then P V of chunk g.
      int stage = 0;                 // K/V stage of chunk g
      uint32_t kv_phase = 0;         // ... and its ring phase
      uint32_t g = 0, it = 0;
      int t = blockIdx.x;
      int nchunks = (t < p.n_tiles) ? chunks_of(t) : 0;
      if (t < p.n_tiles) {
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_qk(0, nchunks == 1);
      }
      for (; t < p.n_tiles; t += gridDim.x, ++it) {
        const int t_next = t + gridDim.x;
        const int nchunks_next = (t_next < p.n_tiles) ? chunks_of(t_next) : 0;
        for (int j = 0; j < nchunks; ++j, ++g) {
          const uint32_t par = g & 1;
          const uint32_t v_addr = smem_u32(smem + S::KV_OFF + stage * S::KV_STAGE_BYTES) + S::NPL * AT_TILE;
          int nstage = stage + 1;
          uint32_t nphase = kv_phase;
          if (nstage == KV_STAGES) {
            nstage = 0;
            nphase ^= 1;
          }
          const bool same_tile = j + 1 < nchunks;
          if (same_tile || nchunks_next > 0) {
            // S(g) has been copied to registers -> overwrite it with Q K(g+1)^T while the softmax math runs
            if (!same_tile) mbar_wait(q_full, (it + 1) & 1);   // first chunk of the next tile: its Q has landed
            mbar_wait(&kv_full[nstage], nphase);
            mbar_wait(s_empty, par);
            tc_fence_after();
            issue_qk(nstage, same_tile ? (j + 2 == nchunks) : (nchunks_next == 1));
          }
          // ---- O += P V   (the first MMA of a tile overwrites O: the softmax warps finished reading the previous tile's O
          //                  before they arrived on p_full for this chunk)
          mbar_wait(p_full, par);
          tc_fence_after();
          if (P_IN_TMEM) {
            // single-pass mode: P was written to TMEM columns [192, 256) by tcgen05.st (two bf16 per column) and is the A operand
#pragma unroll
            for (int ks = 0; ks < AT_BN / 16; ++ks) {
              const uint64_t dv = desc_mnmajor_sw128(v_addr + ks * 2048, 1024, 1024);
              umma_f16_tmem_a(tmem_o, tmem_p + ks * 8, dv, idesc_pv, (j | ks) != 0);
            }
          } else {
#pragma unroll
            for (int pass = 0; pass < PASSES; ++pass) {
              const uint32_t pa = p_addr + ((pass == 1) ? 2 * AT_TILE : 0);
              const uint32_t va = v_addr + ((pass == 2) ? AT_TILE : 0);
#pragma unroll
              for (int ks = 0; ks < AT_BN / 16; ++ks) {
                const uint64_t dp = desc_kmajor_sw128(pa + (ks >> 2) * AT_TILE) + 2 * (ks & 3);
                const uint64_t dv = desc_mnmajor_sw128(va + ks * 2048, 1024, 1024);
                umma_f16(tmem_o, dp, dv, idesc_pv, (j | pass | ks) != 0);
              }
            }
          }
          umma_commit(pv_done);
          umma_commit(p_empty);
          umma_commit(&kv_empty[stage]);
          stage = nstage;
          kv_phase = nphase;
        }
        nchunks = nchunks_next;
      }
    }
  }
  } else {
  // one CTA per SM in the 3-pass modes (launch allocation 256 registers per thread): the softmax warps take 240 and stop spilling
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"((PASSES == 1) ? AT_REGS_SOFTMAX : 240));
  {
    // ---------------------------------------------------------------- softmax / output (one row per thread)
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    uint8_t* p_hi = smem + S::P_OFF;
    uint8_t* p_lo = p_hi + 2 * AT_TILE;
    const uint32_t row_off = (uint32_t)r * 128u;
    const uint32_t swz = (uint32_t)(r & 7);
    const uint32_t s_addr = tmem_s + lane_sel;
    const uint32_t o_addr = tmem_o + lane_sel;
    uint32_t g = 0;             // chunk counter across the tiles of this CTA: the S / P hand-shakes of chunk g complete phase g & 1

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    int q0, h, b;
    tile_coords(tile, q0, h, b);
    const int kv_len = tile_kv_len(b);
    const int nchunks = (kv_len + AT_BN - 1) / AT_BN;
    float m_used = -INFINITY;   // max the exponent is taken against (may lag the true running max by < 2^8)
    float l_run = 0.0f;
    STAMP(tile, 0);
#ifdef AT_STAMPS
    if (threadIdx.x == 0 && tile < 4096) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); g_attn_stamps[tile * 32 + 15] = sm; }
#endif
    for (int j = 0; j < nchunks; ++j, ++g) {
      const uint32_t par = g & 1;
      const int key0 = j * AT_BN;
      const bool partial = key0 + AT_BN > kv_len;
      uint32_t sr[4][32];   // the whole S row of this chunk
      mbar_wait(s_full, par);
      if (j < 6) STAMP(tile, 3 + j);
      tc_fence_after();
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) tmem_ld_32x32b_x32(s_addr + pc * 32, sr[pc]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);  // S is in registers: the MMA warp may start Q K(j+1)^T
      // ---- chunk max, 4 independent 3-input chains; masked keys (ragged last chunk only) become -inf
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
          // rare: rescale O (TMEM) and l by 2^(m_used - m_new); rows that did not grow get alpha == 1
          const float m_new = fmaxf(m_used, cmax);
          const float alpha = ex2_approx((m_used - m_new) * LOG2E);
          mbar_wait(pv_done, par ^ 1);         // PV of chunk j-1 has landed in O
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
      // ---- probabilities (in place), row sum; exponent argument and sum on the packed fp32x2 pipe
      const uint64_t l2e2 = pack2(LOG2E, LOG2E), mneg2 = pack2(mneg, mneg);
      uint64_t sum2[4] = {pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f), pack2(0.f, 0.f)};
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t arg = fma2(pack2(__uint_as_float(sr[pc][i]), __uint_as_float(sr[pc][i + 1])), l2e2, mneg2);
          float a0, a1;
          unpack2(arg, a0, a1);
          a0 = ex2_approx(a0);
          a1 = ex2_approx(a1);
          sum2[(i >> 1) & 3] = add2(sum2[(i >> 1) & 3], pack2(a0, a1));
          sr[pc][i] = __float_as_uint(a0);
          sr[pc][i + 1] = __float_as_uint(a1);
        }
      }
      {
        float s0, s1, s2, s3, s4, s5, s6, s7;
        unpack2(sum2[0], s0, s1);
        unpack2(sum2[1], s2, s3);
        unpack2(sum2[2], s4, s5);
        unpack2(sum2[3], s6, s7);
        l_run += ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
      }
      if (p.drop.thr16) {
        // training: dropout on the probabilities AFTER the softmax (the row sum above is the undropped one)
        const uint64_t rg = attn_row_group(b * p.H + h, q0 + r, p.T) + (uint64_t)(key0 >> 2);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const uint64_t bits = drop_bits4(p.drop, rg + pc * 8 + i4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v = __uint_as_float(sr[pc][4 * i4 + e]);
              sr[pc][4 * i4 + e] = drop_keep(bits, e, p.drop.thr16) ? __float_as_uint(v * p.drop.scale) : 0u;
            }
          }
        }
      }
      // ---- P -> smem (bf16, SW128 K-major image) once the previous PV has consumed the buffer
      mbar_wait(p_empty, par ^ 1);
      if (P_IN_TMEM) {
        // P (bf16 pairs) -> TMEM: no smem round trip, no proxy fence; the PV MMAs read it as their A operand
        tc_fence_after();
        uint32_t pk[2][32];
#pragma unroll
        for (int c = 0; c < 64; ++c)
          pk[c >> 5][c & 31] = FP16 ? pack_f16x2(__uint_as_float(sr[c >> 4][(2 * c) & 31]), __uint_as_float(sr[c >> 4][(2 * c + 1) & 31]))
                                    : pack_bf16x2(__uint_as_float(sr[c >> 4][(2 * c) & 31]), __uint_as_float(sr[c >> 4][(2 * c + 1) & 31]));
        tmem_st_32x32b_x32(tmem_p + lane_sel, pk[0]);
        tmem_st_32x32b_x32(tmem_p + lane_sel + 32, pk[1]);
        tmem_st_wait();
        tc_fence_before();
      } else {
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t half_off = (uint32_t)(pc >> 1) * AT_TILE + row_off;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              hi[e] = FP16 ? split_f16x2(__uint_as_float(sr[pc][8 * q + 2 * e]), __uint_as_float(sr[pc][8 * q + 2 * e + 1]), lo[e])
                           : split_bf16x2(__uint_as_float(sr[pc][8 * q + 2 * e]), __uint_as_float(sr[pc][8 * q + 2 * e + 1]), lo[e]);
            const uint32_t chunk = (uint32_t)((pc & 1) * 4 + q);
            const uint32_t off = half_off + ((chunk ^ swz) << 4);
            *reinterpret_cast<uint4*>(p_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PASSES == 3) *reinterpret_cast<uint4*>(p_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor core (async proxy)
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l
    STAMP(tile, 9);
    mbar_wait(pv_done, (g - 1) & 1);
    STAMP(tile, 10);
    tc_fence_after();
    const int t = q0 + r;
    // O sits at the scale of v (2^4 in the fp16 modes); the fp16 output planes want value * 2^4 again
    const float inv = (1.0f / l_run) * ((FP16 ? 1.0f / ACT_SCALE : 1.0f) * (p.out_format != 0 ? ACT_SCALE : 1.0f));
    const size_t off = ((size_t)b * p.T + t) * p.d + (size_t)h * AT_DH;
    uint32_t ro[2][32];          // both halves of the O row in flight, one wait
    tmem_ld_32x32b_x32(o_addr, ro[0]);
    tmem_ld_32x32b_x32(o_addr + 32, ro[1]);
    tmem_ld_wait();
#pragma unroll
    for (int piece = 0; piece < 2; ++piece) {
      uint32_t (&rr)[32] = ro[piece];
      if (t < p.T) {
        if (p.out_format == 2) {
          // fp16 plane + e4m3 pair plane: this head's 64 columns are one 128-byte group of the byte plane [rows][2 d]
          uint8_t* p8 = reinterpret_cast<uint8_t*>(p.out_lo) + ((size_t)b * p.T + t) * p.d * 2 + (size_t)h * 128 + piece * 32;
#pragma unroll
          for (int q = 0; q < 2; ++q) {       // 16 columns per iteration
            uint32_t hi[8];
            uint16_t l8[8], h8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              hi[e] = split_f16_f8x2(__uint_as_float(rr[16 * q + 2 * e]) * inv, __uint_as_float(rr[16 * q + 2 * e + 1]) * inv, l8[e], h8[e]);
            *reinterpret_cast<uint4*>(p.out_hi + off + piece * 32 + 16 * q) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(p.out_hi + off + piece * 32 + 16 * q + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(p8 + 16 * q) = make_uint4(l8[0] | ((uint32_t)l8[1] << 16), l8[2] | ((uint32_t)l8[3] << 16),
                                                                l8[4] | ((uint32_t)l8[5] << 16), l8[6] | ((uint32_t)l8[7] << 16));
            *reinterpret_cast<uint4*>(p8 + 64 + 16 * q) = make_uint4(h8[0] | ((uint32_t)h8[1] << 16), h8[2] | ((uint32_t)h8[3] << 16),
                                                                     h8[4] | ((uint32_t)h8[5] << 16), h8[6] | ((uint32_t)h8[7] << 16));
          }
        } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = __uint_as_float(rr[8 * q + 2 * e]) * inv, v1 = __uint_as_float(rr[8 * q + 2 * e + 1]) * inv;
            hi[e] = (p.out_format == 0) ? split_bf16x2(v0, v1, lo[e]) : split_f16x2(v0, v1, lo[e]);
          }
          *reinterpret_cast<uint4*>(p.out_hi + off + piece * 32 + 8 * q) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (p.out_lo != nullptr)
            *reinterpret_cast<uint4*>(p.out_lo + off + piece * 32 + 8 * q) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        }
      }
    }
    tc_fence_before();   // the O reads above are ordered before this thread's next p_full arrive (-> next tile's first P V)
    STAMP(tile, 11);
    }  // tiles
  }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int PASSES, bool FP16>
static int launch_attn(const void* qkv_hi, const void* qkv_lo, int B, int T, int H, const int* kv_len, void* out_hi,
                       void* out_lo, int out_format, DropSpec drop, cudaStream_t stream) {
  using S = AttnSmem<PASSES>;
  const int d = H * AT_DH;
  CUtensorMap tm_hi, tm_lo;
  const uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
  const uint32_t box[3] = {AT_DH, AT_BM, 1};
  int rc = make_tmap(&tm_hi, qkv_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tm_lo = tm_hi;
  if (PASSES == 3) {
    rc = make_tmap(&tm_lo, qkv_lo, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  AttnParams p;
  p.T = T;
  p.d = d;
  p.kv_len = kv_len;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
  p.drop = drop;
  p.H = H;
  p.out_format = out_format;
  p.q_tiles = (T + AT_BM - 1) / AT_BM;
  p.n_tiles = p.q_tiles * H * B;
  auto kern = attn_fwd_kernel<PASSES, FP16>;
  static unsigned long long smem_attr_done = 0;   // per template instantiation, one bit per device
  W2V2_CUDA(ensure_dyn_smem(kern, S::TOTAL, smem_attr_done));
  int dev = 0, sms = 0;
  W2V2_CUDA(cudaGetDevice(&dev));
  W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // persistent: one CTA per residency slot (W2V2_ATTN_PERSIST=0: one tile per CTA, the hardware scheduler hands them out)
  static const bool persist = [] { const char* e = getenv("W2V2_ATTN_PERSIST"); return e ? atoi(e) != 0 : true; }();
  const int resident = persist ? sms * ((PASSES == 1) ? 2 : 1) : p.n_tiles;
  dim3 grid(p.n_tiles < resident ? p.n_tiles : resident);
  W2V2_CUDA(launch_pdl(kern, grid, dim3(AT_THREADS), (size_t)S::TOTAL, stream, 0, tm_hi, tm_lo, p));
  return 0;
}

}  // namespace w2v2

namespace w2v2 {
// attn2.cu: the two-tile, one-CTA-per-SM kernel of the single-pass modes
int attn_fwd2(const void* qkv_hi, int B, int T, int H, const int* kv_len, void* out_hi, void* out_lo, bool fp16, int out_format,
              DropSpec drop, cudaStream_t stream);
}

static int attn_fwd_impl(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads, int head_size,
                         const int32_t* kv_len, void* out_hi, void* out_lo, int passes, int out_format, w2v2::DropSpec drop,
                         void* stream) {
  using namespace w2v2;
  W2V2_CHECK_ARG(qkv_hi && out_hi, "null pointer");
  W2V2_CHECK_ARG(head_size == AT_DH, "only head_size == 64 is implemented (base: 768/12, large: 1024/16)");
  W2V2_CHECK_ARG(passes == 1 || passes == 3 || passes == 17 || passes == 19, "passes must be 1, 3 (bf16) or 17, 19 (fp16)");
  const int np = mode_passes(passes);
  W2V2_CHECK_ARG(np == 1 || qkv_lo, "3-pass modes need the lo planes of q / k / v");
  W2V2_CHECK_ARG(out_format >= 0 && out_format <= 2 && (out_format != 2 || out_lo), "out_format must be 0, 1 or 2 (2 writes out_lo)");
  W2V2_CHECK_ARG(batch > 0 && frames > 0 && num_heads > 0, "batch, frames, num_heads must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // single-pass modes: the two-tile kernel of attn2.cu (W2V2_ATTN_KERNEL=1 selects the one-tile kernel of this file: A/B switch)
  static const int which = [] { const char* e = getenv("W2V2_ATTN_KERNEL"); return e ? atoi(e) : 2; }();
  if (np == 1 && which == 2)
    return attn_fwd2(qkv_hi, batch, frames, num_heads, kv_len, out_hi, out_lo, (passes & 16) != 0, out_format, drop, s);
#define AT_LAUNCH(P, F) launch_attn<P, F>(qkv_hi, qkv_lo, batch, frames, num_heads, kv_len, out_hi, out_lo, out_format, drop, s)
  if (passes == 1) return AT_LAUNCH(1, false);
  if (passes == 17) return AT_LAUNCH(1, true);
  if (passes == 3) return AT_LAUNCH(3, false);
  return AT_LAUNCH(3, true);
#undef AT_LAUNCH
}

extern "C" int w2v2_attn_fwd_ex(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads,
                                int head_size, const int32_t* kv_len, void* out_hi, void* out_lo, int passes, int out_format,
                                void* stream) {
  return attn_fwd_impl(qkv_hi, qkv_lo, batch, frames, num_heads, head_size, kv_len, out_hi, out_lo, passes, out_format,
                       w2v2::make_drop(0.0f, 0, 0), stream);
}

extern "C" int w2v2_attn_fwd(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads,
                             int head_size, const int32_t* kv_len, void* out_hi, void* out_lo, int passes,
                             void* stream) {
  return attn_fwd_impl(qkv_hi, qkv_lo, batch, frames, num_heads, head_size, kv_len, out_hi, out_lo, passes,
                       (passes & 16) ? 1 : 0, w2v2::make_drop(0.0f, 0, 0), stream);
}

extern "C" int w2v2_attn_fwd_train(const void* qkv_hi, const void* qkv_lo, int batch, int frames, int num_heads,
                                   int head_size, const int32_t* kv_len, void* out_hi, void* out_lo, int passes,
                                   float drop_p, uint64_t seed, uint32_t site, void* stream) {
  if (!(drop_p >= 0.0f && drop_p < 1.0f)) return w2v2::fail(-1, "%s: drop_p must be in [0, 1)", __func__);
  return attn_fwd_impl(qkv_hi, qkv_lo, batch, frames, num_heads, head_size, kv_len, out_hi, out_lo, passes,
                       (passes & 16) ? 1 : 0, w2v2::make_drop(drop_p, seed, site), stream);
}

#ifdef AT_STAMPS
extern "C" int w2v2_attn_debug_stamps(unsigned long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, w2v2::g_attn_stamps, sizeof(unsigned long long) * n);
}
#endif
