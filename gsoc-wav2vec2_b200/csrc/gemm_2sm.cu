// cta_group::2 bf16 GEMM for sm_100a (see gemm_common.cuh for the epilogue, gemm_tcgen05.cu for the C-ABI entry).
#include "gemm_common.cuh"

namespace w2v2 {

// =====================================================================================================
// cta_group::2 variant: a CTA PAIR (two SMs of one TPC) owns a 256 x 256 output tile.  CTA r holds A rows
// [128 r, 128 r + 128) and weight rows [128 r, 128 r + 128) of the 256-wide n-tile in ITS smem; one
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) issued by the leader reads both halves and writes a
// 128 x 256 fp32 accumulator into EACH CTA's TMEM.  Per 64-deep k-block an SM ingests 16 KB (A) + 16 KB (B half)
// instead of 48 KB - the 1-SM kernel is limited by exactly that operand ingest - and the 32 KB stages allow a
// 6-deep ring.  Barriers: full[s] lives in the leader and collects both CTAs' TMA bytes; empty[s] and tmem_full[a]
// are signalled in both CTAs by multicast commits; tmem_empty[a] (leader) collects all 16 epilogue warps.
// =====================================================================================================
constexpr int GEMM2_STAGES = 5;
constexpr int GEMM2_BLOCK_N = 256;
constexpr int GEMM2_EPI_WARPS = 16;                                  // 4 TMEM lane quadrants x 4 column groups of 64
constexpr int GEMM2_THREADS = 128 + GEMM2_EPI_WARPS * 32;            // + warpgroup 0: TMA / MMA / TMEM-alloc warps
// setmaxnreg only redistributes what the CTA got at launch (640 threads x 96 regs): 128 x 32 + 512 x 112 = 61440
constexpr int GEMM2_REGS_CONTROL = 32, GEMM2_REGS_EPILOGUE = 112;

// The fp32 + residual recipe (out-proj / FFN2: x + Dense(.)) fetches its residual slabs with TMA into the per-warp
// staging blocks (two per warp, double-buffered) instead of per-lane row loads; it trades one ring stage for them.
template <int EPI>
struct Gemm2Smem {
  // (+ EPI_HI [| EPI_LO]: the same recipe also writes the operand planes and row statistics of the sum - LayerNorm-fold producer)
  static constexpr bool TMA_RES = (EPI >= 0) && ((EPI & ~(EPI_HI | EPI_LO)) == (EPI_RESID | EPI_F32 | EPI_TMARES));
  static constexpr int STAGES = GEMM2_STAGES;
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;           // 16 KB
  static constexpr int B_BYTES = (GEMM2_BLOCK_N / 2) * GEMM_BLOCK_K * 2;     // 16 KB: this CTA's half of the n-tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFF = RING_BYTES;                   // 16 epilogue warps x staging blocks (1024-aligned)
  static constexpr int EPI_BLOCK_BYTES = 2048;                 // 32 rows x 64 B, 64-byte swizzle
  static constexpr int EPI_WARP_BYTES = (TMA_RES ? 2 : 1) * EPI_BLOCK_BYTES;
  static constexpr int EPI_BYTES = GEMM2_EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int BAR_OFF = EPI_OFF + EPI_BYTES;
  static constexpr int RES_BAR_OFF = BAR_OFF + 128;            // 2 residual-slab barriers per epilogue warp
  static constexpr int BAR_BYTES = 384;
  static constexpr int BIAS_OFF = BAR_OFF + BAR_BYTES;
  // per accumulator stage: bias slice, then scale slices (the residual recipe has no scale: half the space)
  static constexpr int BIAS_BYTES = (TMA_RES ? 2 : 4) * GEMM2_BLOCK_N * 4;
  // the double staging blocks of the residual recipe only fit next to the 5-stage ring without alignment slack: that
  // instance requires (and checks) a 1024-byte aligned dynamic smem base
  static constexpr int ALIGN_SLACK = TMA_RES ? 0 : 1024;
  static constexpr int TOTAL = BIAS_OFF + BIAS_BYTES + ALIGN_SLACK;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};

// TMA bulk store of one staged [rows x 128 B] block (smem, SW128 image) to a 3-D tensor {N, rows_per_batch, batch}.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct Gemm2OutMaps {
  CUtensorMap f32, hi, lo;  // {N, rows_per_batch, batch}; boxes {16 fp32 | 32 bf16, 32 rows, 1} = 64-byte rows, SW64
  CUtensorMap res;          // fp32 residual, same geometry as f32 (TMA_RES recipe)
};

// Epilogue of one warp: TMEM lane quadrant `ew` (32 rows, lane = row) x column group `cg` (64 columns = 2 chunks).
// Both chunks are requested from TMEM before one wait, the accumulator stage is released, then bias/scale,
// erf-GELU, fp32 residual (coalesced block load through the staging block) and row masking run in registers.
// Every output leaves through a 2 KB staging block (32 rows x 64 B, 64-byte swizzle) and a TMA bulk store per
// 64-byte column slab (32 bf16 or 16 fp32 columns); rows past the utterance are clipped by the tensor map.
template <int EPI, int PASSES>
__device__ __forceinline__ void gemm2_epilogue_warp(const GemmParams& p, const Gemm2OutMaps& om, uint32_t taddr, int cg,
                                                    int n0, int t_warp0, int b, int rows_valid, bool zero_row,
                                                    const float* sb, uint8_t* stage, uint64_t* res_bar,
                                                    uint32_t tmem_empty_cluster_addr, float4 rowc) {
  constexpr int BLOCK_N = GEMM2_BLOCK_N;
  const bool f_gelu = (EPI >= 0) ? bool(EPI & EPI_GELU) : (p.gelu != 0);
  const bool f_fast = (EPI >= 0) ? bool(EPI & EPI_FASTGELU) : (p.gelu == 2);  // tanh-form GELU: single-pass mode
  const bool f_res = (EPI >= 0) ? bool(EPI & EPI_RESID) : (p.residual != nullptr);
  const bool f_f32 = (EPI >= 0) ? bool(EPI & EPI_F32) : (p.out_f32 != nullptr);
  const bool f_hi = (EPI >= 0) ? bool(EPI & EPI_HI) : (p.out_hi != nullptr);
  const bool f_lo = (EPI >= 0) ? bool(EPI & EPI_LO) : (p.out_lo != nullptr);
  const bool f_scale = (EPI >= 0) ? bool(EPI & EPI_SCALE) : (p.scale != nullptr);
  const bool f_fold = (EPI >= 0) ? bool(EPI & EPI_LNFOLD) : (p.ln_fold_stats != nullptr);
  const int lane = lane_id();
  const int c_base = 64 * cg;       // first column of this warp inside the tile
  const int n = n0 + c_base;        // ... and in the output

  uint32_t r[2][32];
  if (!W2V2_DBG(p, 3)) {
    tmem_ld_32x32b_x32(taddr + c_base, r[0]);
    tmem_ld_32x32b_x32(taddr + c_base + 32, r[1]);
    tmem_ld_wait();
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tmem_empty_cluster_addr);  // accumulator stage is free again
  if (W2V2_DBG(p, 3)) return;
  if (W2V2_DBG(p, 1)) {
    if (__uint_as_float(r[0][0] ^ r[1][31]) == 1.2345e-30f) p.out_f32[0] = 0.0f;
    return;
  }
  if (rows_valid <= 0 || n >= p.N) return;
  const size_t orow0 = (size_t)b * p.rows_per_batch + t_warp0;
  const bool f_ln = f_res && p.ln_stats != nullptr;
  const float ln_mean = rowc.x, ln_rstd = rowc.y;
  const float fold_rs = rowc.z, fold_nm = rowc.w;   // LayerNorm fold: v = acc * (rstd * acc_scale) + (-rstd * mean) * colsum + bias

  if constexpr (Gemm2Smem<EPI>::TMA_RES) {
    // out = acc + bias + residual, fp32.  Slab q (16 columns, 64-byte rows) of the residual was TMA-loaded into block
    // q & 1 (slabs 0 / 1 before the accumulator was ready, see the caller); the sum is written back IN PLACE and leaves
    // with a TMA store; as soon as that store has read the block, slab q + 2 is fetched into it.  N % 64 == 0 here.
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = q >> 1, hf = q & 1;
      const int c0 = c_base + 16 * q;
      uint8_t* blk = stage + (q & 1) * Gemm2Smem<EPI>::EPI_BLOCK_BYTES;
      auto bslot = [&](int piece) { return blk + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4); };
      mbar_wait(&res_bar[q & 1], (uint32_t)(q >> 1));   // each barrier completes exactly twice per tile
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4 rr = *reinterpret_cast<const float4*>(bslot(j));
        if (f_ln) rr = ln_of_residual(rr, ln_mean, ln_rstd, p.ln_gamma, p.ln_beta, n0 + c0 + 4 * j);
        const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + 4 * j);
        float4 o;
        o.x = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 0]), p.acc_scale, bb.x) + rr.x;   // fmaf(a, 1, b) == a + b
        o.y = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 1]), p.acc_scale, bb.y) + rr.y;
        o.z = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 2]), p.acc_scale, bb.z) + rr.z;
        o.w = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 3]), p.acc_scale, bb.w) + rr.w;
        if (zero_row) o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        *reinterpret_cast<float4*>(bslot(j)) = o;
        r[i][16 * hf + 4 * j + 0] = __float_as_uint(o.x);   // kept for the row statistics / operand planes below
        r[i][16 * hf + 4 * j + 1] = __float_as_uint(o.y);
        r[i][16 * hf + 4 * j + 2] = __float_as_uint(o.z);
        r[i][16 * hf + 4 * j + 3] = __float_as_uint(o.w);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&om.f32, blk, n + 16 * q, t_warp0, b);
        bulk_commit();
        if (q < 2) {
          bulk_wait_read0();
          mbar_arrive_expect_tx(&res_bar[q & 1], Gemm2Smem<EPI>::EPI_BLOCK_BYTES);
          tma_load_3d(blk, &om.res, &res_bar[q & 1], n + 16 * (q + 2), t_warp0, b);
        }
      }
      __syncwarp();
    }
    if (p.row_stats_out != nullptr && lane < rows_valid) {
      // (sum, sum of squares) of this row over the warp's 64 columns: the LayerNorm statistics of the NEXT GEMM's folded LayerNorm
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[i][j]);
          s1 += x;
          s2 = fmaf(x, x, s2);
        }
      }
      reinterpret_cast<float2*>(p.row_stats_out)[(size_t)(n >> 6) * ((size_t)p.batch * p.rows_per_batch) + orow0 + lane] = make_float2(s1, s2);
    }
    if (p.row_stats_out != nullptr && p.stats_final != nullptr) row_stats_finalize_last(p, orow0, lane, rows_valid);
    if constexpr ((EPI & EPI_HI) != 0) {
      // operand planes of the SUM (the un-normalised LayerNorm input the next GEMM consumes): the two staging blocks are free
      // once the four fp32 slab stores have read them.  PASSES == 2 (fp16f8): fp16 + e4m3 pair plane; else bf16 or fp16 (out_format)
      const int fmt = (PASSES == 2) ? 2 : p.out_format;
      auto pslot = [&](int blk_i, int piece) { return stage + blk_i * Gemm2Smem<EPI>::EPI_BLOCK_BYTES + lane * 64 + ((piece ^ ((lane >> 1) & 3)) << 4); };
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = __uint_as_float(r[i][8 * j + 2 * e]), v1 = __uint_as_float(r[i][8 * j + 2 * e + 1]);
            h[e] = (fmt == 0) ? split_bf16x2(v0, v1, l[e]) : split_f16x2(v0 * ACT_SCALE, v1 * ACT_SCALE, l[e]);
          }
          *reinterpret_cast<uint4*>(pslot(i, j)) = make_uint4(h[0], h[1], h[2], h[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&om.hi, stage, n, t_warp0, b);
        tma_store_3d(&om.hi, stage + Gemm2Smem<EPI>::EPI_BLOCK_BYTES, n + 32, t_warp0, b);
        bulk_commit();
      }
      if constexpr (PASSES == 2 && (EPI & EPI_LO) != 0) {
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {      // 16 columns -> 16 bytes
            uint16_t l8[8], h8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              split_f16_f8x2(__uint_as_float(r[j >> 1][16 * (j & 1) + 2 * e]) * ACT_SCALE,
                             __uint_as_float(r[j >> 1][16 * (j & 1) + 2 * e + 1]) * ACT_SCALE, l8[e], h8[e]);
            const uint16_t* s8 = half == 0 ? l8 : h8;
            *reinterpret_cast<uint4*>(pslot(half, j)) = make_uint4(s8[0] | ((uint32_t)s8[1] << 16), s8[2] | ((uint32_t)s8[3] << 16),
                                                                   s8[4] | ((uint32_t)s8[5] << 16), s8[6] | ((uint32_t)s8[7] << 16));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&om.lo, stage, 2 * n, t_warp0, b);
          tma_store_3d(&om.lo, stage + Gemm2Smem<EPI>::EPI_BLOCK_BYTES, 2 * n + 64, t_warp0, b);
          bulk_commit();
        }
      }
      __syncwarp();
    }
    return;
  }

  // staging block: row-major 64-byte rows, 16-byte slot s of row r lives at slot s ^ ((r >> 1) & 3)  (SWIZZLE_64B)
  auto slot = [&](int row, int piece) { return stage + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4); };
  auto stage_acquire = [&]() {  // the block may still be read by the previous bulk store of this warp
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
  };
  auto stage_store = [&](const CUtensorMap* m, int col) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(m, stage, col, t_warp0, b);
      bulk_commit();
    }
  };

  // ---- pass A: scale / bias / GELU / residual / mask, in place, 16 columns (one 64-byte fp32 slab) at a time
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = q >> 1, hf = q & 1;
    const int c0 = c_base + 16 * q;
    if (n0 + c0 >= p.N) continue;
    float4 rr[4];
    if (f_res && lane < rows_valid) {
      // this row's 64-byte residual slab (pulled into L2 by the per-tile prefetch)
      const float4* gres = reinterpret_cast<const float4*>(p.residual + (orow0 + lane) * p.N + n0 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rr[j] = __ldg(gres + j);
        if (f_ln) rr[j] = ln_of_residual(rr[j], ln_mean, ln_rstd, p.ln_gamma, p.ln_beta, n0 + c0 + 4 * j);
      }
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + 4 * j);
      if (f_fold) {
        // packed fp32x2: 2 FFMA2 per column pair (same issue count as the plain bias add of the unfolded epilogue)
        const float4 cs = *reinterpret_cast<const float4*>(sb + 2 * BLOCK_N + c0 + 4 * j);
        const uint64_t rs2 = pack2(fold_rs, fold_rs), nm2 = pack2(fold_nm, fold_nm);
        unpack2(fma2(pack2(__uint_as_float(r[i][16 * hf + 4 * j + 0]), __uint_as_float(r[i][16 * hf + 4 * j + 1])), rs2,
                     fma2(nm2, pack2(cs.x, cs.y), pack2(bb.x, bb.y))), v[4 * j + 0], v[4 * j + 1]);
        unpack2(fma2(pack2(__uint_as_float(r[i][16 * hf + 4 * j + 2]), __uint_as_float(r[i][16 * hf + 4 * j + 3])), rs2,
                     fma2(nm2, pack2(cs.z, cs.w), pack2(bb.z, bb.w))), v[4 * j + 2], v[4 * j + 3]);
      } else if (f_scale) {
        const float4 sc = *reinterpret_cast<const float4*>(sb + 2 * BLOCK_N + c0 + 4 * j);
        v[4 * j + 0] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 0]), sc.x, bb.x);
        v[4 * j + 1] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 1]), sc.y, bb.y);
        v[4 * j + 2] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 2]), sc.z, bb.z);
        v[4 * j + 3] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 3]), sc.w, bb.w);
      } else {
        // fmaf(a, 1, b) == a + b bit for bit; packed: one FFMA2 per column pair
        const uint64_t as2 = pack2(p.acc_scale, p.acc_scale);
        unpack2(fma2(pack2(__uint_as_float(r[i][16 * hf + 4 * j + 0]), __uint_as_float(r[i][16 * hf + 4 * j + 1])), as2, pack2(bb.x, bb.y)),
                v[4 * j + 0], v[4 * j + 1]);
        unpack2(fma2(pack2(__uint_as_float(r[i][16 * hf + 4 * j + 2]), __uint_as_float(r[i][16 * hf + 4 * j + 3])), as2, pack2(bb.z, bb.w)),
                v[4 * j + 2], v[4 * j + 3]);
      }
    }
    if (f_gelu) {
      if (EPI < 0 && p.gelu == 3) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_tanh_tf(v[j]);
      } else if (f_fast) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) gelu_x2<true>(v[j], v[j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 2) gelu_x2<false>(v[j], v[j + 1]);
      }
    }
    if (EPI < 0 && p.drop.thr16) epilogue_dropout16(p.drop, (orow0 + lane) * p.N + n0 + c0, v);
    if (f_res && lane < rows_valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[4 * j + 0] += rr[j].x;
        v[4 * j + 1] += rr[j].y;
        v[4 * j + 2] += rr[j].z;
        v[4 * j + 3] += rr[j].w;
      }
    }
    if (EPI < 0 && p.row_replace != nullptr && lane < rows_valid && p.row_replace[orow0 + lane]) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __ldg(p.row_value + n0 + c0 + j);   // N % 8 == 0 and n0 + c0 < N: a full 16-column slab unless N % 16
    }
    if (p.row_valid != nullptr) {     // uniform: without a frame mask no per-element select is issued
#pragma unroll
      for (int j = 0; j < 16; ++j) r[i][16 * hf + j] = zero_row ? 0u : __float_as_uint(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[i][16 * hf + j] = __float_as_uint(v[j]);
    }
  }

  if (EPI < 0 && p.row_stats_out != nullptr && lane < rows_valid) {
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = __uint_as_float(r[i][j]);
        s1 += x;
        s2 = fmaf(x, x, s2);
      }
    }
    reinterpret_cast<float2*>(p.row_stats_out)[(size_t)(n >> 6) * ((size_t)p.batch * p.rows_per_batch) + orow0 + lane] = make_float2(s1, s2);
  }
  if (EPI < 0 && p.row_stats_out != nullptr && p.stats_final != nullptr) row_stats_finalize_last(p, orow0, lane, rows_valid);

  // ---- pass B: outputs, one 64-byte column slab per bulk store
  if (f_f32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (n + 16 * q >= p.N) continue;
      const int i = q >> 1, hf = q & 1;
      stage_acquire();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(slot(lane, j)) = make_uint4(r[i][16 * hf + 4 * j], r[i][16 * hf + 4 * j + 1],
                                                              r[i][16 * hf + 4 * j + 2], r[i][16 * hf + 4 * j + 3]);
      stage_store(&om.f32, n + 16 * q);
    }
  }
  if (f_hi) {
    const int fmt = p.out_format;   // 0: bf16 hi(/lo); 1: fp16 hi(/lo) of value * 2^4; 2: fp16 hi + e4m3 pair plane
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
      if (plane == 1 && !f_lo) continue;
      if (fmt == 2 && plane == 1) {
        // e4m3 pair plane [rows][2 N] bytes: this warp's 64 columns are ONE 128-byte group = slab 0: e4m3((v - hi) 2^6) of the
        // 64 columns, slab 1: e4m3(hi 2^-6); each slab is a 64-byte-row staging block and one bulk store
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          stage_acquire();
#pragma unroll
          for (int j = 0; j < 4; ++j) {      // 16 columns -> 16 bytes
            uint16_t l8[8], h8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              split_f16_f8x2(__uint_as_float(r[j >> 1][16 * (j & 1) + 2 * e]) * ACT_SCALE,
                             __uint_as_float(r[j >> 1][16 * (j & 1) + 2 * e + 1]) * ACT_SCALE, l8[e], h8[e]);
            const uint16_t* s8 = half == 0 ? l8 : h8;
            *reinterpret_cast<uint4*>(slot(lane, j)) = make_uint4(s8[0] | ((uint32_t)s8[1] << 16), s8[2] | ((uint32_t)s8[3] << 16),
                                                                  s8[4] | ((uint32_t)s8[5] << 16), s8[6] | ((uint32_t)s8[7] << 16));
          }
          stage_store(&om.lo, 2 * n + 64 * half);
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (n + 32 * i >= p.N) continue;
        stage_acquire();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = __uint_as_float(r[i][8 * j + 2 * e]), v1 = __uint_as_float(r[i][8 * j + 2 * e + 1]);
            h[e] = (fmt == 0) ? split_bf16x2(v0, v1, l[e]) : split_f16x2(v0 * ACT_SCALE, v1 * ACT_SCALE, l[e]);
          }
          *reinterpret_cast<uint4*>(slot(lane, j)) = plane == 0 ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(l[0], l[1], l[2], l[3]);
        }
        stage_store(plane == 0 ? &om.hi : &om.lo, n + 32 * i);
      }
    }
  }
}

template <int PASSES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                     const __grid_constant__ Gemm2OutMaps om, const GemmParams p) {
  using S = Gemm2Smem<EPI>;
  constexpr int BLOCK_N = GEMM2_BLOCK_N;
  constexpr int ACC_STAGES = 2;
  constexpr int STAGES = S::STAGES;
  constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (S::ALIGN_SLACK == 0) {
    smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();   // SW128 tiles need 1024-byte alignment
  }
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  float* s_bias = reinterpret_cast<float*>(smem + S::BIAS_OFF);

  const int warp = threadIdx.x >> 5;
  const int total_kb = PASSES * p.num_kb;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total_m_tiles = p.batch * p.tiles_per_batch;
  const int total_work = ((total_m_tiles + 1) / 2) * p.n_tiles;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (PASSES != 1) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's arrive.expect_tx covers both CTAs' bytes (used in the leader only)
      mbar_init(&empty_bar[i], 1);  // leader's multicast commit
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);                       // leader's multicast commit
      mbar_init(&tmem_empty[i], 2 * GEMM2_EPI_WARPS);    // every epilogue warp of both CTAs (used in the leader only)
    }
    if (S::TMA_RES) {
      uint64_t* rb = reinterpret_cast<uint64_t*>(smem + S::RES_BAR_OFF);
      for (int i = 0; i < 2 * GEMM2_EPI_WARPS; ++i) mbar_init(&rb[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are visible
  if (warp < 4) {
    // warpgroup 0 (TMA / MMA / TMEM-alloc warps) gives registers away ...
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM2_REGS_CONTROL));
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs)
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int w = pair_id; w < total_work; w += num_pairs) {
          const int n_tile = w % p.n_tiles;
          const int m_tile = min((w / p.n_tiles) * 2 + crank, total_m_tiles - 1);
          const int b = m_tile / p.tiles_per_batch;
          const int t0 = (m_tile - b * p.tiles_per_batch) * GEMM_BLOCK_M;
          const int n0 = n_tile * BLOCK_N + crank * (BLOCK_N / 2);
          for (int it = 0; it < total_kb; ++it) {
            const int pass = (PASSES == 1) ? 0 : it / p.num_kb;
            const int kb = it - pass * p.num_kb;
            // PASSES == 2 (fp16f8): pass 0 = fp16 planes, pass 1 = the e4m3 pair planes of BOTH operands - byte tensors whose
            // 128-byte k-block holds [lo8 x 64 | hi8 x 64] (A) / [hi8 x 64 | lo8 x 64] (W), i.e. the two cross terms as ONE K = 128 product
            const bool f8pass = (PASSES == 2) && pass == 1;
            const CUtensorMap* ma = (pass == 1) ? &tmA_lo : &tmA_hi;
            const CUtensorMap* mb = (pass == 2 || f8pass) ? &tmB_lo : &tmB_hi;
            const int kw = f8pass ? 2 * GEMM_BLOCK_K : GEMM_BLOCK_K;     // k-block width in elements of that plane
            int kc = kb * kw, trow = t0;
            if (kb >= p.kb_split) {
              kc = (kb - p.kb_split) * kw;
              trow = t0 + 1;
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * S::STAGE_BYTES;
            uint8_t* sb = sa + S::A_BYTES;
            const uint32_t lead_full = mapa_cluster(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);  // both CTAs' TMA bytes land here
            tma_load_3d_2sm(sa, ma, lead_full, kc, trow, b);
            tma_load_2d_2sm(sb, mb, lead_full, kb * kw, n0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA only)
      if (leader && elect_one()) {
        const uint32_t idesc = idesc_16bit(p.fp16 != 0, 2 * GEMM_BLOCK_M, BLOCK_N, 0, 0);
        constexpr uint32_t idesc8 = idesc_fmt0(2 * GEMM_BLOCK_M, BLOCK_N, 0, 0);     // e4m3 x e4m3 under kind::f8f6f4
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int w = pair_id; w < total_work; w += num_pairs) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int it = 0; it < total_kb; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
            const uint64_t da = desc_kmajor_sw128(sa);
            const uint64_t db = desc_kmajor_sw128(sa + S::A_BYTES);
            if (PASSES == 2 && it >= p.num_kb) {
              // 128 e4m3 per 128-byte swizzle row: 4 MMAs of K = 32 (32 bytes per step, like 16 fp16)
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f8_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc8, 1);
            } else {
#pragma unroll
              for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
            }
            umma_commit_2sm_mcast(&empty_bar[stage], 3);  // slot free in both CTAs once these MMAs retire
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_2sm_mcast(&tmem_full[acc], 3);  // both CTAs' epilogues may read their accumulator half
          if (++acc == ACC_STAGES) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ... to the 16 epilogue warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GEMM2_REGS_EPILOGUE));
    const int ew = warp & 3;          // TMEM lane quadrant
    const int cg = (warp - 4) >> 2;   // column group (64 columns)
    const int lane = lane_id();
    const int et = threadIdx.x - 128;  // 0..511
    uint8_t* stage = smem + S::EPI_OFF + (warp - 4) * S::EPI_WARP_BYTES;
    uint64_t* res_bar = reinterpret_cast<uint64_t*>(smem + S::RES_BAR_OFF) + 2 * (warp - 4);
    const bool f_scale = (EPI >= 0) ? bool(EPI & EPI_SCALE) : (p.scale != nullptr);
    // bias (threads 0..255) / scale (threads 256..511) slice of a tile: fetched ONE TILE AHEAD into a register and
    // parked in smem (double-buffered with the accumulator stage) so that its global-load latency never sits on the
    // per-tile critical path
    const int bcol = et & (BLOCK_N - 1);
    auto fetch_bias = [&](int w) -> float {
      if (w >= total_work) return 0.0f;
      const int n0 = (w % p.n_tiles) * BLOCK_N;
      const int m_tile = min((w / p.n_tiles) * 2 + crank, total_m_tiles - 1);
      const size_t boff = (size_t)(m_tile / p.tiles_per_batch) * p.bias_bstride + n0 + bcol;
      if (et < BLOCK_N) return (p.bias != nullptr && n0 + bcol < p.N) ? __ldg(p.bias + boff) : 0.0f;
      return (f_scale && n0 + bcol < p.N) ? __ldg(p.scale + boff) : 1.0f;
    };
    auto park_bias = [&](int acc, float v) {
      float* sb = s_bias + acc * BLOCK_N;
      if (et < BLOCK_N) sb[bcol] = v;
      else if (f_scale) sb[2 * BLOCK_N + bcol] = v;
    };
    const bool f_res = (EPI >= 0) ? bool(EPI & EPI_RESID) : (p.residual != nullptr);
    int acc = 0;
    uint32_t acc_phase = 0;
    park_bias(0, fetch_bias(pair_id));
    asm volatile("bar.sync 1, 512;" ::: "memory");
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const int n_tile = w % p.n_tiles;
      const int m_raw = (w / p.n_tiles) * 2 + crank;
      const int m_tile = min(m_raw, total_m_tiles - 1);
      const int b = m_tile / p.tiles_per_batch;
      const int t_warp0 = (m_tile - b * p.tiles_per_batch) * GEMM_BLOCK_M + ew * 32;
      const int n0 = n_tile * BLOCK_N;
      const int rows_valid = (m_raw < total_m_tiles) ? min(32, p.rows_per_batch - t_warp0) : 0;
      const bool zero_row = p.row_valid != nullptr && t_warp0 + lane >= p.row_valid[b];
      const float* sb = s_bias + acc * BLOCK_N;
      const float next_bias = fetch_bias(w + num_pairs);   // in flight during this tile's epilogue
      if (S::TMA_RES) {
        // residual slabs 0 and 1 of this warp's 32 x 64 slice: in flight while the MMAs of this tile run
        if (lane == 0 && rows_valid > 0 && n0 + 64 * cg < p.N) {
          bulk_wait_read0();   // the previous tile's stores have finished reading both staging blocks
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            mbar_arrive_expect_tx(&res_bar[q], S::EPI_BLOCK_BYTES);
            tma_load_3d(stage + q * S::EPI_BLOCK_BYTES, &om.res, &res_bar[q], n0 + 64 * cg + 16 * q, t_warp0, b);
          }
        }
      } else if (f_res && lane < rows_valid && n0 + 64 * cg < p.N) {
        const float* rp = p.residual + ((size_t)b * p.rows_per_batch + t_warp0 + lane) * p.N + n0 + 64 * cg;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 32));
      }
      // per-row LayerNorm constants (residual LayerNorm / folded LayerNorm): their loads fly while this tile's MMAs run
      const float4 rowc = epilogue_row_constants<EPI>(p, (size_t)b * p.rows_per_batch + t_warp0 + lane, lane < rows_valid);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
      gemm2_epilogue_warp<EPI, PASSES>(p, om, taddr, cg, n0, t_warp0, b, rows_valid, zero_row, sb, stage, res_bar,
                               mapa_cluster(smem_u32(&tmem_empty[acc]), 0), rowc);
      park_bias(acc ^ 1, next_bias);
      asm volatile("bar.sync 1, 512;" ::: "memory");   // every warp is done with this tile's slice; the next one is visible
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (lane == 0) bulk_wait0();  // all bulk stores of this warp have completed before the CTA may exit
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

template <int PASSES, int EPI>
static int launch_gemm_2sm_t(const w2v2_gemm_args* a, cudaStream_t stream) {
  using S = Gemm2Smem<EPI>;
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  const uint64_t a_dims[3] = {(uint64_t)a->a_row_len, (uint64_t)a->a_rows, (uint64_t)a->batch};
  const uint64_t a_strides[2] = {(uint64_t)a->a_row_stride * 2, (uint64_t)a->a_batch_stride * 2};
  const uint32_t a_box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
  int rc = make_tmap(&tmA_hi, a->a_hi, 3, a_dims, a_strides, a_box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tmA_lo = tmA_hi;
  if (PASSES == 3 && (rc = make_tmap(&tmA_lo, a->a_lo, 3, a_dims, a_strides, a_box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const uint64_t b_dims[2] = {(uint64_t)a->K, (uint64_t)a->w_rows};
  const uint64_t b_strides[1] = {(uint64_t)a->K * 2};
  const uint32_t b_box[2] = {GEMM_BLOCK_K, (uint32_t)(GEMM2_BLOCK_N / 2)};
  rc = make_tmap(&tmB_hi, a->w_hi, 2, b_dims, b_strides, b_box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tmB_lo = tmB_hi;
  if (PASSES == 3 && (rc = make_tmap(&tmB_lo, a->w_lo, 2, b_dims, b_strides, b_box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if (PASSES == 2) {
    // e4m3 pair planes: two bytes per element, so the SAME byte strides as the 16-bit planes; 128-byte k-blocks
    const uint64_t a8_dims[3] = {(uint64_t)a->a_row_len * 2, (uint64_t)a->a_rows, (uint64_t)a->batch};
    const uint32_t a8_box[3] = {2 * GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
    if ((rc = make_tmap(&tmA_lo, a->a_lo, 3, a8_dims, a_strides, a8_box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_UINT8))) return rc;
    const uint64_t b8_dims[2] = {(uint64_t)a->K * 2, (uint64_t)a->w_rows};
    const uint32_t b8_box[2] = {2 * GEMM_BLOCK_K, (uint32_t)(GEMM2_BLOCK_N / 2)};
    if ((rc = make_tmap(&tmB_lo, a->w_lo, 2, b8_dims, b_strides, b8_box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_UINT8))) return rc;
  }
  GemmParams p = make_gemm_params(a, GEMM2_BLOCK_N);
  // output tensor maps {N, rows_per_batch, batch}: rows past an utterance are clipped by the TMA store
  Gemm2OutMaps om;
  memset(&om, 0, sizeof(om));
  {
    const uint64_t c_dims[3] = {(uint64_t)a->N, (uint64_t)a->rows_per_batch, (uint64_t)a->batch};
    const uint64_t s16[2] = {(uint64_t)a->N * 2, (uint64_t)a->rows_per_batch * a->N * 2};
    const uint64_t s32[2] = {(uint64_t)a->N * 4, (uint64_t)a->rows_per_batch * a->N * 4};
    const uint32_t box16[3] = {32, 32, 1}, box32[3] = {16, 32, 1};   // 64-byte rows
    if (a->out_f32 && (rc = make_tmap(&om.f32, a->out_f32, 3, c_dims, s32, box32, CU_TENSOR_MAP_SWIZZLE_64B,
                                      CU_TENSOR_MAP_DATA_TYPE_FLOAT32))) return rc;
    if (S::TMA_RES && (rc = make_tmap(&om.res, a->residual, 3, c_dims, s32, box32, CU_TENSOR_MAP_SWIZZLE_64B,
                                      CU_TENSOR_MAP_DATA_TYPE_FLOAT32))) return rc;
    if (a->out_hi && (rc = make_tmap(&om.hi, a->out_hi, 3, c_dims, s16, box16, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if (a->out_lo && a->out_format == 2) {   // byte plane [rows][2 N]: 64-byte slabs
      const uint64_t c8_dims[3] = {(uint64_t)a->N * 2, (uint64_t)a->rows_per_batch, (uint64_t)a->batch};
      const uint32_t box8[3] = {64, 32, 1};
      if ((rc = make_tmap(&om.lo, a->out_lo, 3, c8_dims, s16, box8, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_DATA_TYPE_UINT8))) return rc;
    } else if (a->out_lo && (rc = make_tmap(&om.lo, a->out_lo, 3, c_dims, s16, box16, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
  }
  auto kern = gemm_bf16_2sm_kernel<PASSES, EPI>;
  static unsigned long long smem_attr_done = 0;   // per template instantiation, one bit per device
  W2V2_CUDA(ensure_dyn_smem(kern, S::TOTAL, smem_attr_done));
  int dev = 0, sms = 0;
  W2V2_CUDA(cudaGetDevice(&dev));
  W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total_m_tiles = p.batch * p.tiles_per_batch;
  const int total_work = ((total_m_tiles + 1) / 2) * p.n_tiles;
  int grid = total_work * 2 < sms ? total_work * 2 : sms;
  if (a->max_ctas > 0 && grid > a->max_ctas) grid = a->max_ctas;
  grid -= grid % 2;
  if (grid < 2) grid = 2;
  W2V2_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM2_THREADS), (size_t)S::TOTAL, stream, 0, tmA_hi, tmA_lo, tmB_hi, tmB_lo, om, p));
  return 0;
}

// Picks the compile-time epilogue recipe matching the call (the combinations the Wav2Vec2 forward uses);
// anything else runs the generic run-time-flag instance.
template <int PASSES, bool FASTG>
static int dispatch_2sm(const w2v2_gemm_args* a, cudaStream_t s) {
  const bool gelu = (a->flags & W2V2_GEMM_GELU) != 0, res = a->residual != nullptr, sc = a->scale != nullptr;
  const bool f32 = a->out_f32 != nullptr, hi = a->out_hi != nullptr, lo = a->out_lo != nullptr;
  constexpr int LO = (PASSES != 1) ? EPI_LO : 0;     // the model writes two planes exactly in the multi-plane modes
  // bf16 single-pass mode: bf16-grade tanh-form GELU; every mode that is more precise than that keeps the erf-exact form
  constexpr int G = EPI_GELU | (FASTG ? EPI_FASTGELU : 0);
  // the rarely used epilogue options (tf-approximate GELU, dropout, SpecAugment row replacement) only exist in the run-time instance
  if ((a->flags & W2V2_GEMM_GELU_TANH) || a->row_replace_mask != nullptr || a->drop_p > 0.0f)
    return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
  // row statistics are written by the TMA-residual recipes and by the run-time instance only
  if (a->row_stats_out != nullptr && !(!gelu && res && f32 && a->N % 64 == 0 && a->ln_fold_stats == nullptr && !sc))
    return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
  if (a->ln_fold_stats != nullptr) {   // LayerNorm folded into the GEMM (QKV, FFN1): the scale slot carries the column sums
    if (!res && !f32 && hi && lo == (PASSES != 1) && a->row_stats_out == nullptr) {
      if (gelu) return launch_gemm_2sm_t<PASSES, EPI_LNFOLD | EPI_SCALE | G | EPI_HI | LO>(a, s);
      return launch_gemm_2sm_t<PASSES, EPI_LNFOLD | EPI_SCALE | EPI_HI | LO>(a, s);
    }
    return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
  }
  if (sc) {
    if (gelu && !res && !f32 && hi && lo == (PASSES != 1)) return launch_gemm_2sm_t<PASSES, EPI_SCALE | G | EPI_HI | LO>(a, s);
    return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
  }
  if (lo == (PASSES != 1) || !hi) {
    if (gelu && !res && !f32 && hi) return launch_gemm_2sm_t<PASSES, G | EPI_HI | LO>(a, s);
    if (gelu && !res && f32 && !hi) return launch_gemm_2sm_t<PASSES, G | EPI_F32>(a, s);
    if (!gelu && !res && f32 && !hi) return launch_gemm_2sm_t<PASSES, EPI_F32>(a, s);
    if (!gelu && !res && f32 && hi) return launch_gemm_2sm_t<PASSES, EPI_F32 | EPI_HI | LO>(a, s);
    if (!gelu && !res && !f32 && hi) return launch_gemm_2sm_t<PASSES, EPI_HI | LO>(a, s);
    // residual slabs fetched by TMA into double staging blocks (W2V2_RES_TMA_MAXK bounds K; default: every K)
    static const int maxk = [] { const char* e = getenv("W2V2_RES_TMA_MAXK"); return e ? atoi(e) : (1 << 30); }();
    if (!gelu && res && f32 && !hi && a->N % 64 == 0 && a->K * PASSES <= maxk)
      return launch_gemm_2sm_t<PASSES, EPI_RESID | EPI_F32 | EPI_TMARES>(a, s);
  }
  // fp32 sum + its operand planes + row statistics (the producer side of a folded LayerNorm): same recipe, planes written last
  if (PASSES != 3 && !gelu && res && f32 && hi && lo == (PASSES == 2) && a->N % 64 == 0 && (PASSES != 2 || a->out_format == 2)) {
    return launch_gemm_2sm_t<PASSES, EPI_RESID | EPI_F32 | EPI_TMARES | EPI_HI | ((PASSES == 2) ? EPI_LO : 0)>(a, s);
  }
  if (lo == (PASSES != 1) || !hi) {
    if (!gelu && res && f32 && !hi && a->row_stats_out == nullptr) return launch_gemm_2sm_t<PASSES, EPI_RESID | EPI_F32>(a, s);
  }
  return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
}

thread_local bool g_stats_final_in_kernel = false;   // set when the launched kernel finalises row_stats_final itself

int launch_gemm_2sm(const w2v2_gemm_args* a, cudaStream_t stream) {
  g_stats_final_in_kernel = (a->row_stats_final != nullptr);
  if (mode_f8(a->passes)) return dispatch_2sm<2, false>(a, stream);
  if (mode_passes(a->passes) == 3) return dispatch_2sm<3, false>(a, stream);
  return mode_fp16(a->passes) ? dispatch_2sm<1, false>(a, stream) : dispatch_2sm<1, true>(a, stream);
}

}  // namespace w2v2
