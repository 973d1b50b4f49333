// Shared pieces of the sm_100a bf16 GEMM kernels (gemm_tcgen05.cu: 1-SM tiles, gemm_2sm.cu: cta_group::2 pairs):
// tile constants, kernel parameters and the warp-level epilogue.
#pragma once
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

// The epilogue ablation switches (GemmParams::debug) are profiling aids: release builds compile them out
#ifdef W2V2_GEMM_DEBUG
#define W2V2_DBG(p, v) ((p).debug == (v))
#else
#define W2V2_DBG(p, v) false
#endif

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_STAGES_MAX = 4;
constexpr int GEMM_EPI_STAGE_BYTES = 8 * 4096;  // 8 epilogue warps x (32 rows x 128 B) transpose buffers
constexpr int GEMM_THREADS = 384;  // warpgroup 0: TMA / MMA / TMEM-alloc warps; warpgroups 1-2: epilogue
constexpr int GEMM_REGS_CONTROL = 56;   // setmaxnreg split: 128 x 56 + 256 x 224 <= 64K registers
constexpr int GEMM_REGS_EPILOGUE = 224;

// epilogue recipe bits (template parameter EPI of the kernels; EPI < 0 = decide from GemmParams at run time)
constexpr int EPI_GELU = 1, EPI_RESID = 2, EPI_F32 = 4, EPI_HI = 8, EPI_LO = 16, EPI_SCALE = 32,
              EPI_FASTGELU = 64,     // tanh-form GELU (single-pass mode only; see gelu_tanh_p2)
              EPI_TMARES = 128,    // 2-SM kernel: residual slabs fetched by TMA into the staging blocks
              EPI_LNFOLD = 256;    // the A operand is the UN-normalised LayerNorm input: the epilogue applies mean / rstd per row
                                   // (needs EPI_SCALE: the scale slot then carries colsum(gamma o W) instead of a multiplier)
constexpr int EPI_RUNTIME = -1;

struct GemmParams {
  int num_kb;           // K / 64 (per pass)
  int kb_split;         // k-blocks >= kb_split are fetched from (k - kb_split*64, row + 1)  [pair-row conv fallback]
  int rows_per_batch;   // valid rows per batch entry
  int tiles_per_batch;  // ceil(rows_per_batch / 128)
  int batch;
  int n_tiles;          // ceil(N / BLOCK_N)
  int N;                // valid output columns == leading dimension of every output / residual
  int gelu;            // 0 = none, 1 = erf-exact, 2 = tanh-form fit of the erf GELU (single-pass mode), 3 = tf "approximate" GELU
  int vec_ok;           // N % 8 == 0: 16-byte vector stores are aligned
  int debug;            // profiling aid, only compiled with -DW2V2_GEMM_DEBUG: 1 = epilogue only drains TMEM, 2 = no global stores, 3 = no TMEM read
  int atomic_f32;       // 1: out_f32 += result with fp32 atomics (split-K partial sums into a zeroed buffer)
  int mn_major;         // 1: both operands are MN-major (reduction over the ROW index of two row-major matrices: wgrad)
  const float* bias;      // [N] (or [batch][N] with bias_bstride = N) or null
  const float* scale;     // optional per-column scale applied before the bias, [N] or [batch][N]
  int bias_bstride;       // elements between batch entries of bias / scale (0 = shared)
  const float* residual;  // fp32 [batch*rows_per_batch, N] or null
  const float* ln_stats;  // optional [rows][2] (mean, rstd): the residual term is LayerNorm(residual) recomputed on the fly
  const float* ln_gamma;  // [N]
  const float* ln_beta;   // [N]
  const int* row_valid;   // [batch] or null: rows >= row_valid[b] are written as zeros
  const uint8_t* row_replace;  // [batch*rows_per_batch] or null: rows with a non-zero byte are written as row_value[0:N] (SpecAugment)
  const float* row_value;      // [N]
  DropSpec drop;               // dropout on the activation (after bias / GELU, before the residual); thr16 == 0: off
  // LayerNorm folded into this GEMM (w2v2.h ln_fold_*): out = rstd_r * acc - rstd_r * mean_r * colsum[n] + bias[n]; (mean, rstd) of
  // row r come from partial (sum, sum of squares) pairs over 64-column groups of the K input columns
  const float* ln_fold_stats;  // [ln_fold_parts][rows][2] or null
  int ln_fold_parts;
  float ln_fold_inv_dim;       // 1 / K
  float ln_eps;
  float* row_stats_out;        // optional [N / 64][rows][2]: (sum, sum of squares) of the fp32 output per 64-column group
  float2* stats_final;         // optional [rows]: (mean, rstd) over all N columns, written by the last column group of a 32-row block
  unsigned* stats_counter;     // [ceil(rows / 32)] arrival counters of those blocks (zero between launches)
  int stats_parts;             // N / 64
  int res_ln_parts;            // 0: ln_stats holds (mean, rstd) per row; > 0: partial sums over N columns in that many groups
  int fp16;               // 1: the 16-bit operand planes are fp16 (idesc format 0) instead of bf16
  float acc_scale;        // accumulator -> value: 1, or 2^-15 for the scaled fp16 operand planes (ACT_SCALE * WGT_SCALE)
  int out_format;         // bf16/16-bit outputs: 0 = bf16 hi(/lo), 1 = fp16 hi(/lo) of value * 2^4, 2 = fp16 hi + e4m3 pair plane
  float* out_f32;         // optional outputs, all [batch*rows_per_batch, N] row-major
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
};

// (mean, rstd) of one row from its partial sums: `parts` (sum, sum of squares) pairs, added in index order (deterministic)
// Layout [parts][rows][2]: the 32 rows of a warp are contiguous, so each of the `parts` loads is two full lines per warp.
// NOTE the hot path does not come here: 12 predicated loads + their address arithmetic per thread and tile cost the FFN1 epilogue
// ~5 extra instructions per element (ncu: 42 M vs 30 M warp instructions, 57 % vs 75 % tensor-pipe activity), so the model reduces
// the partial sums to (mean, rstd) once per row with w2v2_row_stats_finalize and the GEMMs read ONE float2 per row.
__device__ __forceinline__ float2 mean_rstd_from_parts(const float* stats, size_t row, size_t rows_total, int parts, float inv_dim,
                                                       float eps) {
  // all (<= 16) loads are issued before the first add: a rolled load -> add loop serialises `parts` L2 latencies per tile
  const float2* p2 = reinterpret_cast<const float2*>(stats) + row;
  float2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (i < parts) ? __ldg(p2 + (size_t)i * rows_total) : make_float2(0.0f, 0.0f);
  float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    s1 += v[i].x;
    s2 += v[i].y;
  }
  const float mean = s1 * inv_dim;
  const float var = fmaxf(fmaf(-mean, mean, s2 * inv_dim), 0.0f);
  return make_float2(mean, rsqrtf(var + eps));
}

// Producer-side finalisation of the row statistics (w2v2.h row_stats_final): called by ALL lanes of a warp that has just written
// its (sum, sum of squares) partial for the 32 rows orow0 .. orow0 + 31 (one 64-column group).  The warp that brings the block's
// arrival counter to `stats_parts` sums the partials in index order - the arithmetic of w2v2_row_stats_finalize, bit for bit - writes
// (mean, rstd) and re-arms the counter.  (threadFenceReduction pattern: partials -> fence -> counter; last arriver: fence -> L2 loads.)
__device__ __forceinline__ void row_stats_finalize_last(const GemmParams& p, size_t orow0, int lane, int rows_valid) {
  __threadfence();
  __syncwarp();
  unsigned prev = 0;
  if (lane == 0) prev = atomicAdd(p.stats_counter + (orow0 >> 5), 1u);
  prev = __shfl_sync(0xffffffffu, prev, 0);
  if (prev + 1u != (unsigned)p.stats_parts) return;
  __threadfence();
  if (lane < rows_valid) {
    const size_t rows_total = (size_t)p.batch * p.rows_per_batch;
    const float2* p2 = reinterpret_cast<const float2*>(p.row_stats_out) + orow0 + lane;
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (i < p.stats_parts) ? __ldcg(p2 + (size_t)i * rows_total) : make_float2(0.0f, 0.0f);
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s1 += v[i].x;
      s2 += v[i].y;
    }
    const float inv_dim = 1.0f / (float)p.N;
    const float mean = s1 * inv_dim;
    const float var = fmaxf(fmaf(-mean, mean, s2 * inv_dim), 0.0f);
    p.stats_final[orow0 + lane] = make_float2(mean, rsqrtf(var + p.ln_eps));
  }
  if (lane == 0) p.stats_counter[orow0 >> 5] = 0u;
}

// Per-row LayerNorm constants of one output row, fetched BEFORE the epilogue waits for its accumulator (their load latency - up
// to 2 x 16 scattered partial sums per thread - then hides behind the MMAs of the tile): x = mean, y = rstd of the residual's
// LayerNorm; z = rstd * acc_scale, w = -rstd * mean of the LayerNorm folded into this GEMM (z = acc_scale, w = 0 without fold).
template <int EPI>
__device__ __forceinline__ float4 epilogue_row_constants(const GemmParams& p, size_t orow, bool row_ok) {
  const bool f_res = (EPI >= 0) ? bool(EPI & EPI_RESID) : (p.residual != nullptr);
  const bool f_fold = (EPI >= 0) ? bool(EPI & EPI_LNFOLD) : (p.ln_fold_stats != nullptr);
  float4 c = make_float4(0.0f, 0.0f, p.acc_scale, 0.0f);
  const size_t rows_total = (size_t)p.batch * p.rows_per_batch;
  if (f_res && p.ln_stats != nullptr && row_ok) {
    const float2 st = (p.res_ln_parts > 0) ? mean_rstd_from_parts(p.ln_stats, orow, rows_total, p.res_ln_parts, 1.0f / (float)p.N, p.ln_eps)
                                           : __ldg(reinterpret_cast<const float2*>(p.ln_stats) + orow);
    c.x = st.x;
    c.y = st.y;
  }
  if (f_fold && row_ok) {
    const float2 st = (p.ln_fold_parts > 0) ? mean_rstd_from_parts(p.ln_fold_stats, orow, rows_total, p.ln_fold_parts, p.ln_fold_inv_dim, p.ln_eps)
                                            : __ldg(reinterpret_cast<const float2*>(p.ln_fold_stats) + orow);
    c.z = st.y * p.acc_scale;
    c.w = -st.y * st.x;
  }
  return c;
}

// residual term "LayerNorm(r)" with the statistics written by w2v2_ln_rows_stats: the SAME expression as ln_rows_kernel, so
// the value is bit-identical to the fp32 LayerNorm output it replaces
__device__ __forceinline__ float4 ln_of_residual(float4 r, float mean, float rstd, const float* gamma, const float* beta, int col) {
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
  const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
  return make_float4(fmaf((r.x - mean) * rstd, g.x, b.x), fmaf((r.y - mean) * rstd, g.y, b.y),
                     fmaf((r.z - mean) * rstd, g.z, b.z), fmaf((r.w - mean) * rstd, g.w, b.w));
}

// Dropout on 16 consecutive elements of one output row starting at flat element index `e0` (a multiple of 4): the same
// stateless stream as w2v2_dropout_rows - 64 bits per group of four consecutive elements (w2v2_common.cuh).
__device__ __forceinline__ void epilogue_dropout16(const DropSpec& d, size_t e0, float (&v)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint64_t bits = drop_bits4(d, (uint64_t)(e0 >> 2) + g);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[4 * g + e] = drop_keep(bits, e, d.thr16) ? v[4 * g + e] * d.scale : 0.0f;
  }
}

// Epilogue of one 128 x BLOCK_N accumulator tile for one warp (32 rows; lane = row).  The two warps that share a
// TMEM lane quadrant take alternate PAIRS of 32-column chunks (grp = 0 / 1).
//  * tcgen05.ld has a long latency while the tensor pipe is busy, so ALL of this warp's chunks are requested
//    before a single wait (a serial ld/wait per chunk costs ~5.7k cycles per tile and caps K = 768 GEMMs at ~55 %
//    of peak).  Once the slice sits in registers the accumulator stage is handed back to the MMA warp at once.
//  * A thread owns one ROW, so direct stores would scatter 16-byte pieces over 32 different lines per
//    instruction (measured: ~35 % of the kernel).  Instead every 32 x 128-byte block is transposed through a
//    per-warp, XOR-swizzled smem buffer and written with 8 lanes per row: 4 full 128-byte lines per instruction.
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, uint32_t taddr, int grp, int n0, size_t orow,
                                                   int rows_valid, bool zero_row, const float* sb, uint8_t* stage,
                                                   uint32_t tmem_empty_cluster_addr, float4 rowc) {
  // compile-time epilogue recipe (EPI >= 0) or run-time flags (EPI < 0)
  const bool f_gelu = (EPI >= 0) ? bool(EPI & EPI_GELU) : (p.gelu != 0);
  const bool f_fast = (EPI >= 0) ? bool(EPI & EPI_FASTGELU) : (p.gelu == 2);  // tanh-form GELU: single-pass mode
  const bool f_res = (EPI >= 0) ? bool(EPI & EPI_RESID) : (p.residual != nullptr);
  const bool f_f32 = (EPI >= 0) ? bool(EPI & EPI_F32) : (p.out_f32 != nullptr);
  const bool f_hi = (EPI >= 0) ? bool(EPI & EPI_HI) : (p.out_hi != nullptr);
  const bool f_lo = (EPI >= 0) ? bool(EPI & EPI_LO) : (p.out_lo != nullptr);
  const bool f_scale = (EPI >= 0) ? bool(EPI & EPI_SCALE) : (p.scale != nullptr);
  const bool f_fold = (EPI >= 0) ? bool(EPI & EPI_LNFOLD) : (p.ln_fold_stats != nullptr);
  constexpr int NCH = BLOCK_N / 32;
  constexpr int NMINE = (NCH >= 4) ? NCH / 2 : NCH;  // chunks per warp (grp 1 idles when NCH < 4)
  const int lane = lane_id();
  // chunk index of my i-th chunk: pairs (0,1),(4,5),.. for grp 0 and (2,3),(6,7),.. for grp 1
  auto chunk_of = [&](int i) { return 4 * (i >> 1) + 2 * grp + (i & 1); };
  uint32_t r[NMINE][32];
  if (!W2V2_DBG(p, 3)) {
#pragma unroll
    for (int i = 0; i < NMINE; ++i)
      if (chunk_of(i) < NCH) tmem_ld_32x32b_x32(taddr + chunk_of(i) * 32, r[i]);
    tmem_ld_wait();
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster(tmem_empty_cluster_addr);  // accumulator stage is free again
  if (W2V2_DBG(p, 3)) return;
  if (W2V2_DBG(p, 1)) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < NMINE; ++i) x ^= r[i][i];
    if (__uint_as_float(x) == 1.2345e-30f) p.out_f32[0] = 0.0f;
    return;
  }
  if (rows_valid <= 0) return;
  const bool row_ok = lane < rows_valid;
  const size_t orow0 = orow - lane;  // first row of this warp's block
  const bool f_ln = f_res && p.ln_stats != nullptr;
  const float ln_mean = rowc.x, ln_rstd = rowc.y;
  const float fold_rs = rowc.z, fold_nm = rowc.w;   // LayerNorm fold: v = acc * (rstd * acc_scale) + (-rstd * mean) * colsum + bias

  // staged 32 x 128 B block -> global, 8 lanes per row (PIECES = 8) or 4 lanes per row (PIECES = 4: 64-byte rows)
  auto flush = [&](uint8_t* gbase, size_t row_stride_bytes, int pieces, bool atomic = false) {
    __syncwarp();
    if (pieces == 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3), pc = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((pc ^ (row & 7)) << 4));
        if (row < rows_valid) {
          if (atomic) {
            float* g = reinterpret_cast<float*>(gbase + row * row_stride_bytes + pc * 16);
            // one 16-byte vector reduction instead of four scalar atomics (the split-K wgrad tiles issue 2.4 M adds per launch)
            asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g), "f"(__uint_as_float(val.x)),
                         "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                         : "memory");
          } else {
            *reinterpret_cast<uint4*>(gbase + row * row_stride_bytes + pc * 16) = val;
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int row = 8 * k + (lane >> 2), pc = lane & 3;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + row * 128 + ((pc ^ (row & 7)) << 4));
        if (row < rows_valid) *reinterpret_cast<uint4*>(gbase + row * row_stride_bytes + pc * 16) = val;
      }
    }
    __syncwarp();
  };
  auto put = [&](int piece, uint4 val) {
    *reinterpret_cast<uint4*>(stage + lane * 128 + ((piece ^ (lane & 7)) << 4)) = val;
  };

  // ---- pass A: bias / GELU / residual / mask, in place in r[][] (as fp32 bit patterns)
#pragma unroll
  for (int i = 0; i < NMINE; ++i) {
    const int ch = chunk_of(i);
    const int c0 = ch * 32;
    const int n = n0 + c0;
    if (ch >= NCH || n >= p.N) continue;
    const bool full_chunk = (n + 32 <= p.N) && p.vec_ok;
    if (f_res && full_chunk) {
      // residual block (32 rows x 128 B): coalesced loads, 8 lanes per row, transposed through the staging buffer
      const uint8_t* gres = reinterpret_cast<const uint8_t*>(p.residual + orow0 * p.N + n);
      uint4 t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3), pc = lane & 7;
        t[k] = (row < rows_valid) ? __ldg(reinterpret_cast<const uint4*>(gres + (size_t)row * p.N * 4 + pc * 16))
                                  : make_uint4(0u, 0u, 0u, 0u);
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = 4 * k + (lane >> 3), pc = lane & 7;
        *reinterpret_cast<uint4*>(stage + row * 128 + ((pc ^ (row & 7)) << 4)) = t[k];
      }
      __syncwarp();
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float rf[16];
      if (f_res) {
        if (full_chunk) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 q = *reinterpret_cast<const float4*>(stage + lane * 128 + (((4 * hf + j) ^ (lane & 7)) << 4));
            rf[4 * j + 0] = q.x;
            rf[4 * j + 1] = q.y;
            rf[4 * j + 2] = q.z;
            rf[4 * j + 3] = q.w;
          }
        } else {
          const size_t off = orow * p.N + n + 16 * hf;
#pragma unroll
          for (int j = 0; j < 16; ++j) rf[j] = (row_ok && n + 16 * hf + j < p.N) ? __ldg(p.residual + off + j) : 0.0f;
        }
        if (f_ln) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = n + 16 * hf + j;
            if (col < p.N) rf[j] = fmaf((rf[j] - ln_mean) * ln_rstd, __ldg(p.ln_gamma + col), __ldg(p.ln_beta + col));
          }
        }
      }
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = *reinterpret_cast<const float4*>(sb + c0 + 16 * hf + 4 * j);
        if (f_fold) {
          const float4 cs = *reinterpret_cast<const float4*>(sb + 2 * BLOCK_N + c0 + 16 * hf + 4 * j);
          v[4 * j + 0] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 0]), fold_rs, fmaf(fold_nm, cs.x, bb.x));
          v[4 * j + 1] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 1]), fold_rs, fmaf(fold_nm, cs.y, bb.y));
          v[4 * j + 2] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 2]), fold_rs, fmaf(fold_nm, cs.z, bb.z));
          v[4 * j + 3] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 3]), fold_rs, fmaf(fold_nm, cs.w, bb.w));
        } else if (f_scale) {
          const float4 sc = *reinterpret_cast<const float4*>(sb + 2 * BLOCK_N + c0 + 16 * hf + 4 * j);
          v[4 * j + 0] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 0]), sc.x, bb.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 1]), sc.y, bb.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 2]), sc.z, bb.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 3]), sc.w, bb.w);
        } else {
          // fmaf(acc, 1, b) == acc + b bit for bit; the fp16 modes un-scale their 2^15 accumulators here for free
          v[4 * j + 0] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 0]), p.acc_scale, bb.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 1]), p.acc_scale, bb.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 2]), p.acc_scale, bb.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[i][16 * hf + 4 * j + 3]), p.acc_scale, bb.w);
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
      if (EPI < 0 && p.drop.thr16) epilogue_dropout16(p.drop, orow * p.N + n + 16 * hf, v);
      if (f_res) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += rf[j];
      }
      if (EPI < 0 && p.row_replace != nullptr && row_ok && p.row_replace[orow]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (n + 16 * hf + j < p.N) ? __ldg(p.row_value + n + 16 * hf + j) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) r[i][16 * hf + j] = zero_row ? 0u : __float_as_uint(v[j]);
    }
  }
  if (W2V2_DBG(p, 2)) {
    if (__uint_as_float(r[0][0] ^ r[NMINE - 1][31]) == 1.2345e-30f) p.out_f32[0] = 0.0f;
    return;
  }
  if (EPI < 0 && p.row_stats_out != nullptr && row_ok) {
    // (sum, sum of squares) of this row's output per 64-column group = per chunk pair (N % 64 == 0: checked by the launcher)
#pragma unroll
    for (int i = 0; i + 1 < NMINE; i += 2) {
      const int n = n0 + chunk_of(i) * 32;
      if (chunk_of(i) >= NCH || n >= p.N) continue;
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[i + u][j]);
          s1 += x;
          s2 = fmaf(x, x, s2);
        }
      }
      reinterpret_cast<float2*>(p.row_stats_out)[(size_t)(n >> 6) * ((size_t)p.batch * p.rows_per_batch) + orow] = make_float2(s1, s2);
    }
  }

  // ---- pass B: stores
#pragma unroll
  for (int i = 0; i < NMINE; ++i) {
    const int ch = chunk_of(i);
    const int n = n0 + ch * 32;
    if (ch >= NCH || n >= p.N) continue;
    const bool full_chunk = (n + 32 <= p.N) && p.vec_ok;
    if (!full_chunk) {
      // ragged right edge (N not a multiple of 32, or unaligned N): scalar row-per-thread path
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (n + j >= p.N) continue;
          const float v = __uint_as_float(r[i][j]);
          if (f_f32) {
            if (p.atomic_f32) atomicAdd(p.out_f32 + orow * p.N + n + j, v);
            else p.out_f32[orow * p.N + n + j] = v;
          }
          if (f_hi && p.out_format == 0) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            p.out_hi[orow * p.N + n + j] = h;
            if (f_lo) p.out_lo[orow * p.N + n + j] = __float2bfloat16_rn(v - __bfloat162float(h));
          } else if (f_hi) {   // fp16 planes (format 2 needs N % 64 == 0 and never gets here)
            uint32_t l;
            const uint32_t h = split_f16x2(v * ACT_SCALE, 0.0f, l);
            reinterpret_cast<uint16_t*>(p.out_hi)[orow * p.N + n + j] = (uint16_t)(h & 0xFFFFu);
            if (f_lo) reinterpret_cast<uint16_t*>(p.out_lo)[orow * p.N + n + j] = (uint16_t)(l & 0xFFFFu);
          }
        }
      }
      continue;
    }
    if (f_f32) {  // 32 fp32 columns = one 128-byte row segment
#pragma unroll
      for (int j = 0; j < 8; ++j) put(j, make_uint4(r[i][4 * j], r[i][4 * j + 1], r[i][4 * j + 2], r[i][4 * j + 3]));
      flush(reinterpret_cast<uint8_t*>(p.out_f32 + orow0 * p.N + n), (size_t)p.N * 4, 8, p.atomic_f32 != 0);
    }
    if (f_hi) {
      // bf16: an even/odd chunk pair forms one 128-byte row segment; a lone chunk is a 64-byte segment
      const bool pair_lo = (i & 1) == 0 && (i + 1 < NMINE) && (chunk_of(i + 1) < NCH) && (n + 64 <= p.N);
      const bool pair_hi = (i & 1) == 1 && (n0 + chunk_of(i - 1) * 32 + 64 <= p.N);
      if (pair_hi) continue;  // already written together with chunk i-1
      const int fmt = p.out_format;
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        if (plane == 1 && !f_lo) continue;
        __nv_bfloat16* dst = plane == 0 ? p.out_hi : p.out_lo;
        const int nch = pair_lo ? 2 : 1;
        if (fmt == 2 && plane == 1) {
          // e4m3 pair plane [rows][2 N] bytes: per 64-column group 64 bytes of e4m3((v - hi) 2^6) then 64 bytes of e4m3(hi 2^-6)
          // (N % 64 == 0 is checked by the launcher, so chunks always come in pairs here)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int iu = (i + u < NMINE) ? i + u : i;
#pragma unroll
            for (int j = 0; j < 2; ++j) {        // 16 columns per 16-byte piece
              uint16_t l8[8], h8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                split_f16_f8x2(__uint_as_float(r[iu][16 * j + 2 * e]) * ACT_SCALE, __uint_as_float(r[iu][16 * j + 2 * e + 1]) * ACT_SCALE,
                               l8[e], h8[e]);
              put(2 * u + j, make_uint4(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16), l8[4] | (l8[5] << 16), l8[6] | (l8[7] << 16)));
              put(4 + 2 * u + j, make_uint4(h8[0] | (h8[1] << 16), h8[2] | (h8[3] << 16), h8[4] | (h8[5] << 16), h8[6] | (h8[7] << 16)));
            }
          }
          flush(reinterpret_cast<uint8_t*>(p.out_lo) + (orow0 * p.N + n) * 2, (size_t)p.N * 2, 8);
          continue;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (u >= nch) continue;
          const int iu = (i + u < NMINE) ? i + u : i;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float v0 = __uint_as_float(r[iu][8 * j + 2 * e]), v1 = __uint_as_float(r[iu][8 * j + 2 * e + 1]);
              h[e] = (fmt == 0) ? split_bf16x2(v0, v1, l[e]) : split_f16x2(v0 * ACT_SCALE, v1 * ACT_SCALE, l[e]);
            }
            put(4 * u + j, plane == 0 ? make_uint4(h[0], h[1], h[2], h[3]) : make_uint4(l[0], l[1], l[2], l[3]));
          }
        }
        flush(reinterpret_cast<uint8_t*>(dst + orow0 * p.N + n), (size_t)p.N * 2, pair_lo ? 8 : 4);
      }
    }
  }
}

// Per-tile epilogue preamble, issued BEFORE the accumulator is ready so its latency hides behind the MMAs:
// the bias slice goes to smem (one element per epilogue thread), this thread's residual lines are pulled into L2.
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_prepare(const GemmParams& p, int et, int grp, int n0, size_t orow,
                                                      bool row_ok, float* sb, int b) {
  const bool f_scale = (EPI >= 0) ? bool(EPI & EPI_SCALE) : (p.scale != nullptr);
  const size_t boff = (size_t)b * p.bias_bstride + n0 + et;
  if (et < BLOCK_N) {
    sb[et] = (p.bias != nullptr && n0 + et < p.N) ? __ldg(p.bias + boff) : 0.0f;
    if (f_scale) sb[2 * BLOCK_N + et] = (n0 + et < p.N) ? __ldg(p.scale + boff) : 1.0f;  // scale (or LayerNorm-fold colsum) slices follow the bias slices
  }
  const bool f_res = (EPI >= 0) ? bool(EPI & EPI_RESID) : (p.residual != nullptr);
  if (f_res && row_ok) {
#pragma unroll
    for (int c = grp; c < BLOCK_N / 32; c += 2)
      if (n0 + c * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + orow * p.N + n0 + c * 32));
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");  // bias slice visible to all 8 epilogue warps
}

GemmParams make_gemm_params(const w2v2_gemm_args* a, int block_n);
int launch_gemm_2sm(const w2v2_gemm_args* a, cudaStream_t stream);  // gemm_2sm.cu

}  // namespace w2v2
