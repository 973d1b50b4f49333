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

struct Gemm2Smem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;           // 16 KB
  static constexpr int B_BYTES = (GEMM2_BLOCK_N / 2) * GEMM_BLOCK_K * 2;     // 16 KB: this CTA's half of the n-tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = GEMM2_STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = 2 * GEMM2_BLOCK_N * 4;
  static constexpr int EPI_OFF = RING_BYTES + BAR_BYTES + BIAS_BYTES;
  static constexpr int TOTAL = EPI_OFF + GEMM_EPI_STAGE_BYTES + 1024;
};

template <int PASSES, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                     const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                     const GemmParams p) {
  using S = Gemm2Smem;
  constexpr int BLOCK_N = GEMM2_BLOCK_N;
  constexpr int ACC_STAGES = 2;
  constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;  // 512

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::RING_BYTES);
  uint64_t* empty_bar = full_bar + GEMM2_STAGES;
  uint64_t* tmem_full = empty_bar + GEMM2_STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  float* s_bias = reinterpret_cast<float*>(smem + S::RING_BYTES + S::BAR_BYTES);

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
    if (PASSES == 3) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < GEMM2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);   // leader's arrive.expect_tx covers both CTAs' bytes (used in the leader only)
      mbar_init(&empty_bar[i], 1);  // leader's multicast commit
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);    // leader's multicast commit
      mbar_init(&tmem_empty[i], 16);  // 8 epilogue warps x 2 CTAs (used in the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {
  // warpgroup 0 (TMA / MMA / TMEM-alloc warps) gives registers away ...
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CONTROL));
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
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
          const CUtensorMap* ma = (pass == 1) ? &tmA_lo : &tmA_hi;
          const CUtensorMap* mb = (pass == 2) ? &tmB_lo : &tmB_hi;
          int kc = kb * GEMM_BLOCK_K, trow = t0;
          if (kb >= p.kb_split) {
            kc = (kb - p.kb_split) * GEMM_BLOCK_K;
            trow = t0 + 1;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          const uint32_t lead_full = mapa_cluster(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::STAGE_BYTES);  // both CTAs' TMA bytes land here
          tma_load_3d_2sm(sa, ma, lead_full, kc, trow, b);
          tma_load_2d_2sm(sb, mb, lead_full, kb * GEMM_BLOCK_K, n0);
          if (++stage == GEMM2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = idesc_bf16(2 * GEMM_BLOCK_M, BLOCK_N, 0, 0);
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
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) umma_f16_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
          umma_commit_2sm_mcast(&empty_bar[stage], 3);  // slot free in both CTAs once these MMAs retire
          if (++stage == GEMM2_STAGES) {
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
  // ... to the epilogue warpgroups, which keep their whole accumulator slice (128 registers) in flight
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_EPILOGUE));
  {
    // ------------------------------------------------------------------ epilogue (8 warps per CTA)
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int lane = lane_id();
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = pair_id; w < total_work; w += num_pairs) {
      const int n_tile = w % p.n_tiles;
      const int m_raw = (w / p.n_tiles) * 2 + crank;
      const int m_tile = min(m_raw, total_m_tiles - 1);
      const int b = m_tile / p.tiles_per_batch;
      const int t = (m_tile - b * p.tiles_per_batch) * GEMM_BLOCK_M + ew * 32 + lane;
      const int n0 = n_tile * BLOCK_N;
      const bool row_ok = t < p.rows_per_batch && m_raw < total_m_tiles;
      const int rows_valid = (m_raw < total_m_tiles) ? min(32, p.rows_per_batch - (t - lane)) : 0;
      uint8_t* stage = smem + S::EPI_OFF + (warp - 4) * 4096;
      const bool zero_row = p.row_valid != nullptr && t >= p.row_valid[b];
      const size_t orow = (size_t)b * p.rows_per_batch + t;
      float* sb = s_bias + acc * BLOCK_N;
      gemm_epilogue_prepare<BLOCK_N, EPI>(p, et, grp, n0, orow, row_ok, sb);

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
      gemm_epilogue_tile<BLOCK_N, EPI>(p, taddr, grp, n0, orow, rows_valid, zero_row, sb, stage,
                                  mapa_cluster(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
}

template <int PASSES, int EPI>
static int launch_gemm_2sm_t(const w2v2_gemm_args* a, cudaStream_t stream) {
  using S = Gemm2Smem;
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
  GemmParams p = make_gemm_params(a, GEMM2_BLOCK_N);
  auto kern = gemm_bf16_2sm_kernel<PASSES, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    W2V2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  int dev = 0, sms = 0;
  W2V2_CUDA(cudaGetDevice(&dev));
  W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total_m_tiles = p.batch * p.tiles_per_batch;
  const int total_work = ((total_m_tiles + 1) / 2) * p.n_tiles;
  int grid = total_work * 2 < sms ? total_work * 2 : sms;
  if (a->max_ctas > 0 && grid > a->max_ctas) grid = a->max_ctas;
  grid -= grid % 2;
  if (grid < 2) grid = 2;
  kern<<<grid, GEMM_THREADS, S::TOTAL, stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
  W2V2_CUDA(cudaGetLastError());
  return 0;
}

// Picks the compile-time epilogue recipe matching the call (the combinations the Wav2Vec2 forward uses);
// anything else runs the generic run-time-flag instance.
template <int PASSES>
static int dispatch_2sm(const w2v2_gemm_args* a, cudaStream_t s) {
  const bool gelu = (a->flags & W2V2_GEMM_GELU) != 0, res = a->residual != nullptr;
  const bool f32 = a->out_f32 != nullptr, hi = a->out_hi != nullptr, lo = a->out_lo != nullptr;
  constexpr int LO = (PASSES == 3) ? EPI_LO : 0;     // the model writes hi+lo planes exactly in 3-pass mode
  if (lo == (PASSES == 3) || !hi) {
    if (gelu && !res && !f32 && hi) return launch_gemm_2sm_t<PASSES, EPI_GELU | EPI_HI | LO>(a, s);
    if (gelu && !res && f32 && !hi) return launch_gemm_2sm_t<PASSES, EPI_GELU | EPI_F32>(a, s);
    if (!gelu && !res && f32 && !hi) return launch_gemm_2sm_t<PASSES, EPI_F32>(a, s);
    if (!gelu && !res && f32 && hi) return launch_gemm_2sm_t<PASSES, EPI_F32 | EPI_HI | LO>(a, s);
    if (!gelu && !res && !f32 && hi) return launch_gemm_2sm_t<PASSES, EPI_HI | LO>(a, s);
    if (!gelu && res && f32 && !hi) return launch_gemm_2sm_t<PASSES, EPI_RESID | EPI_F32>(a, s);
  }
  return launch_gemm_2sm_t<PASSES, EPI_RUNTIME>(a, s);
}

int launch_gemm_2sm(const w2v2_gemm_args* a, cudaStream_t stream) {
  return a->passes == 1 ? dispatch_2sm<1>(a, stream) : dispatch_2sm<3>(a, stream);
}

}  // namespace w2v2
