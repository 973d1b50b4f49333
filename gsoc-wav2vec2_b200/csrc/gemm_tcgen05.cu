// Persistent warp-specialised bf16 GEMM for sm_100a:  D[rows, N] = A[rows, K] * W[N, K]^T  (+ epilogue)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected lane, tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16)
//   warp 2      : TMEM allocator (2 accumulator stages of BLOCK_N fp32 columns)
//   warps 4..11 : epilogue       (tcgen05.ld 32x32b -> bias / erf-GELU / residual / mask -> fp32 and/or bf16 hi,lo)
//
// The same kernel serves every dense contraction on the Wav2Vec2 path:
//   * encoder Dense layers (reference: encoder.py:24-31,127-128; feature_extractor.py:94; modeling.py:254)
//   * the strided Conv1D layers 1..6 as implicit GEMMs (reference: feature_extractor.py:55): in channels-last
//     layout a k-tap, stride-s window is k*Cin CONTIGUOUS elements whose start advances by s*Cin per frame, so
//     the A operand is a 3-D TMA tensor map {k*Cin, T_out, B} with an overlapping row stride - no im2col.
// "Ragged" rows: each batch entry has rows_per_batch valid rows; tiles never straddle two entries, TMA
// zero-fills rows past the entry and the epilogue masks the stores.
//
// PASSES = 3 is the parity mode: operands arrive as bf16 hi/lo planes and the kernel accumulates
// A_hi*W_hi + A_lo*W_hi + A_hi*W_lo into one fp32 TMEM accumulator (~16 mantissa bits per operand).
#include "gemm_common.cuh"

namespace w2v2 {

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // narrow tiles serve the small problems, where a CTA's serial k-loop is bound by TMA latency / ring depth: as deep as smem allows
  static constexpr int STAGES = (BLOCK_N == 256) ? 3 : (BLOCK_N == 128) ? 5 : 7;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = 4 * BLOCK_N * 4;  // per accumulator stage: bias slice; scale slices follow
  static constexpr int EPI_OFF = RING_BYTES + BAR_BYTES + BIAS_BYTES;
  static constexpr int TOTAL = EPI_OFF + GEMM_EPI_STAGE_BYTES + 1024;  // + slack for 1024-byte alignment
};

// CLUSTER = 2: two CTAs on neighbouring SMs take two consecutive m-tiles of the SAME n-tile and share the
// weight tile: each CTA fetches half of it and TMA-multicasts that half into both CTAs' smem, which cuts the
// L2 -> SM traffic per 128x256x64 block from 48 KB to 32 KB (the kernel is L2-bandwidth bound otherwise).
template <int BLOCK_N, int PASSES, int CLUSTER>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                         const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                         const GemmParams p) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int ACC_STAGES = 2;
  constexpr int TMEM_COLS = (ACC_STAGES * BLOCK_N) < 32 ? 32 : (ACC_STAGES * BLOCK_N);
  static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::RING_BYTES);
  uint64_t* empty_bar = full_bar + S::STAGES;
  uint64_t* tmem_full = empty_bar + S::STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  float* s_bias = reinterpret_cast<float*>(smem + S::RING_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int total_kb = PASSES * p.num_kb;
  // work item = (group of CLUSTER consecutive m-tiles, n-tile); n fastest so the A rows stay hot in L2
  const int crank = (CLUSTER > 1) ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CLUSTER;
  const int num_clusters = gridDim.x / CLUSTER;
  const int total_m_tiles = p.batch * p.tiles_per_batch;
  const int total_work = ((total_m_tiles + CLUSTER - 1) / CLUSTER) * p.n_tiles;
  constexpr uint16_t kMask = (uint16_t)((1u << CLUSTER) - 1);
  constexpr int B_SLICE_ROWS = BLOCK_N / CLUSTER;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmB_hi);
    if (PASSES != 1) {
      tma_prefetch_desc(&tmA_lo);
      tma_prefetch_desc(&tmB_lo);
    }
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < S::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], CLUSTER);  // every CTA of the cluster must have released the slot
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  if (CLUSTER > 1) cluster_sync_all(); else __syncthreads();  // peers' barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // prologue above overlapped the previous kernel; from here on its results are visible
  if (warp < 4) {
  // warpgroup 0 (TMA / MMA / TMEM-alloc warps) gives registers away ...
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GEMM_REGS_CONTROL));
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const int n_tile = w % p.n_tiles;
        const int m_tile = min((w / p.n_tiles) * CLUSTER + crank, total_m_tiles - 1);  // clamp: idle half still feeds B
        const int b = m_tile / p.tiles_per_batch;
        const int t0 = (m_tile - b * p.tiles_per_batch) * GEMM_BLOCK_M;
        const int n0 = n_tile * BLOCK_N;
        for (int it = 0; it < total_kb; ++it) {
          const int pass = (PASSES == 1) ? 0 : it / p.num_kb;
          const int kb = it - pass * p.num_kb;
          // PASSES == 2 (fp16f8): pass 1 reads the e4m3 pair planes of both operands (128-byte k-blocks), see gemm_2sm.cu
          const bool f8pass = (PASSES == 2) && pass == 1;
          const CUtensorMap* ma = (pass == 1) ? &tmA_lo : &tmA_hi;
          const CUtensorMap* mb = (pass == 2 || f8pass) ? &tmB_lo : &tmB_hi;
          const int kw = f8pass ? 2 * GEMM_BLOCK_K : GEMM_BLOCK_K;
          int kc = kb * kw, trow = t0;
          if (kb >= p.kb_split) {
            kc = (kb - p.kb_split) * kw;
            trow = t0 + 1;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::STAGE_BYTES;
          uint8_t* sb = sa + S::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], S::STAGE_BYTES);
          if (CLUSTER == 1 && p.mn_major) {
            // D = X^T Y over the row index: X [rows][m], Y [rows][n] row-major.  A k-block is 64 ROWS; each operand tile is a
            // set of [64 rows x 64 columns] boxes = 128-byte-swizzled MN-major atoms of 8 KB, 64 columns apart.  The batch index is
            // the split-K slice: slice b reduces rows [b K, (b + 1) K) and all slices add into the same output.
#pragma unroll
            for (int at = 0; at < GEMM_BLOCK_M / 64; ++at) tma_load_2d(sa + at * 8192, ma, &full_bar[stage], t0 + 64 * at, (b * p.num_kb + kb) * GEMM_BLOCK_K);
#pragma unroll
            for (int at = 0; at < BLOCK_N / 64; ++at) tma_load_2d(sb + at * 8192, mb, &full_bar[stage], n0 + 64 * at, (b * p.num_kb + kb) * GEMM_BLOCK_K);
          } else {
          tma_load_3d(sa, ma, &full_bar[stage], kc, trow, b);
          if (CLUSTER == 1)
            tma_load_2d(sb, mb, &full_bar[stage], kb * kw, n0);
          else  // my slice of the weight tile, delivered to every CTA of the cluster
            tma_load_2d_mcast(sb + crank * (B_SLICE_ROWS * 128), mb, &full_bar[stage], kb * kw,
                              n0 + crank * B_SLICE_ROWS, kMask);
          }
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const bool mn = (CLUSTER == 1) && p.mn_major;
      const uint32_t idesc = idesc_16bit(p.fp16 != 0, GEMM_BLOCK_M, BLOCK_N, mn ? 1 : 0, mn ? 1 : 0);
      constexpr uint32_t idesc8 = idesc_fmt0(GEMM_BLOCK_M, BLOCK_N, 0, 0);     // e4m3 x e4m3 under kind::f8f6f4
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int it = 0; it < total_kb; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t da = desc_kmajor_sw128(sa);
          const uint64_t db = desc_kmajor_sw128(sa + S::A_BYTES);
          if (mn) {
            // MN-major atoms: 64 columns per 128-byte row, 8 rows per 1 KB group (SBO), atoms 8 KB apart (LBO); one UMMA_K
            // step = 16 rows = 2 KB further
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              umma_f16(d_tmem, desc_mnmajor_sw128(sa + k * 2048, 8192, 1024), desc_mnmajor_sw128(sa + S::A_BYTES + k * 2048, 8192, 1024),
                       idesc, (it | k) != 0);
          } else if (PASSES == 2 && it >= p.num_kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f8(d_tmem, da + 2 * k, db + 2 * k, idesc8, 1);   // K = 32 e4m3 = 32 bytes per step
          } else {
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // +32 bytes per UMMA_K step inside the swizzle row -> +2 in the (addr >> 4) field
            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0);
          }
          }
          if (CLUSTER == 1) umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          else umma_commit_mcast(&empty_bar[stage], kMask);   // ... in every CTA that multicasts into it
          if (++stage == S::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
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
    // ------------------------------------------------------------------ epilogue (8 warps)
    // warp w reads TMEM lane quadrant (w % 4); the two warps sharing a quadrant take alternate 32-column
    // chunks.  Per tile: bias slice -> smem (issued before the accumulator is ready), residual lines
    // prefetched to L2, then a software pipeline: tcgen05.ld of chunk c+1 is in flight while chunk c is
    // processed and stored.
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int lane = lane_id();
    const int et = threadIdx.x - 128;  // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      const int n_tile = w % p.n_tiles;
      const int m_raw = (w / p.n_tiles) * CLUSTER + crank;
      const int m_tile = min(m_raw, total_m_tiles - 1);
      const int b = m_tile / p.tiles_per_batch;
      const int t = (m_tile - b * p.tiles_per_batch) * GEMM_BLOCK_M + ew * 32 + lane;
      const int n0 = n_tile * BLOCK_N;
      const bool row_ok = t < p.rows_per_batch && m_raw < total_m_tiles;
      const int rows_valid = (m_raw < total_m_tiles) ? min(32, p.rows_per_batch - (t - lane)) : 0;
      uint8_t* stage = smem + S::EPI_OFF + (warp - 4) * 4096;
      const bool zero_row = p.row_valid != nullptr && t >= p.row_valid[b];
      const size_t orow = (p.mn_major ? (size_t)0 : (size_t)b * p.rows_per_batch) + t;
      float* sb = s_bias + acc * BLOCK_N;
      gemm_epilogue_prepare<BLOCK_N, EPI_RUNTIME>(p, et, grp, n0, orow, row_ok, sb, b);

      const float4 rowc = epilogue_row_constants<EPI_RUNTIME>(p, orow, row_ok);   // in flight while the MMAs of this tile run

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;

      gemm_epilogue_tile<BLOCK_N, EPI_RUNTIME>(p, taddr, grp, n0, orow, rows_valid, zero_row, sb, stage, smem_u32(&tmem_empty[acc]), rowc);
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  }

  tc_fence_before();
  if (CLUSTER > 1) cluster_sync_all(); else __syncthreads();  // no CTA exits while a peer may still signal it
  tc_fence_after();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------- host
thread_local char g_last_error[512] = "";

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("W2V2_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, CUtensorMapSwizzle swz, CUtensorMapDataType dt) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return fail(-2, "%s: cuTensorMapEncodeTiled entry point unavailable", __func__);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "%s: cuTensorMapEncodeTiled failed (CUresult %ld, rank %ld)", __func__, (long)r, rank);
  return 0;
}

GemmParams make_gemm_params(const w2v2_gemm_args* a, int block_n) {
  GemmParams p;
  p.num_kb = a->K / GEMM_BLOCK_K;
  p.kb_split = a->kb_split > 0 ? a->kb_split : p.num_kb;
  p.rows_per_batch = a->rows_per_batch;
  p.tiles_per_batch = (a->rows_per_batch + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  p.batch = a->batch;
  p.n_tiles = (a->N + block_n - 1) / block_n;
  p.N = a->N;
  p.gelu = (a->flags & W2V2_GEMM_GELU_TANH) ? 3 : (a->flags & W2V2_GEMM_GELU) ? (a->passes == 1 ? 2 : 1) : 0;   // 2 = tanh-form fit (bf16 single-pass mode only)
  p.fp16 = mode_fp16(a->passes) ? 1 : 0;
  p.acc_scale = mode_fp16(a->passes) ? ACC_UNSCALE : 1.0f;
  p.out_format = a->out_format;
  p.ln_fold_stats = a->ln_fold_stats;
  p.ln_fold_parts = a->ln_fold_parts;
  p.ln_fold_inv_dim = 1.0f / (float)a->K;
  p.ln_eps = a->ln_eps;
  p.row_stats_out = a->row_stats_out;
  p.stats_final = reinterpret_cast<float2*>(a->row_stats_final);
  p.stats_counter = a->row_stats_counter;
  p.stats_parts = a->N / 64;
  p.res_ln_parts = a->res_ln_parts;
  p.vec_ok = (a->N % 8 == 0) ? 1 : 0;
  p.debug = (int)(a->flags >> 8) & 3;
  p.mn_major = (a->flags & W2V2_GEMM_MN_MAJOR) ? 1 : 0;
  p.atomic_f32 = (p.mn_major && a->batch > 1) ? 1 : 0;
  p.bias = a->bias;
  p.scale = a->scale;
  p.bias_bstride = a->bias_batch_stride;
  p.residual = a->residual;
  p.ln_stats = a->res_ln_stats;
  p.ln_gamma = a->res_ln_gamma;
  p.ln_beta = a->res_ln_beta;
  p.row_valid = a->row_valid;
  p.row_replace = a->row_replace_mask;
  p.row_value = a->row_replace_value;
  p.drop = make_drop(a->drop_p, a->drop_seed, a->drop_site);
  p.out_f32 = a->out_f32;
  p.out_hi = reinterpret_cast<__nv_bfloat16*>(a->out_hi);
  p.out_lo = reinterpret_cast<__nv_bfloat16*>(a->out_lo);

  return p;
}

template <int BLOCK_N, int PASSES, int CLUSTER>
static int launch_gemm(const w2v2_gemm_args* a, cudaStream_t stream) {
  using S = GemmSmem<BLOCK_N>;
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  int rc;
  if (a->flags & W2V2_GEMM_MN_MAJOR) {
    // X [a_rows][rows_per_batch] and Y [a_rows][N] row-major, reduced over their a_rows rows: boxes of 64 rows x 64 columns
    const uint64_t xd[2] = {(uint64_t)a->rows_per_batch, (uint64_t)a->a_rows}, yd[2] = {(uint64_t)a->N, (uint64_t)a->a_rows};
    const uint64_t xs[1] = {(uint64_t)a->a_row_stride * 2}, ys[1] = {(uint64_t)a->w_row_stride * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap(&tmA_hi, a->a_hi, 2, xd, xs, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap(&tmB_hi, a->w_hi, 2, yd, ys, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    tmA_lo = tmA_hi;
    tmB_lo = tmB_hi;
  } else {
  const uint64_t a_dims[3] = {(uint64_t)a->a_row_len, (uint64_t)a->a_rows, (uint64_t)a->batch};
  const uint64_t a_strides[2] = {(uint64_t)a->a_row_stride * 2, (uint64_t)a->a_batch_stride * 2};
  const uint32_t a_box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
  rc = make_tmap(&tmA_hi, a->a_hi, 3, a_dims, a_strides, a_box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tmA_lo = tmA_hi;
  if (PASSES == 3) {
    rc = make_tmap(&tmA_lo, a->a_lo, 3, a_dims, a_strides, a_box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const uint64_t b_dims[2] = {(uint64_t)a->K, (uint64_t)a->w_rows};
  const uint64_t b_strides[1] = {(uint64_t)a->K * 2};
  const uint32_t b_box[2] = {GEMM_BLOCK_K, (uint32_t)(BLOCK_N / CLUSTER)};
  rc = make_tmap(&tmB_hi, a->w_hi, 2, b_dims, b_strides, b_box, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  tmB_lo = tmB_hi;
  if (PASSES == 3) {
    rc = make_tmap(&tmB_lo, a->w_lo, 2, b_dims, b_strides, b_box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  if (PASSES == 2) {   // e4m3 pair planes: same byte strides as the 16-bit planes, 128-byte k-blocks
    const uint64_t a8_dims[3] = {(uint64_t)a->a_row_len * 2, (uint64_t)a->a_rows, (uint64_t)a->batch};
    const uint32_t a8_box[3] = {2 * GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
    if ((rc = make_tmap(&tmA_lo, a->a_lo, 3, a8_dims, a_strides, a8_box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_UINT8))) return rc;
    const uint64_t b8_dims[2] = {(uint64_t)a->K * 2, (uint64_t)a->w_rows};
    const uint32_t b8_box[2] = {2 * GEMM_BLOCK_K, (uint32_t)(BLOCK_N / CLUSTER)};
    if ((rc = make_tmap(&tmB_lo, a->w_lo, 2, b8_dims, b_strides, b8_box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_UINT8))) return rc;
  }
  }
  GemmParams p = make_gemm_params(a, BLOCK_N);
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, PASSES, CLUSTER>;
  static unsigned long long smem_attr_done = 0;   // per template instantiation, one bit per device
  W2V2_CUDA(ensure_dyn_smem(kern, S::TOTAL, smem_attr_done));
  int dev = 0, sms = 0;
  W2V2_CUDA(cudaGetDevice(&dev));
  W2V2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total_m_tiles = p.batch * p.tiles_per_batch;
  const int total_work = ((total_m_tiles + CLUSTER - 1) / CLUSTER) * p.n_tiles;
  int grid = total_work * CLUSTER < sms ? total_work * CLUSTER : sms;
  if (a->max_ctas > 0 && grid > a->max_ctas) grid = a->max_ctas;
  grid -= grid % CLUSTER;
  if (grid < CLUSTER) grid = CLUSTER;
  W2V2_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), (size_t)S::TOTAL, stream, CLUSTER > 1 ? CLUSTER : 0, tmA_hi, tmA_lo,
                       tmB_hi, tmB_lo, p));
  return 0;
}

}  // namespace w2v2

namespace w2v2 { extern thread_local bool g_stats_final_in_kernel; }   // gemm_2sm.cu
extern "C" int w2v2_row_stats_finalize(const float* parts, int nparts, int64_t rows, int dim, float eps, float* stats, void* stream);

static int gemm_bf16_dispatch(const w2v2_gemm_args* a, void* stream) {
  using namespace w2v2;
  W2V2_CHECK_ARG(a != nullptr, "args is null");
  W2V2_CHECK_ARG(a->a_hi && a->w_hi, "A / W pointers must be non-null");
  const int np = mode_passes(a->passes);
  W2V2_CHECK_ARG((np == 1 || np == 3) && (a->passes & ~(3 | MODE_FP16 | MODE_F8)) == 0 && (!mode_f8(a->passes) || (np == 1 && mode_fp16(a->passes))),
                 "passes must be 1 (bf16), 3 (bf16x3), 17 (fp16), 19 (fp16x3) or 25 (fp16 + e4m3 cross terms)");
  W2V2_CHECK_ARG((np == 1 && !mode_f8(a->passes)) || (a->a_lo && a->w_lo), "multi-plane modes need the second planes");
  W2V2_CHECK_ARG(a->out_format >= 0 && a->out_format <= 2, "out_format must be 0 (bf16), 1 (fp16) or 2 (fp16 + e4m3 pairs)");
  W2V2_CHECK_ARG(a->out_format != 2 || (a->out_hi && a->out_lo && a->N % 64 == 0), "out_format 2 writes both planes and needs N % 64 == 0");
  W2V2_CHECK_ARG(!mode_fp16(a->passes) || a->scale == nullptr || a->ln_fold_stats != nullptr, "per-column scale is not combined with the scaled fp16 planes");
  W2V2_CHECK_ARG(a->ln_fold_stats == nullptr || (a->scale && a->bias && a->bias_batch_stride == 0),
                 "ln_fold_stats needs scale (= colsum(gamma o W)) and bias (= beta W + b), shared by all batch entries");
  W2V2_CHECK_ARG(a->ln_fold_parts == 0 || a->ln_fold_stats != nullptr, "ln_fold_parts > 0 needs ln_fold_stats");
  W2V2_CHECK_ARG(a->row_stats_out == nullptr || (a->N % 64 == 0 && a->out_f32 != nullptr && !(a->flags & W2V2_GEMM_MN_MAJOR)),
                 "row_stats_out needs out_f32 and N % 64 == 0");
  W2V2_CHECK_ARG(a->res_ln_parts >= 0 && a->ln_fold_parts >= 0 && a->res_ln_parts <= 16 && a->ln_fold_parts <= 16,
                 "partial-sum counts must be in [0, 16] (LayerNorm widths up to 1024)");
  W2V2_CHECK_ARG(a->K > 0 && a->K % GEMM_BLOCK_K == 0, "K must be a positive multiple of 64");
  W2V2_CHECK_ARG(a->N > 0 && a->rows_per_batch > 0 && a->batch > 0, "N, rows_per_batch, batch must be positive");
  W2V2_CHECK_ARG(a->a_row_stride % 8 == 0 && a->a_batch_stride % 8 == 0, "A strides must be multiples of 8 elements (16 B)");
  W2V2_CHECK_ARG(a->a_row_len >= a->K || a->kb_split > 0, "a_row_len must cover K");
  W2V2_CHECK_ARG(a->out_f32 || a->out_hi, "at least one output is required");
  W2V2_CHECK_ARG(a->out_lo == nullptr || a->out_hi != nullptr, "out_lo requires out_hi");
  W2V2_CHECK_ARG(a->row_replace_mask == nullptr || (a->row_replace_value != nullptr && a->N % 16 == 0),
                 "row_replace_mask needs row_replace_value and N % 16 == 0");
  W2V2_CHECK_ARG(a->drop_p >= 0.0f && a->drop_p < 1.0f && (a->drop_p == 0.0f || a->N % 16 == 0), "drop_p must be in [0, 1) (and N % 16 == 0)");
  W2V2_CHECK_ARG(a->res_ln_stats == nullptr || (a->residual && a->res_ln_gamma && a->res_ln_beta && a->N % 4 == 0),
                 "res_ln_stats needs residual, res_ln_gamma, res_ln_beta and N % 4 == 0");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int bn = a->block_n;
  if (a->flags & W2V2_GEMM_MN_MAJOR) {
    // D[m][n] = sum_r X[r][m] Y[r][n]  (weight gradients: X = layer input, Y = output gradient, both row-major as stored)
    W2V2_CHECK_ARG(a->passes == 1 && a->kb_split == 0 && a->out_format == 0, "MN-major mode: single bf16 pass");
    W2V2_CHECK_ARG(a->batch == 1 || (a->out_f32 && !a->out_hi && !a->bias && !a->residual),
                   "MN-major split-K (batch > 1): only out_f32 (accumulated atomically into a zeroed buffer), no bias / residual");
    W2V2_CHECK_ARG(a->a_row_stride % 8 == 0 && a->w_row_stride % 8 == 0 && a->w_row_stride >= a->N && a->a_row_stride >= a->rows_per_batch,
                   "MN-major mode: leading dimensions must cover the tile and be multiples of 8 elements");
    W2V2_CHECK_ARG(a->a_rows > 0 && (int64_t)a->K * a->batch >= a->a_rows, "MN-major mode: K x batch must cover the a_rows reduction rows");
    if (bn == 0 || bn > 128) bn = 128;
    if (bn == 128) return launch_gemm<128, 1, 1>(a, s);
    if (bn == 64) return launch_gemm<64, 1, 1>(a, s);
    return fail(-1, "%s: MN-major mode supports block_n 64 or 128", __func__);
  }
  if (bn == 0) {
    bn = (a->N % 256 == 0) ? 256 : (a->N % 128 == 0) ? 128 : (a->N % 64 == 0) ? 64 : 32;
    // Small problems (fewer 256-wide tiles than a quarter of the SMs): every CTA walks its whole K loop alone, bound by TMA
    // latency / ring depth, and most SMs idle.  64-wide tiles give 4x the CTAs and a 7-stage ring (FFN2 at 49 frames: 3 CTAs x 48
    // k-blocks through 3 stages -> 12 CTAs through 7).  Same accumulation order per output element, so results are bit-identical.
    const long tiles256 = (long)a->batch * ((a->rows_per_batch + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * ((a->N + 255) / 256);
    if (bn > 64 && a->N % 64 == 0 && tiles256 < 37) bn = 64;
  }
  W2V2_CHECK_ARG(a->w_rows >= ((a->N + bn - 1) / bn) * bn, "weight matrix must be padded to a multiple of block_n rows");
  // clusters of 2 (weight-tile multicast) for the wide tiles whenever there are at least two m-tiles
  const long m_tiles = (long)a->batch * ((a->rows_per_batch + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  const bool pair = (bn >= 128) && m_tiles >= 2 && a->cluster != 1;
  const bool can_2sm = (a->N % 8 == 0);
  W2V2_CHECK_ARG(a->out_format != 2 || bn >= 64, "out_format 2 needs block_n >= 64");
  if (mode_f8(a->passes)) {
    switch (bn) {
      case 256: return pair ? (a->cluster == 3 || !can_2sm ? launch_gemm<256, 2, 2>(a, s) : launch_gemm_2sm(a, s)) : launch_gemm<256, 2, 1>(a, s);
      case 128: return pair ? launch_gemm<128, 2, 2>(a, s) : launch_gemm<128, 2, 1>(a, s);
      case 64: return launch_gemm<64, 2, 1>(a, s);
      case 32: return launch_gemm<32, 2, 1>(a, s);
    }
  } else if (np == 1) {
    switch (bn) {
      case 256: return pair ? (a->cluster == 3 || !can_2sm ? launch_gemm<256, 1, 2>(a, s) : launch_gemm_2sm(a, s)) : launch_gemm<256, 1, 1>(a, s);
      case 128: return pair ? launch_gemm<128, 1, 2>(a, s) : launch_gemm<128, 1, 1>(a, s);
      case 64: return launch_gemm<64, 1, 1>(a, s);
      case 32: return launch_gemm<32, 1, 1>(a, s);
    }
  } else {
    switch (bn) {
      case 256: return pair ? (a->cluster == 3 || !can_2sm ? launch_gemm<256, 3, 2>(a, s) : launch_gemm_2sm(a, s)) : launch_gemm<256, 3, 1>(a, s);
      case 128: return pair ? launch_gemm<128, 3, 2>(a, s) : launch_gemm<128, 3, 1>(a, s);
      case 64: return launch_gemm<64, 3, 1>(a, s);
      case 32: return launch_gemm<32, 3, 1>(a, s);
    }
  }
  return fail(-1, "%s: unsupported block_n %ld", __func__, bn);
}

extern "C" int w2v2_gemm_bf16(const w2v2_gemm_args* a, void* stream) {
  using namespace w2v2;
  W2V2_CHECK_ARG(a != nullptr, "args is null");
  W2V2_CHECK_ARG(a->row_stats_final == nullptr || (a->row_stats_out != nullptr && a->row_stats_counter != nullptr &&
                                                   (a->batch == 1 || a->rows_per_batch % 32 == 0)),
                 "row_stats_final needs row_stats_out, row_stats_counter and batch == 1 or rows_per_batch % 32 == 0");
  g_stats_final_in_kernel = false;
  const int rc = gemm_bf16_dispatch(a, stream);
  if (rc != 0 || a->row_stats_final == nullptr || g_stats_final_in_kernel) return rc;
  // tile shapes without the in-kernel finalisation: the separate launch, same stream, same arithmetic
  return w2v2_row_stats_finalize(a->row_stats_out, a->N / 64, (int64_t)a->batch * a->rows_per_batch, a->N, a->ln_eps, a->row_stats_final,
                                 stream);
}

extern "C" const char* w2v2_last_error_string(void) { return w2v2::g_last_error; }
extern "C" int w2v2_version(void) { return 141; }   // 1.41: row_stats_final; 1.4: precision modes 17 / 19 / 25, output formats, LayerNorm fold (ln_fold_*, row_stats_out)
