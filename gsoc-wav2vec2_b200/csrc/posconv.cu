// Grouped positional convolution for sm_100a:
//     out[b,t,:] = resid[b,t,:] + gelu( bias + sum_{j<k} W_j . x[b, t + j - k/2, group slice] )
// Reference: PositionalConvEmbedding.call encoder.py:177-181 (weight-normalised grouped Conv1D with explicit
// zero padding k/2 on both sides, VALID conv, drop the last frame when k is even, GELU) and the residual add
// of Wav2Vec2Encoder.call encoder.py:265.  The weight normalisation (tensorflow_addons.py:16-21) is a
// per-tap rescale of the kernel and is folded into the packed weights by the host at load time.
//
// Tensor-core formulation without im2col: for one group (cpg = 48 or 64 channels) the A operand of tap j
// is the SAME smem-resident window of x shifted down by j rows.  The window is kept in the canonical
// no-swizzle K-major core-matrix layout  [channel chunk of 8][row][8 elements]  (rows 16 B apart), so
// "shift by j rows" is just +16*j bytes on the descriptor start address.  One CTA owns MT*128 frames of one
// (batch, group): the window has MT*128 + k - 1 rows, the k taps' weights stream through a bulk-copy ring,
// and k * cpg/16 MMAs of shape 128 x cpg x 16 per m-tile accumulate into TMEM.
#include "host_util.h"
#include "w2v2_common.cuh"
#include "../../include/w2v2.h"

namespace w2v2 {

constexpr int PC_THREADS = 256;
constexpr int PC_TAPS_PER_STAGE = 4;
constexpr int PC_HALF_ROWS = 192;              // TMA box rows; the window is loaded as 192-row pieces
constexpr int PC_WIN_ROWS = 2 * PC_HALF_ROWS;  // 384 >= 2*128 + 127

struct PosconvParams {
  int T, d, cpg, groups, ktaps, mt;  // mt = m-tiles (128 frames each) per CTA
  int stages;
  int shift;              // window starts at frame tf0 - ktaps/2 + shift (0 = forward; +1 = transposed conv of the backward)
  int linear;             // 1: out = resid + acc (no bias, no GELU): the dgrad of the conv
  int gelu_approx;        // 1: tf.nn.gelu(approximate=True) (config.py:14, encoder.py:181)
  int fp16;               // 1: x / w planes are fp16 (x * 2^4, w * 2^11): accumulators are un-scaled by acc_scale
  float acc_scale;
  float* pre_out;         // optional fp32 [B, T, d]: bias + conv (pre-activation kept for the backward)
  const float* bias;      // [d]
  const float* resid;     // fp32 [B, T, d]
  float* out_f32;         // fp32 [B, T, d]
  const __nv_bfloat16* w_hi;  // packed [groups][ktaps][cpg/8][cpg][8]
  const __nv_bfloat16* w_lo;
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int PASSES>
__global__ void __launch_bounds__(PC_THREADS, 2)   // two CTAs per SM when the smem budget allows: one's epilogue / window load overlaps the other's MMAs
posconv_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const PosconvParams p) {
  constexpr int NPL = (PASSES == 3) ? 2 : 1;
  constexpr int MAX_STAGES = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int cpc = p.cpg / 8;                               // 16-byte channel chunks per group
  const uint32_t a_lbo = PC_WIN_ROWS * 16;                 // bytes between channel chunks of the window
  const uint32_t a_plane = (uint32_t)cpc * a_lbo;          // one window plane
  const uint32_t tap_bytes = (uint32_t)p.cpg * p.cpg * 2;  // one tap of one group
  const uint32_t stage_plane = PC_TAPS_PER_STAGE * tap_bytes;
  const uint32_t stage_bytes = NPL * stage_plane;
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + NPL * a_plane;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + p.stages * stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* full_bar = bars + 1;
  uint64_t* empty_bar = bars + 1 + MAX_STAGES;
  uint64_t* acc_full = bars + 1 + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * MAX_STAGES);

  const int warp = threadIdx.x >> 5;
  const int tf0 = blockIdx.x * (p.mt * 128);  // first frame of this CTA
  const int g = blockIdx.y;
  const int b = blockIdx.z;
  const int nsteps = p.ktaps / PC_TAPS_PER_STAGE;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_hi);
    if (PASSES == 3) tma_prefetch_desc(&tm_lo);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(a_full, 1);
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // window: rows [tf0 - k/2, tf0 - k/2 + 384), zero-filled outside [0, T)
      mbar_arrive_expect_tx(a_full, NPL * a_plane);
      for (int pl = 0; pl < NPL; ++pl) {
        const CUtensorMap* tm = pl ? &tm_lo : &tm_hi;
        for (int c = 0; c < cpc; ++c)
          for (int q = 0; q < 2; ++q)
            tma_load_4d(a_smem + pl * a_plane + c * a_lbo + q * PC_HALF_ROWS * 16, tm, a_full, 0,
                        tf0 - p.ktaps / 2 + p.shift + q * PC_HALF_ROWS, g * cpc + c, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < nsteps; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        const size_t goff = ((size_t)g * p.ktaps + (size_t)s * PC_TAPS_PER_STAGE) * (size_t)(p.cpg * p.cpg);
        bulk_g2s(b_smem + stage * stage_bytes, p.w_hi + goff, stage_plane, &full_bar[stage]);
        if (PASSES == 3) bulk_g2s(b_smem + stage * stage_bytes + stage_plane, p.w_lo + goff, stage_plane, &full_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = idesc_16bit(p.fp16 != 0, 128, p.cpg, 0, 0);
      const uint32_t b_lbo = (uint32_t)p.cpg * 16;
      const int ksteps = p.cpg / 16;
      mbar_wait(a_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int s = 0; s < nsteps; ++s) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(b_smem + stage * stage_bytes);
        const uint32_t sa = smem_u32(a_smem);
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
          const uint32_t a_pl = sa + ((pass == 1) ? a_plane : 0);
          const uint32_t b_pl = sb + ((pass == 2) ? stage_plane : 0);
          for (int jj = 0; jj < PC_TAPS_PER_STAGE; ++jj) {
            const int j = s * PC_TAPS_PER_STAGE + jj;
            for (int mt = 0; mt < p.mt; ++mt) {
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t da = desc_kmajor_noswz(a_pl + (2 * kk) * a_lbo + (mt * 128 + j) * 16, a_lbo, 128);
                const uint64_t db = desc_kmajor_noswz(b_pl + jj * tap_bytes + (2 * kk) * b_lbo, b_lbo, 128);
                // the first MMA into each m-tile's accumulator overwrites, everything after accumulates
                umma_f16(tmem_base + mt * p.cpg, da, db, idesc, (s | pass | jj | kk) != 0);
              }
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int lane = lane_id();
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int mt = 0; mt < p.mt; ++mt) {
      const int t = tf0 + mt * 128 + ew * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + mt * p.cpg;
      for (int c0 = 0; c0 < p.cpg; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(taddr + c0, r);
        tmem_ld_wait();
        if (t < p.T) {
          const int n = g * p.cpg + c0;
          const size_t o = ((size_t)b * p.T + t) * p.d + n;
          float v[16];
          if (p.linear) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * p.acc_scale;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + n) + i);
              v[4 * i + 0] = fmaf(__uint_as_float(r[4 * i + 0]), p.acc_scale, bb.x);   // fmaf(a, 1, b) == a + b
              v[4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), p.acc_scale, bb.y);
              v[4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), p.acc_scale, bb.z);
              v[4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), p.acc_scale, bb.w);
            }
            if (p.pre_out != nullptr) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(p.pre_out + o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.gelu_approx) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = gelu_tanh_tf(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 2) gelu_erf_x2(v[i], v[i + 1]);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(p.resid + o) + i);
            reinterpret_cast<float4*>(p.out_f32 + o)[i] =
                make_float4(v[4 * i] + rr.x, v[4 * i + 1] + rr.y, v[4 * i + 2] + rr.z, v[4 * i + 3] + rr.w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<128>(tmem_base);
}

template <int PASSES>
static int launch_posconv(const w2v2_posconv_args* a, cudaStream_t stream) {
  constexpr int NPL = (PASSES == 3) ? 2 : 1;
  const int cpg = a->hidden / a->groups;
  const int cpc = cpg / 8;
  CUtensorMap tm_hi, tm_lo;
  // x viewed as {8 elements, T rows, d/8 chunks, B}
  const uint64_t dims[4] = {8, (uint64_t)a->frames, (uint64_t)a->hidden / 8, (uint64_t)a->batch};
  const uint64_t strides[3] = {(uint64_t)a->hidden * 2, 16, (uint64_t)a->frames * a->hidden * 2};
  const uint32_t box[4] = {8, PC_HALF_ROWS, 1, 1};
  int rc = make_tmap(&tm_hi, a->x_hi, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc) return rc;
  tm_lo = tm_hi;
  if (PASSES == 3) {
    rc = make_tmap(&tm_lo, a->x_lo, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  PosconvParams p;
  p.T = a->frames;
  p.d = a->hidden;
  p.cpg = cpg;
  p.groups = a->groups;
  p.ktaps = a->ktaps;
  p.mt = (a->frames > 128) ? 2 : 1;
  p.shift = a->shift;
  p.linear = a->linear;
  p.gelu_approx = a->gelu_approx;
  p.fp16 = mode_fp16(a->passes) ? 1 : 0;
  p.acc_scale = mode_fp16(a->passes) ? ACC_UNSCALE : 1.0f;
  p.pre_out = a->pre_out;
  p.bias = a->bias;
  p.resid = a->resid;
  p.out_f32 = a->out_f32;
  p.w_hi = reinterpret_cast<const __nv_bfloat16*>(a->w_hi);
  p.w_lo = reinterpret_cast<const __nv_bfloat16*>(a->w_lo);
  const int a_bytes = NPL * cpc * PC_WIN_ROWS * 16;
  const int stage_bytes = NPL * PC_TAPS_PER_STAGE * cpg * cpg * 2;
  // prefer a footprint that lets TWO CTAs share an SM (113 KB each, TMEM 2 x 128 columns): the kernel has no intra-CTA overlap
  // between its window load, its 128-tap MMA chain and its epilogue
  int stages = (115712 - 1280 - a_bytes) / stage_bytes;
  if (stages < 2) stages = (232448 - 1280 - a_bytes) / stage_bytes;
  if (stages > 4) stages = 4;
  if (stages < 1) return fail(-1, "%s: shared memory budget too small for %ld channels per group", __func__, cpg);
  p.stages = stages;
  const int smem_bytes = a_bytes + stages * stage_bytes + 256 + 1024;
  auto kern = posconv_kernel<PASSES>;
  W2V2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  dim3 grid((a->frames + p.mt * 128 - 1) / (p.mt * 128), a->groups, a->batch);
  W2V2_CUDA(launch_pdl(kern, grid, dim3(PC_THREADS), (size_t)smem_bytes, stream, 0, tm_hi, tm_lo, p));
  return 0;
}

}  // namespace w2v2

extern "C" int w2v2_posconv(const w2v2_posconv_args* a, void* stream) {
  using namespace w2v2;
  W2V2_CHECK_ARG(a != nullptr, "args is null");
  W2V2_CHECK_ARG(a->x_hi && a->w_hi && (a->bias || a->linear) && a->resid && a->out_f32, "null pointer");
  W2V2_CHECK_ARG(a->passes == 1 || a->passes == 3 || a->passes == 17 || a->passes == 19, "passes must be 1, 3 (bf16) or 17, 19 (fp16)");
  W2V2_CHECK_ARG(mode_passes(a->passes) == 1 || (a->x_lo && a->w_lo), "3-pass modes need the lo planes");
  W2V2_CHECK_ARG(a->groups > 0 && a->hidden % a->groups == 0, "hidden must be divisible by groups");
  const int cpg = a->hidden / a->groups;
  W2V2_CHECK_ARG(cpg % 16 == 0 && cpg >= 16 && cpg <= 64, "channels per group must be 16, 32, 48 or 64");
  W2V2_CHECK_ARG(a->ktaps % PC_TAPS_PER_STAGE == 0 && a->ktaps % 2 == 0 && a->ktaps <= 128,
                 "ktaps must be even, a multiple of 4 and at most 128");
  W2V2_CHECK_ARG(a->batch > 0 && a->frames > 0, "batch and frames must be positive");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (mode_passes(a->passes) == 1) return launch_posconv<1>(a, s);
  return launch_posconv<3>(a, s);
}
