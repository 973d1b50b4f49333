// Shared device helpers for the sm_100a Wav2Vec2 kernels: raw PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM, packed-fp32 math, the erf-GELU evaluation and
// bf16 hi/lo splitting.  No CUTLASS, no library calls.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace w2v2 {

// --------------------------------------------------------------------------- misc
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// --------------------------------------------------------------------------- programmatic dependent launch
// Every forward-path kernel is launched with programmatic stream serialization: it may START while its predecessor
// is still draining (prologue: barrier init, TMEM alloc, tensor-map prefetch overlap the predecessor's tail), and
// must call pdl_wait() before touching global memory.  pdl_trigger() lets the successor start as early as possible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// --------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trapped launch (cudaErrorLaunchFailure) after ~2 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
}

// --------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// 2-D tile load multicast to every CTA of the cluster named in cta_mask (same smem offset, same mbarrier offset).
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// --------------------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive (optionally with expected transaction bytes) on an mbarrier that may live in a peer CTA
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
// cta_group::2 tile loads: data lands in THIS CTA's smem, the transaction bytes are reported to the mbarrier at
// `bar_cluster_addr` (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// --------------------------------------------------------------------------- tcgen05 / TMEM
template <int kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows per CTA), B's N rows split across the CTAs.
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16/fp16 in, fp32 accumulate), one CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// e4m3 x e4m3 -> fp32 (K = 32 per instruction), one CTA
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows = TMEM lanes, K packed two 16-bit elements per
// 32-bit column) is read from tensor memory - used for O += P V with the probabilities written by tcgen05.st.
__device__ __forceinline__ void umma_f16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, but the arrive is multicast to the mbarrier at the same offset in every CTA of cta_mask.
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- shared-memory matrix descriptors (PTX ISA "tcgen05 shared memory descriptor") ----
// K-major operand, 128-byte swizzle: rows are 128 B apart, 8-row groups are SBO = 1024 B apart.
// start address may be advanced by k*32 B inside the 128 B swizzle row (UMMA_K = 16 bf16).
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major operand, no swizzle: 8x16B core matrices; rows of a core matrix 16 B apart,
// 8-row groups SBO apart, the two 16-byte K chunks of one MMA LBO apart.
__device__ __forceinline__ uint64_t desc_kmajor_noswz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// MN-major operand, 128-byte swizzle (64 contiguous MN elements per row of 128 B; 8 K-rows per
// 1024 B atom): LBO = distance between 64-wide MN atoms, SBO = distance between 8-row K groups.
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same with the operand format fields cleared: fp16 x fp16 -> fp32 under kind::f16, and e4m3 x e4m3 -> fp32 under
// kind::f8f6f4 (format code 0 is F16 in the first table and E4M3 in the second).
__host__ __device__ constexpr uint32_t idesc_fmt0(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t idesc_16bit(bool fp16, int M, int N, int a_mn_major, int b_mn_major) {
  return fp16 ? idesc_fmt0(M, N, a_mn_major, b_mn_major) : idesc_bf16(M, N, a_mn_major, b_mn_major);
}

// ---- precision modes (the `passes` argument of the C ABI, see include/w2v2.h) ----
//   1 = bf16, 3 = bf16x3 (hi*hi + lo*hi + hi*lo), | 16 = fp16 operands (activations stored x 2^4, weights x 2^11),
//   | 8 = fp16 main product + the two cross terms as e4m3 MMAs ("fp16f8", planes: fp16 + interleaved e4m3 pairs)
constexpr int MODE_FP16 = 16, MODE_F8 = 8;
__host__ __device__ constexpr int mode_passes(int mode) { return mode & 3; }
__host__ __device__ constexpr bool mode_fp16(int mode) { return (mode & MODE_FP16) != 0; }
__host__ __device__ constexpr bool mode_f8(int mode) { return (mode & MODE_F8) != 0; }
constexpr float ACT_SCALE = 16.0f;          // fp16 activation planes hold x * 2^4 ...
constexpr float WGT_SCALE = 2048.0f;        // ... fp16 weight planes w * 2^11: accumulators are at 2^15
constexpr float ACC_UNSCALE = 1.0f / 32768.0f;
constexpr float F8_UP = 64.0f, F8_DOWN = 1.0f / 64.0f;   // e4m3 cross-term operands: lo * 2^6, hi * 2^-6

// --------------------------------------------------------------------------- math
// erf-GELU without erff():  gelu(x) = max(x,0) - 0.5*|x|*erfc(|x|/sqrt2),
// erfc(a/sqrt2) = exp2(a*R(a)) with R a degree-5 minimax fit on a in [0,6] (a clamped to 6).
// fp32 max-abs error vs 0.5x(1+erf(x/sqrt2)): 2.6e-7 over |x| <= 14 (fit script: DESIGN.md).
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 6.0f);
  float r = 3.2121541153173894e-05f;
  r = fmaf(r, a, -0.0007558754878118634f);
  r = fmaf(r, a, 0.008020005188882351f);
  r = fmaf(r, a, -0.053288985043764114f);
  r = fmaf(r, a, -0.45888903737068176f);
  r = fmaf(r, a, -1.1511517763137817f);
  const float e = exp2f(r * a);  // -> MUFU.EX2 (plus range scaling) with fast-math off; see ex2_approx
  return fmaf(-0.5f * fabsf(x), e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float ax = fabsf(x);
  const float a = fminf(ax, 6.0f);
  float r = 3.2121541153173894e-05f;
  r = fmaf(r, a, -0.0007558754878118634f);
  r = fmaf(r, a, 0.008020005188882351f);
  r = fmaf(r, a, -0.053288985043764114f);
  r = fmaf(r, a, -0.45888903737068176f);
  r = fmaf(r, a, -1.1511517763137817f);
  const float e = ex2_approx(r * a);
  return fmaf(-0.5f * ax, e, fmaxf(x, 0.0f));
}

// ---- packed fp32x2 (Blackwell FFMA2 / FMUL2 / FADD2) ----
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// Two GELUs at once on the packed fp32x2 pipe.  Epilogue form: 0.5*erfc(a/sqrt2) = exp2(a*R4(a) - 1) with a degree-4
// R (fp32 max-abs error of gelu 1.7e-6, relative 1.6e-4; tools/fit_gelu.py), so gelu(x) = max(x,0) - |x| * exp2(.):
// per pair 5 FFMA2 + 2 MUFU.EX2 + 4 FMNMX + 2 FFMA.
__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) {
  const float ax0 = fabsf(x0), ax1 = fabsf(x1);
  const uint64_t a = pack2(fminf(ax0, 6.0f), fminf(ax1, 6.0f));
  uint64_t r = pack2(-0.0004135944473091513f, -0.0004135944473091513f);
  r = fma2(r, a, pack2(0.006748747080564499f, 0.006748747080564499f));
  r = fma2(r, a, pack2(-0.051224492490291595f, -0.051224492490291595f));
  r = fma2(r, a, pack2(-0.4603409171104431f, -0.4603409171104431f));
  r = fma2(r, a, pack2(-1.1508060693740845f, -1.1508060693740845f));
  r = fma2(r, a, pack2(-1.0f, -1.0f));
  float p0, p1;
  unpack2(r, p0, p1);
  const float e0 = ex2_approx(p0), e1 = ex2_approx(p1);
  x0 = fmaf(-ax0, e0, fmaxf(x0, 0.0f));
  x1 = fmaf(-ax1, e1, fmaxf(x1, 0.0f));
}

// bf16-grade GELU for the single-pass (throughput) mode, two at once on the packed pipe: gelu(x) ~= hx + hx * tanh(x (A + B x^2)),
// hx = x / 2, (A, B) a minimax refit of the tanh form against the erf form (formula error 2.7e-4, + MUFU.TANH's 2^-11:
// < 5e-4 absolute for |x| < 2, i.e. below the 2^-9 relative rounding of the bf16 output; tools/fit_gelu.py --tanh).
// 3 FMUL2 + 2 FFMA2 + 2 MUFU.TANH per pair - about half the issue slots of gelu_erf_x2.  Parity mode never uses it.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t gelu_tanh_p2(uint64_t v) {
  const uint64_t x2 = mul2(v, v);
  const uint64_t p = fma2(x2, pack2(0.03471371f, 0.03471371f), pack2(0.80012897f, 0.80012897f));
  const uint64_t u = mul2(v, p);
  float u0, u1;
  unpack2(u, u0, u1);
  const uint64_t t = pack2(tanh_approx(u0), tanh_approx(u1));
  const uint64_t h = mul2(v, pack2(0.5f, 0.5f));
  return fma2(h, t, h);
}
// tf.nn.gelu(approximate=True) - the reference's `is_gelu_approx` switch (config.py:14, feature_extractor.py:58,
// encoder.py:127): 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3), evaluated as x / (1 + exp(-2u)) with
// ex2.approx (2^-22 relative) so that it is fp32-grade in every precision mode (MUFU.TANH alone is only 2^-11).
__device__ __forceinline__ float gelu_tanh_tf(float x) {
  const float u = x * fmaf(x * x, 0.0356774081f, 0.7978845608f);
  const float e = ex2_approx(-2.8853900818f * u);   // exp(-2u)
  return __fdividef(x, 1.0f + e);
}
template <bool FAST>
__device__ __forceinline__ void gelu_x2(float& x0, float& x1) {
  if (FAST) {
    unpack2(gelu_tanh_p2(pack2(x0, x1)), x0, x1);
  } else {
    gelu_erf_x2(x0, x1);
  }
}

// --------------------------------------------------------------------------- dropout RNG
// Counter-based: 64 random bits per group of FOUR consecutive elements = hash(seed, site, element_index / 4); element e
// of the group is KEPT when its 16-bit lane >= thr16 (thr16 = round(p * 65536)).  Stateless, so the backward pass (and the
// test oracle, through w2v2_dropout_mask) regenerates exactly the mask of the forward pass.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {   // murmur3 finaliser: full avalanche on 32 bits
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool drop_keep(uint64_t bits, int e, uint32_t thr16) {
  return ((uint32_t)(bits >> (16 * e)) & 0xFFFFu) >= thr16;
}
struct DropSpec {       // thr16 == 0: dropout off
  uint32_t key0, key1;  // host-side hashes of (seed, site): the per-stream keys
  uint32_t thr16;
  float scale;          // 1 / (1 - p)
};
// 64 bits for group `group` of stream (key0, key1): two murmur3 finalisers on 32-bit integer multiplies (a 64-bit splitmix
// per element pair cost 2.2 ms per train step inside the attention backward)
__device__ __forceinline__ uint64_t drop_bits4(const DropSpec& d, uint64_t group) {
  const uint32_t h = (uint32_t)group ^ ((uint32_t)(group >> 32) * 0x85EBCA77u);
  const uint32_t r0 = mix32(h ^ d.key0);
  const uint32_t r1 = mix32((h + 0x9E3779B9u) ^ d.key1);
  return (uint64_t)r0 | ((uint64_t)r1 << 32);
}
inline DropSpec make_drop(float p, uint64_t seed, uint32_t site) {
  DropSpec d;
  d.key0 = mix32((uint32_t)seed ^ mix32(site * 0x9E3779B9u + 0x7F4A7C15u));
  d.key1 = mix32((uint32_t)(seed >> 32) + 0x6A09E667u + mix32(d.key0 ^ site));
  d.thr16 = (p > 0.0f) ? (uint32_t)(p * 65536.0f + 0.5f) : 0u;
  if (d.thr16 > 65535u) d.thr16 = 65535u;
  d.scale = (d.thr16 > 0u) ? 65536.0f / (65536.0f - (float)d.thr16) : 1.0f;
  return d;
}
// attention-probability dropout (encoder.py:42): the group index of key k of query row (bh, q) is row_group + (k >> 2),
// row_group = (bh * T + q) * ceil(T / 4); the lane is k & 3.
__device__ __forceinline__ uint64_t attn_row_group(int bh, int q, int T) {
  return ((uint64_t)bh * T + q) * (uint64_t)((T + 3) >> 2);
}

// ---- bf16 packing and hi/lo splitting (x ~= hi + lo, both bf16: ~16 mantissa bits) ----
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return r;
}
__device__ __forceinline__ float bf16_lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }
// Returns packed hi pair; writes packed lo pair (residual after rounding to bf16).
__device__ __forceinline__ uint32_t split_bf16x2(float v0, float v1, uint32_t& lo_pair) {
  const uint32_t hi = pack_bf16x2(v0, v1);
  lo_pair = pack_bf16x2(v0 - bf16_lo_to_f32(hi), v1 - bf16_hi_to_f32(hi));
  return hi;
}

// ---- fp16 / e4m3 operand planes (fp16 and fp16f8 modes) ----
// h = fp16(sat(x)) pair; x is ALREADY multiplied by the plane scale (ACT_SCALE for activations)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
__device__ __forceinline__ float f16_lo_to_f32(uint32_t packed) {
  return __half2float(__ushort_as_half((unsigned short)(packed & 0xFFFFu)));
}
__device__ __forceinline__ float f16_hi_to_f32(uint32_t packed) {
  return __half2float(__ushort_as_half((unsigned short)(packed >> 16)));
}
// Returns the packed fp16 hi pair of (v0, v1); lo_pair = packed fp16 residuals (fp16x3 mode)
__device__ __forceinline__ uint32_t split_f16x2(float v0, float v1, uint32_t& lo_pair) {
  const uint32_t hi = pack_f16x2(v0, v1);
  lo_pair = pack_f16x2(v0 - f16_lo_to_f32(hi), v1 - f16_hi_to_f32(hi));
  return hi;
}
// two e4m3 bytes (v0 -> low byte)
__device__ __forceinline__ uint16_t pack_e4m3x2(float v0, float v1) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(v1), "f"(v0));   // first source -> upper byte
  return r;
}
// fp16f8 planes of a value pair: returns the packed fp16 hi pair; l8 = e4m3((v - hi) * 2^6) pair, h8 = e4m3(hi * 2^-6) pair
__device__ __forceinline__ uint32_t split_f16_f8x2(float v0, float v1, uint16_t& l8, uint16_t& h8) {
  const uint32_t hi = pack_f16x2(v0, v1);
  const float h0 = f16_lo_to_f32(hi), h1 = f16_hi_to_f32(hi);
  l8 = pack_e4m3x2((v0 - h0) * F8_UP, (v1 - h1) * F8_UP);
  h8 = pack_e4m3x2(h0 * F8_DOWN, h1 * F8_DOWN);
  return hi;
}

// one half of that: HALF == 0 -> the two residual bytes, HALF == 1 -> the two hi bytes (v already scaled by ACT_SCALE)
template <int HALF>
__device__ __forceinline__ uint32_t f8_pair_of(float v0, float v1) {
  const uint32_t hi = pack_f16x2(v0, v1);
  const float h0 = f16_lo_to_f32(hi), h1 = f16_hi_to_f32(hi);
  return HALF == 0 ? pack_e4m3x2((v0 - h0) * F8_UP, (v1 - h1) * F8_UP) : pack_e4m3x2(h0 * F8_DOWN, h1 * F8_DOWN);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace w2v2
