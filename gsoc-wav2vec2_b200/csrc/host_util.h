// Host-side helpers shared by the C-ABI launchers: error reporting and TMA tensor-map encoding
// through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace w2v2 {

extern thread_local char g_last_error[512];

inline int fail(int code, const char* fmt, const char* a = "", long x = 0, long y = 0) {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, x, y);
  return code;
}

#define W2V2_CHECK_ARG(cond, msg)                                                       \
  do {                                                                                  \
    if (!(cond)) return ::w2v2::fail(-1, "%s: argument check failed: " msg, __func__);  \
  } while (0)

#define W2V2_CUDA(call)                                                                            \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return ::w2v2::fail((int)e__, "%s: CUDA error %ld", __func__, (long)e__); \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn();

// rank-N bf16 (or other 2-byte) tensor map.  dims/box innermost first; strides in BYTES for dims 1..rank-1.
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, CUtensorMapSwizzle swz, CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);

// The dynamic-smem opt-in is a PER-DEVICE function attribute: remember it per device (bit d of `done_mask`), not per process.
template <typename K>
inline cudaError_t ensure_dyn_smem(K kern, int bytes, unsigned long long& done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done_mask |= bit;
  return e;
}

bool pdl_enabled();  // W2V2_PDL=0 disables programmatic dependent launch (A/B switch)

// Launch with programmatic stream serialization (+ optional static cluster size).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 0) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace w2v2
