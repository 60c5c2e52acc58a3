// Host-side plumbing shared by the C-ABI translation units: error reporting, driver entry points,
// TMA tensor-map construction (cached).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hoigen_b200.h"

namespace hoigen {

void set_error(const char* fmt, ...);
int num_sms();   // of the calling thread's current device

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device and several
// devices may be driven from one process (one module per device), possibly from several host threads.
int set_max_dynamic_smem(const void* kernel, int bytes);

// 2-D / 3-D bf16 row-major tensor maps with 128-byte swizzle. dims/box innermost-first, strides in
// BYTES for dims 1.. (dim 0 is contiguous).  Returns nullptr (and sets the error) on failure.
const CUtensorMap* get_tmap_2d_bf16(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                                    uint32_t box0, uint32_t box1);
const CUtensorMap* get_tmap_3d_bf16(const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                                    uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                                    uint32_t box1, uint32_t box2);

// Scratch of the CTA-pair GEMM's stream-K split: fp32 partial-accumulator slots + one flag per slot (zeroed).  One
// workspace per stream (launches on one stream are ordered, launches on different streams must not share slots).
struct StreamKWorkspace {
  float* slots;
  int* flags;
};
StreamKWorkspace get_streamk_workspace(cudaStream_t stream, size_t slot_bytes, int num_flags);

// CTA-pair form of the fused cache kernel (cache_fused2.cu); fills parts [3 * nsplit][ktot_pad][c_pad]
int launch_cache_fused_pair(const hoigen_score_weights* w, const void* pair_feat_bf16, const float* const* cache_bias, int ktot,
                            int affinity, float beta, float* parts, int* nsplit_out, int* ktot_pad_out, int* c_pad_out,
                            cudaStream_t s);

// RoIAlign + mean as a 3 x bf16-split tensor-core product (roi_tc.cu); computes the per-axis weights from the boxes itself
int launch_roi_features_tc(const float* tokens, const float* boxes, const int* box_off, const int* pair_off, int batch,
                           float spatial_scale, float* single_feat, float* union_feat, cudaStream_t s);

// Counts every kernel launch of this library and, when profiling is enabled (hoigen_profile_enable), brackets
// the launch with CUDA events on the launching stream. flops / bytes are the ALGORITHMIC work of the launch.
struct KernelScope {
  KernelScope(const char* tag, cudaStream_t stream, double flops = 0.0, double bytes = 0.0);
  ~KernelScope();
  int slot_;
  cudaStream_t stream_;
};

#define HOIGEN_CHECK_ARG(cond, ...)        \
  do {                                     \
    if (!(cond)) {                         \
      ::hoigen::set_error(__VA_ARGS__);    \
      return HOIGEN_ERR_INVALID;           \
    }                                      \
  } while (0)

#define HOIGEN_CHECK_CUDA(expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::hoigen::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return HOIGEN_ERR_CUDA;                                                                \
    }                                                                                        \
  } while (0)

#define HOIGEN_TRY_RC(expr)               \
  do {                                    \
    int _rc = (expr);                     \
    if (_rc != HOIGEN_OK) return _rc;     \
  } while (0)

#define HOIGEN_CHECK_LAUNCH() HOIGEN_CHECK_CUDA(cudaGetLastError())

}  // namespace hoigen
