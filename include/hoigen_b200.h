/*
 * hoigen_b200 — C ABI of the B200 (sm_100a) HOI-scoring forward path.
 *
 * The reference (soberguo/HOIGen) is pure Python/PyTorch and has no FFI of its own; its "operator
 * API" for this path is three Python call surfaces:
 *   - build_detector(...)                      upt_tip_cache_model_free_finetune_distill3.py:1712
 *   - UPT.forward(images, targets=None)        upt_tip_cache_model_free_finetune_distill3.py:1543
 *   - VisionTransformer.forward(x, prior)      CLIP_models_adapter_prior2.py:489
 * Every entry point below replaces the library calls one stage of those functions dispatches to;
 * the file:line each one replaces is cited on the declaration.  hoigen_b200/ (Python) mirrors the
 * three surfaces on top of this ABI through ctypes; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers + sizes + a cudaStream_t (passed as void*).
 *   - never allocates device memory, never synchronises; the caller (torch) owns every buffer.
 *   - returns 0 on success, a negative hoigen_status otherwise; hoigen_last_error() gives the text.
 *   - "bf16" buffers are raw uint16 (__nv_bfloat16) arrays; "tokens" are rows b*197+t.
 */
#ifndef HOIGEN_B200_H_
#define HOIGEN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOIGEN_ABI_VERSION 1
#define HOIGEN_API __attribute__((visibility("default")))

typedef void* hoigen_stream_t; /* cudaStream_t */

enum hoigen_status {
  HOIGEN_OK = 0,
  HOIGEN_ERR_INVALID = -1,   /* bad argument (shape / alignment / null pointer) */
  HOIGEN_ERR_CUDA = -2,      /* CUDA runtime / driver error, see hoigen_last_error() */
  HOIGEN_ERR_ARCH = -3,      /* device is not sm_100 */
  HOIGEN_ERR_CAPACITY = -4   /* caller-provided output buffer too small */
};

enum hoigen_act { HOIGEN_ACT_NONE = 0, HOIGEN_ACT_QUICKGELU = 1, HOIGEN_ACT_RELU = 2 };

HOIGEN_API int hoigen_abi_version(void);
HOIGEN_API const char* hoigen_last_error(void);
/* Resolve driver entry points, check the device is sm_100, raise shared-memory limits. */
HOIGEN_API int hoigen_init(int device);

/* ------------------------------------------------------------------------------------------------
 * GEMM: out = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, fp32 accumulate (TMA + tcgen05 + TMEM).
 * Replaces every nn.Linear / F.linear / `@` on the path: conv1-as-GEMM C:491, in_proj/out_proj
 * C:443-445, c_fc/c_proj C:428-432, adapter down/up C:184,201, `@ proj` C:505, and the cache /
 * text GEMMs U:1156-1163.
 *   v = acc (+ bias[n]) ; v = act(v) ; v *= colscale[n] ; v += residual[m,n]
 *   then out_f32[m,n] = v and/or out_bf16[m,n] = bf16(v).   residual may alias out_f32.
 * lda/ldw in elements, multiples of 8 (16-byte TMA strides); a, w 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;         /* bf16 [M, lda] */
  const void* w;         /* bf16 [N, ldw] */
  int32_t M, N, K;
  int32_t lda, ldw;
  const float* bias;     /* [N] or NULL */
  const float* colscale; /* [N] or NULL */
  int32_t act;           /* hoigen_act */
  const float* residual; /* fp32 [M, ld_res] or NULL */
  int32_t ld_res;
  float* out_f32;        /* fp32 [M, ld_f32] or NULL */
  int32_t ld_f32;
  void* out_bf16;        /* bf16 [M, ld_bf16] or NULL */
  int32_t ld_bf16;
  int32_t block_n;       /* 0 = choose; else 64 / 128 / 256 */
} hoigen_gemm_params;

HOIGEN_API int hoigen_gemm_bf16(const hoigen_gemm_params* p, hoigen_stream_t stream);
/* Test-only SIMT cross-check of the same contract (never used by the product path). */
HOIGEN_API int hoigen_debug_gemm_simt(const hoigen_gemm_params* p, hoigen_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HOIGEN_B200_H_ */
