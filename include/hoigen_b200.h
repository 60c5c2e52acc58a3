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

#define HOIGEN_ABI_VERSION 2   /* 2: hoigen_gemm_params gained a2 / lda2 / k2 / conv_stride */
#define HOIGEN_API __attribute__((visibility("default")))

typedef void* hoigen_stream_t; /* cudaStream_t */

enum hoigen_status {
  HOIGEN_OK = 0,
  HOIGEN_ERR_INVALID = -1,   /* bad argument (shape / alignment / null pointer) */
  HOIGEN_ERR_CUDA = -2,      /* CUDA runtime / driver error, see hoigen_last_error() */
  HOIGEN_ERR_ARCH = -3,      /* device is not sm_100 */
  HOIGEN_ERR_CAPACITY = -4   /* caller-provided output buffer too small */
};

enum hoigen_act { HOIGEN_ACT_NONE = 0, HOIGEN_ACT_QUICKGELU = 1, HOIGEN_ACT_RELU = 2, HOIGEN_ACT_EXP = 3 /* exp(act_param * v) */ };

HOIGEN_API int hoigen_abi_version(void);
HOIGEN_API const char* hoigen_last_error(void);
/* Resolve driver entry points, check the device is sm_100, raise shared-memory limits. */
HOIGEN_API int hoigen_init(int device);

/* Launch accounting + optional per-launch CUDA-event timing (used by bench.py for gpu_launches / roofline).
 * hoigen_profile_read writes one line per recorded launch, "tag ms flops bytes start_ms", and returns the byte count. */
HOIGEN_API long long hoigen_launch_count(void);
HOIGEN_API int hoigen_profile_enable(int on);
HOIGEN_API int hoigen_profile_reset(void);
HOIGEN_API long long hoigen_profile_read(char* buf, long long cap);

/* ------------------------------------------------------------------------------------------------
 * GEMM: out = epilogue(A[M,K] @ W[N,K]^T), bf16 operands, fp32 accumulate (TMA + tcgen05 + TMEM).
 * Replaces every nn.Linear / F.linear / `@` on the path: conv1-as-GEMM C:491, in_proj/out_proj
 * C:443-445, c_fc/c_proj C:428-432, adapter down/up C:184,201, `@ proj` C:505, and the cache /
 * text GEMMs U:1156-1163.
 *   v = acc (+ bias[n]) (+ res_bf16[m,n]) ; v = act(v) ; v *= colscale[n] ; v += residual[m,n]
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
  int32_t block_n;       /* 0 = choose; 64/128/256 = one-CTA tiles 128xBN; 2128/2192/2256 = CTA-pair tiles 256xBN;
                            +10000 (tests) = split leftover tiles across SM pairs by k-blocks even when K is short
                            (by default only K >= 2048 is split, where it pays) */
  int32_t split_k;       /* one-CTA tiles only: 0 = choose, n >= 1 = cut every tile's k-range into n units that run on
                            different SMs; partials meet in a workspace and are summed in split order (deterministic) */
  float act_param;       /* HOIGEN_ACT_EXP: beta */
  /* LayerNorm folded into the GEMM (C:443-445, C:457-458: ln_1 -> attn.in_proj, ln_2 -> mlp.c_fc).  a = bf16 copy of
   * the RAW rows x, w = W . diag(gamma) (bf16); out = act(rstd[m] * (a w^T - mean[m] * ln_colsum[n]) + bias[n]) with
   * ln_stats[m] = {mean, rstd} of row m (fp32, eps 1e-5; hoigen_add_rowstats768), ln_colsum[n] = sum_k w[n][k] and
   * bias = b + W beta.  NULL = plain GEMM.  Excludes colscale. */
  const float* ln_stats;   /* (M, 2) */
  const float* ln_colsum;  /* (N) */
  /* Convolutions of the ResNet-50 branch (U:1616-1618: dino_model = torchvision resnet50, eval) on NHWC bf16 activations
   * stored WITH a one-pixel zero halo: rows = the pixels of (B, halo_h, halo_w) = (B, H + 2, W + 2), columns = channels.
   *   1x1 convolution = this GEMM as it is.   3x3 / stride 1 / pad 1 (conv_taps = 9) = nine accumulated products of
   *   ROW-SHIFTED views of the same matrix: a = [M, conv_cin], w = [N, 9 * conv_cin] with k = (ky * 3 + kx) * conv_cin + c,
   *   K = 9 * conv_cin; tap t reads rows m + (t / 3 - 1) * halo_w + (t % 3 - 1) (rows outside [0, M) read as zero) --
   *   an implicit GEMM: no im2col matrix exists.
   * halo_w > 0: output rows on the halo ring are written as 0 (they are the next convolution's zero padding).
   * res_bf16: bf16 [M, ld_resb] added BEFORE the activation (Bottleneck: relu(conv3 + identity)).
   * All zero / NULL = plain GEMM. */
  int32_t conv_taps;       /* 0, 1 or 9 */
  int32_t conv_cin;        /* multiple of 64 when conv_taps = 9 */
  int32_t halo_h, halo_w;
  const void* res_bf16;
  int32_t ld_resb;
  /* Second A source (CTA-pair tiles only): columns [K - k2, K) of the product come from a2 [M, lda2] instead of a -- one GEMM for
   * `conv3(t) + downsample(x)` of a Bottleneck's first block (w = [W3 | Wds] along K, bias = b3 + bds).  K - k2 must be a multiple
   * of 64; excludes res_bf16 and conv_taps = 9.  NULL = off. */
  const void* a2;
  int32_t lda2, k2;
  /* conv_taps = 9 with conv_stride = 2: the 3x3 / stride 2 / pad 1 convolution as an implicit GEMM.  `a` holds the FOUR PHASES of
   * the input, [4][M, lda]: phase (py, px) = 2 py + px is the image in[2 y' + py][2 x' + px] laid out in the OUTPUT's haloed
   * geometry (halo_h, halo_w), zero where the source pixel does not exist (hoigen_conv_gather_s2 with taps = 4).  Tap (ky, kx) of
   * output row q then reads phase (ky != 1, kx != 1) at row q - (ky == 0) halo_w - (kx == 0): a constant row shift per tap,
   * as in the stride-1 form.  0 / 1 = stride 1. */
  int32_t conv_stride;
} hoigen_gemm_params;

HOIGEN_API int hoigen_gemm_bf16(const hoigen_gemm_params* p, hoigen_stream_t stream);
/* Test-only SIMT cross-check of the same contract (never used by the product path). */
HOIGEN_API int hoigen_debug_gemm_simt(const hoigen_gemm_params* p, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Encoder row kernels (memory bound).  tokens = rows b*197+t of (B*197, 768) fp32 residual stream.
 * ---------------------------------------------------------------------------------------------- */
/* images (B,3,224,224) fp32 -> patches (B*196, 768) bf16, column = c*256+ky*16+kx: im2col of conv1, C:491 */
HOIGEN_API int hoigen_patchify_bf16(const float* images, void* patches_bf16, int32_t batch, hoigen_stream_t stream);
/* x[b,t] = ln_pre((t==0 ? class_embedding : patch_emb[b,t-1]) + positional_embedding[t])   C:494-496 */
HOIGEN_API int hoigen_embed_lnpre(const float* patch_emb, const float* class_embedding, const float* positional_embedding,
                                  const float* gamma, const float* beta, float* x_f32, void* x_bf16, int32_t batch,
                                  hoigen_stream_t stream);
/* LayerNorm over 768 columns, fp32 statistics, eps 1e-5 (C:409-415) -> bf16 and/or fp32 */
HOIGEN_API int hoigen_layernorm768(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                                   int32_t rows, hoigen_stream_t stream);
/* x += delta (+ delta2) (+ col_bias, a (768) row vector) in place on the fp32 residual stream, then LayerNorm(x) ->
 * out_bf16, and optionally a bf16 copy of the updated stream (x_bf16): the residual adds of C:456-458 deferred from the
 * producing GEMMs into the LayerNorm pass that streams the same rows anyway.  delta2, col_bias, x_bf16 may be NULL. */
HOIGEN_API int hoigen_add_layernorm768(float* x, const void* delta_bf16, const void* delta2_bf16, const float* col_bias,
                                       const float* gamma, const float* beta, void* out_bf16, void* x_bf16, int32_t rows,
                                       hoigen_stream_t stream);
/* The same pass with the LayerNorm itself left to the consuming GEMM (hoigen_gemm_params.ln_stats): x += delta (+ delta2)
 * (+ col_bias row) in fp32, x_bf16 = bf16 copy of the updated rows (the GEMM's A operand), stats (rows,2) = {mean, rstd}
 * of each updated row (fp32, eps 1e-5: C:409-415). */
HOIGEN_API int hoigen_add_rowstats768(float* x, const void* delta_bf16, const void* delta2_bf16, const float* col_bias,
                                      void* x_bf16, float* stats, int32_t rows, hoigen_stream_t stream);
/* Adapter cross-attention K/V of the prior tokens for all layers: kv[l][tok][0:64]=K, [64:128]=V  (C:63-66).
 * in_proj_w (layers,192,64) rows [q;k;v], in_proj_b (layers,192); prior (tokens,64). */
HOIGEN_API int hoigen_adapter_kv(const float* prior, const float* in_proj_w, const float* in_proj_b, float* kv,
                                 int32_t tokens, int32_t layers, hoigen_stream_t stream);

typedef struct {
  const void* wd;          /* bf16 (64,768)  adaptermlp.down_proj.weight */
  const float* down_b;     /* (64)           adaptermlp.down_proj.bias */
  const void* wq;          /* bf16 (64,64)   q rows of multihead_attn.in_proj_weight */
  const void* wo;          /* bf16 (64,64)   multihead_attn.out_proj.weight */
  const void* w1;          /* bf16 (128,64)  linear1.weight */
  const void* w2;          /* bf16 (64,128)  linear2.weight */
  const float* in_proj_b;  /* (192) multihead_attn.in_proj_bias (q part used here) */
  const float* out_proj_b; /* (64) */
  const float* linear1_b;  /* (128) */
  const float* linear2_b;  /* (64) */
  const float* norm2_w; const float* norm2_b; const float* norm3_w; const float* norm3_b; /* (64) */
  const void* wup;         /* bf16 (768, 64)  scale[:,None] * up_proj.weight (the adapter's output scale folded in) */
} hoigen_adapter_weights;
/* One whole adapter block on the tensor cores — Adapter.forward C:183-203 with forward_post C:51-72.  The adapter input is xb + delta_c: xb = bf16 copy of the stream (B*197,768), delta_c = the
 * still-pending bf16 residual of the previous block's MLP output (C:458) or NULL; the sum is never materialised
 * ((xb + delta_c) Wd^T = xb Wd^T + delta_c Wd^T inside one TMEM accumulation).
 *   d = relu(down_proj(xb + delta_c)) ; t = LN2(d + MHA_2h(d, prior, mask)) ; o = LN3(t + FFN(t))   (B*197,64)
 *   delta_out = o wup^T -> bf16 (B*197,768) = scale * (up_proj(o) - up_proj.bias): the residual of C:456 less its
 *   constant row scale * up_proj.bias, which the caller adds with hoigen_add_layernorm768(col_bias)
 * bottleneck_bf16 (B*197,64) receives o when not NULL (tests).
 * kv_layer (B*n_max,128) from hoigen_adapter_kv; mask (B,n_max) uint8, 1 = padding. n_max <= 32. */
HOIGEN_API int hoigen_adapter_block(const void* xb, const void* delta_c, const float* kv_layer, const uint8_t* mask,
                                    const hoigen_adapter_weights* w, void* bottleneck_bf16, void* delta_out_bf16,
                                    int32_t batch, int32_t n_max, hoigen_stream_t stream);
/* Diagnostics only: later hoigen_adapter_block launches write CTA 0's clock64() phase stamps (16 x int64) to trace. */
HOIGEN_API int hoigen_debug_adapter_trace(int64_t* trace);
/* 12-head attention over 197 tokens (tcgen05): qkv bf16 (B*197, 2304) = [q|k|v] -> out bf16 (B*197, 768). C:443-445 */
HOIGEN_API int hoigen_attention(const void* qkv_bf16, void* out_bf16, int32_t batch, hoigen_stream_t stream);
/* Diagnostics only: the same launch; CTA 0 writes clock64() stamps [16 items][8 phases] of its softmax loop to trace. */
HOIGEN_API int hoigen_debug_attention_trace(const void* qkv, void* out_bf16, int32_t batch, int64_t* trace,
                                            hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Whole visual encoder = VisionTransformer.forward(x, prior), C:489-506.  One call per batch.
 * Weights are the reference parameters re-packed once at build time (bf16 GEMM operands, fp32 vectors),
 * stacked over the 12 layers; see hoigen_b200/encoder.py::pack_encoder_weights for the exact mapping.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* conv_w;                 /* bf16 (768, 768)      conv1.weight.view(768,-1) */
  const float* class_embedding;       /* (768) */
  const float* positional_embedding;  /* (197,768) */
  const float* ln_pre_w; const float* ln_pre_b; const float* ln_post_w; const float* ln_post_b; /* (768) */
  const void* proj_t;                 /* bf16 (512, 768)      proj^T */
  const float* ln1_w; const float* ln1_b; const float* ln2_w; const float* ln2_b;  /* (12,768) */
  const void* qkv_w; const float* qkv_b;     /* bf16 (12,2304,768), (12,2304)   attn.in_proj_* */
  const void* out_w; const float* out_b;     /* bf16 (12,768,768),  (12,768)    attn.out_proj.* */
  const void* fc_w; const float* fc_b;       /* bf16 (12,3072,768), (12,3072)   mlp.c_fc.* */
  const void* proj_w; const float* proj_b;   /* bf16 (12,768,3072), (12,768)    mlp.c_proj.* */
  const void* ad_down_w; const float* ad_down_b;  /* bf16 (12,64,768), (12,64)  adaptermlp.down_proj.* */
  const void* ad_up_w; const float* ad_up_b;      /* bf16 (12,768,64), (12,768) adaptermlp.scale * up_proj.{weight,bias} */
  const float* ad_in_proj_w; const float* ad_in_proj_b;    /* (12,192,64), (12,192)  mhsa_layers.0.multihead_attn */
  const void* ad_wq; const void* ad_wo;                    /* bf16 (12,64,64) each: q rows of in_proj, out_proj */
  const void* ad_w1; const void* ad_w2;                    /* bf16 (12,128,64), (12,64,128): linear1, linear2 */
  const float* ad_out_proj_b;                              /* (12,64) */
  const float* ad_linear1_b;                               /* (12,128) */
  const float* ad_linear2_b;                               /* (12,64) */
  const float* ad_norm2_w; const float* ad_norm2_b; const float* ad_norm3_w; const float* ad_norm3_b; /* (12,64) */
  /* LayerNorm folded into the QKV / c_fc GEMMs (north_star item 1; C:457-458).  All six NULL = unfused LayerNorm passes. */
  const void* qkv_wf; const float* qkv_colsum; const float* qkv_bf;   /* bf16 (12,2304,768) = in_proj_weight . diag(ln_1.weight); (12,2304) row sums of it; (12,2304) in_proj_bias + in_proj_weight ln_1.bias */
  const void* fc_wf; const float* fc_colsum; const float* fc_bf;      /* the same for mlp.c_fc and ln_2: (12,3072,768), (12,3072), (12,3072) */
} hoigen_encoder_weights;

typedef struct {            /* caller-owned workspace, M = B*197 */
  void* patches;            /* bf16 (B*196, 768) */
  float* patch_emb;         /* f32  (B*196, 768) */
  float* x;                 /* f32  (M, 768)   residual stream */
  void* xb;                 /* bf16 (M, 768)   bf16 copy of x as of the last LayerNorm pass (adapter down-proj operand) */
  void* h;                  /* bf16 (M, 768)   LayerNorm output */
  void* qkv;                /* bf16 (M, 2304) */
  void* attn;               /* bf16 (M, 768) */
  void* mlp;                /* bf16 (M, 3072) */
  void* delta;              /* bf16 (M, 768)   adapter up-proj / attention out-proj output awaiting its residual add */
  void* delta2;             /* bf16 (M, 768)   MLP c_proj output awaiting its residual add (applied by the next adapter block) */
  float* adapter_kv;        /* f32  (12, B*n_max, 128) */
  float* row_stats;         /* f32  (M, 2)     {mean, rstd} of the stream rows (LayerNorm folded into the next GEMM) */
  float* tokens_out;        /* f32  (M, 512)   OUTPUT: ln_post(x) @ proj for all tokens; row b*197 = feat_global[b],
                                               rows b*197+1.. = feat_local[b] token-major (C:503-506) */
} hoigen_encoder_buffers;

/* prior (B,n_max,64) fp32 and mask (B,n_max) uint8 (1 = padding) as produced by hoigen_prior_tokens. */
HOIGEN_API int hoigen_encoder_forward(const hoigen_encoder_weights* w, const hoigen_encoder_buffers* buf,
                                      const float* images, const float* prior, const uint8_t* mask, int32_t batch,
                                      int32_t n_max, int32_t num_layers, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * HOI head.  Boxes/scores/labels of all images are concatenated; box_off (B+1) and pair_off (B+1) are CSR
 * offsets (pairs per image K_b = n_h*(n_b-1), or 0 when n_h == 0 or n_b <= 1; humans lead: U:988-1016).
 * ---------------------------------------------------------------------------------------------- */
/* get_prior U:1445-1495 (prior_type 'cbe', prior_method 0) incl. the 517->128->128->64 MLP (U:40-52, U:520).
 * w*t are the Linear weights TRANSPOSED to (in,out). Outputs prior (B,n_max,64), mask (B,n_max) 1=padding. */
HOIGEN_API int hoigen_prior_tokens(const float* boxes, const float* scores, const int64_t* labels, const int32_t* box_off,
                                   const float* object_embedding, const float* w0t, const float* b0, const float* w1t,
                                   const float* b1, const float* w2t, const float* b2, float img_w, float img_h,
                                   int32_t batch, int32_t n_max, int32_t num_objects /* rows of object_embedding */,
                                   float* prior, uint8_t* mask, hoigen_stream_t stream);
/* compute_roi_embeddings geometry U:981-1057: RoIAlign(7x7, sampling_ratio=-1, aligned=True)+mean of every single
 * box and every union box on the 14x14 token grid (torchvision.ops.roi_align call sites U:1028-1029), then
 * f_H = single[x]/|.|, f_O = single[y]/|.|, f_U = union/|.|  -> pair_feat [3][Ktot][512] (H,O,U) bf16 (+fp32).
 * RoIAlign + mean runs as one fp32-accurate tensor-core product per image (3 x bf16-split operands, roi_tc.cu; <= 32 boxes
 * per image); roi_weights is scratch for the fp32 SIMT form (HOIGEN_ROI_SIMT=1) and otherwise untouched. */
HOIGEN_API int hoigen_roi_pair_features(const float* tokens, const float* boxes, const int32_t* box_off,
                                        const int32_t* pair_off, int32_t batch, int32_t ntot, int32_t ktot,
                                        float spatial_scale, float* roi_weights /* workspace (ntot+ktot, 32) */,
                                        float* single_feat, float* union_feat, void* pair_feat_bf16,
                                        float* pair_feat_f32, hoigen_stream_t stream);
/* n 32-bit words dst <- src by a kernel; either side may be mapped pinned host memory (cudaHostAlloc under UVA).  Used
 * for the path's two tiny transfers (CSR layout host->device, per-image triplet offsets device->host, U:1421-1425's
 * list lengths) so that they never queue behind bulk copies on the copy engines. */
HOIGEN_API int hoigen_copy_words(void* dst, const void* src, int32_t n_words, hoigen_stream_t stream);
/* n 32-bit words from HOST memory (read before the call returns) -> device memory, carried in the kernel parameters:
 * no DMA and no PCIe read, so the CSR layout of a forward cannot queue behind a bulk upload that shares the link. */
HOIGEN_API int hoigen_set_words(void* dst, const void* host_values, int32_t n_words, hoigen_stream_t stream);
HOIGEN_API int hoigen_rows_to_bf16(const float* in, int64_t ld_in, int32_t rows, int32_t cols, int32_t normalize,
                                   void* out_bf16, hoigen_stream_t stream);
HOIGEN_API int hoigen_broadcast_image_logits(const float* img_logits, const int32_t* pair_off, int32_t batch,
                                             int32_t ktot, int32_t num_classes, int32_t ld_logits, float* logits,
                                             hoigen_stream_t stream);

typedef struct {                   /* scoring weights, packed once at build time (U:1156-1163, 1112-1115, 1133-1138) */
  int32_t num_classes;             /* C */
  int32_t cache_rows;              /* N (multiple of 8) */
  const void* cache_keys[3];       /* bf16 (N,512)   gen_adapter_{H,O,U}_weight            order: H, O, U */
  const float* bias_term[3];       /* f32 (C)        gen_adapter_X_bias @ gen_label_X */
  const void* label_t[3];          /* bf16 (C,N)     gen_label_X^T (multi-hot, exact in bf16) */
  const float* colscale[3];        /* f32 (C)        gen_logit_scale_X / sample_lens_X */
  const void* global_keys;         /* bf16 (N,512)   global_cache^T */
  const float* global_bias_term;   /* f32 (C)        global_cache_bias @ gen_label_U */
  const float* colscale_global;    /* f32 (C)        clip_cache_logit / global_sample_len */
  const void* dino_keys;           /* bf16 (N,2048)  dino_cache^T, or NULL */
  const float* dino_bias_term;     /* f32 (C) */
  const float* colscale_dino;      /* f32 (C)        dino_cache_logit / dino_sample_len */
  const void* text_w;              /* bf16 (C,512)   adapter_union_weight */
  const float* colscale_text;      /* f32 (C)        logit_scale_text broadcast */
  /* cache affinity: 0 = linear phi = f W^T + b (THE REFERENCE, U:1156-1158; b enters through the bias_term carriers),
   * 1 = exp(beta (f W^T + b)) (textbook Tip-Adapter, north_star item 3; not a parity mode).  The fields below are read
   * only for affinity 1. */
  int32_t affinity;
  float beta;
  const float* cache_bias[3];      /* f32 (N)        gen_adapter_{H,O,U}_bias, zero-padded */
  const float* global_bias;        /* f32 (N)        global_cache_bias */
  const float* dino_bias;          /* f32 (N)        dino_cache_bias */
} hoigen_score_weights;

typedef struct {                   /* caller-owned workspace */
  const void* pair_feat_bf16;      /* bf16 [3][Ktot][512] from hoigen_roi_pair_features */
  void* phi;                       /* bf16 (Ktot, N)   two-GEMM form only (may be NULL when cache_parts is given) */
  void* phi_img;                   /* bf16 (B, N) */
  void* g_bf16;                    /* bf16 (B, 512) */
  void* d_bf16;                    /* bf16 (B, 2048) */
  float* img_logits;               /* f32 (B, C) */
  float* logits;                   /* f32 (Ktot, ld_logits)   OUTPUT: columns [0, C) of every row */
  int64_t ld_logits;               /* floats between logits rows, >= C (0 = C).  A multiple of 4 keeps the accumulating
                                      GEMM epilogues (6 terms summed into this buffer) on their float4 path */
  float* cache_parts;              /* f32, hoigen_cache_fused_workspace_bytes(Ktot, C) bytes: when given and C <= 128 the three
                                      cache branches run as ONE fused GEMM-f-GEMM kernel (hoigen_score_cache_fused) and `phi`
                                      is not touched; NULL = two GEMMs per branch through `phi` */
} hoigen_score_buffers;

/* logits = sum_X scale_X * ((f_X W_X^T + b_X) Y_X)/s_X + scale_T f_U T^T + per-image global/DINO cache terms.
 * tokens = encoder output (B*197,512); dino_feats (B,2048) L2-normalised fp32 or NULL. */
HOIGEN_API int hoigen_score_pairs(const hoigen_score_weights* w, const hoigen_score_buffers* buf, const float* tokens,
                                  const float* dino_feats, const int32_t* pair_off, int32_t batch, int32_t ktot,
                                  hoigen_stream_t stream);

/* The three cache branches of hoigen_score_pairs as ONE fused GEMM - f - GEMM tcgen05 kernel (north_star item 3; U:1156-1163):
 * per 64 cache rows S = f W^T lives in TMEM, is converted to bf16 in registers and consumed from TMEM by the second MMA
 * against the label tile — the (Ktot x N) affinity matrix never exists in memory.  affinity 0 = the reference's linear
 * phi = f W^T + b (bias carried exactly in fp32 by the combine pass), 1 = exp(beta (f W^T + b)) (textbook Tip-Adapter;
 * cache_bias[3] = gen_adapter_{H,O,U}_bias, (N) fp32, zero-padded).  Writes logits[i][c] = img_logits[image(i)][c] (or 0 if
 * NULL) + sum_X colscale_X[c] ((bias_term_X[c]) + L_X[i][c]) in a fixed summation order.  num_classes <= 128.
 * parts: fp32 workspace of hoigen_cache_fused_workspace_bytes(ktot, num_classes) bytes. */
HOIGEN_API int64_t hoigen_cache_fused_workspace_bytes(int32_t ktot, int32_t num_classes);
HOIGEN_API int hoigen_score_cache_fused(const hoigen_score_weights* w, const void* pair_feat_bf16,
                                        const float* const* cache_bias, const float* img_logits, const int32_t* pair_off,
                                        int32_t batch, int32_t ktot, int32_t affinity, float beta, float* parts,
                                        float* logits, int32_t ld_logits, hoigen_stream_t stream);

/* fp32 rows -> three bf16 planes (hi, mid, lo; hi + mid + lo == x to 2^-24) laid side by side along K:
 * pattern 6 = [hi|hi|hi|mid|mid|lo] (multiplied against a weight packed as [hi|mid|lo|hi|mid|hi]: every cross term down
 * to 2^-16), pattern 3 = [hi|mid|lo] (against an operand exact in bf16, tiled 3x).  out: bf16 (rows, pattern*cols).
 * normalize != 0: the row is divided by its L2 norm first (U:960, U:1618). */
HOIGEN_API int hoigen_rows_split3(const float* in, int64_t ld_in, int32_t rows, int32_t cols, int32_t normalize,
                                  int32_t pattern, void* out, hoigen_stream_t stream);

/* fp32-accurate scoring (north_star: "fp32 <= 1e-4" for RoI + scoring given identical features):
 * hoigen_score_pairs' chain with every GEMM evaluated through the 3 x bf16 split on the tcgen05 kernel and phi kept in
 * fp32.  Weights as below are packed once from the fp32 parameters. */
typedef struct {
  int32_t num_classes;             /* C */
  int32_t cache_rows;              /* N (multiple of 8) */
  const void* cache_keys6[3];      /* bf16 (N, 6*512)   [hi|mid|lo|hi|mid|hi] planes of gen_adapter_{H,O,U}_weight */
  const float* bias_term[3];       /* f32 (C) */
  const void* label3_t[3];         /* bf16 (C, 3N)      gen_label_X^T tiled three times along K */
  const float* colscale[3];        /* f32 (C) */
  const void* global_keys6;        /* bf16 (N, 6*512) */
  const float* global_bias_term;   /* f32 (C) */
  const float* colscale_global;    /* f32 (C) */
  const void* dino_keys6;          /* bf16 (N, 6*2048), or NULL */
  const float* dino_bias_term;     /* f32 (C) */
  const float* colscale_dino;      /* f32 (C) */
  const void* text_w6;             /* bf16 (C, 6*512) */
  const float* colscale_text;      /* f32 (C) */
} hoigen_score_weights_fp32;

typedef struct {                   /* caller-owned workspace */
  void* feat6;                     /* bf16 [3][Ktot][6*512] */
  float* phi;                      /* f32  (Ktot, N) */
  void* phi3;                      /* bf16 (Ktot, 3N) */
  float* phi_img;                  /* f32  (B, N) */
  void* phi_img3;                  /* bf16 (B, 3N) */
  void* g6;                        /* bf16 (B, 6*512) */
  void* d6;                        /* bf16 (B, 6*2048) */
  float* img_logits;               /* f32 (B, C) */
  float* logits;                   /* f32 (Ktot, ld_logits)   OUTPUT */
  int64_t ld_logits;
} hoigen_score_buffers_fp32;

/* pair_feat_f32 = the fp32 [3][Ktot][512] output of hoigen_roi_pair_features. */
HOIGEN_API int hoigen_score_pairs_fp32(const hoigen_score_weights_fp32* w, const hoigen_score_buffers_fp32* buf,
                                       const float* tokens, const float* dino_feats, const float* pair_feat_f32,
                                       const int32_t* pair_off, int32_t batch, int32_t ktot, hoigen_stream_t stream);

/* Opt-in "folded cache" form of hoigen_score_pairs.  The reference's cache affinity is linear (no exp: U:1156-1158), so
 * ((f W^T + b) Y) s / L = f (W^T Y s / L) + (b Y) s / L : the host contracts every cache with its label matrix ONCE and
 * the six logit terms of U:1185-1186 become one (C x 1536) matrix for the pair features [H | O | U], one (C x 512) for
 * the global feature, one (C x 2048) for the DINO feature and a constant row.  Same outputs (tested), no 4096-wide
 * intermediate.  NOT the default of the Python surface: bench.py's headline runs the unfolded path above. */
typedef struct {
  int32_t num_classes;             /* C */
  const void* pair_w;              /* bf16 (C, 1536)  [E_H | E_O | E_U + s_T W_T],  E_X = s_X/L_X . (Y_X^T W_X) */
  const void* global_w;            /* bf16 (C, 512)   or NULL */
  const void* dino_w;              /* bf16 (C, 2048)  or NULL */
  const float* bias_total;         /* f32 (C)         sum over present branches of s_X/L_X . (b_X Y_X) */
} hoigen_folded_weights;
HOIGEN_API int hoigen_score_pairs_folded(const hoigen_folded_weights* w, const hoigen_score_buffers* buf, const float* tokens,
                                         const float* dino_feats, const int32_t* pair_off, int32_t batch, int32_t ktot,
                                         hoigen_stream_t stream);

/* compute_prior_scores U:806-833 + postprocessing U:1408-1427, whole batch, reference (row-major) order.
 * table_bits (80, table_words) uint32 bitmask of object_class_to_target_class. Outputs are packed over images:
 * scores/labels/objects [Mtot]; pairing holds, per image b, a contiguous [2][M_b] block at 2*img_off[b];
 * img_off (B+1) int32 triplet offsets. capacity = allocated Mtot; entries beyond it are dropped (check img_off[B]). */
HOIGEN_API int hoigen_emit_triplets(const float* logits, int32_t num_classes, int32_t ld_logits, const float* scores,
                                    const int64_t* labels,
                                    const int32_t* box_off, const int32_t* pair_off, int32_t batch, int32_t ktot,
                                    const uint32_t* table_bits, int32_t table_words,
                                    int32_t table_rows /* rows of table_bits; labels outside [0,rows) emit nothing */,
                                    float hyper_lambda, int32_t* work_counts, int32_t* work_offsets, float* work_pr, int64_t capacity,
                                    float* out_scores, int64_t* out_labels, int64_t* out_objects, int64_t* out_pairing,
                                    int32_t* img_off, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: compact wire record of one batch's detections for the path's one collective, the gather of
 * per-image detections (no reference counterpart: single-GPU eval, main_tip_finetune.py:383-388).  9 bytes per triplet
 * (f32 score, u16 label, u8 object class, u8 human index, u8 object index) instead of the 36 the reference's dtypes take
 * (U:1421-1425); the record is built and parsed on the device, sizes travel in its header:
 *   int32 header[4 + 2*(max_images+1)] = magic, nimg, M (-1: did not fit), nbox, triplet_off[0..nimg], box_off[0..nimg]
 *   f32 boxes[nbox*4] (16-byte aligned) | f32 scores[M] | u16 labels[M] | u8 objects[M] | u8 human_idx[M] | u8 object_idx[M]
 * ---------------------------------------------------------------------------------------------- */
/* bytes a record needs for the given bounds (multiple of 16), or -1 */
HOIGEN_API int64_t hoigen_wire_record_bytes(int32_t max_images, int64_t max_triplets, int64_t max_boxes);
/* inputs = the packed outputs of hoigen_emit_triplets (+ boxes, box_off); record: cap_bytes bytes, 16-byte aligned */
HOIGEN_API int hoigen_pack_wire(const float* scores, const int64_t* labels, const int64_t* objects, const int64_t* pairing,
                                const float* boxes, const int32_t* img_off, const int32_t* box_off, int32_t nimg,
                                int32_t max_images, int64_t cap_bytes, void* record, hoigen_stream_t stream);
/* n_records records `record_stride` bytes apart -> the reference's dtypes.  bases (n_records, 2) int64 on the device:
 * {first triplet, first box} of each record in the output arrays, or -1 to skip it; pairing gets per-image [2][M_b]
 * blocks at 2*(base + triplet_off[b]). */
HOIGEN_API int hoigen_unpack_wire(const void* records, int32_t n_records, int64_t record_stride, int32_t max_images,
                                  const int64_t* bases, float* scores, int64_t* labels, int64_t* objects,
                                  int64_t* pairing, float* boxes, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * f1 (the path's immediate consumer): batched detection <-> ground-truth association of the eval caller,
 * CustomisedDLE.test_hico utils_tip_cache_and_union_finetune.py:375-407 + BoxPairAssociation
 * pocket/pocket/utils/association.py:51-125, for all images of a batch in one launch.
 *   interactions[d] = conversion[objects[d]][verbs[d]] (-1 where the table has none; = verbs[d] if conversion is NULL)
 *   labels[d] = 1 iff d is the highest-scoring detection (first on ties) among those whose best-overlapping
 *               ground-truth pair of the same HOI id (first on ties) is g, with pair IoU
 *               min(IoU(human boxes), IoU(object boxes)) > min_iou; else 0.
 * Detections are the packed outputs of hoigen_emit_triplets (boxes + box_off, pairing blocks, objects, verbs = labels,
 * scores, trip_off); ground truth is CSR over images (gt_off), boxes xyxy in the detections' frame (UPT.recover_boxes
 * already applied).  IoU arithmetic is bit-identical to torchvision.ops.box_iou in fp32.  Scores must be >= 0. */
HOIGEN_API int hoigen_associate_pairs(const float* boxes, const int32_t* box_off, const int64_t* pairing,
                                      const int64_t* objects, const int64_t* verbs, const float* scores,
                                      const int32_t* trip_off, const int32_t* conversion /* (80, num_verbs) or NULL */,
                                      int32_t num_verbs, const float* gt_boxes_h, const float* gt_boxes_o,
                                      const int64_t* gt_hoi, const int32_t* gt_off, int32_t max_gt_per_image,
                                      int32_t batch, float min_iou, int64_t* interactions, float* labels,
                                      hoigen_stream_t stream);

/* f2: per-class 11-point interpolated AP of a sweep — DetectionAPMeter (algorithm '11P', precision 64),
 * pocket/pocket/utils/meters.py:561-583 + 255-270, all classes in one launch.  labels_sorted: the 0/1 labels of every
 * collected detection ordered by (class ascending, score descending); class_off (C+1) int64 CSR offsets into it;
 * num_gt (C) fp64, < 0 = "not given" (recall over the collected true positives); thresholds = the 11 fp64 values of
 * torch.linspace(0, 1, 11).  Outputs ap (C), max_rec (C) in fp64; an empty class gives 0, 0 as the reference does. */
HOIGEN_API int hoigen_ap_11point(const float* labels_sorted, const int64_t* class_off, const double* num_gt,
                                 const double* thresholds, int32_t num_classes, double* ap, double* max_rec,
                                 hoigen_stream_t stream);

/* f3 (the non-network half of the proposal stage): UPT.prepare_region_proposals U:1361-1406 for a whole batch —
 * batched_nms(boxes, scores, labels, nms_iou) with torchvision's coordinate trick and NMS arithmetic, score threshold,
 * min / max-instance rule, humans first.  Inputs (B, Q) scores, (B, Q) int64 labels, (B, Q, 4) boxes, Q <= 256.
 * Outputs are padded per image to 2*max_instances slots: the first counts[2b] are humans, the next counts[2b+1] objects,
 * each group in descending score order; out_boxes (B, 2*max_instances, 4), out_scores, out_labels; counts (B, 2). */
HOIGEN_API int hoigen_prepare_proposals(const float* scores, const int64_t* labels, const float* boxes, int32_t batch,
                                        int32_t num_queries, int64_t human_idx, float box_score_thresh,
                                        int32_t min_instances, int32_t max_instances, float nms_iou, float* out_boxes,
                                        float* out_scores, int64_t* out_labels, int32_t* counts, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a8: the ResNet-50 branch of the forward -- `dino_model(images_clip)` then `/ norm` (U:1616-1618; dino_model is a
 * torchvision resnet50 with fc = Identity in eval mode, main_tip_finetune.py:393,404-405).  BatchNorms are folded into the
 * convolutions at pack time; activations are NHWC bf16 with a one-pixel zero halo (rows = pixels of (B, H+2, W+2));
 * 1x1 convolutions and the 3x3 / stride-1 ones run on hoigen_gemm_bf16 (conv_taps / halo_* / res_bf16 of
 * hoigen_gemm_params: bias + identity + ReLU + halo zeroing in the epilogue); these are the remaining row kernels and the
 * runner that executes a packed list of such steps in order on one stream.
 * ---------------------------------------------------------------------------------------------- */
/* images (B,3,224,224) fp32 -> rows (B*112*112, 160) bf16: im2col of conv1 (7x7, stride 2, pad 3), column = (ky*7+kx)*3+c,
 * columns 147..159 zero. */
HOIGEN_API int hoigen_stem_im2col(const float* images, void* rows_bf16, int32_t batch, hoigen_stream_t stream);
/* The same im2col for (B,3,h,w) images of any size -> rows (B*ceil(h/2)*ceil(w/2), 160): the stem of DETR's ResNet-50 backbone
 * (detr/models/backbone.py:83-91, U:1594) is the same convolution on larger, padded images. */
HOIGEN_API int hoigen_stem_im2col_hw(const float* images, void* rows_bf16, int32_t batch, int32_t h, int32_t w,
                                     hoigen_stream_t stream);
/* The same convolution fused with its BatchNorm-folded bias and ReLU as ONE tensor-core kernel without the im2col matrix:
 * images (B,3,224,224) fp32 -> out (B*112*112, 64) bf16 NHWC rows.  w: bf16 (64, 192), column = ky*24 + kx*3 + c (each
 * ky run of 21 taps padded to 24, 168..191 zero); bias (64) fp32. */
HOIGEN_API int hoigen_stem_conv(const float* images, const void* w_bf16, const float* bias, void* out_bf16, int32_t batch,
                                hoigen_stream_t stream);
/* The same kernel for (B,3,h,w) images of any size -> out (B*ceil(h/2)*ceil(w/2), 64)  (DETR's backbone stem, U:1594). */
HOIGEN_API int hoigen_stem_conv_hw(const float* images, const void* w_bf16, const float* bias, void* out_bf16, int32_t batch,
                                   int32_t h, int32_t w, hoigen_stream_t stream);
/* MaxPool2d(3, stride 2, padding 1): in (B,h,w,c) bf16 without halo -> out (B, ceil(h/2)+2, ceil(w/2)+2, c) with the zero halo */
HOIGEN_API int hoigen_maxpool3x3s2_halo(const void* in_bf16, void* out_bf16, int32_t batch, int32_t h, int32_t w, int32_t c,
                                        hoigen_stream_t stream);
/* A operand of the stride-2 convolutions: in (B, h+2, w+2, c) -> rows (B*(ceil(h/2)+2)*(ceil(w/2)+2), taps*c); taps = 9: 3x3 / pad 1,
 * column = (ky*3+kx)*c + channel; taps = 1: the 1x1 shortcut.  Rows of the output ring are zeros.
 * taps = 4: the four-phase split [4][B*(ceil(h/2)+2)*(ceil(w/2)+2)][c] that hoigen_gemm_params.conv_stride = 2 consumes. */
HOIGEN_API int hoigen_conv_gather_s2(const void* in_bf16, void* rows_bf16, int32_t batch, int32_t h, int32_t w, int32_t c,
                                     int32_t taps, hoigen_stream_t stream);
/* AdaptiveAvgPool2d(1) over the interior of (B, h+2, w+2, c), then x / ||x||_2 (U:1618) -> (B, c) fp32; c % 256 == 0, <= 2048 */
HOIGEN_API int hoigen_avgpool_l2norm(const void* in_bf16, float* out, int32_t batch, int32_t h, int32_t w, int32_t c,
                                     hoigen_stream_t stream);

typedef enum {
  HOIGEN_CONV_OP_GEMM = 0,            /* gemm                                   */
  HOIGEN_CONV_OP_STEM_IM2COL = 1,     /* in = images, out = rows, batch, (h, w: image size, 0 = 224) */
  HOIGEN_CONV_OP_MAXPOOL = 2,         /* in, out, batch, h, w, c                 */
  HOIGEN_CONV_OP_GATHER_S2 = 3,       /* in, out, batch, h, w, c, taps           */
  HOIGEN_CONV_OP_AVGPOOL_L2NORM = 4,  /* in, out (fp32), batch, h, w, c          */
  HOIGEN_CONV_OP_STEM_CONV = 5        /* in = images, out, batch, gemm.w, gemm.bias, (h, w: image size, 0 = 224) */
} hoigen_conv_op_kind;
typedef struct {
  int32_t kind;
  const void* in;
  void* out;
  int32_t batch, h, w, c, taps;
  hoigen_gemm_params gemm;
} hoigen_conv_op;
/* Launches ops[0..n_ops) in order on `stream`; stops at the first error. */
HOIGEN_API int hoigen_conv_plan_run(const hoigen_conv_op* ops, int32_t n_ops, hoigen_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * f3, second half: the DETR transformer behind U:1596 (detr/models/transformer.py; d_model 256, 8 heads of 32, post-norm).
 * Every nn.Linear is hoigen_gemm_bf16; these are the two row kernels between them.
 * ---------------------------------------------------------------------------------------------- */
/* x (rows,256) fp32 in place:  x += delta (bf16, may be NULL) ; x = LayerNorm(x) * gamma + beta (eps 1e-5; gamma = beta = NULL:
 * no normalisation) -- `src = norm(src + dropout(src2))` of forward_post (transformer.py:143-149, 200-208).  Side outputs for the
 * next products: x_bf16 = bf16(x) and xpos_bf16 = bf16(x + pos[row % pos_rows]) = `with_pos_embed` (transformer.py:124-125); either may be NULL. */
HOIGEN_API int hoigen_add_layernorm256(float* x, const void* delta_bf16, const float* gamma, const float* beta, const float* pos,
                                       int32_t pos_rows, void* x_bf16, void* xpos_bf16, int32_t rows, hoigen_stream_t stream);
/* nn.MultiheadAttention's core for head_dim 32: out[b, i, h*32:(h+1)*32] = softmax_j(q_i . k_j * scale  [-inf where
 * key_mask[b, j] != 0]) v_j, per image b and head h.  q (batch*lq, ldq), k / v (batch*lk, ldk / ldv), out (batch*lq, ldo): bf16 rows,
 * head h in columns [32 h, 32 h + 32); pitches in elements (multiples of 8).  key_mask (batch, lk) uint8 or NULL
 * (key_padding_mask, transformer.py:139,193).  Online-softmax tcgen05 kernel (S = Q K^T in TMEM, P V with V read MN-major from its
 * key rows); HOIGEN_ATT32_SIMT=1 selects the fp32 SIMT form. */
HOIGEN_API int hoigen_attention_heads32(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                        void* out, int32_t ldo, const uint8_t* key_mask, int32_t batch, int32_t lq, int32_t lk,
                                        int32_t heads, float scale, hoigen_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HOIGEN_B200_H_ */
