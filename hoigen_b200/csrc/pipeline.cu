// Host-side drivers that chain the kernels of one stage into ONE C-ABI call per batch (the Python caller would
// otherwise pay ~120 ctypes round trips per encoder forward):
//   hoigen_encoder_forward : VisionTransformer.forward(x, prior)      CLIP_models_adapter_prior2.py:489-506
//   hoigen_score_pairs     : cache-model + text logits                upt_..._distill3.py:1111-1186
#include <stdlib.h>

#include "common.h"

namespace hoigen {

static int gemm(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias, int act,
                const float* colscale, const float* residual, int ld_res, float* out_f32, int ld_f32, void* out_bf16,
                int ld_bf16, cudaStream_t s) {
  hoigen_gemm_params p = {};
  p.a = a; p.w = w; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw;
  p.bias = bias; p.colscale = colscale; p.act = act;
  p.residual = residual; p.ld_res = ld_res;
  p.out_f32 = out_f32; p.ld_f32 = ld_f32;
  p.out_bf16 = out_bf16; p.ld_bf16 = ld_bf16;
  p.split_k = 0;
  p.block_n = 0;
  p.act_param = 0.f;
  p.ln_stats = nullptr;
  p.ln_colsum = nullptr;
  return hoigen_gemm_bf16(&p, s);
}

// bf16-output GEMM with the preceding LayerNorm folded into its epilogue (see hoigen_gemm_params.ln_stats)
static int gemm_ln(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias, int act,
                   const float* ln_stats, const float* ln_colsum, void* out_bf16, int ld_bf16, cudaStream_t s) {
  hoigen_gemm_params p = {};
  p.a = a; p.w = w; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw;
  p.bias = bias; p.colscale = nullptr; p.act = act;
  p.residual = nullptr; p.ld_res = 0;
  p.out_f32 = nullptr; p.ld_f32 = 0;
  p.out_bf16 = out_bf16; p.ld_bf16 = ld_bf16;
  p.split_k = 0;
  p.block_n = 0;
  p.act_param = 0.f;
  p.ln_stats = ln_stats;
  p.ln_colsum = ln_colsum;
  return hoigen_gemm_bf16(&p, s);
}

static int gemm_exp(const void* a, int lda, const void* w, int ldw, int M, int N, int K, const float* bias, float beta,
                    void* out_bf16, int ld_bf16, cudaStream_t s) {
  hoigen_gemm_params p = {};
  p.a = a; p.w = w; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldw = ldw;
  p.bias = bias; p.colscale = nullptr; p.act = HOIGEN_ACT_EXP;
  p.residual = nullptr; p.ld_res = 0;
  p.out_f32 = nullptr; p.ld_f32 = 0;
  p.out_bf16 = out_bf16; p.ld_bf16 = ld_bf16;
  p.split_k = 0;
  p.block_n = 0;
  p.act_param = beta;
  p.ln_stats = nullptr;
  p.ln_colsum = nullptr;
  return hoigen_gemm_bf16(&p, s);
}

#define HOIGEN_TRY(expr)          \
  do {                            \
    int _rc = (expr);             \
    if (_rc != HOIGEN_OK) return _rc; \
  } while (0)

}  // namespace hoigen

extern "C" {

int hoigen_encoder_forward(const hoigen_encoder_weights* w, const hoigen_encoder_buffers* buf, const float* images,
                           const float* prior, const uint8_t* mask, int32_t batch, int32_t n_max, int32_t num_layers,
                           hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(w && buf && images && prior && mask, "encoder_forward: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && num_layers >= 0 && num_layers <= 12, "encoder_forward: bad batch/layers");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int D = 768, T = 197, M = batch * T;
  // ---- patch embedding (conv1 as GEMM) + cls/pos + ln_pre -------------------------------------------------
  HOIGEN_TRY(hoigen_patchify_bf16(images, buf->patches, batch, s));
  HOIGEN_TRY(gemm(buf->patches, D, w->conv_w, D, batch * 196, D, D, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                  buf->patch_emb, D, nullptr, 0, s));
  HOIGEN_TRY(hoigen_embed_lnpre(buf->patch_emb, w->class_embedding, w->positional_embedding, w->ln_pre_w, w->ln_pre_b,
                                buf->x, buf->xb, batch, s));
  // ---- adapter K/V of the prior tokens, all layers at once ------------------------------------------------
  HOIGEN_TRY(hoigen_adapter_kv(prior, w->ad_in_proj_w, w->ad_in_proj_b, buf->adapter_kv, batch * n_max, 12, s));

  // Residual adds are DEFERRED out of the GEMM epilogues: every GEMM that feeds the fp32 stream writes a bf16 delta
  // (TMA-store epilogue) and the next LayerNorm pass (which streams x anyway) applies it.  The MLP output of layer
  // l-1 (delta2) is consumed twice: by linearity inside the adapter block's down-projection, and by ln_1's pass.
  // LayerNorm folded into the QKV / c_fc GEMMs (north_star item 1) when the folded weights were packed: the residual
  // pass then writes the bf16 copy of the RAW stream + per-row (mean, rstd) instead of a normalised `h`.
  const bool fold = w->qkv_wf && w->qkv_colsum && w->qkv_bf && w->fc_wf && w->fc_colsum && w->fc_bf && buf->row_stats;
  for (int l = 0; l < num_layers; ++l) {
    const size_t o768 = size_t(l) * D, o64 = size_t(l) * 64;
    // (1) x += mlp(l-1) ; adapter: down-proj + body (one tensor-core kernel) ; up-proj * scale -> delta
    hoigen_adapter_weights aw;
    aw.wd = (const uint16_t*)w->ad_down_w + size_t(l) * 64 * D; aw.down_b = w->ad_down_b + o64;
    aw.wq = (const uint16_t*)w->ad_wq + size_t(l) * 64 * 64; aw.wo = (const uint16_t*)w->ad_wo + size_t(l) * 64 * 64;
    aw.w1 = (const uint16_t*)w->ad_w1 + size_t(l) * 128 * 64; aw.w2 = (const uint16_t*)w->ad_w2 + size_t(l) * 64 * 128;
    aw.in_proj_b = w->ad_in_proj_b + size_t(l) * 192;
    aw.out_proj_b = w->ad_out_proj_b + o64;
    aw.linear1_b = w->ad_linear1_b + size_t(l) * 128;
    aw.linear2_b = w->ad_linear2_b + o64;
    aw.norm2_w = w->ad_norm2_w + o64; aw.norm2_b = w->ad_norm2_b + o64;
    aw.norm3_w = w->ad_norm3_w + o64; aw.norm3_b = w->ad_norm3_b + o64;
    aw.wup = (const uint16_t*)w->ad_up_w + size_t(l) * D * 64;
    HOIGEN_TRY(hoigen_adapter_block(buf->xb, l == 0 ? nullptr : buf->delta2, buf->adapter_kv + size_t(l) * batch * n_max * 128,
                                    mask, &aw, nullptr, buf->delta, batch, n_max, s));
    // (2) x += adapter ; x += out_proj(attention(ln_1(x)))
    if (fold) {
      HOIGEN_TRY(hoigen_add_rowstats768(buf->x, buf->delta, l == 0 ? nullptr : buf->delta2, w->ad_up_b + o768, buf->h,
                                        buf->row_stats, M, s));
      HOIGEN_TRY(gemm_ln(buf->h, D, (const uint16_t*)w->qkv_wf + size_t(l) * 3 * D * D, D, M, 3 * D, D,
                         w->qkv_bf + size_t(l) * 3 * D, HOIGEN_ACT_NONE, buf->row_stats, w->qkv_colsum + size_t(l) * 3 * D,
                         buf->qkv, 3 * D, s));
    } else {
      HOIGEN_TRY(hoigen_add_layernorm768(buf->x, buf->delta, l == 0 ? nullptr : buf->delta2, w->ad_up_b + o768,
                                         w->ln1_w + o768, w->ln1_b + o768, buf->h, nullptr, M, s));
      HOIGEN_TRY(gemm(buf->h, D, (const uint16_t*)w->qkv_w + size_t(l) * 3 * D * D, D, M, 3 * D, D,
                      w->qkv_b + size_t(l) * 3 * D, HOIGEN_ACT_NONE, nullptr, nullptr, 0, nullptr, 0, buf->qkv, 3 * D, s));
    }
    HOIGEN_TRY(hoigen_attention(buf->qkv, buf->attn, batch, s));
    HOIGEN_TRY(gemm(buf->attn, D, (const uint16_t*)w->out_w + size_t(l) * D * D, D, M, D, D, w->out_b + o768,
                    HOIGEN_ACT_NONE, nullptr, nullptr, 0, nullptr, 0, buf->delta, D, s));
    // (3) mlp: c_proj(quickgelu(c_fc(ln_2(x)))) -> delta2, added by the next adapter block (or the final LayerNorm)
    if (fold) {
      HOIGEN_TRY(hoigen_add_rowstats768(buf->x, buf->delta, nullptr, nullptr, buf->xb, buf->row_stats, M, s));
      HOIGEN_TRY(gemm_ln(buf->xb, D, (const uint16_t*)w->fc_wf + size_t(l) * 4 * D * D, D, M, 4 * D, D,
                         w->fc_bf + size_t(l) * 4 * D, HOIGEN_ACT_QUICKGELU, buf->row_stats, w->fc_colsum + size_t(l) * 4 * D,
                         buf->mlp, 4 * D, s));
    } else {
      HOIGEN_TRY(hoigen_add_layernorm768(buf->x, buf->delta, nullptr, nullptr, w->ln2_w + o768, w->ln2_b + o768, buf->h,
                                         l + 1 < num_layers ? buf->xb : nullptr, M, s));
      HOIGEN_TRY(gemm(buf->h, D, (const uint16_t*)w->fc_w + size_t(l) * 4 * D * D, D, M, 4 * D, D,
                      w->fc_b + size_t(l) * 4 * D, HOIGEN_ACT_QUICKGELU, nullptr, nullptr, 0, nullptr, 0, buf->mlp, 4 * D, s));
    }
    HOIGEN_TRY(gemm(buf->mlp, 4 * D, (const uint16_t*)w->proj_w + size_t(l) * D * 4 * D, 4 * D, M, D, 4 * D,
                    w->proj_b + o768, HOIGEN_ACT_NONE, nullptr, nullptr, 0, nullptr, 0, buf->delta2, D, s));
  }
  // ---- ln_post on ALL tokens (with the last pending residual), @ proj (768 -> 512) ---------------------------
  if (num_layers > 0) {
    HOIGEN_TRY(hoigen_add_layernorm768(buf->x, buf->delta2, nullptr, nullptr, w->ln_post_w, w->ln_post_b, buf->h, nullptr, M, s));
  } else {
    HOIGEN_TRY(hoigen_layernorm768(buf->x, w->ln_post_w, w->ln_post_b, nullptr, buf->h, M, s));
  }
  HOIGEN_TRY(gemm(buf->h, D, w->proj_t, D, M, 512, D, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0, buf->tokens_out, 512,
                  nullptr, 0, s));
  return HOIGEN_OK;
}

int hoigen_score_pairs(const hoigen_score_weights* w, const hoigen_score_buffers* buf, const float* tokens,
                       const float* dino_feats, const int32_t* pair_off, int32_t batch, int32_t ktot,
                       hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(w && buf && tokens && pair_off, "score_pairs: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ktot >= 0, "score_pairs: bad sizes");
  HOIGEN_CHECK_ARG(w->num_classes > 0 && w->cache_rows > 0 && (w->cache_rows % 8) == 0,
                   "score_pairs: cache_rows must be a positive multiple of 8 (got %d)", w->cache_rows);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int C = w->num_classes, N = w->cache_rows;
  const int L = buf->ld_logits > 0 ? int(buf->ld_logits) : C;   // row pitch of the logits accumulator
  HOIGEN_CHECK_ARG(L >= C, "score_pairs: ld_logits (%d) < num_classes (%d)", L, C);
  // The affinity is LINEAR in the reference (phi = f W^T + b, no exp: U:1156-1158), so the bias is carried exactly
  // in fp32 through the second GEMM's epilogue: ((f W^T + b) Y)/s = (f W^T) Y / s + (b Y)/s, bias_term = b Y.
  const bool exp_aff = w->affinity == 1;     // textbook Tip-Adapter exp(beta (f W^T + b)) instead of the reference's linear phi
  HOIGEN_CHECK_ARG(w->affinity == 0 || w->affinity == 1, "score_pairs: affinity must be 0 (linear) or 1 (exp)");
  // ---- per-image terms: global-CLIP cache (U:1133-1138) and DINO cache (U:1112-1115) -------------------------
  // g = feat_global / |feat_global|  (U:960) = token row 0 of each image
  HOIGEN_TRY(hoigen_rows_to_bf16(tokens, 197L * 512, batch, 512, 1, buf->g_bf16, s));
  if (exp_aff) {
    HOIGEN_CHECK_ARG(w->global_bias != nullptr, "score_pairs: the exp affinity needs global_bias");
    HOIGEN_TRY(gemm_exp(buf->g_bf16, 512, w->global_keys, 512, batch, N, 512, w->global_bias, w->beta, buf->phi_img, N, s));
  } else {
    HOIGEN_TRY(gemm(buf->g_bf16, 512, w->global_keys, 512, batch, N, 512, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                    nullptr, 0, buf->phi_img, N, s));
  }
  HOIGEN_TRY(gemm(buf->phi_img, N, w->label_t[2], N, batch, C, N, exp_aff ? nullptr : w->global_bias_term, HOIGEN_ACT_NONE,
                  w->colscale_global, nullptr, 0, buf->img_logits, C, nullptr, 0, s));
  if (dino_feats) {
    HOIGEN_CHECK_ARG(w->dino_keys != nullptr, "score_pairs: dino features given but no dino cache");
    HOIGEN_TRY(hoigen_rows_to_bf16(dino_feats, 2048, batch, 2048, 0, buf->d_bf16, s));
    if (exp_aff) {
      HOIGEN_CHECK_ARG(w->dino_bias != nullptr, "score_pairs: the exp affinity needs dino_bias");
      HOIGEN_TRY(gemm_exp(buf->d_bf16, 2048, w->dino_keys, 2048, batch, N, 2048, w->dino_bias, w->beta, buf->phi_img, N, s));
    } else {
      HOIGEN_TRY(gemm(buf->d_bf16, 2048, w->dino_keys, 2048, batch, N, 2048, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                      nullptr, 0, buf->phi_img, N, s));
    }
    HOIGEN_TRY(gemm(buf->phi_img, N, w->label_t[2], N, batch, C, N, exp_aff ? nullptr : w->dino_bias_term, HOIGEN_ACT_NONE,
                    w->colscale_dino, buf->img_logits, C, buf->img_logits, C, nullptr, 0, s));
  }
  if (ktot == 0) return HOIGEN_OK;
  // ---- pair terms: three cache branches (H, O, U) + text classifier ---------------------------------------------
  static const bool unfused = getenv("HOIGEN_CACHE_UNFUSED") != nullptr;      // A/B switch: the two-GEMM form through `phi`
  if (buf->cache_parts && C <= 128 && !unfused) {
    // ONE fused GEMM - f - GEMM kernel for the three branches (no `phi` in memory) + the fixed-order combine pass
    HOIGEN_TRY(hoigen_score_cache_fused(w, buf->pair_feat_bf16, exp_aff ? w->cache_bias : nullptr, buf->img_logits, pair_off, batch,
                                        ktot, w->affinity, w->beta, buf->cache_parts, buf->logits, L, s));
  } else {
    HOIGEN_CHECK_ARG(!exp_aff, "score_pairs: the exp affinity is implemented by the fused kernel (num_classes <= 128, cache_parts)");
    HOIGEN_CHECK_ARG(buf->phi != nullptr, "score_pairs: the two-GEMM form needs the phi workspace");
    HOIGEN_TRY(hoigen_broadcast_image_logits(buf->img_logits, pair_off, batch, ktot, C, L, buf->logits, s));
    for (int x = 0; x < 3; ++x) {
      const uint16_t* f = (const uint16_t*)buf->pair_feat_bf16 + size_t(x) * ktot * 512;
      HOIGEN_TRY(gemm(f, 512, w->cache_keys[x], 512, ktot, N, 512, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                      nullptr, 0, buf->phi, N, s));
      HOIGEN_TRY(gemm(buf->phi, N, w->label_t[x], N, ktot, C, N, w->bias_term[x], HOIGEN_ACT_NONE, w->colscale[x], buf->logits, L,
                      buf->logits, L, nullptr, 0, s));
    }
  }
  const uint16_t* fu = (const uint16_t*)buf->pair_feat_bf16 + size_t(2) * ktot * 512;
  HOIGEN_TRY(gemm(fu, 512, w->text_w, 512, ktot, C, 512, nullptr, HOIGEN_ACT_NONE, w->colscale_text, buf->logits, L,
                  buf->logits, L, nullptr, 0, s));
  return HOIGEN_OK;
}

// fp32-accurate form of hoigen_score_pairs (north_star "fp32 <= 1e-4" for RoI + scoring on identical features): the same
// chain, every product evaluated through the 3 x bf16 split on the SAME tcgen05 GEMM (operands K-concatenated, fp32
// accumulation in TMEM), phi kept in fp32.  ~6x the FLOPs of the bf16 path: a parity mode, not the benchmarked one.
int hoigen_score_pairs_fp32(const hoigen_score_weights_fp32* w, const hoigen_score_buffers_fp32* buf, const float* tokens,
                            const float* dino_feats, const float* pair_feat_f32, const int32_t* pair_off, int32_t batch,
                            int32_t ktot, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(w && buf && tokens && pair_off, "score_pairs_fp32: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ktot >= 0, "score_pairs_fp32: bad sizes");
  HOIGEN_CHECK_ARG(w->num_classes > 0 && w->cache_rows > 0 && (w->cache_rows % 8) == 0,
                   "score_pairs_fp32: cache_rows must be a positive multiple of 8 (got %d)", w->cache_rows);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int C = w->num_classes, N = w->cache_rows;
  const int L = buf->ld_logits > 0 ? int(buf->ld_logits) : C;
  HOIGEN_CHECK_ARG(L >= C, "score_pairs_fp32: ld_logits (%d) < num_classes (%d)", L, C);
  // ---- per-image terms (U:1133-1138, U:1112-1115) ------------------------------------------------------------
  HOIGEN_TRY(hoigen_rows_split3(tokens, 197L * 512, batch, 512, 1, 6, buf->g6, s));
  HOIGEN_TRY(gemm(buf->g6, 3072, w->global_keys6, 3072, batch, N, 3072, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                  buf->phi_img, N, nullptr, 0, s));
  HOIGEN_TRY(hoigen_rows_split3(buf->phi_img, N, batch, N, 0, 3, buf->phi_img3, s));
  HOIGEN_TRY(gemm(buf->phi_img3, 3 * N, w->label3_t[2], 3 * N, batch, C, 3 * N, w->global_bias_term, HOIGEN_ACT_NONE,
                  w->colscale_global, nullptr, 0, buf->img_logits, C, nullptr, 0, s));
  if (dino_feats) {
    HOIGEN_CHECK_ARG(w->dino_keys6 != nullptr, "score_pairs_fp32: dino features given but no dino cache");
    HOIGEN_TRY(hoigen_rows_split3(dino_feats, 2048, batch, 2048, 0, 6, buf->d6, s));
    HOIGEN_TRY(gemm(buf->d6, 6 * 2048, w->dino_keys6, 6 * 2048, batch, N, 6 * 2048, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                    buf->phi_img, N, nullptr, 0, s));
    HOIGEN_TRY(hoigen_rows_split3(buf->phi_img, N, batch, N, 0, 3, buf->phi_img3, s));
    HOIGEN_TRY(gemm(buf->phi_img3, 3 * N, w->label3_t[2], 3 * N, batch, C, 3 * N, w->dino_bias_term, HOIGEN_ACT_NONE,
                    w->colscale_dino, buf->img_logits, C, buf->img_logits, C, nullptr, 0, s));
  }
  if (ktot == 0) return HOIGEN_OK;
  HOIGEN_CHECK_ARG(pair_feat_f32 != nullptr, "score_pairs_fp32: needs the fp32 pair features");
  HOIGEN_TRY(hoigen_broadcast_image_logits(buf->img_logits, pair_off, batch, ktot, C, L, buf->logits, s));
  // ---- pair terms ----------------------------------------------------------------------------------------------
  HOIGEN_TRY(hoigen_rows_split3(pair_feat_f32, 512, 3 * ktot, 512, 0, 6, buf->feat6, s));
  for (int x = 0; x < 3; ++x) {
    const uint16_t* f6 = (const uint16_t*)buf->feat6 + size_t(x) * ktot * 3072;
    HOIGEN_TRY(gemm(f6, 3072, w->cache_keys6[x], 3072, ktot, N, 3072, nullptr, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                    buf->phi, N, nullptr, 0, s));
    HOIGEN_TRY(hoigen_rows_split3(buf->phi, N, ktot, N, 0, 3, buf->phi3, s));
    HOIGEN_TRY(gemm(buf->phi3, 3 * N, w->label3_t[x], 3 * N, ktot, C, 3 * N, w->bias_term[x], HOIGEN_ACT_NONE, w->colscale[x],
                    buf->logits, L, buf->logits, L, nullptr, 0, s));
  }
  const uint16_t* fu6 = (const uint16_t*)buf->feat6 + size_t(2) * ktot * 3072;
  HOIGEN_TRY(gemm(fu6, 3072, w->text_w6, 3072, ktot, C, 3072, nullptr, HOIGEN_ACT_NONE, w->colscale_text, buf->logits, L,
                  buf->logits, L, nullptr, 0, s));
  return HOIGEN_OK;
}

int hoigen_score_pairs_folded(const hoigen_folded_weights* w, const hoigen_score_buffers* buf, const float* tokens,
                              const float* dino_feats, const int32_t* pair_off, int32_t batch, int32_t ktot,
                              hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(w && buf && tokens && pair_off, "score_pairs_folded: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ktot >= 0 && w->num_classes > 0 && w->pair_w && w->bias_total, "score_pairs_folded: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int C = w->num_classes;
  const int L = buf->ld_logits > 0 ? int(buf->ld_logits) : C;
  HOIGEN_CHECK_ARG(L >= C, "score_pairs_folded: ld_logits (%d) < num_classes (%d)", L, C);
  // per-image terms: img_logits = bias_total + g E_G + d E_D
  bool have_img = false;
  if (w->global_w) {
    HOIGEN_TRY(hoigen_rows_to_bf16(tokens, 197L * 512, batch, 512, 1, buf->g_bf16, s));
    HOIGEN_TRY(gemm(buf->g_bf16, 512, w->global_w, 512, batch, C, 512, w->bias_total, HOIGEN_ACT_NONE, nullptr, nullptr, 0,
                    buf->img_logits, C, nullptr, 0, s));
    have_img = true;
  }
  if (w->dino_w) {
    HOIGEN_CHECK_ARG(dino_feats != nullptr, "score_pairs_folded: the folded weights include a DINO branch but no features were given");
    HOIGEN_TRY(hoigen_rows_to_bf16(dino_feats, 2048, batch, 2048, 0, buf->d_bf16, s));
    HOIGEN_TRY(gemm(buf->d_bf16, 2048, w->dino_w, 2048, batch, C, 2048, have_img ? nullptr : w->bias_total, HOIGEN_ACT_NONE,
                    nullptr, have_img ? buf->img_logits : nullptr, C, buf->img_logits, C, nullptr, 0, s));
    have_img = true;
  }
  if (ktot == 0) return HOIGEN_OK;
  if (have_img) HOIGEN_TRY(hoigen_broadcast_image_logits(buf->img_logits, pair_off, batch, ktot, C, L, buf->logits, s));
  // pair terms: the planar [3][Ktot][512] features against the three 512-column blocks of pair_w, accumulated in place
  for (int x = 0; x < 3; ++x) {
    const uint16_t* f = (const uint16_t*)buf->pair_feat_bf16 + size_t(x) * ktot * 512;
    const bool first = (x == 0 && !have_img);
    HOIGEN_TRY(gemm(f, 512, (const uint16_t*)w->pair_w + x * 512, 1536, ktot, C, 512, first ? w->bias_total : nullptr,
                    HOIGEN_ACT_NONE, nullptr, first ? nullptr : buf->logits, L, buf->logits, L, nullptr, 0, s));
  }
  return HOIGEN_OK;
}

}  // extern "C"
