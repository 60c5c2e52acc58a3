// Row-wise (memory-bound) kernels of the ViT-B/16 + InsAdapter encoder: patch extraction, cls/pos
// embedding + ln_pre, LayerNorm -> bf16 (optionally fused with a deferred residual add) and the adapter K/V
// projection of the prior tokens.  (The adapter bottleneck body lives in adapter_tc.cu.)
//
// Reference: CLIP_models_adapter_prior2.py:489-496 (embed + ln_pre), :409-415 (LayerNorm, fp32, eps 1e-5),
// :183-203 + :51-72 (Adapter.forward / TransformerDecoderLayer.forward_post).
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int WIDTH = 768;
constexpr int TOKENS = 197;
constexpr int GRID14 = 14;
constexpr int AD = 64;  // adapter bottleneck width

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

// ------------------------------------------------------------------------------------------------
// images (B,3,224,224) fp32 -> patches (B*196, 768) bf16, column = c*256 + ky*16 + kx   (C:491 as im2col)
// one thread = 8 consecutive pixels of one image row
// ------------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B) {
  const long total = long(B) * 3 * 224 * 28;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int xc = int(i % 28);
    const int y = int((i / 28) % 224);
    const int c = int((i / (28 * 224)) % 3);
    const int b = int(i / (28 * 224 * 3));
    const float4* src = reinterpret_cast<const float4*>(img + ((long(b) * 3 + c) * 224 + y) * 224 + xc * 8);
    const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
    const int py = y >> 4, ky = y & 15, px = xc >> 1, half = xc & 1;
    uint4 pk;
    pk.x = pack_bf16x2(v0.x, v0.y);
    pk.y = pack_bf16x2(v0.z, v0.w);
    pk.z = pack_bf16x2(v1.x, v1.y);
    pk.w = pack_bf16x2(v1.z, v1.w);
    __nv_bfloat16* dst = out + (long(b) * 196 + py * GRID14 + px) * WIDTH + c * 256 + ky * 16 + half * 8;
    *reinterpret_cast<uint4*>(dst) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over 768 columns, one warp per row (24 values per lane, two-pass statistics in registers).
//   EMBED = false: x = in[row]                                         -> out_bf16 (and optionally out_f32)
//   EMBED = true : x = (t == 0 ? cls : in[b*196 + t - 1]) + pos[t]      (C:494-496), row = b*197 + t
// ------------------------------------------------------------------------------------------------
//   delta != nullptr (EMBED = false): x = in[row] + delta[row] (bf16), and x is written back to `stream_out` (the
//                  deferred residual add of the preceding GEMM: X += up-proj / out-proj output)
//   STATS = true: the LayerNorm itself is folded into the consuming GEMM (gemm.cu, ln_stats): this pass only applies the
//                  pending residuals, writes the fp32 stream + its bf16 copy and emits (mean, rstd) per row.
template <bool EMBED, bool STATS = false>
__global__ void __launch_bounds__(256)
layernorm768_kernel(const float* __restrict__ in, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_bf16, int rows, const __nv_bfloat16* __restrict__ delta = nullptr,
                    float* __restrict__ stream_out = nullptr, const __nv_bfloat16* __restrict__ delta_b = nullptr,
                    __nv_bfloat16* __restrict__ stream_bf16 = nullptr, const float* __restrict__ col_bias = nullptr,
                    float2* __restrict__ stats_out = nullptr, int keep_l2 = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;   // one warp per row; rows per CTA = launch-time choice
  if (row >= rows) return;
  float4 v[6];
  if (EMBED) {
    const int b = row / TOKENS, t = row % TOKENS;
    const float4* src = (t == 0) ? reinterpret_cast<const float4*>(cls)
                                 : reinterpret_cast<const float4*>(in + (long(b) * 196 + (t - 1)) * WIDTH);
    const float4* p = reinterpret_cast<const float4*>(pos + long(t) * WIDTH);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float4 a = __ldg(src + lane + 32 * j), q = __ldg(p + lane + 32 * j);
      v[j] = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(in + long(row) * WIDTH);
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = keep_l2 ? src[lane + 32 * j] : __ldcs(src + lane + 32 * j);   // streamed: the fp32 stream is not re-read before ~100 MB of other traffic
    if (delta) {
      const uint2* dsrc = reinterpret_cast<const uint2*>(delta + long(row) * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint2 d = __ldg(dsrc + lane + 32 * j);
        v[j].x += __uint_as_float(d.x << 16); v[j].y += __uint_as_float(d.x & 0xffff0000u);
        v[j].z += __uint_as_float(d.y << 16); v[j].w += __uint_as_float(d.y & 0xffff0000u);
      }
      if (delta_b) {
        const uint2* dsrc2 = reinterpret_cast<const uint2*>(delta_b + long(row) * WIDTH);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const uint2 d = __ldg(dsrc2 + lane + 32 * j);
          v[j].x += __uint_as_float(d.x << 16); v[j].y += __uint_as_float(d.x & 0xffff0000u);
          v[j].z += __uint_as_float(d.y << 16); v[j].w += __uint_as_float(d.y & 0xffff0000u);
        }
      }
      if (col_bias) {   // bias row of the GEMM that produced delta (kept out of that GEMM's epilogue)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(col_bias) + lane + 32 * j);
          v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;
        }
      }
      float4* dst = reinterpret_cast<float4*>(stream_out + long(row) * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) {   // evict-first: keep L2 for h / delta / qkv, which ARE re-read
        if (keep_l2) dst[lane + 32 * j] = v[j]; else __stcs(dst + lane + 32 * j, v[j]);
      }
      if (stream_bf16) {   // bf16 copy of the updated stream: the next adapter block's tensor-core operand
        uint2* dstb = reinterpret_cast<uint2*>(stream_bf16 + long(row) * WIDTH);
#pragma unroll
        for (int j = 0; j < 6; ++j) dstb[lane + 32 * j] = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / WIDTH);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / WIDTH) + 1e-5f);
  if (STATS) {
    if (lane == 0) stats_out[row] = make_float2(mean, rstd);
    return;
  }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 g = __ldg(g4 + lane + 32 * j), bb = __ldg(b4 + lane + 32 * j);
    float4 y;
    y.x = (v[j].x - mean) * rstd * g.x + bb.x;
    y.y = (v[j].y - mean) * rstd * g.y + bb.y;
    y.z = (v[j].z - mean) * rstd * g.z + bb.z;
    y.w = (v[j].w - mean) * rstd * g.w + bb.w;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + long(row) * WIDTH)[lane + 32 * j] = y;
    if (out_bf16) {
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + long(row) * WIDTH)[lane + 32 * j] = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The encoder's residual pass (25 launches per forward): x += delta (+ delta_b) (+ col_bias), write x (fp32) [+ its bf16
// copy], then LayerNorm -> bf16 h  (STATS: emit (mean, rstd) instead of h).  PERSISTENT and software-pipelined: a fixed grid
// of resident CTAs (no partial last wave: 1576 CTAs on 148 x 4 slots used to run 2.66 waves as 3), one warp per row, and
// the loads of the warp's NEXT row are in flight while the current row is reduced, normalised and stored.
// ------------------------------------------------------------------------------------------------
struct RowRegs {
  float4 x[6];
  uint2 d[6];
  uint2 e[6];
};

// How the fp32 stream is cached by the residual pass (HOIGEN_X_L2, A/B in profiles/r02_ab_notes.md):
//   0 = streaming (ld/st .cs, evict-first; the default)   1 = default caching   2 = L2 evict_last policy on the stream's
//   loads and stores.  Measured: all three within noise (3.74-3.76 ms/step); a persisting access-policy window over x on
//   the launching stream (82.9 MB carve-out) was WORSE (3.97 ms/step, pass 25 -> 29.6 us) and is not kept.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_f4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f4_hint(float4* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ void residual_row_load(RowRegs& r, const float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                                                  const __nv_bfloat16* __restrict__ delta_b, long row, int lane, int xmode,
                                                  uint64_t pol) {
  const float4* src = reinterpret_cast<const float4*>(x + row * WIDTH);
  const uint2* d1 = reinterpret_cast<const uint2*>(delta + row * WIDTH);
  if (xmode == 0) {
#pragma unroll
    for (int j = 0; j < 6; ++j) r.x[j] = __ldcs(src + lane + 32 * j);
  } else if (xmode == 2) {
#pragma unroll
    for (int j = 0; j < 6; ++j) r.x[j] = ld_f4_hint(src + lane + 32 * j, pol);
  } else {
#pragma unroll
    for (int j = 0; j < 6; ++j) r.x[j] = src[lane + 32 * j];
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) r.d[j] = __ldg(d1 + lane + 32 * j);
  if (delta_b) {
    const uint2* d2 = reinterpret_cast<const uint2*>(delta_b + row * WIDTH);
#pragma unroll
    for (int j = 0; j < 6; ++j) r.e[j] = __ldg(d2 + lane + 32 * j);
  }
}

#ifndef HOIGEN_LN_THREADS
#define HOIGEN_LN_THREADS 256       // 128: a CTA is 16 Ki registers and fits beside a register-capped GEMM / attention CTA
#endif
template <bool STATS>
__global__ void __launch_bounds__(HOIGEN_LN_THREADS, 512 / HOIGEN_LN_THREADS)
residual_ln768_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ delta, const __nv_bfloat16* __restrict__ delta_b,
                      const float* __restrict__ col_bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ out_bf16, __nv_bfloat16* __restrict__ stream_bf16,
                      float2* __restrict__ stats_out, int rows, int xmode) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long stride = long(gridDim.x) * (blockDim.x >> 5);
  long row = long(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const uint64_t pol = xmode == 2 ? l2_policy_evict_last() : 0;
  RowRegs cur, nxt;
  residual_row_load(cur, x, delta, delta_b, row, lane, xmode, pol);
  for (; row < rows; row += stride) {
    const long nrow = row + stride;
    if (nrow < rows) residual_row_load(nxt, x, delta, delta_b, nrow, lane, xmode, pol);      // in flight during this row's work
    float4 v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j] = cur.x[j];
      v[j].x += __uint_as_float(cur.d[j].x << 16); v[j].y += __uint_as_float(cur.d[j].x & 0xffff0000u);
      v[j].z += __uint_as_float(cur.d[j].y << 16); v[j].w += __uint_as_float(cur.d[j].y & 0xffff0000u);
    }
    if (delta_b) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v[j].x += __uint_as_float(cur.e[j].x << 16); v[j].y += __uint_as_float(cur.e[j].x & 0xffff0000u);
        v[j].z += __uint_as_float(cur.e[j].y << 16); v[j].w += __uint_as_float(cur.e[j].y & 0xffff0000u);
      }
    }
    if (col_bias) {   // bias row of the GEMM that produced delta (kept out of that GEMM's epilogue); L1-resident
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(col_bias) + lane + 32 * j);
        v[j].x += b.x; v[j].y += b.y; v[j].z += b.z; v[j].w += b.w;
      }
    }
    float4* dst = reinterpret_cast<float4*>(x + row * WIDTH);
    if (xmode == 0) {
#pragma unroll
      for (int j = 0; j < 6; ++j) __stcs(dst + lane + 32 * j, v[j]);   // evict-first: keep L2 for h / delta / qkv, which ARE re-read
    } else if (xmode == 2) {
#pragma unroll
      for (int j = 0; j < 6; ++j) st_f4_hint(dst + lane + 32 * j, v[j], pol);
    } else {
#pragma unroll
      for (int j = 0; j < 6; ++j) dst[lane + 32 * j] = v[j];
    }
    if (stream_bf16) {   // bf16 copy of the updated stream: the next adapter block's tensor-core operand
      uint2* dstb = reinterpret_cast<uint2*>(stream_bf16 + row * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) dstb[lane + 32 * j] = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / WIDTH);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / WIDTH) + 1e-5f);
    if (STATS) {
      if (lane == 0) stats_out[row] = make_float2(mean, rstd);
    } else {
      const float4* g4 = reinterpret_cast<const float4*>(gamma);
      const float4* b4 = reinterpret_cast<const float4*>(beta);
      uint2* oh = reinterpret_cast<uint2*>(out_bf16 + row * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j), bb = __ldg(b4 + lane + 32 * j);
        oh[lane + 32 * j] = make_uint2(pack_bf16x2((v[j].x - mean) * rstd * g.x + bb.x, (v[j].y - mean) * rstd * g.y + bb.y),
                                       pack_bf16x2((v[j].z - mean) * rstd * g.z + bb.z, (v[j].w - mean) * rstd * g.w + bb.w));
      }
    }
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------
// Adapter K/V of the prior tokens for every layer:  KV[l][tok][0:64] = Wk_l p + bk_l, [64:128] = Wv_l p + bv_l
// (rows 64..191 of multihead_attn.in_proj_weight, C:63-66).  grid = (ceil(tokens/16), layers), 128 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
adapter_kv_kernel(const float* __restrict__ prior, const float* __restrict__ in_w, const float* __restrict__ in_b,
                  float* __restrict__ kv, int tokens) {
  __shared__ float wT[AD][128 + 1];
  __shared__ float x[16][AD];
  const int l = blockIdx.y;
  const int t0 = blockIdx.x * 16;
  const float* w = in_w + long(l) * 192 * AD + AD * AD;  // skip the q rows
  for (int i = threadIdx.x; i < 128 * AD; i += 128) wT[i % AD][i / AD] = __ldg(w + i);
  for (int i = threadIdx.x; i < 16 * AD; i += 128) {
    const int t = t0 + i / AD;
    x[i / AD][i % AD] = t < tokens ? __ldg(prior + long(t) * AD + (i % AD)) : 0.f;
  }
  __syncthreads();
  const int o = threadIdx.x;
  const float bias = __ldg(in_b + l * 192 + AD + o);
  float acc[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) acc[r] = bias;
  for (int i = 0; i < AD; ++i) {
    const float wv = wT[i][o];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = fmaf(wv, x[r][i], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 16; ++r)
    if (t0 + r < tokens) kv[(long(l) * tokens + t0 + r) * 128 + o] = acc[r];
}

// experiment switch (HOIGEN_LN_KEEP_L2=1): default caching of the fp32 stream in the residual passes instead of streaming it
static int residual_x_l2_mode() {
  static const int v = getenv("HOIGEN_X_L2") ? atoi(getenv("HOIGEN_X_L2")) : 0;
  return v;
}
static int ln_keep_l2() {
  static const int v = getenv("HOIGEN_LN_KEEP_L2") ? atoi(getenv("HOIGEN_LN_KEEP_L2")) : 0;
  return v;
}

}  // namespace hoigen

extern "C" {

int hoigen_patchify_bf16(const float* images, void* patches, int32_t batch, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(images && patches && batch > 0, "patchify: bad arguments");
  const long total = long(batch) * 3 * 224 * 28;
  const int blocks = int(min(long(num_sms()) * 8, (total + 255) / 256));
  KernelScope ks("patchify", reinterpret_cast<cudaStream_t>(stream), 0, double(batch) * 3 * 224 * 224 * 6);
  patchify_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      images, reinterpret_cast<__nv_bfloat16*>(patches), batch);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_embed_lnpre(const float* patch_emb, const float* cls, const float* pos, const float* gamma,
                       const float* beta, float* x_f32, void* x_bf16, int32_t batch, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(patch_emb && cls && pos && gamma && beta && batch > 0, "embed_lnpre: bad arguments");
  const int rows = batch * TOKENS;
  KernelScope ks("embed_lnpre", reinterpret_cast<cudaStream_t>(stream), 0, double(rows) * WIDTH * (4 + (x_f32 ? 4 : 0) + (x_bf16 ? 2 : 0)));
  layernorm768_kernel<true><<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      patch_emb, cls, pos, gamma, beta, x_f32, reinterpret_cast<__nv_bfloat16*>(x_bf16), rows);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_layernorm768(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                        int32_t rows, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(x && gamma && beta && rows > 0 && (out_f32 || out_bf16), "layernorm768: bad arguments");
  KernelScope ks("layernorm768", reinterpret_cast<cudaStream_t>(stream), 0, double(rows) * WIDTH * (4 + (out_f32 ? 4 : 0) + (out_bf16 ? 2 : 0)));
  layernorm768_kernel<false><<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, nullptr, nullptr, gamma, beta, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_add_layernorm768(float* x, const void* delta_bf16, const void* delta2_bf16, const float* col_bias,
                            const float* gamma, const float* beta, void* out_bf16, void* x_bf16, int32_t rows,
                            hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(x && delta_bf16 && gamma && beta && out_bf16 && rows > 0, "add_layernorm768: bad arguments");
  KernelScope ks("add_layernorm768", reinterpret_cast<cudaStream_t>(stream), 0,
                 double(rows) * WIDTH * (4 + 2 + 4 + 2 + (delta2_bf16 ? 2 : 0) + (x_bf16 ? 2 : 0)));
  static const bool legacy = getenv("HOIGEN_LN_LEGACY") != nullptr;      // A/B switch: the one-wave-per-8-rows kernel
  if (legacy) {
    constexpr int rpc = 8;
    layernorm768_kernel<false><<<(rows + rpc - 1) / rpc, rpc * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, nullptr, nullptr, gamma, beta, nullptr, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows,
        reinterpret_cast<const __nv_bfloat16*>(delta_bf16), x, reinterpret_cast<const __nv_bfloat16*>(delta2_bf16),
        reinterpret_cast<__nv_bfloat16*>(x_bf16), col_bias, nullptr, ln_keep_l2());
  } else {
    constexpr int wpc = HOIGEN_LN_THREADS / 32;
    const int grid = std::min((rows + wpc - 1) / wpc, (16 / wpc) * num_sms());
    residual_ln768_kernel<false><<<grid, HOIGEN_LN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, reinterpret_cast<const __nv_bfloat16*>(delta_bf16), reinterpret_cast<const __nv_bfloat16*>(delta2_bf16), col_bias, gamma,
        beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), reinterpret_cast<__nv_bfloat16*>(x_bf16), nullptr, rows, residual_x_l2_mode());
  }
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_add_rowstats768(float* x, const void* delta_bf16, const void* delta2_bf16, const float* col_bias, void* x_bf16,
                           float* stats, int32_t rows, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(x && delta_bf16 && x_bf16 && stats && rows > 0, "add_rowstats768: bad arguments");
  KernelScope ks("add_rowstats768", reinterpret_cast<cudaStream_t>(stream), 0,
                 double(rows) * WIDTH * (4 + 2 + 4 + 2 + (delta2_bf16 ? 2 : 0)) + double(rows) * 8);
  constexpr int wpc = HOIGEN_LN_THREADS / 32;
  const int grid = std::min((rows + wpc - 1) / wpc, (16 / wpc) * num_sms());
  residual_ln768_kernel<true><<<grid, HOIGEN_LN_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(delta_bf16), reinterpret_cast<const __nv_bfloat16*>(delta2_bf16), col_bias, nullptr,
      nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(x_bf16), reinterpret_cast<float2*>(stats), rows, residual_x_l2_mode());
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_adapter_kv(const float* prior, const float* in_proj_w, const float* in_proj_b, float* kv,
                      int32_t tokens, int32_t layers, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(prior && in_proj_w && in_proj_b && kv && tokens > 0 && layers > 0, "adapter_kv: bad arguments");
  dim3 grid((tokens + 15) / 16, layers);
  KernelScope ks("adapter_kv", reinterpret_cast<cudaStream_t>(stream), 2.0 * tokens * 128 * 64 * layers, double(tokens) * (64 + 128.0 * layers) * 4);
  adapter_kv_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(prior, in_proj_w, in_proj_b, kv, tokens);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
