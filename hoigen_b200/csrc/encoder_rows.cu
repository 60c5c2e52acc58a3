// Row-wise (memory-bound) kernels of the ViT-B/16 + InsAdapter encoder: patch extraction, cls/pos
// embedding + ln_pre, LayerNorm -> bf16, adapter K/V projection of the prior tokens and the adapter
// bottleneck body (2-head cross-attention over <= 32 prior tokens, LN, FFN, LN).
//
// Reference: CLIP_models_adapter_prior2.py:489-496 (embed + ln_pre), :409-415 (LayerNorm, fp32, eps 1e-5),
// :183-203 + :51-72 (Adapter.forward / TransformerDecoderLayer.forward_post).
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
template <bool EMBED>
__global__ void __launch_bounds__(256)
layernorm768_kernel(const float* __restrict__ in, const float* __restrict__ cls, const float* __restrict__ pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_bf16, int rows, const __nv_bfloat16* __restrict__ delta = nullptr,
                    float* __restrict__ stream_out = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
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
    for (int j = 0; j < 6; ++j) v[j] = src[lane + 32 * j];
    if (delta) {
      const uint2* dsrc = reinterpret_cast<const uint2*>(delta + long(row) * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint2 d = __ldg(dsrc + lane + 32 * j);
        v[j].x += __uint_as_float(d.x << 16); v[j].y += __uint_as_float(d.x & 0xffff0000u);
        v[j].z += __uint_as_float(d.y << 16); v[j].w += __uint_as_float(d.y & 0xffff0000u);
      }
      float4* dst = reinterpret_cast<float4*>(stream_out + long(row) * WIDTH);
#pragma unroll
      for (int j = 0; j < 6; ++j) dst[lane + 32 * j] = v[j];
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

// ------------------------------------------------------------------------------------------------
// Adapter body for one layer (C:186-200 with forward_post C:51-72), input D = relu(down_proj(x)) fp32 (M,64):
//   q = Wq D + bq ; 2 heads x 32 ; softmax over the image's unmasked prior tokens ; A = Wo (P V) + bo
//   T = LN_norm2(D + A) ; T = LN_norm3(T + W2 relu(W1 T + b1) + b2)          -> out bf16 (M,64)
// grid = (B, ROW_SPLIT); a block owns a contiguous slice of one image's 197 rows, a warp 4 rows at a time;
// lane owns elements (lane, lane+32) of every 64-vector.  Weights live transposed in shared memory.
// ------------------------------------------------------------------------------------------------
struct AdapterMidW {
  const uint32_t* packed;  // AM_W_WORDS words: the shared-memory image of the four matrices (see below)
  const float* in_b;       // (192) in_proj bias, q part = first 64
  const float* out_b;      // (64)
  const float* l1_b;       // (128)
  const float* l2_b;       // (64)
  const float* n2_w; const float* n2_b; const float* n3_w; const float* n3_b;  // (64) each
};

constexpr int AM_ROWS = 4;       // rows per warp pass
constexpr int AM_THREADS = 256;
constexpr int AM_MAXKEYS = 32;
// Shared memory (65 KiB -> 3 CTAs / SM): the four weight matrices as bf16 PAIRS, transposed to [input][output] with
// inputs (i, i + half) packed in one 32-bit word (halves the LDS count, fp32 FMAs), then K and V (fp32, padded rows).
constexpr int AM_W_WORDS = 32 * 64 * 2 + 32 * 128 + 64 * 64;   // WqP, WoP, W1P, W2P
constexpr int AM_SMEM_BYTES = (AM_W_WORDS + 2 * AM_MAXKEYS * 65) * 4;

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t bf_pack(float lo, float hi) { return pack_bf16x2(lo, hi); }

__device__ __forceinline__ void ln64(float (&a)[AM_ROWS], float (&b)[AM_ROWS], float g0, float g1, float b0, float b1) {
#pragma unroll
  for (int r = 0; r < AM_ROWS; ++r) {
    const float mean = warp_sum(a[r] + b[r]) * (1.0f / 64);
    const float da = a[r] - mean, db = b[r] - mean;
    const float rstd = rsqrtf(warp_sum(da * da + db * db) * (1.0f / 64) + 1e-5f);
    a[r] = da * rstd * g0 + b0;
    b[r] = db * rstd * g1 + b1;
  }
}

// y(lane), y(lane+32) += W[64 in][64 out] x, weights packed as P[i][o] = (W[i][o], W[i+32][o]), i < 32
__device__ __forceinline__ void matvec64(const uint32_t* __restrict__ P, int lane, const float (&x0)[AM_ROWS],
                                         const float (&x1)[AM_ROWS], float (&y0)[AM_ROWS], float (&y1)[AM_ROWS]) {
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const uint32_t wa = P[i * 64 + lane], wb = P[i * 64 + lane + 32];
    const float a0 = bf_lo(wa), b0 = bf_hi(wa), a1 = bf_lo(wb), b1 = bf_hi(wb);
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) {
      const float xa = __shfl_sync(0xffffffffu, x0[r], i), xb = __shfl_sync(0xffffffffu, x1[r], i);
      y0[r] = fmaf(a0, xa, fmaf(b0, xb, y0[r]));
      y1[r] = fmaf(a1, xa, fmaf(b1, xb, y1[r]));
    }
  }
}

__global__ void __launch_bounds__(AM_THREADS, 3)
adapter_mid_kernel(const float* __restrict__ D, const float* __restrict__ kv /* (tokens_total,128) this layer */,
                   const uint8_t* __restrict__ mask /* (B,n_max) 1 = padding */, AdapterMidW w,
                   __nv_bfloat16* __restrict__ out, int n_max, int row_split) {
  extern __shared__ uint32_t smw[];
  uint32_t* WqP = smw;                   // [32][64]
  uint32_t* WoP = WqP + 32 * 64;         // [32][64]
  uint32_t* W1P = WoP + 32 * 64;         // [32][128]   (W1[i][o], W1[i+32][o])
  uint32_t* W2P = W1P + 32 * 128;        // [64][64]    (W2[j][o], W2[j+64][o]),  j < 64
  float* Ks = reinterpret_cast<float*>(W2P + 64 * 64);   // [j][65]
  float* Vs = Ks + AM_MAXKEYS * 65;                      // [j][65]
  __shared__ int s_nkeys;
  __shared__ int s_keyidx[AM_MAXKEYS];

  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the packed image was laid out at build time (hoigen_b200/encoder.py::pack_adapter_mid): straight 16-byte copy
  for (int e = tid; e < AM_W_WORDS / 4; e += AM_THREADS)
    reinterpret_cast<uint4*>(smw)[e] = __ldg(reinterpret_cast<const uint4*>(w.packed) + e);
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < n_max; ++j)
      if (!mask[b * n_max + j]) s_keyidx[n++] = j;  // compact the unmasked keys (key_padding_mask, C:66)
    s_nkeys = n;
  }
  __syncthreads();
  const int nkeys = s_nkeys;
  for (int i = tid; i < nkeys * 128; i += AM_THREADS) {
    const int j = i / 128, c = i % 128;
    const float v = __ldg(kv + (long(b) * n_max + s_keyidx[j]) * 128 + c);
    if (c < 64) Ks[j * 65 + c] = v; else Vs[j * 65 + (c - 64)] = v;
  }
  __syncthreads();

  const float bq0 = __ldg(w.in_b + lane), bq1 = __ldg(w.in_b + lane + 32);
  const float bo0 = __ldg(w.out_b + lane), bo1 = __ldg(w.out_b + lane + 32);
  const float b2_0 = __ldg(w.l2_b + lane), b2_1 = __ldg(w.l2_b + lane + 32);
  float b1v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) b1v[k] = __ldg(w.l1_b + lane + 32 * k);
  const float n2g0 = __ldg(w.n2_w + lane), n2g1 = __ldg(w.n2_w + lane + 32);
  const float n2b0 = __ldg(w.n2_b + lane), n2b1 = __ldg(w.n2_b + lane + 32);
  const float n3g0 = __ldg(w.n3_w + lane), n3g1 = __ldg(w.n3_w + lane + 32);
  const float n3b0 = __ldg(w.n3_b + lane), n3b1 = __ldg(w.n3_b + lane + 32);
  const float qscale = 0.17677669529663687f;  // 32^-0.5

  const int rows_per_blk = (TOKENS + row_split - 1) / row_split;
  const int t_begin = blockIdx.y * rows_per_blk;
  const int t_end = min(TOKENS, t_begin + rows_per_blk);
  for (int t0 = t_begin + warp * AM_ROWS; t0 < t_end; t0 += (AM_THREADS / 32) * AM_ROWS) {
    float d0[AM_ROWS], d1[AM_ROWS];
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) {
      const int t = min(t0 + r, t_end - 1);  // tail rows recompute the last row (not stored)
      const float* src = D + (long(b) * TOKENS + t) * AD;
      d0[r] = src[lane];
      d1[r] = src[lane + 32];
    }
    // ---- q = Wq d + bq ----
    float q0[AM_ROWS], q1[AM_ROWS];
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) { q0[r] = bq0; q1[r] = bq1; }
    matvec64(WqP, lane, d0, d1, q0, q1);
    // ---- 2-head attention over the compacted keys: lane j scores key j ----
    float a0[AM_ROWS], a1[AM_ROWS];  // attention output (head 0 -> dims 0..31 held as a0, head 1 -> a1)
    {
      float s0[AM_ROWS], s1[AM_ROWS];
#pragma unroll
      for (int r = 0; r < AM_ROWS; ++r) { s0[r] = 0.f; s1[r] = 0.f; }
      const int jk = min(lane, max(nkeys - 1, 0));
#pragma unroll 8
      for (int dd = 0; dd < 32; ++dd) {
        const float k0 = Ks[jk * 65 + dd], k1 = Ks[jk * 65 + 32 + dd];
#pragma unroll
        for (int r = 0; r < AM_ROWS; ++r) {
          s0[r] = fmaf(__shfl_sync(0xffffffffu, q0[r], dd), k0, s0[r]);
          s1[r] = fmaf(__shfl_sync(0xffffffffu, q1[r], dd), k1, s1[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < AM_ROWS; ++r) {
        const bool valid = lane < nkeys;
        float v0 = valid ? s0[r] * qscale : -INFINITY, v1 = valid ? s1[r] * qscale : -INFINITY;
        const float m0 = warp_max(v0), m1 = warp_max(v1);
        float e0 = valid ? __expf(v0 - m0) : 0.f, e1 = valid ? __expf(v1 - m1) : 0.f;
        const float z0 = warp_sum(e0), z1 = warp_sum(e1);
        s0[r] = e0 / z0;  // all keys masked -> 0/0 = NaN, as in the reference
        s1[r] = e1 / z1;
        a0[r] = 0.f; a1[r] = 0.f;
      }
      for (int j = 0; j < nkeys; ++j) {
        const float vv0 = Vs[j * 65 + lane], vv1 = Vs[j * 65 + 32 + lane];
#pragma unroll
        for (int r = 0; r < AM_ROWS; ++r) {
          a0[r] = fmaf(__shfl_sync(0xffffffffu, s0[r], j), vv0, a0[r]);
          a1[r] = fmaf(__shfl_sync(0xffffffffu, s1[r], j), vv1, a1[r]);
        }
      }
      if (nkeys == 0) {
#pragma unroll
        for (int r = 0; r < AM_ROWS; ++r) { a0[r] = s0[r]; a1[r] = s1[r]; }  // propagate NaN
      }
    }
    // ---- out-proj + residual + norm2 ----
    float t0v[AM_ROWS], t1v[AM_ROWS];
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) { t0v[r] = bo0; t1v[r] = bo1; }
    matvec64(WoP, lane, a0, a1, t0v, t1v);
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) { t0v[r] += d0[r]; t1v[r] += d1[r]; }
    ln64(t0v, t1v, n2g0, n2g1, n2b0, n2b1);
    // ---- FFN 64 -> 128 (relu) -> 64, residual, norm3 ----
    float h[4][AM_ROWS];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int r = 0; r < AM_ROWS; ++r) h[k][r] = b1v[k];
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      uint32_t wv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wv[k] = W1P[i * 128 + lane + 32 * k];
#pragma unroll
      for (int r = 0; r < AM_ROWS; ++r) {
        const float xa = __shfl_sync(0xffffffffu, t0v[r], i), xb = __shfl_sync(0xffffffffu, t1v[r], i);
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k][r] = fmaf(bf_lo(wv[k]), xa, fmaf(bf_hi(wv[k]), xb, h[k][r]));
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int r = 0; r < AM_ROWS; ++r) h[k][r] = fmaxf(h[k][r], 0.f);
    float f0[AM_ROWS], f1[AM_ROWS];
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) { f0[r] = b2_0; f1[r] = b2_1; }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const uint32_t w0 = W2P[(kk * 32 + i) * 64 + lane], w1 = W2P[(kk * 32 + i) * 64 + lane + 32];
#pragma unroll
        for (int r = 0; r < AM_ROWS; ++r) {
          const float xl = __shfl_sync(0xffffffffu, h[kk][r], i), xh = __shfl_sync(0xffffffffu, h[kk + 2][r], i);
          f0[r] = fmaf(bf_lo(w0), xl, fmaf(bf_hi(w0), xh, f0[r]));
          f1[r] = fmaf(bf_lo(w1), xl, fmaf(bf_hi(w1), xh, f1[r]));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) { f0[r] += t0v[r]; f1[r] += t1v[r]; }
    ln64(f0, f1, n3g0, n3g1, n3b0, n3b1);
#pragma unroll
    for (int r = 0; r < AM_ROWS; ++r) {
      const int t = t0 + r;
      if (t < t_end) {
        __nv_bfloat16* dst = out + (long(b) * TOKENS + t) * AD;
        dst[lane] = __float2bfloat16_rn(f0[r]);
        dst[lane + 32] = __float2bfloat16_rn(f1[r]);
      }
    }
  }
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

int hoigen_add_layernorm768(float* x, const void* delta_bf16, const float* gamma, const float* beta, void* out_bf16,
                            int32_t rows, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(x && delta_bf16 && gamma && beta && out_bf16 && rows > 0, "add_layernorm768: bad arguments");
  KernelScope ks("add_layernorm768", reinterpret_cast<cudaStream_t>(stream), 0, double(rows) * WIDTH * (4 + 2 + 4 + 2));
  layernorm768_kernel<false><<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, nullptr, nullptr, gamma, beta, nullptr, reinterpret_cast<__nv_bfloat16*>(out_bf16), rows,
      reinterpret_cast<const __nv_bfloat16*>(delta_bf16), x);
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

int hoigen_adapter_mid(const float* d, const float* kv_layer, const uint8_t* mask, const hoigen_adapter_mid_weights* w,
                       void* out_bf16, int32_t batch, int32_t n_max, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(d && kv_layer && mask && w && w->packed && out_bf16 && batch > 0, "adapter_mid: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(w->packed) & 15) == 0, "adapter_mid: packed weights must be 16-byte aligned");
  HOIGEN_CHECK_ARG(n_max > 0 && n_max <= AM_MAXKEYS, "adapter_mid: n_max must be in [1,%d] (got %d)", AM_MAXKEYS, n_max);
  static bool attr_set = false;
  if (!attr_set) {
    HOIGEN_CHECK_CUDA(cudaFuncSetAttribute(adapter_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AM_SMEM_BYTES));
    attr_set = true;
  }
  AdapterMidW mw;
  mw.packed = w->packed; mw.in_b = w->in_proj_b; mw.out_b = w->out_proj_b;
  mw.l1_b = w->linear1_b; mw.l2_b = w->linear2_b;
  mw.n2_w = w->norm2_w; mw.n2_b = w->norm2_b; mw.n3_w = w->norm3_w; mw.n3_b = w->norm3_b;
  // split each image's 197 rows so that the grid covers the SMs at least ~2x
  int split = 1;
  while (batch * split < 3 * num_sms() && split < 8) split *= 2;
  dim3 grid(batch, split);
  KernelScope ks("adapter_mid", reinterpret_cast<cudaStream_t>(stream), 2.0 * batch * TOKENS * (64 * 64 * 2 + 2 * 64 * 128 + 2 * 64 * n_max),
                 double(batch) * TOKENS * 64 * (4 + 2));
  adapter_mid_kernel<<<grid, AM_THREADS, AM_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
      d, kv_layer, mask, mw, reinterpret_cast<__nv_bfloat16*>(out_bf16), n_max, split);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
