// Row kernels of the DETR transformer (second half; reference detr/models/transformer.py: post-norm encoder
// layers 127-150, decoder layers 187-209, call site U:1596): every nn.Linear runs on the tcgen05 GEMM (gemm.cu); these are the two
// kernels in between.
//
//   hoigen_add_layernorm256   x += delta ; x = LayerNorm_256(x) ; side outputs bf16(x) and bf16(x + pos): the residual + norm of
//                             forward_post and the `with_pos_embed` operand of the NEXT attention's q / k projection in one pass
//   hoigen_attention_heads32  softmax(q k^T * scale + key_padding_mask) v for head_dim 32, any number of queries / keys
//                             (nn.MultiheadAttention inside the encoder / decoder layers; online softmax, fp32 arithmetic)
//
// d_model = 256 and head_dim = 32 are DETR's (detr/models/detr.py:308-, hidden_dim 256, nheads 8) and compile-time constants here.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int DM = 256;        // d_model
constexpr int DH = 32;         // head dimension
constexpr int ATT_Q = 128;     // queries per CTA (one per thread)
constexpr int ATT_K = 64;      // keys per shared-memory tile

__global__ void __launch_bounds__(256) add_layernorm256_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ delta,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ pos, int pos_rows,
                                                                __nv_bfloat16* __restrict__ x_bf16,
                                                                __nv_bfloat16* __restrict__ xpos_bf16, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + size_t(row) * DM) + lane * 2;
    const float4 a = xr[0], b = xr[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  if (delta != nullptr) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(delta + size_t(row) * DM) + lane);
    const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += __uint_as_float(w4[e] << 16);
      v[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
    }
  }
  if (gamma != nullptr) {   // nn.LayerNorm(256), eps 1e-5, biased variance, two passes in registers
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / DM);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = v[j] - mean; ss += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / DM) + 1e-5f);
    const float4* gp = reinterpret_cast<const float4*>(gamma) + lane * 2;
    const float4* bp = reinterpret_cast<const float4*>(beta) + lane * 2;
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] - mean) * rstd * gg[j] + bb[j];
  }
  {
    float4* xw = reinterpret_cast<float4*>(x + size_t(row) * DM) + lane * 2;
    xw[0] = make_float4(v[0], v[1], v[2], v[3]);
    xw[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  auto pack8 = [](const float (&f)[8]) {
    uint32_t p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
      p[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(p[0], p[1], p[2], p[3]);
  };
  if (x_bf16 != nullptr) *(reinterpret_cast<uint4*>(x_bf16 + size_t(row) * DM) + lane) = pack8(v);
  if (xpos_bf16 != nullptr) {
    float w[8];
    const float4* pp = reinterpret_cast<const float4*>(pos + size_t(row % pos_rows) * DM) + lane * 2;
    const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
    w[0] = v[0] + p0.x; w[1] = v[1] + p0.y; w[2] = v[2] + p0.z; w[3] = v[3] + p0.w;
    w[4] = v[4] + p1.x; w[5] = v[5] + p1.y; w[6] = v[6] + p1.z; w[7] = v[7] + p1.w;
    *(reinterpret_cast<uint4*>(xpos_bf16 + size_t(row) * DM) + lane) = pack8(w);
  }
}

// One thread per query, 128 queries of one (image, head) per CTA; keys / values stream through shared memory 64 at a time as
// fp32; the scores of 8 keys are formed, the running maximum / sum / output rescaled once per 8 (online softmax).  All lanes of
// a warp read the same key / value element, so every shared-memory read is a broadcast.
__global__ void __launch_bounds__(ATT_Q) attention_heads32_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                                   const __nv_bfloat16* __restrict__ k, int ldk,
                                                                   const __nv_bfloat16* __restrict__ v, int ldv,
                                                                   __nv_bfloat16* __restrict__ out, int ldo,
                                                                   const uint8_t* __restrict__ key_mask, int lq, int lk, float scale_log2e) {
  __shared__ __align__(16) float s_k[ATT_K][DH];
  __shared__ __align__(16) float s_v[ATT_K][DH];
  __shared__ float s_bias[ATT_K];          // 0 or -inf (masked / past the end)
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * ATT_Q + threadIdx.x;
  const bool q_ok = qi < lq;
  float qr[DH], o[DH];
  {
    const __nv_bfloat16* qp = q + (size_t(b) * lq + (q_ok ? qi : 0)) * ldq + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(qp) + c);
      const uint32_t w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qr[c * 8 + 2 * e] = __uint_as_float(w4[e] << 16) * scale_log2e;
        qr[c * 8 + 2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u) * scale_log2e;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < lk; k0 += ATT_K) {
    __syncthreads();                       // the previous tile has been consumed
    // ---- stage 64 keys / values of this head: 64 rows x 4 chunks of 8 bf16 each for K and for V = 512 chunk loads ----
    for (int i = threadIdx.x; i < ATT_K * (DH / 8) * 2; i += ATT_Q) {
      const int which = i / (ATT_K * (DH / 8)), r = (i / (DH / 8)) % ATT_K, c = i % (DH / 8);
      const int kj = k0 + r;
      uint4 w = make_uint4(0, 0, 0, 0);
      if (kj < lk) {
        const __nv_bfloat16* src = which ? v + (size_t(b) * lk + kj) * ldv : k + (size_t(b) * lk + kj) * ldk;
        w = __ldg(reinterpret_cast<const uint4*>(src + h * DH) + c);
      }
      float* dst = which ? &s_v[r][c * 8] : &s_k[r][c * 8];
      const uint32_t w4[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dst[2 * e] = __uint_as_float(w4[e] << 16);
        dst[2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
      }
    }
    if (threadIdx.x < ATT_K) {
      const int kj = k0 + threadIdx.x;
      const bool dead = kj >= lk || (key_mask != nullptr && key_mask[size_t(b) * lk + kj] != 0);
      s_bias[threadIdx.x] = dead ? -INFINITY : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int j0 = 0; j0 < ATT_K; j0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 kk = *reinterpret_cast<const float4*>(&s_k[j0 + j][d4 * 4]);
          acc = fmaf(qr[d4 * 4], kk.x, acc); acc = fmaf(qr[d4 * 4 + 1], kk.y, acc);
          acc = fmaf(qr[d4 * 4 + 2], kk.z, acc); acc = fmaf(qr[d4 * 4 + 3], kk.w, acc);
        }
        s[j] = acc + s_bias[j0 + j];
        cmax = fmaxf(cmax, s[j]);
      }
      if (cmax == -INFINITY) continue;     // (warp-uniform: the mask does not depend on the query)
      const float m_new = fmaxf(m, cmax);
      const float corr = exp2f(m - m_new); // m = -inf on the first live chunk: exp2(-inf) = 0
      m = m_new;
      l *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = exp2f(s[j] - m);
        l += p;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 vv = *reinterpret_cast<const float4*>(&s_v[j0 + j][d4 * 4]);
          o[d4 * 4] = fmaf(p, vv.x, o[d4 * 4]); o[d4 * 4 + 1] = fmaf(p, vv.y, o[d4 * 4 + 1]);
          o[d4 * 4 + 2] = fmaf(p, vv.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(p, vv.w, o[d4 * 4 + 3]);
        }
      }
    }
  }
  if (!q_ok) return;
  const float inv = l > 0.f ? 1.0f / l : 0.f;   // every key masked: zeros (torch gives NaN; DETR never masks a whole row)
  __nv_bfloat16* op = out + (size_t(b) * lq + qi) * ldo + h * DH;
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    uint32_t p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(o[c * 8 + 2 * e] * inv, o[c * 8 + 2 * e + 1] * inv);
      p[e] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    *(reinterpret_cast<uint4*>(op) + c) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// The same attention on the tensor cores: one CTA = 128 queries of one (image, head), keys in blocks of 128.
//   S = Q K^T      UMMA 128 x 128 x 16, two k-steps (head_dim 32), operands staged by the threads into 128B-swizzled tiles
//   softmax        thread = query row: S row read from TMEM, online max / sum, P = exp2(.) written back as packed bf16 OVER the
//                  row's own S columns (tcgen05.st; 128 keys = 64 columns), so P never passes through shared memory
//   O_blk = P V    UMMA 128 x 64 x 16 with A = P from TMEM, eight k-steps, V consumed MN-major straight from its key rows (rows
//                  padded to 64 columns); the block's 128 x 32 result is read back and accumulated in registers with the
//                  running rescale
// The next block's key / value rows are fetched into registers while the current block is being processed.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_fast(float x) {   // MUFU.EX2 without exp2f's range fix-up (ex2(-inf) = 0, no denormal care needed)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int AT2_THREADS = 128;
constexpr int AT2_KB = 128;                 // keys per block
constexpr int AT2_Q = 0;                    // 16 KiB  Q tile   [128 rows x 128 B], 64 B of each row used
constexpr int AT2_K = 16384;                // 16 KiB  K tile   [128 keys x 128 B]
constexpr int AT2_V = 32768;                // 16 KiB  V tile   [128 keys x 128 B], columns 32..63 zero
constexpr int AT2_BIAS = 49152;             // 128 floats   (P never touches shared memory: it overlays S in TMEM)
constexpr int AT2_BAR = AT2_BIAS + 512;
constexpr int AT2_SMEM_BYTES = AT2_BAR + 64 + 1024;

__global__ void __launch_bounds__(AT2_THREADS) attention_heads32_tc_kernel(const __nv_bfloat16* __restrict__ q, int ldq,
                                                                            const __nv_bfloat16* __restrict__ k, int ldk,
                                                                            const __nv_bfloat16* __restrict__ v, int ldv,
                                                                            __nv_bfloat16* __restrict__ out, int ldo,
                                                                            const uint8_t* __restrict__ key_mask, int lq, int lk,
                                                                            float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  float* s_bias = reinterpret_cast<float*>(sm + AT2_BIAS);
  const uint32_t bar = base + AT2_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + AT2_BAR + 16);
  const int t = threadIdx.x, warp = t >> 5;
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * 128 + t;

  if (t == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  // Q row -> A tile (this thread's row, four 16-byte chunks); the V tile's unused upper half is zeroed once
  {
    const __nv_bfloat16* qp = q + (size_t(b) * lq + (qi < lq ? qi : 0)) * ldq + h * DH;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(sm + AT2_Q + sw128_offset(uint32_t(t), uint32_t(c))) = __ldg(reinterpret_cast<const uint4*>(qp) + c);
#pragma unroll
    for (int c = 4; c < 8; ++c) *reinterpret_cast<uint4*>(sm + AT2_V + sw128_offset(uint32_t(t), uint32_t(c))) = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_row = tmem + (uint32_t(warp * 32) << 16);

  float o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) o[d] = 0.f;
  float m = -INFINITY, l = 0.f;
  uint32_t phase = 0;

  // this thread's key / value row of a block -> registers
  uint4 kr[4], vr[4];
  bool dead = true;
  auto fetch = [&](int k0) {
    const int kj = k0 + t;
    dead = kj >= lk || (key_mask != nullptr && key_mask[size_t(b) * lk + kj] != 0);
    if (kj < lk) {
      const uint4* kp = reinterpret_cast<const uint4*>(k + (size_t(b) * lk + kj) * ldk + h * DH);
      const uint4* vp = reinterpret_cast<const uint4*>(v + (size_t(b) * lk + kj) * ldv + h * DH);
#pragma unroll
      for (int c = 0; c < 4; ++c) { kr[c] = __ldg(kp + c); vr[c] = __ldg(vp + c); }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) { kr[c] = make_uint4(0, 0, 0, 0); vr[c] = make_uint4(0, 0, 0, 0); }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < lk; k0 += AT2_KB) {
    // ---- stage this block's rows (the previous block's MMAs have completed: both were waited for below) ----
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      *reinterpret_cast<uint4*>(sm + AT2_K + sw128_offset(uint32_t(t), uint32_t(c))) = kr[c];
      *reinterpret_cast<uint4*>(sm + AT2_V + sw128_offset(uint32_t(t), uint32_t(c))) = vr[c];
    }
    s_bias[t] = dead ? -INFINITY : 0.f;
    if (k0 + AT2_KB < lk) fetch(k0 + AT2_KB);          // in flight during the MMAs and the softmax
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16_ss(tmem, make_sdesc_sw128(base + AT2_Q + ks * 32), make_sdesc_sw128(base + AT2_K + ks * 32), idesc_s, ks > 0 ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // ---- online softmax of this row over the block's 128 keys; P -> A tile of the second product ----
    float bmax = -INFINITY;
    uint32_t r[32];
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {                   // pass 1: the block maximum
      tmem_ld_32x32b_x32(t_row + uint32_t(c4 * 32), r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) bmax = fmaxf(bmax, fmaf(__uint_as_float(r[j]), scale_log2e, s_bias[c4 * 32 + j]));
    }
    const float m_new = fmaxf(m, bmax);
    const bool live = m_new != -INFINITY;              // false only while every key so far is masked
    const float corr = live ? exp2f(m - m_new) : 1.f;
    m = m_new;
    l *= corr;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {                   // pass 2: probabilities; P chunk c4 (columns 16 c4 ..) lies inside S chunks
      tmem_ld_32x32b_x32(t_row + uint32_t(c4 * 32), r);   // <= c4 of this row, all of which are in registers or consumed by now
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t pk[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = g * 16 + 2 * e;
          const float p0 = live ? ex2_fast(fmaf(__uint_as_float(r[j]), scale_log2e, s_bias[c4 * 32 + j]) - m_new) : 0.f;
          const float p1 = live ? ex2_fast(fmaf(__uint_as_float(r[j + 1]), scale_log2e, s_bias[c4 * 32 + j + 1]) - m_new) : 0.f;
          l += p0 + p1;
          pk[e] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x8(t_row + uint32_t(c4 * 16 + g * 8), pk);   // keys 32 c4 + 16 g .. + 15 -> 8 packed columns
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, /*a_mn_major=*/0, /*b_mn_major=*/1);
#pragma unroll
      for (int ks = 0; ks < AT2_KB / 16; ++ks)
        umma_bf16_ts(tmem + 64u, tmem + uint32_t(ks * 8), make_sdesc_sw128(base + AT2_V + ks * 2048), idesc_o, ks > 0 ? 1u : 0u);
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    tmem_ld_32x32b_x32(t_row + 64u, r);
    tmem_wait_ld();
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] += __uint_as_float(r[d]);
    tc_fence_before();
  }
  if (qi < lq) {
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    __nv_bfloat16* op = out + (size_t(b) * lq + qi) * ldo + h * DH;
#pragma unroll
    for (int c = 0; c < DH / 8; ++c)
      *(reinterpret_cast<uint4*>(op) + c) = make_uint4(pack_bf16x2(o[c * 8] * inv, o[c * 8 + 1] * inv), pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv),
                                                       pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv), pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv));
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace hoigen

extern "C" {

using namespace hoigen;

int hoigen_add_layernorm256(float* x, const void* delta_bf16, const float* gamma, const float* beta, const float* pos,
                            int32_t pos_rows, void* x_bf16, void* xpos_bf16, int32_t rows, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(x != nullptr && rows > 0, "add_layernorm256: bad arguments");
  HOIGEN_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "add_layernorm256: gamma and beta go together");
  HOIGEN_CHECK_ARG(xpos_bf16 == nullptr || (pos != nullptr && pos_rows > 0), "add_layernorm256: xpos output needs pos");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("add_layernorm256", s, 0, double(rows) * DM * (8 + (delta_bf16 ? 2 : 0) + (x_bf16 ? 2 : 0) + (xpos_bf16 ? 6 : 0)));
  add_layernorm256_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, reinterpret_cast<const __nv_bfloat16*>(delta_bf16), gamma, beta, pos,
                                                         pos_rows > 0 ? pos_rows : 1, reinterpret_cast<__nv_bfloat16*>(x_bf16),
                                                         reinterpret_cast<__nv_bfloat16*>(xpos_bf16), rows);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_attention_heads32(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                             int32_t ldo, const uint8_t* key_mask, int32_t batch, int32_t lq, int32_t lk, int32_t heads,
                             float scale, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(q && k && v && out && batch > 0 && lq > 0 && lk > 0 && heads > 0, "attention_heads32: bad arguments");
  HOIGEN_CHECK_ARG((ldq % 8) == 0 && (ldk % 8) == 0 && (ldv % 8) == 0 && (ldo % 8) == 0 && ldq >= heads * DH && ldk >= heads * DH &&
                       ldv >= heads * DH && ldo >= heads * DH,
                   "attention_heads32: row pitches must be multiples of 8 and >= heads * 32");
  HOIGEN_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                     reinterpret_cast<uintptr_t>(out)) & 15) == 0, "attention_heads32: operands must be 16-byte aligned");
  HOIGEN_CHECK_ARG(batch <= 65535 && heads <= 65535, "attention_heads32: batch / heads too large");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("attention_heads32", s, 4.0 * batch * heads * double(lq) * lk * DH,
                 2.0 * batch * heads * DH * (2.0 * lq + 2.0 * lk * ((lq + ATT_Q - 1) / ATT_Q)));
  static const bool simt = getenv("HOIGEN_ATT32_SIMT") != nullptr;      // A/B switch: the fp32 SIMT form
  if (!simt) {
    HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(attention_heads32_tc_kernel), AT2_SMEM_BYTES));
    attention_heads32_tc_kernel<<<dim3((lq + 127) / 128, heads, batch), AT2_THREADS, AT2_SMEM_BYTES, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k), ldk,
        reinterpret_cast<const __nv_bfloat16*>(v), ldv, reinterpret_cast<__nv_bfloat16*>(out), ldo, key_mask, lq, lk,
        scale * 1.4426950408889634f);
    HOIGEN_CHECK_LAUNCH();
    return HOIGEN_OK;
  }
  attention_heads32_kernel<<<dim3((lq + ATT_Q - 1) / ATT_Q, heads, batch), ATT_Q, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), ldq, reinterpret_cast<const __nv_bfloat16*>(k), ldk,
      reinterpret_cast<const __nv_bfloat16*>(v), ldv, reinterpret_cast<__nv_bfloat16*>(out), ldo, key_mask, lq, lk,
      scale * 1.4426950408889634f);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
