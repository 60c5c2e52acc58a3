// TMA-fed tcgen05/TMEM bf16 GEMM with fused epilogue (bias / QuickGELU / ReLU / column scale /
// fp32 residual add / dual fp32+bf16 store).  out = epi(A[M,K] @ W[N,K]^T).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0    : TMA producer   (A tile 128x64, W tile BNx64, 128B swizzle, STAGES-deep mbarrier ring)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accum in TMEM)
//   warps 2-9 : epilogue       (tcgen05.ld 32 lanes x 16 columns -> registers -> global / smem staging + TMA store)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Two kernels: gemm_bf16_kernel (one CTA per 128 x BN tile, optional deterministic split-K for long-K / few-tile shapes)
// and gemm2_bf16_kernel (a cluster of two CTAs per 256 x BN tile, cta_group::2, stream-K split of the leftover tiles);
// choose_config picks per call.
//
// Replaces the cuBLAS SGEMMs the reference dispatches from CLIP_models_adapter_prior2.py:443-445 (in/out
// proj), :428-432 (c_fc/c_proj), :184/:201 (adapter down/up), :491 (conv1 as GEMM), :505 (@ proj) and
// upt_tip_cache_model_free_finetune_distill3.py:1156-1163 (cache / text GEMMs).
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int GEMM2_THREADS = 352; // CTA-pair kernel: + warp 10 = output DMA (TMA stores / identity loads of the staging panels)

// Hand-over of the 64-column staging panels between the epilogue warps and the DMA warp of the CTA-pair kernel:
// ready[p] completes once per tile when panel p may be written (its previous TMA store has been read out, and -- identity
// recipes -- the identity tile has landed in it); written[p] when every thread that writes panel p has done so.
struct PanelSync {
  uint32_t ready, written;   // smem addresses of the two barrier arrays (8 bytes per panel); 0 = no staging (no-op)
  uint32_t phase;
  __device__ __forceinline__ void enter(int p) const {
    if (ready) mbar_wait(ready + 8u * p, phase);
  }
  __device__ __forceinline__ void leave(int p) const {
    if (ready) {
      fence_proxy_async_smem();          // this thread's staging writes -> visible to the TMA engine
      mbar_arrive(written + 8u * p);
    }
  }
};

struct GemmArgs {
  int M, N, K;
  const float* bias;
  const float* colscale;
  int act;
  const float* residual;
  int ld_res;
  float* out_f32;
  int ld_f32;
  __nv_bfloat16* out_bf16;
  int ld_bf16;
  float act_param;          // HOIGEN_ACT_EXP: beta
  // LayerNorm folded into the epilogue (C:443-445 / C:457-458 ln_1 -> in_proj, ln_2 -> c_fc): A = bf16 copy of the RAW
  // stream x, W = W . diag(gamma); out = rstd[row] * (acc - mean[row] * colsum[col]) + bias'[col], colsum = sum_k W'[col][k]
  // (staged where the column scale would be), bias' = bias + W beta.  ln_stats = (mean, rstd) per row from the residual pass.
  const float2* ln_stats;
  // convolution forms (hoigen_gemm_params.conv_taps / halo_* / res_bf16)
  int tap_kb;               // k-blocks per 3x3 tap (0 = plain GEMM): k-block kb reads A columns (kb % tap_kb) * 64 of the
                            // rows shifted by (t / 3 - 1) * halo_w + (t % 3 - 1), t = kb / tap_kb
  int phase_rows;           // stride-2 form (conv_stride = 2): rows per input phase block (= M); 0 = stride 1
  int halo_h, halo_w;       // rows on the ring of each (halo_h, halo_w) image are written as 0 (halo_w = 0: off)
  const __nv_bfloat16* res_bf16;   // added before the activation
  int ld_resb;
  int a2_kb0;               // CTA-pair kernel: k-blocks >= a2_kb0 load their A tile from the second source (tmR); 0 = off
  const __nv_bfloat16* a2;  // (SIMT cross-check only)
  int lda2;
  int debug;  // diagnostics only (HOIGEN_GEMM_DEBUG): 1 = TMA loads without MMAs, 2 = MMAs without TMA loads
  // CTA-pair kernel work split (stream-K): work unit = one k-block of one tile; pair p owns units [bound(p), bound(p+1))
  float* sk_ws;     // fp32 partial-accumulator slots, one per (pair, cta rank)
  int* sk_flags;    // one flag per slot: 1 = partial published
  int sk_snap;      // unit-range boundaries closer than this to a tile edge snap to it
  int sk_tiles;     // the LAST sk_tiles tiles are split at k-block granularity; the others are dealt out round-robin
  int split_k;      // one-CTA kernel: every tile's k-range is cut into split_k units (sk_ws slots + sk_flags tile counters)
};

// A-operand TMA coordinates of k-block kb: plain GEMM (kb * 64, row0) or the row-shifted view of a 3x3 tap
__device__ __forceinline__ void a_coords(const GemmArgs& g, int kb, int row0, int& col, int& row) {
  col = kb * BK; row = row0;
  if (g.tap_kb > 0) {
    const int t = kb / g.tap_kb;
    col = (kb - t * g.tap_kb) * BK;
    const int ky = t / 3, kx = t - ky * 3;
    if (g.phase_rows > 0) row = row0 + ((ky != 1) * 2 + (kx != 1)) * g.phase_rows - (ky == 0) * g.halo_w - (kx == 0);
    else row = row0 + (ky - 1) * g.halo_w + (kx - 1);
  }
}

// true when output row `row` is a halo pixel of its (halo_h, halo_w) image
__device__ __forceinline__ bool halo_row(const GemmArgs& g, int row) {
  if (g.halo_w == 0) return false;
  const int q = row % (g.halo_h * g.halo_w);
  const int y = q / g.halo_w, x = q - y * g.halo_w;
  return x == 0 || x == g.halo_w - 1 || y == 0 || y == g.halo_h - 1;
}

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KiB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (power of two >= 32)
  // [tiles][256 B barriers][2 x 2 x BN floats: double-buffered bias / colscale staging]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * BN * 4;
};

__device__ __forceinline__ float apply_act(float v, int act, float param = 0.f) {
  if (act == HOIGEN_ACT_QUICKGELU) {
    // x * sigmoid(1.702 x)   (CLIP_models_adapter_prior2.py:420); sigmoid(z) = 0.5 + 0.5 tanh(z/2): one MUFU op
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.851f * v));
    return v * fmaf(0.5f, th, 0.5f);
  } else if (act == HOIGEN_ACT_RELU) {
    return fmaxf(v, 0.0f);
  } else if (act == HOIGEN_ACT_EXP) {
    return exp2f(v * (param * 1.4426950408889634f));    // exp(param * v): the Tip-Adapter exp affinity of the per-image cache terms
  }
  return v;
}

// Epilogue-warp helpers.  Thread = one output row; a warp owns CHUNKS x 32 consecutive columns of the tile.
// Latency is the enemy here (few warps, long dependent chains), so everything that does not depend on the MMA is
// fetched early: the tile's bias / column-scale vectors are staged once per tile in shared memory (broadcast LDS
// instead of per-chunk global loads), the fp32 residual of chunk c+1 is prefetched while chunk c is processed (the
// first chunk's before the accumulator is ready), and the TMEM load of chunk c+1 is in flight during chunk c.
constexpr int EPI_THREADS = 256;

// stage bias (0 if absent) and colscale (1 if absent) of columns [col_begin, col_begin + BN) ; tid in [0, 256)
template <int BN>
__device__ __forceinline__ void epilogue_stage_vectors(const GemmArgs& g, int col_begin, float* s_bias, float* s_cs, int tid) {
  for (int i = tid; i < BN; i += EPI_THREADS) {
    const int col = col_begin + i;
    const bool ok = col < g.N;
    s_bias[i] = (g.bias && ok) ? __ldg(g.bias + col) : 0.f;
    s_cs[i] = (g.colscale && ok) ? __ldg(g.colscale + col) : 1.f;     // (LN-folded GEMMs: colscale points at colsum)
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only
}

// Pull one row's residual segment of a tile into L2 ahead of the read-modify-write epilogue (one tile of lead): the
// epilogue's own loads then see L2 latency instead of HBM latency, which its few bytes in flight could not cover.
__device__ __forceinline__ void prefetch_residual_row(const GemmArgs& g, int row, int col_begin, int bn) {
  if (g.residual == nullptr || row >= g.M || col_begin >= g.N || (g.ld_res & 3) != 0) return;
  const int cols = min(bn, g.N - col_begin) & ~3;
  if (cols > 0) prefetch_l2_bulk(g.residual + size_t(row) * g.ld_res + col_begin, uint32_t(cols) * 4u);
}

constexpr int CW = 16;  // columns per epilogue chunk (register budget: two chunks of accumulators + residual in flight)

// out_tile != nullptr: bf16 results go to the 128B-swizzled smem staging tile (64-column panels of 16 KiB, row = rrow)
// for a TMA store instead of per-thread global stores; tile_col = column of this chunk inside the tile.
template <int EPI>
__device__ __forceinline__ void epilogue_process_chunk(const GemmArgs& g, const uint32_t (&r)[CW], const float4 (&res)[CW / 4],
                                                       bool res_vec, const float* res_row, const float* s_bias,
                                                       const float* s_cs, int row, bool row_ok, int col0,
                                                       uint8_t* out_tile = nullptr, int rrow = 0, int tile_col = 0,
                                                       float ln_mean = 0.f, float ln_rstd = 1.f, bool halo = false,
                                                       const uint4* resb = nullptr /* prefetched bf16 identity, 2 x uint4 */) {
  const bool full_chunk = (col0 + CW <= g.N);
  // EPI >= 0: the epilogue recipe is a compile-time constant (act | colscale << 4 | bias << 5 | layernorm << 6 |
  // bf16 identity << 7 | halo zeroing << 8), no work for absent terms
  const int act = EPI >= 0 ? (EPI & 15) : g.act;
  const bool has_ln = EPI >= 0 ? ((EPI >> 6) & 1) != 0 : (g.ln_stats != nullptr);
  const bool has_cs = (EPI >= 0 ? ((EPI >> 4) & 1) != 0 : true) && !has_ln;
  const bool has_bias = EPI >= 0 ? ((EPI >> 5) & 1) != 0 : true;
  float v[CW];
#pragma unroll
  for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
  if (has_ln) {
    // rstd * (x W'^T - mean * colsum) + bias' = fma(rstd, acc, fma(-rstd * mean, colsum, bias')): two FMAs per element
    // (s_cs holds colsum here)
    const float nmr = -ln_mean * ln_rstd;
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(s_cs + j);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_bias) b = *reinterpret_cast<const float4*>(s_bias + j);
      v[j] = fmaf(ln_rstd, v[j], fmaf(nmr, c.x, b.x)); v[j + 1] = fmaf(ln_rstd, v[j + 1], fmaf(nmr, c.y, b.y));
      v[j + 2] = fmaf(ln_rstd, v[j + 2], fmaf(nmr, c.z, b.z)); v[j + 3] = fmaf(ln_rstd, v[j + 3], fmaf(nmr, c.w, b.w));
    }
  } else if (has_bias) {
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(s_bias + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  const bool has_resb = EPI >= 0 ? ((EPI >> 7) & 1) != 0 : (g.res_bf16 != nullptr);
  if (EPI >= 0 && has_resb && out_tile != nullptr) {
    // identity tile already in the staging buffer (TMA-loaded, same swizzled slots this thread is about to overwrite)
    const uint8_t* panel = out_tile + (tile_col >> 6) * 16384;
    const uint32_t ch = uint32_t(tile_col & 63) >> 3;
#pragma unroll
    for (int h = 0; h < CW / 8; ++h) {
      const uint4 q = *reinterpret_cast<const uint4*>(panel + sw128_offset(rrow, ch + h));
      const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[8 * h + 2 * e] += __uint_as_float(w4[e] << 16);
        v[8 * h + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
      }
    }
  } else if (has_resb && row_ok) {   // Bottleneck identity: added before the activation
    const __nv_bfloat16* rp = g.res_bf16 + size_t(row) * g.ld_resb + col0;
    if (full_chunk && (g.ld_resb & 7) == 0) {
#pragma unroll
      for (int h = 0; h < CW / 8; ++h) {
        const uint4 q = resb ? resb[h] : __ldg(reinterpret_cast<const uint4*>(rp) + h);
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[8 * h + 2 * e] += __uint_as_float(w4[e] << 16);
          v[8 * h + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < g.N) v[j] += __bfloat162float(rp[j]);
    }
  }
  if (act != HOIGEN_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = apply_act(v[j], act, g.act_param);
  }
  if ((EPI < 0 || ((EPI >> 8) & 1) != 0) && halo) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = 0.f;
  }
  if (has_cs) {
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      const float4 c = *reinterpret_cast<const float4*>(s_cs + j);
      v[j] *= c.x; v[j + 1] *= c.y; v[j + 2] *= c.z; v[j + 3] *= c.w;
    }
  }
  if (out_tile) {
    uint8_t* panel = out_tile + (tile_col >> 6) * 16384;
    const uint32_t ch = uint32_t(tile_col & 63) >> 3;
    *reinterpret_cast<uint4*>(panel + sw128_offset(rrow, ch)) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(panel + sw128_offset(rrow, ch + 1)) =
        make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
    return;
  }
  if (!row_ok || g.debug == 3) return;
  if (g.residual) {
    if (res_vec && full_chunk) {
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        v[4 * j] += res[j].x; v[4 * j + 1] += res[j].y; v[4 * j + 2] += res[j].z; v[4 * j + 3] += res[j].w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < g.N) v[j] += res_row[col0 + j];
    }
  }
  if (g.out_f32) {
    float* op = g.out_f32 + size_t(row) * g.ld_f32 + col0;
    if (full_chunk && (g.ld_f32 & 3) == 0) {
#pragma unroll
      for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < g.N) op[j] = v[j];
    }
  }
  if (g.out_bf16) {
    __nv_bfloat16* op = g.out_bf16 + size_t(row) * g.ld_bf16 + col0;
    if (full_chunk && (g.ld_bf16 & 7) == 0) {
#pragma unroll
      for (int j = 0; j < CW; j += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(v[j], v[j + 1]);
        pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
        pk.z = pack_bf16x2(v[j + 4], v[j + 5]);
        pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(op + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (col0 + j < g.N) op[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// s_bias / s_cs point at this warp's first column (colbase) inside the staged vectors. COLS = columns per warp.
template <int COLS, int EPI, typename WaitFn>
__device__ __forceinline__ void epilogue_warp(const GemmArgs& g, int row, int colbase, uint32_t t_addr, const float* s_bias,
                                              const float* s_cs, WaitFn wait_acc, uint8_t* out_tile = nullptr, int rrow = 0,
                                              int tile_col0 = 0, const float* part = nullptr, int n_parts = 0,
                                              size_t part_stride = 0, PanelSync ps = PanelSync{0, 0, 0}) {
  constexpr int CHUNKS = COLS / CW;
  const bool row_ok = row < g.M;
  const bool res_vec = g.residual && row_ok && (g.ld_res & 3) == 0;
  const float* res_row = g.residual ? g.residual + size_t(row_ok ? row : 0) * g.ld_res : nullptr;
  float4 res[2][CW / 4];
  uint32_t r[2][CW];
  float ln_mean = 0.f, ln_rstd = 1.f;
  if ((EPI >= 0 ? ((EPI >> 6) & 1) != 0 : g.ln_stats != nullptr) && row_ok) {
    const float2 st = __ldg(g.ln_stats + row);
    ln_mean = st.x; ln_rstd = st.y;
  }
  const bool halo = (EPI < 0 || ((EPI >> 8) & 1) != 0) && halo_row(g, row);
  // bf16 identity of the convolution epilogues: chunk c + 1's 32 bytes are fetched while chunk c is processed
  // (compile-time recipes of the TMA-store kernels get the identity tile by TMA into the staging buffer instead: see RES_TMA)
  const bool resb_vec = (EPI >= 0 ? (((EPI >> 7) & 1) != 0 && out_tile == nullptr) : g.res_bf16 != nullptr) && row_ok && (g.ld_resb & 7) == 0;
  const __nv_bfloat16* resb_row = resb_vec ? g.res_bf16 + size_t(row) * g.ld_resb : nullptr;
  uint4 resb[2][CW / 8];
  if (resb_vec && colbase + CW <= g.N) {
#pragma unroll
    for (int h = 0; h < CW / 8; ++h) resb[0][h] = __ldg(reinterpret_cast<const uint4*>(resb_row + colbase) + h);
  }
  if (res_vec && colbase + CW <= g.N) {
#pragma unroll
    for (int j = 0; j < CW / 4; ++j) res[0][j] = *reinterpret_cast<const float4*>(res_row + colbase + 4 * j);
  }
  wait_acc();
  if (g.debug == 4) return;
  const int last_panel = (tile_col0 + COLS - 1) >> 6;
  if (colbase >= g.N) {  // warp-uniform: nothing to compute for this half; its panels are handed over untouched
    for (int p = tile_col0 >> 6; p <= last_panel; ++p) { ps.enter(p); ps.leave(p); }
    return;
  }
  int cur_panel = -1;
  tmem_ld_32x32b_x16(t_addr, r[0]);
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) {
    const int col0 = colbase + c * CW;
    const int pc = (tile_col0 + c * CW) >> 6;
    if (pc != cur_panel) {               // first chunk of a staging panel: the previous one is complete, this one must be free
      if (cur_panel >= 0) ps.leave(cur_panel);
      ps.enter(pc);
      cur_panel = pc;
    }
    tmem_wait_ld();   // chunk c is in registers
    const bool more = (c + 1 < CHUNKS) && (col0 + CW < g.N);
    if (more) {
      tmem_ld_32x32b_x16(t_addr + uint32_t((c + 1) * CW), r[(c + 1) & 1]);   // in flight while chunk c is processed
      if (res_vec && col0 + 2 * CW <= g.N) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j)
          res[(c + 1) & 1][j] = *reinterpret_cast<const float4*>(res_row + col0 + CW + 4 * j);
      }
      if (resb_vec && col0 + 2 * CW <= g.N) {
#pragma unroll
        for (int h = 0; h < CW / 8; ++h) resb[(c + 1) & 1][h] = __ldg(reinterpret_cast<const uint4*>(resb_row + col0 + CW) + h);
      }
    }
    if (n_parts > 0) {   // stream-K fix-up: add the other pairs' partial accumulators of this tile (fixed order)
      const float* pp = part + size_t(c) * (BM * CW);
      for (int q = 0; q < n_parts; ++q, pp += part_stride) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(pp) + j);
          uint32_t* rr = r[c & 1] + 4 * j;
          rr[0] = __float_as_uint(__uint_as_float(rr[0]) + a.x); rr[1] = __float_as_uint(__uint_as_float(rr[1]) + a.y);
          rr[2] = __float_as_uint(__uint_as_float(rr[2]) + a.z); rr[3] = __float_as_uint(__uint_as_float(rr[3]) + a.w);
        }
      }
    }
    epilogue_process_chunk<EPI>(g, r[c & 1], res[c & 1], res_vec, res_row, s_bias + c * CW, s_cs + c * CW, row, row_ok, col0,
                           out_tile, rrow, tile_col0 + c * CW, ln_mean, ln_rstd, halo,
                           (resb_vec && col0 + CW <= g.N) ? resb[c & 1] : nullptr);
    if (!more) break;
  }
  tmem_wait_ld();
  ps.leave(cur_panel);
  for (int p = cur_panel + 1; p <= last_panel; ++p) { ps.enter(p); ps.leave(p); }   // panels past the matrix edge
}

// Split-K finisher: the accumulator of a tile is the sum of its n_parts partial slots (fixed order -> bit-reproducible),
// then the normal epilogue.  Slot layout as written by epilogue_dump_partial: [16-column chunk][128 rows][16 floats].
template <int COLS, int EPI>
__device__ __forceinline__ void epilogue_from_parts(const GemmArgs& g, int row, int colbase, const float* s_bias, const float* s_cs,
                                                    const float* part, int n_parts, size_t part_stride) {
  const bool row_ok = row < g.M;
  const bool res_vec = g.residual && row_ok && (g.ld_res & 3) == 0;
  const float* res_row = g.residual ? g.residual + size_t(row_ok ? row : 0) * g.ld_res : nullptr;
  if (colbase >= g.N) return;
#pragma unroll 1
  for (int c = 0; c < COLS / CW; ++c) {
    const int col0 = colbase + c * CW;
    if (col0 >= g.N) break;
    float acc[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) acc[j] = 0.f;
    const float* pp = part + size_t(c) * (BM * CW);
    for (int q = 0; q < n_parts; ++q, pp += part_stride) {
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(pp) + j);
        acc[4 * j] += a.x; acc[4 * j + 1] += a.y; acc[4 * j + 2] += a.z; acc[4 * j + 3] += a.w;
      }
    }
    uint32_t r[CW];
    float4 res[CW / 4];
#pragma unroll
    for (int j = 0; j < CW; ++j) r[j] = __float_as_uint(acc[j]);
    if (res_vec && col0 + CW <= g.N) {
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) res[j] = *reinterpret_cast<const float4*>(res_row + col0 + 4 * j);
    }
    float ln_mean = 0.f, ln_rstd = 1.f;
    if (g.ln_stats != nullptr && row_ok) {
      const float2 st = __ldg(g.ln_stats + row);
      ln_mean = st.x; ln_rstd = st.y;
    }
    epilogue_process_chunk<EPI>(g, r, res, res_vec, res_row, s_bias + c * CW, s_cs + c * CW, row, row_ok, col0, nullptr, 0, 0,
                                ln_mean, ln_rstd, EPI < 0 && halo_row(g, row));
  }
}

template <int COLS>
__device__ __forceinline__ void epilogue_dump_partial(uint32_t t_addr, float* slot);

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* tiles = smem_raw + (tiles_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::STAGE_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, then tmem base slot
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8u * STAGES;
  const uint32_t bar_tfull = bar_full + 16u * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (g.M + BM - 1) / BM;
  const int num_n = (g.N + BN - 1) / BN;
  const int num_k = (g.K + BK - 1) / BK;
  // work unit = (tile, k-split): unit u -> tile u / S, k-blocks [sp * num_k / S, (sp + 1) * num_k / S)
  const int S = g.split_k;
  const int num_tiles = num_m * num_n * S;    // units (== tiles when S == 1)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8u * a, 1);
      mbar_init(bar_tempty + 8u * a, 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x) {
        const int tile = unit / S, sp = unit - tile * S;
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = sp * num_k / S; kb < (sp + 1) * num_k / S; ++kb) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t a_dst = tiles_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + Cfg::A_BYTES;
          const uint32_t full = bar_full + 8u * stage;
          mbar_arrive_expect_tx(full, Cfg::STAGE_BYTES);
          int a_col, a_row;
          a_coords(g, kb, m_blk * BM, a_col, a_row);
          tma_load_2d(a_dst, &tmA, full, a_col, a_row);
          tma_load_2d(b_dst, &tmB, full, kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
        const int sp = unit % S;
        const int kb0 = sp * num_k / S, kb1 = (sp + 1) * num_k / S;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = tiles_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_sdesc_sw128(a_addr + k * (UMMA_K * 2));
            const uint64_t bdesc = make_sdesc_sw128(b_addr + k * (UMMA_K * 2));
            umma_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(bar_empty + 8u * stage);  // frees this smem stage once the MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(bar_tfull + 8u * acc);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue warps (8: two per TMEM lane quadrant, each owning half of the BN columns) ======
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;          // 0 / 1: which half of the tile's columns
    int it = 0;
    int* s_last = reinterpret_cast<int*>(tiles + STAGES * Cfg::STAGE_BYTES + 192);   // split-K: "this CTA finishes the tile"
    constexpr size_t SLOT = size_t(BN) * BM;
    const size_t slot_off = (size_t(half * (BN / 2) / CW) * BM + size_t(quad * 32 + lane)) * CW;
    for (int unit = blockIdx.x; unit < num_tiles; unit += gridDim.x, ++it) {
      const int tile = unit / S;
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int row = m_blk * BM + quad * 32 + lane;
      const int colbase = n_blk * BN + half * (BN / 2);
      const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * BN + half * (BN / 2));
      float* s_bias = reinterpret_cast<float*>(tiles + STAGES * Cfg::STAGE_BYTES + 256) + (it & 1) * 2 * BN;
      float* s_cs = s_bias + BN;
      epilogue_stage_vectors<BN>(g, n_blk * BN, s_bias, s_cs, threadIdx.x - 64);
      if (S == 1) {
        epilogue_warp<BN / 2, -1>(g, row, colbase, t_addr, s_bias + half * (BN / 2), s_cs + half * (BN / 2), [&]() {
          mbar_wait(bar_tfull + 8u * acc, acc_phase);
          tc_fence_after();
        });
        // release the accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
        continue;
      }
      // ---- split-K: publish this unit's raw accumulator; whoever arrives last at the tile's counter sums all of the
      //      tile's slots in split order (so the result does not depend on who that is) and runs the epilogue ----
      mbar_wait(bar_tfull + 8u * acc, acc_phase);
      tc_fence_after();
      epilogue_dump_partial<BN / 2>(t_addr, g.sk_ws + size_t(unit) * SLOT + slot_off);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 64) *s_last = (atomicAdd(g.sk_flags + tile, 1) == S - 1) ? 1 : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (*s_last) {
        __threadfence();
        epilogue_from_parts<BN / 2, -1>(g, row, colbase, s_bias + half * (BN / 2), s_cs + half * (BN / 2),
                                        g.sk_ws + size_t(tile) * S * SLOT + slot_off, S, SLOT);
        if (threadIdx.x == 64) g.sk_flags[tile] = 0;      // re-arm for the next launch
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // s_last is rewritten by the next unit
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes one 256 x BN tile. Each CTA TMA-loads its own
// 128 rows of A and HALF of the W tile (BN/2 rows) per k-block, the leader CTA's single MMA thread issues
// UMMA 256 x BN x 16 over both CTAs' shared memory, and every CTA drains its own 128 x BN half of the accumulator
// from its own TMEM.  Against the one-CTA kernel this cuts the L2 -> SM operand traffic per FLOP by 1.5x
// (BN = 256), which is what bounds the one-CTA kernel on these shapes, and frees smem for a deeper TMA ring.
//   full[s]      (leader's)  2 arrivals: leader producer arrive.expect_tx(bytes of BOTH CTAs) + peer producer arrive;
//                            both CTAs' TMA loads credit their bytes to the leader's barrier
//   empty[s]     (per CTA)   tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a] (per CTA)   tcgen05.commit multicast
//   tmem_empty[a](leader's)  16 arrivals: 8 epilogue warps of each CTA
// ---------------------------------------------------------------------------------------------
// TMA_OUT: bf16-only outputs are staged in a swizzled smem tile and written with TMA stores (full-line writes, and
// the TMEM accumulator is released as soon as it has been read, not when the global stores have been issued).
// Stream-K work split of the CTA-pair kernel.  Wave quantisation is what the data-parallel schedule loses on the
// encoder's shapes (12608 rows = 49.25 row panels: 150 / 450 / 600 tiles on 74 pairs = 2.03 / 6.08 / 8.1 rounds), so
// the unit of work is one k-block of one tile and pair p owns the contiguous unit range [sk_bound(p), sk_bound(p+1))
// of the tile-major order.  A range therefore is: [tail of a tile] [whole tiles] [head of a tile].
//   tail (k-blocks [kb0 > 0, ...))   processed FIRST: the pair dumps its raw fp32 accumulator into its slot and
//                                    publishes a flag (contributor);
//   head (k-blocks [0, kb1 < num_k)) processed LAST: the pair waits for the flags of the pairs that hold the rest of the
//                                    tile (they produced it at the start of their ranges), adds their slots in pair
//                                    order (deterministic) and runs the normal epilogue (finisher), then clears the flags.
// All pairs are co-resident (grid <= SM pairs, one CTA per SM), so the wait cannot deadlock.
__device__ __forceinline__ int sk_bound(int p, int P, int total, int num_k, int snap) {
  if (p >= P) return total;
  int b = int((long long)total * p / P);
  const int r = b % num_k;
  if (r < snap) b -= r;
  else if (num_k - r <= snap) b += num_k - r;
  return b;
}

// Per-pair work sequence: first the pair's unit range of the split region (so its fix-up traffic hides behind the
// whole tiles that follow), then whole tiles dealt round-robin (neighbouring pairs share operand panels in L2).
struct PairWork {
  int u, u_end, num_k, sk_tile0, tile_dp, dp_end, stride;
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (u < u_end) {
      const int t = u / num_k;
      kb0 = u - t * num_k;
      kb1 = min(num_k, kb0 + (u_end - u));
      u += kb1 - kb0;
      tile = sk_tile0 + t;
      return true;
    }
    if (tile_dp < dp_end) {
      tile = tile_dp; tile_dp += stride; kb0 = 0; kb1 = num_k;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// contributor epilogue: raw accumulator -> slot, layout [chunk of 16 columns][128 rows][16 floats] (a warp writes 2 KiB
// contiguous per chunk).  COLS = columns this warp owns; slot points at this thread's row of the warp's first chunk.
template <int COLS>
__device__ __forceinline__ void epilogue_dump_partial(uint32_t t_addr, float* slot) {
  uint32_t r[2][CW];
  tmem_ld_32x32b_x16(t_addr, r[0]);
#pragma unroll
  for (int c = 0; c < COLS / CW; ++c) {
    tmem_wait_ld();
    if (c + 1 < COLS / CW) tmem_ld_32x32b_x16(t_addr + uint32_t((c + 1) * CW), r[(c + 1) & 1]);
    float4* dst = reinterpret_cast<float4*>(slot + size_t(c) * (BM * CW));
#pragma unroll
    for (int j = 0; j < CW / 4; ++j)
      __stcg(dst + j, make_float4(__uint_as_float(r[c & 1][4 * j]), __uint_as_float(r[c & 1][4 * j + 1]),
                                  __uint_as_float(r[c & 1][4 * j + 2]), __uint_as_float(r[c & 1][4 * j + 3])));
  }
}

template <int BN, bool TMA_OUT>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KiB: this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * BK * 2;    // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
#ifdef HOIGEN_EXP_OUT_BYTES
  static constexpr int OUT_BYTES = HOIGEN_EXP_OUT_BYTES;   // experiment only (epilogue disabled): deeper operand ring
#else
  static constexpr int OUT_BYTES = TMA_OUT ? (BN / 64) * 16384 : 0;
#endif
  static constexpr int STAGES_FIT = (196608 - OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr int ACC_STRIDE = 256;               // TMEM columns between the two accumulator stages
  static constexpr int TMEM_COLS = 512;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES + OUT_BYTES;
  static constexpr int BAR_BYTES = 384;                // mbarriers + TMEM slot, then the bias / column-scale staging
  static constexpr int SMEM_BYTES = BAR_OFF + 1024 + BAR_BYTES + 4 * BN * 4;
};

template <int BN, bool TMA_OUT, int EPI>
// HOIGEN_GEMM2_REGCAP_THREADS (512) declares a larger thread bound than the 352 the kernel is launched with: the compiler then
// keeps it within 64 Ki / 512 = 128 registers (no spills in the encoder's recipes), so that a 128-thread CTA of the HBM-bound
// residual + LayerNorm pass of the OTHER batch in flight fits on the same SM (352 x 128 + 128 x 128 <= 64 Ki registers).
#ifndef HOIGEN_GEMM2_REGCAP_THREADS
#define HOIGEN_GEMM2_REGCAP_THREADS GEMM2_THREADS
#endif
__global__ void __launch_bounds__(HOIGEN_GEMM2_REGCAP_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, GemmArgs g) {
  // convolution recipe with a bf16 identity: its tile is TMA-loaded into the output staging buffer and updated in place
  constexpr bool RES_TMA = TMA_OUT && EPI >= 0 && ((EPI >> 7) & 1) != 0;
  using Cfg = Gemm2Cfg<BN, TMA_OUT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (tiles_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + Cfg::BAR_OFF);
  uint8_t* out_stage = tiles + STAGES * Cfg::STAGE_BYTES;   // TMA_OUT staging tile (1024-aligned)
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8u * STAGES;
  const uint32_t bar_tfull = bar_full + 16u * STAGES;
  const uint32_t bar_tempty = bar_tfull + 16u;
  const uint32_t bar_ready = bar_full + 256u;     // [BN / 64] staging panel may be written   (PanelSync)
  const uint32_t bar_written = bar_full + 288u;   // [BN / 64] staging panel has been written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (g.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (g.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (g.K + BK - 1) / BK;
  const int total_units = g.sk_tiles * num_k;   // of the split region = tiles [num_tiles - sk_tiles, num_tiles)
  PairWork work0;
  work0.u = sk_bound(cluster_id, num_clusters, total_units, num_k, g.sk_snap);
  work0.u_end = sk_bound(cluster_id + 1, num_clusters, total_units, num_k, g.sk_snap);
  work0.num_k = num_k; work0.sk_tile0 = num_tiles - g.sk_tiles;
  work0.tile_dp = cluster_id; work0.dp_end = num_tiles - g.sk_tiles; work0.stride = num_clusters;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8u * s, 2);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8u * a, 1);
      mbar_init(bar_tempty + 8u * a, 16);
    }
    if (TMA_OUT) {
      // threads writing staging panel p: 128 per column half that reaches into it
      for (int p = 0; p < BN / 64; ++p) {
        int halves = 0;
        for (int h = 0; h < 2; ++h) halves += ((h * (BN / 2)) >> 6) <= p && p <= ((h * (BN / 2) + BN / 2 - 1) >> 6) ? 1 : 0;
        mbar_init(bar_ready + 8u * p, 1);
        mbar_init(bar_written + 8u * p, 128u * halves);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible to the peer before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      PairWork work = work0;
      for (int tile, kb0, kb1; work.next(tile, kb0, kb1);) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
          const uint32_t a_dst = tiles_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t b_dst = a_dst + Cfg::A_BYTES;
          const uint32_t full = bar_full + 8u * stage;
          if (g.debug == 2) {   // diagnostics: no loads, just hand the stage over
            if (rank == 0) mbar_arrive(full); else mbar_arrive_leader(full);
          } else {
            if (rank == 0) mbar_arrive_expect_tx(full, 2 * Cfg::STAGE_BYTES);
            else mbar_arrive_leader(full);
            int a_col, a_row;
            a_coords(g, kb, m_blk * (2 * BM) + int(rank) * BM, a_col, a_row);
            if (g.a2_kb0 > 0 && kb >= g.a2_kb0) tma_load_2d_2sm(a_dst, &tmR, full, (kb - g.a2_kb0) * BK, a_row);   // second A source
            else tma_load_2d_2sm(a_dst, &tmA, full, a_col, a_row);
            tma_load_2d_2sm(b_dst, &tmB, full, kb * BK, n_blk * BN + int(rank) * (BN / 2));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      PairWork work = work0;
      for (int tile, kb0, kb1; work.next(tile, kb0, kb1); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * Cfg::ACC_STRIDE);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8u * stage, phase);
          tc_fence_after();
          const uint32_t a_addr = tiles_addr + stage * Cfg::STAGE_BYTES;
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          if (g.debug != 1) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adesc = make_sdesc_sw128(a_addr + k * (UMMA_K * 2));
              const uint64_t bdesc = make_sdesc_sw128(b_addr + k * (UMMA_K * 2));
              umma_bf16_ss_2sm(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit_2sm(bar_empty + 8u * stage);   // frees the stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2sm(bar_tfull + 8u * acc);       // accumulator halves complete in BOTH CTAs
      }
    }
  } else if (warp < 10) {
    // ===================== epilogue warps (both CTAs: own 128 rows) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rrow = quad * 32 + lane;
    constexpr size_t SLOT = size_t(BN) * BM;                 // floats per (pair, rank) slot
    // this thread's row inside the warp's first chunk of a slot
    const size_t slot_off = (size_t(half * (BN / 2) / CW) * BM + rrow) * CW;
    int it = 0;
    int res_it = 0;   // tiles this thread has run the epilogue of (phase of the staging-panel barriers)
    PairWork work = work0;
    for (int tile, kb0, kb1; work.next(tile, kb0, kb1); ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const uint32_t t_addr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * Cfg::ACC_STRIDE + half * (BN / 2));
      if (kb0 > 0) {
        // ---- contributor: tail of a tile another pair finishes ----
        mbar_wait(bar_tfull + 8u * acc, acc_phase);
        tc_fence_after();
        if (g.debug != 8) epilogue_dump_partial<BN / 2>(t_addr, g.sk_ws + (size_t(cluster_id) * 2 + rank) * SLOT + slot_off);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar_tempty + 8u * acc);
        if (g.debug != 8) __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 64) st_release_gpu(g.sk_flags + cluster_id * 2 + int(rank), 1);
        continue;
      }
      int n_parts = 0;
      if (kb1 < num_k) {
        // ---- finisher of a split tile: the pairs after this one hold the rest ----
        const int tile_end = (tile - work0.sk_tile0 + 1) * num_k;
        for (int q = cluster_id + 1; q < num_clusters && sk_bound(q, num_clusters, total_units, num_k, g.sk_snap) < tile_end; ++q) {
          const int* flag = g.sk_flags + q * 2 + int(rank);
          uint32_t spins = 0;
          while (g.debug != 7 && ld_acquire_gpu(flag) != 1) {
            __nanosleep(64);
            if (++spins > (1u << 24)) __trap();   // a contributor that never arrives is a scheduling bug, not a wait
          }
          ++n_parts;
        }
      }
      const int row = m_blk * (2 * BM) + int(rank) * BM + rrow;
      const int colbase = n_blk * BN + half * (BN / 2);
      float* s_bias = reinterpret_cast<float*>(tiles + Cfg::BAR_OFF + Cfg::BAR_BYTES) + (it & 1) * 2 * BN;
      float* s_cs = s_bias + BN;
      epilogue_stage_vectors<BN>(g, n_blk * BN, s_bias, s_cs, threadIdx.x - 64);   // (contains the epilogue-wide barrier)
      epilogue_warp<BN / 2, EPI>(g, row, colbase, t_addr, s_bias + half * (BN / 2), s_cs + half * (BN / 2), [&]() {
        mbar_wait(bar_tfull + 8u * acc, acc_phase);
        tc_fence_after();
      }, TMA_OUT ? out_stage : nullptr, rrow, half * (BN / 2),
      n_parts ? g.sk_ws + (size_t(cluster_id + 1) * 2 + rank) * SLOT + slot_off : nullptr,
      (g.debug == 6 || g.debug == 7) ? 0 : n_parts, 2 * SLOT,
      TMA_OUT ? PanelSync{bar_ready, bar_written, uint32_t(res_it) & 1u} : PanelSync{0, 0, 0});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_tempty + 8u * acc);   // TMEM drained: the next-but-one tile's MMAs may start
      if (n_parts) asm volatile("bar.sync 1, 256;" ::: "memory");
      if (n_parts && threadIdx.x == 64)   // every epilogue thread is past its slot reads: re-arm the flags
        for (int q = 1; q <= n_parts; ++q) g.sk_flags[(cluster_id + q) * 2 + int(rank)] = 0;
      ++res_it;
    }
  } else if (TMA_OUT && warp == 10 && lane == 0 && g.debug != 4) {
    // ===================== output DMA (both CTAs: own staging panels) =====================
    // Per finished tile: as each 64-column panel is handed over, TMA-store it; then, as each store has been read out of
    // shared memory, hand the panel back for the next tile -- after TMA-loading that tile's bf16 identity into it for the
    // convolution recipes (the epilogue then updates the panel in place).  The epilogue warps never wait for a store.
    constexpr int NP = BN / 64;
    // panels in the order they are completed: with four, the two column halves finish their first panels together
    constexpr int ORDER[4] = {0, NP == 4 ? 2 : 1, NP == 4 ? 1 : 2, 3};
    PairWork work = work0;
    int tile = 0, kb0 = 0, kb1 = 0;
    auto next_tile = [&]() {             // tiles this pair runs the epilogue of (stream-K contributions have none)
      while (work.next(tile, kb0, kb1))
        if (kb0 == 0) return true;
      return false;
    };
    auto hand_back = [&](int p) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      if (RES_TMA && n_blk * BN + p * 64 < g.N) {
        mbar_arrive_expect_tx(bar_ready + 8u * p, 16384u);
        tma_load_2d(smem_u32(out_stage) + p * 16384, &tmR, bar_ready + 8u * p, n_blk * BN + p * 64,
                    m_blk * (2 * BM) + int(rank) * BM);
      } else {
        mbar_arrive(bar_ready + 8u * p);
      }
    };
    bool have = next_tile();
    if (have)
      for (int p = 0; p < NP; ++p) hand_back(p);
    for (uint32_t n_ep = 0; have; ++n_ep) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row0 = m_blk * (2 * BM) + int(rank) * BM;
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const int p = ORDER[j];
        mbar_wait(bar_written + 8u * p, n_ep & 1u);
        if (n_blk * BN + p * 64 < g.N && g.debug != 5)   // 5: whole epilogue but no global writes
          tma_store_2d(&tmC, smem_u32(out_stage) + p * 16384, n_blk * BN + p * 64, row0);
        tma_store_commit();                              // one bulk group per panel, empty or not
      }
      have = next_tile();
      if (have) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          // groups complete in order: at most NP - 1 - j younger ones pending <=> the store of ORDER[j] has been read out
          if (NP - 1 - j == 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
          else if (NP - 1 - j == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          else if (NP - 1 - j == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          hand_back(ORDER[j]);
        }
      }
    }
    tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the peer can still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// test-only SIMT cross-check
// ---------------------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ w, int lda,
                                 int ldw, GemmArgs g) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= g.N || m >= g.M) return;
  float acc = 0.f;
  if (g.tap_kb > 0) {
    const int cin = g.tap_kb * BK;
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      const long long src = g.phase_rows > 0
                                ? (long long)m + (long long)((ky != 1) * 2 + (kx != 1)) * g.phase_rows - (ky == 0) * g.halo_w - (kx == 0)
                                : (long long)m + (ky - 1) * g.halo_w + (kx - 1);
      if (src < 0 || src >= (g.phase_rows > 0 ? 4LL * g.M : (long long)g.M)) continue;
      for (int k = 0; k < cin; ++k)
        acc = fmaf(__bfloat162float(a[size_t(src) * lda + k]), __bfloat162float(w[size_t(n) * ldw + t * cin + k]), acc);
    }
  } else {
    const int k1 = g.a2_kb0 > 0 ? g.a2_kb0 * BK : g.K;
    for (int k = 0; k < k1; ++k)
      acc = fmaf(__bfloat162float(a[size_t(m) * lda + k]), __bfloat162float(w[size_t(n) * ldw + k]), acc);
    for (int k = k1; k < g.K; ++k)
      acc = fmaf(__bfloat162float(g.a2[size_t(m) * g.lda2 + k - k1]), __bfloat162float(w[size_t(n) * ldw + k]), acc);
  }
  if (g.ln_stats) acc = g.ln_stats[m].y * (acc - g.ln_stats[m].x * g.colscale[n]);
  if (g.bias) acc += g.bias[n];
  if (g.res_bf16) acc += __bfloat162float(g.res_bf16[size_t(m) * g.ld_resb + n]);
  acc = apply_act(acc, g.act, g.act_param);
  if (halo_row(g, m)) acc = 0.f;
  if (g.colscale && !g.ln_stats) acc *= g.colscale[n];
  if (g.residual) acc += g.residual[size_t(m) * g.ld_res + n];
  if (g.out_f32) g.out_f32[size_t(m) * g.ld_f32 + n] = acc;
  if (g.out_bf16) g.out_bf16[size_t(m) * g.ld_bf16 + n] = __float2bfloat16_rn(acc);
}

static int validate(const hoigen_gemm_params* p) {
  HOIGEN_CHECK_ARG(p != nullptr, "gemm: null params");
  HOIGEN_CHECK_ARG(p->a && p->w, "gemm: null operand");
  HOIGEN_CHECK_ARG(p->M > 0 && p->N > 0 && p->K > 0, "gemm: bad shape M=%d N=%d K=%d", p->M, p->N, p->K);
  HOIGEN_CHECK_ARG((p->conv_taps == 9 || p->lda >= (p->a2 ? p->K - p->k2 : p->K)) && p->ldw >= p->K, "gemm: lda/ldw < K");
  HOIGEN_CHECK_ARG((p->lda % 8) == 0 && (p->ldw % 8) == 0, "gemm: lda/ldw must be multiples of 8 (got %d, %d)",
                   p->lda, p->ldw);
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(p->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w) & 15) == 0,
                   "gemm: operands must be 16-byte aligned");
  HOIGEN_CHECK_ARG(p->out_f32 || p->out_bf16, "gemm: no output");
  HOIGEN_CHECK_ARG(!p->out_f32 || p->ld_f32 >= p->N, "gemm: ld_f32 < N");
  HOIGEN_CHECK_ARG(!p->out_bf16 || p->ld_bf16 >= p->N, "gemm: ld_bf16 < N");
  HOIGEN_CHECK_ARG(!p->residual || p->ld_res >= p->N, "gemm: ld_res < N");
  HOIGEN_CHECK_ARG(p->conv_taps == 0 || p->conv_taps == 1 || p->conv_taps == 9, "gemm: conv_taps must be 0, 1 or 9 (got %d)", p->conv_taps);
  HOIGEN_CHECK_ARG(p->conv_stride == 0 || p->conv_stride == 1 || (p->conv_stride == 2 && p->conv_taps == 9 && p->M <= (1 << 28)),
                   "gemm: conv_stride must be 0 / 1, or 2 together with conv_taps = 9 (got %d)", p->conv_stride);
  HOIGEN_CHECK_ARG(p->conv_taps != 9 || (p->conv_cin > 0 && p->conv_cin % 64 == 0 && p->K == 9 * p->conv_cin && p->halo_w > 0 && p->lda >= p->conv_cin),
                   "gemm: 3x3 form needs conv_cin %% 64 == 0, K == 9 * conv_cin, halo_w > 0 (cin=%d K=%d)", p->conv_cin, p->K);
  HOIGEN_CHECK_ARG(p->halo_w == 0 || (p->halo_h >= 3 && p->halo_w >= 3 && p->M % (p->halo_h * p->halo_w) == 0),
                   "gemm: M must be a whole number of (halo_h, halo_w) images");
  HOIGEN_CHECK_ARG(!p->res_bf16 || (p->ld_resb >= p->N && !p->ln_stats), "gemm: bad res_bf16 layout");
  HOIGEN_CHECK_ARG(!p->a2 || (p->k2 > 0 && p->k2 < p->K && (p->K - p->k2) % 64 == 0 && p->lda2 >= p->k2 && (p->lda2 % 8) == 0 &&
                              (reinterpret_cast<uintptr_t>(p->a2) & 15) == 0 && !p->res_bf16 && p->conv_taps != 9 &&
                              (p->block_n == 0 || p->block_n > 2000) && p->M > 128),
                   "gemm: second A source needs K - k2 %% 64 == 0, CTA-pair tiles, no res_bf16 / 3x3 form (K=%d k2=%d)", p->K, p->k2);
  HOIGEN_CHECK_ARG(p->act >= 0 && p->act <= 3, "gemm: bad act %d", p->act);
  HOIGEN_CHECK_ARG(!p->ln_stats || (p->ln_colsum && !p->colscale), "gemm: ln_stats needs ln_colsum and excludes colscale");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(p->ln_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(p->ln_colsum) & 15) == 0,
                   "gemm: ln_stats / ln_colsum alignment");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(p->bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->colscale) & 15) == 0,
                   "gemm: bias/colscale must be 16-byte aligned");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(p->residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out_f32) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(p->out_bf16) & 15) == 0,
                   "gemm: residual/outputs must be 16-byte aligned");
  return HOIGEN_OK;
}

static GemmArgs to_args(const hoigen_gemm_params* p) {
  GemmArgs g;
  g.M = p->M; g.N = p->N; g.K = p->K;
  g.bias = p->bias; g.colscale = p->ln_stats ? p->ln_colsum : p->colscale; g.act = p->act;
  g.act_param = p->act_param;
  g.ln_stats = reinterpret_cast<const float2*>(p->ln_stats);
  g.residual = p->residual; g.ld_res = p->ld_res;
  g.out_f32 = p->out_f32; g.ld_f32 = p->ld_f32;
  g.out_bf16 = reinterpret_cast<__nv_bfloat16*>(p->out_bf16); g.ld_bf16 = p->ld_bf16;
  static const int dbg = getenv("HOIGEN_GEMM_DEBUG") ? atoi(getenv("HOIGEN_GEMM_DEBUG")) : 0;
  g.debug = dbg;
  g.sk_ws = nullptr; g.sk_flags = nullptr; g.sk_snap = 0; g.sk_tiles = 0; g.split_k = 1;
  g.tap_kb = p->conv_taps == 9 ? p->conv_cin / BK : 0;
  g.phase_rows = (p->conv_taps == 9 && p->conv_stride == 2) ? p->M : 0;
  g.halo_h = p->halo_h; g.halo_w = p->halo_w;
  g.res_bf16 = reinterpret_cast<const __nv_bfloat16*>(p->res_bf16); g.ld_resb = p->ld_resb;
  g.a2_kb0 = p->a2 ? (p->K - p->k2) / BK : 0;
  g.a2 = reinterpret_cast<const __nv_bfloat16*>(p->a2); g.lda2 = p->lda2;
  return g;
}

// interned "gemm<1|2>_n<N>_k<K>" tags for the launch profiler (1 = one-CTA kernel, 2 = CTA-pair kernel)
static const char* gemm_tag(int N, int K, int kind = 1) {
  static std::map<std::pair<std::pair<int, int>, int>, std::string> tags;   // node-based: c_str() of an entry stays valid
  static std::mutex mu;                                                       // callers may launch from several host threads
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(std::make_pair(N, K), kind);
  auto it = tags.find(key);
  if (it == tags.end())
    it = tags.emplace(key, "gemm" + std::to_string(kind) + "_n" + std::to_string(N) + "_k" + std::to_string(K)).first;
  return it->second.c_str();
}

// Split-K of the one-CTA kernel: a long-K GEMM with few tiles (the M = 64 per-image cache terms: 2 tiles x 64 k-blocks)
// is bound by ONE SM's TMA service rate (~0.4 us per 24 KiB k-block); cutting every tile's k-range over idle SMs is the
// only parallelism there is.  Only when all units still fit one wave and every unit keeps >= 8 k-blocks.
static int auto_split_k(int num_tiles, int num_k) {
  if (num_k < 32 || num_tiles >= num_sms()) return 1;
  int s = num_sms() / num_tiles;
  if (s > num_k / 8) s = num_k / 8;
  if (s > 8) s = 8;
  return s < 2 ? 1 : s;
}

template <int BN>
static int launch_gemm(const hoigen_gemm_params* p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(gemm_bf16_kernel<BN>), Cfg::SMEM_BYTES));
  const CUtensorMap* ta = get_tmap_2d_bf16(p->a, uint64_t(p->conv_taps == 9 ? p->conv_cin : (p->a2 ? p->K - p->k2 : p->K)), uint64_t(p->M) * (p->conv_taps == 9 && p->conv_stride == 2 ? 4 : 1), uint64_t(p->lda) * 2, BK, BM);
  if (!ta) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tb = get_tmap_2d_bf16(p->w, uint64_t(p->K), uint64_t(p->N), uint64_t(p->ldw) * 2, BK, BN);
  if (!tb) return HOIGEN_ERR_CUDA;
  const int num_tiles = ((p->M + BM - 1) / BM) * ((p->N + BN - 1) / BN);
  const int num_k = (p->K + BK - 1) / BK;
  GemmArgs args = to_args(p);
  int S = p->split_k > 0 ? p->split_k : auto_split_k(num_tiles, num_k);
  if (S > num_k) S = num_k;
  if (S > 1) {
    StreamKWorkspace ws = get_streamk_workspace(stream, size_t(num_tiles) * S * BN * BM * sizeof(float), num_tiles);
    if (!ws.slots) return HOIGEN_ERR_CUDA;
    args.sk_ws = ws.slots; args.sk_flags = ws.flags; args.split_k = S;
  }
  const int units = num_tiles * S;
  const int grid = units < num_sms() ? units : num_sms();
  KernelScope ks(gemm_tag(p->N, p->K), stream, 2.0 * p->M * p->N * p->K,
                 2.0 * (double(p->M) * p->K + double(p->N) * p->K) + double(p->M) * p->N * ((p->out_f32 ? 4 : 0) + (p->out_bf16 ? 2 : 0) + (p->residual ? 4 : 0)));
  gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(*ta, *tb, args);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

template <int BN, bool TMA_OUT, int EPI>
static int launch_gemm2_impl(const hoigen_gemm_params* p, cudaStream_t stream, bool force_split) {
  using Cfg = Gemm2Cfg<BN, TMA_OUT>;
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(gemm2_bf16_kernel<BN, TMA_OUT, EPI>), Cfg::SMEM_BYTES));
  const CUtensorMap* ta = get_tmap_2d_bf16(p->a, uint64_t(p->conv_taps == 9 ? p->conv_cin : (p->a2 ? p->K - p->k2 : p->K)), uint64_t(p->M) * (p->conv_taps == 9 && p->conv_stride == 2 ? 4 : 1), uint64_t(p->lda) * 2, BK, BM);
  if (!ta) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tb = get_tmap_2d_bf16(p->w, uint64_t(p->K), uint64_t(p->N), uint64_t(p->ldw) * 2, BK, BN / 2);
  if (!tb) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tc = ta;   // unused unless TMA_OUT
  if (TMA_OUT) {
    tc = get_tmap_2d_bf16(p->out_bf16, uint64_t(p->N), uint64_t(p->M), uint64_t(p->ld_bf16) * 2, 64, BM);
    if (!tc) return HOIGEN_ERR_CUDA;
  }
  const int num_tiles = ((p->M + 2 * BM - 1) / (2 * BM)) * ((p->N + BN - 1) / BN);
  const int num_k = (p->K + BK - 1) / BK;
  const int max_clusters = num_sms() / 2;
  GemmArgs args = to_args(p);
  // Tiles that do not fill the last round: split the last (1 + remainder) rounds' tiles across all pairs by k-blocks
  static const bool no_streamk = getenv("HOIGEN_GEMM_NO_STREAMK") != nullptr;
  const int rem = num_tiles % max_clusters;
  int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  if (!no_streamk && rem != 0) {
    const int sk_tiles = num_tiles < max_clusters ? num_tiles : rem + max_clusters;
    // Measured (B200): the split's fixed cost (partial dump + fix-up, ~8 us) only pays when a round is long; the
    // K = 768 shapes are epilogue-bound and lose.  HOIGEN_GEMM_STREAMK_MIN_K overrides the threshold (k-blocks).
    static const int min_k = getenv("HOIGEN_GEMM_STREAMK_MIN_K") ? atoi(getenv("HOIGEN_GEMM_STREAMK_MIN_K")) : 32;
    if ((num_k >= min_k || force_split) && (long long)sk_tiles * num_k / max_clusters >= 6) {
      StreamKWorkspace ws = get_streamk_workspace(stream, size_t(max_clusters) * 2 * 256 * BM * sizeof(float), max_clusters * 2);
      if (!ws.slots) return HOIGEN_ERR_CUDA;
      args.sk_ws = ws.slots; args.sk_flags = ws.flags;
      args.sk_tiles = sk_tiles;
      args.sk_snap = num_k >= 12 ? 2 : (num_k >= 6 ? 1 : 0);
      clusters = max_clusters;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM2_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KernelScope ks(gemm_tag(p->N, p->K, 2), stream, 2.0 * p->M * p->N * p->K,
                 2.0 * (double(p->M) * p->K + double(p->N) * p->K) + double(p->M) * p->N * ((p->out_f32 ? 4 : 0) + (p->out_bf16 ? 2 : 0) + (p->residual ? 4 : 0)));
  const CUtensorMap* tr = tc;   // identity tiles (RES_TMA recipes) or the second A source
  if (p->a2) {
    tr = get_tmap_2d_bf16(p->a2, uint64_t(p->k2), uint64_t(p->M), uint64_t(p->lda2) * 2, BK, BM);
    if (!tr) return HOIGEN_ERR_CUDA;
  } else if (TMA_OUT && EPI >= 0 && ((EPI >> 7) & 1) != 0) {
    tr = get_tmap_2d_bf16(p->res_bf16, uint64_t(p->N), uint64_t(p->M), uint64_t(p->ld_resb) * 2, 64, BM);
    if (!tr) return HOIGEN_ERR_CUDA;
  }
  HOIGEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm2_bf16_kernel<BN, TMA_OUT, EPI>, *ta, *tb, *tc, *tr, args));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

template <int BN>
static int launch_gemm2(const hoigen_gemm_params* p, cudaStream_t stream, bool force_split) {
  // bf16-only outputs with TMA-compatible strides go through the smem-staged TMA-store epilogue
  static const bool no_tma_out = getenv("HOIGEN_GEMM_NO_TMA_STORE") != nullptr;
  const bool tma_out = !no_tma_out && p->out_bf16 && !p->out_f32 && !p->residual && (p->ld_bf16 % 8) == 0 && (p->N % 8) == 0;
  if (!tma_out) return launch_gemm2_impl<BN, false, -1>(p, stream, force_split);

  // compile-time epilogue recipes of the encoder's bf16-output GEMMs; anything else uses the runtime-flag epilogue
  const bool b = p->bias != nullptr, c = p->colscale != nullptr;
  if (p->halo_w > 0 || p->res_bf16) {   // convolution epilogues of the ResNet-50 branch (bias folded from the BatchNorm)
    constexpr int HALO = 1 << 8, RESB = 1 << 7, BIAS = 1 << 5;
    const bool plain = b && !p->colscale && !p->ln_stats && p->halo_w > 0;
    if (plain && !p->res_bf16 && p->act == HOIGEN_ACT_RELU) return launch_gemm2_impl<BN, true, HALO | BIAS | HOIGEN_ACT_RELU>(p, stream, force_split);
    if (plain && !p->res_bf16 && p->act == HOIGEN_ACT_NONE) return launch_gemm2_impl<BN, true, HALO | BIAS>(p, stream, force_split);
    if (plain && p->res_bf16 && p->act == HOIGEN_ACT_RELU && (p->ld_resb % 8) == 0 && (reinterpret_cast<uintptr_t>(p->res_bf16) & 15) == 0) return launch_gemm2_impl<BN, true, HALO | RESB | BIAS | HOIGEN_ACT_RELU>(p, stream, force_split);
    return launch_gemm2_impl<BN, true, -1>(p, stream, force_split);
  }
  if (p->ln_stats) {   // LayerNorm-folded QKV / c_fc
    if (b && p->act == HOIGEN_ACT_NONE) return launch_gemm2_impl<BN, true, (1 << 6) | (1 << 5)>(p, stream, force_split);
    if (b && p->act == HOIGEN_ACT_QUICKGELU) return launch_gemm2_impl<BN, true, (1 << 6) | (1 << 5) | HOIGEN_ACT_QUICKGELU>(p, stream, force_split);
    return launch_gemm2_impl<BN, true, -1>(p, stream, force_split);
  }
  if (b && !c && p->act == HOIGEN_ACT_NONE) return launch_gemm2_impl<BN, true, (1 << 5)>(p, stream, force_split);                          // QKV, out-proj
  if (b && !c && p->act == HOIGEN_ACT_QUICKGELU) return launch_gemm2_impl<BN, true, (1 << 5) | HOIGEN_ACT_QUICKGELU>(p, stream, force_split);  // c_fc
  if (b && c && p->act == HOIGEN_ACT_NONE) return launch_gemm2_impl<BN, true, (1 << 5) | (1 << 4)>(p, stream, force_split);                 // adapter up-proj
  if (!b && !c && p->act == HOIGEN_ACT_NONE) return launch_gemm2_impl<BN, true, 0>(p, stream, force_split);                                 // cache affinity
  return launch_gemm2_impl<BN, true, -1>(p, stream, force_split);
}

// Pick (cta pair?, BN): fewest (possibly fractional, see stream-K) scheduling rounds x relative tile time. The one-CTA kernel is L2-operand-bound
// (~87 FLOP per L2 byte at BN = 256) and stores through per-thread writes: measured 550-680 TFLOP/s against 900-1060 for the
// pair kernel on the K = 768 shapes at M = 25 216, so its tiles are charged 1.6x (1.35x let a one-round quantisation
// advantage pick it for BASELINE config 4).
static int auto_split_k(int num_tiles, int num_k);
static void choose_config(int M, int N, int K, int* pair, int* bn) {
  const int sms = num_sms();
  const int nk = (K + BK - 1) / BK;
  if (nk >= 32 && N <= 128 && ((M + BM - 1) / BM) * ((N + 63) / 64) <= sms / 2) {
    // long K, skinny N, few tiles (measured: M = 64 x N = 117 x K = 4096 runs 25 -> 19 us with the split; with 120 tiles
    // the unsplit BN = 64 schedule below stays the best): compare bytes per CTA chain,
    // rounds x k-blocks per unit x (A 16 KiB + W tile), with the k-split the launcher will pick
    double best = 1e30;
    const int cand[2] = {64, 128};
    for (int i = 0; i < 2; ++i) {
      const int b = cand[i];
      if (b > 64 && N <= b / 2) continue;
      const int tiles = ((M + BM - 1) / BM) * ((N + b - 1) / b);
      const int S = auto_split_k(tiles, nk);
      const int rounds = (tiles * S + sms - 1) / sms;
      const double cost = double(rounds) * (double(nk) / S) * (16384.0 + b * 128.0) + (S > 1 ? 40000.0 : 0.0);
      if (cost < best) { best = cost; *pair = 0; *bn = b; }
    }
    return;
  }
  double best = 1e30;
  *pair = 0; *bn = 256;
  const int c1[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int b = c1[i];
    if (b > 64 && N <= b / 2) continue;
    const int tiles = ((M + BM - 1) / BM) * ((N + b - 1) / b);
    const int rounds = (tiles + sms - 1) / sms;
    const double cost = rounds * (double(b) * 1.6 + 24.0);
    if (cost < best) { best = cost; *pair = 0; *bn = b; }
  }
  if (M > BM) {
    const int c2[3] = {256, 192, 128};
    for (int i = 0; i < 3; ++i) {
      const int b = c2[i];
      if (N <= b / 2) continue;
      const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + b - 1) / b);
      const int pairs = sms / 2, num_k = (K + BK - 1) / BK;
      // the stream-K split (launch_gemm2_impl) balances to a k-block: fractional rounds plus the fix-up's share
      const int sk_tiles = tiles < pairs ? tiles : tiles % pairs + pairs;
      const bool split = num_k >= 32 && tiles % pairs != 0 && (long long)sk_tiles * num_k / pairs >= 6;
      const double rounds = split ? double(tiles) / pairs + 0.4 : double((tiles + pairs - 1) / pairs);
      const double cost = rounds * (double(b) + 96.0);
      if (cost < best) { best = cost; *pair = 1; *bn = b; }
    }
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_gemm_bf16(const hoigen_gemm_params* p, hoigen_stream_t stream) {
  using namespace hoigen;
  int rc = validate(p);
  if (rc != HOIGEN_OK) return rc;
  // block_n: 0 = choose; 64/128/256 = one-CTA kernel; 2128/2192/2256 = CTA-pair kernel (256 x {128,192,256} tiles)
  int bn = p->block_n, pair = 0;
  const bool force_split = bn >= 10000;   // testing: +10000 forces the stream-K split whatever K is
  if (force_split) bn -= 10000;
  if (bn == 0) {
    choose_config(p->M, p->N, p->K, &pair, &bn);
    if (p->a2 && !pair) { pair = 1; bn = p->N > 128 ? 256 : 128; }   // the second A source exists in the CTA-pair kernel only
  } else if (bn > 2000) { pair = 1; bn -= 2000; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (pair) {
    switch (bn) {
      case 256: return launch_gemm2<256>(p, s, force_split);
      case 192: return launch_gemm2<192>(p, s, force_split);
      case 128: return launch_gemm2<128>(p, s, force_split);
      default: break;
    }
  } else {
    switch (bn) {
      case 256: return launch_gemm<256>(p, s);
      case 128: return launch_gemm<128>(p, s);
      case 64: return launch_gemm<64>(p, s);
      default: break;
    }
  }
  set_error("gemm: block_n must be 0, 64/128/256 (one CTA) or 2128/2192/2256 (CTA pair); got %d", p->block_n);
  return HOIGEN_ERR_INVALID;
}

int hoigen_debug_gemm_simt(const hoigen_gemm_params* p, hoigen_stream_t stream) {
  using namespace hoigen;
  int rc = validate(p);
  if (rc != HOIGEN_OK) return rc;
  dim3 grid((p->N + 127) / 128, p->M);
  KernelScope ks("gemm_simt_debug", reinterpret_cast<cudaStream_t>(stream), 2.0 * p->M * p->N * p->K, 0);
  gemm_simt_kernel<<<grid, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(p->a), reinterpret_cast<const __nv_bfloat16*>(p->w), p->lda, p->ldw,
      to_args(p));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
