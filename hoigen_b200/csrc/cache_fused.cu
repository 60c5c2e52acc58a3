// Tip-Adapter cache-model scoring as ONE fused GEMM - f - GEMM kernel on tcgen05 (north_star item 3):
//
//     L_X[i, c] = sum_n  aff( f_X[i, :] . W_X[n, :] )  *  Y_X[n, c]          X in {H, O, U}
//
// with aff(s) = s (the reference's LINEAR affinity: phi = f W^T + b, U:1156-1158; the bias is carried exactly in fp32 as
// (b Y) by the combine pass) or aff(s) = exp(beta (s + b_n)) (the textbook Tip-Adapter exp(-beta (1 - q K^T)) for b = -1;
// compile-time switch).  The (Ktot x N) affinity matrix `phi` NEVER exists in memory: per 64 cache rows it lives as an
// fp32 tile in TMEM, is converted in registers to bf16 and written back over the same TMEM columns as the A operand of the
// second MMA (tcgen05.mma, A from TMEM) against the label tile.
//
// Work unit = (branch X, cache split sp, 128-pair row tile); persistent CTAs, one per SM, units dealt round-robin in
// (X, sp)-major order so that concurrently running CTAs stream the same cache slice through L2.
//   warp 0      TMA producer : W k-blocks [64 cache rows x 64 k] (8 KiB, ring of 16) and label chunks [C_pad x 64] (ring of 3)
//   warp 1      MMA issuer   : S_j = F W_j^T  (32 x UMMA 128x64x16, A = F from TMEM) one chunk ahead of
//                              L  += P_{j-1} Y_{j-1}  (4 x UMMA 128xC_padx16, A = P from TMEM)
//   warps 2-9   two threads per pair row: F tile global -> TMEM (once per unit), S -> aff -> bf16 P -> TMEM per chunk,
//               and the unit's epilogue (raw fp32 partial sums -> parts[X][sp])
// TMEM (512 columns): F (128 rows x 512 k, bf16 pairs) [0,256) | S/P buffers [256,320) [320,384) | L [384,512).
// A second pass (cache_combine_kernel) adds, in a FIXED order (bit-reproducible, no atomics), the per-image global / DINO
// terms, the exact bias carriers and the parts of every branch and split into the logits.
//
// Replaces U:1156-1163 (gen_feat cache branches) for C <= 128; wider classifiers (600 HOI triplets) keep the two-GEMM form.
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int CF_THREADS = 320;
constexpr int CF_ROWS = 128;        // pairs per unit
constexpr int CF_CHUNK = 64;        // cache rows per chunk
constexpr int CF_K = 512;           // feature width
constexpr int CF_KB = CF_K / 64;    // 8 k-blocks per chunk
constexpr int CF_KB_BYTES = CF_CHUNK * 128;         // 8 KiB: one [64 cache rows x 64 k] block
constexpr int CF_W_BYTES = CF_KB * CF_KB_BYTES;     // 64 KiB: one chunk of keys
constexpr int CF_Y_BYTES = 128 * 128;               // 16 KiB (C_pad <= 128 rows of 128 B)
constexpr int CF_SMEM_Y = 2 * CF_W_BYTES;
constexpr int CF_SMEM_BAR = CF_SMEM_Y + 2 * CF_Y_BYTES;
constexpr int CF_SMEM_BYTES = CF_SMEM_BAR + 256 + 1024;
constexpr uint32_t CF_TM_F = 0, CF_TM_S = 256, CF_TM_L = 384;

struct CacheFusedArgs {
  const __nv_bfloat16* feat;      // [3][ktot][512]
  const float* cache_bias[3];     // (N) per branch — exp affinity only
  float* parts;                   // [3 * nsplit][ktot_pad][c_pad]
  int ktot, ktot_pad, n_rows, c_pad, nsplit, chunks_per_split, num_tiles;
  float beta_log2e;
  int debug;   // diagnostics (HOIGEN_CF_DEBUG): 1 = no S MMAs, 2 = no L MMAs, 3 = no F-tile load, 4 = no S -> P conversion, 5 = no TMA loads
};

// 4 x 32 bytes of one pair row in ONE statement, so that the eight loads are in flight together (separate volatile asm
// statements would be kept in program order with the TMEM stores between them: one L2 round trip each)
__device__ __forceinline__ void ld_global_128B(const void* p, uint32_t (&r)[32]) {
  asm volatile(
      "ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%32];\n\t"
      "ld.global.nc.v4.u32 {%4, %5, %6, %7}, [%32 + 16];\n\t"
      "ld.global.nc.v4.u32 {%8, %9, %10, %11}, [%32 + 32];\n\t"
      "ld.global.nc.v4.u32 {%12, %13, %14, %15}, [%32 + 48];\n\t"
      "ld.global.nc.v4.u32 {%16, %17, %18, %19}, [%32 + 64];\n\t"
      "ld.global.nc.v4.u32 {%20, %21, %22, %23}, [%32 + 80];\n\t"
      "ld.global.nc.v4.u32 {%24, %25, %26, %27}, [%32 + 96];\n\t"
      "ld.global.nc.v4.u32 {%28, %29, %30, %31}, [%32 + 112];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "l"(p));
}

// Barriers are per CHUNK (64 cache rows), not per k-block: a tcgen05.commit costs a few hundred cycles of the single MMA
// thread's time, and with 128x64x16 MMAs (32-64 cycles each) one commit per four MMAs was measured to dominate the
// kernel (185 us of 266 with the MMAs themselves disabled).  Chunk c uses key stage / label stage / S-P buffer c & 1:
//   wfull[b], yfull[b]  TMA landed                      sdone[b]  S_c complete: converters read it, producer refills keys
//   pready[b]           P_c written (8 warps)           ldone[b]  L += P_c Y_c complete: S/P buffer + label stage free
template <bool EXP>
__global__ void __launch_bounds__(CF_THREADS, 1)
cache_fused_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY0,
                   const __grid_constant__ CUtensorMap tmY1, const __grid_constant__ CUtensorMap tmY2, CacheFusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF_SMEM_BAR);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_wfull = bar0;             // [2]
  const uint32_t bar_yfull = bar0 + 16;        // [2]
  const uint32_t bar_sdone = bar0 + 32;        // [2]
  const uint32_t bar_pready = bar0 + 48;       // [2]
  const uint32_t bar_ldone = bar0 + 64;        // [2]
  const uint32_t bar_fready = bar0 + 80;       // F tile of the unit in TMEM (8 warps)
  const uint32_t bar_lfree = bar0 + 88;        // L drained by the unit's epilogue (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + CF_SMEM_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmY0); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_wfull + 8u * s, 1); mbar_init(bar_yfull + 8u * s, 1); mbar_init(bar_sdone + 8u * s, 1);
      mbar_init(bar_pready + 8u * s, 8); mbar_init(bar_ldone + 8u * s, 1);
    }
    mbar_init(bar_fready, 8);
    mbar_init(bar_lfree, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int total_units = 3 * g.nsplit * g.num_tiles;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < total_units ? (total_units - first + stride - 1) / stride : 0;
  const int C = g.chunks_per_split;
  // unit u -> (part p = X * nsplit + sp, tile)
  auto decode = [&](int u, int& x, int& sp, int& tile) {
    const int p = u / g.num_tiles;
    tile = u - p * g.num_tiles;
    x = p / g.nsplit;
    sp = p - x * g.nsplit;
  };
  // chunk c is the (c >> 1)-th user of buffer c & 1: its barriers complete with parity (c >> 1) & 1
  auto par = [](long c) { return uint32_t(c >> 1) & 1u; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      long cj = 0;
      for (int i = 0; i < n_local; ++i) {
        int x, sp, tile;
        decode(first + i * stride, x, sp, tile);
        const CUtensorMap* tw = x == 0 ? &tmW0 : (x == 1 ? &tmW1 : &tmW2);
        const CUtensorMap* ty = x == 0 ? &tmY0 : (x == 1 ? &tmY1 : &tmY2);
        for (int j = 0; j < C; ++j, ++cj) {
          const int b = int(cj & 1);
          const int n0 = (sp * C + j) * CF_CHUNK;
          if (cj >= 2) mbar_wait(bar_sdone + 8u * b, par(cj - 2));      // S_{c-2} has read key stage b
          if (g.debug == 5) mbar_arrive(bar_wfull + 8u * b);
          else {
            mbar_arrive_expect_tx(bar_wfull + 8u * b, CF_W_BYTES);
#pragma unroll
            for (int kb = 0; kb < CF_KB; ++kb)
              tma_load_2d(base + b * CF_W_BYTES + kb * CF_KB_BYTES, tw, bar_wfull + 8u * b, kb * 64, n0);
          }
          if (cj >= 2) mbar_wait(bar_ldone + 8u * b, par(cj - 2));      // L of chunk c-2 has read label stage b
          if (g.debug == 5) mbar_arrive(bar_yfull + 8u * b);
          else {
            mbar_arrive_expect_tx(bar_yfull + 8u * b, uint32_t(g.c_pad) * 128u);
            tma_load_2d(base + CF_SMEM_Y + b * CF_Y_BYTES, ty, bar_yfull + 8u * b, n0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(CF_ROWS, CF_CHUNK);
      const uint32_t idesc_l = make_idesc_bf16(CF_ROWS, g.c_pad);
      long cj = 0;
      for (int i = 0; i < n_local; ++i) {
        mbar_wait(bar_fready, i & 1u);
        tc_fence_after();
        auto issue_l = [&](int j, long c) {
          const int b = int(c & 1);
          mbar_wait(bar_pready + 8u * b, par(c));
          mbar_wait(bar_yfull + 8u * b, par(c));
          if (j == 0 && i > 0) mbar_wait(bar_lfree, (i - 1) & 1u);      // previous unit's L drained
          tc_fence_after();
          const uint32_t yaddr = base + CF_SMEM_Y + b * CF_Y_BYTES;
          const uint32_t pbase = tmem + CF_TM_S + uint32_t(b * 64);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (g.debug == 2) break;
            // P k-step t: k-values 16t..16t+15 = packed columns {0, 8, 32, 40}[t] of the buffer (each converter thread
            // overlays ITS OWN 32 S columns with its 16 packed P columns)
            const uint32_t pcol = uint32_t((k >> 1) * 32 + (k & 1) * 8);
            umma_bf16_ts(tmem + CF_TM_L, pbase + pcol, make_sdesc_sw128(yaddr + k * 32), idesc_l, (j > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(bar_ldone + 8u * b);
        };
        for (int j = 0; j < C; ++j, ++cj) {
          const int b = int(cj & 1);
          if (cj >= 2) mbar_wait(bar_ldone + 8u * b, par(cj - 2));     // P_{c-2} consumed: buffer b may take S_c
          mbar_wait(bar_wfull + 8u * b, par(cj));
          tc_fence_after();
          const uint32_t waddr = base + b * CF_W_BYTES;
#pragma unroll
          for (int kk = 0; kk < CF_KB * 4; ++kk) {
            if (g.debug == 1) break;
            umma_bf16_ts(tmem + CF_TM_S + uint32_t(b * 64), tmem + CF_TM_F + uint32_t(kk * 8),
                         make_sdesc_sw128(waddr + (kk >> 2) * CF_KB_BYTES + (kk & 3) * 32), idesc_s, kk > 0 ? 1u : 0u);
          }
          tc_commit(bar_sdone + 8u * b);
          if (j > 0) issue_l(j - 1, cj - 1);       // lags one chunk: the conversion of S_{j-1} ran under the MMAs of S_j
        }
        issue_l(C - 1, cj - 1);
      }
    }
  } else {
    // ===================== converter / epilogue warps: two threads per pair row =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rrow = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    long cj = 0;
    auto load_f = [&](int i) {
      int x, sp, tile;
      decode(first + i * stride, x, sp, tile);
      const int row = tile * CF_ROWS + rrow;
      const bool ok = row < g.ktot;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(g.feat + (size_t(x) * g.ktot + (ok ? row : 0)) * CF_K) + half * 512;
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        if (g.debug == 3) break;
        uint32_t v[32];
        if (ok) ld_global_128B(src + it * 128, v);
        else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w8[e] = v[q * 8 + e];
          tmem_st_32x32b_x8(tmem + lane_addr + CF_TM_F + uint32_t(half * 128 + it * 32 + q * 8), w8);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fready);
    };
    auto epilogue = [&](int i, long c_last) {
      int x, sp, tile;
      decode(first + i * stride, x, sp, tile);
      mbar_wait(bar_ldone + 8u * uint32_t(c_last & 1), par(c_last));     // L of the unit's last chunk complete
      tc_fence_after();
      const int row = tile * CF_ROWS + rrow;
      const int cols = g.c_pad / 2;                        // columns of this thread (multiple of 8)
      float* dst = g.parts + (size_t(x * g.nsplit + sp) * g.ktot_pad + row) * g.c_pad + half * cols;
      for (int c0 = 0; c0 < cols; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + lane_addr + CF_TM_L + uint32_t(half * cols + c0))
                     : "memory");
        tmem_wait_ld();
        if (row < g.ktot_pad) {
          reinterpret_cast<float4*>(dst + c0)[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
          reinterpret_cast<float4*>(dst + c0)[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_lfree);
    };
    if (n_local > 0) load_f(0);
    for (int i = 0; i < n_local; ++i) {
      int x, sp, tile;
      decode(first + i * stride, x, sp, tile);
      const float* bias = EXP ? g.cache_bias[x] : nullptr;
      for (int j = 0; j < C; ++j, ++cj) {
        const int b = int(cj & 1);
        const uint32_t sbuf = tmem + lane_addr + CF_TM_S + uint32_t(b * 64 + half * 32);
        mbar_wait(bar_sdone + 8u * b, par(cj));
        tc_fence_after();
        if (g.debug == 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_pready + 8u * b);
          continue;
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(sbuf, r);
        tmem_wait_ld();
        uint32_t pk[16];
        if (EXP) {
          const int n0 = (sp * C + j) * CF_CHUNK + half * 32;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int n = n0 + 2 * q;
            const float b0 = n < g.n_rows ? __ldg(bias + n) : 0.f, b1 = n + 1 < g.n_rows ? __ldg(bias + n + 1) : 0.f;
            float e0, e1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"((__uint_as_float(r[2 * q]) + b0) * g.beta_log2e));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"((__uint_as_float(r[2 * q + 1]) + b1) * g.beta_log2e));
            pk[q] = pack_bf16x2(e0, e1);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) pk[q] = pack_bf16x2(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
        }
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { lo[q] = pk[q]; hi[q] = pk[8 + q]; }
        tmem_st_32x32b_x8(sbuf, lo);               // overlays this thread's own (already loaded) S columns
        tmem_st_32x32b_x8(sbuf + 8, hi);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pready + 8u * b);
      }
      // unit boundary: every S MMA of this unit is complete (sdone of its last chunk was observed above), so the F columns
      // may take the next unit's tile now — its S MMAs then start while this unit's last L MMA and epilogue finish
      if (i + 1 < n_local) load_f(i + 1);
      epilogue(i, cj - 1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// logits[i][c] = img_logits[image(i)][c] + sum_X colscale_X[c] * (bias_term_X[c] + sum_sp parts[X][sp][i][c])   (fixed order)
__global__ void __launch_bounds__(256)
cache_combine_kernel(const float* __restrict__ parts, int nsplit, int ktot, int ktot_pad, int c_pad, int num_classes,
                     const float* __restrict__ img_logits, const int* __restrict__ pair_off, int nimg,
                     const float* __restrict__ bt0, const float* __restrict__ bt1, const float* __restrict__ bt2,
                     const float* __restrict__ cs0, const float* __restrict__ cs1, const float* __restrict__ cs2,
                     int use_bias, int ld, float* __restrict__ logits) {
  const long total = long(ktot) * num_classes;
  for (long e = blockIdx.x * long(blockDim.x) + threadIdx.x; e < total; e += long(gridDim.x) * blockDim.x) {
    const int i = int(e / num_classes), c = int(e % num_classes);
    int lo = 0, hi = nimg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (pair_off[mid] <= i) lo = mid; else hi = mid;
    }
    float acc = img_logits ? img_logits[size_t(lo) * num_classes + c] : 0.f;
    const float* bts[3] = {bt0, bt1, bt2};
    const float* css[3] = {cs0, cs1, cs2};
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      float s = use_bias ? bts[x][c] : 0.f;
      for (int sp = 0; sp < nsplit; ++sp) s += parts[(size_t(x * nsplit + sp) * ktot_pad + i) * c_pad + c];
      acc += css[x][c] * s;
    }
    logits[size_t(i) * ld + c] = acc;
  }
}

}  // namespace hoigen

extern "C" {

int64_t hoigen_cache_fused_workspace_bytes(int32_t ktot, int32_t num_classes) {
  if (ktot < 0 || num_classes <= 0 || num_classes > 128) return -1;
  const int64_t c_pad = (num_classes + 15) / 16 * 16;
  const int64_t ktot_pad = (int64_t(ktot) + 255) / 256 * 256;
  return 3 * 4 /*max nsplit*/ * ktot_pad * c_pad * 4;
}

int hoigen_score_cache_fused(const hoigen_score_weights* w, const void* pair_feat_bf16, const float* const* cache_bias,
                             const float* img_logits, const int32_t* pair_off, int32_t batch, int32_t ktot, int32_t affinity,
                             float beta, float* parts, float* logits, int32_t ld_logits, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(w && pair_feat_bf16 && pair_off && parts && logits, "score_cache_fused: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ktot > 0, "score_cache_fused: bad sizes");
  const int C = w->num_classes, N = w->cache_rows;
  HOIGEN_CHECK_ARG(C > 0 && C <= 128, "score_cache_fused: num_classes must be in [1,128] (got %d); use hoigen_score_pairs", C);
  HOIGEN_CHECK_ARG(N > 0 && (N % 8) == 0, "score_cache_fused: cache_rows must be a positive multiple of 8 (got %d)", N);
  HOIGEN_CHECK_ARG(ld_logits >= C, "score_cache_fused: ld_logits < num_classes");
  HOIGEN_CHECK_ARG(affinity == 0 || affinity == 1, "score_cache_fused: affinity must be 0 (linear) or 1 (exp)");
  HOIGEN_CHECK_ARG(affinity == 0 || (cache_bias && cache_bias[0] && cache_bias[1] && cache_bias[2]),
                   "score_cache_fused: the exp affinity needs the per-row cache biases");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static const bool one_cta = getenv("HOIGEN_CF_ONE_CTA") != nullptr;     // A/B switch: the one-CTA form below
  if (!one_cta) {
    int nsplit = 1, ktot_pad = 0, c_pad = 0;
    HOIGEN_TRY_RC(launch_cache_fused_pair(w, pair_feat_bf16, cache_bias, ktot, affinity, beta, parts, &nsplit, &ktot_pad, &c_pad, s));
    const long total = long(ktot) * C;
    const int blocks = int(std::min<long>(long(num_sms()) * 8, (total + 255) / 256));
    KernelScope ks("cache_combine", s, 0, double(3 * nsplit) * ktot * C * 4 + double(ktot) * C * 4);
    cache_combine_kernel<<<blocks, 256, 0, s>>>(parts, nsplit, ktot, ktot_pad, c_pad, C, img_logits, pair_off, batch,
                                                w->bias_term[0], w->bias_term[1], w->bias_term[2], w->colscale[0],
                                                w->colscale[1], w->colscale[2], affinity == 0 ? 1 : 0, ld_logits, logits);
    HOIGEN_CHECK_LAUNCH();
    return HOIGEN_OK;
  }
  CacheFusedArgs g;
  g.feat = reinterpret_cast<const __nv_bfloat16*>(pair_feat_bf16);
  for (int x = 0; x < 3; ++x) g.cache_bias[x] = cache_bias ? cache_bias[x] : nullptr;
  g.parts = parts;
  g.ktot = ktot;
  g.ktot_pad = (ktot + 127) / 128 * 128;
  g.n_rows = N;
  g.c_pad = (C + 15) / 16 * 16;
  g.num_tiles = g.ktot_pad / 128;
  g.beta_log2e = beta * 1.4426950408889634f;
  g.debug = getenv("HOIGEN_CF_DEBUG") ? atoi(getenv("HOIGEN_CF_DEBUG")) : 0;
  const int force_split = getenv("HOIGEN_CF_NSPLIT") ? atoi(getenv("HOIGEN_CF_NSPLIT")) : 0;
  // cache split: fewest rounds x (unit cost + per-unit F-tile load), units = 3 * nsplit * tiles over one CTA per SM
  const int chunks = (N + CF_CHUNK - 1) / CF_CHUNK;
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 4; ns *= 2) {
    if (chunks / ns < 4 && ns > 1) break;
    const int cps = (chunks + ns - 1) / ns;
    const int units = 3 * ns * g.num_tiles;
    const int rounds = (units + num_sms() - 1) / num_sms();
    const double cost = double(rounds) * (double(cps) * (1024.0 + 2.0 * g.c_pad) + 3500.0);
    if (cost < best_cost) { best_cost = cost; best = ns; }
  }
  if (force_split == 1 || force_split == 2 || force_split == 4) best = force_split;
  g.nsplit = best;
  g.chunks_per_split = (chunks + best - 1) / best;
  const CUtensorMap* tw[3];
  const CUtensorMap* ty[3];
  for (int x = 0; x < 3; ++x) {
    tw[x] = get_tmap_2d_bf16(w->cache_keys[x], CF_K, uint64_t(N), uint64_t(CF_K) * 2, 64, CF_CHUNK);
    if (!tw[x]) return HOIGEN_ERR_CUDA;
    ty[x] = get_tmap_2d_bf16(w->label_t[x], uint64_t(N), uint64_t(C), uint64_t(N) * 2, 64, uint32_t(g.c_pad));
    if (!ty[x]) return HOIGEN_ERR_CUDA;
  }
  const int units = 3 * g.nsplit * g.num_tiles;
  const int grid = units < num_sms() ? units : num_sms();
  {
    KernelScope ks("cache_fused", s, 3.0 * 2.0 * ktot * double(N) * (CF_K + C),
                   3.0 * (double(ktot) * CF_K * 2 + double(N) * (CF_K + C) * 2 + double(ktot) * C * 4));
    if (affinity == 1) {
      HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(cache_fused_kernel<true>), CF_SMEM_BYTES));
      cache_fused_kernel<true><<<grid, CF_THREADS, CF_SMEM_BYTES, s>>>(*tw[0], *tw[1], *tw[2], *ty[0], *ty[1], *ty[2], g);
    } else {
      HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(cache_fused_kernel<false>), CF_SMEM_BYTES));
      cache_fused_kernel<false><<<grid, CF_THREADS, CF_SMEM_BYTES, s>>>(*tw[0], *tw[1], *tw[2], *ty[0], *ty[1], *ty[2], g);
    }
    HOIGEN_CHECK_LAUNCH();
  }
  {
    const long total = long(ktot) * C;
    const int blocks = int(std::min<long>(long(num_sms()) * 8, (total + 255) / 256));
    KernelScope ks("cache_combine", s, 0, double(3 * g.nsplit) * ktot * C * 4 + double(ktot) * C * 4);
    cache_combine_kernel<<<blocks, 256, 0, s>>>(parts, g.nsplit, ktot, g.ktot_pad, g.c_pad, C, img_logits, pair_off, batch,
                                                w->bias_term[0], w->bias_term[1], w->bias_term[2], w->colscale[0],
                                                w->colscale[1], w->colscale[2], affinity == 0 ? 1 : 0, ld_logits, logits);
    HOIGEN_CHECK_LAUNCH();
  }
  return HOIGEN_OK;
}

}  // extern "C"
