// CTA-PAIR form of the fused GEMM - f - GEMM cache kernel (cache_fused.cu has the algorithm and the one-CTA form):
// a cluster of two CTAs scores a 256-pair row tile with cta_group::2 MMAs.
//
// Why pairs, and why the first MMA takes BOTH operands from shared memory here (measured on the one-CTA form, B200):
//   * with F in TMEM the 128x64x16 MMA runs at ~62 clk instead of its 32-clk floor — fetching 4 KiB of A from TMEM per
//     instruction is the limit, whatever the tile's N — so S = F W^T alone cost 74 us of a 175 us kernel;
//   * per row tile the whole W_X + Y_X (5 MiB) streams through the SM: 943 MB of L2 -> SM traffic at 128-row tiles.
// A pair halves the key / label bytes per pair row (each CTA loads half of every W chunk and half of the label rows) and
// doubles the rows per MMA instruction; F stays resident in SHARED memory (128 rows x 512 k = 128 KiB per CTA) so
// the S MMAs are SS-form, and the TMEM that F occupied holds FOUR S / P buffers: the MMA thread never waits for a P -> L
// round trip before it may start the next S.
//
// Per CTA: smem = F 128 KiB | 2 x W half-chunk 32 KiB | 2 x Y half 8 KiB = 208 KiB; TMEM = 4 x 64 (S/P) + 128 (L).
// Barriers (same offsets in both CTAs; cluster protocol of gemm2_bf16_kernel):
//   ffull, wfull[2], yfull[2]   leader's, 2 arrivals + the TMA bytes of BOTH CTAs
//   sdone[4], ldone[4]          per CTA, tcgen05.commit multicast from the leader's MMA thread
//   pready[4], lfree            leader's, 16 arrivals (8 converter warps of each CTA)
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int CP_THREADS = 320;
constexpr int CP_CHUNK = 64;                       // cache rows per chunk (32 per CTA)
constexpr int CP_K = 512;
constexpr int CP_KB = CP_K / 64;                   // 8 k-blocks
constexpr int CP_NBUF = 4;                         // S / P buffers in TMEM
constexpr int CP_F_BYTES = CP_KB * 16384;          // 128 KiB: [8 k-blocks][128 rows x 64 k]
constexpr int CP_WKB_BYTES = 32 * 128;             // 4 KiB: one [32 cache rows x 64 k] block
constexpr int CP_W_BYTES = CP_KB * CP_WKB_BYTES;   // 32 KiB
constexpr int CP_Y_BYTES = 64 * 128;               // 8 KiB: [<= 64 class rows x 64 k]
constexpr int CP_SMEM_W = CP_F_BYTES;
constexpr int CP_SMEM_Y = CP_SMEM_W + 2 * CP_W_BYTES;
constexpr int CP_SMEM_BAR = CP_SMEM_Y + 2 * CP_Y_BYTES;
constexpr int CP_SMEM_BYTES = CP_SMEM_BAR + 256 + 1024;
constexpr uint32_t CP_TM_S = 0, CP_TM_L = 256;

struct CachePairArgs {
  const float* cache_bias[3];
  float* parts;                    // [3 * nsplit][ktot_pad][c_pad]
  int ktot, ktot_pad, n_rows, c_pad, nsplit, chunks_per_split, num_tiles;   // num_tiles: 256-row tiles
  float beta_log2e;
  int debug;   // diagnostics (HOIGEN_CF_DEBUG): 1 = no S MMAs, 2 = no L MMAs, 4 = no S -> P conversion, 5 = no key/label TMA loads, 6 = S MMAs with N = 16
};

template <bool EXP>
__global__ void __launch_bounds__(CP_THREADS, 1)
cache_fused_pair_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmW0,
                        const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                        const __grid_constant__ CUtensorMap tmY0, const __grid_constant__ CUtensorMap tmY1,
                        const __grid_constant__ CUtensorMap tmY2, CachePairArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  const uint32_t bar0 = smem_u32(sm + CP_SMEM_BAR);
  const uint32_t bar_ffull = bar0;
  const uint32_t bar_wfull = bar0 + 8;          // [2]
  const uint32_t bar_yfull = bar0 + 24;         // [2]
  const uint32_t bar_sdone = bar0 + 40;         // [4]
  const uint32_t bar_ldone = bar0 + 72;         // [4]
  const uint32_t bar_pready = bar0 + 104;       // [4]
  const uint32_t bar_lfree = bar0 + 136;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + CP_SMEM_BAR + 192);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmF);
    tma_prefetch_desc(&tmW0); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmY0); tma_prefetch_desc(&tmY1); tma_prefetch_desc(&tmY2);
    mbar_init(bar_ffull, 2);
    for (int s = 0; s < 2; ++s) { mbar_init(bar_wfull + 8u * s, 2); mbar_init(bar_yfull + 8u * s, 2); }
    for (int s = 0; s < CP_NBUF; ++s) {
      mbar_init(bar_sdone + 8u * s, 1); mbar_init(bar_ldone + 8u * s, 1); mbar_init(bar_pready + 8u * s, 16);
    }
    mbar_init(bar_lfree, 16);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int total_units = 3 * g.nsplit * g.num_tiles;
  const int n_local = cluster_id < total_units ? (total_units - cluster_id + num_clusters - 1) / num_clusters : 0;
  const int C = g.chunks_per_split;
  auto decode = [&](int i, int& x, int& sp, int& tile) {
    const int u = cluster_id + i * num_clusters;
    const int p = u / g.num_tiles;
    tile = u - p * g.num_tiles;
    x = p / g.nsplit;
    sp = p - x * g.nsplit;
  };
  // chunk c (global counter of this pair): key / label stage c & 1, S/P buffer c % 4
  auto par2 = [](long c) { return uint32_t(c >> 1) & 1u; };
  auto par4 = [](long c) { return uint32_t(c >> 2) & 1u; };
  const int half_c = g.c_pad / 2;              // label rows (classes) this CTA loads

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own F rows, own half of the keys / labels) =====================
    if (lane == 0) {
      long cj = 0;
      for (int i = 0; i < n_local; ++i) {
        int x, sp, tile;
        decode(i, x, sp, tile);
        const CUtensorMap* tw = x == 0 ? &tmW0 : (x == 1 ? &tmW1 : &tmW2);
        const CUtensorMap* ty = x == 0 ? &tmY0 : (x == 1 ? &tmY1 : &tmY2);
        // F tile: free once every S MMA of the previous unit has completed (= sdone of its last chunk)
        if (i > 0) mbar_wait(bar_sdone + 8u * uint32_t((cj - 1) & 3), par4(cj - 1));
        if (rank == 0) mbar_arrive_expect_tx(bar_ffull, 2 * CP_F_BYTES); else mbar_arrive_leader(bar_ffull);
#pragma unroll
        for (int kb = 0; kb < CP_KB; ++kb)
          tma_load_3d_2sm(base + kb * 16384, &tmF, bar_ffull, kb * 64, tile * 256 + int(rank) * 128, x);
        for (int j = 0; j < C; ++j, ++cj) {
          const int st = int(cj & 1);
          const int n0 = (sp * C + j) * CP_CHUNK;
          if (cj >= 2) mbar_wait(bar_sdone + 8u * uint32_t((cj - 2) & 3), par4(cj - 2));    // S_{c-2} has read key stage st
          if (g.debug == 5) { if (rank == 0) mbar_arrive(bar_wfull + 8u * st); else mbar_arrive_leader(bar_wfull + 8u * st); }
          else {
            if (rank == 0) mbar_arrive_expect_tx(bar_wfull + 8u * st, 2 * CP_W_BYTES); else mbar_arrive_leader(bar_wfull + 8u * st);
#pragma unroll
            for (int kb = 0; kb < CP_KB; ++kb)
              tma_load_2d_2sm(base + CP_SMEM_W + st * CP_W_BYTES + kb * CP_WKB_BYTES, tw, bar_wfull + 8u * st, kb * 64,
                              n0 + int(rank) * 32);
          }
          if (cj >= 2) mbar_wait(bar_ldone + 8u * uint32_t((cj - 2) & 3), par4(cj - 2));    // L of chunk c-2 has read label stage st
          if (g.debug == 5) { if (rank == 0) mbar_arrive(bar_yfull + 8u * st); else mbar_arrive_leader(bar_yfull + 8u * st); }
          else {
            if (rank == 0) mbar_arrive_expect_tx(bar_yfull + 8u * st, 2u * uint32_t(half_c) * 128u); else mbar_arrive_leader(bar_yfull + 8u * st);
            tma_load_2d_2sm(base + CP_SMEM_Y + st * CP_Y_BYTES, ty, bar_yfull + 8u * st, n0, int(rank) * half_c);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc_s = make_idesc_bf16(256, g.debug == 6 ? 16 : CP_CHUNK);
      const uint32_t idesc_l = make_idesc_bf16(256, g.c_pad);
      long cj = 0;
      for (int i = 0; i < n_local; ++i) {
        mbar_wait(bar_ffull, i & 1u);
        tc_fence_after();
        auto issue_l = [&](int j, long c) {
          const int b = int(c & 3), st = int(c & 1);
          mbar_wait(bar_pready + 8u * b, par4(c));
          mbar_wait(bar_yfull + 8u * st, par2(c));
          if (j == 0 && i > 0) mbar_wait(bar_lfree, (i - 1) & 1u);       // previous unit's L drained (both CTAs)
          tc_fence_after();
          const uint32_t yaddr = base + CP_SMEM_Y + st * CP_Y_BYTES;
          const uint32_t pbase = tmem + CP_TM_S + uint32_t(b * 64);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (g.debug == 2) break;
            const uint32_t pcol = uint32_t((k >> 1) * 32 + (k & 1) * 8);     // see cache_fused.cu: P overlays its own S columns
            umma_bf16_ts_2sm(tmem + CP_TM_L, pbase + pcol, make_sdesc_sw128(yaddr + k * 32), idesc_l, (j > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_2sm(bar_ldone + 8u * b);
        };
        for (int j = 0; j < C; ++j, ++cj) {
          const int b = int(cj & 3), st = int(cj & 1);
          if (cj >= CP_NBUF) mbar_wait(bar_ldone + 8u * b, par4(cj - CP_NBUF));    // P_{c-4} consumed: buffer b may take S_c
          mbar_wait(bar_wfull + 8u * st, par2(cj));
          tc_fence_after();
          const uint32_t waddr = base + CP_SMEM_W + st * CP_W_BYTES;
#pragma unroll
          for (int kk = 0; kk < CP_KB * 4; ++kk) {
            if (g.debug == 1) break;
            umma_bf16_ss_2sm(tmem + CP_TM_S + uint32_t(b * 64), make_sdesc_sw128(base + (kk >> 2) * 16384 + (kk & 3) * 32),
                             make_sdesc_sw128(waddr + (kk >> 2) * CP_WKB_BYTES + (kk & 3) * 32), idesc_s, kk > 0 ? 1u : 0u);
          }
          tc_commit_2sm(bar_sdone + 8u * b);
          if (j > 0) issue_l(j - 1, cj - 1);
        }
        issue_l(C - 1, cj - 1);
      }
    }
  } else {
    // ===================== converter / epilogue warps (both CTAs: own 128 rows) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int rrow = quad * 32 + lane;
    const uint32_t lane_addr = uint32_t(quad * 32) << 16;
    long cj = 0;
    for (int i = 0; i < n_local; ++i) {
      int x, sp, tile;
      decode(i, x, sp, tile);
      const float* bias = EXP ? g.cache_bias[x] : nullptr;
      for (int j = 0; j < C; ++j, ++cj) {
        const int b = int(cj & 3);
        const uint32_t sbuf = tmem + lane_addr + CP_TM_S + uint32_t(b * 64 + half * 32);
        mbar_wait(bar_sdone + 8u * b, par4(cj));
        tc_fence_after();
        if (g.debug == 4) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (rank == 0) mbar_arrive(bar_pready + 8u * b); else mbar_arrive_leader(bar_pready + 8u * b); }
          continue;
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(sbuf, r);
        tmem_wait_ld();
        uint32_t pk[16];
        if (EXP) {
          const int n0 = (sp * C + j) * CP_CHUNK + half * 32;
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
        tmem_st_32x32b_x8(sbuf, lo);
        tmem_st_32x32b_x8(sbuf + 8, hi);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (rank == 0) mbar_arrive(bar_pready + 8u * b); else mbar_arrive_leader(bar_pready + 8u * b); }
      }
      // ---- unit epilogue: raw fp32 partial sums of this CTA's 128 rows -> parts[X * nsplit + sp] ----
      const long c_last = cj - 1;
      mbar_wait(bar_ldone + 8u * uint32_t(c_last & 3), par4(c_last));
      tc_fence_after();
      const int row = tile * 256 + int(rank) * 128 + rrow;
      const int cols = g.c_pad / 2;
      float* dst = g.parts + (size_t(x * g.nsplit + sp) * g.ktot_pad + row) * g.c_pad + half * cols;
      for (int c0 = 0; c0 < cols; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + lane_addr + CP_TM_L + uint32_t(half * cols + c0))
                     : "memory");
        tmem_wait_ld();
        if (row < g.ktot_pad) {
          reinterpret_cast<float4*>(dst + c0)[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          reinterpret_cast<float4*>(dst + c0)[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (rank == 0) mbar_arrive(bar_lfree); else mbar_arrive_leader(bar_lfree); }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem, 512);
  }
}

int launch_cache_fused_pair(const hoigen_score_weights* w, const void* pair_feat_bf16, const float* const* cache_bias, int ktot,
                            int affinity, float beta, float* parts, int* nsplit_out, int* ktot_pad_out, int* c_pad_out,
                            cudaStream_t s) {
  const int C = w->num_classes, N = w->cache_rows;
  CachePairArgs g;
  for (int x = 0; x < 3; ++x) g.cache_bias[x] = cache_bias ? cache_bias[x] : nullptr;
  g.parts = parts;
  g.ktot = ktot;
  g.ktot_pad = (ktot + 255) / 256 * 256;
  g.n_rows = N;
  g.c_pad = (C + 15) / 16 * 16;
  g.num_tiles = g.ktot_pad / 256;
  g.beta_log2e = beta * 1.4426950408889634f;
  g.debug = getenv("HOIGEN_CF_DEBUG") ? atoi(getenv("HOIGEN_CF_DEBUG")) : 0;
  const int pairs = num_sms() / 2;
  const int chunks = (N + CP_CHUNK - 1) / CP_CHUNK;
  const int force_split = getenv("HOIGEN_CF_NSPLIT") ? atoi(getenv("HOIGEN_CF_NSPLIT")) : 0;
  int best = 1;
  double best_cost = 1e30;
  for (int ns = 1; ns <= 4; ns *= 2) {
    if (chunks / ns < 4 && ns > 1) break;
    const int cps = (chunks + ns - 1) / ns;
    const int units = 3 * ns * g.num_tiles;
    const int rounds = (units + pairs - 1) / pairs;
    const double cost = double(rounds) * (double(cps) * (1300.0 + 2.0 * g.c_pad) + 4500.0);   // + the unit's F-tile load
    if (cost < best_cost) { best_cost = cost; best = ns; }
  }
  if (force_split == 1 || force_split == 2 || force_split == 4) best = force_split;
  g.nsplit = best;
  g.chunks_per_split = (chunks + best - 1) / best;
  const CUtensorMap* tf = get_tmap_3d_bf16(pair_feat_bf16, CP_K, uint64_t(ktot), 3, uint64_t(CP_K) * 2, uint64_t(ktot) * CP_K * 2, 64, 128, 1);
  if (!tf) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tw[3];
  const CUtensorMap* ty[3];
  for (int x = 0; x < 3; ++x) {
    tw[x] = get_tmap_2d_bf16(w->cache_keys[x], CP_K, uint64_t(N), uint64_t(CP_K) * 2, 64, 32);
    if (!tw[x]) return HOIGEN_ERR_CUDA;
    ty[x] = get_tmap_2d_bf16(w->label_t[x], uint64_t(N), uint64_t(C), uint64_t(N) * 2, 64, uint32_t(g.c_pad / 2));
    if (!ty[x]) return HOIGEN_ERR_CUDA;
  }
  const int units = 3 * g.nsplit * g.num_tiles;
  const int clusters = std::min(units, pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(CP_THREADS);
  cfg.dynamicSmemBytes = CP_SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KernelScope ks("cache_fused", s, 3.0 * 2.0 * ktot * double(N) * (CP_K + C),
                 3.0 * (double(ktot) * CP_K * 2 + double(N) * (CP_K + C) * 2 + double(ktot) * C * 4));
  if (affinity == 1) {
    HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(cache_fused_pair_kernel<true>), CP_SMEM_BYTES));
    HOIGEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cache_fused_pair_kernel<true>, *tf, *tw[0], *tw[1], *tw[2], *ty[0], *ty[1], *ty[2], g));
  } else {
    HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(cache_fused_pair_kernel<false>), CP_SMEM_BYTES));
    HOIGEN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cache_fused_pair_kernel<false>, *tf, *tw[0], *tw[1], *tw[2], *ty[0], *ty[1], *ty[2], g));
  }
  HOIGEN_CHECK_LAUNCH();
  *nsplit_out = g.nsplit;
  *ktot_pad_out = g.ktot_pad;
  *c_pad_out = g.c_pad;
  return HOIGEN_OK;
}

}  // namespace hoigen
