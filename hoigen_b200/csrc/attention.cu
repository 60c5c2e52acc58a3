// 12-head self-attention over the 197-token sequence on tcgen05: PERSISTENT, warp-specialised CTAs (one per SM) that
// stream (image, head, 128-row tile) work items through a TMA -> UMMA -> softmax -> UMMA pipeline:
//
//   warp 0      TMA producer: Q tile (128x64), K (208x64), V (208x64) of item i+1 land while item i is computed
//   warp 1      MMA issuer:   S_i = Q K^T  (UMMA 128x208x16, 4 k-steps, fp32 in TMEM buffer i&1), issued one item
//                              ahead of the softmax;  O_i = P_i V (13 k-steps of UMMA 128x64x16 with the A operand P read
//                              straight from TMEM and V consumed MN-major straight from the QKV tile)
//   warps 2-9   softmax + epilogue, TWO threads per query row (key halves [0,112) and [112,208)):
//                              row max -> exp2 -> P (bf16 pairs) back into TMEM over the dead S columns -> ... ->
//                              O / rowsum -> bf16 -> swizzled smem tile -> ONE TMA store per item
// The whole key range (197 -> 208) fits one tile, so no online-softmax rescaling is needed.
//
// TMEM buffer b (256 columns): S in [0,208); once every thread holds its half row of S in registers the columns are
// dead, P (104 packed columns) overlays [0,104) and O accumulates into [128,192).  P never touches shared memory.
// Per item the softmax threads run  load S / max  ->  first exponentials  ->  drain O of the PREVIOUS item  ->  the
// remaining exponentials: the P V MMA of item i-1 completes under the first half, and draining O_{i-1} frees its TMEM
// buffer so S of item i+1 is computed under the second half (measured with hoigen_debug_attention_trace: the old
// order stalled ~700 clocks per item on P V and spent ~1700 on uncoalesced 16-byte row stores).
//
// Shared memory: 2 stages x [Q 16K | K 26K | V 26K] + O staging 16K = 152 KiB.  TMEM: 2 x 256 columns.
// Barriers: full/empty[s] (Q,K: released right after S), vfull/vempty[s] (V: released after PV), s_ready[b] (S in TMEM),
// p_ready (P in TMEM), o_ready (PV done), epi_done[b] (TMEM buffer b drained).
//
// Replaces F.multi_head_attention_forward -> SDPA at CLIP_models_adapter_prior2.py:443-445 (no mask,
// scale = 64^-0.5, dropout off).
#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int ATT_TOKENS = 197;
constexpr int ATT_KEYS = 208;      // 197 padded to a multiple of 16 (UMMA N / K granularity)
constexpr int ATT_DH = 64;
constexpr int ATT_HEADS = 12;
constexpr int ATT_WIDTH = 768;
constexpr int ATT_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 softmax/epilogue
constexpr int ATT_STAGE_Q = 0;
constexpr int ATT_STAGE_K = 16384;
constexpr int ATT_STAGE_V = 16384 + ATT_KEYS * 128;            // 43008
constexpr int ATT_STAGE_BYTES = 16384 + 2 * ATT_KEYS * 128;    // 69632
constexpr int ATT_SMEM_O = 2 * ATT_STAGE_BYTES;                // 139264 (1024-aligned): O tile staged for the TMA store
constexpr int ATT_SMEM_MISC = ATT_SMEM_O + 16384;              // barriers + row-stat exchange
constexpr int ATT_O_COL = 128;     // O accumulator columns [128,192) of a TMEM buffer; P overlays [0,104)
constexpr int ATT_SMEM_BYTES = ATT_SMEM_MISC + 128 + 4 * 256 * 4 + 1024;
constexpr int ATT_TMEM_COLS = 512;
constexpr int ATT_SPLIT = 112;     // keys [0,112) -> column half 0 (7 chunks of 16), [112,208) -> half 1 (6 chunks)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifndef HOIGEN_ATT_REGCAP_THREADS
#define HOIGEN_ATT_REGCAP_THREADS ATT_THREADS     // 448: <= 144 registers (see HOIGEN_GEMM2_REGCAP_THREADS in gemm.cu)
#endif
__global__ void __launch_bounds__(HOIGEN_ATT_REGCAP_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ /* box 64 x 128 x 1 */,
                 const __grid_constant__ CUtensorMap tmKV /* box 64 x 208 x 1 */,
                 const __grid_constant__ CUtensorMap tmO /* box 64 x 128 x 1 on the (B,197,768) output */,
                 int num_items, long long* __restrict__ trace /* diagnostics: CTA 0 phase timestamps, or nullptr */) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + ATT_SMEM_MISC);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t bar_full = bar0;            // [2]  Q + K of a stage landed
  const uint32_t bar_empty = bar0 + 16;      // [2]  Q + K consumed (S MMA done)  -> reload two items ahead, early
  const uint32_t bar_sready = bar0 + 32;     // [2]
  const uint32_t bar_epi = bar0 + 48;        // [2]
  const uint32_t bar_pready = bar0 + 64;
  const uint32_t bar_oready = bar0 + 72;
  const uint32_t bar_vfull = bar0 + 80;      // [2]  V of a stage landed
  const uint32_t bar_vempty = bar0 + 96;     // [2]  V consumed (PV MMA done)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* stat_max = reinterpret_cast<float*>(sm + ATT_SMEM_MISC + 128);   // [2 item parities][2 halves][128 rows]
  float* stat_sum = stat_max + 512;                                       // [2 item parities][2 halves][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
      mbar_init(bar_sready + 8u * s, 1);
      mbar_init(bar_vfull + 8u * s, 1);
      mbar_init(bar_vempty + 8u * s, 1);
      mbar_init(bar_epi + 8u * s, 8);     // one arrive per softmax warp
    }
    mbar_init(bar_pready, 8);
    mbar_init(bar_oready, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // item -> (image b, head h, row tile mt); consecutive items of a CTA stride by gridDim.x
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_local = first < num_items ? (num_items - first + stride - 1) / stride : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int item = first + i * stride;
        const int mt = item & 1, h = (item >> 1) % ATT_HEADS, b = (item >> 1) / ATT_HEADS;
        const int s = i & 1;
        const uint32_t st = base + s * ATT_STAGE_BYTES;
        // Q and K are released as soon as S_{i-2} has been computed (early), V only after P V of item i-2
        mbar_wait(bar_empty + 8u * s, ((i >> 1) & 1u) ^ 1u);
        const uint32_t full = bar_full + 8u * s;
        mbar_arrive_expect_tx(full, 16384 + ATT_KEYS * 128);
        tma_load_3d(st + ATT_STAGE_Q, &tmQ, full, h * ATT_DH, mt * 128, b);
        tma_load_3d(st + ATT_STAGE_K, &tmKV, full, ATT_WIDTH + h * ATT_DH, 0, b);
        mbar_wait(bar_vempty + 8u * s, ((i >> 1) & 1u) ^ 1u);
        const uint32_t vfull = bar_vfull + 8u * s;
        mbar_arrive_expect_tx(vfull, ATT_KEYS * 128);
        tma_load_3d(st + ATT_STAGE_V, &tmKV, vfull, 2 * ATT_WIDTH + h * ATT_DH, 0, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && n_local > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_KEYS);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, ATT_DH, /*a_mn_major=*/0, /*b_mn_major=*/1);
      auto issue_s = [&](int i) {
        const int s = i & 1;
        mbar_wait(bar_full + 8u * s, (i >> 1) & 1u);
        if (i >= 2) mbar_wait(bar_epi + 8u * s, ((i >> 1) - 1) & 1u);   // TMEM buffer s drained by item i-2
        tc_fence_after();
        const uint32_t st = base + s * ATT_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < ATT_DH / 16; ++k)
          umma_bf16_ss(tmem + uint32_t(s * 256), make_sdesc_sw128(st + ATT_STAGE_Q + k * 32),
                       make_sdesc_sw128(st + ATT_STAGE_K + k * 32), idesc_s, k > 0 ? 1u : 0u);
        tc_commit(bar_sready + 8u * s);
        tc_commit(bar_empty + 8u * s);     // Q, K of stage s free for item i+2
      };
      issue_s(0);
      for (int i = 0; i < n_local; ++i) {
        if (i + 1 < n_local) issue_s(i + 1);       // S of the next item overlaps the softmax of this one
        const int s = i & 1;
        mbar_wait(bar_pready, i & 1u);
        mbar_wait(bar_vfull + 8u * s, (i >> 1) & 1u);
        tc_fence_after();
        const uint32_t st = base + s * ATT_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < ATT_KEYS / 16; ++ks) {
          // A = P from TMEM (16 keys = 8 packed columns per k-step); V tile rows are keys (MN-major B operand): one
          // k-step = 16 keys = two 8-row groups = 2048 bytes
          const uint64_t bdesc = make_sdesc_sw128(st + ATT_STAGE_V + ks * 2048);
          umma_bf16_ts(tmem + uint32_t(s * 256 + ATT_O_COL), tmem + uint32_t(s * 256 + ks * 8), bdesc, idesc_o, ks > 0 ? 1u : 0u);
        }
        tc_commit(bar_oready);             // O_i complete (and P consumed)
        tc_commit(bar_vempty + 8u * s);    // V of stage s free for item i+2
      }
    }
  } else {
    // ===================== softmax + epilogue (256 threads, two per query row) =====================
    const int quad = warp & 3;
    const int colhalf = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int c_begin = colhalf == 0 ? 0 : ATT_SPLIT / 16;            // 16-column chunk range of this thread
    const int c_end = colhalf == 0 ? ATT_SPLIT / 16 : ATT_KEYS / 16;
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // 64^-0.5 * log2(e)
    uint8_t* smO = sm + ATT_SMEM_O;
    const bool tracing = trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64;
    auto stamp = [&](int i, int k) { if (tracing && i < 16) trace[i * 8 + k] = clock64(); };

    // O_i / rowsum -> bf16 -> 128B-swizzled staging tile [128 rows x 64 cols] -> one TMA store (rows >= 197 clipped)
    auto epilogue = [&](int i) {
      const int item = first + i * stride;
      const int mt = item & 1, h = (item >> 1) % ATT_HEADS, b = (item >> 1) / ATT_HEADS;
      const float inv = 1.0f / (stat_sum[(i & 1) * 256 + row] + stat_sum[(i & 1) * 256 + 128 + row]);
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem + (uint32_t(quad * 32) << 16) + uint32_t((i & 1) * 256 + ATT_O_COL + colhalf * 32), r);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_epi + 8u * (i & 1));     // TMEM buffer drained: S of item i+2 may be issued
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
        *reinterpret_cast<uint4*>(smO + sw128_offset(row, uint32_t(colhalf * 4 + (j >> 3)))) = pk;
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (threadIdx.x == 64) {
        tma_store_3d(&tmO, smem_u32(smO), h * ATT_DH, mt * 128, b);
        tma_store_commit();
      }
    };

    // exponentials of chunks [cc0, cc1) of this thread's half row -> packed bf16 pairs -> TMEM (P overlays S)
    float sum4[4];
    auto exp_chunks = [&](uint32_t (&r)[7][16], int cc0, int cc1, float mxs, uint32_t p_row) {
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        if (cc >= cc0 && cc < cc1 && cc < c_end - c_begin) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = ex2_approx(fmaf(__uint_as_float(r[cc][2 * j]), scale_log2, -mxs));
            const float e1 = ex2_approx(fmaf(__uint_as_float(r[cc][2 * j + 1]), scale_log2, -mxs));
            sum4[j & 3] += e0 + e1;
            pk[j] = pack_bf16x2(e0, e1);
          }
          tmem_st_32x32b_x8(p_row + uint32_t((c_begin + cc) * 8), pk);   // keys 16c .. 16c+15 -> columns 8c .. 8c+7
        }
      }
    };

    for (int i = 0; i < n_local; ++i) {
      const int sb = i & 1;
      const uint32_t t_row = tmem + (uint32_t(quad * 32) << 16) + uint32_t(sb * 256);
      stamp(i, 0);
      mbar_wait(bar_sready + 8u * sb, (i >> 1) & 1u);
      tc_fence_after();
      stamp(i, 1);
      // ---- this thread's half row of S (7 or 6 chunks of 16 fp32) -> registers with ONE TMEM round trip ----
      uint32_t r[7][16];
#pragma unroll
      for (int cc = 0; cc < 7; ++cc)
        if (cc < c_end - c_begin) tmem_ld_32x32b_x16(t_row + (c_begin + cc) * 16, r[cc]);
      tmem_wait_ld();
      // keys 197..207 of the last chunk are padding: force them to -inf once, so no per-element masking is needed
      if (colhalf == 1) {
#pragma unroll
        for (int j = ATT_TOKENS - 192; j < 16; ++j) r[5][j] = 0xff800000u;   // chunk 12 = keys 192..207
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains, not one of 112
#pragma unroll
      for (int cc = 0; cc < 7; ++cc) {
        if (cc < c_end - c_begin) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(r[cc][j]));
        }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      stat_max[sb * 256 + colhalf * 128 + row] = mx;
      if (threadIdx.x == 64) tma_store_wait_read();   // the O staging tile has been read by the previous item's store
      stamp(i, 2);
      // after this barrier: the partner's row max is visible, every thread holds its S (the columns are dead and may
      // take P), the partner's row sum of item i-1 is visible, and the staging tile is free
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tc_fence_after();
      stamp(i, 3);
      mx = fmaxf(mx, stat_max[sb * 256 + (colhalf ^ 1) * 128 + row]);
      const float mxs = mx * scale_log2;
      sum4[0] = sum4[1] = sum4[2] = sum4[3] = 0.f;
      exp_chunks(r, 0, 3, mxs, t_row);
      stamp(i, 4);
      if (i > 0) {
        mbar_wait(bar_oready, (i - 1) & 1u);
        tc_fence_after();
        epilogue(i - 1);
      }
      stamp(i, 5);
      exp_chunks(r, 3, 7, mxs, t_row);
      stat_sum[sb * 256 + colhalf * 128 + row] = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      tmem_wait_st();                 // P is in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pready);
      stamp(i, 6);
    }
    if (n_local > 0) {
      mbar_wait(bar_oready, (n_local - 1) & 1u);
      tc_fence_after();
      if (threadIdx.x == 64) tma_store_wait_read();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // partner's partial row sum of the last item is visible; staging free
      epilogue(n_local - 1);
      if (threadIdx.x == 64) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, ATT_TMEM_COLS);
  }
}

}  // namespace hoigen

extern "C" {

static int attention_launch(const void* qkv, void* out, int32_t batch, long long* trace, hoigen_stream_t stream);

int hoigen_attention(const void* qkv, void* out, int32_t batch, hoigen_stream_t stream) {
  return attention_launch(qkv, out, batch, nullptr, stream);
}

/* diagnostics: same launch, CTA 0's softmax-leader thread writes clock64() stamps [16 items][8 phases] to trace */
int hoigen_debug_attention_trace(const void* qkv, void* out, int32_t batch, int64_t* trace, hoigen_stream_t stream) {
  return attention_launch(qkv, out, batch, reinterpret_cast<long long*>(trace), stream);
}

static int attention_launch(const void* qkv, void* out, int32_t batch, long long* trace, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(qkv && out && batch > 0, "attention: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "attention: buffers must be 16-byte aligned");
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(attention_kernel), ATT_SMEM_BYTES));
  const uint64_t row_bytes = 3ull * ATT_WIDTH * 2;
  const CUtensorMap* tq = get_tmap_3d_bf16(qkv, 3 * ATT_WIDTH, ATT_TOKENS, uint64_t(batch), row_bytes,
                                           row_bytes * ATT_TOKENS, 64, 128, 1);
  if (!tq) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tkv = get_tmap_3d_bf16(qkv, 3 * ATT_WIDTH, ATT_TOKENS, uint64_t(batch), row_bytes,
                                            row_bytes * ATT_TOKENS, 64, ATT_KEYS, 1);
  if (!tkv) return HOIGEN_ERR_CUDA;
  const uint64_t orow_bytes = uint64_t(ATT_WIDTH) * 2;
  const CUtensorMap* to = get_tmap_3d_bf16(out, ATT_WIDTH, ATT_TOKENS, uint64_t(batch), orow_bytes, orow_bytes * ATT_TOKENS,
                                           64, 128, 1);
  if (!to) return HOIGEN_ERR_CUDA;
  const int items = batch * ATT_HEADS * 2;
  const int grid = items < num_sms() ? items : num_sms();
  KernelScope ks("attention", reinterpret_cast<cudaStream_t>(stream), 4.0 * ATT_TOKENS * ATT_TOKENS * ATT_DH * ATT_HEADS * batch,
                 double(batch) * ATT_TOKENS * ATT_WIDTH * 2 * 4);
  attention_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
      *tq, *tkv, *to, items, trace);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
