// 12-head self-attention over the 197-token sequence as ONE tcgen05 pass per (image, head, 128-row tile):
//   S = Q K^T  (UMMA 128 x 208 x 16, 4 k-steps, fp32 in TMEM)  ->  softmax in registers (tcgen05.ld)
//   -> P (bf16) into 128B-swizzled shared memory  ->  O = P V  (UMMA 128 x 64 x 16, 13 k-steps,
//   V consumed MN-major straight from the QKV buffer, no transpose)  ->  O / rowsum -> bf16.
// The whole key range (197 -> 208) fits one tile, so no online-softmax rescaling is needed.
//
// Shared memory (90 KiB => 2 CTAs / SM; P overlays Q and K once S has been produced):
//   [0,16K)  Q tile / P atom 0     [16K,48K) K tile (26 KiB used) / P atoms 1,2     [48K,64K) P atom 3
//   [64K,90K) V tile
// TMEM: 256 columns (S uses 208; O overlays columns [0,64) after the softmax has consumed S).
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
constexpr int ATT_THREADS = 128;
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_K = 16384;
constexpr int ATT_SMEM_P3 = 49152;
constexpr int ATT_SMEM_V = 65536;
constexpr int ATT_SMEM_TILES = 65536 + ATT_KEYS * 128;  // 92160
constexpr int ATT_SMEM_BYTES = ATT_SMEM_TILES + 1024 + 64;
constexpr int ATT_TMEM_COLS = 256;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ /* box 64 x 128 x 1 */,
                 const __grid_constant__ CUtensorMap tmKV /* box 64 x 208 x 1 */, __nv_bfloat16* __restrict__ out,
                 int batch) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + ATT_SMEM_TILES);
  const uint32_t bar_load = smem_u32(bars);
  const uint32_t bar_s = bar_load + 8;
  const uint32_t bar_o = bar_load + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x & 1;            // 128-row tile of the 197 queries
  const int h = (blockIdx.x >> 1) % ATT_HEADS;
  const int b = (blockIdx.x >> 1) / ATT_HEADS;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    // ---- loads: Q (128 x 64), K (208 x 64), V (208 x 64); rows past token 196 are zero-filled by TMA ----
    mbar_arrive_expect_tx(bar_load, 128 * 128 + 2 * ATT_KEYS * 128);
    tma_load_3d(base + ATT_SMEM_Q, &tmQ, bar_load, h * ATT_DH, mt * 128, b);
    tma_load_3d(base + ATT_SMEM_K, &tmKV, bar_load, ATT_WIDTH + h * ATT_DH, 0, b);
    tma_load_3d(base + ATT_SMEM_V, &tmKV, bar_load, 2 * ATT_WIDTH + h * ATT_DH, 0, b);
    mbar_wait(bar_load, 0);
    tc_fence_after();
    // ---- S = Q K^T ----
    constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_KEYS);
#pragma unroll
    for (int k = 0; k < ATT_DH / 16; ++k) {
      umma_bf16_ss(tmem, make_sdesc_sw128(base + ATT_SMEM_Q + k * 32), make_sdesc_sw128(base + ATT_SMEM_K + k * 32),
                   idesc_s, k > 0 ? 1u : 0u);
    }
    tc_commit(bar_s);
  }

  // ---- softmax: thread r owns query row r of the tile ----
  mbar_wait(bar_s, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  const uint32_t t_row = tmem + (uint32_t(warp * 32) << 16);
  const float scale_log2 = 0.125f * 1.4426950408889634f;  // 64^-0.5 * log2(e)
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < ATT_KEYS / 16; ++c) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c * 16, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c * 16 + j < ATT_TOKENS) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  const float mxs = mx * scale_log2;
  float sum = 0.f;
#pragma unroll 1
  for (int c = 0; c < ATT_KEYS / 16; ++c) {
    uint32_t r[16];
    tmem_ld_32x32b_x16(t_row + c * 16, r);
    tmem_wait_ld();
    float p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float e = exp2f(fmaf(__uint_as_float(r[j]), scale_log2, -mxs));
      p[j] = (c * 16 + j < ATT_TOKENS) ? e : 0.f;
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // the row sum is taken over the bf16-rounded probabilities that the PV MMA actually consumes
      const __nv_bfloat162 v2 = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
      sum += __low2float(v2) + __high2float(v2);
      pk[j] = *reinterpret_cast<const uint32_t*>(&v2);
    }
    // P[row][c*16 .. c*16+15] -> K-major SW128 atoms: atom = c/4 (64 keys each), 16-byte chunks (c%4)*2, +1
    const int atom = c >> 2;
    const uint32_t atom_off = (atom == 0) ? ATT_SMEM_Q : (atom == 3 ? ATT_SMEM_P3 : ATT_SMEM_K + (atom - 1) * 16384);
    uint8_t* pa = tiles + atom_off;
    const uint32_t ch = uint32_t(c & 3) * 2;
    *reinterpret_cast<uint4*>(pa + sw128_offset(row, ch)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(pa + sw128_offset(row, ch + 1)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  // make the generic-proxy smem writes visible to the tensor core (async proxy), and order the TMEM reads of S
  // before the PV MMA overwrites columns [0,64)
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_o = make_idesc_bf16(128, ATT_DH, /*a_mn_major=*/0, /*b_mn_major=*/1);
#pragma unroll
    for (int ks = 0; ks < ATT_KEYS / 16; ++ks) {
      const int atom = ks >> 2;
      const uint32_t atom_off = (atom == 0) ? ATT_SMEM_Q : (atom == 3 ? ATT_SMEM_P3 : ATT_SMEM_K + (atom - 1) * 16384);
      const uint64_t adesc = make_sdesc_sw128(base + atom_off + (ks & 3) * 32);
      // V tile rows are keys (MN-major B operand): one k-step = 16 keys = two 8-row groups = 2048 bytes
      const uint64_t bdesc = make_sdesc_sw128(base + ATT_SMEM_V + ks * 2048);
      umma_bf16_ss(tmem, adesc, bdesc, idesc_o, ks > 0 ? 1u : 0u);
    }
    tc_commit(bar_o);
  }

  // ---- epilogue: O / rowsum -> bf16 -> out[b*197 + t][h*64 .. +63] ----
  mbar_wait(bar_o, 0);
  tc_fence_after();
  const int t = mt * 128 + row;
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < ATT_DH / 32; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_row + c * 32, r);
    tmem_wait_ld();
    if (t < ATT_TOKENS) {
      __nv_bfloat16* dst = out + (size_t(b) * ATT_TOKENS + t) * ATT_WIDTH + h * ATT_DH + c * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(r[j + 4]) * inv, __uint_as_float(r[j + 5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(r[j + 6]) * inv, __uint_as_float(r[j + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + j) = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, ATT_TMEM_COLS);
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_attention(const void* qkv, void* out, int32_t batch, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(qkv && out && batch > 0, "attention: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                   "attention: buffers must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    HOIGEN_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    attr_set = true;
  }
  const uint64_t row_bytes = 3ull * ATT_WIDTH * 2;
  const CUtensorMap* tq = get_tmap_3d_bf16(qkv, 3 * ATT_WIDTH, ATT_TOKENS, uint64_t(batch), row_bytes,
                                           row_bytes * ATT_TOKENS, 64, 128, 1);
  if (!tq) return HOIGEN_ERR_CUDA;
  const CUtensorMap* tkv = get_tmap_3d_bf16(qkv, 3 * ATT_WIDTH, ATT_TOKENS, uint64_t(batch), row_bytes,
                                            row_bytes * ATT_TOKENS, 64, ATT_KEYS, 1);
  if (!tkv) return HOIGEN_ERR_CUDA;
  KernelScope ks("attention", reinterpret_cast<cudaStream_t>(stream), 4.0 * ATT_TOKENS * ATT_TOKENS * ATT_DH * ATT_HEADS * batch,
                 double(batch) * ATT_TOKENS * ATT_WIDTH * 2 * 4);
  attention_kernel<<<batch * ATT_HEADS * 2, ATT_THREADS, ATT_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
      *tq, *tkv, reinterpret_cast<__nv_bfloat16*>(out), batch);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
