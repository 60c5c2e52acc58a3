// Stem of the ResNet-50 branch (torchvision resnet50.conv1 + bn1 + relu behind U:1616): the 7x7 / stride 2 /
// pad 3 convolution of (B, 3, H, W) fp32 images (224 x 224 for the DINO branch, any padded size for DETR's backbone) ->
// (B, ceil(H/2), ceil(W/2), 64) bf16 NHWC rows, BatchNorm folded, ReLU applied -- as
// ONE tensor-core kernel with no im2col matrix in memory (the two-step form, hoigen_stem_im2col + GEMM, writes and re-reads
// 257 MB per 64 images for 103 MB of output).
//
// Persistent CTAs walk (image, group of four output rows, block of 128 output columns) items.  Per item the 13 input rows it needs are staged once in shared
// memory, channel-interleaved and zero-padded ([row][col -3 .. 226][c] bf16), so the 21 taps (kx, c) of one (pixel, ky) are
// contiguous there.  Per output row the threads copy those runs into a K-major 128B-swizzled A tile (128 pixel slots x K = 192,
// k = ky * 24 + kx * 3 + c: runs padded to 24 so every run is three aligned 16-byte chunks; the pad taps read real neighbouring
// pixels and meet zero weights), one thread issues 11 UMMA 128x64x16 against the resident weight tile, and four warps turn the
// previous row's accumulator (TMEM, double-buffered) into bias + ReLU + bf16 and store it while the next tile is being built.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "ptx.cuh"

namespace hoigen {

constexpr int SC_THREADS = 256;
constexpr int SC_ROWS = 4;                       // output rows per work item
constexpr int SC_IN_ROWS = 2 * SC_ROWS + 5;      // 13 input rows: 2 oy - 3 .. 2 oy + 3 for the four oy
constexpr int SC_COLS = 2 * 128 + 6;             // staged input columns of a 128-pixel block: 2 ox0 - 3 .. 2 ox0 + 258
constexpr int SC_SROW = SC_COLS * 3;             // staged row: 3 channels interleaved (786 bf16)
constexpr int SC_K = 192;                        // 7 runs of 24 (21 taps + 3 pad) = 168, padded to three 64-wide k-blocks
constexpr int SC_KSTEPS = 11;                    // 176 >= 168 columns carry data
constexpr int SC_A_TILE = 3 * 16384;             // 128 rows x 192 k bf16
constexpr int SC_W_TILE = 3 * 8192;              // 64 rows x 192 k bf16
constexpr int SC_SMEM_W = 0;
constexpr int SC_SMEM_A = SC_W_TILE;             // two A buffers
constexpr int SC_SMEM_IN = SC_SMEM_A + 2 * SC_A_TILE;
constexpr int SC_SMEM_BIAS = SC_SMEM_IN + ((SC_IN_ROWS * SC_SROW * 2 + 15) & ~15);
constexpr int SC_SMEM_BAR = SC_SMEM_BIAS + 256;
constexpr int SC_SMEM_BYTES = SC_SMEM_BAR + 64 + 1024;   // + alignment slack

__global__ void __launch_bounds__(SC_THREADS, 1)
stem_conv_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ w /* (64, 192) */, const float* __restrict__ bias,
                 __nv_bfloat16* __restrict__ out /* (B * Ho * Wo, 64) */, int batch, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(sm + SC_SMEM_IN);
  float* s_bias = reinterpret_cast<float*>(sm + SC_SMEM_BIAS);
  const uint32_t bar0 = base + SC_SMEM_BAR;                 // bar0 + 8 * buf: MMAs of the tile in A / accumulator buffer `buf` done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SC_SMEM_BAR + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 128);
    tmem_relinquish();
  }
  // resident weight tile (K-major, 128B swizzle) and bias; the never-written tail chunks of both A buffers are zeroed once
  for (int i = threadIdx.x; i < 64 * (SC_K / 8); i += SC_THREADS) {
    const int n = i / (SC_K / 8), chunk = i % (SC_K / 8);
    *reinterpret_cast<uint4*>(sm + SC_SMEM_W + (chunk >> 3) * 8192 + sw128_offset(uint32_t(n), uint32_t(chunk & 7))) =
        __ldg(reinterpret_cast<const uint4*>(w + size_t(n) * SC_K) + chunk);
  }
  for (int i = threadIdx.x; i < 2 * 128 * 3; i += SC_THREADS) {
    const int buf = i / 384, row = (i % 384) / 3, chunk = 21 + i % 3;
    *reinterpret_cast<uint4*>(sm + SC_SMEM_A + buf * SC_A_TILE + 2 * 16384 + sw128_offset(uint32_t(row), uint32_t(chunk & 7))) =
        make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = __ldg(bias + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // epilogue of one finished tile: accumulator buffer `buf` -> out rows of (b, oy); warps 0..3 = TMEM lane quadrants = pixels
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int row_groups = (Ho + SC_ROWS - 1) / SC_ROWS, col_blocks = (Wo + 127) / 128;
  auto epilogue = [&](int buf, uint32_t phase, int b, int oy, int ox0) {
    mbar_wait(bar0 + 8u * buf, phase);
    tc_fence_after();
    const int ox = ox0 + warp * 32 + lane;
    const uint32_t t_addr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(buf * 64);
    __nv_bfloat16* dst = out + ((size_t(b) * Ho + oy) * Wo + ox) * 64;
    uint32_t r[2][16];
    tmem_ld_32x32b_x16(t_addr, r[0]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      tmem_wait_ld();
      if (c + 1 < 4) tmem_ld_32x32b_x16(t_addr + uint32_t((c + 1) * 16), r[(c + 1) & 1]);
      if (ox < Wo) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v0 = fmaxf(__uint_as_float(r[c & 1][2 * j]) + s_bias[c * 16 + 2 * j], 0.f);
          const float v1 = fmaxf(__uint_as_float(r[c & 1][2 * j + 1]) + s_bias[c * 16 + 2 * j + 1], 0.f);
          const __nv_bfloat162 p = __floats2bfloat162_rn(v0, v1);
          pk[j] = *reinterpret_cast<const uint32_t*>(&p);
        }
        *reinterpret_cast<uint4*>(dst + c * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + c * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
  };

  const int items = batch * row_groups * col_blocks;
  int tt = 0;                      // tiles issued by this CTA so far (buffer = tt & 1, barrier phase = (tt >> 1) & 1)
  int prev_b = 0, prev_oy = 0, prev_ox0 = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int cb = item % col_blocks, rg = (item / col_blocks) % row_groups, b = item / (col_blocks * row_groups);
    const int oy0 = rg * SC_ROWS, ox0 = cb * 128;
    const int npix = min(128, Wo - ox0);
    const float* ibase = img + size_t(b) * 3 * H * W;
    // ---- stage input rows 2 oy0 - 3 .. 2 oy0 + 9, columns 2 ox0 - 3 .. 2 ox0 + 258 (reads coalesced along x; outside = 0) ----
    // (the previous item's last tile was built before its __syncthreads, so s_in is free)
    for (int i = threadIdx.x; i < SC_IN_ROWS * 3 * SC_COLS; i += SC_THREADS) {
      const int col = i % SC_COLS, rc = i / SC_COLS;
      const int c = rc % 3, r = rc / 3;
      const int iy = 2 * oy0 - 3 + r, ix = 2 * ox0 - 3 + col;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(ibase + (size_t(c) * H + iy) * W + ix);
      s_in[r * SC_SROW + col * 3 + c] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    const int nrows = min(SC_ROWS, Ho - oy0);
    for (int dy = 0; dy < nrows; ++dy, ++tt) {
      const int buf = tt & 1;
      if (tt >= 2) {               // the MMAs that read this A buffer two tiles ago are done
        mbar_wait(bar0 + 8u * buf, uint32_t((tt - 2) >> 1) & 1u);
      }
      // ---- A tile: pixel p, run ky = 24 consecutive staged elements starting at staged column 2 p of input row 2 dy + ky ----
      uint8_t* a_tile = sm + SC_SMEM_A + buf * SC_A_TILE;
      for (int i = threadIdx.x; i < npix * 7; i += SC_THREADS) {
        const int p = i % npix, ky = i / npix;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(s_in + (2 * dy + ky) * SC_SROW + 6 * p);
        uint32_t wds[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) wds[j] = src[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int chunk = ky * 3 + j;
          *reinterpret_cast<uint4*>(a_tile + (chunk >> 3) * 16384 + sw128_offset(uint32_t(p), uint32_t(chunk & 7))) =
              make_uint4(wds[4 * j], wds[4 * j + 1], wds[4 * j + 2], wds[4 * j + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 64);
        const uint32_t a_addr = base + SC_SMEM_A + buf * SC_A_TILE, w_addr = base + SC_SMEM_W;
#pragma unroll
        for (int k = 0; k < SC_KSTEPS; ++k)
          umma_bf16_ss(tmem + uint32_t(buf * 64), make_sdesc_sw128(a_addr + (k >> 2) * 16384 + (k & 3) * 32),
                       make_sdesc_sw128(w_addr + (k >> 2) * 8192 + (k & 3) * 32), idesc, k > 0 ? 1u : 0u);
        tc_commit(bar0 + 8u * buf);
      }
      // the previous tile's accumulator drains while this tile's MMAs run and the next tile is built
      if (tt >= 1 && warp < 4) epilogue(buf ^ 1, uint32_t((tt - 1) >> 1) & 1u, prev_b, prev_oy, prev_ox0);
      prev_b = b; prev_oy = oy0 + dy; prev_ox0 = ox0;
    }
  }
  if (tt >= 1 && warp < 4) epilogue((tt - 1) & 1, uint32_t((tt - 1) >> 1) & 1u, prev_b, prev_oy, prev_ox0);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_stem_conv_hw(const float* images, const void* w_bf16, const float* bias, void* out_bf16, int32_t batch, int32_t h, int32_t w,
                        hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(images && w_bf16 && bias && out_bf16 && batch > 0 && h > 0 && w > 0, "stem_conv: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(w_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0,
                   "stem_conv: weights / output must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(stem_conv_kernel), SC_SMEM_BYTES));
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long items = (long long)batch * ((ho + SC_ROWS - 1) / SC_ROWS) * ((wo + 127) / 128);
  HOIGEN_CHECK_ARG(items < 0x7fffffffLL, "stem_conv: batch too large");
  const int grid = items < num_sms() ? int(items) : num_sms();
  KernelScope ks("stem_conv", s, 2.0 * batch * ho * wo * 64 * 147, double(batch) * (3.0 * h * w * 4 + double(ho) * wo * 64 * 2));
  stem_conv_kernel<<<grid, SC_THREADS, SC_SMEM_BYTES, s>>>(images, reinterpret_cast<const __nv_bfloat16*>(w_bf16), bias,
                                                           reinterpret_cast<__nv_bfloat16*>(out_bf16), batch, h, w);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_stem_conv(const float* images, const void* w_bf16, const float* bias, void* out_bf16, int32_t batch, hoigen_stream_t stream) {
  return hoigen_stem_conv_hw(images, w_bf16, bias, out_bf16, batch, 224, 224, stream);
}

}  // extern "C"
