// RoIAlign(7x7, adaptive, aligned) + mean over the 49 bins as a TENSOR-CORE product (replaces the
// torchvision.ops.roi_align call sites U:1028-1029 + the mean U:1032-1037):
//
//     feature[job, c] = inv_count[job] * sum_t  (Wy[job, t / 14] * Wx[job, t % 14]) * token[t, c]         t < 196
//
// i.e. per image a (n + K) x 196 weight matrix times the 196 x 512 token map.  The SIMT form of this (hoi_head.cu) walks
// the bounding window of four boxes per warp and is FP32-FMA-issue bound (68 us per 64-image batch for 43 MB of
// compulsory traffic).  Here both operands are split into THREE bf16 planes (hi + mid + lo = the fp32 value to 2^-24; every
// bf16 x bf16 product is exact in the fp32 accumulator) and the six significant cross terms run on tcgen05:
//
//     out^T[c, job] = sum_{(a,b) in {hh, hm, mh, hl, lh, mm}}  F_a^T[c, t] . W_b[job, t]^T
//
// with M = channels (two 128-row tiles per CTA), N = the image's jobs (<= 192 per pass, a multiple of 16: no padding to a
// power of two), K = tokens in chunks of 64.  grid = (image, channel half).  Per token chunk every thread builds the bf16
// planes straight into 128B-swizzled K-major shared-memory tiles (token planes transposed on the way: a thread owns one
// channel and packs 8 consecutive tokens into one 16-byte store; weight planes from the per-axis weights the CTA
// computes itself from the boxes), one thread issues 2 x 6 x 4 UMMA 128 x N x 16, and the epilogue scales by 1 / (49 count) and writes
// the feature rows with coalesced 128-byte stores.
#include "common.h"
#include "ptx.cuh"
#include "roi_common.cuh"

namespace hoigen {

constexpr int RT_THREADS = 512;                  // 16 warps: the plane construction is ALU work (3-way splits), issue-bound with fewer
constexpr int RT_MAXJOBS = 192;                  // jobs (N of the MMA) per pass
constexpr int RT_A_TILE = 128 * 128;             // 16 KiB: [128 channels x 64 tokens] bf16
constexpr int RT_B_TILE = RT_MAXJOBS * 128;      // 24 KiB: [<= 192 jobs x 64 tokens] bf16
constexpr int RT_SMEM_A = 0;                     // [2 M-tiles][3 planes]
constexpr int RT_SMEM_B = 6 * RT_A_TILE;         // [3 planes]
constexpr int RT_SMEM_W = RT_SMEM_B + 3 * RT_B_TILE;          // axis weights of the pass's jobs: [192][32] fp32
constexpr int RT_SMEM_BAR = RT_SMEM_W + RT_MAXJOBS * 128;
constexpr int RT_SMEM_BYTES = RT_SMEM_BAR + 64 + 1024;

// two fp32 values -> their (hi, mid, lo) bf16 pairs; packed converts (cvt.rn.bf16x2.f32), a bf16 -> fp32 is a shift
__device__ __forceinline__ void split3x2(float x0, float x1, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  hi = pack_bf16x2(x0, x1);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  mid = pack_bf16x2(r0, r1);
  lo = pack_bf16x2(r0 - __uint_as_float(mid << 16), r1 - __uint_as_float(mid & 0xffff0000u));
}

// eight fp32 values -> the three planes' 16-byte chunks
__device__ __forceinline__ void split8(const float (&v)[8], uint4& h, uint4& m, uint4& l) {
  split3x2(v[0], v[1], h.x, m.x, l.x);
  split3x2(v[2], v[3], h.y, m.y, l.y);
  split3x2(v[4], v[5], h.z, m.z, l.z);
  split3x2(v[6], v[7], h.w, m.w, l.w);
}

__global__ void __launch_bounds__(RT_THREADS, 1)
roi_tc_kernel(const float* __restrict__ tokens /* (B*197, 512) */, const float* __restrict__ boxes /* (Ntot, 4) xyxy */,
              const int* __restrict__ box_off, const int* __restrict__ pair_off, float spatial_scale,
              float* __restrict__ single_feat, float* __restrict__ union_feat) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_addr);
  float* s_w = reinterpret_cast<float*>(sm + RT_SMEM_W);
  const uint32_t bar = smem_u32(sm + RT_SMEM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + RT_SMEM_BAR + 16);

  const int b = blockIdx.x, chalf = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bbase = box_off[b], n = box_off[b + 1] - bbase;
  const int pbase = pair_off[b], K = pair_off[b + 1] - pbase;
  const int njobs = n + K;
  if (njobs == 0) return;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
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
  uint32_t phase = 0;

  // this thread's share of the A planes: channel c = tid % 256 (M-tile c / 128, row c % 128), token groups 4 gh .. 4 gh + 3
  const int a_c = threadIdx.x & 255, a_gh = threadIdx.x >> 8;
  const int a_mt = a_c >> 7, a_row = a_c & 127;
  const float* tok_col = tokens + (size_t(b) * 197 + 1) * 512 + chalf * 256 + a_c;   // token t at + t * 512

  for (int j0 = 0; j0 < njobs; j0 += RT_MAXJOBS) {
    const int nj = min(RT_MAXJOBS, njobs - j0);
    const int NT = (nj + 15) & ~15;                       // MMA N (multiple of 16)
    // ---- the pass's per-axis RoIAlign weights, computed here from the boxes (row r: Wy[0..14) at 0, Wx[0..14) at 16,
    //      1 / (49 count) at 31; single boxes first, then the unions min / max of (human x, box y), U:1007-1023) ----
    //      one owner thread per (row, axis): 2 NT <= 384 tasks in one round of the 512 threads
    for (int i = threadIdx.x; i < NT * 2; i += RT_THREADS) {
      const int r = i >> 1, axis = i & 1, job = j0 + r;
      float* row = s_w + r * 32 + axis * 16;
      if (r >= nj) {
#pragma unroll
        for (int t = 0; t < 16; ++t) row[t] = 0.f;
        continue;
      }
      float4 bx;
      if (job < n) {
        bx = __ldg(reinterpret_cast<const float4*>(boxes) + bbase + job);
      } else {
        const int li = job - n;
        const int px = li / (n - 1), rr = li % (n - 1);
        const int py = rr < px ? rr : rr + 1;           // row-major enumeration of (x, y != x), x < n_h
        const float4 bh = __ldg(reinterpret_cast<const float4*>(boxes) + bbase + px);
        const float4 bo = __ldg(reinterpret_cast<const float4*>(boxes) + bbase + py);
        bx = make_float4(fminf(bh.x, bo.x), fminf(bh.y, bo.y), fmaxf(bh.z, bo.z), fmaxf(bh.w, bo.w));
      }
      // same expressions as roi_axis_entry: extent = (end * scale - 0.5) - (start * scale - 0.5), g = ceil(extent / 7)
      const float sx = bx.x * spatial_scale - 0.5f, sy = bx.y * spatial_scale - 0.5f;
      const float rw = (bx.z * spatial_scale - 0.5f) - sx, rh = (bx.w * spatial_scale - 0.5f) - sy;
      const int gw = int(ceilf(rw / float(POOL))), gh = int(ceilf(rh / float(POOL)));
      if (axis == 0) axis_weights_row(row, sy, rh / float(POOL), gh);
      else           axis_weights_row(row, sx, rw / float(POOL), gw);
      row[14] = 0.f;
      row[15] = axis ? 1.0f / (float(max(gh * gw, 1)) * float(POOL * POOL)) : 0.f;
    }
    __syncthreads();
    // this thread's 32 tokens of chunk kc -> registers (32 loads in flight per thread; tokens >= 196 are zero)
    float tv[4][8];
    auto load_tokens = [&](int kc) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int t = kc * 64 + (a_gh * 4 + g) * 8 + i;
          tv[g][i] = t < 196 ? __ldg(tok_col + size_t(t) * 512) : 0.f;
        }
      }
    };
    load_tokens(0);
    for (int kc = 0; kc < 4; ++kc) {
      // ---- A planes: token map slice, transposed to K-major [channel][token] ----
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 h, m, l;
        split8(tv[g], h, m, l);
        const uint32_t off = sw128_offset(uint32_t(a_row), uint32_t(a_gh * 4 + g));
        uint8_t* at = sm + RT_SMEM_A + a_mt * 3 * RT_A_TILE + off;
        *reinterpret_cast<uint4*>(at) = h;
        *reinterpret_cast<uint4*>(at + RT_A_TILE) = m;
        *reinterpret_cast<uint4*>(at + 2 * RT_A_TILE) = l;
      }
      // the next chunk's tokens are fetched while this chunk's weight planes are built and its MMAs run
      if (kc + 1 < 4) load_tokens(kc + 1);
      // ---- B planes: bilinear weight matrix rows W[job][t] = Wy[t / 14] * Wx[t % 14] (rows >= nj are zero) ----
      for (int i = threadIdx.x; i < NT * 8; i += RT_THREADS) {
        const int r = i >> 3, g = i & 7;
        const int t0 = kc * 64 + g * 8;
        uint4 h = make_uint4(0, 0, 0, 0), m = h, l = h;
        if (r < nj && t0 < 196) {
          const float* wr = s_w + r * 32;
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int t = t0 + q;
            v[q] = t < 196 ? wr[t / 14] * wr[16 + t % 14] : 0.f;
          }
          split8(v, h, m, l);
        }
        const uint32_t off = sw128_offset(uint32_t(r), uint32_t(g));
        uint8_t* bt = sm + RT_SMEM_B + off;
        *reinterpret_cast<uint4*>(bt) = h;
        *reinterpret_cast<uint4*>(bt + RT_B_TILE) = m;
        *reinterpret_cast<uint4*>(bt + 2 * RT_B_TILE) = l;
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, NT);
        const int ksteps = kc < 3 ? 4 : 1;                 // tokens 192..195 live in the first k-step of the last chunk
        // (A plane, B plane) of the six cross terms down to 2^-16 of the leading one: hh, hm, mh, hl, lh, mm
        const int pa[6] = {0, 0, 1, 0, 2, 1}, pb[6] = {0, 1, 0, 2, 0, 1};
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int term = 0; term < 6; ++term) {
            const uint32_t aaddr = base + RT_SMEM_A + (mt * 3 + pa[term]) * RT_A_TILE;
            const uint32_t baddr = base + RT_SMEM_B + pb[term] * RT_B_TILE;
            for (int k = 0; k < ksteps; ++k)
              umma_bf16_ss(tmem + uint32_t(mt * 256), make_sdesc_sw128(aaddr + k * 32), make_sdesc_sw128(baddr + k * 32), idesc,
                           (kc > 0 || term > 0 || k > 0) ? 1u : 0u);
          }
        }
        tc_commit(bar);
      }
      mbar_wait(bar, phase);                               // the MMAs have read the planes: the next chunk may overwrite them
      phase ^= 1u;
      tc_fence_after();
    }
    // ---- epilogue: D^T [channel lane][job column] * inv_count -> feature rows (a warp writes 32 consecutive channels) ----
    {
      const int q = warp & 3, mt = (warp >> 2) & 1, chf = warp >> 3;      // TMEM lane quadrant, M-tile, half of the job columns
      const int ch = chalf * 256 + mt * 128 + q * 32 + lane;
      const uint32_t t_addr = tmem + (uint32_t(q * 32) << 16) + uint32_t(mt * 256);
      const int cbeg = chf * (NT / 2), cend = cbeg + NT / 2;              // NT / 2 is a multiple of 8
      for (int c0 = cbeg; c0 < cend; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(t_addr + uint32_t(c0))
                     : "memory");
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int rj = c0 + j;
          if (rj < nj) {
            const int job = j0 + rj;
            float* dst = job < n ? single_feat + size_t(bbase + job) * 512 : union_feat + size_t(pbase + job - n) * 512;
            dst[ch] = __uint_as_float(r[j]) * s_w[rj * 32 + 31];
          }
        }
      }
      tc_fence_before();
    }
    __syncthreads();                                       // s_w and the accumulators are reused by the next pass
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_roi_features_tc(const float* tokens, const float* boxes, const int* box_off, const int* pair_off, int batch,
                           float spatial_scale, float* single_feat, float* union_feat, cudaStream_t s) {
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(roi_tc_kernel), RT_SMEM_BYTES));
  roi_tc_kernel<<<dim3(batch, 2), RT_THREADS, RT_SMEM_BYTES, s>>>(tokens, boxes, box_off, pair_off, spatial_scale, single_feat, union_feat);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // namespace hoigen
