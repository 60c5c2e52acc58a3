// Row kernels of the ResNet-50 branch (reference U:1616-1618: `dino_model(images_clip)`, a torchvision
// resnet50 with fc = Identity in eval mode, followed by an L2 normalisation) and the plan runner that strings them together
// with the tcgen05 GEMM (gemm.cu: 1x1 convolutions as plain GEMMs, 3x3 / stride-1 convolutions as nine row-shifted
// accumulated products -- see hoigen_gemm_params.conv_taps).
//
// Activation layout between the convolutions: NHWC bf16 WITH a one-pixel zero halo, i.e. a row-major matrix whose rows are
// the pixels of (B, H + 2, W + 2) and whose columns are the channels.  The halo is the 3x3 convolutions' zero padding, so a
// shifted row read never needs a bounds test; every producer writes zeros on the ring.
//
// Everything here is bandwidth work (gathers, a max, a mean): 16-byte accesses, one pass, no shared memory.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"

namespace hoigen {

constexpr int STEM_K = 160;   // 7 * 7 * 3 = 147 taps, zero-padded to a multiple of 8 (16-byte TMA rows)

// images (B, 3, 224, 224) fp32 -> rows (B * 112 * 112, 160) bf16: row = output pixel (b, oy, ox) of the 7x7 / stride 2 /
// pad 3 stem convolution, column k = (ky * 7 + kx) * 3 + c.  One CTA per (image, pair of output rows): the nine input rows
// it needs are staged once in shared memory, channel-interleaved and zero-padded ([row][col -3 .. 226][c] bf16), so that the
// 21 entries (kx, c) of one (pixel, ky) are CONTIGUOUS there; the threads then emit the output as coalesced 16-byte chunks.
constexpr int STEM_SCOLS = 224 + 6;            // input columns -3 .. 226
constexpr int STEM_SROW = STEM_SCOLS * 3;      // bf16 elements per staged row
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ rows) {
  __shared__ __nv_bfloat16 s_in[9 * STEM_SROW];
  const int b = blockIdx.y, oy0 = blockIdx.x * 2;
  const float* base = img + size_t(b) * 3 * 224 * 224;
  // stage input rows 2 oy0 - 3 .. 2 oy0 + 5 (coalesced 4-byte reads along x; rows / columns outside the image are zero)
  for (int i = threadIdx.x; i < 9 * 3 * STEM_SCOLS; i += 256) {
    const int col = i % STEM_SCOLS, rc = i / STEM_SCOLS;
    const int c = rc % 3, r = rc / 3;
    const int iy = 2 * oy0 - 3 + r, ix = col - 3;
    float v = 0.f;
    if (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) v = __ldg(base + (size_t(c) * 224 + iy) * 224 + ix);
    s_in[r * STEM_SROW + col * 3 + c] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  __nv_bfloat16* out = rows + (size_t(b) * 112 + oy0) * 112 * STEM_K;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  for (int i = threadIdx.x; i < 2 * 112 * (STEM_K / 8); i += 256) {
    const int g8 = i % (STEM_K / 8), pix = i / (STEM_K / 8);
    const int ox = pix % 112, dy = pix / 112;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g8 * 8 + e;
      const int ky = k / 21, r = k - ky * 21;
      v[e] = k < 147 ? s_in[(2 * dy + ky) * STEM_SROW + 6 * ox + r] : zero;
    }
    *reinterpret_cast<uint4*>(out + size_t(pix) * STEM_K + g8 * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// The same im2col for any image size (DETR's backbone runs the identical stem on up to 800 x 1333 inputs): one thread per 8
// columns of one output pixel of (B, ceil(H / 2), ceil(W / 2)); reads hit L1 / L2 (neighbouring pixels share 5 of 7 columns).
__global__ void stem_im2col_hw_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ rows, int batch, int H, int W) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * Ho * Wo * (STEM_K / 8);
  if (gid >= total) return;
  const int g8 = int(gid % (STEM_K / 8));
  const long long pix = gid / (STEM_K / 8);
  const int ox = int(pix % Wo), oy = int((pix / Wo) % Ho), b = int(pix / ((long long)Wo * Ho));
  const float* base = img + size_t(b) * 3 * H * W;
  uint32_t packed[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = g8 * 8 + e * 2 + h;
      float v = 0.f;
      if (k < 147) {
        const int tap = k / 3, c = k - tap * 3;
        const int ky = tap / 7, kx = tap - ky * 7;
        const int iy = 2 * oy - 3 + ky, ix = 2 * ox - 3 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(base + (size_t(c) * H + iy) * W + ix);
      }
      v2[h] = v;
    }
    const __nv_bfloat162 p = __floats2bfloat162_rn(v2[0], v2[1]);
    packed[e] = *reinterpret_cast<const uint32_t*>(&p);
  }
  *reinterpret_cast<uint4*>(rows + pix * STEM_K + g8 * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

// 3x3 / stride 2 / pad 1 max pooling: in (B, h, w, c) bf16 WITHOUT halo (the stem GEMM's rows) -> out (B, ho + 2, wo + 2, c),
// ho = ceil(h / 2), wo = ceil(w / 2), with the zero halo written.  One thread per (output pixel incl. ring, 8 channels).
__global__ void maxpool3x3s2_halo_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int batch, int h,
                                         int w, int c) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2, hp = ho + 2, wp = wo + 2, c8 = c / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * hp * wp * c8;
  if (gid >= total) return;
  const int g8 = int(gid % c8);
  const long long pix = gid / c8;
  const int px = int(pix % wp), py = int((pix / wp) % hp), b = int(pix / ((long long)wp * hp));
  uint4 best = make_uint4(0, 0, 0, 0);
  if (px >= 1 && px <= wo && py >= 1 && py <= ho) {
    const int oy = py - 1, ox = px - 1;
    bool first = true;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = 2 * oy - 1 + dy;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = 2 * ox - 1 + dx;
        if (ix < 0 || ix >= w) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t(b) * h + iy) * w + ix) * c) + g8);
        best = first ? v : bf16x8_max(best, v);
        first = false;
      }
    }
  }
  *(reinterpret_cast<uint4*>(out + pix * c) + g8) = best;
}

// Stride-2 gathers: in (B, h + 2, w + 2, c) with halo -> rows (B * (ho + 2) * (wo + 2), taps * c), ho = ceil(h / 2), the A operand of the
// stride-2 convolutions (taps = 9: 3x3 / pad 1, column = (ky * 3 + kx) * c + channel; taps = 1: the 1x1 down-sampling
// shortcut).  Ring rows of the output are written as zeros.  One thread per (row, tap, 8 channels).
// taps = 4: the FOUR-PHASE split for the implicit stride-2 3x3 convolution (hoigen_gemm_params.conv_stride = 2):
// rows [p][B * (ho + 2) * (wo + 2)][c], p = 2 py + px, holding in[2 y' + py][2 x' + px] at haloed position (y' + 1, x' + 1) and
// zero where that source pixel does not exist -- 4/9 of the bytes of the nine-tap gather, written and read.
__global__ void conv_phase_split_s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ rows, int batch, int h,
                                           int w, int c) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2, hp = ho + 2, wp = wo + 2, c8 = c / 8;
  const int hin = h + 2, win = w + 2;
  const long long per_phase = (long long)batch * hp * wp;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= 4 * per_phase * c8) return;
  const int g8 = int(gid % c8);
  const long long r = gid / c8;                   // row of the [4 * per_phase, c] matrix
  const int ph = int(r / per_phase);
  const long long pix = r - ph * per_phase;
  const int X = int(pix % wp), Y = int((pix / wp) % hp), b = int(pix / ((long long)wp * hp));
  const int iy = 2 * (Y - 1) + (ph >> 1), ix = 2 * (X - 1) + (ph & 1);       // unpadded source pixel
  uint4 v = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < h && ix >= 0 && ix < w)
    v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t(b) * hin + iy + 1) * win + ix + 1) * c) + g8);
  *(reinterpret_cast<uint4*>(rows + r * c) + g8) = v;
}

__global__ void conv_gather_s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ rows, int batch, int h,
                                      int w, int c, int taps) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2, hp = ho + 2, wp = wo + 2, c8 = c / 8;   // odd h: tap row 2 ho = h + 1 is the halo
  const int hin = h + 2, win = w + 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)batch * hp * wp * taps * c8;
  if (gid >= total) return;
  const int g8 = int(gid % c8);
  const int t = int((gid / c8) % taps);
  const long long pix = gid / ((long long)c8 * taps);
  const int px = int(pix % wp), py = int((pix / wp) % hp), b = int(pix / ((long long)wp * hp));
  uint4 v = make_uint4(0, 0, 0, 0);
  if (px >= 1 && px <= wo && py >= 1 && py <= ho) {
    // unpadded output (py-1, px-1) -> unpadded input centre (2(py-1), 2(px-1)); + 1 for the input halo; 3x3 taps at -1..1
    const int ky = taps == 9 ? t / 3 - 1 : 0, kx = taps == 9 ? t % 3 - 1 : 0;
    const int iy = 2 * (py - 1) + 1 + ky, ix = 2 * (px - 1) + 1 + kx;
    v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t(b) * hin + iy) * win + ix) * c) + g8);
  }
  *(reinterpret_cast<uint4*>(rows + (pix * taps + t) * c) + g8) = v;
}

// Global average pooling over the interior of (B, h + 2, w + 2, c) followed by the L2 normalisation of U:1618:
// out (B, c) fp32.  One CTA per image, 256 threads, each owning c / 256 channels.
__global__ void __launch_bounds__(256) avgpool_l2norm_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int h,
                                                              int w, int c) {
  __shared__ float s_part[8];
  const int b = blockIdx.x;
  const int hp = h + 2, wp = w + 2;
  const float inv = 1.0f / float(h * w);
  float ssq = 0.f;
  const int per = c / 256;   // 8 for c = 2048: one 16-byte load per pixel
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int y = 1; y <= h; ++y)
    for (int x = 1; x <= w; ++x) {
      const __nv_bfloat16* p = in + ((size_t(b) * hp + y) * wp + x) * c + threadIdx.x * per;
      if (per == 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(w4[e] << 16);
          acc[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
        }
      } else {
        for (int j = 0; j < per; ++j) acc[j] += __bfloat162float(p[j]);
      }
    }
  for (int j = 0; j < per; ++j) { acc[j] *= inv; ssq += acc[j] * acc[j]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ssq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_part[i];
  const float rn = 1.0f / sqrtf(tot);
  for (int j = 0; j < per; ++j) out[size_t(b) * c + threadIdx.x * per + j] = acc[j] * rn;
}

}  // namespace hoigen

extern "C" {

using namespace hoigen;

int hoigen_stem_im2col(const float* images, void* rows_bf16, int32_t batch, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(images && rows_bf16 && batch > 0, "stem_im2col: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("stem_im2col", s, 0, double(batch) * (3.0 * 224 * 224 * 4 + 112.0 * 112 * STEM_K * 2));
  stem_im2col_kernel<<<dim3(56, batch), 256, 0, s>>>(images, reinterpret_cast<__nv_bfloat16*>(rows_bf16));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_stem_im2col_hw(const float* images, void* rows_bf16, int32_t batch, int32_t h, int32_t w, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(images && rows_bf16 && batch > 0 && h > 0 && w > 0, "stem_im2col_hw: bad arguments");
  if (h == 224 && w == 224) return hoigen_stem_im2col(images, rows_bf16, batch, stream);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)batch * ((h + 1) / 2) * ((w + 1) / 2) * (STEM_K / 8);
  HOIGEN_CHECK_ARG(total / 256 < 0x7fffffffLL, "stem_im2col_hw: batch too large");
  KernelScope ks("stem_im2col", s, 0, double(batch) * (3.0 * h * w * 4 + double((h + 1) / 2) * ((w + 1) / 2) * STEM_K * 2));
  stem_im2col_hw_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(images, reinterpret_cast<__nv_bfloat16*>(rows_bf16), batch, h, w);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_maxpool3x3s2_halo(const void* in_bf16, void* out_bf16, int32_t batch, int32_t h, int32_t w, int32_t c,
                             hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(in_bf16 && out_bf16 && batch > 0 && h > 0 && w > 0 && c > 0 && (c % 8) == 0,
                   "maxpool3x3s2_halo: bad arguments (h=%d w=%d c=%d)", h, w, c);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = (long long)batch * ((h + 1) / 2 + 2) * ((w + 1) / 2 + 2) * (c / 8);
  KernelScope ks("maxpool3x3s2", s, 0, double(batch) * c * 2 * (double(h) * w + double((h + 1) / 2 + 2) * ((w + 1) / 2 + 2)));
  maxpool3x3s2_halo_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(in_bf16), reinterpret_cast<__nv_bfloat16*>(out_bf16), batch, h, w, c);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_conv_gather_s2(const void* in_bf16, void* rows_bf16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t taps,
                          hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(in_bf16 && rows_bf16 && batch > 0 && h > 0 && w > 0 && c > 0 && (c % 8) == 0 && (taps == 1 || taps == 9 || taps == 4),
                   "conv_gather_s2: bad arguments (h=%d w=%d c=%d taps=%d)", h, w, c, taps);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (taps == 4) {
    const long long total4 = 4LL * batch * ((h + 1) / 2 + 2) * ((w + 1) / 2 + 2) * (c / 8);
    KernelScope ks4("conv_phase_split_s2", s, 0, 2.0 * double(total4) * 16);
    conv_phase_split_s2_kernel<<<unsigned((total4 + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(in_bf16), reinterpret_cast<__nv_bfloat16*>(rows_bf16), batch, h, w, c);
    HOIGEN_CHECK_LAUNCH();
    return HOIGEN_OK;
  }
  const long long total = (long long)batch * ((h + 1) / 2 + 2) * ((w + 1) / 2 + 2) * taps * (c / 8);
  KernelScope ks("conv_gather_s2", s, 0, 2.0 * double(total) * 16);
  conv_gather_s2_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(in_bf16), reinterpret_cast<__nv_bfloat16*>(rows_bf16), batch, h, w, c, taps);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_avgpool_l2norm(const void* in_bf16, float* out, int32_t batch, int32_t h, int32_t w, int32_t c, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(in_bf16 && out && batch > 0 && h > 0 && w > 0 && c > 0 && (c % 256) == 0 && c / 256 <= 8,
                   "avgpool_l2norm: c must be a multiple of 256, at most 2048 (got %d)", c);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("avgpool_l2norm", s, 0, double(batch) * c * (double(h) * w * 2 + 4));
  avgpool_l2norm_kernel<<<batch, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(in_bf16), out, h, w, c);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_conv_plan_run(const hoigen_conv_op* ops, int32_t n_ops, hoigen_stream_t stream) {
  HOIGEN_CHECK_ARG(ops != nullptr && n_ops > 0, "conv_plan_run: empty plan");
  for (int i = 0; i < n_ops; ++i) {
    const hoigen_conv_op& o = ops[i];
    int rc = HOIGEN_ERR_INVALID;
    switch (o.kind) {
      case HOIGEN_CONV_OP_GEMM: rc = hoigen_gemm_bf16(&o.gemm, stream); break;
      case HOIGEN_CONV_OP_STEM_IM2COL:
        rc = (o.h > 0 && o.w > 0) ? hoigen_stem_im2col_hw(reinterpret_cast<const float*>(o.in), o.out, o.batch, o.h, o.w, stream)
                                  : hoigen_stem_im2col(reinterpret_cast<const float*>(o.in), o.out, o.batch, stream);
        break;
      case HOIGEN_CONV_OP_STEM_CONV:
        rc = hoigen_stem_conv_hw(reinterpret_cast<const float*>(o.in), o.gemm.w, o.gemm.bias, o.out, o.batch, o.h > 0 ? o.h : 224,
                                 o.w > 0 ? o.w : 224, stream);
        break;
      case HOIGEN_CONV_OP_MAXPOOL: rc = hoigen_maxpool3x3s2_halo(o.in, o.out, o.batch, o.h, o.w, o.c, stream); break;
      case HOIGEN_CONV_OP_GATHER_S2: rc = hoigen_conv_gather_s2(o.in, o.out, o.batch, o.h, o.w, o.c, o.taps, stream); break;
      case HOIGEN_CONV_OP_AVGPOOL_L2NORM:
        rc = hoigen_avgpool_l2norm(o.in, reinterpret_cast<float*>(o.out), o.batch, o.h, o.w, o.c, stream);
        break;
      default: set_error("conv_plan_run: op %d has unknown kind %d", i, o.kind); return HOIGEN_ERR_INVALID;
    }
    if (rc != HOIGEN_OK) return rc;
  }
  return HOIGEN_OK;
}

}  // extern "C"
