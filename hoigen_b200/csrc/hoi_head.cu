// HOI head kernels (HBM / latency bound, fp32): prior-token MLP, RoIAlign(7x7, adaptive, aligned) + mean over the
// 14x14 patch-token grid for single and union boxes, pairwise human/object/union feature assembly with L2
// normalisation, per-image logit broadcast, and prior-score + ordered triplet emission.
//
// Reference (U = upt_tip_cache_model_free_finetune_distill3.py):
//   get_prior U:1445-1495, compute_roi_embeddings U:981-1057 (torchvision.ops.roi_align call sites U:1028-1029),
//   compute_prior_scores U:806-833, postprocessing U:1408-1427.
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "ptx.cuh"
#include "roi_common.cuh"

namespace hoigen {

constexpr int FEAT = 512;
constexpr int TOK = 197;

__device__ __forceinline__ float warp_sum_h(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// prior tokens: [score, box/(w,h,w,h), object_embedding[label]] (517) -> 128 -> 128 -> 64 (ReLU between)
// one block per (image, 8 tokens), 256 threads; thread (o, half) owns output feature o of four of those tokens.
// Padding tokens (t >= n_b) are MLP(0) constants and mask = 1 (U:1448-1450, 1469, 1495).
// Weights are passed TRANSPOSED (in, out) and stream through a three-stage cp.async ring of 32-row chunks (16 KiB): the
// chunk loads of the whole three-layer chain stay two ahead of the FMAs, so the kernel runs at its weight-stream rate
// instead of one exposed L2 round trip per chunk (41 -> ~10 us at B = 64).  Every output element is still the k-ascending
// fmaf chain  acc = fmaf(w[k][o], x[t][k], acc)  from the bias: results are bit-identical to the first form.
// ------------------------------------------------------------------------------------------------
constexpr int PRIOR_IN = 517;
constexpr int PRIOR_MAXTOK = 32;
constexpr int PRIOR_TPB = 8;            // tokens per block
constexpr int PRIOR_LDX = 520;          // row pitch of the staged inputs (16-byte aligned rows)
constexpr int PRIOR_CHUNK = 32;         // weight rows per ring stage
constexpr int PRIOR_STAGES = 3;
constexpr int PRIOR_SMEM_BYTES = (PRIOR_TPB * PRIOR_LDX + 2 * PRIOR_TPB * 128 + PRIOR_STAGES * PRIOR_CHUNK * 128) * 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One weight chunk of the chain: layer L (0, 1, 2), rows [k0, k0 + kn) of its (K x NOUT) transposed matrix.
struct PriorChunk {
  const float* src;   // first row of the chunk
  int kn;             // rows (<= 32)
  int nout;           // 128 or 64
};

// acc[t] = fmaf(w[k][o], x[t][k], acc[t]) for k ascending over one chunk (TOK tokens of this thread); full chunks have a
// compile-time trip count so that the next quads' shared-memory loads are in flight under the current quad's FMAs
template <int NOUT, int TOK>
__device__ __forceinline__ void prior_chunk_fma(const float* __restrict__ ws, const float* __restrict__ x, int ldx, int kn, int o,
                                                float (&acc)[TOK]) {
  auto quad = [&](int k) {
    const float wa = ws[k * NOUT + o], wb = ws[(k + 1) * NOUT + o], wc = ws[(k + 2) * NOUT + o], wd = ws[(k + 3) * NOUT + o];
#pragma unroll
    for (int t = 0; t < TOK; ++t) {
      const float4 xv = *reinterpret_cast<const float4*>(x + t * ldx + k);    // broadcast
      acc[t] = fmaf(wa, xv.x, acc[t]);
      acc[t] = fmaf(wb, xv.y, acc[t]);
      acc[t] = fmaf(wc, xv.z, acc[t]);
      acc[t] = fmaf(wd, xv.w, acc[t]);
    }
  };
  if (kn == PRIOR_CHUNK) {
#pragma unroll
    for (int k = 0; k < PRIOR_CHUNK; k += 4) quad(k);
  } else {
    const int k4 = kn & ~3;
    for (int k = 0; k < k4; k += 4) quad(k);
    for (int k = k4; k < kn; ++k) {
      const float w = ws[k * NOUT + o];
#pragma unroll
      for (int t = 0; t < TOK; ++t) acc[t] = fmaf(w, x[t * ldx + k], acc[t]);
    }
  }
}

constexpr int PRIOR_THREADS = 256;      // thread = (output feature o, token half): 8 warps per CTA to cover the LDS latency
constexpr int PRIOR_TPT = PRIOR_TPB / 2;   // tokens per thread

__global__ void __launch_bounds__(PRIOR_THREADS)
prior_tokens_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ labels,
                    const int* __restrict__ box_off, const float* __restrict__ obj_emb, const float* __restrict__ w0t,
                    const float* __restrict__ b0, const float* __restrict__ w1t, const float* __restrict__ b1,
                    const float* __restrict__ w2t, const float* __restrict__ b2, float img_w, float img_h, int n_max,
                    int num_objects, float* __restrict__ prior, uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) float prior_smem[];
  float* xin = prior_smem;                                   // [8][520]
  float* h1 = xin + PRIOR_TPB * PRIOR_LDX;                   // [8][128]
  float* h2 = h1 + PRIOR_TPB * 128;                          // [8][128]
  float* ring = h2 + PRIOR_TPB * 128;                        // [3][32 x 128]
  constexpr int C0 = (PRIOR_IN + PRIOR_CHUNK - 1) / PRIOR_CHUNK;   // 17 chunks of layer 0
  constexpr int C1 = 128 / PRIOR_CHUNK;                            // 4 of layer 1, 4 of layer 2
  constexpr int NCHUNK = C0 + 2 * C1;
  const int b = blockIdx.x;
  const int tbase = blockIdx.y * PRIOR_TPB;
  const int o = threadIdx.x & 127, th = threadIdx.x >> 7;     // tokens [4 th, 4 th + 4) of the block
  const int base = box_off[b];
  const int n = box_off[b + 1] - base;

  auto chunk_of = [&](int c) {
    PriorChunk ch;
    if (c < C0) { ch.src = w0t + size_t(c) * PRIOR_CHUNK * 128; ch.kn = min(PRIOR_CHUNK, PRIOR_IN - c * PRIOR_CHUNK); ch.nout = 128; }
    else if (c < C0 + C1) { ch.src = w1t + size_t(c - C0) * PRIOR_CHUNK * 128; ch.kn = PRIOR_CHUNK; ch.nout = 128; }
    else { ch.src = w2t + size_t(c - C0 - C1) * PRIOR_CHUNK * 64; ch.kn = PRIOR_CHUNK; ch.nout = 64; }
    return ch;
  };
  auto issue = [&](int c) {     // chunk c -> ring stage c % 3 (16-byte pieces, coalesced); one commit group per call
    if (c < NCHUNK) {
      const PriorChunk ch = chunk_of(c);
      float* dst = ring + (c % PRIOR_STAGES) * PRIOR_CHUNK * 128;
      const int pieces = ch.kn * ch.nout / 4;
      for (int i = threadIdx.x; i < pieces; i += PRIOR_THREADS) cp_async16(dst + 4 * i, ch.src + 4 * i);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);

  // staged inputs of the block's tokens (zero rows for padding tokens).  The embedding gather depends on the label load:
  // one owner thread per token fetches the label (and writes the five scalar columns), then every thread issues its eight
  // INDEPENDENT 16-byte row loads at once -- the element-per-iteration form paid one exposed round trip per iteration.
  __shared__ int s_label[PRIOR_TPB];
  if (threadIdx.x < PRIOR_TPB) {
    const int t = tbase + threadIdx.x;
    float* row = xin + threadIdx.x * PRIOR_LDX;
    int lab = -1;
    float sc = 0.f, x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (t < n) {
      const int64_t lb = labels[base + t];   // a label outside the embedding table reads nothing (the host surface validates)
      lab = (lb >= 0 && lb < num_objects) ? int(lb) : -1;
      const float* bx = boxes + size_t(base + t) * 4;
      sc = scores[base + t];
      x1 = bx[0] / img_w; y1 = bx[1] / img_h; x2 = bx[2] / img_w; y2 = bx[3] / img_h;
    }
    s_label[threadIdx.x] = lab;
    row[0] = sc; row[1] = x1; row[2] = y1; row[3] = x2; row[4] = y2;
    row[PRIOR_IN] = 0.f; row[PRIOR_IN + 1] = 0.f; row[PRIOR_IN + 2] = 0.f;
    if (t < n_max) mask[b * n_max + t] = t < n ? 0 : 1;
  }
  __syncthreads();
  {
    float4 v[PRIOR_TPT];
#pragma unroll
    for (int t = 0; t < PRIOR_TPT; ++t) {          // token 4 th + t, floats [4 o, 4 o + 4) of its embedding row
      const int lab = s_label[th * PRIOR_TPT + t];
      v[t] = lab >= 0 ? __ldg(reinterpret_cast<const float4*>(obj_emb + size_t(lab) * FEAT) + o) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int t = 0; t < PRIOR_TPT; ++t) {
      float* dst = xin + (th * PRIOR_TPT + t) * PRIOR_LDX + 5 + 4 * o;
      dst[0] = v[t].x; dst[1] = v[t].y; dst[2] = v[t].z; dst[3] = v[t].w;
    }
  }

  float acc[PRIOR_TPT];
  {
    const float bias = b0[o];
#pragma unroll
    for (int t = 0; t < PRIOR_TPT; ++t) acc[t] = bias;
  }
  const int t0 = th * PRIOR_TPT;
  for (int c = 0; c < NCHUNK; ++c) {
    issue(c + 2);
    cp_async_wait<2>();        // this thread's pieces of chunk c have landed ...
    __syncthreads();           // ... and everyone's (also orders the staged inputs / previous layer's activations)
    const PriorChunk ch = chunk_of(c);
    const float* ws = ring + (c % PRIOR_STAGES) * PRIOR_CHUNK * 128;
    if (c < C0) prior_chunk_fma<128, PRIOR_TPT>(ws, xin + t0 * PRIOR_LDX + c * PRIOR_CHUNK, PRIOR_LDX, ch.kn, o, acc);
    else if (c < C0 + C1) prior_chunk_fma<128, PRIOR_TPT>(ws, h1 + t0 * 128 + (c - C0) * PRIOR_CHUNK, 128, ch.kn, o, acc);
    else if (o < 64) prior_chunk_fma<64, PRIOR_TPT>(ws, h2 + t0 * 128 + (c - C0 - C1) * PRIOR_CHUNK, 128, ch.kn, o, acc);
    // layer boundaries: activations out, next layer's bias in
    if (c == C0 - 1) {
#pragma unroll
      for (int t = 0; t < PRIOR_TPT; ++t) h1[(t0 + t) * 128 + o] = fmaxf(acc[t], 0.f);
      const float bias = b1[o];
#pragma unroll
      for (int t = 0; t < PRIOR_TPT; ++t) acc[t] = bias;
    } else if (c == C0 + C1 - 1) {
#pragma unroll
      for (int t = 0; t < PRIOR_TPT; ++t) h2[(t0 + t) * 128 + o] = fmaxf(acc[t], 0.f);
      const float bias = o < 64 ? b2[o] : 0.f;
#pragma unroll
      for (int t = 0; t < PRIOR_TPT; ++t) acc[t] = bias;
    }
    __syncthreads();           // the stage is refilled by the next iteration's issue
  }
  if (o < 64) {
#pragma unroll
    for (int t = 0; t < PRIOR_TPT; ++t)
      if (tbase + t0 + t < n_max) prior[(size_t(b) * n_max + tbase + t0 + t) * 64 + o] = acc[t];
  }
}

// ------------------------------------------------------------------------------------------------
// RoIAlign + mean.  Because the 7x7xg^2 sampling lattice is a product lattice and bilinear weights factorise,
//   mean_bins(RoIAlign(F, box))[c] = (1 / (49 * count)) * sum_y sum_x Wy[y] * Wx[x] * F[y][x][c]
// with per-axis weights Wy / Wx accumulated over the 7*g samples of that axis (out-of-range samples,
// coordinate < -1 or > 14, contribute zero exactly as in torchvision's kernel).
// grid = (B, 4): a CTA stages a 128-channel slice of one image's 14x14x512 token map in shared memory (coalesced
// float4 loads) and produces that slice for every single box and every union box of the image; one warp per group of
// four boxes, lane = 4 channels.
// ------------------------------------------------------------------------------------------------
constexpr int ROI_THREADS = 256;
constexpr int ROI_SLICE = 128;
constexpr int ROI_SMEM_BYTES = 196 * ROI_SLICE * 4;

// Per-axis RoIAlign weights of every single / union box, computed ONCE (not once per channel slice): one warp per box,
// lanes 0..13 -> Wy, lanes 16..29 -> Wx, lane 31 -> 1 / (49 * count).  job index: [0, Ntot) singles, [Ntot, Ntot+Ktot)
// unions, in global (CSR) numbering.
__global__ void __launch_bounds__(256)
roi_weights_kernel(const float* __restrict__ boxes, const int* __restrict__ box_off, const int* __restrict__ pair_off,
                   int nimg, int ntot, int ktot, float spatial_scale, float* __restrict__ wts /* (Ntot+Ktot, 32) */);

// A warp takes FOUR consecutive boxes of the image at a time (unions are enumerated pair-major, so the four usually
// share their human and their windows nest) and walks the bounding window of the four ONCE: per cell one LDS.128 of the
// token slice feeds 16 FMAs, and the four boxes' axis weights come as one broadcast LDS.128 from a per-warp scratch row.
// (One box per warp with a SHFL per weight was bound by the shared-memory / shuffle pipe, 4 + 1 wavefront cycles per 4
// FMAs: ncu smem pipe 52 %, 80 us at B = 64; this form measures 69 us.)  Cells outside a box's own window carry weight
// 0 for it: fmaf(0, f, acc) == acc, so every box's sum is the sequence of its own non-zero terms.
constexpr int ROI_GROUP = 4;
constexpr int ROI_SMEM_BYTES_V2 = ROI_SMEM_BYTES + (ROI_THREADS / 32) * 32 * 16;   // + one float4[32] weight row per warp

__global__ void __launch_bounds__(ROI_THREADS)
roi_features_grouped_kernel(const float* __restrict__ tokens /* (B*197, 512) */, const float* __restrict__ wts,
                            const int* __restrict__ box_off, const int* __restrict__ pair_off, int ntot,
                            float* __restrict__ single_feat /* (Ntot,512) */, float* __restrict__ union_feat /* (Ktot,512) */) {
  extern __shared__ float4 fs4[];  // [196][32] float4 token slice, then [8 warps][32] float4 weight rows
  const int b = blockIdx.x, slice = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* wq = fs4 + 196 * 32 + warp * 32;
  const float4* src = reinterpret_cast<const float4*>(tokens + (size_t(b) * TOK + 1) * FEAT + slice * ROI_SLICE);
  for (int i = threadIdx.x; i < 196 * 32; i += ROI_THREADS) fs4[i] = __ldg(src + (i >> 5) * (FEAT / 4) + (i & 31));
  __syncthreads();
  const int bbase = box_off[b];
  const int n = box_off[b + 1] - bbase;
  const int pbase = pair_off[b];
  const int K = pair_off[b + 1] - pbase;
  const int njobs = n + K;
  for (int j0 = warp * ROI_GROUP; j0 < njobs; j0 += (ROI_THREADS / 32) * ROI_GROUP) {
    float w[ROI_GROUP];
    float* dst[ROI_GROUP];
#pragma unroll
    for (int r = 0; r < ROI_GROUP; ++r) {
      const int job = j0 + r;
      w[r] = 0.f;
      dst[r] = nullptr;
      if (job < njobs) {
        const int gjob = job < n ? bbase + job : ntot + pbase + (job - n);
        dst[r] = job < n ? single_feat + size_t(bbase + job) * FEAT : union_feat + size_t(pbase + job - n) * FEAT;
        w[r] = __ldg(wts + size_t(gjob) * 32 + lane);
      }
    }
    __syncwarp();                       // the previous group's reads of this warp's weight row are done
    wq[lane] = make_float4(w[0], w[1], w[2], w[3]);
    __syncwarp();
    const bool any = (w[0] != 0.f) | (w[1] != 0.f) | (w[2] != 0.f) | (w[3] != 0.f);
    const unsigned nzy = __ballot_sync(0xffffffffu, lane < G14 && any);
    const unsigned nzx = __ballot_sync(0xffffffffu, lane >= 16 && lane < 16 + G14 && any) >> 16;
    float4 acc[ROI_GROUP];
#pragma unroll
    for (int r = 0; r < ROI_GROUP; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nzy && nzx) {
      const int ylo = __ffs(nzy) - 1, yhi = 31 - __clz(nzy);
      const int xlo = __ffs(nzx) - 1, xhi = 31 - __clz(nzx);
      for (int y = ylo; y <= yhi; ++y) {
        float4 row[ROI_GROUP];
#pragma unroll
        for (int r = 0; r < ROI_GROUP; ++r) row[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* frow = fs4 + (y * G14) * 32 + lane;
        for (int x = xlo; x <= xhi; ++x) {
          const float4 wx = wq[16 + x];             // broadcast: the four boxes' x-weights of this column
          const float4 f = frow[x * 32];
          const float wxr[ROI_GROUP] = {wx.x, wx.y, wx.z, wx.w};
#pragma unroll
          for (int r = 0; r < ROI_GROUP; ++r) {
            row[r].x = fmaf(wxr[r], f.x, row[r].x); row[r].y = fmaf(wxr[r], f.y, row[r].y);
            row[r].z = fmaf(wxr[r], f.z, row[r].z); row[r].w = fmaf(wxr[r], f.w, row[r].w);
          }
        }
        const float4 wy = wq[y];
        const float wyr[ROI_GROUP] = {wy.x, wy.y, wy.z, wy.w};
#pragma unroll
        for (int r = 0; r < ROI_GROUP; ++r) {
          acc[r].x = fmaf(wyr[r], row[r].x, acc[r].x); acc[r].y = fmaf(wyr[r], row[r].y, acc[r].y);
          acc[r].z = fmaf(wyr[r], row[r].z, acc[r].z); acc[r].w = fmaf(wyr[r], row[r].w, acc[r].w);
        }
      }
    }
    const float4 inv4 = wq[31];
    const float inv[ROI_GROUP] = {inv4.x, inv4.y, inv4.z, inv4.w};
#pragma unroll
    for (int r = 0; r < ROI_GROUP; ++r) {
      if (dst[r]) {
        acc[r].x *= inv[r]; acc[r].y *= inv[r]; acc[r].z *= inv[r]; acc[r].w *= inv[r];
        reinterpret_cast<float4*>(dst[r] + slice * ROI_SLICE)[lane] = acc[r];
      }
    }
  }
}

__device__ __forceinline__ int find_image(const int* __restrict__ off, int nimg, int idx);

__global__ void __launch_bounds__(256)
roi_weights_kernel(const float* __restrict__ boxes, const int* __restrict__ box_off, const int* __restrict__ pair_off,
                   int nimg, int ntot, int ktot, float spatial_scale, float* __restrict__ wts) {
  const int gjob = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gjob >= ntot + ktot) return;
  float x1, y1, x2, y2;
  if (gjob < ntot) {
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + gjob);
    x1 = bx.x; y1 = bx.y; x2 = bx.z; y2 = bx.w;
  } else {
    const int gi = gjob - ntot;
    const int b = find_image(pair_off, nimg, gi);
    const int bbase = box_off[b], n = box_off[b + 1] - bbase;
    const int i = gi - pair_off[b];
    const int px = i / (n - 1), r = i % (n - 1);
    const int py = r < px ? r : r + 1;           // row-major enumeration of (x, y != x), x < n_h  (U:1007-1012)
    const float4 bh = __ldg(reinterpret_cast<const float4*>(boxes) + bbase + px);
    const float4 bo = __ldg(reinterpret_cast<const float4*>(boxes) + bbase + py);
    x1 = fminf(bh.x, bo.x); y1 = fminf(bh.y, bo.y); x2 = fmaxf(bh.z, bo.z); y2 = fmaxf(bh.w, bo.w);  // U:1021-1023
  }
  const float sx = x1 * spatial_scale - 0.5f, sy = y1 * spatial_scale - 0.5f;
  const float ex = x2 * spatial_scale - 0.5f, ey = y2 * spatial_scale - 0.5f;
  const float rw = ex - sx, rh = ey - sy;
  const float bw = rw / float(POOL), bh_ = rh / float(POOL);
  const int gw = int(ceilf(rw / float(POOL))), gh = int(ceilf(rh / float(POOL)));
  const float count = float(max(gh * gw, 1));
  float wl = 0.f;
  if (lane < G14) wl = axis_weight(lane, sy, bh_, gh);
  else if (lane >= 16 && lane < 16 + G14) wl = axis_weight(lane - 16, sx, bw, gw);
  else if (lane == 31) wl = 1.0f / (count * float(POOL * POOL));
  wts[size_t(gjob) * 32 + lane] = wl;
}

// ------------------------------------------------------------------------------------------------
// pair assembly: f_H[i] = single[x_i]/|.|, f_O[i] = single[y_i]/|.|, f_U[i] = union[i]/|.|   (U:1044-1050)
// one warp per pair; outputs bf16 (GEMM operands) and optionally fp32. out layout: [3][Ktot][512] (H, O, U).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_image(const int* __restrict__ off, int nimg, int idx) {
  int lo = 0, hi = nimg;  // off[lo] <= idx < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
pair_assemble_kernel(const float* __restrict__ single_feat, const float* __restrict__ union_feat,
                     const int* __restrict__ box_off, const int* __restrict__ pair_off, int nimg, int ktot,
                     __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ktot) return;
  const int b = find_image(pair_off, nimg, warp);
  const int n = box_off[b + 1] - box_off[b];
  const int i = warp - pair_off[b];
  const int px = i / (n - 1), r = i % (n - 1);
  const int py = r < px ? r : r + 1;
  const float* srcs[3] = {single_feat + size_t(box_off[b] + px) * FEAT, single_feat + size_t(box_off[b] + py) * FEAT,
                          union_feat + size_t(warp) * FEAT};
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    float4 v[4];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = reinterpret_cast<const float4*>(srcs[s])[lane + 32 * j];
      ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float nrm = sqrtf(warp_sum_h(ss));  // 0 -> x/0 = NaN, as in the reference
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 y = make_float4(v[j].x / nrm, v[j].y / nrm, v[j].z / nrm, v[j].w / nrm);
      const size_t o = (size_t(s) * ktot + warp) * FEAT;
      if (out_f32) reinterpret_cast<float4*>(out_f32 + o)[lane + 32 * j] = y;
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_bf16 + o)[lane + 32 * j] = pk;
    }
  }
}

// A few 32-bit words between device memory and MAPPED pinned host memory (either direction) by a kernel on the
// compute stream.  The path's tiny transfers (CSR layout up, triplet offsets down) must not go through the copy
// engines: there they queue behind a caller's bulk image upload / detection download and stall the compute stream.
__global__ void __launch_bounds__(256)
copy_words_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// Host values -> device memory THROUGH THE KERNEL PARAMETERS (no DMA, no PCIe read by the kernel): while a caller's bulk
// image upload saturates the link, a memcpy queues behind it on the copy engine and a kernel reading mapped host memory
// queues behind it on the link — both were measured to stall the compute stream 0.3-0.9 ms per step.
constexpr int SET_WORDS_MAX = 960;   // 3840 of the 4096 parameter bytes
struct WordPack { uint32_t v[SET_WORDS_MAX]; };
__global__ void __launch_bounds__(256)
set_words_kernel(uint32_t* __restrict__ dst, const __grid_constant__ WordPack pack, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = pack.v[i];
}

// rows of `in` (ld_in floats apart) -> L2-normalised (optional) bf16 rows (cols wide); one warp per row.
__global__ void __launch_bounds__(256)
rows_to_bf16_kernel(const float* __restrict__ in, long ld_in, int rows, int cols, int normalize,
                    __nv_bfloat16* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* src = in + size_t(warp) * ld_in;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) ss += src[c] * src[c];
  const float nrm = normalize ? sqrtf(warp_sum_h(ss)) : 1.0f;
  for (int c = lane; c < cols; c += 32) out[size_t(warp) * cols + c] = __float2bfloat16_rn(src[c] / nrm);
}

// fp32 rows -> THREE bf16 planes hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid) (8 + 8 + 8 mantissa bits: the sum
// restores the fp32 value), written side by side along K so that ONE bf16 tensor-core GEMM with fp32 accumulation evaluates
// the fp32 product (the "3 x bf16 split"; every partial product of two bf16 numbers is exact in fp32):
//   pattern 6: out row = [hi | hi | hi | mid | mid | lo]   against a packed weight row [hi | mid | lo | hi | mid | hi]
//              (all cross terms whose weight is >= 2^-16 of the leading one)
//   pattern 3: out row = [hi | mid | lo]                    against an operand that is EXACT in bf16, tiled three times
// optional L2 normalisation of the row first (U:960, U:1618).  One warp per row.
__global__ void __launch_bounds__(256)
rows_split3_kernel(const float* __restrict__ in, long ld_in, int rows, int cols, int normalize, int pattern,
                   __nv_bfloat16* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* src = in + size_t(warp) * ld_in;
  float ss = 0.f;
  if (normalize)
    for (int c = lane; c < cols; c += 32) ss += src[c] * src[c];
  const float nrm = normalize ? sqrtf(warp_sum_h(ss)) : 1.0f;
  __nv_bfloat16* dst = out + size_t(warp) * cols * pattern;
  for (int c = lane; c < cols; c += 32) {
    const float x = normalize ? src[c] / nrm : src[c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    if (pattern == 6) {
      dst[c] = hi; dst[cols + c] = hi; dst[2 * cols + c] = hi;
      dst[3 * cols + c] = mid; dst[4 * cols + c] = mid; dst[5 * cols + c] = lo;
    } else {
      dst[c] = hi; dst[cols + c] = mid; dst[2 * cols + c] = lo;
    }
  }
}

// logits[i][:] = img_logits[image(i)][:]   (the per-image global-CLIP + DINO cache terms, U:1115, U:1138)
__global__ void broadcast_rows_kernel(const float* __restrict__ img_logits, const int* __restrict__ pair_off, int nimg,
                                      int ktot, int C, int ld, float* __restrict__ logits) {
  const long total = long(ktot) * C;
  for (long e = blockIdx.x * long(blockDim.x) + threadIdx.x; e < total; e += long(gridDim.x) * blockDim.x) {
    const int i = int(e / C), c = int(e % C);
    const int b = find_image(pair_off, nimg, i);
    logits[size_t(i) * ld + c] = img_logits[size_t(b) * C + c];
  }
}

// ------------------------------------------------------------------------------------------------
// prior scores + ordered triplet emission  (U:806-833, U:1408-1427)
//   pr_i = score[x_i]^lambda * score[y_i]^lambda for every class c in table[label[y_i]] (bitmask rows), else 0
//   emit (i, c) in row-major order wherever pr != 0: score = sigmoid(logit[i,c]) * pr, label = c,
//   pairing = (x_i, y_i), object = label[y_i].  Compiled WITHOUT fast-math: denormal products stay non-zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pair_count_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels, const int* __restrict__ box_off,
                  const int* __restrict__ pair_off, int nimg, int ktot, const uint32_t* __restrict__ table_bits,
                  int words, int table_rows, float lambda, int* __restrict__ counts, float* __restrict__ pr_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ktot) return;
  const int b = find_image(pair_off, nimg, i);
  const int base = box_off[b];
  const int n = box_off[b + 1] - base;
  const int li = i - pair_off[b];
  const int px = li / (n - 1), r = li % (n - 1);
  const int py = r < px ? r : r + 1;
  const float sh = powf(scores[base + px], lambda);
  const float so = powf(scores[base + py], lambda);
  float pr = sh * so;
  const int64_t obj = labels[base + py];
  int cnt = 0;
  if (obj < 0 || obj >= table_rows) pr = 0.f;   // no row in the object -> target-class table: emits nothing, reads nothing
  if (pr != 0.f) {
    const uint32_t* bits = table_bits + size_t(obj) * words;
    for (int w = 0; w < words; ++w) cnt += __popc(bits[w]);
  }
  counts[i] = cnt;
  pr_out[i] = pr;
}

// single-block exclusive scan of counts[0..n) -> offsets[0..n], plus per-image offsets.
__global__ void __launch_bounds__(1024)
scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ offsets, const int* __restrict__ pair_off,
            int nimg, int* __restrict__ img_off) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < n ? counts[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const int prefix = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + x - v;
    if (idx < n) offsets[idx] = prefix;
    __syncthreads();
    if (threadIdx.x == 1023) carry = prefix + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry;
  __syncthreads();
  for (int b = threadIdx.x; b <= nimg; b += 1024) img_off[b] = (pair_off[b] < n) ? offsets[pair_off[b]] : carry;
}

__global__ void __launch_bounds__(256)
emit_kernel(const float* __restrict__ logits, int C, int ld, const float* __restrict__ pr_in, const int* __restrict__ offsets,
            const int64_t* __restrict__ labels, const int* __restrict__ box_off, const int* __restrict__ pair_off,
            const int* __restrict__ img_off, int nimg, int ktot, const uint32_t* __restrict__ table_bits, int words,
            long capacity, float* __restrict__ out_scores, int64_t* __restrict__ out_labels,
            int64_t* __restrict__ out_objects, int64_t* __restrict__ out_pairing) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= ktot) return;
  const float pr = pr_in[i];
  if (pr == 0.f) return;
  const int b = find_image(pair_off, nimg, i);
  const int base = box_off[b];
  const int n = box_off[b + 1] - base;
  const int li = i - pair_off[b];
  const int px = li / (n - 1), r = li % (n - 1);
  const int py = r < px ? r : r + 1;
  const int64_t obj = labels[base + py];
  const uint32_t* bits = table_bits + size_t(obj) * words;
  int pos = offsets[i];
  // per-image pairing block: [2][M_b] at 2*img_off[b]
  const long ioff = img_off[b];
  const long mb = long(img_off[b + 1]) - ioff;
  for (int w = 0; w < words; ++w) {
    const uint32_t m = bits[w];
    const int c = w * 32 + lane;
    if ((m >> lane) & 1u) {
      const long p = long(pos) + __popc(m & ((1u << lane) - 1u));
      if (p < capacity) {
        const float x = logits[size_t(i) * ld + c];
        out_scores[p] = (1.0f / (1.0f + expf(-x))) * pr;
        out_labels[p] = c;
        out_objects[p] = obj;
        const long j = p - ioff;
        out_pairing[2 * ioff + j] = px;
        out_pairing[2 * ioff + mb + j] = py;
      }
    }
    pos += __popc(m);
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_prior_tokens(const float* boxes, const float* scores, const int64_t* labels, const int32_t* box_off,
                        const float* obj_emb, const float* w0t, const float* b0, const float* w1t, const float* b1,
                        const float* w2t, const float* b2, float img_w, float img_h, int32_t batch, int32_t n_max,
                        int32_t num_objects, float* prior, uint8_t* mask, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(boxes && scores && labels && box_off && obj_emb && prior && mask && batch > 0, "prior_tokens: bad arguments");
  HOIGEN_CHECK_ARG(n_max > 0 && n_max <= PRIOR_MAXTOK, "prior_tokens: n_max must be in [1,%d] (got %d)", PRIOR_MAXTOK, n_max);
  HOIGEN_CHECK_ARG(num_objects > 0, "prior_tokens: num_objects (rows of object_embedding) must be positive");
  KernelScope ks("prior_tokens", reinterpret_cast<cudaStream_t>(stream), 2.0 * batch * n_max * (517 * 128 + 128 * 128 + 128 * 64),
                 double(batch) * n_max * (517 + 64) * 4);
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(prior_tokens_kernel), PRIOR_SMEM_BYTES));
  HOIGEN_CHECK_ARG(((reinterpret_cast<uintptr_t>(w0t) | reinterpret_cast<uintptr_t>(w1t) | reinterpret_cast<uintptr_t>(w2t)) & 15) == 0,
                   "prior_tokens: the transposed weight matrices must be 16-byte aligned");
  prior_tokens_kernel<<<dim3(batch, (n_max + PRIOR_TPB - 1) / PRIOR_TPB), PRIOR_THREADS, PRIOR_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(
      boxes, scores, labels, box_off, obj_emb, w0t, b0, w1t, b1, w2t, b2, img_w, img_h, n_max, num_objects, prior, mask);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_roi_pair_features(const float* tokens, const float* boxes, const int32_t* box_off, const int32_t* pair_off,
                             int32_t batch, int32_t ntot, int32_t ktot, float spatial_scale, float* roi_weights,
                             float* single_feat, float* union_feat, void* pair_feat_bf16, float* pair_feat_f32,
                             hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(tokens && boxes && box_off && pair_off && roi_weights && single_feat && union_feat && pair_feat_bf16,
                   "roi_pair_features: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ntot > 0 && ktot >= 0, "roi_pair_features: bad sizes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  HOIGEN_TRY_RC(set_max_dynamic_smem(reinterpret_cast<const void*>(roi_features_grouped_kernel), ROI_SMEM_BYTES_V2));
  static const bool simt = getenv("HOIGEN_ROI_SIMT") != nullptr;      // A/B switch: the fp32 SIMT form (four boxes per warp)
  if (simt) {
    {
      KernelScope ks("roi_weights", s, 0, double(ntot + ktot) * (16 + 128));
      roi_weights_kernel<<<((ntot + ktot) * 32 + 255) / 256, 256, 0, s>>>(boxes, box_off, pair_off, batch, ntot, ktot,
                                                                          spatial_scale, roi_weights);
    }
    HOIGEN_CHECK_LAUNCH();
    KernelScope ks("roi_features", s, 2.0 * double(ntot + ktot) * 196 * FEAT, double(batch) * 196 * FEAT * 4 + double(ntot + ktot) * (16 + FEAT * 4));
    roi_features_grouped_kernel<<<dim3(batch, FEAT / ROI_SLICE), ROI_THREADS, ROI_SMEM_BYTES_V2, s>>>(
          tokens, roi_weights, box_off, pair_off, ntot, single_feat, union_feat);
  } else {
    // algorithmic bytes: token map read once + boxes + one fp32 feature row per single / union box
    KernelScope ks("roi_features", s, 2.0 * double(ntot + ktot) * 196 * FEAT, double(batch) * 196 * FEAT * 4 + double(ntot + ktot) * (16 + FEAT * 4));
    HOIGEN_TRY_RC(launch_roi_features_tc(tokens, boxes, box_off, pair_off, batch, spatial_scale, single_feat, union_feat, s));
  }
  HOIGEN_CHECK_LAUNCH();
  if (ktot > 0) {
    KernelScope ks("pair_assemble", s, 0, double(ktot) * FEAT * (3 * 4 + 3 * 2 + (pair_feat_f32 ? 12 : 0)));
    pair_assemble_kernel<<<(ktot * 32 + 255) / 256, 256, 0, s>>>(single_feat, union_feat, box_off, pair_off, batch, ktot,
                                                                 reinterpret_cast<__nv_bfloat16*>(pair_feat_bf16), pair_feat_f32);
    HOIGEN_CHECK_LAUNCH();
  }
  return HOIGEN_OK;
}

int hoigen_copy_words(void* dst, const void* src, int32_t n_words, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(dst && src && n_words > 0, "copy_words: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0,
                   "copy_words: pointers must be 4-byte aligned");
  KernelScope ks("copy_words", reinterpret_cast<cudaStream_t>(stream), 0, double(n_words) * 8);
  copy_words_kernel<<<(n_words + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(src), n_words);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_set_words(void* dst, const void* host_values, int32_t n_words, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(dst && host_values && n_words > 0, "set_words: bad arguments");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 3) == 0, "set_words: dst must be 4-byte aligned");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(host_values);
  for (int done = 0; done < n_words; done += SET_WORDS_MAX) {
    const int n = n_words - done < SET_WORDS_MAX ? n_words - done : SET_WORDS_MAX;
    WordPack pack;
    memcpy(pack.v, src + done, size_t(n) * 4);     // read on the host NOW: the caller may reuse host_values on return
    KernelScope ks("set_words", reinterpret_cast<cudaStream_t>(stream), 0, double(n) * 4);
    set_words_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<uint32_t*>(dst) + done, pack, n);
    HOIGEN_CHECK_LAUNCH();
  }
  return HOIGEN_OK;
}

int hoigen_rows_to_bf16(const float* in, int64_t ld_in, int32_t rows, int32_t cols, int32_t normalize, void* out,
                        hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(in && out && rows > 0 && cols > 0, "rows_to_bf16: bad arguments");
  KernelScope ks("rows_to_bf16", reinterpret_cast<cudaStream_t>(stream), 0, double(rows) * cols * 6);
  rows_to_bf16_kernel<<<(rows * 32 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ld_in, rows, cols, normalize, reinterpret_cast<__nv_bfloat16*>(out));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_rows_split3(const float* in, int64_t ld_in, int32_t rows, int32_t cols, int32_t normalize, int32_t pattern,
                       void* out, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(in && out && rows > 0 && cols > 0 && (pattern == 3 || pattern == 6), "rows_split3: bad arguments");
  KernelScope ks("rows_split3", reinterpret_cast<cudaStream_t>(stream), 0, double(rows) * cols * (4 + 2 * pattern));
  rows_split3_kernel<<<(rows * 32 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in, ld_in, rows, cols, normalize, pattern, reinterpret_cast<__nv_bfloat16*>(out));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_broadcast_image_logits(const float* img_logits, const int32_t* pair_off, int32_t batch, int32_t ktot,
                                  int32_t num_classes, int32_t ld_logits, float* logits, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(img_logits && pair_off && logits && batch > 0 && num_classes > 0, "broadcast_image_logits: bad arguments");
  HOIGEN_CHECK_ARG(ld_logits >= num_classes, "broadcast_image_logits: ld_logits < num_classes");
  if (ktot == 0) return HOIGEN_OK;
  const long total = long(ktot) * num_classes;
  const int blocks = int(min(long(num_sms()) * 8, (total + 255) / 256));
  KernelScope ks("broadcast_image_logits", reinterpret_cast<cudaStream_t>(stream), 0, double(total) * 4 + double(batch) * num_classes * 4);
  broadcast_rows_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img_logits, pair_off, batch, ktot,
                                                                                      num_classes, ld_logits, logits);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_emit_triplets(const float* logits, int32_t num_classes, int32_t ld_logits, const float* scores, const int64_t* labels,
                         const int32_t* box_off, const int32_t* pair_off, int32_t batch, int32_t ktot,
                         const uint32_t* table_bits, int32_t table_words, int32_t table_rows, float hyper_lambda, int32_t* work_counts,
                         int32_t* work_offsets, float* work_pr, int64_t capacity, float* out_scores,
                         int64_t* out_labels, int64_t* out_objects, int64_t* out_pairing, int32_t* img_off,
                         hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(logits && scores && labels && box_off && pair_off && table_bits && work_counts && work_offsets &&
                       work_pr && out_scores && out_labels && out_objects && out_pairing && img_off,
                   "emit_triplets: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && ktot >= 0 && num_classes > 0 && table_words * 32 >= num_classes && table_rows > 0,
                   "emit_triplets: bad sizes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (ktot > 0) {
    KernelScope ks("pair_count", s, 0, double(ktot) * 24);
    pair_count_kernel<<<(ktot + 255) / 256, 256, 0, s>>>(scores, labels, box_off, pair_off, batch, ktot, table_bits,
                                                         table_words, table_rows, hyper_lambda, work_counts, work_pr);
    HOIGEN_CHECK_LAUNCH();
  }
  {
    KernelScope ks("emit_scan", s, 0, double(ktot) * 8);
    scan_kernel<<<1, 1024, 0, s>>>(work_counts, ktot, work_offsets, pair_off, batch, img_off);
  }
  HOIGEN_CHECK_LAUNCH();
  if (ktot > 0) {
    // reads the logits row of every pair; writes 36 bytes per emitted triplet (count unknown on the host: the
    // per-class average of the object->target table bounds it; bench.py recomputes the exact figure from img_off)
    KernelScope ks("emit_triplets", s, 0, double(ktot) * num_classes * 4);
    emit_kernel<<<(ktot * 32 + 255) / 256, 256, 0, s>>>(logits, num_classes, ld_logits, work_pr, work_offsets, labels, box_off, pair_off,
                                                        img_off, batch, ktot, table_bits, table_words, long(capacity),
                                                        out_scores, out_labels, out_objects, out_pairing);
    HOIGEN_CHECK_LAUNCH();
  }
  return HOIGEN_OK;
}

}  // extern "C"
