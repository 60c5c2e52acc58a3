// Proposal stage for a whole batch in one launch (the part upstream of the hot path that is NOT
// the DETR network): UPT.prepare_region_proposals, upt_tip_cache_model_free_finetune_distill3.py:1361-1406 =
// torchvision batched_nms(boxes, scores, labels, 0.5) + score threshold + min / max-instance rule, humans first.
// The reference runs ~25 tiny torch launches and several host syncs PER IMAGE here.
//
// One CTA per image, Q <= 256 candidates (DETR: 100 queries), everything in shared memory:
//   1. batched_nms' coordinate trick: shifted = box + float(label) * (max coordinate + 1)      (fp32, as torchvision)
//   2. stable order by descending score (rank by counting)
//   3. greedy NMS in that order: ovr = inter / (area_i + area_j - inter) > thr suppresses j    (torchvision's cpu/cuda
//      kernels' arithmetic, round-to-nearest intrinsics so no FMA contraction changes a decision)
//   4. survivors are in descending-score order; humans (label == human_idx) and objects separately: if fewer than
//      min_instances pass the score threshold take the first min_instances, if more than max_instances take the first
//      max_instances, else those that pass (U:1374-1395 — every branch is a prefix of the score order)
//   5. outputs, humans first (slots by ballot / popc prefix counts, all positions in parallel): boxes / scores / labels into a (B, 2*max_instances) padded block + counts (B,2).
#include "common.h"

namespace hoigen {

constexpr int PROP_MAXQ = 256;

__global__ void __launch_bounds__(PROP_MAXQ)
prepare_proposals_kernel(const float* __restrict__ scores, const long long* __restrict__ labels, const float4* __restrict__ boxes,
                         int Q, long long human_idx, float score_thresh, int min_inst, int max_inst, float nms_thr,
                         float4* __restrict__ out_boxes, float* __restrict__ out_scores, long long* __restrict__ out_labels,
                         int* __restrict__ counts) {
  __shared__ float sx1[PROP_MAXQ], sy1[PROP_MAXQ], sx2[PROP_MAXQ], sy2[PROP_MAXQ], sarea[PROP_MAXQ], sscore[PROP_MAXQ];
  __shared__ int sorder[PROP_MAXQ];
  __shared__ unsigned char ssupp[PROP_MAXQ], shuman[PROP_MAXQ];
  __shared__ float sred[PROP_MAXQ / 32];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const bool live = t < Q;
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  long long lab = 0;
  float sc = 0.f;
  if (live) {
    bx = boxes[size_t(b) * Q + t];
    lab = labels[size_t(b) * Q + t];
    sc = scores[size_t(b) * Q + t];
  }
  // 1. max coordinate over the image's boxes
  float mx = live ? fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)) : -INFINITY;
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sred[warp] = mx;
  __syncthreads();
  mx = sred[0];
  for (int w = 1; w < PROP_MAXQ / 32; ++w) mx = fmaxf(mx, sred[w]);
  const float off = __fmul_rn(float(lab), __fadd_rn(mx, 1.0f));
  if (live) {
    const float x1 = __fadd_rn(bx.x, off), y1 = __fadd_rn(bx.y, off), x2 = __fadd_rn(bx.z, off), y2 = __fadd_rn(bx.w, off);
    sx1[t] = x1; sy1[t] = y1; sx2[t] = x2; sy2[t] = y2;
    sarea[t] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    sscore[t] = sc;
    ssupp[t] = 0;
    shuman[t] = lab == human_idx;
  }
  __syncthreads();
  // 2. stable descending order: rank = #{j : s_j > s_i  or  (s_j == s_i and j < i)}
  //    (a NaN score orders first, as torch's descending sort puts it, so the ranks stay a permutation)
  if (live) {
    const float key = sc != sc ? INFINITY : sc;
    int rank = 0;
    for (int j = 0; j < Q; ++j) {
      const float sj = sscore[j];
      const float kj = sj != sj ? INFINITY : sj;
      rank += (kj > key) || (kj == key && j < t);
    }
    sorder[rank] = t;
  }
  __syncthreads();
  // 3. greedy NMS; thread t owns position t of the order
  const int me = live ? sorder[t] : 0;
  for (int pi = 0; pi < Q; ++pi) {
    const int i = sorder[pi];
    if (!ssupp[i] && live && t > pi && !ssupp[me]) {
      const float w = fmaxf(0.f, __fsub_rn(fminf(sx2[i], sx2[me]), fmaxf(sx1[i], sx1[me])));
      const float h = fmaxf(0.f, __fsub_rn(fminf(sy2[i], sy2[me]), fmaxf(sy1[i], sy1[me])));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sarea[i], sarea[me]), inter));
      if (ovr > nms_thr) ssupp[me] = 1;
    }
    __syncthreads();
  }
  // 4. / 5. thread t owns position t of the order: its slot is its rank among the surviving humans (objects), found with
  //          warp ballots + per-warp totals; the selection is a prefix of each group (header), so no second pass
  __shared__ int swh[PROP_MAXQ / 32], swo[PROP_MAXQ / 32], swokh[PROP_MAXQ / 32], swoko[PROP_MAXQ / 32];
  const bool alive = live && !ssupp[me];
  const bool hum = alive && shuman[me];
  const bool obj = alive && !shuman[me];
  const bool pass = alive && sscore[me] >= score_thresh;
  const unsigned bh = __ballot_sync(0xffffffffu, hum), bo = __ballot_sync(0xffffffffu, obj);
  const unsigned bph = __ballot_sync(0xffffffffu, hum && pass), bpo = __ballot_sync(0xffffffffu, obj && pass);
  if (lane == 0) { swh[warp] = __popc(bh); swo[warp] = __popc(bo); swokh[warp] = __popc(bph); swoko[warp] = __popc(bpo); }
  __syncthreads();
  int all_h = 0, all_o = 0, ok_h = 0, ok_o = 0, before_h = 0, before_o = 0;
#pragma unroll
  for (int w = 0; w < PROP_MAXQ / 32; ++w) {
    if (w < warp) { before_h += swh[w]; before_o += swo[w]; }
    all_h += swh[w]; all_o += swo[w]; ok_h += swokh[w]; ok_o += swoko[w];
  }
  const int kh = min(all_h, ok_h < min_inst ? min_inst : (ok_h > max_inst ? max_inst : ok_h));
  const int ko = min(all_o, ok_o < min_inst ? min_inst : (ok_o > max_inst ? max_inst : ok_o));
  if (t == 0) { counts[2 * b] = kh; counts[2 * b + 1] = ko; }
  const unsigned below = (1u << lane) - 1u;
  const int ph = before_h + __popc(bh & below), po = before_o + __popc(bo & below);
  int slot = -1;
  if (hum && ph < kh) slot = ph;
  if (obj && po < ko) slot = kh + po;
  if (slot >= 0) {
    const size_t base = size_t(b) * 2 * max_inst;
    out_boxes[base + slot] = boxes[size_t(b) * Q + me];           // the ORIGINAL (unshifted) box
    out_scores[base + slot] = sscore[me];
    out_labels[base + slot] = labels[size_t(b) * Q + me];
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_prepare_proposals(const float* scores, const int64_t* labels, const float* boxes, int32_t batch, int32_t num_queries,
                             int64_t human_idx, float box_score_thresh, int32_t min_instances, int32_t max_instances,
                             float nms_iou, float* out_boxes, float* out_scores, int64_t* out_labels, int32_t* counts,
                             hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(scores && labels && boxes && out_boxes && out_scores && out_labels && counts, "prepare_proposals: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && num_queries > 0 && num_queries <= PROP_MAXQ,
                   "prepare_proposals: 1..%d candidates per image (got %d)", PROP_MAXQ, num_queries);
  HOIGEN_CHECK_ARG(min_instances >= 0 && max_instances >= min_instances && max_instances > 0, "prepare_proposals: bad instance limits");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0,
                   "prepare_proposals: box arrays must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("prepare_proposals", s, 0, 0);
  prepare_proposals_kernel<<<batch, PROP_MAXQ, 0, s>>>(
      scores, reinterpret_cast<const long long*>(labels), reinterpret_cast<const float4*>(boxes), num_queries, human_idx,
      box_score_thresh, min_instances, max_instances, nms_iou, reinterpret_cast<float4*>(out_boxes), out_scores,
      reinterpret_cast<long long*>(out_labels), counts);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
