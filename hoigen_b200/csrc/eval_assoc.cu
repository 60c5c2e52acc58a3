// Batched HOI association for evaluation: the inner loop of the reference's eval caller,
// utils_tip_cache_and_union_finetune.py:375-407 with pocket/pocket/utils/association.py:51-125, for a whole batch of
// packed detections in one launch instead of a D2H copy + Python loops per image and per HOI class.
//
// One CTA per image.  For every detection d of the image:
//   interaction(d) = conversion[object(d)][verb(d)]           (-1 = not an HOI class; or = verb when conversion == NULL)
//   among the image's ground-truth pairs g with hoi(g) == interaction(d):
//       iou(g,d) = min(IoU(gt_h[g], det_h[d]), IoU(gt_o[g], det_o[d]))           pair IoU (BoxPairAssociation._iou)
//       d is assigned to the g of maximal iou (FIRST g on ties, as torch.max does)
//   a ground truth's true positive is its assigned detection with iou > min_iou and the highest score (FIRST d on ties,
//   as argmax does) -> labels[d] = 1, every other detection 0.
// The per-ground-truth winner is a 64-bit atomicMax on (score bits, ~d): scores are positive floats, whose IEEE-754
// bit patterns order like unsigned integers.
//
// IoU arithmetic restates torchvision.ops.box_iou operation by operation with round-to-nearest intrinsics (no FMA
// contraction), so the > min_iou decisions and the ties are bit-identical to the reference's fp32 results.
#include "common.h"

namespace hoigen {

constexpr int ASSOC_MAX_GT = 1024;     // ground-truth pairs per image held in shared memory

__device__ __forceinline__ float box_area_rn(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

__device__ __forceinline__ float box_iou_rn(const float4 a, float area_a, const float4 b, float area_b) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  return __fdiv_rn(inter, uni);
}

__global__ void __launch_bounds__(256)
associate_pairs_kernel(const float4* __restrict__ boxes, const int* __restrict__ box_off,
                       const long long* __restrict__ pairing, const long long* __restrict__ objects,
                       const long long* __restrict__ verbs, const float* __restrict__ scores,
                       const int* __restrict__ trip_off, const int* __restrict__ conversion, int num_verbs,
                       const float4* __restrict__ gt_h, const float4* __restrict__ gt_o,
                       const long long* __restrict__ gt_hoi, const int* __restrict__ gt_off, float min_iou,
                       long long* __restrict__ interactions, float* __restrict__ labels) {
  __shared__ unsigned long long s_win[ASSOC_MAX_GT];
  const int b = blockIdx.x;
  const int t0 = trip_off[b], m = trip_off[b + 1] - t0;
  const int g0 = gt_off[b], ng = gt_off[b + 1] - g0;
  const int bb = box_off[b];
  const long long* pair_h = pairing + 2ll * t0;      // [2][m] block of this image
  const long long* pair_o = pair_h + m;
  for (int g = threadIdx.x; g < ng; g += blockDim.x) s_win[g] = 0ull;
  __syncthreads();

  // pass 1: HOI id, best ground truth, candidacy
  for (int d = threadIdx.x; d < m; d += blockDim.x) {
    const long long obj = objects[t0 + d], verb = verbs[t0 + d];
    long long hoi = verb;
    if (conversion != nullptr) hoi = (obj >= 0 && obj < 80 && verb >= 0 && verb < num_verbs) ? conversion[obj * num_verbs + verb] : -1;
    interactions[t0 + d] = hoi;
    labels[t0 + d] = 0.f;
    if (hoi < 0 || ng == 0) continue;
    const float4 dh = boxes[bb + pair_h[d]], dob = boxes[bb + pair_o[d]];
    const float adh = box_area_rn(dh), ado = box_area_rn(dob);
    float best = -1.f;
    int best_g = -1;
    for (int g = 0; g < ng; ++g) {
      if (gt_hoi[g0 + g] != hoi) continue;
      const float4 gh = gt_h[g0 + g], go = gt_o[g0 + g];
      // gt is boxes_1, the detection boxes_2 in the reference's call (association.py:118-124)
      const float iou = fminf(box_iou_rn(gh, box_area_rn(gh), dh, adh), box_iou_rn(go, box_area_rn(go), dob, ado));
      if (iou > best) { best = iou; best_g = g; }       // strict: the first ground truth wins a tie
    }
    if (best_g >= 0 && best > min_iou) {
      const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(scores[t0 + d])) << 32) |
                                     static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned>(d));
      atomicMax(&s_win[best_g], key);
    }
  }
  __syncthreads();
  // pass 2: the winners
  for (int g = threadIdx.x; g < ng; g += blockDim.x) {
    const unsigned long long key = s_win[g];
    if (key != 0ull) labels[t0 + int(0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull))] = 1.f;
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_associate_pairs(const float* boxes, const int32_t* box_off, const int64_t* pairing, const int64_t* objects,
                           const int64_t* verbs, const float* scores, const int32_t* trip_off, const int32_t* conversion,
                           int32_t num_verbs, const float* gt_boxes_h, const float* gt_boxes_o, const int64_t* gt_hoi,
                           const int32_t* gt_off, int32_t max_gt_per_image, int32_t batch, float min_iou,
                           int64_t* interactions, float* labels, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(boxes && box_off && pairing && objects && verbs && scores && trip_off && gt_off && interactions && labels,
                   "associate_pairs: null argument");
  HOIGEN_CHECK_ARG(batch > 0 && num_verbs > 0, "associate_pairs: bad sizes");
  HOIGEN_CHECK_ARG(max_gt_per_image >= 0 && max_gt_per_image <= ASSOC_MAX_GT,
                   "associate_pairs: at most %d ground-truth pairs per image (got %d)", ASSOC_MAX_GT, max_gt_per_image);
  HOIGEN_CHECK_ARG(max_gt_per_image == 0 || (gt_boxes_h && gt_boxes_o && gt_hoi), "associate_pairs: null ground truth");
  HOIGEN_CHECK_ARG((reinterpret_cast<uintptr_t>(boxes) & 15) == 0 && (reinterpret_cast<uintptr_t>(gt_boxes_h) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(gt_boxes_o) & 15) == 0,
                   "associate_pairs: box arrays must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("associate_pairs", s, 0, 0);
  associate_pairs_kernel<<<batch, 256, 0, s>>>(
      reinterpret_cast<const float4*>(boxes), box_off, reinterpret_cast<const long long*>(pairing),
      reinterpret_cast<const long long*>(objects), reinterpret_cast<const long long*>(verbs), scores, trip_off, conversion,
      num_verbs, reinterpret_cast<const float4*>(gt_boxes_h), reinterpret_cast<const float4*>(gt_boxes_o),
      reinterpret_cast<const long long*>(gt_hoi), gt_off, min_iou, reinterpret_cast<long long*>(interactions), labels);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
