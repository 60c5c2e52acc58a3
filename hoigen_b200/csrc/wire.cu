// Compact wire format of one batch's detections for the path's single collective (the NCCL gather of per-image
// detections; SURVEY.md 8e — the reference has no counterpart: single-GPU sequential eval, main_tip_finetune.py:383-388).
//
// The forward emits the reference's own dtypes (U:1421-1425: int64 labels / objects / pairing, fp32 scores) = 36 bytes per
// triplet, of which 32 are int64 indices whose values fit in 5: verb / HOI class < 65536, object class < 256, box index
// < 256.  One record = one batch of one rank, fixed capacity, self-describing, built and parsed ON THE DEVICE (no host
// synchronisation to learn the sizes):
//
//   int32 header[4 + 2*(max_images+1)] : magic 'HOIW', nimg, M (triplets), nbox, triplet_off[0..nimg], box_off[0..nimg]
//   f32   boxes  [nbox*4]
//   f32   scores [M]
//   u16   labels [M]
//   u8    objects[M], u8 human_idx[M], u8 object_idx[M]          (pairing rows, planar)
//
// = 9 bytes per triplet.  hoigen_pack_wire writes a record from the forward's packed outputs, hoigen_unpack_wire widens
// any number of records back to the reference's dtypes and per-image [2][M_b] pairing blocks.
#include "common.h"

namespace hoigen {

constexpr uint32_t WIRE_MAGIC = 0x57494F48u;   // "HOIW"

__device__ __forceinline__ int wire_find(const int* __restrict__ off, int n, int idx) {
  int lo = 0, hi = n;  // off[lo] <= idx < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= idx) lo = mid; else hi = mid;
  }
  return lo;
}

__host__ __device__ __forceinline__ size_t wire_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WireLayout {
  size_t boxes, scores, labels, objects, ph, po, end;
};
__host__ __device__ __forceinline__ WireLayout wire_layout(int hdr_words, int m, int nbox) {
  WireLayout l;
  l.boxes = wire_align(size_t(hdr_words) * 4, 16);
  l.scores = l.boxes + size_t(nbox) * 16;
  l.labels = l.scores + size_t(m) * 4;
  l.objects = l.labels + size_t(m) * 2;
  l.ph = l.objects + size_t(m);
  l.po = l.ph + size_t(m);
  l.end = l.po + size_t(m);
  return l;
}

__global__ void __launch_bounds__(256)
pack_wire_kernel(const float* __restrict__ scores, const int64_t* __restrict__ labels, const int64_t* __restrict__ objects,
                 const int64_t* __restrict__ pairing, const float* __restrict__ boxes, const int* __restrict__ img_off,
                 const int* __restrict__ box_off, int nimg, int max_images, long cap_bytes, uint8_t* __restrict__ rec) {
  const int hdr_words = 4 + 2 * (max_images + 1);
  const int m = img_off[nimg], nbox = box_off[nimg];
  const WireLayout l = wire_layout(hdr_words, m, nbox);
  int32_t* hdr = reinterpret_cast<int32_t*>(rec);
  const bool fits = l.end <= size_t(cap_bytes);
  const long tid = blockIdx.x * long(blockDim.x) + threadIdx.x, nth = long(gridDim.x) * blockDim.x;
  if (tid == 0) {
    hdr[0] = int32_t(WIRE_MAGIC);
    hdr[1] = nimg;
    hdr[2] = fits ? m : -1;        // -1: the record did not fit its capacity (the reader raises)
    hdr[3] = nbox;
  }
  for (long i = tid; i <= nimg; i += nth) {
    hdr[4 + i] = img_off[i];
    hdr[4 + (max_images + 1) + i] = box_off[i];
  }
  if (!fits) return;
  float* o_boxes = reinterpret_cast<float*>(rec + l.boxes);
  float* o_scores = reinterpret_cast<float*>(rec + l.scores);
  uint16_t* o_labels = reinterpret_cast<uint16_t*>(rec + l.labels);
  uint8_t* o_obj = rec + l.objects;
  uint8_t* o_ph = rec + l.ph;
  uint8_t* o_po = rec + l.po;
  for (long i = tid; i < long(nbox) * 4; i += nth) o_boxes[i] = boxes[i];
  for (long i = tid; i < m; i += nth) {
    const int b = wire_find(img_off, nimg, int(i));
    const long ioff = img_off[b], mb = long(img_off[b + 1]) - ioff, j = i - ioff;
    o_scores[i] = scores[i];
    o_labels[i] = uint16_t(labels[i]);
    o_obj[i] = uint8_t(objects[i]);
    o_ph[i] = uint8_t(pairing[2 * ioff + j]);
    o_po[i] = uint8_t(pairing[2 * ioff + mb + j]);
  }
}

// grid.y = record; bases[r] = {triplet base, box base} of record r in the output arrays (-1 = skip the record)
__global__ void __launch_bounds__(256)
unpack_wire_kernel(const uint8_t* __restrict__ records, long record_stride, int max_images, const int64_t* __restrict__ bases,
                   float* __restrict__ scores, int64_t* __restrict__ labels, int64_t* __restrict__ objects,
                   int64_t* __restrict__ pairing, float* __restrict__ boxes) {
  const int r = blockIdx.y;
  const int64_t tbase = bases[2 * r], bbase = bases[2 * r + 1];
  if (tbase < 0) return;
  const uint8_t* rec = records + size_t(r) * record_stride;
  const int hdr_words = 4 + 2 * (max_images + 1);
  const int32_t* hdr = reinterpret_cast<const int32_t*>(rec);
  const int nimg = hdr[1], m = hdr[2], nbox = hdr[3];
  if (uint32_t(hdr[0]) != WIRE_MAGIC || m < 0) return;
  const int* img_off = hdr + 4;
  const WireLayout l = wire_layout(hdr_words, m, nbox);
  const float* i_boxes = reinterpret_cast<const float*>(rec + l.boxes);
  const float* i_scores = reinterpret_cast<const float*>(rec + l.scores);
  const uint16_t* i_labels = reinterpret_cast<const uint16_t*>(rec + l.labels);
  const uint8_t* i_obj = rec + l.objects;
  const uint8_t* i_ph = rec + l.ph;
  const uint8_t* i_po = rec + l.po;
  const long tid = blockIdx.x * long(blockDim.x) + threadIdx.x, nth = long(gridDim.x) * blockDim.x;
  for (long i = tid; i < long(nbox) * 4; i += nth) boxes[bbase * 4 + i] = i_boxes[i];
  for (long i = tid; i < m; i += nth) {
    const int b = wire_find(img_off, nimg, int(i));
    const long ioff = img_off[b], mb = long(img_off[b + 1]) - ioff, j = i - ioff;
    scores[tbase + i] = i_scores[i];
    labels[tbase + i] = int64_t(i_labels[i]);
    objects[tbase + i] = int64_t(i_obj[i]);
    pairing[2 * (tbase + ioff) + j] = int64_t(i_ph[i]);
    pairing[2 * (tbase + ioff) + mb + j] = int64_t(i_po[i]);
  }
}

}  // namespace hoigen

extern "C" {

int64_t hoigen_wire_record_bytes(int32_t max_images, int64_t max_triplets, int64_t max_boxes) {
  using namespace hoigen;
  if (max_images < 0 || max_triplets < 0 || max_boxes < 0 || max_triplets > 0x7fffffff || max_boxes > 0x7fffffff) return -1;
  return int64_t(wire_align(wire_layout(4 + 2 * (max_images + 1), int(max_triplets), int(max_boxes)).end, 16));
}

int hoigen_pack_wire(const float* scores, const int64_t* labels, const int64_t* objects, const int64_t* pairing,
                     const float* boxes, const int32_t* img_off, const int32_t* box_off, int32_t nimg, int32_t max_images,
                     int64_t cap_bytes, void* record, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(scores && labels && objects && pairing && boxes && img_off && box_off && record, "pack_wire: null argument");
  HOIGEN_CHECK_ARG(nimg > 0 && nimg <= max_images, "pack_wire: nimg (%d) must be in [1, max_images = %d]", nimg, max_images);
  HOIGEN_CHECK_ARG(cap_bytes >= int64_t(wire_align(size_t(4 + 2 * (max_images + 1)) * 4, 16)) && (reinterpret_cast<uintptr_t>(record) & 15) == 0,
                   "pack_wire: record must be 16-byte aligned and hold at least the header");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("pack_wire", s, 0, 0);
  pack_wire_kernel<<<2 * num_sms(), 256, 0, s>>>(scores, labels, objects, pairing, boxes, img_off, box_off, nimg, max_images,
                                                  long(cap_bytes), reinterpret_cast<uint8_t*>(record));
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

int hoigen_unpack_wire(const void* records, int32_t n_records, int64_t record_stride, int32_t max_images, const int64_t* bases,
                       float* scores, int64_t* labels, int64_t* objects, int64_t* pairing, float* boxes,
                       hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(records && bases && scores && labels && objects && pairing && boxes, "unpack_wire: null argument");
  HOIGEN_CHECK_ARG(n_records > 0 && n_records <= 65535 && record_stride > 0 && max_images > 0, "unpack_wire: bad sizes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("unpack_wire", s, 0, 0);
  unpack_wire_kernel<<<dim3(64, n_records), 256, 0, s>>>(reinterpret_cast<const uint8_t*>(records), long(record_stride), max_images,
                                                         bases, scores, labels, objects, pairing, boxes);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
