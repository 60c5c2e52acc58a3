// Per-class 11-point interpolated average precision: DetectionAPMeter.compute_pr_for_each +
// AveragePrecisionMeter.compute_per_class_ap_with_11_point_interpolation, pocket/pocket/utils/meters.py:561-583, 255-270,
// for ALL classes in one launch (the reference loops over 600 classes on the host, optionally in a process pool).
//
// Input: the sweep's detections already ordered by (class ascending, score descending) — `labels` (0/1, fp32) in that
// order and CSR class offsets.  One CTA per class:
//   tp_i   = inclusive prefix sum of the labels           (block scan, chunk by chunk with a carry; exact integers)
//   prec_i = tp_i / (i + 1)          rec_i = tp_i / num_gt[c]   (or / total true positives when num_gt[c] < 0;
//                                            x / 0 := 0 as the reference's `div` helper defines)                  fp64
//   ap     = sum over the 11 thresholds t of  max{prec_i : rec_i >= t} / 11, skipping thresholds nothing reaches,
//            accumulated in threshold order exactly as the reference's Python loop does (fp64)
//   max_rec = rec_{N-1}
// The thresholds are passed in (torch.linspace(0, 1, 11, dtype=float64) on the host) so they are the reference's bits.
#include "common.h"

namespace hoigen {

constexpr int AP_THREADS = 256;
constexpr int AP_T = 11;

__global__ void __launch_bounds__(AP_THREADS)
ap11_kernel(const float* __restrict__ labels, const long long* __restrict__ class_off, const double* __restrict__ num_gt,
            const double* __restrict__ thresholds, double* __restrict__ ap, double* __restrict__ max_rec) {
  __shared__ long long s_warp[AP_THREADS / 32];
  __shared__ long long s_carry;
  __shared__ double s_max[AP_T][AP_THREADS / 32];
  const int c = blockIdx.x;
  const long long beg = class_off[c], n = class_off[c + 1] - beg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (n <= 0) {
    if (tid == 0) { ap[c] = 0.0; max_rec[c] = 0.0; }
    return;
  }
  // total true positives (needed up front when num_gt is not given)
  long long local = 0;
  for (long long i = tid; i < n; i += AP_THREADS) local += labels[beg + i] != 0.f;
  for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if (lane == 0) s_warp[warp] = local;
  __syncthreads();
  long long total_tp = 0;
  for (int w = 0; w < AP_THREADS / 32; ++w) total_tp += s_warp[w];
  __syncthreads();
  const double denom = num_gt[c] >= 0.0 ? num_gt[c] : double(total_tp);
  double thr[AP_T], best[AP_T];
#pragma unroll
  for (int k = 0; k < AP_T; ++k) { thr[k] = thresholds[k]; best[k] = -1.0; }
  if (tid == 0) s_carry = 0;
  __syncthreads();
  double last_rec = 0.0;
  for (long long base = 0; base < n; base += AP_THREADS) {
    const long long i = base + tid;
    const long long v = (i < n && labels[beg + i] != 0.f) ? 1 : 0;
    long long incl = v;                                   // warp inclusive scan
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long offset = s_carry;
    for (int w = 0; w < warp; ++w) offset += s_warp[w];
    const long long tp = offset + incl;
    if (i < n) {
      const double prec = double(tp) / double(i + 1);
      const double rec = denom == 0.0 ? 0.0 : double(tp) / denom;     // meters.py:24-30 `div`: x / 0 := 0
#pragma unroll
      for (int k = 0; k < AP_T; ++k)
        if (rec >= thr[k] && prec > best[k]) best[k] = prec;
      if (i == n - 1) last_rec = rec;
    }
    __syncthreads();
    if (tid == AP_THREADS - 1) s_carry = tp;              // thread 255's inclusive value = running total
    __syncthreads();
  }
  // block-wide max per threshold
#pragma unroll
  for (int k = 0; k < AP_T; ++k) {
    double b = best[k];
    for (int o = 16; o; o >>= 1) b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane == 0) s_max[k][warp] = b;
  }
  // the thread that saw the last element publishes max_rec
  if (((n - 1) % AP_THREADS) == tid) max_rec[c] = last_rec;
  __syncthreads();
  if (tid == 0) {
    double acc = 0.0;
    for (int k = 0; k < AP_T; ++k) {
      double b = -1.0;
      for (int w = 0; w < AP_THREADS / 32; ++w) b = fmax(b, s_max[k][w]);
      if (b >= 0.0) acc += b / 11.0;                      // `ap += prec[inds].max() / 11` in threshold order
    }
    ap[c] = acc;
  }
}

}  // namespace hoigen

extern "C" {

int hoigen_ap_11point(const float* labels_sorted, const int64_t* class_off, const double* num_gt, const double* thresholds,
                      int32_t num_classes, double* ap, double* max_rec, hoigen_stream_t stream) {
  using namespace hoigen;
  HOIGEN_CHECK_ARG(class_off && num_gt && thresholds && ap && max_rec && num_classes > 0, "ap_11point: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  KernelScope ks("ap_11point", s, 0, 0);
  ap11_kernel<<<num_classes, AP_THREADS, 0, s>>>(labels_sorted, reinterpret_cast<const long long*>(class_off), num_gt,
                                                 thresholds, ap, max_rec);
  HOIGEN_CHECK_LAUNCH();
  return HOIGEN_OK;
}

}  // extern "C"
