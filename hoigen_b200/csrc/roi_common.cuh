// RoIAlign sampling rule shared by the axis-weight kernel (hoi_head.cu) and the tensor-core RoIAlign kernel (roi_tc.cu).
// torchvision.ops.roi_align semantics, aligned = True, sampling_ratio = -1, 7 x 7 bins on the 14 x 14 token grid
// (call sites U:1028-1029).
#pragma once

namespace hoigen {

constexpr int G14 = 14;
constexpr int POOL = 7;

// Sum of the bilinear weights that grid line `t` (0..13) of one axis receives from the 7 * g sample positions of a box whose
// scaled extent on that axis starts at `start` with bin size `bin` (g = samples per bin).  Samples with coordinate < -1 or
// > 14 contribute nothing, exactly as in torchvision's kernel.
__device__ __forceinline__ float axis_weight(int t, float start, float bin, int g) {
  float w = 0.f;
  const float gf = float(g);
  for (int p = 0; p < POOL; ++p) {
    for (int i = 0; i < g; ++i) {
      float c = start + float(p) * bin + (float(i) + 0.5f) * bin / gf;
      if (c < -1.0f || c > float(G14)) continue;
      c = fmaxf(c, 0.f);
      int lo = int(c), hi;
      if (lo >= G14 - 1) { lo = hi = G14 - 1; c = float(lo); } else { hi = lo + 1; }
      const float l = c - float(lo);
      if (t == lo) w += 1.0f - l;
      if (t == hi) w += l;
    }
  }
  return w;
}

// (Wy[0..14), Wx[0..14), 1 / (49 count)) of one box for lane-like index `idx` in [0, 32): 0..13 -> Wy, 16..29 -> Wx, 31 -> inv
__device__ __forceinline__ float roi_axis_entry(int idx, float x1, float y1, float x2, float y2, float spatial_scale) {
  const float sx = x1 * spatial_scale - 0.5f, sy = y1 * spatial_scale - 0.5f;
  const float ex = x2 * spatial_scale - 0.5f, ey = y2 * spatial_scale - 0.5f;
  const float rw = ex - sx, rh = ey - sy;
  const float bw = rw / float(POOL), bh_ = rh / float(POOL);
  const int gw = int(ceilf(rw / float(POOL))), gh = int(ceilf(rh / float(POOL)));
  if (idx < G14) return axis_weight(idx, sy, bh_, gh);
  if (idx >= 16 && idx < 16 + G14) return axis_weight(idx - 16, sx, bw, gw);
  if (idx == 31) return 1.0f / (float(max(gh * gw, 1)) * float(POOL * POOL));
  return 0.f;
}

// The same weights written a whole axis at a time by ONE owner thread into `row[0..14)` (shared memory): every sample adds
// its two bilinear weights to the grid lines it touches, in the (bin, sample) order axis_weight visits them, so each entry
// is the bit-identical sum -- at 1/14 of the work of evaluating axis_weight per entry.
__device__ __forceinline__ void axis_weights_row(float* row, float start, float bin, int g) {
#pragma unroll
  for (int t = 0; t < G14; ++t) row[t] = 0.f;
  const float gf = float(g);
  for (int p = 0; p < POOL; ++p) {
    for (int i = 0; i < g; ++i) {
      float c = start + float(p) * bin + (float(i) + 0.5f) * bin / gf;
      if (c < -1.0f || c > float(G14)) continue;
      c = fmaxf(c, 0.f);
      int lo = int(c), hi;
      if (lo >= G14 - 1) { lo = hi = G14 - 1; c = float(lo); } else { hi = lo + 1; }
      const float l = c - float(lo);
      row[lo] += 1.0f - l;
      row[hi] += l;
    }
  }
}

}  // namespace hoigen
