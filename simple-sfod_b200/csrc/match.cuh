// match.cuh -- fused pairwise_iou + Matcher (student-side proposal / anchor labelling).
//
// Replaces, per image, `pairwise_iou(gt_boxes, boxes)` followed by detectron2's `Matcher.__call__` as reached from
// RPN.label_and_sample_anchors (called by reference daod/modeling/proposal_generator/rpn.py:45) and from
// label_and_sample_proposals (reference daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:165-215).
// The (M x N) quality matrix never exists: every prediction walks the M ground-truth boxes (staged in shared memory) once
// for its (max, first argmax, label) and -- with allow_low_quality_matches -- once more against the per-gt maxima.
// IoU arithmetic is detectron2's, separately rounded: inter = w * h (clamped at 0), iou = inter > 0 ?
// inter / ((area_gt + area_box) - inter) : 0; max/argmax follow torch.max(dim=0) (first maximal index).
#pragma once
#include "common.cuh"

namespace matchk {

constexpr int kThreads = 256;
constexpr int kGtTile = 512;     // ground-truth boxes staged per pass
constexpr int kMaxThresholds = 8;

struct Thresholds { float t[kMaxThresholds]; int label[kMaxThresholds + 1]; int count; };

__device__ __forceinline__ float pair_iou(const float4 g, const float ag, const float4 b, const float ab) {
  const float w = fmaxf(__fsub_rn(fminf(g.z, b.z), fmaxf(g.x, b.x)), 0.0f);
  const float h = fmaxf(__fsub_rn(fminf(g.w, b.w), fmaxf(g.y, b.y)), 0.0f);
  const float inter = __fmul_rn(w, h);
  return inter > 0.0f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(ag, ab), inter)) : 0.0f;
}
__device__ __forceinline__ float box_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// pass 1: per prediction max / first argmax / threshold label; per-gt maxima (IoU >= 0, so the int order of the bits is the
// float order and atomicMax on the bit pattern is exact)
__global__ void __launch_bounds__(kThreads) match_kernel(const float4 *__restrict__ gt, const float4 *__restrict__ boxes, int M, int N,
                                                         Thresholds th, long long *__restrict__ matches,
                                                         signed char *__restrict__ labels, float *__restrict__ vals,
                                                         int *__restrict__ gt_max_bits) {
  __shared__ float4 sg[kGtTile];
  __shared__ float sa[kGtTile];
  const int j = blockIdx.x * kThreads + threadIdx.x;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < N) b = boxes[j];
  const float ab = box_area(b);
  float best = -1.0f;   // every IoU is >= 0 (or NaN, which never wins): the first gt always takes over
  int best_i = 0;
  for (int i0 = 0; i0 < M; i0 += kGtTile) {
    const int m = min(kGtTile, M - i0);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += kThreads) { const float4 g = gt[i0 + i]; sg[i] = g; sa[i] = box_area(g); }
    __syncthreads();
    if (j < N) {
      for (int i = 0; i < m; ++i) {
        const float v = pair_iou(sg[i], sa[i], b, ab);
        if (v > best) { best = v; best_i = i0 + i; }
        if (gt_max_bits && v > 0.0f) atomicMax(&gt_max_bits[i0 + i], __float_as_int(v));
      }
    }
  }
  if (j >= N) return;
  int k = 0;
#pragma unroll
  for (int q = 0; q < kMaxThresholds; ++q) k += (q < th.count && best >= th.t[q]) ? 1 : 0;
  matches[j] = best_i;
  labels[j] = (signed char)th.label[k];
  if (vals) vals[j] = best;
}

// pass 2 (allow_low_quality_matches): predictions whose IoU with some gt EQUALS that gt's maximum get label 1 -- including
// every prediction with IoU 0 to a gt that overlaps nothing (detectron2's `mqm == mqm.max(dim=1)[:, None]`, literally)
__global__ void __launch_bounds__(kThreads) low_quality_kernel(const float4 *__restrict__ gt, const float4 *__restrict__ boxes, int M,
                                                               int N, const int *__restrict__ gt_max_bits,
                                                               signed char *__restrict__ labels) {
  __shared__ float4 sg[kGtTile];
  __shared__ float sa[kGtTile];
  __shared__ float smax[kGtTile];
  const int j = blockIdx.x * kThreads + threadIdx.x;
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < N) b = boxes[j];
  const float ab = box_area(b);
  bool hit = false;
  for (int i0 = 0; i0 < M; i0 += kGtTile) {
    const int m = min(kGtTile, M - i0);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += kThreads) {
      const float4 g = gt[i0 + i]; sg[i] = g; sa[i] = box_area(g); smax[i] = __int_as_float(gt_max_bits[i0 + i]);
    }
    __syncthreads();
    if (j < N && !hit) {
      for (int i = 0; i < m; ++i)
        if (pair_iou(sg[i], sa[i], b, ab) == smax[i]) { hit = true; break; }
    }
  }
  if (j < N && hit) labels[j] = 1;
}

}  // namespace matchk
