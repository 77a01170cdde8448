// frcnn.cuh -- Fast R-CNN box decode + softmax + score filter + per-class NMS + top-k + pseudo-label filter.
// Replaces FastRCNNOutputLayers.inference (reference ...source_free_adaptive_teacher_roi_heads.py:161), whose
// per-image body is on disk at reference daod/modeling/roi_heads/fast_rcnn.py:108-142, and the confidence
// filter threshold_bbox (reference daod/engine/trainers/source_free_adaptive_teacher.py:167-181).
// All images and classes are processed by the same launches; nothing synchronises with the host.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "nms.cuh"

namespace frk {

constexpr int kMaxK = 128;  // class id must fit the 8-bit key field with room for the sentinel

// key = class(8) | ~score_key(32) | lin(24), lin = row_in_image * K + class.
__device__ __forceinline__ unsigned long long make_key(int cls, float score, int lin) {
  return ((unsigned long long)cls << 56) | ((unsigned long long)(~sfod_score_key(score)) << 24) | (unsigned)lin;
}

__device__ __forceinline__ int image_of_row(const int *__restrict__ row_offsets, int N, int g) {
  int lo = 0, hi = N;  // find i with off[i] <= g < off[i+1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_offsets[mid] <= g) lo = mid; else hi = mid; }
  return lo;
}

// One thread per proposal row.
__global__ void __launch_bounds__(128) frcnn_decode_keys_kernel(
    const float *__restrict__ cls_logits, const float *__restrict__ deltas, const float4 *__restrict__ proposals,
    const int *__restrict__ row_offsets, const int *__restrict__ image_hw, int N, int R, int K, int class_agnostic,
    int rows_stride, int Rmax, int P, float wx, float wy, float ww, float wh, float scale_clamp, float score_thresh,
    float4 *__restrict__ cand_boxes, unsigned long long *__restrict__ keys, int *__restrict__ maxc_bits,
    float *__restrict__ probs_out, float *__restrict__ boxes_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= R) return;
  int img, rl;
  if (rows_stride > 0) {   // packed layout: image i owns rows [i * stride, i * stride + count[i]); the rest is padding
    img = g / rows_stride; rl = g - img * rows_stride;
    if (img >= N || rl >= row_offsets[img]) return;
  } else {
    img = image_of_row(row_offsets, N, g);
    rl = g - row_offsets[img];
  }
  if (rl >= Rmax) return;  // caller bug guard; rows beyond the declared maximum are ignored
  const float img_h = (float)image_hw[2 * img], img_w = (float)image_hw[2 * img + 1];
  const int K1 = K + 1;
  const float *x = cls_logits + (size_t)g * K1;
  // softmax (DESIGN.md "softmax"): max, correctly rounded exp, sequential fp32 sum, division
  float m = x[0];
  for (int k = 1; k < K1; ++k) { const float v = x[k]; if (v > m || v != v) m = v; }
  float s = 0.f;
  for (int k = 0; k < K1; ++k) s = __fadd_rn(s, sfod_exp_cr(__fsub_rn(x[k], m)));
  bool row_ok = true;
  for (int k = 0; k < K1; ++k) {
    const float p = __fdiv_rn(sfod_exp_cr(__fsub_rn(x[k], m)), s);
    if (!isfinite(p)) row_ok = false;
    if (probs_out) probs_out[(size_t)g * K1 + k] = p;
  }
  const float4 pb = proposals[g];
  const int nb = class_agnostic ? 1 : K;
  float4 *cb = cand_boxes + ((size_t)img * Rmax + rl) * K;
  for (int k = 0; k < nb; ++k) {
    const float4 d = *reinterpret_cast<const float4 *>(deltas + ((size_t)g * nb + k) * 4);
    const float4 b = sfod_decode_box(pb, d, wx, wy, ww, wh, scale_clamp);
    if (!sfod_finite4(b)) row_ok = false;
    if (boxes_out) *reinterpret_cast<float4 *>(boxes_out + ((size_t)g * nb + k) * 4) = b;
    const float4 c = sfod_clip_box(b, img_h, img_w);
    if (class_agnostic) { for (int kk = 0; kk < K; ++kk) cb[kk] = c; } else cb[k] = c;
  }
  unsigned long long *kk = keys + (size_t)img * P + (size_t)rl * K;
  float mx = 0.f; bool any = false;
  for (int k = 0; k < K; ++k) {
    const float p = __fdiv_rn(sfod_exp_cr(__fsub_rn(x[k], m)), s);
    const bool cand = row_ok && (p > score_thresh);
    kk[k] = cand ? make_key(k, p, rl * K + k) : bsort::kSentinel;
    if (cand) {
      const float4 c = cb[k];
      mx = fmaxf(mx, fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w))); any = true;
    }
  }
  if (any) atomicMax(maxc_bits + img, __float_as_int(mx));
}

// Per image: class segment boundaries by binary search in the sorted keys.
__global__ void frcnn_segments_kernel(const unsigned long long *__restrict__ keys, int P, int K,
                                      nmsk::Seg *__restrict__ segs, int *__restrict__ cand_count) {
  extern __shared__ int lb[];  // K+1 lower bounds
  const int img = blockIdx.x;
  const unsigned long long *kk = keys + (size_t)img * P;
  for (int k = threadIdx.x; k <= K; k += blockDim.x) {
    // first j with class(key_j) >= k   (sentinel has class 0xFF)
    int lo = 0, hi = P;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(kk[mid] >> 56) >= k) hi = mid; else lo = mid + 1; }
    lb[k] = lo;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    segs[img * K + k].start = img * P + lb[k];
    segs[img * K + k].len = lb[k + 1] - lb[k];
  }
  if (threadIdx.x == 0) cand_count[img] = lb[K];
}

// Gather candidate boxes into sorted order; apply torchvision's coordinate-trick offsets when the CPU path would.
__global__ void __launch_bounds__(256) frcnn_gather_sorted_kernel(const unsigned long long *__restrict__ keys, int P, int K,
                                                                  int Rmax, const float4 *__restrict__ cand_boxes,
                                                                  const int *__restrict__ cand_count,
                                                                  const int *__restrict__ maxc_bits, long long trick_max_n,
                                                                  float4 *__restrict__ sboxes) {
  const int img = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = cand_count[img];
  if (j >= M) return;
  const unsigned long long key = keys[(size_t)img * P + j];
  const int lin = (int)(key & 0xFFFFFFull);
  const int k = (int)(key >> 56);
  float4 b = cand_boxes[(size_t)img * Rmax * K + lin];
  if ((long long)M <= trick_max_n) {
    // offsets = idxs.to(boxes) * (max_coordinate + 1); boxes_for_nms = boxes + offsets[:, None]
    const float off = __fmul_rn((float)k, __fadd_rn(__int_as_float(maxc_bits[img]), 1.0f));
    b = make_float4(__fadd_rn(b.x, off), __fadd_rn(b.y, off), __fadd_rn(b.z, off), __fadd_rn(b.w, off));
  }
  sboxes[(size_t)img * P + j] = b;
}

// Collect the kept candidates of all classes of an image into a second key array (score desc, lin asc).
__global__ void __launch_bounds__(256) frcnn_merge_keys_kernel(const unsigned long long *__restrict__ keys, int P, int K,
                                                               const nmsk::Seg *__restrict__ segs,
                                                               const int *__restrict__ keep_rank,
                                                               const int *__restrict__ keep_count, int topk, int P2,
                                                               unsigned long long *__restrict__ keys2,
                                                               int *__restrict__ total_kept) {
  const int img = blockIdx.x;
  unsigned long long *out = keys2 + (size_t)img * P2;
  int base = 0;
  for (int k = 0; k < K; ++k) {
    const int s = img * K + k;
    const int c = keep_count[s];
    const int j0 = segs[s].start - img * P;
    for (int q = threadIdx.x; q < c; q += blockDim.x) {
      const unsigned long long key = keys[(size_t)img * P + j0 + keep_rank[(size_t)s * topk + q]];
      out[base + q] = (((key >> 24) & 0xFFFFFFFFull) << 32) | (key & 0xFFFFFFull);
    }
    base += c;
  }
  for (int q = base + threadIdx.x; q < P2; q += blockDim.x) out[q] = bsort::kSentinel;
  if (threadIdx.x == 0) total_kept[img] = base;
}

__global__ void frcnn_emit_kernel(const unsigned long long *__restrict__ keys2, int P2, int K, int Rmax,
                                  const float4 *__restrict__ cand_boxes, const int *__restrict__ total_kept, int topk,
                                  float pseudo_thresh, float4 *__restrict__ det_boxes, float *__restrict__ det_scores,
                                  long long *__restrict__ det_classes, long long *__restrict__ det_rows,
                                  int *__restrict__ det_count, int *__restrict__ pseudo_count) {
  const int img = blockIdx.x;
  const int cnt = min(total_kept[img], topk);
  int npseudo = 0;
  for (int q0 = 0; q0 < topk; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool is_pseudo = false;
    if (q < topk) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f); float sc = 0.f; long long cl = -1, row = -1;
      if (q < cnt) {
        const unsigned long long key = keys2[(size_t)img * P2 + q];
        const int lin = (int)(key & 0xFFFFFFull);
        sc = sfod_key_score(~(uint32_t)(key >> 32));
        cl = lin % K; row = lin / K;
        b = cand_boxes[(size_t)img * Rmax * K + lin];
        is_pseudo = sc > pseudo_thresh;
      }
      det_boxes[(size_t)img * topk + q] = b;
      det_scores[(size_t)img * topk + q] = sc;
      det_classes[(size_t)img * topk + q] = cl;
      det_rows[(size_t)img * topk + q] = row;
    }
    npseudo += __syncthreads_count(is_pseudo);
  }
  if (threadIdx.x == 0) { det_count[img] = cnt; pseudo_count[img] = npseudo; }
}

struct Plan { int P, P2, wstride, Rmax; };
struct Buffers {
  float4 *cand_boxes; unsigned long long *keys; int *maxc; nmsk::Seg *segs; int *cand_count; float4 *sboxes;
  unsigned long long *mask; int *keep_rank; int *keep_count; unsigned long long *keys2; int *total_kept;
};

template <typename WS>
static inline void carve(WS &ws, const sfod_frcnn_params *p, Plan &pl, Buffers *b) {
  const size_t N = (size_t)p->N, K = (size_t)p->K;
  pl.Rmax = p->max_rows_per_image;
  pl.P = bsort::next_pow2((long long)pl.Rmax * p->K);
  pl.P2 = bsort::next_pow2((long long)p->K * p->topk);
  pl.wstride = (pl.Rmax + 63) / 64;
  auto cb = ws.template take<float4>(N * pl.Rmax * K);
  auto ky = ws.template take<unsigned long long>(N * pl.P);
  auto mc = ws.template take<int>(N);
  auto sg = ws.template take<nmsk::Seg>(N * K);
  auto cc = ws.template take<int>(N);
  auto sb = ws.template take<float4>(N * pl.P);
  auto mk = ws.template take<unsigned long long>(N * K * (size_t)pl.Rmax * pl.wstride);
  auto kr = ws.template take<int>(N * K * (size_t)p->topk);
  auto kc = ws.template take<int>(N * K);
  auto k2 = ws.template take<unsigned long long>(N * pl.P2);
  auto tk = ws.template take<int>(N);
  if (b) { b->cand_boxes = cb; b->keys = ky; b->maxc = mc; b->segs = sg; b->cand_count = cc; b->sboxes = sb; b->mask = mk;
           b->keep_rank = kr; b->keep_count = kc; b->keys2 = k2; b->total_kept = tk; }
}

}  // namespace
