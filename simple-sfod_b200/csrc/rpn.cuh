// rpn.cuh -- RPN proposal selection kernels (decode + top-k + clip + NMS + post top-k), all images per launch.
// Replaces d2 RPN.predict_proposals as called at reference daod/modeling/proposal_generator/rpn.py:54-56
// (SURVEY.md A-2/A-3).  The per-image Python loop of find_top_rpn_proposals (with its .item() syncs) becomes
// four launches over all images: key build -> segmented sort -> decode/clip/compact -> NMS mask -> NMS scan+gather.
#pragma once
#include "common.cuh"
#include "sort.cuh"
#include "nms.cuh"

namespace rpnk {

struct CellAnchors { float v[64 * 4]; };

// key = (~score_key << 32) | flat_anchor_index : ascending key order == score descending, index ascending.
__global__ void __launch_bounds__(256) rpn_make_keys_kernel(const float *__restrict__ logits, int HWA, int P,
                                                            unsigned long long *__restrict__ keys) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  unsigned long long k = bsort::kSentinel;
  if (i < HWA) {
    const float s = logits[(size_t)n * HWA + i];
    k = ((unsigned long long)(~sfod_score_key(s)) << 32) | (unsigned)i;
  }
  keys[(size_t)n * P + i] = k;
}

// One CTA per image walks the sorted keys in rank order (chunks of blockDim), decodes the anchor/delta pair,
// applies d2's finite filter, Boxes.clip and nonempty filter, and compacts the survivors in order.
constexpr int kDecodeThreads = 1024;
__global__ void __launch_bounds__(kDecodeThreads) rpn_decode_compact_kernel(
    const unsigned long long *__restrict__ keys, int P, const float *__restrict__ logits,
    const float4 *__restrict__ deltas, const float4 *__restrict__ anchors, CellAnchors cell, int HWA, int A, int Wf,
    int stride, float anchor_offset, float wx, float wy, float ww, float wh, float scale_clamp, int topk,
    float min_box_size, const int *__restrict__ image_hw, float4 *__restrict__ sboxes, float *__restrict__ sscores,
    int *__restrict__ ssrc, nmsk::Seg *__restrict__ segs, int *__restrict__ invalid_count) {
  __shared__ int warp_tot[kDecodeThreads / 32];
  __shared__ int running;
  __shared__ int n_invalid;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { running = 0; n_invalid = 0; }
  __syncthreads();
  const float img_h = (float)image_hw[2 * n], img_w = (float)image_hw[2 * n + 1];
  for (int base = 0; base < topk; base += kDecodeThreads) {
    const int j = base + tid;
    bool keep = false;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float score = 0.f; int idx = 0;
    if (j < topk) {
      const unsigned long long key = keys[(size_t)n * P + j];
      idx = (int)(unsigned)(key & 0xFFFFFFFFull);
      score = logits[(size_t)n * HWA + idx];
      float4 a;
      if (anchors) {
        a = anchors[idx];
      } else {  // d2 DefaultAnchorGenerator grid in closed form: fl(shift + cell), (H, W, A) order
        const int ai = idx % A, cellpos = idx / A;
        const int gx = cellpos % Wf, gy = cellpos / Wf;
        const float sx = __fadd_rn(__fmul_rn(anchor_offset, (float)stride), (float)(gx * stride));
        const float sy = __fadd_rn(__fmul_rn(anchor_offset, (float)stride), (float)(gy * stride));
        a = make_float4(__fadd_rn(sx, cell.v[4 * ai]), __fadd_rn(sy, cell.v[4 * ai + 1]),
                        __fadd_rn(sx, cell.v[4 * ai + 2]), __fadd_rn(sy, cell.v[4 * ai + 3]));
      }
      const float4 d = deltas[(size_t)n * HWA + idx];
      box = sfod_decode_box(a, d, wx, wy, ww, wh, scale_clamp);
      const bool finite = sfod_finite4(box) && isfinite(score);
      if (!finite) atomicAdd(&n_invalid, 1);
      box = sfod_clip_box(box, img_h, img_w);
      keep = finite && (__fsub_rn(box.z, box.x) > min_box_size) && (__fsub_rn(box.w, box.y) > min_box_size);
    }
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (keep) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      sboxes[(size_t)n * topk + pos] = box;
      sscores[(size_t)n * topk + pos] = score;
      ssrc[(size_t)n * topk + pos] = idx;
    }
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w = 0; w < kDecodeThreads / 32; ++w) tot += warp_tot[w]; running += tot; }
    __syncthreads();
  }
  if (tid == 0) {
    segs[n].start = n * topk; segs[n].len = running;
    invalid_count[n] = n_invalid;
  }
}

// Gather the kept ranks into the padded output tensors (zero-filled beyond the count).
__global__ void __launch_bounds__(256) rpn_gather_kernel(const float4 *__restrict__ sboxes, const float *__restrict__ sscores,
                                                         const int *__restrict__ ssrc, const int *__restrict__ keep_rank,
                                                         const int *__restrict__ keep_count, int topk, int post,
                                                         float4 *__restrict__ out_boxes, float *__restrict__ out_logits,
                                                         long long *__restrict__ out_src, int *__restrict__ out_count) {
  const int n = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= post) return;
  const int cnt = keep_count[n];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f); float s = 0.f; long long src = -1;
  if (q < cnt) {
    const int r = keep_rank[(size_t)n * post + q];
    b = sboxes[(size_t)n * topk + r]; s = sscores[(size_t)n * topk + r]; src = ssrc[(size_t)n * topk + r];
  }
  out_boxes[(size_t)n * post + q] = b;
  out_logits[(size_t)n * post + q] = s;
  out_src[(size_t)n * post + q] = src;
  if (q == 0) out_count[n] = cnt;
}

struct Plan {
  int P, topk, wstride;
  size_t bytes;
};
struct Buffers {
  unsigned long long *keys; float4 *sboxes; float *sscores; int *ssrc; nmsk::Seg *segs;
  unsigned long long *mask; int *keep_rank; int *keep_count;
};

template <typename WS>
static inline void carve(WS &ws, const sfod_rpn_params *p, Plan &pl, Buffers *b) {
  pl.P = bsort::next_pow2(p->HWA);
  pl.topk = p->HWA < p->pre_nms_topk ? p->HWA : p->pre_nms_topk;
  pl.wstride = (pl.topk + 63) / 64;
  const size_t N = (size_t)p->N;
  auto k = ws.template take<unsigned long long>(N * pl.P);
  auto sb = ws.template take<float4>(N * pl.topk);
  auto ss = ws.template take<float>(N * pl.topk);
  auto sr = ws.template take<int>(N * pl.topk);
  auto sg = ws.template take<nmsk::Seg>(N);
  auto mk = ws.template take<unsigned long long>(N * pl.topk * pl.wstride);
  auto kr = ws.template take<int>(N * p->post_nms_topk);
  auto kc = ws.template take<int>(N);
  if (b) { b->keys = k; b->sboxes = sb; b->sscores = ss; b->ssrc = sr; b->segs = sg; b->mask = mk; b->keep_rank = kr; b->keep_count = kc; }
}

}  // namespace
