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
// kNative: the logits lie as the head wrote them, (N, A, Hf*Wf); slot j = a * HW + pos is read coalesced and carries the
// FLATTENED index pos * A + a in its key (the sort does not care where a key starts), so the permute copy of
// reference rpn.py:28-33 is never made.
template <bool kNative>
__global__ void __launch_bounds__(256) rpn_make_keys_kernel(const float *__restrict__ logits, int HWA, int P, int A, int HW,
                                                            unsigned long long *__restrict__ keys) {
  const int n = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  unsigned long long k = bsort::kSentinel;
  if (i < HWA) {
    const float s = logits[(size_t)n * HWA + i];
    const unsigned flat = kNative ? (unsigned)((i % HW) * A + i / HW) : (unsigned)i;
    k = ((unsigned long long)(~sfod_score_key(s)) << 32) | flat;
  }
  keys[(size_t)n * P + i] = k;
}

// One CTA per image walks the sorted keys in rank order, decodes the anchor/delta pair, applies d2's finite filter,
// Boxes.clip and nonempty filter, and compacts the survivors in order.  A round covers kDecodeE x 1024 consecutive ranks:
// every thread has kDecodeE independent key -> (logit, delta) gather chains in flight and the order-preserving offsets of
// the whole round come from one warp-level scan of the kDecodeE x 32 ballot counts (3 barriers per 4096 ranks).
constexpr int kDecodeThreads = 1024;
constexpr int kDecodeE = 4;
// kNative: logits (N, A, HW) and deltas (N, 4A, HW) as the head wrote them -- the (N, HWA, 4) copy of reference
// rpn.py:34-41 is replaced by four 4-byte gathers per selected anchor.
template <bool kNative>
__global__ void __launch_bounds__(kDecodeThreads) rpn_decode_compact_kernel(
    const unsigned long long *__restrict__ keys, int P, const float *__restrict__ logits,
    const float *__restrict__ deltas, const float4 *__restrict__ anchors, CellAnchors cell, int HWA, int A, int Wf, int HW,
    int stride, float anchor_offset, float wx, float wy, float ww, float wh, float scale_clamp, int topk,
    float min_box_size, const int *__restrict__ image_hw, float4 *__restrict__ sboxes, float *__restrict__ sscores,
    int *__restrict__ ssrc, nmsk::Seg *__restrict__ segs, int *__restrict__ invalid_count) {
  __shared__ int warp_off[kDecodeE][kDecodeThreads / 32];   // ballot counts, then exclusive offsets within the round
  __shared__ int round_total;
  __shared__ int n_invalid;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) n_invalid = 0;
  __syncthreads();
  const float img_h = (float)image_hw[2 * n], img_w = (float)image_hw[2 * n + 1];
  int running = 0;   // uniform: survivors of the rounds before
  for (int base = 0; base < topk; base += kDecodeE * kDecodeThreads) {
    unsigned long long key[kDecodeE];
#pragma unroll
    for (int e = 0; e < kDecodeE; ++e) {
      const int j = base + e * kDecodeThreads + tid;
      key[e] = j < topk ? keys[(size_t)n * P + j] : 0ull;
    }
    float score[kDecodeE]; float4 dl[kDecodeE], an[kDecodeE]; int idx[kDecodeE];
#pragma unroll
    for (int e = 0; e < kDecodeE; ++e) {
      const int j = base + e * kDecodeThreads + tid;
      idx[e] = (int)(unsigned)(key[e] & 0xFFFFFFFFull);
      if (j >= topk) idx[e] = 0;
      if (kNative) {
        const int ai = idx[e] % A, cellpos = idx[e] / A;
        score[e] = logits[(size_t)n * HWA + (size_t)ai * HW + cellpos];
        const float *d = deltas + ((size_t)n * HWA + (size_t)ai * HW) * 4 + cellpos;
        dl[e] = make_float4(d[0], d[HW], d[2 * (size_t)HW], d[3 * (size_t)HW]);
      } else {
        score[e] = logits[(size_t)n * HWA + idx[e]];
        dl[e] = reinterpret_cast<const float4 *>(deltas)[(size_t)n * HWA + idx[e]];
      }
      if (anchors) an[e] = anchors[idx[e]];
    }
    bool keep[kDecodeE]; float4 box[kDecodeE]; unsigned bal[kDecodeE];
#pragma unroll
    for (int e = 0; e < kDecodeE; ++e) {
      const int j = base + e * kDecodeThreads + tid;
      float4 a;
      if (anchors) {
        a = an[e];
      } else {  // d2 DefaultAnchorGenerator grid in closed form: fl(shift + cell), (H, W, A) order
        const int ai = idx[e] % A, cellpos = idx[e] / A;
        const int gx = cellpos % Wf, gy = cellpos / Wf;
        const float sx = __fadd_rn(__fmul_rn(anchor_offset, (float)stride), (float)(gx * stride));
        const float sy = __fadd_rn(__fmul_rn(anchor_offset, (float)stride), (float)(gy * stride));
        a = make_float4(__fadd_rn(sx, cell.v[4 * ai]), __fadd_rn(sy, cell.v[4 * ai + 1]),
                        __fadd_rn(sx, cell.v[4 * ai + 2]), __fadd_rn(sy, cell.v[4 * ai + 3]));
      }
      float4 bx = sfod_decode_box(a, dl[e], wx, wy, ww, wh, scale_clamp);
      const bool finite = sfod_finite4(bx) && isfinite(score[e]);
      if (j < topk && !finite) atomicAdd(&n_invalid, 1);
      bx = sfod_clip_box(bx, img_h, img_w);
      keep[e] = j < topk && finite && (__fsub_rn(bx.z, bx.x) > min_box_size) && (__fsub_rn(bx.w, bx.y) > min_box_size);
      box[e] = bx;
      bal[e] = __ballot_sync(0xFFFFFFFFu, keep[e]);
      if (lane == 0) warp_off[e][warp] = __popc(bal[e]);
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the kDecodeE x 32 counts in rank order (e major, warp minor)
      int carry = 0;
#pragma unroll
      for (int e = 0; e < kDecodeE; ++e) {
        const int v = warp_off[e][lane];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        warp_off[e][lane] = carry + incl - v;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
      }
      if (lane == 0) round_total = carry;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < kDecodeE; ++e) {
      if (keep[e]) {
        const int pos = running + warp_off[e][warp] + __popc(bal[e] & ((1u << lane) - 1u));
        sboxes[(size_t)n * topk + pos] = box[e];
        sscores[(size_t)n * topk + pos] = score[e];
        ssrc[(size_t)n * topk + pos] = idx[e];
      }
    }
    running += round_total;
    __syncthreads();   // warp_off / round_total are rewritten by the next round
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
