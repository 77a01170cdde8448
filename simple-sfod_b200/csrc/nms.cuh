// nms.cuh -- segmented greedy NMS on score-sorted boxes (torchvision semantics, bit-exact keep sets).
//
// Serves d2 find_top_rpn_proposals -> batched_nms (reference rpn.py:54-56; one segment per image) and
// fast_rcnn_inference_single_image -> batched_nms (reference roi_heads/fast_rcnn.py:133; one segment per
// (image, class)), plus the torchvision-compatible nms/batched_nms operators.
//
// Two kernels:
//  1. nms_mask_kernel: 64x64 tiles of the upper triangle; each thread owns one row box and ballots a
//     64-bit suppression word against 64 column boxes staged in shared memory (float4 + area).
//     IoU in torchvision's fp32 operation order, division skipped when the intersection is empty.
//  2. nms_scan_kernel: one CTA per segment walks the tiles in order.  Warp 0 resolves the 64x64 diagonal
//     block with a fully unrolled register chain (alive &= ~row_i if alive_i), the other warps then OR
//     the kept rows into the shared `removed` bitmap (coalesced 8-byte loads).  The scan stops as soon
//     as `max_keep` boxes are kept (post_nms_topk / detections-per-image), which is what bounds its
//     latency in the low-suppression case.
#pragma once
#include "common.cuh"

namespace nmsk {

struct Seg { int start; int len; };  // range in the sorted box array

// mask layout: segment s owns rows [row_base[s], row_base[s]+len) of a matrix with `wstride` words/row.
// Row r of segment s, column word w: mask[(row_base(s) + r) * wstride + w], row_base(s) = seg.start + slab offset.

template <bool kClassAware>
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4 *__restrict__ boxes, const int *__restrict__ cls,
                                                      const Seg *__restrict__ segs, const int *__restrict__ seg_of_block_z,
                                                      int rows_per_slab, int wstride, double thr,
                                                      unsigned long long *__restrict__ mask) {
  const int s = blockIdx.z;
  (void)seg_of_block_z;
  const Seg seg = segs[s];
  const int rt = blockIdx.y, ct = blockIdx.x;
  if (ct < rt) return;
  const int row0 = rt * 64, col0 = ct * 64;
  if (row0 >= seg.len || col0 >= seg.len) return;
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  __shared__ int cc[64];
  const int tid = threadIdx.x;
  const int ncol = min(64, seg.len - col0);
  if (tid < ncol) {
    const float4 b = boxes[seg.start + col0 + tid];
    cb[tid] = b; ca[tid] = sfod_box_area(b);
    if (kClassAware) cc[tid] = cls[seg.start + col0 + tid];
  }
  __syncthreads();
  const int r = row0 + tid;
  if (r >= seg.len) return;
  const float4 a = boxes[seg.start + r];
  const float aa = sfod_box_area(a);
  const int ac = kClassAware ? cls[seg.start + r] : 0;
  unsigned long long word = 0;
  const int jstart = (rt == ct) ? tid + 1 : 0;
  const bool neg_thr = thr < 0.0;
  for (int j = jstart; j < ncol; ++j) {
    const float4 b = cb[j];
    if (kClassAware && cc[j] != ac) continue;
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    if (inter > 0.0f || neg_thr) {  // inter == 0 => iou is 0, -0 or NaN: never > thr for thr >= 0
      const float uni = __fsub_rn(__fadd_rn(aa, ca[j]), inter);
      const float ovr = __fdiv_rn(inter, uni);
      if ((double)ovr > thr) word |= 1ull << j;
    }
  }
  mask[((size_t)s * rows_per_slab + r) * wstride + ct] = word;
}

constexpr int kScanThreads = 512;

// keep_rank: (S, keep_stride) int32 ranks (position in the segment's sorted order) of kept boxes;
// keep_count: (S) int32.
__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(const unsigned long long *__restrict__ mask,
                                                                const Seg *__restrict__ segs, int rows_per_slab,
                                                                int wstride, int max_keep, int keep_stride,
                                                                int *__restrict__ keep_rank, int *__restrict__ keep_count) {
  extern __shared__ unsigned long long removed[];  // wstride words
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long kept_bits_sh;
  __shared__ int count_sh;
  const int s = blockIdx.x;
  const Seg seg = segs[s];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = (seg.len + 63) >> 6;
  for (int w = tid; w < W; w += kScanThreads) removed[w] = 0;
  if (tid == 0) count_sh = 0;
  const unsigned long long *m = mask + (size_t)s * rows_per_slab * wstride;
  int *out = keep_rank + (size_t)s * keep_stride;
  // prefetch the diagonal block of tile 0
  unsigned long long dnext = 0;
  if (tid < 64 && tid < seg.len) dnext = m[(size_t)tid * wstride + 0];
  __syncthreads();
  for (int t = 0; t < W; ++t) {
    if (tid < 64) diag[tid] = dnext;
    __syncthreads();
    // prefetch next diagonal block (independent of this tile's outcome)
    if (tid < 64) {
      const int r = (t + 1) * 64 + tid;
      dnext = (t + 1 < W && r < seg.len) ? m[(size_t)r * wstride + (t + 1)] : 0ull;
    }
    if (warp == 0) {
      const int nvalid = min(64, seg.len - t * 64);
      unsigned long long alive = (nvalid == 64 ? ~0ull : ((1ull << nvalid) - 1ull)) & ~removed[t];
      unsigned alo = (unsigned)alive, ahi = (unsigned)(alive >> 32);
      // rows into registers (broadcast LDS), then a fully unrolled dependent chain
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long row = diag[i];
        const unsigned rlo = (unsigned)row, rhi = (unsigned)(row >> 32);
        const unsigned msk = 0u - ((alo >> i) & 1u);
        alo &= ~(rlo & msk);
        ahi &= ~(rhi & msk);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned long long row = diag[32 + i];
        const unsigned rhi = (unsigned)(row >> 32);
        const unsigned msk = 0u - ((ahi >> i) & 1u);
        ahi &= ~(rhi & msk);  // rows >= 32 only suppress bits > 32+i (upper triangle)
      }
      unsigned long long kept = ((unsigned long long)ahi << 32) | alo;
      // cap at max_keep, emit ranks in order
      const int base = count_sh;
      int nk = __popcll(kept);
      if (base + nk > max_keep) {
        // drop the highest bits beyond the cap
        int allow = max_keep - base;
        unsigned long long kk = kept, keep2 = 0;
        for (int q = 0; q < allow; ++q) { unsigned long long low = kk & (0ull - kk); keep2 |= low; kk ^= low; }
        kept = keep2; nk = allow;
      }
      // lane handles bits lane and lane+32
      const unsigned klo = (unsigned)kept, khi = (unsigned)(kept >> 32);
      if ((klo >> lane) & 1u) out[base + __popc(klo & ((1u << lane) - 1u))] = t * 64 + lane;
      if ((khi >> lane) & 1u) out[base + __popc(klo) + __popc(khi & ((1u << lane) - 1u))] = t * 64 + 32 + lane;
      __syncwarp();
      if (lane == 0) { kept_bits_sh = kept; count_sh = base + nk; }
    }
    __syncthreads();
    const unsigned long long kept = kept_bits_sh;
    const int cnt = count_sh;
    if (cnt >= max_keep) break;  // uniform
    // push: OR the kept rows of this tile into removed[t+1 .. W)
    // warp q takes kept rows q, q+16, ... (in bit order); lanes stride over words
    {
      unsigned long long kk = kept;
      int idx = 0;
      while (kk) {
        const int bit = __ffsll((long long)kk) - 1;
        kk &= kk - 1;
        if ((idx & 15) == warp) {
          const unsigned long long *row = m + (size_t)(t * 64 + bit) * wstride;
          for (int w = t + 1 + lane; w < W; w += 32) {
            const unsigned long long v = row[w];
            if (v) atomicOr(&removed[w], v);
          }
        }
        ++idx;
      }
    }
    __syncthreads();
  }
  if (tid == 0) keep_count[s] = count_sh < max_keep ? count_sh : max_keep;
}

// host helpers
static inline int launch_mask(const float4 *boxes, const int *cls, const Seg *segs, int S, int max_len, int rows_per_slab,
                              int wstride, double thr, unsigned long long *mask, cudaStream_t stream) {
  if (S <= 0 || max_len <= 0) return SFOD_OK;
  const int tiles = (max_len + 63) / 64;
  dim3 grid(tiles, tiles, S);
  if (cls) nms_mask_kernel<true><<<grid, 64, 0, stream>>>(boxes, cls, segs, nullptr, rows_per_slab, wstride, thr, mask);
  else nms_mask_kernel<false><<<grid, 64, 0, stream>>>(boxes, nullptr, segs, nullptr, rows_per_slab, wstride, thr, mask);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}
static inline int launch_scan(const unsigned long long *mask, const Seg *segs, int S, int rows_per_slab, int wstride,
                              int max_keep, int keep_stride, int *keep_rank, int *keep_count, cudaStream_t stream) {
  if (S <= 0) return SFOD_OK;
  const size_t smem = (size_t)(wstride > 0 ? wstride : 1) * sizeof(unsigned long long);
  if (smem > 48 * 1024) return SFOD_ERR_UNSUPPORTED;  // > 393k boxes per segment
  nms_scan_kernel<<<S, kScanThreads, smem, stream>>>(mask, segs, rows_per_slab, wstride, max_keep, keep_stride, keep_rank,
                                                     keep_count);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

}  // namespace nmsk
