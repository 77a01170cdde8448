// nms.cuh -- segmented greedy NMS on score-sorted boxes (torchvision semantics, bit-exact keep sets).
//
// Serves d2 find_top_rpn_proposals -> batched_nms (reference rpn.py:54-56; one segment per image) and
// fast_rcnn_inference_single_image -> batched_nms (reference roi_heads/fast_rcnn.py:133; one segment per
// (image, class)), plus the torchvision-compatible nms/batched_nms operators.
//
// Three implementations, chosen by run_segmented():
//  A. nms_deferred_kernel (default): one thread-block cluster per segment; every CTA keeps the complete kept list in
//     shared memory, tiles meet it only near their turn, owners publish through DSMEM + one mbarrier per tile.
//  B. nms_lazy_kernel: the same cluster organisation with every kept list applied to all later boxes at once; used when
//     the kept list cannot fit in shared memory (plain nms of many boxes with max_keep = n).
//  C. nms_mask_kernel + nms_scan_kernel (fallback when clusters cannot launch):
//     mask: 64x64 tiles of the upper triangle; each thread owns one row box and ballots a 64-bit suppression word
//     against 64 column boxes staged in shared memory; scan: one CTA per segment walks the tiles in order, warp 0 resolves
//     the diagonal block, the other warps OR the kept rows into the shared `removed` bitmap.
// All of them use torchvision's fp32 IoU operation order, skip the division when the intersection is empty, give the same
// keep set and order, and stop as soon as `max_keep` boxes are kept (post_nms_topk / detections-per-image).
#pragma once
#include "common.cuh"

namespace nmsk {

struct Seg { int start; int len; };  // range in the sorted box array

// mask layout: segment s owns rows [row_base[s], row_base[s]+len) of a matrix with `wstride` words/row.
// Row r of segment s, column word w: mask[(row_base(s) + r) * wstride + w], row_base(s) = seg.start + slab offset.

template <bool kClassAware>
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4 *__restrict__ boxes, const int *__restrict__ cls,
                                                      const Seg *__restrict__ segs,
                                                      int rows_per_slab, int wstride, double thr,
                                                      unsigned long long *__restrict__ mask) {
  const int s = blockIdx.z;
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

// ---------------------------------------------------------------------------------------------------------------
// Lazy-row cluster NMS (implementation B).  Greedy NMS only ever needs the IoU rows of boxes that end up KEPT, and a
// box stops being tested the moment something suppresses it; the mask kernel above computes all n^2/2 pairs (50 M for
// the 9 990 RPN candidates of one VGG image) although at most max_keep * n of them matter.  Here one thread-block
// CLUSTER owns one segment (image, or image x class) and evaluates pairs on the fly from shared memory:
//   * the sorted boxes are dealt to the CL CTAs of the cluster tile-cyclically (tile t = 64 boxes -> CTA t % CL); each CTA
//     keeps its own boxes, areas, a `removed` bitmap and the 64x64 intra-tile suppression bits of its tiles (computed
//     once, up front) in shared memory -- the n x n/64 bitmask never exists;
//   * tile t is resolved by its owner with a warp-parallel fixed-point iteration (kept_j = alive_j and no kept
//     suppressor among the lower bits; converges in (suppression-chain depth + 1) ballots instead of a 64-step serial
//     chain), and the compacted kept boxes of the tile are PUSHED into every CTA's shared memory (DSMEM stores);
//   * one split cluster barrier per tile: the owner of tile t+1 applies the kept list of tile t to that tile first,
//     resolves and publishes it, ARRIVES, and only then tests its remaining boxes, so the serial resolve of the next
//     tile overlaps the parallel apply phase of the current one; kept lists are triple-buffered;
//   * every CTA tests its own later, still-alive boxes against the kept list (early exit on the first hit; when few
//     boxes remain several threads share one box and split the list).
// The scan stops as soon as max_keep boxes are kept.  IoU arithmetic and tie order are those of the mask kernel.
constexpr int kLazyThreads = 1024;

struct LazyPub {           // kept boxes of one tile, compacted in rank order
  float4 box[64];
  float area[64];
  int cls[64];
  int count;
  int pad[3];
};

// `thr` is the largest float <= the caller's double threshold: for a float q, (double)q > thr_double <=> q > thr, so the
// comparison stays in fp32 with the result of torchvision's float-vs-double compare.
__device__ __forceinline__ bool lazy_iou_sup(const float4 a, const float aa, const float4 b, const float ab, const float thr,
                                             const bool neg_thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  if (inter > 0.0f || neg_thr) {  // inter == 0 => iou is 0, -0 or NaN: never > thr for thr >= 0
    const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
    // fl(inter / uni) > thr is decided without the division unless inter is within 2^-19 of thr * uni: the quotient's
    // rounding (2^-24) and that of the product cannot move a comparison that far.  Ordinary magnitudes only.
    const float pr = __fmul_rn(thr, uni);
    if (pr > 1e-30f && pr < 1e30f) {
      if (inter > __fmul_rn(pr, 1.000002f)) return true;
      if (inter < __fmul_rn(pr, 0.999998f)) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
  }
  return false;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// generic address of `p` (a shared-memory address of this CTA) in CTA `rank` of the cluster
template <typename T> __device__ __forceinline__ T *cluster_map(T *p, unsigned rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
  return reinterpret_cast<T *>(out);
}

__host__ __device__ inline size_t lazy_smem_bytes(int own_cap_tiles) {
  // boxes + areas + classes + intra-tile suppressor words, removed words (padded to 16 B), three kept lists
  return (size_t)own_cap_tiles * 64 * (16 + 4 + 4 + 8) + (((size_t)own_cap_tiles * 8 + 15) & ~(size_t)15) + 3 * sizeof(LazyPub);
}

template <bool kClassAware>
__global__ void __launch_bounds__(kLazyThreads) nms_lazy_kernel(const float4 *__restrict__ boxes, const int *__restrict__ cls,
                                                                const Seg *__restrict__ segs, float thr, int max_keep,
                                                                int keep_stride, int own_cap, int *__restrict__ keep_rank,
                                                                int *__restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char lazy_raw[];
  float4 *ob = reinterpret_cast<float4 *>(lazy_raw);
  unsigned long long *col = reinterpret_cast<unsigned long long *>(ob + (size_t)own_cap * 64);   // suppressors (lower bits) per box
  float *oa = reinterpret_cast<float *>(col + (size_t)own_cap * 64);
  int *oc = reinterpret_cast<int *>(oa + (size_t)own_cap * 64);
  unsigned long long *removed = reinterpret_cast<unsigned long long *>(oc + (size_t)own_cap * 64);
  LazyPub *loc = reinterpret_cast<LazyPub *>(removed + ((own_cap + 1) & ~1));                    // 16-byte aligned

  const int CL = (int)cluster_nctarank(), c = (int)cluster_ctarank();
  const int s = blockIdx.y;
  const Seg seg = segs[s];
  const int n = seg.len;
  const int W = (n + 63) >> 6;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool neg_thr = thr < 0.0f;
  const int own_tiles = W > c ? (W - c + CL - 1) / CL : 0;   // tiles t = q*CL + c, q < own_tiles (<= own_cap by construction)
  const int own_n = own_tiles * 64;
  int *out = keep_rank + (size_t)s * keep_stride;

  for (int idx = tid; idx < own_n; idx += kLazyThreads) {
    const int q = idx >> 6, r = (q * CL + c) * 64 + (idx & 63);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    int cc = 0;
    if (r < n) { b = boxes[seg.start + r]; if (kClassAware) cc = cls[seg.start + r]; }
    ob[idx] = b; oa[idx] = sfod_box_area(b); oc[idx] = cc;
  }
  for (int q = tid; q < own_tiles; q += kLazyThreads) {
    const int nvalid = min(64, n - (q * CL + c) * 64);
    removed[q] = nvalid >= 64 ? 0ull : ~((1ull << nvalid) - 1ull);   // lanes beyond the segment never participate
  }
  __syncthreads();
  // intra-tile suppressor words of every own tile: byte ic of col[q*64 + j] = { i in [8 ic, 8 ic + 8) : i < j, IoU(i, j) > thr }
  for (int idx = tid; idx < own_n * 8; idx += kLazyThreads) {
    const int q = idx >> 9, j = (idx >> 3) & 63, ic = idx & 7;
    const int nvalid = min(64, n - (q * CL + c) * 64);
    unsigned bits = 0;
    if (j < nvalid && ic * 8 < j) {
      const float4 bj = ob[q * 64 + j]; const float aj = oa[q * 64 + j];
      const int cj = oc[q * 64 + j];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = ic * 8 + ii;
        if (i < j && (!kClassAware || oc[q * 64 + i] == cj) && lazy_iou_sup(ob[q * 64 + i], oa[q * 64 + i], bj, aj, thr, neg_thr))
          bits |= 1u << ii;
      }
    }
    reinterpret_cast<unsigned char *>(col)[(size_t)(q * 64 + j) * 8 + ic] = (unsigned char)bits;
  }
  __syncthreads();

  // Resolve own tile q (global tile t): fixed-point iteration in warp 0, compacted list into loc[slot] of EVERY CTA.
  auto resolve_publish = [&](int q, int t, int slot, int base) {
    LazyPub &mine = loc[slot];
    if (warp == 0) {
      const unsigned long long alive = ~removed[q];
      const unsigned long long c0 = col[q * 64 + lane], c1 = col[q * 64 + 32 + lane];
      const bool a0 = (alive >> lane) & 1ull, a1 = (alive >> (32 + lane)) & 1ull;
      unsigned long long kept = alive;
      for (;;) {   // kept_j = alive_j && no kept suppressor below j; bits 0..k-1 are final after k rounds
        const unsigned lo = __ballot_sync(0xFFFFFFFFu, a0 && !(c0 & kept));
        const unsigned hi = __ballot_sync(0xFFFFFFFFu, a1 && !(c1 & kept));
        const unsigned long long nk = ((unsigned long long)hi << 32) | lo;
        if (nk == kept) break;
        kept = nk;
      }
      int nk = __popcll(kept);
      if (base + nk > max_keep) {   // keep only the first (max_keep - base) survivors
        const int allow = max_keep - base;
        unsigned long long kk = kept, keep2 = 0;
        for (int z = 0; z < allow; ++z) { const unsigned long long low = kk & (0ull - kk); keep2 |= low; kk ^= low; }
        kept = keep2; nk = allow;
      }
      const unsigned klo = (unsigned)kept, khi = (unsigned)(kept >> 32);
      if ((klo >> lane) & 1u) {
        const int pos = __popc(klo & ((1u << lane) - 1u));
        out[base + pos] = t * 64 + lane;
        mine.box[pos] = ob[q * 64 + lane]; mine.area[pos] = oa[q * 64 + lane]; mine.cls[pos] = oc[q * 64 + lane];
      }
      if ((khi >> lane) & 1u) {
        const int pos = __popc(klo) + __popc(khi & ((1u << lane) - 1u));
        out[base + pos] = t * 64 + 32 + lane;
        mine.box[pos] = ob[q * 64 + 32 + lane]; mine.area[pos] = oa[q * 64 + 32 + lane]; mine.cls[pos] = oc[q * 64 + 32 + lane];
      }
      if (lane == 0) mine.count = nk;
    }
    __syncthreads();
    const int kc = mine.count;
    for (int p = warp; p < CL; p += kLazyThreads / 32) {   // one warp per peer: push the list through DSMEM
      if (p == c) continue;
      LazyPub *remote = cluster_map(&mine, (unsigned)p);
      for (int k = lane; k < kc; k += 32) {
        remote->box[k] = mine.box[k]; remote->area[k] = mine.area[k];
        if (kClassAware) remote->cls[k] = mine.cls[k];
      }
      if (lane == 0) remote->count = kc;
    }
  };

  // test own boxes [b_lo, b_hi) that are still alive against the kc kept boxes of `L`: 8 lanes share one box and split
  // the list (short dependent chains, balanced work); a lane stops at its own first hit
  auto apply_range = [&](const LazyPub &L, int b_lo, int b_hi, int kc) {
    if (kc <= 0) return;
    const int sub = tid & 7;
    for (int b = b_lo + (tid >> 3); b < b_hi; b += kLazyThreads / 8) {
      const int q = b >> 6, bit = b & 63;
      if ((removed[q] >> bit) & 1ull) continue;
      const float4 bx = ob[b]; const float ab = oa[b];
      const int cb = oc[b];
      for (int k = sub; k < kc; k += 8) {
        if (kClassAware && L.cls[k] != cb) continue;
        if (lazy_iou_sup(L.box[k], L.area[k], bx, ab, thr, neg_thr)) { atomicOr(&removed[q], 1ull << bit); break; }
      }
    }
  };

  int total = 0;
  if (c == 0 && W > 0) resolve_publish(0, 0, 0, 0);
  cluster_arrive(); cluster_wait();
  for (int t = 0; t < W; ++t) {
    const int slot = t % 3;
    const LazyPub &L = loc[slot];
    const int kc = L.count;
    total += kc;
    const bool done = (total >= max_keep) || (t + 1 >= W);
    int q_first = t < c ? 0 : (t - c) / CL + 1;       // first own tile with global index > t
    if (!done && c == (t + 1) % CL) {                 // owner of the next tile: finish that tile first and publish it
      const int qn = (t + 1) / CL;                    // == q_first
      apply_range(L, qn * 64, qn * 64 + 64, kc);
      __syncthreads();
      resolve_publish(qn, t + 1, (t + 1) % 3, total);
      q_first = qn + 1;
    }
    cluster_arrive();
    if (!done) apply_range(L, q_first * 64, own_n, kc);
    cluster_wait();
    if (done) break;
  }
  if (c == 0 && tid == 0) keep_count[s] = total < max_keep ? total : max_keep;
  cluster_arrive(); cluster_wait();   // no CTA may exit while a peer can still push into its shared memory
}

// ---------------------------------------------------------------------------------------------------------------
// Deferred-apply cluster NMS (implementation A: the default whenever the kept list fits in shared memory).
// The lazy-row kernel above applies every kept list to ALL later boxes of the segment as soon as it is published.  When
// the scan stops early -- RPN keeps post_nms_topk = 2000 of 9 990 candidates and reaches them after ~2 700 boxes -- three
// quarters of those IoU tests are spent on boxes that are never visited (17 M of 20 M pairs per VGG image), and they sit
// between consecutive cluster steps.  Here every CTA keeps the COMPLETE kept list of the segment (boxes, areas, classes;
// owners push their tile's survivors into every CTA's copy) and a tile meets the kept boxes only when its turn comes
// near: each CTA works ahead on its NEXT own tile only, a chunk of the backlog per cluster step sized so that it is done
// exactly when the CTA becomes owner again.  The pair set evaluated for a visited box is the one of the lazy kernel
// (first hit ends the box); unvisited boxes cost nothing; the step itself is the critical path (apply the newest list to
// one tile, fixed-point resolve, push, split cluster barrier).  Same arithmetic, tie order and results.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// one arrival on the barrier at the same shared-memory offset of CTA `rank`; release at cluster scope: the remote stores
// this thread issued before it are visible to whoever observes the completed phase with acquire.cluster
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long *bar, unsigned rank) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

__host__ __device__ inline size_t deferred_smem_bytes(int own_cap_tiles, int cap) {
  // own boxes (16 B) + kept boxes (16 B) | suppressor words, removed words (8 B) | own areas, classes, kept areas, classes, counts
  return (size_t)own_cap_tiles * 64 * 16 + (size_t)cap * 16 + (size_t)own_cap_tiles * 64 * 8 + (((size_t)own_cap_tiles * 8 + 15) & ~(size_t)15) +
         (size_t)own_cap_tiles * 64 * 8 + (size_t)cap * 8;
}
// + one arrival barrier and one count per tile of the segment
__host__ __device__ inline size_t deferred_smem_bytes(int own_cap_tiles, int cap, int wcap) {
  return deferred_smem_bytes(own_cap_tiles, cap) + (size_t)wcap * 12 + 16;
}

template <bool kClassAware>
__global__ void __launch_bounds__(kLazyThreads) nms_deferred_kernel(const float4 *__restrict__ boxes, const int *__restrict__ cls,
                                                                    const Seg *__restrict__ segs, float thr, int max_keep,
                                                                    int keep_stride, int own_cap, int cap, int wcap,
                                                                    int *__restrict__ keep_rank, int *__restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char lazy_raw[];
  float4 *ob = reinterpret_cast<float4 *>(lazy_raw);
  float4 *KB = ob + (size_t)own_cap * 64;
  unsigned long long *col = reinterpret_cast<unsigned long long *>(KB + cap);
  unsigned long long *removed = col + (size_t)own_cap * 64;
  unsigned long long *mbar = removed + ((own_cap + 1) & ~1);   // mbar[t]: "the kept list of tile t has landed in this CTA"
  float *oa = reinterpret_cast<float *>(mbar + wcap);
  int *oc = reinterpret_cast<int *>(oa + (size_t)own_cap * 64);
  float *KA = reinterpret_cast<float *>(oc + (size_t)own_cap * 64);
  int *KC = reinterpret_cast<int *>(KA + cap);
  int *cnt = KC + cap;   // cnt[t]: survivors of tile t

  const int CL = (int)cluster_nctarank(), c = (int)cluster_ctarank();
  const int s = blockIdx.y;
  const Seg seg = segs[s];
  const int n = seg.len;
  const int W = (n + 63) >> 6;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool neg_thr = thr < 0.0f;
  const int own_tiles = W > c ? (W - c + CL - 1) / CL : 0;
  const int own_n = own_tiles * 64;
  int *out = keep_rank + (size_t)s * keep_stride;

  for (int idx = tid; idx < own_n; idx += kLazyThreads) {
    const int q = idx >> 6, r = (q * CL + c) * 64 + (idx & 63);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    int cc = 0;
    if (r < n) { b = boxes[seg.start + r]; if (kClassAware) cc = cls[seg.start + r]; }
    ob[idx] = b; oa[idx] = sfod_box_area(b); oc[idx] = cc;
  }
  for (int q = tid; q < own_tiles; q += kLazyThreads) {
    const int nvalid = min(64, n - (q * CL + c) * 64);
    removed[q] = nvalid >= 64 ? 0ull : ~((1ull << nvalid) - 1ull);
  }
  __syncthreads();
  for (int idx = tid; idx < own_n * 8; idx += kLazyThreads) {   // intra-tile suppressor words (see the lazy kernel)
    const int q = idx >> 9, j = (idx >> 3) & 63, ic = idx & 7;
    const int nvalid = min(64, n - (q * CL + c) * 64);
    unsigned bits = 0;
    if (j < nvalid && ic * 8 < j) {
      const float4 bj = ob[q * 64 + j]; const float aj = oa[q * 64 + j];
      const int cj = oc[q * 64 + j];
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int i = ic * 8 + ii;
        if (i < j && (!kClassAware || oc[q * 64 + i] == cj) && lazy_iou_sup(ob[q * 64 + i], oa[q * 64 + i], bj, aj, thr, neg_thr))
          bits |= 1u << ii;
      }
    }
    reinterpret_cast<unsigned char *>(col)[(size_t)(q * 64 + j) * 8 + ic] = (unsigned char)bits;
  }
  __syncthreads();

  // Resolve own tile q (global tile t) and append its survivors to the kept list of EVERY CTA at [base, base + nk).
  auto resolve_publish = [&](int q, int t, int base) {
    if (warp == 0) {
      const unsigned long long alive = ~removed[q];
      const unsigned long long c0 = col[q * 64 + lane], c1 = col[q * 64 + 32 + lane];
      const bool a0 = (alive >> lane) & 1ull, a1 = (alive >> (32 + lane)) & 1ull;
      unsigned long long kept = alive;
      for (;;) {
        const unsigned lo = __ballot_sync(0xFFFFFFFFu, a0 && !(c0 & kept));
        const unsigned hi = __ballot_sync(0xFFFFFFFFu, a1 && !(c1 & kept));
        const unsigned long long nk = ((unsigned long long)hi << 32) | lo;
        if (nk == kept) break;
        kept = nk;
      }
      int nk = __popcll(kept);
      if (base + nk > max_keep) {
        const int allow = max_keep - base;
        unsigned long long kk = kept, keep2 = 0;
        for (int z = 0; z < allow; ++z) { const unsigned long long low = kk & (0ull - kk); keep2 |= low; kk ^= low; }
        kept = keep2; nk = allow;
      }
      const unsigned klo = (unsigned)kept, khi = (unsigned)(kept >> 32);
      if ((klo >> lane) & 1u) {
        const int pos = base + __popc(klo & ((1u << lane) - 1u));
        out[pos] = t * 64 + lane;
        KB[pos] = ob[q * 64 + lane]; KA[pos] = oa[q * 64 + lane]; KC[pos] = oc[q * 64 + lane];
      }
      if ((khi >> lane) & 1u) {
        const int pos = base + __popc(klo) + __popc(khi & ((1u << lane) - 1u));
        out[pos] = t * 64 + 32 + lane;
        KB[pos] = ob[q * 64 + 32 + lane]; KA[pos] = oa[q * 64 + 32 + lane]; KC[pos] = oc[q * 64 + 32 + lane];
      }
      if (lane == 0) cnt[t] = nk;
    }
    __syncthreads();
    const int kc = cnt[t];
    for (int p = warp; p < CL; p += kLazyThreads / 32) {   // one warp per CTA of the cluster: push through DSMEM, then signal
      if (p != c) {
        float4 *rKB = cluster_map(KB, (unsigned)p);
        float *rKA = cluster_map(KA, (unsigned)p);
        int *rKC = cluster_map(KC, (unsigned)p);
        for (int k = lane; k < kc; k += 32) {
          rKB[base + k] = KB[base + k]; rKA[base + k] = KA[base + k];
          if (kClassAware) rKC[base + k] = KC[base + k];
        }
        if (lane == 0) cluster_map(cnt, (unsigned)p)[t] = kc;
      }
      mbar_arrive_remote(&mbar[t], (unsigned)p);   // all 32 lanes: each releases its own stores (barrier count 32)
    }
  };

  // own tile q against kept entries [k_lo, k_hi): 16 lanes share one box and stride through the list
  auto apply = [&](int q, int k_lo, int k_hi) {
    if (k_hi <= k_lo) return;
    const int bi = tid >> 4, sub = tid & 15;
    const unsigned long long bit = 1ull << bi;
    if (removed[q] & bit) return;
    const float4 bx = ob[q * 64 + bi]; const float ab = oa[q * 64 + bi];
    const int cb = oc[q * 64 + bi];
    for (int k = k_lo + sub, it = 0; k < k_hi; k += 16, ++it) {
      if ((it & 3) == 3 && (removed[q] & bit)) break;   // another lane already removed this box
      if (kClassAware && KC[k] != cb) continue;
      if (lazy_iou_sup(KB[k], KA[k], bx, ab, thr, neg_thr)) { atomicOr(&removed[q], bit); break; }
    }
  };

  for (int t = tid; t < W; t += kLazyThreads) mbar_init(&mbar[t], 32u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  cluster_arrive(); cluster_wait();   // every CTA's barriers exist before anybody signals them

  // No cluster-wide barrier per step: the owner of tile t pushes its list and signals mbar[t] in every CTA (one-way
  // DSMEM latency); nothing is ever reused (lists append, one barrier and one count per tile), so no flow control either.
  int total = 0;
  int next_q = 0;   // next own tile to resolve (global tile next_q * CL + c)
  int done = 0;     // kept entries [0, done) have been applied to it
  if (c == 0 && W > 0) { resolve_publish(0, 0, 0); next_q = 1; }
  const int cl_mask = CL - 1, cl_shift = __ffs(CL) - 1;   // CL is a power of two
  for (int t = 0; t < W; ++t) {
    if (warp == 0) mbar_wait(&mbar[t], 0u);   // one warp polls, the others sleep at the barrier
    __syncthreads();
    const int new_total = total + cnt[t];
    const bool finished = (new_total >= max_keep) || (t + 1 >= W);
    if (!finished && c == ((t + 1) & cl_mask)) {   // owner of the next tile: finish its backlog + the newest list, resolve, publish
      apply(next_q, done, new_total);
      __syncthreads();
      resolve_publish(next_q, t + 1, new_total);
      ++next_q; done = 0;
    }
    if (!finished && next_q < own_tiles) {   // work ahead on the next own tile: a share of the backlog per step
      const int turn = (next_q << cl_shift) + c - 1;   // loop iteration at which this CTA resolves it
      const int steps_left = max(1, turn - t - 1);   // be done one step early: the step before the turn only sees the newest list
      const int avail = new_total - done;    // the newest list is complete in this CTA's copy (barrier wait above)
      int chunk = (int)(__fdividef((float)avail, (float)steps_left)) + 1;   // ~ceil(avail / steps_left): a pacing heuristic only
      chunk = min(avail, (chunk + 15) & ~15);
      apply(next_q, done, done + chunk);
      done += chunk;
    }
    total = new_total;
    if (finished) break;
  }
  if (c == 0 && tid == 0) keep_count[s] = total < max_keep ? total : max_keep;
  cluster_arrive(); cluster_wait();   // no CTA may exit while a peer can still push into its shared memory
}

// host helpers
static inline int lazy_cluster_size(int S, int max_len, size_t *smem_out, int *own_cap_out) {
  int CL = 16;
  while (CL > 1 && (long long)S * CL > 144) CL >>= 1;          // keep every cluster resident at once when possible
  const int W = (max_len + 63) / 64;
  while (CL > 1 && CL > W) CL >>= 1;
  int own_cap = (W + CL - 1) / CL; if (own_cap < 1) own_cap = 1;
  *smem_out = lazy_smem_bytes(own_cap); *own_cap_out = own_cap;
  return CL;
}

// Largest cluster size the device will actually schedule for this kernel (probed once per process and kernel variant; an
// idempotent capability cache, not state that any result depends on).  0 = the kernel cannot be launched as a cluster.
template <bool kCls>
static inline int lazy_max_cluster() {
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = nms_lazy_kernel<kCls>;
  (void)cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int best = 0;
  for (int CL = 16; CL >= 1; CL >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 1, 1); cfg.blockDim = dim3(kLazyThreads, 1, 1); cfg.dynamicSmemBytes = lazy_smem_bytes(16);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0) { best = CL; break; }
    (void)cudaGetLastError();
  }
  cached = best;
  return best;
}

template <bool kCls>
static inline int launch_lazy_t(const float4 *boxes, const int *cls, const Seg *segs, int S, int max_len, double thr, int max_keep,
                                int keep_stride, int *keep_rank, int *keep_count, cudaStream_t stream, bool *launched) {
  *launched = false;
  const int max_cl = lazy_max_cluster<kCls>();
  if (max_cl <= 0) return SFOD_OK;         // caller falls back to mask + scan
  size_t smem; int own_cap;
  int CL = lazy_cluster_size(S, max_len, &smem, &own_cap);
  if (CL > max_cl) {
    CL = max_cl;
    const int W = (max_len + 63) / 64;
    own_cap = (W + CL - 1) / CL; if (own_cap < 1) own_cap = 1;
    smem = lazy_smem_bytes(own_cap);
  }
  if (smem > 200 * 1024) return SFOD_OK;   // segment too long for the shared-memory resident form: caller falls back
  auto kern = nms_lazy_kernel<kCls>;
  if (smem > 48 * 1024) SFOD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, S, 1); cfg.blockDim = dim3(kLazyThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  float thr_f = (float)thr;                                  // largest float <= thr (see lazy_iou_sup)
  if ((double)thr_f > thr) thr_f = nextafterf(thr_f, -INFINITY);
  if (cudaLaunchKernelEx(&cfg, kern, boxes, cls, segs, thr_f, max_keep, keep_stride, own_cap, keep_rank, keep_count) != cudaSuccess) {
    (void)cudaGetLastError();
    return SFOD_OK;                        // not launchable in this configuration: caller falls back
  }
  sfod_count_launch();
  *launched = true;
  return SFOD_OK;
}

template <bool kCls>
static inline int deferred_max_cluster() {   // see lazy_max_cluster
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = nms_deferred_kernel<kCls>;
  (void)cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  (void)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int best = 0;
  for (int CL = 16; CL >= 1; CL >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL, 1, 1); cfg.blockDim = dim3(kLazyThreads, 1, 1); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) == cudaSuccess && max_clusters > 0) { best = CL; break; }
    (void)cudaGetLastError();
  }
  cached = best;
  return best;
}

template <bool kCls>
static inline int launch_deferred_t(const float4 *boxes, const int *cls, const Seg *segs, int S, int max_len, double thr, int max_keep,
                                    int keep_stride, int *keep_rank, int *keep_count, cudaStream_t stream, bool *launched) {
  *launched = false;
  const int max_cl = deferred_max_cluster<kCls>();
  if (max_cl <= 0) return SFOD_OK;
  size_t unused; int own_cap;
  int CL = lazy_cluster_size(S, max_len, &unused, &own_cap);
  if (CL > max_cl) {
    CL = max_cl;
    const int W = (max_len + 63) / 64;
    own_cap = (W + CL - 1) / CL; if (own_cap < 1) own_cap = 1;
  }
  int cap = max_keep < max_len ? max_keep : max_len;   // the kept list never grows beyond min(max_keep, segment length)
  cap = (cap + 3) & ~3; if (cap < 4) cap = 4;
  const int wcap = (max_len + 63) / 64;
  const size_t smem = deferred_smem_bytes(own_cap, cap, wcap);
  if (smem > 200 * 1024) return SFOD_OK;   // kept list too long for shared memory: caller uses the lazy-row kernel
  auto kern = nms_deferred_kernel<kCls>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL, S, 1); cfg.blockDim = dim3(kLazyThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  float thr_f = (float)thr;                                  // largest float <= thr (see lazy_iou_sup)
  if ((double)thr_f > thr) thr_f = nextafterf(thr_f, -INFINITY);
  if (cudaLaunchKernelEx(&cfg, kern, boxes, cls, segs, thr_f, max_keep, keep_stride, own_cap, cap, wcap, keep_rank, keep_count) != cudaSuccess) {
    (void)cudaGetLastError();
    return SFOD_OK;
  }
  sfod_count_launch();
  *launched = true;
  return SFOD_OK;
}

static inline int launch_mask(const float4 *boxes, const int *cls, const Seg *segs, int S, int max_len, int rows_per_slab,
                              int wstride, double thr, unsigned long long *mask, cudaStream_t stream) {
  if (S <= 0 || max_len <= 0) return SFOD_OK;
  const int tiles = (max_len + 63) / 64;
  dim3 grid(tiles, tiles, S);
  if (cls) nms_mask_kernel<true><<<grid, 64, 0, stream>>>(boxes, cls, segs, rows_per_slab, wstride, thr, mask);
  else nms_mask_kernel<false><<<grid, 64, 0, stream>>>(boxes, nullptr, segs, rows_per_slab, wstride, thr, mask);
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

// Segmented NMS entry used by every caller: deferred-apply cluster kernel; lazy-row cluster kernel when the kept list does
// not fit in shared memory; mask + scan when clusters cannot run.
static inline int run_segmented(const float4 *boxes, const int *cls, const Seg *segs, int S, int max_len, int rows_per_slab,
                                int wstride, double thr, int max_keep, int keep_stride, unsigned long long *mask,
                                int *keep_rank, int *keep_count, cudaStream_t stream) {
  if (S <= 0) return SFOD_OK;
  if (max_len > 0) {
    bool launched = false;
    int rc = cls ? launch_deferred_t<true>(boxes, cls, segs, S, max_len, thr, max_keep, keep_stride, keep_rank, keep_count, stream, &launched)
                 : launch_deferred_t<false>(boxes, cls, segs, S, max_len, thr, max_keep, keep_stride, keep_rank, keep_count, stream, &launched);
    if (rc) return rc;
    if (launched) return SFOD_OK;
    rc = cls ? launch_lazy_t<true>(boxes, cls, segs, S, max_len, thr, max_keep, keep_stride, keep_rank, keep_count, stream, &launched)
                 : launch_lazy_t<false>(boxes, cls, segs, S, max_len, thr, max_keep, keep_stride, keep_rank, keep_count, stream, &launched);
    if (rc) return rc;
    if (launched) return SFOD_OK;
  }
  int rc = launch_mask(boxes, cls, segs, S, max_len, rows_per_slab, wstride, thr, mask, stream);
  if (rc) return rc;
  return launch_scan(mask, segs, S, rows_per_slab, wstride, max_keep, keep_stride, keep_rank, keep_count, stream);
}

}  // namespace nmsk
