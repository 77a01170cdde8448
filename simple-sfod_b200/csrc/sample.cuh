// sample.cuh -- batched detectron2 `subsample_labels` (student-side proposal / anchor sampling).
//
// Replaces, for ALL images of a batch in one launch, the per-image
//     positive = nonzero((labels != -1) & (labels != bg_label));  negative = nonzero(labels == bg_label)
//     num_pos = min(positive.numel(), int(num_samples * positive_fraction));  num_neg = min(negative.numel(), num_samples - num_pos)
//     pos_idx = positive[randperm(positive.numel())[:num_pos]];  neg_idx = negative[randperm(negative.numel())[:num_neg]]
// of detectron2.modeling.sampling.subsample_labels, called per image from ROIHeads._sample_proposals (reference
// daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:177-186) and RPN._subsample_labels (via
// label_and_sample_anchors, reference daod/modeling/proposal_generator/rpn.py:45) -- each call with two `nonzero` host syncs.
//
// A uniformly random k-subset in random order = the k elements with the smallest i.i.d. random keys, ordered by key.  The key
// of element i of segment s is a counter-based hash (splitmix64 finaliser of seed, s, i) -- no RNG state, any element's key can be
// recomputed anywhere (the oracle does, on the CPU).  Per segment one CTA: class counts, an 8-bit radix select of the k-th
// smallest 32-bit hash per class (labels are re-read, keys recomputed: nothing is materialised), compaction of the selected
// elements into shared memory, a bitonic sort of those <= 1024 (hash, index) pairs.  Ties in the hash are broken by index.
#pragma once
#include "common.cuh"

namespace samplek {

constexpr int kThreads = 1024;
constexpr int kMaxSamples = 1024;
constexpr int kMaxSegments = 255;     // segment offsets travel in the kernel parameters (host array in, no device copy, no sync)
struct Offsets { int v[kMaxSegments + 1]; };

__host__ __device__ __forceinline__ unsigned int sample_hash(unsigned long long seed, unsigned int seg, unsigned int idx) {
  unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ((((unsigned long long)seg) << 32) | idx);
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (unsigned int)(x >> 32);
}

// class of a label: 0 = positive (foreground), 1 = negative (== bg_label), 2 = ignored (-1)
__device__ __forceinline__ int label_class(long long l, long long bg) { return l == bg ? 1 : (l == -1 ? 2 : 0); }

__global__ void __launch_bounds__(kThreads) subsample_kernel(const long long *__restrict__ labels, const Offsets offsets,
                                                             int num_samples, int max_pos, long long bg_label,
                                                             unsigned long long seed, long long *__restrict__ sampled,
                                                             int *__restrict__ counts) {
  __shared__ unsigned long long sel[kMaxSamples];      // (hash << 32) | index of the selected elements, fg then bg
  __shared__ unsigned int hist[256];
  __shared__ int s_cnt[2];                             // elements per class
  __shared__ unsigned int s_prefix, s_need;            // radix-select state
  __shared__ int s_fill;                               // compaction cursor
  __shared__ unsigned int eq_idx[64];                  // indices whose hash equals the threshold hash (ties; practically <= 1)
  __shared__ int s_eq;
  const int seg = blockIdx.x, tid = threadIdx.x;
  const int beg = offsets.v[seg], n = offsets.v[seg + 1] - beg;
  const long long *lab = labels + beg;
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  {
    int c0 = 0, c1 = 0;
    for (int i = tid; i < n; i += kThreads) { const int c = label_class(lab[i], bg_label); c0 += c == 0; c1 += c == 1; }
    c0 = __reduce_add_sync(0xFFFFFFFFu, c0); c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
    if ((tid & 31) == 0) { if (c0) atomicAdd(&s_cnt[0], c0); if (c1) atomicAdd(&s_cnt[1], c1); }
  }
  __syncthreads();
  const int num_pos = min(s_cnt[0], max_pos);
  const int num_neg = min(s_cnt[1], num_samples - num_pos);
  int base = 0;
  for (int cls = 0; cls < 2; ++cls) {
    const int k = cls == 0 ? num_pos : num_neg;         // CTA-uniform
    const int total = s_cnt[cls];
    if (k > 0) {
      // ---- threshold hash h*: the k-th smallest hash of the class (radix select, 4 x 8 bits); `need` = rank still to find
      unsigned int hstar = 0xFFFFFFFFu;
      if (k < total) {
        if (tid == 0) { s_prefix = 0u; s_need = (unsigned)k; }
        for (int shift = 24; shift >= 0; shift -= 8) {
          if (tid < 256) hist[tid] = 0u;
          __syncthreads();
          const unsigned int prefix = s_prefix;
          const unsigned int pmask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
          for (int i = tid; i < n; i += kThreads) {
            if (label_class(lab[i], bg_label) != cls) continue;
            const unsigned int h = sample_hash(seed, (unsigned)seg, (unsigned)i);
            if ((h & pmask) == prefix) atomicAdd(&hist[(h >> shift) & 255u], 1u);
          }
          __syncthreads();
          if (tid == 0) {
            unsigned int need = s_need, b = 0;
            while (b < 255u && hist[b] < need) { need -= hist[b]; ++b; }
            s_prefix = prefix | (b << shift); s_need = need;
          }
          __syncthreads();
        }
        hstar = s_prefix;
      }
      // ---- compaction: hash < h* always; hash == h* by ascending index until k are selected
      if (tid == 0) { s_fill = 0; s_eq = 0; }
      __syncthreads();
      for (int i = tid; i < n; i += kThreads) {
        if (label_class(lab[i], bg_label) != cls) continue;
        const unsigned int h = sample_hash(seed, (unsigned)seg, (unsigned)i);
        if (k == total || h < hstar) {
          const int p = atomicAdd(&s_fill, 1);
          sel[base + p] = ((unsigned long long)h << 32) | (unsigned)i;
        } else if (h == hstar) {
          const int p = atomicAdd(&s_eq, 1);
          if (p < 64) eq_idx[p] = (unsigned)i;
        }
      }
      __syncthreads();
      if (tid == 0 && k < total) {
        int ne = min(s_eq, 64), f = s_fill;
        for (int a = 1; a < ne; ++a) {   // insertion sort of the (tiny) tie list by index
          const unsigned int v = eq_idx[a]; int b = a - 1;
          while (b >= 0 && eq_idx[b] > v) { eq_idx[b + 1] = eq_idx[b]; --b; }
          eq_idx[b + 1] = v;
        }
        for (int a = 0; a < ne && f < k; ++a) sel[base + f++] = ((unsigned long long)hstar << 32) | eq_idx[a];
        s_fill = f;
      }
      __syncthreads();
      // ---- order by (hash, index): bitonic sort of the k selected pairs (padded with the maximal key)
      int P = 1; while (P < k) P <<= 1;
      for (int i = k + tid; i < P; i += kThreads) sel[base + i] = ~0ull;   // base + P <= kMaxSamples is guaranteed by the host
      __syncthreads();
      for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = tid; t < (P >> 1); t += kThreads) {
            const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
            const bool up = (lo & size) == 0;
            const unsigned long long a = sel[base + lo], b = sel[base + hi];
            if ((a > b) == up) { sel[base + lo] = b; sel[base + hi] = a; }
          }
          __syncthreads();
        }
      for (int i = tid; i < k; i += kThreads) sampled[(size_t)seg * num_samples + base + i] = (long long)(sel[base + i] & 0xFFFFFFFFull);
      __syncthreads();
    }
    base += k;
  }
  for (int i = base + tid; i < num_samples; i += kThreads) sampled[(size_t)seg * num_samples + i] = -1;
  if (tid == 0) { counts[2 * seg] = num_pos; counts[2 * seg + 1] = num_neg; }
}

}  // namespace samplek
