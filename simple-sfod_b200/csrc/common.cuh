// common.cuh -- shared device/host helpers for libsfod_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "../../include/sfod_b200.h"

#define SFOD_API extern "C" __attribute__((visibility("default")))

#define SFOD_CUDA_TRY(expr)                                        \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return SFOD_ERR_CUDA_BASE + (int)_e;    \
  } while (0)

// Diagnostic launch counter (the only process-wide state of the library; monotonically increasing, never read by
// any kernel path): every kernel launch of the library passes through SFOD_LAUNCH_CHECK exactly once.
void sfod_count_launch();

#define SFOD_LAUNCH_CHECK()                                        \
  do {                                                             \
    sfod_count_launch();                                           \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return SFOD_ERR_CUDA_BASE + (int)_e;    \
  } while (0)

static inline cudaStream_t sfod_cu(sfod_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool sfod_aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t sfod_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int SFOD_NUM_SMS = 148;  // B200

// Bump allocator over the caller-provided workspace (256-byte aligned slices).
struct SfodWs {
  char *base; size_t cap; size_t off; bool ok;
  __host__ SfodWs(void *p, size_t c) : base(static_cast<char *>(p)), cap(c), off(0), ok(true) {}
  template <typename T> __host__ T *take(size_t n) {
    size_t bytes = sfod_align_up(n * sizeof(T), 256);
    if (base == nullptr || off + bytes > cap) { ok = false; off += bytes; return nullptr; }
    T *r = reinterpret_cast<T *>(base + off); off += bytes; return r;
  }
};
// Same arithmetic, sizes only.
struct SfodWsSize {
  size_t off = 0;
  template <typename T> T *take(size_t n) { off += sfod_align_up(n * sizeof(T), 256); return nullptr; }
};

// ---- order-preserving key for fp32 scores: larger score -> larger key; NaN is the largest
// (torch.sort semantics); -0.0 == +0.0.
__host__ __device__ __forceinline__ uint32_t sfod_score_key(float s) {
  if (s != s) return 0xFFFFFFFFu;
  if (s == 0.0f) s = 0.0f;  // canonicalise -0
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(s);
#else
  union { float f; uint32_t u; } cv; cv.f = s; u = cv.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float sfod_key_score(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } cv; cv.u = u; return cv.f;
#endif
}

#ifdef __CUDACC__
// Correctly rounded fp32 exponential (via fp64), the arithmetic the oracle defines.
__device__ __forceinline__ float sfod_exp_cr(float x) { return (float)exp((double)x); }

// d2 Box2BoxTransform.apply_deltas for one (box, delta) pair; separately rounded fp32 ops in
// the order of the Python source (this TU is compiled with -fmad=false).
__device__ __forceinline__ float4 sfod_decode_box(float4 b, float4 d, float wx, float wy, float ww, float wh,
                                                  float scale_clamp) {
  float width = __fsub_rn(b.z, b.x), height = __fsub_rn(b.w, b.y);
  float cx = __fadd_rn(b.x, __fmul_rn(0.5f, width));
  float cy = __fadd_rn(b.y, __fmul_rn(0.5f, height));
  float dx = __fdiv_rn(d.x, wx), dy = __fdiv_rn(d.y, wy), dw = __fdiv_rn(d.z, ww), dh = __fdiv_rn(d.w, wh);
  if (dw > scale_clamp) dw = scale_clamp;  // torch.clamp(max=): NaN propagates
  if (dh > scale_clamp) dh = scale_clamp;
  float px = __fadd_rn(__fmul_rn(dx, width), cx);
  float py = __fadd_rn(__fmul_rn(dy, height), cy);
  float pw = __fmul_rn(sfod_exp_cr(dw), width);
  float ph = __fmul_rn(sfod_exp_cr(dh), height);
  float hpw = __fmul_rn(0.5f, pw), hph = __fmul_rn(0.5f, ph);
  return make_float4(__fsub_rn(px, hpw), __fsub_rn(py, hph), __fadd_rn(px, hpw), __fadd_rn(py, hph));
}

// d2 Boxes.clip: x in [0,w], y in [0,h] (torch.clamp semantics: NaN stays NaN).
__device__ __forceinline__ float sfod_clampf(float v, float lo, float hi) {
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}
__device__ __forceinline__ float4 sfod_clip_box(float4 b, float h, float w) {
  return make_float4(sfod_clampf(b.x, 0.f, w), sfod_clampf(b.y, 0.f, h), sfod_clampf(b.z, 0.f, w), sfod_clampf(b.w, 0.f, h));
}
__device__ __forceinline__ bool sfod_finite4(float4 b) { return isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w); }

__device__ __forceinline__ float sfod_box_area(float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }
#endif
