// ema.cu -- mean-teacher EMA as ONE multi-tensor launch.
// Replaces the per-tensor Python loop of reference
//   daod/engine/trainers/source_free_adaptive_teacher.py:593-603 (3 temporaries + 1 copy per tensor,
//   ~36 B/element of traffic) by a single pass at the algorithmic 12 B/element:
//   read student, read teacher, write teacher.
// HBM-bound elementwise kernel: persistent grid (multiple of 148 SMs), 128-bit loads/stores,
// 4 independent 16-byte requests per operand in flight per thread, streaming cache hints
// (each byte is touched exactly once per step; nothing is worth keeping in L1).
// Arithmetic contract (bit-exact vs the reference): t = fl(fl(s*a) + fl(t*b)), a = fp32(1-k),
// b = fp32(k); compiled with -fmad=false and written with explicit _rn intrinsics so no FMA is formed.
#include "common.cuh"

namespace {

constexpr int kChunkElems = 16384;  // elements per plan chunk (64 KiB of fp32 per operand)
constexpr int kThreads = 256;

struct alignas(8) EmaChunk {
  const void *student;
  void *teacher;
  int32_t n;
  int32_t dtype;
};
static_assert(sizeof(EmaChunk) == 24, "plan layout");

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_rw(const float4 *p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float ema1(float s, float t, float a, float b) {
  return __fadd_rn(__fmul_rn(s, a), __fmul_rn(t, b));
}
__device__ __forceinline__ float4 ema4(float4 s, float4 t, float a, float b) {
  return make_float4(ema1(s.x, t.x, a, b), ema1(s.y, t.y, a, b), ema1(s.z, t.z, a, b), ema1(s.w, t.w, a, b));
}

__global__ void __launch_bounds__(kThreads) ema_multi_tensor_kernel(const EmaChunk *__restrict__ plan, int64_t n_chunks,
                                                                    float a, float b) {
  for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const EmaChunk ch = plan[c];
    const int n = ch.n;
    if (ch.dtype == SFOD_F32) {
      const float *s = static_cast<const float *>(ch.student);
      float *t = static_cast<float *>(ch.teacher);
      const bool vec = (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(t)) & 15u) == 0);
      if (vec) {
        const float4 *s4 = reinterpret_cast<const float4 *>(s);
        float4 *t4 = reinterpret_cast<float4 *>(t);
        const int n4 = n >> 2;
        // 4 x 16 B per operand per thread per iteration (kThreads*4 float4 = 4096 elements)
        int i = threadIdx.x;
        for (; i + 3 * kThreads < n4; i += 4 * kThreads) {
          float4 sv0 = ld_stream(s4 + i), sv1 = ld_stream(s4 + i + kThreads);
          float4 sv2 = ld_stream(s4 + i + 2 * kThreads), sv3 = ld_stream(s4 + i + 3 * kThreads);
          float4 tv0 = ld_rw(t4 + i), tv1 = ld_rw(t4 + i + kThreads);
          float4 tv2 = ld_rw(t4 + i + 2 * kThreads), tv3 = ld_rw(t4 + i + 3 * kThreads);
          st_stream(t4 + i, ema4(sv0, tv0, a, b));
          st_stream(t4 + i + kThreads, ema4(sv1, tv1, a, b));
          st_stream(t4 + i + 2 * kThreads, ema4(sv2, tv2, a, b));
          st_stream(t4 + i + 3 * kThreads, ema4(sv3, tv3, a, b));
        }
        for (; i < n4; i += kThreads) st_stream(t4 + i, ema4(ld_stream(s4 + i), ld_rw(t4 + i), a, b));
        for (int j = (n4 << 2) + threadIdx.x; j < n; j += kThreads) t[j] = ema1(s[j], t[j], a, b);
      } else {
        for (int j = threadIdx.x; j < n; j += kThreads) t[j] = ema1(s[j], t[j], a, b);
      }
    } else {  // SFOD_I64: promote to fp32, blend, truncate (load_state_dict copy_ into the int64 buffer)
      const long long *s = static_cast<const long long *>(ch.student);
      long long *t = static_cast<long long *>(ch.teacher);
      for (int j = threadIdx.x; j < n; j += kThreads) {
        float r = ema1(__ll2float_rn(s[j]), __ll2float_rn(t[j]), a, b);
        t[j] = (long long)r;  // C-style truncation toward zero, as ATen's float->int64 copy
      }
    }
  }
}

}  // namespace

SFOD_API int64_t sfod_ema_plan_chunks(const sfod_ema_tensor *tensors, int n_tensors) {
  if (!tensors || n_tensors < 0) return -1;
  int64_t n = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (tensors[i].numel < 0) return -1;
    n += (tensors[i].numel + kChunkElems - 1) / kChunkElems;
  }
  return n;
}

SFOD_API size_t sfod_ema_plan_bytes(int64_t n_chunks) { return n_chunks > 0 ? (size_t)n_chunks * sizeof(EmaChunk) : 0; }

SFOD_API int sfod_ema_plan_build(const sfod_ema_tensor *tensors, int n_tensors, void *host_plan, size_t host_plan_bytes) {
  int64_t need = sfod_ema_plan_chunks(tensors, n_tensors);
  if (need < 0 || (need > 0 && !host_plan)) return SFOD_ERR_INVALID_ARG;
  if (host_plan_bytes < sfod_ema_plan_bytes(need)) return SFOD_ERR_WORKSPACE_TOO_SMALL;
  EmaChunk *out = static_cast<EmaChunk *>(host_plan);
  int64_t k = 0;
  for (int i = 0; i < n_tensors; ++i) {
    const sfod_ema_tensor &t = tensors[i];
    if (t.dtype != SFOD_F32 && t.dtype != SFOD_I64) return SFOD_ERR_UNSUPPORTED;
    if (t.numel > 0 && (!t.student || !t.teacher)) return SFOD_ERR_INVALID_ARG;
    const size_t esz = t.dtype == SFOD_F32 ? 4 : 8;
    for (int64_t off = 0; off < t.numel; off += kChunkElems) {
      int64_t n = t.numel - off < kChunkElems ? t.numel - off : kChunkElems;
      out[k].student = static_cast<const char *>(t.student) + off * esz;
      out[k].teacher = static_cast<char *>(t.teacher) + off * esz;
      out[k].n = (int32_t)n;
      out[k].dtype = t.dtype;
      ++k;
    }
  }
  return SFOD_OK;
}

SFOD_API int sfod_ema_multi_tensor(const void *device_plan, int64_t n_chunks, double keep_rate, sfod_stream_t stream) {
  if (n_chunks < 0) return SFOD_ERR_INVALID_ARG;
  if (n_chunks == 0) return SFOD_OK;
  if (!device_plan) return SFOD_ERR_INVALID_ARG;
  // Python evaluates (1 - keep_rate) in double; ATen then rounds the scalar to the tensor dtype.
  const float a = (float)(1.0 - keep_rate), b = (float)keep_rate;
  int64_t grid = n_chunks < (int64_t)SFOD_NUM_SMS * 8 ? n_chunks : (int64_t)SFOD_NUM_SMS * 8;
  ema_multi_tensor_kernel<<<(unsigned)grid, kThreads, 0, sfod_cu(stream)>>>(static_cast<const EmaChunk *>(device_plan),
                                                                            n_chunks, a, b);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}
