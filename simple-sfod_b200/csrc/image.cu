// image.cu -- detectron2 GeneralizedRCNN.preprocess_image as one pass (the normalise / pad / batch part).
// Replaces `(x - pixel_mean) / pixel_std` followed by ImageList.from_tensors (reference daod/modeling/meta_arch/rcnn.py:92-104):
// reads the uint8 (or float32) CHW image once (1 or 4 B/element) and writes the normalised, zero-padded fp32 batch slot once
// (4 B/element) in NCHW or channels-last order, instead of a uint8->float promotion pass, a subtraction pass, a division pass
// and a padding copy.  Arithmetic: fl(fl(x - mean[c]) / std[c]) in fp32, i.e. exactly ATen's; pad value 0.0 (applied after
// normalisation, as from_tensors does).
#include "common.cuh"

namespace {

constexpr int kMaxChannels = 8;
struct ChanParams { float mean[kMaxChannels]; float stdv[kMaxChannels]; };

__device__ __forceinline__ float norm1(float v, float m, float s) { return __fdiv_rn(__fsub_rn(v, m), s); }

// grid (x blocks, Hp, N * C) for NCHW / (x blocks, Hp, N) for channels-last: only the column comes from the thread index
template <typename T, bool kVec>
__global__ void __launch_bounds__(256) normalize_pad_nchw_kernel(const T *__restrict__ img, long long img_stride, int C, int H, int W,
                                                                 int Hp, int Wp, ChanParams p, float *__restrict__ out) {
  const int y = blockIdx.y, nc = blockIdx.z, n = nc / C, c = nc - n * C;
  const float m = p.mean[c], sd = p.stdv[c];
  const T *row = img + (size_t)n * img_stride + ((size_t)c * H + y) * W;
  float *orow = out + ((size_t)nc * Hp + y) * Wp;
  const bool inside = y < H;
  if (kVec) {   // W, Wp multiples of 4 and 4-element aligned rows: one 4-element load, one 16-byte store per thread
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= Wp) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (inside && x4 < W) {
      if (sizeof(T) == 1) {
        const uchar4 u = *reinterpret_cast<const uchar4 *>(row + x4);
        v = make_float4(norm1((float)u.x, m, sd), norm1((float)u.y, m, sd), norm1((float)u.z, m, sd), norm1((float)u.w, m, sd));
      } else {
        const float4 f = *reinterpret_cast<const float4 *>(row + x4);
        v = make_float4(norm1(f.x, m, sd), norm1(f.y, m, sd), norm1(f.z, m, sd), norm1(f.w, m, sd));
      }
    }
    *reinterpret_cast<float4 *>(orow + x4) = v;
  } else {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= Wp) return;
    orow[x] = (inside && x < W) ? norm1((float)row[x], m, sd) : 0.0f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) normalize_pad_nhwc_kernel(const T *__restrict__ img, long long img_stride, int C, int H, int W,
                                                                 int Hp, int Wp, ChanParams p, float *__restrict__ out) {
  const int y = blockIdx.y, n = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (x, c) in output order
  if (i >= Wp * C) return;
  const int x = i / C, c = i - x * C;
  float v = 0.0f;
  if (y < H && x < W) v = norm1((float)img[(size_t)n * img_stride + ((size_t)c * H + y) * W + x], p.mean[c], p.stdv[c]);
  out[((size_t)n * Hp + y) * Wp * C + i] = v;
}

}  // namespace

SFOD_API int sfod_normalize_pad(const void *images, int dtype, int64_t image_stride, int N, int C, int H, int W, const float *mean,
                                const float *stdv, int Hp, int Wp, int layout, float *out, sfod_stream_t stream) {
  if (N < 0 || C <= 0 || C > kMaxChannels || H <= 0 || W <= 0 || Hp < H || Wp < W || !mean || !stdv) return SFOD_ERR_INVALID_ARG;
  if (layout != SFOD_NCHW && layout != SFOD_NHWC) return SFOD_ERR_INVALID_ARG;
  if (dtype != SFOD_F32 && dtype != SFOD_U8) return SFOD_ERR_INVALID_ARG;
  if (N == 0) return SFOD_OK;
  if (!images || !out) return SFOD_ERR_INVALID_ARG;
  ChanParams p;
  for (int c = 0; c < kMaxChannels; ++c) { p.mean[c] = c < C ? mean[c] : 0.f; p.stdv[c] = c < C ? stdv[c] : 1.f; }
  cudaStream_t st = sfod_cu(stream);
  if (Hp > 65535 || (long long)N * C > 65535) return SFOD_ERR_UNSUPPORTED;
  if (layout == SFOD_NCHW) {
    const size_t esz = dtype == SFOD_U8 ? 1 : 4;
    const bool vec = (W % 4 == 0) && (Wp % 4 == 0) && (image_stride % 4 == 0) && sfod_aligned16(out) &&
                     ((reinterpret_cast<uintptr_t>(images) & (4 * esz - 1)) == 0);
    dim3 grid((unsigned)(((vec ? Wp / 4 : Wp) + 255) / 256), (unsigned)Hp, (unsigned)(N * C));
    if (dtype == SFOD_U8) {
      const unsigned char *im = static_cast<const unsigned char *>(images);
      if (vec) normalize_pad_nchw_kernel<unsigned char, true><<<grid, 256, 0, st>>>(im, image_stride, C, H, W, Hp, Wp, p, out);
      else normalize_pad_nchw_kernel<unsigned char, false><<<grid, 256, 0, st>>>(im, image_stride, C, H, W, Hp, Wp, p, out);
    } else {
      const float *im = static_cast<const float *>(images);
      if (vec) normalize_pad_nchw_kernel<float, true><<<grid, 256, 0, st>>>(im, image_stride, C, H, W, Hp, Wp, p, out);
      else normalize_pad_nchw_kernel<float, false><<<grid, 256, 0, st>>>(im, image_stride, C, H, W, Hp, Wp, p, out);
    }
  } else {
    dim3 grid((unsigned)((Wp * C + 255) / 256), (unsigned)Hp, (unsigned)N);
    if (dtype == SFOD_U8)
      normalize_pad_nhwc_kernel<unsigned char><<<grid, 256, 0, st>>>(static_cast<const unsigned char *>(images), image_stride, C, H, W, Hp, Wp, p, out);
    else
      normalize_pad_nhwc_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(images), image_stride, C, H, W, Hp, Wp, p, out);
  }
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}
