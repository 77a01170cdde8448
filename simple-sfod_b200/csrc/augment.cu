// augment.cu -- strong augmentation of the student's input on the GPU.
// Replaces the CPU / PIL pipeline of reference daod/data/detection_utils.py:7-37 (applied per image in
// daod/data/mappers/two_crop_augmentation_mapper.py:141-157):
//     RandomApply([ColorJitter(0.4, 0.4, 0.4, 0.1)], p=0.8) -> RandomGrayscale(p=0.2) -> RandomApply([GaussianBlur((0.1, 2.0))], p=0.5)
//     -> ToTensor -> 3 x RandomErasing(value="random") -> ToPILImage
// on batches of uint8 CHW images that already live in HBM.  The random DECISIONS (which ops, in which order, which factors,
// sigma, rectangles) are drawn on the host exactly as torchvision draws them and handed over as small per-image parameter
// records; the kernels do the pixel work.
//
// Arithmetic contract (this TU is compiled with -fmad=false; every op is separately rounded):
//   * brightness / contrast / saturation / hue / grayscale follow torchvision.transforms.functional on uint8 TENSORS
//     (torchvision/transforms/_functional_tensor.py: rgb_to_grayscale :148-168, _blend :258-261, _rgb2hsv :264-300,
//     _hsv2rgb :303-321, adjust_hue :196-221) operation by operation -- bit-exact, except that the image mean of the contrast
//     op is the exact integer sum / n here and a float32 cascade sum in ATen (differs in the last ulp: a pixel can move by one
//     level when its blend lands within 1e-5 of an integer);
//   * ... and, by default, the SAME ops in Pillow's arithmetic (jitter_pixel<true>): the reference's mapper hands PIL images to the
//     transforms, so torchvision's PIL path (ImageEnhance / Image.convert) is what the reference executes -- bit-exact;
//   * Gaussian blur: pil_blur_kernel reproduces Pillow's ImageFilter.GaussianBlur (three extended-box passes per axis in 8.24
//     fixed point) bit for bit -- the reference's filter; gauss_blur_kernel is a true separable Gaussian with torchvision's taps,
//     reflect padding, round-half-to-even to uint8 (optional);
//   * erasing writes byte(255 * v) (v ~ N(0, 1) from a counter-based generator, or a caller-provided noise tensor) into the
//     rectangles, which is what ToTensor -> erase(value="random") -> ToPILImage (`pic.mul(255).byte()`) leaves there.
#include "common.cuh"

namespace {

constexpr int kJThreads = 256;
constexpr int kOpBrightness = 0, kOpContrast = 1, kOpSaturation = 2, kOpHue = 3;

struct JitterRec {     // one per image; mirrors sfod_jitter_params
  int n_ops;           // 0..4 (0: ColorJitter not applied)
  int op[4];           // application order (torchvision: a random permutation of 0..3)
  float factor[4];     // factor of op[k] (float32 of the Python float)
  float one_minus[4];  // float32 of (1.0 - factor) evaluated in double, as Python does before the tensor op sees it
  int grayscale;       // RandomGrayscale hit: 3-channel gray AFTER the jitter
};

__device__ __forceinline__ float u8f(unsigned char v) { return (float)v; }
__device__ __forceinline__ unsigned char trunc_u8(float v) { return (unsigned char)(int)fminf(fmaxf(v, 0.0f), 255.0f); }

// torchvision rgb_to_grayscale on uint8: (0.2989 r + 0.587 g + 0.114 b).to(uint8)
__device__ __forceinline__ unsigned char gray_u8(unsigned char r, unsigned char g, unsigned char b) {
  const float l = __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, u8f(r)), __fmul_rn(0.587f, u8f(g))), __fmul_rn(0.114f, u8f(b)));
  return (unsigned char)(int)l;
}
// _blend(img1, img2, ratio) for uint8 img1: (ratio * img1 + (1 - ratio) * img2).clamp(0, 255).to(uint8)
__device__ __forceinline__ unsigned char blend_u8(unsigned char a, float other, float ratio, float one_minus) {
  return trunc_u8(__fadd_rn(__fmul_rn(ratio, u8f(a)), __fmul_rn(one_minus, other)));
}

__device__ __forceinline__ void hue_shift(unsigned char &R, unsigned char &G, unsigned char &B, float hue) {
  const float r = __fdiv_rn(u8f(R), 255.0f), g = __fdiv_rn(u8f(G), 255.0f), b = __fdiv_rn(u8f(B), 255.0f);
  const float maxc = fmaxf(fmaxf(r, g), b), minc = fminf(fminf(r, g), b);
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float s = __fdiv_rn(cr, eqc ? 1.0f : maxc);
  const float div = eqc ? 1.0f : cr;
  const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div), bc = __fdiv_rn(__fsub_rn(maxc, b), div);
  const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.0f;
  const float hg = ((maxc == g) && (maxc != r)) ? __fsub_rn(__fadd_rn(2.0f, rc), bc) : 0.0f;
  const float hb = ((maxc != g) && (maxc != r)) ? __fsub_rn(__fadd_rn(4.0f, gc), rc) : 0.0f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = fmodf(__fadd_rn(__fdiv_rn(h, 6.0f), 1.0f), 1.0f);
  // h = (h + hue_factor) % 1.0  (torch.remainder: result has the sign of the divisor)
  h = __fadd_rn(h, hue);
  h = __fsub_rn(h, floorf(h));           // remainder(h, 1) for |h| < 2: h - floor(h) is exact in this range
  if (h >= 1.0f) h = 0.0f;               // -tiny + 1 rounds to 1.0: remainder returns a value in [0, 1)
  const float v = maxc;
  const float h6 = __fmul_rn(h, 6.0f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  int i = (int)fi; i = i % 6;
  const float p = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.0f, s)), 0.0f), 1.0f);
  const float q = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, f))), 0.0f), 1.0f);
  const float t = fminf(fmaxf(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, __fsub_rn(1.0f, f)))), 0.0f), 1.0f);
  float ro, go, bo;
  switch (i) {
    case 0: ro = v; go = t; bo = p; break;
    case 1: ro = q; go = v; bo = p; break;
    case 2: ro = p; go = v; bo = t; break;
    case 3: ro = p; go = q; bo = v; break;
    case 4: ro = t; go = p; bo = v; break;
    default: ro = v; go = p; bo = q; break;
  }
  // convert_image_dtype(float -> uint8): image.mul(255 + 1.0 - 1e-3).to(uint8)
  R = (unsigned char)(int)__fmul_rn(ro, 255.999f);
  G = (unsigned char)(int)__fmul_rn(go, 255.999f);
  B = (unsigned char)(int)__fmul_rn(bo, 255.999f);
}

// ---- the same ops in PILLOW's arithmetic: the reference hands PIL images to torchvision's transforms
// (daod/data/mappers/two_crop_augmentation_mapper.py:141-157), whose PIL path is ImageEnhance / Image.convert, not the tensor
// code above.  Restated from Pillow's libImaging (Blend.c, Convert.c) and torchvision/transforms/_functional_pil.py and pinned
// against the installed Pillow (oracle/pil_cpu.py, tests): bit-exact.
//   convert("L"):   L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16
//   Image.blend(degenerate, image, alpha): (UINT8)(float)(deg + alpha * (img - deg)), clipped to [0, 255] when alpha is outside [0, 1]
//   Brightness: degenerate = 0;  Contrast: degenerate = int(mean(L) + 0.5);  Color (saturation): degenerate = L
//   adjust_hue: convert("HSV") (colorsys in float / double, see below), h += uint8(int32(255 hue)) mod 256, convert back
__device__ __forceinline__ unsigned char pil_L(unsigned char r, unsigned char g, unsigned char b) {
  return (unsigned char)((r * 19595u + g * 38470u + b * 7471u + 0x8000u) >> 16);
}
__device__ __forceinline__ unsigned char pil_blend(int deg, unsigned char img, float alpha) {
  const float t = __fadd_rn((float)deg, __fmul_rn(alpha, (float)((int)img - deg)));
  if (alpha >= 0.0f && alpha <= 1.0f) return (unsigned char)(int)t;
  return t <= 0.0f ? (unsigned char)0 : t >= 255.0f ? (unsigned char)255 : (unsigned char)(int)t;
}
__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
__device__ __forceinline__ void pil_hue_shift(unsigned char &R, unsigned char &G, unsigned char &B, int shift) {
  // rgb2hsv_row (Convert.c): float h, s, rc, gc, bc, cr; the literals 2.0 / 4.0 / 6.0 / 255.0 are doubles
  const int r = R, g = G, b = B;
  const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
  int uh = 0, us = 0;
  const int uv = maxc;
  if (minc != maxc) {
    const float cr = (float)(maxc - minc);
    const float s = __fdiv_rn(cr, (float)maxc);
    const float rc = __fdiv_rn((float)(maxc - r), cr), gc = __fdiv_rn((float)(maxc - g), cr), bc = __fdiv_rn((float)(maxc - b), cr);
    float h;
    if (r == maxc) h = __fsub_rn(bc, gc);
    else if (g == maxc) h = (float)(__dsub_rn(__dadd_rn(2.0, (double)rc), (double)bc));
    else h = (float)(__dsub_rn(__dadd_rn(4.0, (double)gc), (double)rc));
    const double hd = __dadd_rn(__ddiv_rn((double)h, 6.0), 1.0);
    h = (float)(hd - floor(hd));                                   // fmod(x, 1.0) for x in (0, 2): exact
    uh = clip8((int)__dmul_rn((double)h, 255.0));
    us = clip8((int)__dmul_rn((double)s, 255.0));
  }
  uh = (uh + shift) & 255;
  // hsv2rgb (Convert.c)
  if (us == 0) { R = G = B = (unsigned char)uv; return; }
  const double h6 = __ddiv_rn(__dmul_rn((double)(float)uh, 6.0), 255.0);
  const int i = (int)floor(h6);
  const float f = (float)__dsub_rn(h6, (double)(float)i);
  const float fs = (float)__ddiv_rn((double)(float)us, 255.0);
  const double v = (double)(float)uv;
  const int pp = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, (double)fs))));
  const int qq = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn((double)fs, (double)f)))));
  const int tt = clip8((int)round(__dmul_rn(v, __dsub_rn(1.0, __dmul_rn((double)fs, __dsub_rn(1.0, (double)f))))));
  int ro, go, bo;
  switch (i % 6) {
    case 0: ro = uv; go = tt; bo = pp; break;
    case 1: ro = qq; go = uv; bo = pp; break;
    case 2: ro = pp; go = uv; bo = tt; break;
    case 3: ro = pp; go = qq; bo = uv; break;
    case 4: ro = tt; go = pp; bo = uv; break;
    default: ro = uv; go = pp; bo = qq; break;
  }
  R = (unsigned char)ro; G = (unsigned char)go; B = (unsigned char)bo;
}

template <bool kPil>
__device__ __forceinline__ unsigned char gray_of(unsigned char r, unsigned char g, unsigned char b) { return kPil ? pil_L(r, g, b) : gray_u8(r, g, b); }

// applies ops [0, upto) of the record to one pixel; `mean` = gray mean of the image at the point where contrast is applied
// (kPil: the integer int(mean + 0.5); the hue record carries the shift byte in one_minus)
template <bool kPil>
__device__ __forceinline__ void jitter_pixel(const JitterRec &p, int upto, float mean, unsigned char &r, unsigned char &g, unsigned char &b) {
  for (int k = 0; k < upto; ++k) {
    const float f = p.factor[k], om = p.one_minus[k];
    if (kPil) {
      switch (p.op[k]) {
        case kOpBrightness: r = pil_blend(0, r, f); g = pil_blend(0, g, f); b = pil_blend(0, b, f); break;
        case kOpContrast: { const int m = (int)mean; r = pil_blend(m, r, f); g = pil_blend(m, g, f); b = pil_blend(m, b, f); break; }
        case kOpSaturation: { const int l = pil_L(r, g, b); r = pil_blend(l, r, f); g = pil_blend(l, g, f); b = pil_blend(l, b, f); break; }
        default: pil_hue_shift(r, g, b, (int)om); break;
      }
    } else {
      switch (p.op[k]) {
        case kOpBrightness: r = blend_u8(r, 0.0f, f, om); g = blend_u8(g, 0.0f, f, om); b = blend_u8(b, 0.0f, f, om); break;
        case kOpContrast: r = blend_u8(r, mean, f, om); g = blend_u8(g, mean, f, om); b = blend_u8(b, mean, f, om); break;
        case kOpSaturation: { const float l = u8f(gray_u8(r, g, b)); r = blend_u8(r, l, f, om); g = blend_u8(g, l, f, om); b = blend_u8(b, l, f, om); break; }
        default: hue_shift(r, g, b, f); break;
      }
    }
  }
}

__device__ __forceinline__ int contrast_pos(const JitterRec &p) {
  for (int k = 0; k < p.n_ops; ++k) if (p.op[k] == kOpContrast) return k;
  return -1;
}

// pass A: sum of the gray levels of every image at the point where its contrast op applies (exact integer sum)
template <bool kPil>
__global__ void __launch_bounds__(kJThreads) jitter_gray_sum_kernel(const unsigned char *__restrict__ img, const JitterRec *__restrict__ recs,
                                                                    int HW, unsigned long long *__restrict__ sums) {
  const int n = blockIdx.y;
  const JitterRec p = recs[n];
  const int cp = contrast_pos(p);
  if (cp < 0) return;
  const unsigned char *R = img + (size_t)n * 3 * HW, *G = R + HW, *B = G + HW;
  unsigned int local = 0;
  for (int i = blockIdx.x * kJThreads + threadIdx.x; i < HW; i += gridDim.x * kJThreads) {
    unsigned char r = R[i], g = G[i], b = B[i];
    jitter_pixel<kPil>(p, cp, 0.0f, r, g, b);
    local += gray_of<kPil>(r, g, b);
  }
  local = __reduce_add_sync(0xFFFFFFFFu, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(sums + n, (unsigned long long)local);
}

// pass B: the whole op sequence (+ RandomGrayscale) per pixel
template <bool kPil>
__global__ void __launch_bounds__(kJThreads) jitter_apply_kernel(const unsigned char *__restrict__ img, const JitterRec *__restrict__ recs,
                                                                 const unsigned long long *__restrict__ sums, int HW,
                                                                 unsigned char *__restrict__ out) {
  const int n = blockIdx.y;
  const JitterRec p = recs[n];
  // torch.mean of the float32 gray image; here: exact sum / n rounded once
  // (kPil: ImageStat mean = sum / count as a Python float, then int(mean + 0.5))
  const double dmean = contrast_pos(p) >= 0 ? (double)sums[n] / (double)HW : 0.0;
  const float mean = kPil ? (float)(int)(dmean + 0.5) : (float)dmean;
  const unsigned char *R = img + (size_t)n * 3 * HW, *G = R + HW, *B = G + HW;
  unsigned char *Ro = out + (size_t)n * 3 * HW, *Go = Ro + HW, *Bo = Go + HW;
  for (int i = blockIdx.x * kJThreads + threadIdx.x; i < HW; i += gridDim.x * kJThreads) {
    unsigned char r = R[i], g = G[i], b = B[i];
    jitter_pixel<kPil>(p, p.n_ops, mean, r, g, b);
    if (p.grayscale) { const unsigned char l = gray_of<kPil>(r, g, b); r = g = b = l; }
    Ro[i] = r; Go[i] = g; Bo[i] = b;
  }
}

// ---- Gaussian blur: one CTA = 32 x 32 output pixels of one plane; (32 + 2R)^2 input tile in shared memory (reflect padding),
// horizontal then vertical pass in fp32, round-half-to-even (torch.round) to uint8.  taps: (N, 2*kMaxRadius+1) floats, radius: (N).
constexpr int kMaxRadius = 15;
constexpr int kTile = 32;

__device__ __forceinline__ int reflect(int i, int n) {   // torch 'reflect' padding (no edge repeat); valid for pad < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) gauss_blur_kernel(const unsigned char *__restrict__ img, const float *__restrict__ taps,
                                                         const int *__restrict__ radius, int H, int W, unsigned char *__restrict__ out) {
  extern __shared__ float gb_smem[];
  const int plane = blockIdx.z;          // n * 3 + c
  const int n = plane / 3;
  const int Rr = radius[n];
  const unsigned char *src = img + (size_t)plane * H * W;
  unsigned char *dst = out + (size_t)plane * H * W;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  if (Rr <= 0) {                          // not blurred: copy
    for (int t = threadIdx.x; t < kTile * kTile; t += 256) {
      const int y = y0 + t / kTile, x = x0 + t % kTile;
      if (y < H && x < W) dst[(size_t)y * W + x] = src[(size_t)y * W + x];
    }
    return;
  }
  const int TW = kTile + 2 * Rr;
  float *tile = gb_smem;                  // TW x TW input
  float *hbuf = gb_smem + TW * TW;        // TW rows x kTile columns (horizontal pass)
  float *w = hbuf + TW * kTile;           // 2R + 1 taps
  const float *tp = taps + (size_t)n * (2 * kMaxRadius + 1);
  for (int t = threadIdx.x; t < 2 * Rr + 1; t += 256) w[t] = tp[t];
  for (int t = threadIdx.x; t < TW * TW; t += 256) {
    const int ty = t / TW, tx = t - ty * TW;
    const int y = reflect(y0 + ty - Rr, H), x = reflect(x0 + tx - Rr, W);
    tile[t] = (float)src[(size_t)y * W + x];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TW * kTile; t += 256) {
    const int ty = t / kTile, tx = t - ty * kTile;
    float acc = 0.0f;
    for (int k = 0; k <= 2 * Rr; ++k) acc = __fadd_rn(acc, __fmul_rn(w[k], tile[ty * TW + tx + k]));
    hbuf[t] = acc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kTile * kTile; t += 256) {
    const int ty = t / kTile, tx = t - ty * kTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float acc = 0.0f;
    for (int k = 0; k <= 2 * Rr; ++k) acc = __fadd_rn(acc, __fmul_rn(w[k], hbuf[(ty + k) * kTile + tx]));
    dst[(size_t)y * W + x] = (unsigned char)(int)fminf(fmaxf(rintf(acc), 0.0f), 255.0f);
  }
}

// ---- PIL's ImageFilter.GaussianBlur, bit for bit.  The reference blurs with Pillow (reference daod/data/transforms/augmentations.py:18-21:
// x.filter(ImageFilter.GaussianBlur(radius=sigma))), which is NOT a Gaussian convolution: libImaging/BoxBlur.c approximates it with
// three passes of an "extended box blur" per axis (rows first, then columns), each pass in 8.24 fixed point and rounded to uint8:
//   out[x] = (ww * sum_{|d| <= r} in[c(x + d)] + fw * (in[c(x - r - 1)] + in[c(x + r + 1)]) + 2^23) >> 24,   c = clamp to the line,
// r = int(R), ww = uint32(2^24 / (2 R + 1)), fw = (2^24 - (2 r + 1) ww) / 2, R = the fractional box radius derived from sigma (host,
// ops.pil_blur_params).  All six passes run on one shared-memory tile: 32 x 32 outputs plus a halo of 3 (r + 1) pixels per side;
// every pass clamps in IMAGE coordinates, exactly as the line-by-line C code does, so borders are identical too.
constexpr int kPilMaxRadius = 9;   // box radius; sigma <= ~10

__global__ void __launch_bounds__(256) pil_blur_kernel(const unsigned char *__restrict__ img, const int *__restrict__ radius,
                                                       const unsigned *__restrict__ wws, const unsigned *__restrict__ fws, int H, int W,
                                                       unsigned char *__restrict__ out) {
  extern __shared__ unsigned char pb_smem[];
  const int plane = blockIdx.z;          // n * 3 + c
  const int n = plane / 3;
  const int r = radius[n];
  const unsigned char *src = img + (size_t)plane * H * W;
  unsigned char *dst = out + (size_t)plane * H * W;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  if (r < 0) {                            // not blurred: copy
    for (int t = threadIdx.x; t < kTile * kTile; t += 256) {
      const int y = y0 + t / kTile, x = x0 + t % kTile;
      if (y < H && x < W) dst[(size_t)y * W + x] = src[(size_t)y * W + x];
    }
    return;
  }
  const unsigned ww = wws[n], fw = fws[n];
  const int h = 3 * (r + 1), TW = kTile + 2 * h;
  unsigned char *A = pb_smem, *B = pb_smem + TW * TW;
  const int ox = x0 - h, oy = y0 - h;     // image coordinates of tile cell (0, 0)
  for (int t = threadIdx.x; t < TW * TW; t += 256) {
    const int ty = t / TW, tx = t - ty * TW;
    const int y = min(max(oy + ty, 0), H - 1), x = min(max(ox + tx, 0), W - 1);
    A[t] = src[(size_t)y * W + x];
  }
  __syncthreads();
  // three horizontal passes over every row of the tile that lies inside the image
  for (int p = 1; p <= 3; ++p) {
    const int lo = p * (r + 1), wid = TW - 2 * lo;
    for (int t = threadIdx.x; t < TW * wid; t += 256) {
      const int ty = t / wid, tx = lo + (t - ty * wid);
      const int yi = oy + ty, xi = ox + tx;
      if (yi < 0 || yi >= H || xi < 0 || xi >= W) continue;
      const unsigned char *row = A + ty * TW;
      unsigned acc = 0;
      for (int d = -r; d <= r; ++d) acc += row[min(max(xi + d, 0), W - 1) - ox];
      const unsigned far = (unsigned)row[min(max(xi - r - 1, 0), W - 1) - ox] + (unsigned)row[min(max(xi + r + 1, 0), W - 1) - ox];
      B[ty * TW + tx] = (unsigned char)((acc * ww + far * fw + (1u << 23)) >> 24);
    }
    __syncthreads();
    unsigned char *tmp = A; A = B; B = tmp;
  }
  // three vertical passes over the output columns
  for (int p = 1; p <= 3; ++p) {
    const int lo = p * (r + 1), hgt = TW - 2 * lo;
    for (int t = threadIdx.x; t < hgt * kTile; t += 256) {
      const int ty = lo + t / kTile, tx = h + t % kTile;
      const int yi = oy + ty, xi = ox + tx;
      if (yi < 0 || yi >= H || xi >= W) continue;
      unsigned acc = 0;
      for (int d = -r; d <= r; ++d) acc += A[(min(max(yi + d, 0), H - 1) - oy) * TW + tx];
      const unsigned far = (unsigned)A[(min(max(yi - r - 1, 0), H - 1) - oy) * TW + tx] + (unsigned)A[(min(max(yi + r + 1, 0), H - 1) - oy) * TW + tx];
      B[ty * TW + tx] = (unsigned char)((acc * ww + far * fw + (1u << 23)) >> 24);
    }
    __syncthreads();
    unsigned char *tmp = A; A = B; B = tmp;
  }
  for (int t = threadIdx.x; t < kTile * kTile; t += 256) {
    const int ty = t / kTile, tx = t - ty * kTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) dst[(size_t)y * W + x] = A[(h + ty) * TW + h + tx];
  }
}

// ---- RandomErasing(value="random"): up to kMaxRects rectangles per image, applied in order
constexpr int kMaxRects = 4;
struct EraseRec { int n_rects; int rect[kMaxRects][4]; };   // (i, j, h, w) = top, left, height, width

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// byte(255 * v): float -> int32 (truncation) -> low 8 bits, the value ToPILImage's `pic.mul(255).byte()` stores on x86
__device__ __forceinline__ unsigned char noise_byte(float v) { return (unsigned char)(__float2int_rz(__fmul_rn(v, 255.0f)) & 0xFF); }

__global__ void __launch_bounds__(256) erase_kernel(unsigned char *__restrict__ img, const EraseRec *__restrict__ recs,
                                                    const float *__restrict__ noise /* (N, kMaxRects, 3, H, W) or NULL */, int H, int W,
                                                    unsigned long long seed) {
  const int n = blockIdx.z / kMaxRects, k = blockIdx.z % kMaxRects;
  const EraseRec p = recs[n];
  if (k >= p.n_rects) return;
  const int i0 = p.rect[k][0], j0 = p.rect[k][1], h = p.rect[k][2], w = p.rect[k][3];
  // a later rectangle overwrites an earlier one: pixel ownership = the LAST rectangle that covers it
  const long long total = (long long)3 * h * w;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
    const int c = (int)(t / ((long long)h * w));
    const int rem = (int)(t - (long long)c * h * w);
    const int y = i0 + rem / w, x = j0 + rem % w;
    bool owned = true;
    for (int q = k + 1; q < p.n_rects; ++q)
      owned = owned && !(y >= p.rect[q][0] && y < p.rect[q][0] + p.rect[q][2] && x >= p.rect[q][1] && x < p.rect[q][1] + p.rect[q][3]);
    if (!owned) continue;
    float v;
    if (noise) {
      v = noise[((((size_t)n * kMaxRects + k) * 3 + c) * H + y) * W + x];
    } else {   // Box-Muller on two counter-based uniforms
      const unsigned long long key = mix64(seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)(n * kMaxRects + k) << 40) | ((unsigned long long)c << 36) | (unsigned long long)(y * W + x)));
      const float u1 = ((float)(unsigned int)(key >> 40) + 1.0f) * (1.0f / 16777217.0f);   // (0, 1)
      const float u2 = (float)(unsigned int)((key >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
      v = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    }
    img[((size_t)n * 3 + c) * H * W + (size_t)y * W + x] = noise_byte(v);
  }
}

}  // namespace

// =========================================================================== C ABI
static int color_jitter_impl(bool pil, const uint8_t *images, int N, int H, int W, const sfod_jitter_params *params_dev, void *workspace,
                             size_t workspace_bytes, uint8_t *out, sfod_stream_t stream) {
  static_assert(sizeof(sfod_jitter_params) == sizeof(JitterRec), "parameter record layout");
  if (N < 0 || H <= 0 || W <= 0) return SFOD_ERR_INVALID_ARG;
  if (N == 0) return SFOD_OK;
  if (!images || !params_dev || !out) return SFOD_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < (size_t)N * sizeof(unsigned long long)) return SFOD_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = sfod_cu(stream);
  unsigned long long *sums = static_cast<unsigned long long *>(workspace);
  SFOD_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)N * sizeof(unsigned long long), st));
  const int HW = H * W;
  const unsigned gx = (unsigned)min(SFOD_NUM_SMS * 2, (HW + kJThreads - 1) / kJThreads);
  const JitterRec *recs = reinterpret_cast<const JitterRec *>(params_dev);
  if (pil) jitter_gray_sum_kernel<true><<<dim3(gx, N), kJThreads, 0, st>>>(images, recs, HW, sums);
  else jitter_gray_sum_kernel<false><<<dim3(gx, N), kJThreads, 0, st>>>(images, recs, HW, sums);
  SFOD_LAUNCH_CHECK();
  if (pil) jitter_apply_kernel<true><<<dim3(gx, N), kJThreads, 0, st>>>(images, recs, sums, HW, out);
  else jitter_apply_kernel<false><<<dim3(gx, N), kJThreads, 0, st>>>(images, recs, sums, HW, out);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

SFOD_API int sfod_color_jitter(const uint8_t *images, int N, int H, int W, const sfod_jitter_params *params_dev, void *workspace,
                               size_t workspace_bytes, uint8_t *out, sfod_stream_t stream) {
  return color_jitter_impl(false, images, N, H, W, params_dev, workspace, workspace_bytes, out, stream);
}

SFOD_API int sfod_color_jitter_pil(const uint8_t *images, int N, int H, int W, const sfod_jitter_params *params_dev, void *workspace,
                                   size_t workspace_bytes, uint8_t *out, sfod_stream_t stream) {
  return color_jitter_impl(true, images, N, H, W, params_dev, workspace, workspace_bytes, out, stream);
}

SFOD_API int sfod_gaussian_blur(const uint8_t *images, int N, int H, int W, const float *taps_dev, const int32_t *radius_dev,
                                int max_radius, uint8_t *out, sfod_stream_t stream) {
  if (N < 0 || H <= 0 || W <= 0 || max_radius < 0 || max_radius > kMaxRadius) return SFOD_ERR_INVALID_ARG;
  if (max_radius >= H || max_radius >= W) return SFOD_ERR_INVALID_ARG;   // reflect padding needs pad < size
  if (N == 0) return SFOD_OK;
  if (!images || !taps_dev || !radius_dev || !out || images == out) return SFOD_ERR_INVALID_ARG;
  const int TW = kTile + 2 * max_radius;
  const size_t smem = ((size_t)TW * TW + (size_t)TW * kTile + 2 * kMaxRadius + 1) * sizeof(float);
  SFOD_CUDA_TRY(cudaFuncSetAttribute(gauss_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N * 3);
  gauss_blur_kernel<<<grid, 256, smem, sfod_cu(stream)>>>(images, taps_dev, radius_dev, H, W, out);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

SFOD_API int sfod_gaussian_blur_pil(const uint8_t *images, int N, int H, int W, const int32_t *radius_dev, const uint32_t *ww_dev,
                                    const uint32_t *fw_dev, int max_radius, uint8_t *out, sfod_stream_t stream) {
  if (N < 0 || H <= 0 || W <= 0 || max_radius < 0 || max_radius > kPilMaxRadius) return SFOD_ERR_INVALID_ARG;
  if (N == 0) return SFOD_OK;
  if (!images || !radius_dev || !ww_dev || !fw_dev || !out || images == out) return SFOD_ERR_INVALID_ARG;
  const int TW = kTile + 6 * (max_radius + 1);
  const size_t smem = 2 * (size_t)TW * TW;
  SFOD_CUDA_TRY(cudaFuncSetAttribute(pil_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N * 3);
  pil_blur_kernel<<<grid, 256, smem, sfod_cu(stream)>>>(images, radius_dev, ww_dev, fw_dev, H, W, out);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

SFOD_API int sfod_random_erase(uint8_t *images, int N, int H, int W, const sfod_erase_params *params_dev, const float *noise,
                               uint64_t seed, sfod_stream_t stream) {
  static_assert(sizeof(sfod_erase_params) == sizeof(EraseRec), "parameter record layout");
  if (N < 0 || H <= 0 || W <= 0) return SFOD_ERR_INVALID_ARG;
  if (N == 0) return SFOD_OK;
  if (!images || !params_dev) return SFOD_ERR_INVALID_ARG;
  dim3 grid(64, 1, N * kMaxRects);
  erase_kernel<<<grid, 256, 0, sfod_cu(stream)>>>(images, reinterpret_cast<const EraseRec *>(params_dev), noise, H, W,
                                                 (unsigned long long)seed);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}
