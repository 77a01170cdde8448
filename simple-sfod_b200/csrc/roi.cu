// roi.cu -- ROIAlign (V1/V2) and ROIPool, forward and backward.
// Replaces torchvision.ops.roi_align / roi_pool as reached from d2 ROIPooler, constructed at reference
// daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:42-47 and called at :117.
//
// Forward is write-bound: 100 352 B of output per VGG ROI against a 1.36 MB, L2-resident feature map
// (202.1 MB algorithmic for 2000 ROIs).  The fast kernel therefore
//   * reads a channels-last (NHWC) feature map so that one warp load covers 64 consecutive channels of
//     one pixel (fully coalesced), instead of torchvision's 4 scattered scalar taps per output element;
//   * exploits that bilinear sampling + bin averaging is separable: out = A . F . B^T per channel, with
//     A (PH x H) and B (PW x W) built once per ROI in shared memory -- each feature pixel of the ROI
//     region is loaded once per thread, not once per overlapping sample tap;
//   * stages the (256 channels x 49 bins) output tile in shared memory with a bank-conflict-free lane
//     permutation and writes it with ONE cp.async.bulk (TMA bulk) store of 50 176 contiguous bytes.
// The result differs from torchvision's per-sample summation order by a few ulp (<= 1e-5 relative,
// tested); `exact` selects a gather kernel that reproduces torchvision's order bit for bit.
#include "common.cuh"

namespace {

// --------------------------------------------------------------------------- layout transposes
// (N, C, HW) <-> (N, HW, C), 32x32 shared-memory tiles, coalesced on both sides.
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows,
                                                        int cols) {
  // src: (batch, rows, cols) -> dst: (batch, cols, rows)
  __shared__ float tile[32][33];
  const size_t boff = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = src[boff + (size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[boff + (size_t)c * rows + r] = tile[tx][i];
  }
}

int launch_transpose(const float *src, float *dst, int batch, int rows, int cols, cudaStream_t st) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return SFOD_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_kernel<<<grid, 256, 0, st>>>(src, dst, rows, cols);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

// --------------------------------------------------------------------------- ROI geometry (torchvision order)
struct RoiGeom {
  int n;                 // batch index
  float sh, sw, bh, bw;  // roi start (h, w), bin size
  int gh, gw;            // sampling grid
};

__device__ __forceinline__ RoiGeom roi_geometry(const float *__restrict__ roi, float scale, int aligned, int PH, int PW,
                                                int sampling_ratio) {
  RoiGeom g;
  g.n = (int)roi[0];
  const float offset = aligned ? 0.5f : 0.0f;
  const float rsw = __fsub_rn(__fmul_rn(roi[1], scale), offset);
  const float rsh = __fsub_rn(__fmul_rn(roi[2], scale), offset);
  const float rew = __fsub_rn(__fmul_rn(roi[3], scale), offset);
  const float reh = __fsub_rn(__fmul_rn(roi[4], scale), offset);
  float rw = __fsub_rn(rew, rsw), rh = __fsub_rn(reh, rsh);
  if (!aligned) { rw = fmaxf(rw, 1.0f); rh = fmaxf(rh, 1.0f); }
  g.bh = __fdiv_rn(rh, (float)PH);
  g.bw = __fdiv_rn(rw, (float)PW);
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  g.sh = rsh; g.sw = rsw;
  return g;
}

// sample coordinate: start + p * bin + (i + .5) * bin / grid   (left-to-right, separately rounded)
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
}

// 1-D half of torchvision's bilinear_interpolate: returns false when the sample is outside [-1, size].
__device__ __forceinline__ bool bilinear_1d(float v, int size, int &lo, int &hi, float &l, float &h) {
  if (v < -1.0f || v > (float)size) return false;
  if (v <= 0.f) v = 0.f;
  lo = (int)v;
  if (lo >= size - 1) { hi = lo = size - 1; v = (float)lo; } else hi = lo + 1;
  l = __fsub_rn(v, (float)lo);
  h = __fsub_rn(1.0f, l);
  return true;
}

// --------------------------------------------------------------------------- exact forward (NCHW gather)
// One thread per output element; the summation order is torchvision's:
//   val += ((w1*v1 + w2*v2) + w3*v3) + w4*v4 over (iy, ix), then val / count.
__global__ void __launch_bounds__(256) roi_align_fwd_exact_kernel(const float *__restrict__ input, const float *__restrict__ rois,
                                                                  int N, int C, int H, int W, long long total, int PH, int PW,
                                                                  float scale, int sampling_ratio, int aligned,
                                                                  float *__restrict__ output) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
  const int c = (int)((idx / ((long long)PW * PH)) % C);
  const long long r = idx / ((long long)PW * PH * C);
  const RoiGeom g = roi_geometry(rois + 5 * r, scale, aligned, PH, PW, sampling_ratio);
  float val = 0.f;
  if (g.n >= 0 && g.n < N) {
    const float *in = input + ((size_t)g.n * C + c) * H * W;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float y = sample_coord(g.sh, ph, g.bh, iy, g.gh);
      int yl, yh; float ly, hy;
      const bool yok = bilinear_1d(y, H, yl, yh, ly, hy);
      for (int ix = 0; ix < g.gw; ++ix) {
        const float x = sample_coord(g.sw, pw, g.bw, ix, g.gw);
        int xl, xh; float lx, hx;
        if (!yok || !bilinear_1d(x, W, xl, xh, lx, hx)) continue;
        const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
        float t = __fadd_rn(__fmul_rn(w1, in[yl * W + xl]), __fmul_rn(w2, in[yl * W + xh]));
        t = __fadd_rn(t, __fmul_rn(w3, in[yh * W + xl]));
        t = __fadd_rn(t, __fmul_rn(w4, in[yh * W + xh]));
        val = __fadd_rn(val, t);
      }
    }
  }
  int cnt = g.gh * g.gw; if (cnt < 1) cnt = 1;
  output[idx] = __fdiv_rn(val, (float)cnt);
}

// generic backward (NCHW, one thread per output-gradient element, atomics), torchvision's formulation
__global__ void __launch_bounds__(256) roi_align_bwd_generic_kernel(const float *__restrict__ grad_out,
                                                                    const float *__restrict__ rois, int N, int C, int H, int W,
                                                                    long long total, int PH, int PW, float scale,
                                                                    int sampling_ratio, int aligned, float *__restrict__ grad_in) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
  const int c = (int)((idx / ((long long)PW * PH)) % C);
  const long long r = idx / ((long long)PW * PH * C);
  const RoiGeom g = roi_geometry(rois + 5 * r, scale, aligned, PH, PW, sampling_ratio);
  if (g.n < 0 || g.n >= N) return;
  const float go = grad_out[idx];
  const float count = (float)(g.gh * g.gw);
  float *gi = grad_in + ((size_t)g.n * C + c) * H * W;
  for (int iy = 0; iy < g.gh; ++iy) {
    const float y = sample_coord(g.sh, ph, g.bh, iy, g.gh);
    int yl, yh; float ly, hy;
    if (!bilinear_1d(y, H, yl, yh, ly, hy)) continue;
    for (int ix = 0; ix < g.gw; ++ix) {
      const float x = sample_coord(g.sw, pw, g.bw, ix, g.gw);
      int xl, xh; float lx, hx;
      if (!bilinear_1d(x, W, xl, xh, lx, hx)) continue;
      atomicAdd(gi + yl * W + xl, __fdiv_rn(__fmul_rn(go, __fmul_rn(hy, hx)), count));
      atomicAdd(gi + yl * W + xh, __fdiv_rn(__fmul_rn(go, __fmul_rn(hy, lx)), count));
      atomicAdd(gi + yh * W + xl, __fdiv_rn(__fmul_rn(go, __fmul_rn(ly, hx)), count));
      atomicAdd(gi + yh * W + xh, __fdiv_rn(__fmul_rn(go, __fmul_rn(ly, lx)), count));
    }
  }
}

// --------------------------------------------------------------------------- separable fast path (7x7, NHWC)
// Bilinear sampling + bin averaging is separable: out = A . F . B^T per channel, A (7 x H) and B (7 x W) built once per
// ROI in shared memory.  The kernel is bound by instruction issue and L2 latency, not by DRAM (the feature map is
// L2-resident and the 100 KB/ROI output is the only HBM traffic), so it is organised to minimise instructions per pixel:
//   * one CTA = one ROI x 256 consecutive channels, 2 channels per thread (one LDG.64 per pixel, a warp reads 256
//     contiguous bytes); the weight tables are warp-uniform shared-memory broadcasts (2 LDS.128 per pixel);
//   * the channel stride is a template constant for the common C, so the <= 8 pixel loads of a batch are LDG.64 with
//     immediate offsets from one base pointer (no per-load address arithmetic), issued before any FMA consumes them,
//     and partial batches execute exactly their valid pixels (warp-uniform branches, no padded work);
//   * the 7 bin rows are produced in two passes (bins 0-3, then 4-6), which halves the accumulator registers
//     (28 pairs instead of 49) and lifts occupancy to 4 CTAs/SM; only rows shared by bins 3 and 4 are visited twice;
//   * the (256 x 49) output tile is staged in shared memory with a conflict-free lane permutation and leaves the SM
//     as ONE 50 176 B cp.async.bulk (TMA) store.
constexpr int kPH = 7, kPW = 7, kBins = kPH * kPW;
constexpr int kSepThreads = 128;           // 2 channels per thread
constexpr int kCT = 2 * kSepThreads;       // 256 channels per CTA
constexpr int kTileFloats = kCT * kBins;   // 12 544 floats = 50 176 B
constexpr int kXB = 8;                     // pixels per load batch

struct SepSmem {
  float *tile;   // kTileFloats
  float *Ad;     // H * 8: Ad[y*8 + ph] = sum of y-weights of bin ph on row y, divided by grid_h
  float *Bd;     // W * 8
  int *lim;      // [0..3] ymin, ymax, xmin, xmax; [4..10] first row of bin ph; [11..17] last row of bin ph;
                 // [18..24] first column of the compact window of bin pw; [25] widest bin window in columns (0 = none)
  float *Bc;     // kPW * kNXMax: Bc[pw*kNXMax + j] = Bd[(lim[18+pw] + j)*8 + pw], the nonzero run of column weights of bin pw
};
constexpr int kNXMax = 4;                  // widest per-bin column window handled by the compact (sparse-in-x) forward
__host__ __device__ inline size_t sep_smem_bytes(int H, int W) {
  return (size_t)kTileFloats * 4 + (size_t)(H + W) * 8 * 4 + 32 * 4 + 32 * 4;
}
__device__ __forceinline__ SepSmem sep_carve(float *base, int H, int W) {
  SepSmem s;
  s.tile = base; s.Ad = base + kTileFloats; s.Bd = s.Ad + H * 8;
  s.lim = reinterpret_cast<int *>(s.Bd + W * 8);
  s.Bc = reinterpret_cast<float *>(s.lim + 32);
  return s;
}

// Build the dense separable weight tables of one ROI.  Called by all threads of the CTA.
__device__ __forceinline__ void sep_build_tables(const SepSmem &s, const RoiGeom &g, int H, int W) {
  const int tid = threadIdx.x;
  for (int i = tid; i < (H + W) * 8; i += blockDim.x) s.Ad[i] = 0.f;  // Ad and Bd are contiguous
  if (tid == 0) { s.lim[0] = H; s.lim[1] = -1; s.lim[2] = W; s.lim[3] = -1; s.lim[25] = 0; }
  __syncthreads();
  if (tid < kPH) {
    const int ph = tid; int mn = H, mx = -1;
    const float inv = g.gh > 0 ? __fdiv_rn(1.0f, (float)g.gh) : 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
      int lo, hi; float l, h;
      if (!bilinear_1d(sample_coord(g.sh, ph, g.bh, iy, g.gh), H, lo, hi, l, h)) continue;
      s.Ad[lo * 8 + ph] += h * inv; s.Ad[hi * 8 + ph] += l * inv;
      mn = min(mn, lo); mx = max(mx, hi);
    }
    s.lim[4 + ph] = mn; s.lim[11 + ph] = mx;
    if (mx >= 0) { atomicMin(&s.lim[0], mn); atomicMax(&s.lim[1], mx); }
  } else if (tid >= 32 && tid < 32 + kPW) {
    const int pw = tid - 32; int mn = W, mx = -1;
    const float inv = g.gw > 0 ? __fdiv_rn(1.0f, (float)g.gw) : 0.f;
    for (int ix = 0; ix < g.gw; ++ix) {
      int lo, hi; float l, h;
      if (!bilinear_1d(sample_coord(g.sw, pw, g.bw, ix, g.gw), W, lo, hi, l, h)) continue;
      s.Bd[lo * 8 + pw] += h * inv; s.Bd[hi * 8 + pw] += l * inv;
      mn = min(mn, lo); mx = max(mx, hi);
    }
    if (mx >= 0) { atomicMin(&s.lim[2], mn); atomicMax(&s.lim[3], mx); }
    // compact window of this bin: columns [x0, x0 + kNXMax) inside the row, covering [mn, mx] whenever it is narrow enough
    // (only this thread touched column pw of Bd, so the dense entries can be read back without a barrier)
    const int x0 = mx >= 0 ? max(0, min(mn, W - kNXMax)) : 0;
    s.lim[18 + pw] = x0;
#pragma unroll
    for (int j = 0; j < kNXMax; ++j) s.Bc[pw * kNXMax + j] = (mx >= 0 && x0 + j < W) ? s.Bd[(x0 + j) * 8 + pw] : 0.f;
    if (mx >= 0) atomicMax(&s.lim[25], W >= kNXMax ? mx - x0 + 1 : kNXMax + 1);
  }
  __syncthreads();
}

// T[b] += B[x][b] * f for the two channels of this thread (weights are scalar, warp-uniform shared-memory broadcasts).
__device__ __forceinline__ void sep_row_fma(float2 (&T)[kPW], const float *__restrict__ bw, float2 f) {
  const float4 w0 = *reinterpret_cast<const float4 *>(bw), w1 = *reinterpret_cast<const float4 *>(bw + 4);
  T[0].x = fmaf(w0.x, f.x, T[0].x); T[0].y = fmaf(w0.x, f.y, T[0].y);
  T[1].x = fmaf(w0.y, f.x, T[1].x); T[1].y = fmaf(w0.y, f.y, T[1].y);
  T[2].x = fmaf(w0.z, f.x, T[2].x); T[2].y = fmaf(w0.z, f.y, T[2].y);
  T[3].x = fmaf(w0.w, f.x, T[3].x); T[3].y = fmaf(w0.w, f.y, T[3].y);
  T[4].x = fmaf(w1.x, f.x, T[4].x); T[4].y = fmaf(w1.x, f.y, T[4].y);
  T[5].x = fmaf(w1.y, f.x, T[5].x); T[5].y = fmaf(w1.y, f.y, T[5].y);
  T[6].x = fmaf(w1.z, f.x, T[6].x); T[6].y = fmaf(w1.z, f.y, T[6].y);
}

// One pass over the rows feeding bins [PH0, PH0 + NPH): accumulates and writes those bin rows of the output tile.
// fbase2: float2 pointer to (image, pixel 0, this thread's channel pair); cs2: channel stride in float2 units.
template <int PH0, int NPH, int kC>
__device__ __forceinline__ void sep_fwd_pass(const SepSmem &s, const float2 *__restrict__ fbase2, int C, int W, int xmin, int xmax,
                                             float *__restrict__ t0, float *__restrict__ t1, int half) {
  const int cs2 = (kC ? kC : C) >> 1;
  int y0 = s.lim[4 + PH0], y1 = s.lim[11 + PH0];
#pragma unroll
  for (int a = 1; a < NPH; ++a) { y0 = min(y0, s.lim[4 + PH0 + a]); y1 = max(y1, s.lim[11 + PH0 + a]); }
  float2 acc[NPH][kPW];
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < kPW; ++b) acc[a][b] = make_float2(0.f, 0.f);
  for (int y = y0; y <= y1; ++y) {
    float2 T[kPW];
#pragma unroll
    for (int b = 0; b < kPW; ++b) T[b] = make_float2(0.f, 0.f);
    const float2 *prow = fbase2 + (size_t)y * W * cs2;
    for (int x0 = xmin; x0 <= xmax; x0 += kXB) {
      const float2 *p = prow + (size_t)x0 * cs2;
      const float *bw = s.Bd + x0 * 8;
      const int nv = xmax - x0 + 1;   // warp-uniform
      float2 f[kXB];
      if (nv >= kXB) {
#pragma unroll
        for (int i = 0; i < kXB; ++i) f[i] = __ldg(p + (size_t)i * cs2);
#pragma unroll
        for (int i = 0; i < kXB; ++i) sep_row_fma(T, bw + 8 * i, f[i]);
      } else {
#pragma unroll
        for (int i = 0; i < kXB - 1; ++i) if (i < nv) f[i] = __ldg(p + (size_t)i * cs2);
#pragma unroll
        for (int i = 0; i < kXB - 1; ++i) if (i < nv) sep_row_fma(T, bw + 8 * i, f[i]);
      }
    }
    const float4 av4 = *reinterpret_cast<const float4 *>(s.Ad + y * 8 + PH0);   // PH0 in {0, 4}: 16-byte aligned
    const float av[4] = {av4.x, av4.y, av4.z, av4.w};
#pragma unroll
    for (int a = 0; a < NPH; ++a) {
      if (av[a] != 0.f) {  // warp-uniform (broadcast shared-memory value)
#pragma unroll
        for (int b = 0; b < kPW; ++b) { acc[a][b].x = fmaf(av[a], T[b].x, acc[a][b].x); acc[a][b].y = fmaf(av[a], T[b].y, acc[a][b].y); }
      }
    }
  }
  // registers -> shared tile (flat [channel][49], the global layout).  Lanes 0-15 write their even channel while lanes
  // 16-31 write their odd channel (and vice versa): word index (2*tid + j)*49 + k hits 32 distinct banks per instruction.
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < kPW; ++b) {
      const float v0 = acc[a][b].x, v1 = acc[a][b].y;
      t0[(PH0 + a) * kPW + b] = half ? v1 : v0;
      t1[(PH0 + a) * kPW + b] = half ? v0 : v1;
    }
}

// The same pass with the column weights in compact form.  With the adaptive sampling grid every bin touches at most
// grid_w + 1 consecutive columns, so instead of 7 (mostly zero) weights per pixel of the ROI's bounding box a row costs
// 7 * NX pixel loads with ONE register-resident weight each: no weight traffic on the shared-memory crossbar and
// NX / (7 ns) of the FMAs.  Summation order per bin (ascending x) is that of the dense pass; the skipped terms are exact
// zeros.  NX = widest bin window of this ROI rounded up to a template instance; narrower bins carry zero weights.
template <int PH0, int NPH, int kC, int NX>
__device__ __forceinline__ void sep_fwd_pass_compact(const SepSmem &s, const float2 *__restrict__ fbase2, int C, int W,
                                                     float *__restrict__ t0, float *__restrict__ t1, int half) {
  const int cs2 = (kC ? kC : C) >> 1;
  int y0 = s.lim[4 + PH0], y1 = s.lim[11 + PH0];
#pragma unroll
  for (int a = 1; a < NPH; ++a) { y0 = min(y0, s.lim[4 + PH0 + a]); y1 = max(y1, s.lim[11 + PH0 + a]); }
  float bw[kPW][NX];
  int xo[kPW];
#pragma unroll
  for (int b = 0; b < kPW; ++b) {
    xo[b] = s.lim[18 + b] * cs2;
#pragma unroll
    for (int j = 0; j < NX; ++j) bw[b][j] = s.Bc[b * kNXMax + j];
  }
  float2 acc[NPH][kPW];
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < kPW; ++b) acc[a][b] = make_float2(0.f, 0.f);
  for (int y = y0; y <= y1; ++y) {
    const float2 *prow = fbase2 + (size_t)y * W * cs2;
    float2 T[kPW];
#pragma unroll
    for (int b = 0; b < kPW; ++b) {
      float2 f[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) f[j] = __ldg(prow + xo[b] + j * cs2);
      T[b] = make_float2(bw[b][0] * f[0].x, bw[b][0] * f[0].y);
#pragma unroll
      for (int j = 1; j < NX; ++j) { T[b].x = fmaf(bw[b][j], f[j].x, T[b].x); T[b].y = fmaf(bw[b][j], f[j].y, T[b].y); }
    }
    const float4 av4 = *reinterpret_cast<const float4 *>(s.Ad + y * 8 + PH0);
    const float av[4] = {av4.x, av4.y, av4.z, av4.w};
#pragma unroll
    for (int a = 0; a < NPH; ++a) {
      if (av[a] != 0.f) {
#pragma unroll
        for (int b = 0; b < kPW; ++b) { acc[a][b].x = fmaf(av[a], T[b].x, acc[a][b].x); acc[a][b].y = fmaf(av[a], T[b].y, acc[a][b].y); }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < kPW; ++b) {
      const float v0 = acc[a][b].x, v1 = acc[a][b].y;
      t0[(PH0 + a) * kPW + b] = half ? v1 : v0;
      t1[(PH0 + a) * kPW + b] = half ? v0 : v1;
    }
}

template <int kC>
__global__ void __launch_bounds__(kSepThreads, 4) roi_align_fwd_sep_kernel(const float *__restrict__ feat /* NHWC */,
                                                                           const float *__restrict__ rois, int N, int C, int H,
                                                                           int W, float scale, int sampling_ratio, int aligned,
                                                                           float *__restrict__ output) {
  extern __shared__ __align__(128) float sep_smem[];
  const SepSmem s = sep_carve(sep_smem, H, W);
  const int r = blockIdx.x, cbase = blockIdx.y * kCT;
  const int tid = threadIdx.x;
  RoiGeom g = roi_geometry(rois + 5 * (size_t)r, scale, aligned, kPH, kPW, sampling_ratio);
  if (g.n < 0 || g.n >= N) { g.gh = 0; g.gw = 0; g.n = 0; }  // invalid batch index -> zeros
  sep_build_tables(s, g, H, W);
  const int ymin = s.lim[0], ymax = s.lim[1], xmin = s.lim[2], xmax = s.lim[3];
  const int c0 = cbase + 2 * tid;
  const bool active = c0 < C;
  if (active) {
    const int half = (tid >> 4) & 1;
    float *t0 = s.tile + (size_t)(2 * tid + half) * kBins;
    float *t1 = s.tile + (size_t)(2 * tid + 1 - half) * kBins;
    if (ymax >= ymin && xmax >= xmin) {
      const float2 *fbase2 = reinterpret_cast<const float2 *>(feat + (size_t)g.n * H * W * C + c0);
      const int nx = s.lim[25];   // CTA-uniform
      if (nx <= 2) {
        sep_fwd_pass_compact<0, 4, kC, 2>(s, fbase2, C, W, t0, t1, half);
        sep_fwd_pass_compact<4, 3, kC, 2>(s, fbase2, C, W, t0, t1, half);
      } else if (nx == 3) {
        sep_fwd_pass_compact<0, 4, kC, 3>(s, fbase2, C, W, t0, t1, half);
        sep_fwd_pass_compact<4, 3, kC, 3>(s, fbase2, C, W, t0, t1, half);
      } else if (nx == 4) {
        sep_fwd_pass_compact<0, 4, kC, 4>(s, fbase2, C, W, t0, t1, half);
        sep_fwd_pass_compact<4, 3, kC, 4>(s, fbase2, C, W, t0, t1, half);
      } else {   // wide bins (fixed sampling_ratio with bins wider than a pixel, or maps narrower than the window): dense tables
        sep_fwd_pass<0, 4, kC>(s, fbase2, C, W, xmin, xmax, t0, t1, half);
        sep_fwd_pass<4, 3, kC>(s, fbase2, C, W, xmin, xmax, t0, t1, half);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kBins; ++k) { t0[k] = 0.f; t1[k] = 0.f; }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    const int nch = min(kCT, C - cbase);
    const unsigned bytes = (unsigned)nch * kBins * 4u;
    float *dst = output + ((size_t)r * C + cbase) * kBins;
    const unsigned src = (unsigned)__cvta_generic_to_shared(s.tile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// Backward of the separable form: gF = A^T . gOut . B per channel pair, accumulated into an NHWC gradient with 8-byte
// vector reductions (red.global.add.v2.f32): one coalesced 256 B reduction per warp and pixel; same two-pass structure.
template <int PH0, int NPH, int kC>
__device__ __forceinline__ void sep_bwd_pass(const SepSmem &s, float2 *__restrict__ gbase2, int C, int W, int xmin, int xmax,
                                             const float *__restrict__ t0, const float *__restrict__ t1, int half) {
  const int cs2 = (kC ? kC : C) >> 1;
  int y0 = s.lim[4 + PH0], y1 = s.lim[11 + PH0];
#pragma unroll
  for (int a = 1; a < NPH; ++a) { y0 = min(y0, s.lim[4 + PH0 + a]); y1 = max(y1, s.lim[11 + PH0 + a]); }
  float2 go[NPH][kPW];
#pragma unroll
  for (int a = 0; a < NPH; ++a)
#pragma unroll
    for (int b = 0; b < kPW; ++b) {
      const float u0 = t0[(PH0 + a) * kPW + b], u1 = t1[(PH0 + a) * kPW + b];
      go[a][b] = half ? make_float2(u1, u0) : make_float2(u0, u1);
    }
  for (int y = y0; y <= y1; ++y) {
    const float4 av4 = *reinterpret_cast<const float4 *>(s.Ad + y * 8 + PH0);
    const float av[4] = {av4.x, av4.y, av4.z, av4.w};
    float2 U[kPW];
#pragma unroll
    for (int b = 0; b < kPW; ++b) U[b] = make_float2(0.f, 0.f);
    bool any = false;
#pragma unroll
    for (int a = 0; a < NPH; ++a) {
      if (av[a] != 0.f) {
        any = true;
#pragma unroll
        for (int b = 0; b < kPW; ++b) { U[b].x = fmaf(av[a], go[a][b].x, U[b].x); U[b].y = fmaf(av[a], go[a][b].y, U[b].y); }
      }
    }
    if (!any) continue;
    float2 *grow = gbase2 + (size_t)y * W * cs2;
    for (int x = xmin; x <= xmax; ++x) {
      const float4 w0 = *reinterpret_cast<const float4 *>(s.Bd + x * 8), w1 = *reinterpret_cast<const float4 *>(s.Bd + x * 8 + 4);
      float2 v = make_float2(w0.x * U[0].x, w0.x * U[0].y);
      v.x = fmaf(w0.y, U[1].x, v.x); v.y = fmaf(w0.y, U[1].y, v.y);
      v.x = fmaf(w0.z, U[2].x, v.x); v.y = fmaf(w0.z, U[2].y, v.y);
      v.x = fmaf(w0.w, U[3].x, v.x); v.y = fmaf(w0.w, U[3].y, v.y);
      v.x = fmaf(w1.x, U[4].x, v.x); v.y = fmaf(w1.x, U[4].y, v.y);
      v.x = fmaf(w1.y, U[5].x, v.x); v.y = fmaf(w1.y, U[5].y, v.y);
      v.x = fmaf(w1.z, U[6].x, v.x); v.y = fmaf(w1.z, U[6].y, v.y);
      atomicAdd(grow + (size_t)x * cs2, v);   // result unused -> RED.E.ADD.F32x2
    }
  }
}

template <int kC>
__global__ void __launch_bounds__(kSepThreads, 4) roi_align_bwd_sep_kernel(const float *__restrict__ grad_out,
                                                                           const float *__restrict__ rois, int N, int C, int H,
                                                                           int W, float scale, int sampling_ratio, int aligned,
                                                                           float *__restrict__ grad_nhwc) {
  extern __shared__ __align__(128) float sep_smem[];
  const SepSmem s = sep_carve(sep_smem, H, W);
  const int r = blockIdx.x, cbase = blockIdx.y * kCT;
  const int tid = threadIdx.x;
  RoiGeom g = roi_geometry(rois + 5 * (size_t)r, scale, aligned, kPH, kPW, sampling_ratio);
  if (g.n < 0 || g.n >= N) return;
  const int nch = min(kCT, C - cbase);
  // coalesced 128-bit staging of the contiguous (nch x 49) gradient tile
  {
    const float4 *src = reinterpret_cast<const float4 *>(grad_out + ((size_t)r * C + cbase) * kBins);
    float4 *dst = reinterpret_cast<float4 *>(s.tile);
    const int n4 = nch * kBins / 4;
    for (int i = tid; i < n4; i += kSepThreads) dst[i] = __ldg(src + i);
  }
  sep_build_tables(s, g, H, W);  // contains the barriers that also publish the tile
  const int ymin = s.lim[0], ymax = s.lim[1], xmin = s.lim[2], xmax = s.lim[3];
  const int c0 = cbase + 2 * tid;
  if (c0 >= C || ymax < ymin || xmax < xmin) return;
  const int half = (tid >> 4) & 1;
  const float *t0 = s.tile + (size_t)(2 * tid + half) * kBins;
  const float *t1 = s.tile + (size_t)(2 * tid + 1 - half) * kBins;
  float2 *gbase2 = reinterpret_cast<float2 *>(grad_nhwc + (size_t)g.n * H * W * C + c0);
  sep_bwd_pass<0, 4, kC>(s, gbase2, C, W, xmin, xmax, t0, t1, half);
  sep_bwd_pass<4, 3, kC>(s, gbase2, C, W, xmin, xmax, t0, t1, half);
}

// --------------------------------------------------------------------------- ROIPool
__global__ void __launch_bounds__(256) roi_pool_fwd_kernel(const float *__restrict__ input, const float *__restrict__ rois, int N,
                                                           int C, int H, int W, long long total, int PH, int PW, float scale,
                                                           float *__restrict__ output, int *__restrict__ argmax) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
  const int c = (int)((idx / ((long long)PW * PH)) % C);
  const long long r = idx / ((long long)PW * PH * C);
  const float *roi = rois + 5 * r;
  const int b = (int)roi[0];
  const int rsw = (int)roundf(__fmul_rn(roi[1], scale)), rsh = (int)roundf(__fmul_rn(roi[2], scale));
  const int rew = (int)roundf(__fmul_rn(roi[3], scale)), reh = (int)roundf(__fmul_rn(roi[4], scale));
  const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
  const float bh = __fdiv_rn((float)rh, (float)PH), bw = __fdiv_rn((float)rw, (float)PW);
  int hs = (int)floorf(__fmul_rn((float)ph, bh)), ws = (int)floorf(__fmul_rn((float)pw, bw));
  int he = (int)ceilf(__fmul_rn((float)(ph + 1), bh)), we = (int)ceilf(__fmul_rn((float)(pw + 1), bw));
  hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
  ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
  const bool empty = (he <= hs) || (we <= ws);
  float maxval = empty ? 0.f : -FLT_MAX;
  int maxidx = -1;
  if (b >= 0 && b < N) {
    const float *in = input + ((size_t)b * C + c) * H * W;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float v = in[h * W + w];
        if (v > maxval) { maxval = v; maxidx = h * W + w; }
      }
  } else if (!empty) maxval = 0.f;
  output[idx] = maxval;
  argmax[idx] = maxidx;
}

__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const float *__restrict__ grad_out, const float *__restrict__ rois,
                                                           const int *__restrict__ argmax, int N, int C, int H, int W,
                                                           long long total, int PH, int PW, float *__restrict__ grad_in) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)((idx / ((long long)PW * PH)) % C);
  const long long r = idx / ((long long)PW * PH * C);
  const int b = (int)rois[5 * r];
  const int am = argmax[idx];
  if (am != -1 && b >= 0 && b < N) atomicAdd(grad_in + ((size_t)b * C + c) * H * W + am, grad_out[idx]);
}

bool sep_supported(int C, int H, int W, int PH, int PW) {
  return PH == kPH && PW == kPW && (C % 4) == 0 && sep_smem_bytes(H, W) <= 100 * 1024;
}

}  // namespace

// =========================================================================== C ABI
SFOD_API int sfod_nchw_to_nhwc(const float *src, float *dst, int N, int C, int HW, sfod_stream_t stream) {
  if (!src || !dst) return SFOD_ERR_INVALID_ARG;
  return launch_transpose(src, dst, N, C, HW, sfod_cu(stream));
}
SFOD_API int sfod_nhwc_to_nchw(const float *src, float *dst, int N, int C, int HW, sfod_stream_t stream) {
  if (!src || !dst) return SFOD_ERR_INVALID_ARG;
  return launch_transpose(src, dst, N, HW, C, sfod_cu(stream));
}

SFOD_API size_t sfod_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int layout, int exact) {
  // exact kernel reads NCHW, separable kernel reads NHWC: a converted copy is needed when layouts differ
  const bool need = exact ? (layout == SFOD_NHWC) : (layout == SFOD_NCHW);
  return need ? sfod_align_up((size_t)N * C * H * W * sizeof(float), 256) : 256;
}

SFOD_API int sfod_roi_align_fwd(const float *input, int layout, const float *rois, int N, int C, int H, int W, int R, int PH,
                                int PW, float spatial_scale, int sampling_ratio, int aligned, int exact, float *output,
                                void *workspace, size_t workspace_bytes, sfod_stream_t stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || PH <= 0 || PW <= 0) return SFOD_ERR_INVALID_ARG;
  if (layout != SFOD_NCHW && layout != SFOD_NHWC) return SFOD_ERR_INVALID_ARG;
  if (R == 0) return SFOD_OK;
  if (!input || !rois || !output) return SFOD_ERR_INVALID_ARG;
  cudaStream_t st = sfod_cu(stream);
  const bool fast = !exact && sep_supported(C, H, W, PH, PW) && sfod_aligned16(output);
  const size_t fbytes = (size_t)N * C * H * W * sizeof(float);
  if (fast) {
    const float *feat = input;
    if (layout == SFOD_NCHW) {
      if (!workspace || workspace_bytes < fbytes) return SFOD_ERR_WORKSPACE_TOO_SMALL;
      int rc = launch_transpose(input, static_cast<float *>(workspace), N, C, H * W, st);
      if (rc) return rc;
      feat = static_cast<const float *>(workspace);
    }
    if (!sfod_aligned16(feat)) return SFOD_ERR_ALIGNMENT;
    const size_t smem = sep_smem_bytes(H, W);
    dim3 grid(R, (C + kCT - 1) / kCT);
#define SFOD_ROI_FWD(KC)                                                                                                     \
    do {                                                                                                                     \
      SFOD_CUDA_TRY(cudaFuncSetAttribute(roi_align_fwd_sep_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      roi_align_fwd_sep_kernel<KC><<<grid, kSepThreads, smem, st>>>(feat, rois, N, C, H, W, spatial_scale, sampling_ratio, aligned, output); \
    } while (0)
    switch (C) {
      case 256: SFOD_ROI_FWD(256); break;
      case 512: SFOD_ROI_FWD(512); break;
      case 1024: SFOD_ROI_FWD(1024); break;
      case 2048: SFOD_ROI_FWD(2048); break;
      default: SFOD_ROI_FWD(0); break;
    }
#undef SFOD_ROI_FWD
    SFOD_LAUNCH_CHECK();
    return SFOD_OK;
  }
  const float *in = input;
  if (layout == SFOD_NHWC) {
    if (!workspace || workspace_bytes < fbytes) return SFOD_ERR_WORKSPACE_TOO_SMALL;
    int rc = launch_transpose(input, static_cast<float *>(workspace), N, H * W, C, st);
    if (rc) return rc;
    in = static_cast<const float *>(workspace);
  }
  const long long total = (long long)R * C * PH * PW;
  roi_align_fwd_exact_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, rois, N, C, H, W, total, PH, PW, spatial_scale,
                                                                             sampling_ratio, aligned, output);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

SFOD_API size_t sfod_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int layout) {
  return layout == SFOD_NCHW ? sfod_align_up((size_t)N * C * H * W * sizeof(float), 256) : 256;
}

SFOD_API int sfod_roi_align_bwd(const float *grad_out, const float *rois, int N, int C, int H, int W, int R, int PH, int PW,
                                float spatial_scale, int sampling_ratio, int aligned, float *grad_in, int layout,
                                void *workspace, size_t workspace_bytes, sfod_stream_t stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || PH <= 0 || PW <= 0 || !grad_in) return SFOD_ERR_INVALID_ARG;
  if (layout != SFOD_NCHW && layout != SFOD_NHWC) return SFOD_ERR_INVALID_ARG;
  cudaStream_t st = sfod_cu(stream);
  const size_t fbytes = (size_t)N * C * H * W * sizeof(float);
  if (R == 0) { SFOD_CUDA_TRY(cudaMemsetAsync(grad_in, 0, fbytes, st)); return SFOD_OK; }
  if (!grad_out || !rois) return SFOD_ERR_INVALID_ARG;
  const bool fast = sep_supported(C, H, W, PH, PW) && sfod_aligned16(grad_out) && sfod_aligned16(grad_in);
  if (fast) {
    float *acc = grad_in;
    if (layout == SFOD_NCHW) {
      if (!workspace || workspace_bytes < fbytes) return SFOD_ERR_WORKSPACE_TOO_SMALL;
      acc = static_cast<float *>(workspace);
    }
    SFOD_CUDA_TRY(cudaMemsetAsync(acc, 0, fbytes, st));
    const size_t smem = sep_smem_bytes(H, W);
    dim3 grid(R, (C + kCT - 1) / kCT);
#define SFOD_ROI_BWD(KC)                                                                                                     \
    do {                                                                                                                     \
      SFOD_CUDA_TRY(cudaFuncSetAttribute(roi_align_bwd_sep_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      roi_align_bwd_sep_kernel<KC><<<grid, kSepThreads, smem, st>>>(grad_out, rois, N, C, H, W, spatial_scale, sampling_ratio, aligned, acc); \
    } while (0)
    switch (C) {
      case 256: SFOD_ROI_BWD(256); break;
      case 512: SFOD_ROI_BWD(512); break;
      case 1024: SFOD_ROI_BWD(1024); break;
      case 2048: SFOD_ROI_BWD(2048); break;
      default: SFOD_ROI_BWD(0); break;
    }
#undef SFOD_ROI_BWD
    SFOD_LAUNCH_CHECK();
    if (layout == SFOD_NCHW) return launch_transpose(acc, grad_in, N, H * W, C, st);
    return SFOD_OK;
  }
  // generic: accumulate in NCHW
  float *acc = grad_in;
  if (layout == SFOD_NHWC) {
    if (!workspace || workspace_bytes < fbytes) return SFOD_ERR_WORKSPACE_TOO_SMALL;
    acc = static_cast<float *>(workspace);
  }
  SFOD_CUDA_TRY(cudaMemsetAsync(acc, 0, fbytes, st));
  const long long total = (long long)R * C * PH * PW;
  roi_align_bwd_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(grad_out, rois, N, C, H, W, total, PH, PW,
                                                                               spatial_scale, sampling_ratio, aligned, acc);
  SFOD_LAUNCH_CHECK();
  if (layout == SFOD_NHWC) return launch_transpose(acc, grad_in, N, C, H * W, st);
  return SFOD_OK;
}

SFOD_API int sfod_roi_pool_fwd(const float *input, const float *rois, int N, int C, int H, int W, int R, int PH, int PW,
                               float spatial_scale, float *output, int32_t *argmax, sfod_stream_t stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || PH <= 0 || PW <= 0) return SFOD_ERR_INVALID_ARG;
  if (R == 0) return SFOD_OK;
  if (!input || !rois || !output || !argmax) return SFOD_ERR_INVALID_ARG;
  const long long total = (long long)R * C * PH * PW;
  roi_pool_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, sfod_cu(stream)>>>(input, rois, N, C, H, W, total, PH, PW,
                                                                                   spatial_scale, output, argmax);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}

SFOD_API int sfod_roi_pool_bwd(const float *grad_out, const float *rois, const int32_t *argmax, int N, int C, int H, int W, int R,
                               int PH, int PW, float *grad_in, sfod_stream_t stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || PH <= 0 || PW <= 0 || !grad_in) return SFOD_ERR_INVALID_ARG;
  cudaStream_t st = sfod_cu(stream);
  SFOD_CUDA_TRY(cudaMemsetAsync(grad_in, 0, (size_t)N * C * H * W * sizeof(float), st));
  if (R == 0) return SFOD_OK;
  if (!grad_out || !rois || !argmax) return SFOD_ERR_INVALID_ARG;
  const long long total = (long long)R * C * PH * PW;
  roi_pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(grad_out, rois, argmax, N, C, H, W, total, PH, PW, grad_in);
  SFOD_LAUNCH_CHECK();
  return SFOD_OK;
}
