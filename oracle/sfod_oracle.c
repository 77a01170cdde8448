/*
 * sfod_oracle.c -- plain-C CPU restatement of the native arithmetic on the simple-SFOD
 * pseudo-labelling hot path.  TEST INFRASTRUCTURE ONLY: linked/loaded by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by the product.
 *
 * The reference (EPFL-IMOS/simple-SFOD) has no native code; its arithmetic executes in
 * un-vendored third-party kernels:
 *   - torchvision (0.26.0 installed here; reference README.md:23 pins "pytorch >= 1.12")
 *     csrc/ops/cpu/{nms_kernel,roi_align_kernel,roi_pool_kernel}.cpp
 *   - detectron2 0.6 (README.md:24) box_regression.Box2BoxTransform.apply_deltas
 *   - ATen softmax / batch_norm / elementwise (EMA)
 * Each function below restates the published algorithm of one of those kernels and names
 * the reference call site it serves.  Pinning: tests/test_oracle_cpu.py checks every
 * function against the *installed* torchvision/ATen CPU kernels (bit-exact for nms /
 * roi_align / roi_pool / EMA, <= 1 ulp for exp-based ones) and against tests/golden/.
 *
 * Build (see oracle/Makefile): gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC
 * -ffp-contract=off matters: the parity contract is "separately rounded fp32 operations in
 * source order", which is what torchvision's generic x86-64 build executes.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ NMS
 * torchvision nms_kernel_impl<float> (CPU); reached from reference rpn.py:54-56
 * (find_top_rpn_proposals -> batched_nms) and roi_heads/fast_rcnn.py:133.
 * order = stable sort of scores, descending; suppress j iff ovr > thr (strict);
 * ovr = inter / (iarea + area_j - inter), all fp32, NaN never suppresses. */
typedef struct { float s; int64_t i; } orc_kv;

static void orc_merge_sort_desc(orc_kv *a, orc_kv *tmp, int64_t n) {
  /* bottom-up stable merge sort, descending by s, ties keep lower original position first */
  for (int64_t w = 1; w < n; w *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * w) {
      int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      int64_t p = lo, q = mid, k = lo;
      while (p < mid && q < hi) {
        /* take from right only if strictly greater (NaN treated as largest, like torch) */
        float l = a[p].s, r = a[q].s;
        int r_gt_l = (r != r) ? !(l != l) : (r > l);
        if (r_gt_l) tmp[k++] = a[q++]; else tmp[k++] = a[p++];
      }
      while (p < mid) tmp[k++] = a[p++];
      while (q < hi) tmp[k++] = a[q++];
    }
    memcpy(a, tmp, (size_t)n * sizeof(orc_kv));
  }
}

ORC_API int orc_argsort_desc_stable(const float *scores, int64_t n, int64_t *order) {
  if (n <= 0) return 0;
  orc_kv *a = (orc_kv *)malloc((size_t)n * sizeof(orc_kv));
  orc_kv *t = (orc_kv *)malloc((size_t)n * sizeof(orc_kv));
  if (!a || !t) { free(a); free(t); return -1; }
  for (int64_t i = 0; i < n; ++i) { a[i].s = scores[i]; a[i].i = i; }
  orc_merge_sort_desc(a, t, n);
  for (int64_t i = 0; i < n; ++i) order[i] = a[i].i;
  free(a); free(t);
  return 0;
}

ORC_API int orc_nms(const float *boxes, const float *scores, int64_t n, double iou_threshold,
                    int64_t *keep, int64_t *num_keep) {
  *num_keep = 0;
  if (n <= 0) return 0;
  int64_t *order = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  float *areas = (float *)malloc((size_t)n * sizeof(float));
  uint8_t *suppressed = (uint8_t *)calloc((size_t)n, 1);
  if (!order || !areas || !suppressed) { free(order); free(areas); free(suppressed); return -1; }
  orc_argsort_desc_stable(scores, n, order);
  for (int64_t i = 0; i < n; ++i) {
    float w = boxes[4 * i + 2] - boxes[4 * i + 0];
    float h = boxes[4 * i + 3] - boxes[4 * i + 1];
    areas[i] = w * h;
  }
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (suppressed[i]) continue;
    keep[nk++] = i;
    float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (suppressed[j]) continue;
      float xx1 = ix1 < boxes[4 * j] ? boxes[4 * j] : ix1;          /* std::max(ix1, x1[j]) */
      float yy1 = iy1 < boxes[4 * j + 1] ? boxes[4 * j + 1] : iy1;
      float xx2 = boxes[4 * j + 2] < ix2 ? boxes[4 * j + 2] : ix2;  /* std::min(ix2, x2[j]) */
      float yy2 = boxes[4 * j + 3] < iy2 ? boxes[4 * j + 3] : iy2;
      float dw = xx2 - xx1, dh = yy2 - yy1;
      float w = 0.0f < dw ? dw : 0.0f;                               /* std::max(0, .) */
      float h = 0.0f < dh ? dh : 0.0f;
      float inter = w * h;
      float uni = iarea + areas[j];
      uni = uni - inter;
      float ovr = inter / uni;
      if ((double)ovr > iou_threshold) suppressed[j] = 1;
    }
  }
  *num_keep = nk;
  free(order); free(areas); free(suppressed);
  return 0;
}

/* ------------------------------------------------------------------ ROIAlign
 * torchvision roi_align_forward_kernel_impl<float> + pre_calc_for_bilinear_interpolate (CPU);
 * reached from reference source_free_adaptive_teacher_roi_heads.py:117 via d2 ROIPooler. */
typedef struct { int pos1, pos2, pos3, pos4; float w1, w2, w3, w4; } orc_precalc;

static void orc_roi_geometry(const float *roi, float scale, int aligned, int ph, int pw, int sampling_ratio,
                             float *sh, float *sw, float *bh, float *bw, int *gh, int *gw) {
  float offset = aligned ? 0.5f : 0.0f;
  float roi_start_w = roi[1] * scale - offset;
  float roi_start_h = roi[2] * scale - offset;
  float roi_end_w = roi[3] * scale - offset;
  float roi_end_h = roi[4] * scale - offset;
  float roi_width = roi_end_w - roi_start_w;
  float roi_height = roi_end_h - roi_start_h;
  if (!aligned) {
    roi_width = roi_width > 1.0f ? roi_width : 1.0f;
    roi_height = roi_height > 1.0f ? roi_height : 1.0f;
  }
  *bh = roi_height / (float)ph;
  *bw = roi_width / (float)pw;
  *gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_height / (float)ph);
  *gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_width / (float)pw);
  *sh = roi_start_h; *sw = roi_start_w;
}

static void orc_bilinear(int height, int width, float y, float x, orc_precalc *pc) {
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
    pc->pos1 = pc->pos2 = pc->pos3 = pc->pos4 = -1;
    pc->w1 = pc->w2 = pc->w3 = pc->w4 = 0.0f;
    return;
  }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; } else y_high = y_low + 1;
  if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; } else x_high = x_low + 1;
  float ly = y - (float)y_low, lx = x - (float)x_low;
  float hy = 1.0f - ly, hx = 1.0f - lx;
  pc->w1 = hy * hx; pc->w2 = hy * lx; pc->w3 = ly * hx; pc->w4 = ly * lx;
  pc->pos1 = y_low * width + x_low; pc->pos2 = y_low * width + x_high;
  pc->pos3 = y_high * width + x_low; pc->pos4 = y_high * width + x_high;
}

ORC_API int orc_roi_align_fwd(const float *input, const float *rois, int N, int C, int H, int W, int R,
                              int PH, int PW, float scale, int sampling_ratio, int aligned, float *output) {
  (void)N;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + 5 * n;
    int b = (int)roi[0];
    float sh, sw, bh, bw; int gh, gw;
    orc_roi_geometry(roi, scale, aligned, PH, PW, sampling_ratio, &sh, &sw, &bh, &bw, &gh, &gw);
    int cnt_i = gh * gw; if (cnt_i < 1) cnt_i = 1;
    float count = (float)cnt_i;
    int ghp = gh > 0 ? gh : 0, gwp = gw > 0 ? gw : 0;
    size_t npc = (size_t)ghp * gwp * PH * PW;
    orc_precalc *pc = (orc_precalc *)malloc((npc ? npc : 1) * sizeof(orc_precalc));
    if (!pc) return -1;
    size_t k = 0;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw)
        for (int iy = 0; iy < ghp; ++iy) {
          float yy = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
          for (int ix = 0; ix < gwp; ++ix) {
            float xx = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
            orc_bilinear(H, W, yy, xx, &pc[k++]);
          }
        }
    for (int c = 0; c < C; ++c) {
      const float *in = input + ((size_t)b * C + c) * H * W;
      float *out = output + ((size_t)n * C + c) * PH * PW;
      k = 0;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float v = 0.0f;
          for (int s = 0; s < ghp * gwp; ++s) {
            const orc_precalc *p = &pc[k++];
            if (p->pos1 < 0) continue; /* weights are zero: contributes +0 */
            float t = p->w1 * in[p->pos1] + p->w2 * in[p->pos2];
            t = t + p->w3 * in[p->pos3];
            t = t + p->w4 * in[p->pos4];
            v = v + t;
          }
          out[ph * PW + pw] = v / count;
        }
    }
    free(pc);
  }
  return 0;
}

/* torchvision roi_align_backward_kernel_impl<float> (CPU): autograd of the above, reached
 * from the student's losses.backward() (reference source_free_adaptive_teacher.py:579). */
ORC_API int orc_roi_align_bwd(const float *grad_out, const float *rois, int N, int C, int H, int W, int R,
                              int PH, int PW, float scale, int sampling_ratio, int aligned, float *grad_in) {
  memset(grad_in, 0, (size_t)N * C * H * W * sizeof(float));
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + 5 * n;
    int b = (int)roi[0];
    float sh, sw, bh, bw; int gh, gw;
    orc_roi_geometry(roi, scale, aligned, PH, PW, sampling_ratio, &sh, &sw, &bh, &bw, &gh, &gw);
    float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      float *gi = grad_in + ((size_t)b * C + c) * H * W;
      const float *go = grad_out + ((size_t)n * C + c) * PH * PW;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float g = go[ph * PW + pw];
          for (int iy = 0; iy < gh; ++iy) {
            float y = sh + (float)ph * bh + ((float)iy + 0.5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
              float x = sw + (float)pw * bw + ((float)ix + 0.5f) * bw / (float)gw;
              orc_precalc p;
              orc_bilinear(H, W, y, x, &p);
              if (p.pos1 < 0) continue;
              gi[p.pos1] += g * p.w1 / count;
              gi[p.pos2] += g * p.w2 / count;
              gi[p.pos3] += g * p.w3 / count;
              gi[p.pos4] += g * p.w4 / count;
            }
          }
        }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ ROIPool
 * torchvision roi_pool_forward/backward_kernel_impl<float> (CPU); the "ROIPool" pooler_type of
 * d2 ROIPooler (ctor arg at reference source_free_adaptive_teacher_roi_heads.py:34,46). */
ORC_API int orc_roi_pool_fwd(const float *input, const float *rois, int N, int C, int H, int W, int R,
                             int PH, int PW, float scale, float *output, int32_t *argmax) {
  (void)N;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + 5 * n;
    int b = (int)roi[0];
    int rsw = (int)roundf(roi[1] * scale), rsh = (int)roundf(roi[2] * scale);
    int rew = (int)roundf(roi[3] * scale), reh = (int)roundf(roi[4] * scale);
    int rw = rew - rsw + 1; if (rw < 1) rw = 1;
    int rh = reh - rsh + 1; if (rh < 1) rh = 1;
    float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw) {
        int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
        int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
        hs = hs + rsh; if (hs < 0) hs = 0; if (hs > H) hs = H;
        he = he + rsh; if (he < 0) he = 0; if (he > H) he = H;
        ws = ws + rsw; if (ws < 0) ws = 0; if (ws > W) ws = W;
        we = we + rsw; if (we < 0) we = 0; if (we > W) we = W;
        int empty = (he <= hs) || (we <= ws);
        for (int c = 0; c < C; ++c) {
          const float *in = input + ((size_t)b * C + c) * H * W;
          float maxval = empty ? 0.0f : -FLT_MAX;
          int maxidx = -1;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w)
              if (in[h * W + w] > maxval) { maxval = in[h * W + w]; maxidx = h * W + w; }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          output[o] = maxval; argmax[o] = maxidx;
        }
      }
  }
  return 0;
}

ORC_API int orc_roi_pool_bwd(const float *grad_out, const float *rois, const int32_t *argmax, int N, int C, int H,
                             int W, int R, int PH, int PW, float *grad_in) {
  memset(grad_in, 0, (size_t)N * C * H * W * sizeof(float));
  for (int n = 0; n < R; ++n) {
    int b = (int)rois[5 * n];
    for (int c = 0; c < C; ++c) {
      float *gi = grad_in + ((size_t)b * C + c) * H * W;
      for (int k = 0; k < PH * PW; ++k) {
        size_t o = ((size_t)n * C + c) * PH * PW + k;
        if (argmax[o] != -1) gi[argmax[o]] += grad_out[o];
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ box decode
 * detectron2 0.6 Box2BoxTransform.apply_deltas; RPN weights (1,1,1,1) from
 * reference rpn.py:54, box-head weights (10,10,5,5) from ...roi_heads.py:161.
 * Every op is a separately rounded fp32 op in the order of the Python source; exp is the
 * correctly rounded fp32 exponential (via fp64). deltas (R, 4k) with box r shared by k groups. */
ORC_API int orc_apply_deltas(const float *deltas, const float *boxes, int64_t R, int k, const float *wts,
                             float scale_clamp, float *out) {
  for (int64_t r = 0; r < R; ++r) {
    const float *b = boxes + 4 * r;
    float width = b[2] - b[0], height = b[3] - b[1];
    float hw = 0.5f * width, hh = 0.5f * height;
    float cx = b[0] + hw, cy = b[1] + hh;
    for (int j = 0; j < k; ++j) {
      const float *d = deltas + (r * k + j) * 4;
      float dx = d[0] / wts[0], dy = d[1] / wts[1], dw = d[2] / wts[2], dh = d[3] / wts[3];
      if (dw > scale_clamp) dw = scale_clamp;   /* torch.clamp(max=): NaN propagates */
      if (dh > scale_clamp) dh = scale_clamp;
      float px = dx * width; px = px + cx;
      float py = dy * height; py = py + cy;
      float pw = (float)exp((double)dw) * width;
      float ph = (float)exp((double)dh) * height;
      float hpw = 0.5f * pw, hph = 0.5f * ph;
      float *o = out + (r * k + j) * 4;
      o[0] = px - hpw; o[1] = py - hph; o[2] = px + hpw; o[3] = py + hph;
    }
  }
  return 0;
}

/* softmax over the last dim as defined for the CUDA path (DESIGN.md "softmax"):
 * m = max_k x_k; e_k = fp32(exp_fp64(x_k - m)); s = ((e_0 + e_1) + ...) fp32; p_k = e_k / s.
 * ATen's vectorised softmax differs from this by <= 4e-7 relative (measured). */
ORC_API int orc_softmax(const float *x, int64_t R, int K1, float *out) {
  for (int64_t r = 0; r < R; ++r) {
    const float *xr = x + r * K1; float *o = out + r * K1;
    float m = xr[0];
    for (int k = 1; k < K1; ++k) if (xr[k] > m || xr[k] != xr[k]) m = xr[k];
    float s = 0.0f;
    for (int k = 0; k < K1; ++k) { float d = xr[k] - m; o[k] = (float)exp((double)d); s = s + o[k]; }
    for (int k = 0; k < K1; ++k) o[k] = o[k] / s;
  }
  return 0;
}

/* ------------------------------------------------------------------ EMA
 * reference source_free_adaptive_teacher.py:593-603: new = student*(1-k) + teacher*k with the
 * Python-double scalars rounded to fp32 first (ATen wraps the scalar into the tensor dtype),
 * two separately rounded products and one rounded sum; then load_state_dict copy_()s the fp32
 * result into the teacher tensor (int64 buffers: C-style truncation). */
ORC_API int orc_ema_f32(const float *student, float *teacher, int64_t n, double keep_rate) {
  float a = (float)(1.0 - keep_rate), b = (float)keep_rate;
  for (int64_t i = 0; i < n; ++i) { float p = student[i] * a; float q = teacher[i] * b; teacher[i] = p + q; }
  return 0;
}
ORC_API int orc_ema_i64(const int64_t *student, int64_t *teacher, int64_t n, double keep_rate) {
  float a = (float)(1.0 - keep_rate), b = (float)keep_rate;
  for (int64_t i = 0; i < n; ++i) {
    float p = (float)student[i] * a; float q = (float)teacher[i] * b; float r = p + q;
    teacher[i] = (int64_t)r;
  }
  return 0;
}

/* ------------------------------------------------------------------ BatchNorm statistics (AdaBN)
 * ATen batch_norm_cpu_update_stats (train mode) as driven by reference base.py:290-291:
 * per-channel mean and biased variance over (N, H, W) accumulated in fp64; running stats
 * updated as momentum*stat + (1-momentum)*running in fp64, rounded to fp32 on store;
 * running_var uses the unbiased variance (n/(n-1)). */
ORC_API int orc_bn_stats(const float *x, int N, int C, int64_t HW, float *mean, float *var_biased) {
  for (int c = 0; c < C; ++c) {
    double s = 0.0; double n = (double)N * (double)HW;
    for (int i = 0; i < N; ++i) { const float *p = x + ((size_t)i * C + c) * HW; for (int64_t k = 0; k < HW; ++k) s += p[k]; }
    double m = s / n, v = 0.0;
    for (int i = 0; i < N; ++i) { const float *p = x + ((size_t)i * C + c) * HW; for (int64_t k = 0; k < HW; ++k) { double d = p[k] - m; v += d * d; } }
    mean[c] = (float)m; var_biased[c] = (float)(v / n);
  }
  return 0;
}
ORC_API int orc_bn_update_running(const float *mean, const float *var_biased, int C, double n, double momentum,
                                  float *running_mean, float *running_var) {
  for (int c = 0; c < C; ++c) {
    running_mean[c] = (float)(momentum * (double)mean[c] + (1.0 - momentum) * (double)running_mean[c]);
    double unb = (double)var_biased[c] * n / (n - 1.0);
    running_var[c] = (float)(momentum * unb + (1.0 - momentum) * (double)running_var[c]);
  }
  return 0;
}
ORC_API int orc_bn_apply(const float *x, int N, int C, int64_t HW, const float *mean, const float *var_biased,
                         const float *weight, const float *bias, double eps, float *y) {
  for (int c = 0; c < C; ++c) {
    float invstd = (float)(1.0 / sqrt((double)var_biased[c] + eps));
    float w = weight ? weight[c] : 1.0f, b = bias ? bias[c] : 0.0f;
    for (int i = 0; i < N; ++i) {
      const float *p = x + ((size_t)i * C + c) * HW; float *q = y + ((size_t)i * C + c) * HW;
      for (int64_t k = 0; k < HW; ++k) q[k] = (p[k] - mean[c]) * invstd * w + b;
    }
  }
  return 0;
}

ORC_API int orc_version(void) { return 1; }
