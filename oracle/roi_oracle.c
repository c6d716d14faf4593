/*
 * oracle/roi_oracle.c -- CPU restatement of the reference RoI hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link
 * or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity pin: the reference ships no tests or golden vectors,
 * so this restatement is pinned against the reference's own CPU code compiled
 * in place from /root/reference (oracle/_ref, see oracle/build_oracle.py) and
 * against fixtures generated from it (tests/golden/make_golden.py).
 *
 * Everything is plain C, fp32 arithmetic with the exact operation order of
 * the reference; build with -ffp-contract=off so no multiply-add is fused.
 *
 * Reference files followed (relative to /root/reference/maskrcnn_benchmark):
 *   RoIAlign forward   csrc/cpu/ROIAlign_cpu.cpp:18-217
 *   RoIAlign backward  csrc/cuda/ROIAlign_cuda.cu:126-254 (no CPU version exists)
 *   NMS                csrc/cpu/nms_cpu.cpp:6-65
 *   RoIPool fwd/bwd    csrc/cuda/ROIPool_cuda.cu:17-108  (no CPU version exists)
 *   FPN level mapping  modeling/poolers.py:31-42, structures/bounding_box.py:226-230
 *   box decode / clip  modeling/box_coder.py:52-95, structures/bounding_box.py:214-224
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------ */
/* RoIAlign geometry: one axis of sample positions for one RoI.              */
/* ------------------------------------------------------------------------ */

typedef struct {
  int lo, hi;   /* the two taps along this axis                              */
  float l, h;   /* weight of hi tap (frac) and lo tap (1-frac)               */
  int ok;       /* 0 when the coordinate is outside [-1, extent]             */
} axis_tap;

/* csrc/cpu/ROIAlign_cpu.cpp:36-92, one axis at a time.  The reference tests
 * validity of y and x jointly (:47) and zeroes all four weights; callers do
 * the joint test by and-ing the two `ok` flags. */
static void axis_sample(float start, int p, float bin, int i, int grid,
                        int extent, axis_tap *t) {
  /* start + p*bin + (i+.5f)*bin/grid, left-to-right in fp32 (:36-42) */
  float pos = start + (float)p * bin;
  pos = pos + (float)((float)i + .5f) * bin / (float)grid;
  t->ok = !((double)pos < -1.0 || pos > (float)extent);
  if (pos <= 0.f) pos = 0.f;
  int lo = (int)pos, hi;
  if (lo >= extent - 1) {
    hi = lo = extent - 1;
    pos = (float)lo;
  } else {
    hi = lo + 1;
  }
  t->lo = lo;
  t->hi = hi;
  t->l = pos - (float)lo;
  t->h = (float)(1. - (double)t->l);
}

typedef struct {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
} roi_geom;

/* csrc/cpu/ROIAlign_cpu.cpp:146-169 */
static void roi_geometry(const float *box, float scale, int PH, int PW, int sr,
                         roi_geom *g) {
  float x1 = box[0] * scale, y1 = box[1] * scale;
  float x2 = box[2] * scale, y2 = box[3] * scale;
  float rw = x2 - x1, rh = y2 - y1;
  if (!(rw > 1.f)) rw = 1.f; /* std::max(v, 1): NaN-free inputs assumed */
  if (!(rh > 1.f)) rh = 1.f;
  g->start_w = x1;
  g->start_h = y1;
  g->bin_h = rh / (float)PH;
  g->bin_w = rw / (float)PW;
  g->grid_h = sr > 0 ? sr : (int)ceilf(rh / (float)PH);
  g->grid_w = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
}

/* input [B,C,H,W] contiguous fp32, rois [R,5] (batch,x1,y1,x2,y2),
 * out [R,C,PH,PW].  Returns 0, or -1 on allocation failure. */
int oracle_roi_align_forward(const float *input, int B, int C, int H, int W,
                             const float *rois, int R, float scale, int PH,
                             int PW, int sr, float *out) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int b = (int)roi[0];
    roi_geom g;
    roi_geometry(roi + 1, scale, PH, PW, sr, &g);
    int ny = PH * g.grid_h, nx = PW * g.grid_w;
    axis_tap *ty = (axis_tap *)malloc(sizeof(axis_tap) * (size_t)(ny + nx));
    if (!ty) return -1;
    axis_tap *tx = ty + ny;
    for (int ph = 0; ph < PH; ++ph)
      for (int iy = 0; iy < g.grid_h; ++iy)
        axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, &ty[ph * g.grid_h + iy]);
    for (int pw = 0; pw < PW; ++pw)
      for (int ix = 0; ix < g.grid_w; ++ix)
        axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, &tx[pw * g.grid_w + ix]);
    const float count = (float)(g.grid_h * g.grid_w);
    for (int c = 0; c < C; ++c) {
      const float *plane = input + ((size_t)b * C + c) * H * W;
      float *o = out + ((size_t)n * C + c) * PH * PW;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < g.grid_h; ++iy) {
            const axis_tap *a = &ty[ph * g.grid_h + iy];
            for (int ix = 0; ix < g.grid_w; ++ix) {
              const axis_tap *e = &tx[pw * g.grid_w + ix];
              if (!(a->ok && e->ok)) continue; /* adds +0 in the reference (:47-61) */
              float w1 = a->h * e->h, w2 = a->h * e->l;
              float w3 = a->l * e->h, w4 = a->l * e->l;
              float v1 = plane[a->lo * W + e->lo], v2 = plane[a->lo * W + e->hi];
              float v3 = plane[a->hi * W + e->lo], v4 = plane[a->hi * W + e->hi];
              float s = w1 * v1 + w2 * v2; /* :201-204, left-to-right */
              s = s + w3 * v3;
              s = s + w4 * v4;
              acc += s;
            }
          }
          o[ph * PW + pw] = acc / count;
        }
    }
    free(ty);
  }
  return 0;
}

/* RoIAlign backward, csrc/cuda/ROIAlign_cuda.cu:178-254.  The reference adds
 * with fp32 atomics in undefined order; this restatement produces
 *   grad_in32 : fp32 accumulation in (n,c,ph,pw,iy,ix) order (one legal order)
 *   grad_in64 : the same addends (each rounded to fp32 exactly as the
 *               reference rounds them, :239-242) accumulated in fp64 -- the
 *               order-free value tolerance tests compare against.
 * Either output pointer may be NULL.  Buffers must be zero-initialised
 * (reference: at::zeros, :316). */
int oracle_roi_align_backward(const float *grad_out, const float *rois, int R,
                              float scale, int PH, int PW, int B, int C, int H,
                              int W, int sr, float *grad_in32,
                              double *grad_in64) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int b = (int)roi[0];
    roi_geom g;
    roi_geometry(roi + 1, scale, PH, PW, sr, &g);
    const float count = (float)(g.grid_h * g.grid_w);
    for (int c = 0; c < C; ++c) {
      size_t plane = ((size_t)b * C + c) * H * W;
      const float *go = grad_out + ((size_t)n * C + c) * PH * PW;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          const float top = go[ph * PW + pw];
          for (int iy = 0; iy < g.grid_h; ++iy) {
            axis_tap a;
            axis_sample(g.start_h, ph, g.bin_h, iy, g.grid_h, H, &a);
            for (int ix = 0; ix < g.grid_w; ++ix) {
              axis_tap e;
              axis_sample(g.start_w, pw, g.bin_w, ix, g.grid_w, W, &e);
              if (!(a.ok && e.ok)) continue; /* indices -1 => skipped (:244) */
              float w[4] = {a.h * e.h, a.h * e.l, a.l * e.h, a.l * e.l};
              size_t at[4] = {plane + (size_t)a.lo * W + e.lo, plane + (size_t)a.lo * W + e.hi,
                              plane + (size_t)a.hi * W + e.lo, plane + (size_t)a.hi * W + e.hi};
              for (int k = 0; k < 4; ++k) {
                float gk = top * w[k] / count; /* (top*w)/count, :239-242 */
                if (grad_in32) grad_in32[at[k]] += gk;
                if (grad_in64) grad_in64[at[k]] += (double)gk;
              }
            }
          }
        }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* NMS, csrc/cpu/nms_cpu.cpp:6-65.                                           */
/* ------------------------------------------------------------------------ */

typedef struct {
  float score;
  int64_t idx;
} scored;

/* Descending score; ties -> lower original index first.  The reference's
 * at::sort(descending) leaves tie order unspecified; this is the defined
 * order of the new implementation ("NMS tie-break"). */
static int by_score_desc(const void *pa, const void *pb) {
  const scored *a = (const scored *)pa, *b = (const scored *)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);
}

/* dets [N,4] xyxy, scores [N]; keep[] receives the kept ORIGINAL indices in
 * ascending order (nonzero(suppressed==0), :64).  Returns the count, or -1.
 * `order_in` (optional) supplies the visiting order explicitly (used to
 * replay the reference's own sort when checking tie cases). */
int64_t oracle_nms(const float *dets, const float *scores, int64_t N,
                   float threshold, const int64_t *order_in, int64_t *keep) {
  if (N == 0) return 0;
  float *areas = (float *)malloc(sizeof(float) * (size_t)N);
  scored *ord = (scored *)malloc(sizeof(scored) * (size_t)N);
  uint8_t *dead = (uint8_t *)calloc((size_t)N, 1);
  if (!areas || !ord || !dead) {
    free(areas); free(ord); free(dead);
    return -1;
  }
  for (int64_t i = 0; i < N; ++i) {
    const float *d = dets + i * 4;
    float w = d[2] - d[0]; w = w + 1.f;   /* (x2 - x1 + 1), one rounding per op (:22) */
    float h = d[3] - d[1]; h = h + 1.f;
    areas[i] = w * h;
    ord[i].score = scores[i];
    ord[i].idx = order_in ? order_in[i] : i;
  }
  if (!order_in) qsort(ord, (size_t)N, sizeof(scored), by_score_desc);
  for (int64_t _i = 0; _i < N; ++_i) {
    int64_t i = ord[_i].idx;
    if (dead[i]) continue;
    const float *bi = dets + i * 4;
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < N; ++_j) {
      int64_t j = ord[_j].idx;
      if (dead[j]) continue;
      const float *bj = dets + j * 4;
      float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
      float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
      float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
      float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
      float w = xx2 - xx1; w = w + 1.f;
      float h = yy2 - yy1; h = h + 1.f;
      if (!(w > 0.f)) w = 0.f;
      if (!(h > 0.f)) h = 0.f;
      float inter = w * h;
      float uni = iarea + areas[j];
      uni = uni - inter;
      float ovr = inter / uni;
      if (ovr >= threshold) dead[j] = 1; /* `>=`, :60 (the CUDA file uses `>`) */
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < N; ++i)
    if (!dead[i]) keep[k++] = i;
  free(areas); free(ord); free(dead);
  return k;
}

/* Segmented form used to check the batched kernel: seg_off[S+1], per segment
 * keep indices are LOCAL to the segment, written at keep + seg_off[s], count
 * in keep_cnt[s]; at most max_keep kept per segment if max_keep > 0
 * (structures/boxlist_ops.py:28-29 truncates the ascending list). */
int oracle_nms_batched(const float *dets, const float *scores,
                       const int64_t *seg_off, int64_t S, float threshold,
                       int64_t max_keep, int64_t *keep, int64_t *keep_cnt) {
  for (int64_t s = 0; s < S; ++s) {
    int64_t o = seg_off[s], n = seg_off[s + 1] - o;
    int64_t k = oracle_nms(dets + o * 4, scores + o, n, threshold, NULL, keep + o);
    if (k < 0) return -1;
    if (max_keep > 0 && k > max_keep) k = max_keep;
    keep_cnt[s] = k;
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* FPN level assignment, modeling/poolers.py:31-42.                          */
/* torch evaluates sqrt, /, +, log2, +, floor, clamp one fp32 op at a time.  */
/* ------------------------------------------------------------------------ */
void oracle_level_map(const float *rois5, int64_t R, float k_min, float k_max,
                      float s0, float lvl0, float eps, int32_t *levels) {
  for (int64_t i = 0; i < R; ++i) {
    const float *b = rois5 + i * 5 + 1;
    float w = b[2] - b[0]; w = w + 1.f; /* bounding_box.py:226-230, legacy +1 */
    float h = b[3] - b[1]; h = h + 1.f;
    float s = sqrtf(w * h);
    float q = s / s0;
    q = q + eps;
    float l = log2f(q);
    l = lvl0 + l;
    l = floorf(l);
    if (l < k_min) l = k_min;
    if (l > k_max) l = k_max;
    levels[i] = (int32_t)((int64_t)l - (int64_t)k_min);
  }
}

/* ------------------------------------------------------------------------ */
/* RoIPool (max), csrc/cuda/ROIPool_cuda.cu:17-108.                          */
/* ------------------------------------------------------------------------ */
int oracle_roi_pool_forward(const float *input, int B, int C, int H, int W,
                            const float *rois, int R, float scale, int PH,
                            int PW, float *out, int32_t *argmax) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float *roi = rois + (size_t)n * 5;
    int b = (int)roi[0];
    int sw = (int)roundf(roi[1] * scale), sh = (int)roundf(roi[2] * scale);
    int ew = (int)roundf(roi[3] * scale), eh = (int)roundf(roi[4] * scale);
    int rw = ew - sw + 1, rh = eh - sh + 1;
    if (rw < 1) rw = 1;
    if (rh < 1) rh = 1;
    float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    for (int c = 0; c < C; ++c) {
      const float *plane = input + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
          int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
          hs += sh; he += sh; ws += sw; we += sw;
          hs = hs < 0 ? 0 : (hs > H ? H : hs);
          he = he < 0 ? 0 : (he > H ? H : he);
          ws = ws < 0 ? 0 : (ws > W ? W : ws);
          we = we < 0 ? 0 : (we > W ? W : we);
          int empty = (he <= hs) || (we <= ws);
          float best = empty ? 0.f : -FLT_MAX;
          int besti = -1;
          for (int y = hs; y < he; ++y)
            for (int x = ws; x < we; ++x)
              if (plane[y * W + x] > best) {
                best = plane[y * W + x];
                besti = y * W + x;
              }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          out[o] = best;
          argmax[o] = besti;
        }
    }
  }
  return 0;
}

int oracle_roi_pool_backward(const float *grad_out, const int32_t *argmax,
                             const float *rois, int R, int PH, int PW, int B,
                             int C, int H, int W, float *grad_in) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    int b = (int)rois[(size_t)n * 5];
    for (int c = 0; c < C; ++c) {
      float *plane = grad_in + ((size_t)b * C + c) * H * W;
      size_t o = ((size_t)n * C + c) * PH * PW;
      for (int i = 0; i < PH * PW; ++i)
        if (argmax[o + i] != -1) plane[argmax[o + i]] += grad_out[o + i];
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Box decode + clip, modeling/box_coder.py:52-95 and                         */
/* structures/bounding_box.py:214-224 (legacy +1 / -1 geometry).             */
/* codes [N,4], anchors [N,4] -> boxes [N,4]; img_w/img_h <= 0 skips clip.   */
/* ------------------------------------------------------------------------ */
void oracle_box_decode(const float *codes, const float *anchors, int64_t N,
                       float wx, float wy, float ww, float wh, float clip,
                       float img_w, float img_h, float *boxes) {
  for (int64_t i = 0; i < N; ++i) {
    const float *a = anchors + i * 4, *r = codes + i * 4;
    float w = a[2] - a[0]; w = w + 1.f;
    float h = a[3] - a[1]; h = h + 1.f;
    float hw = 0.5f * w, hh = 0.5f * h;
    float cx = a[0] + hw, cy = a[1] + hh;
    float dx = r[0] / wx, dy = r[1] / wy, dw = r[2] / ww, dh = r[3] / wh;
    if (dw > clip) dw = clip;
    if (dh > clip) dh = clip;
    float px = dx * w; px = px + cx;
    float py = dy * h; py = py + cy;
    float pw = expf(dw) * w, ph = expf(dh) * h;
    float hpw = 0.5f * pw, hph = 0.5f * ph;
    float x1 = px - hpw, y1 = py - hph;
    float x2 = px + hpw; x2 = x2 - 1.f;
    float y2 = py + hph; y2 = y2 - 1.f;
    if (img_w > 0.f && img_h > 0.f) {
      float mx = img_w - 1.f, my = img_h - 1.f;
      x1 = x1 < 0.f ? 0.f : (x1 > mx ? mx : x1);
      y1 = y1 < 0.f ? 0.f : (y1 > my ? my : y1);
      x2 = x2 < 0.f ? 0.f : (x2 > mx ? mx : x2);
      y2 = y2 < 0.f ? 0.f : (y2 > my ? my : y2);
    }
    float *o = boxes + i * 4;
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
  }
}
