/* ============================================================================
 * oracle/mpn_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, single thread, no FMA contraction: built with
 * -ffp-contract=off) of the non-GEMM arithmetic on the multipathnet detection
 * forward path. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   orc_overlap / orc_nms / orc_bbox_vote : PINNED against the literal reference
 *       nms.c compiled into oracle/_ref/libnms_ref.so, and the reference's IoU
 *       known-answer vector (test.lua:40-52).
 *   orc_convert_from : PINNED by the reference round-trip KAT (test.lua:17-38)
 *       via orc_convert_to.
 *   orc_foveal / orc_context_region / orc_bbox_norm / orc_softmax / orc_clamp :
 *       restated line by line from the cited Lua; no reference test holds values
 *       for them and Torch-7 cannot run here => "parity unpinned" (restatement only).
 *   orc_roi_pool : restates inn.ROIPooling (szagoruyko/imagine-nn, luarocks `inn`,
 *       UNPINNED version, source absent from /root/reference). "parity unpinned":
 *       the reference's only test of it (modules/test.lua:60-83) pins
 *       chunk-invariance, not values. Both known variants are implemented.
 *   orc_image_transform : ImageTransformer.lua:19-33 line by line; unpinned (no reference test).
 *   orc_image_scale / orc_get_images_size : restate `image.scale` ('bilinear') of the
 *       third-party torch `image` package (luarocks scm, UNPINNED, source absent from
 *       /root/reference) from generic/image.c as recalled => "parity unpinned".
 * ==========================================================================*/
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- IoU with the +1 pixel convention: nms.c:14-41, op for op ------------- */
float orc_overlap(const float *a, const float *b) {
  float x1 = a[0] > b[0] ? a[0] : b[0];
  float y1 = a[1] > b[1] ? a[1] : b[1];
  float x2 = a[2] < b[2] ? a[2] : b[2];
  float y2 = a[3] < b[3] ? a[3] : b[3];
  float w = x2 - x1 + 1;
  float h = y2 - y1 + 1;
  float inter = w * h;
  float aarea = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float barea = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  float iou = inter / (aarea + barea - inter);
  return (w <= 0 || h <= 0) ? 0 : iou;
}

/* ---- Greedy NMS returning INDICES in selection order: nms.c:59-108 --------
 * Same pointer-permutation walk as the reference (first strict max in the
 * current order, swap to front, order-preserving survivor compaction with the
 * swap at nms.c:95-97), but tracking row indices instead of row pointers.   */
long orc_nms(const float *sb, long N, float thr, int32_t *keep) {
  if (N <= 0) return 0;
  int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * N);
  for (long i = 0; i < N; ++i) idx[i] = (int32_t)i;
  int32_t *cur = idx;
  long num = N, nkeep = 0;
  while (num) {
    long best = -1; float bestS = -10000000;
    for (long i = 0; i < num; ++i)
      if (sb[5 * cur[i] + 4] > bestS) { bestS = sb[5 * cur[i] + 4]; best = i; }
    if (best < 0) break; /* all scores <= -1e7 or NaN: the reference would read boxes[-1] (UB) */
    int32_t b = cur[best]; cur[best] = cur[0]; cur[0] = b;
    cur++; keep[nkeep++] = b;
    long good = 0;
    for (long i = 0; i < num - 1; ++i) {
      float ov = orc_overlap(sb + 5 * b, sb + 5 * cur[i]);
      if (ov <= thr) { int32_t t = cur[good]; cur[good++] = cur[i]; cur[i] = t; }
    }
    num = good;
  }
  free(idx);
  return nkeep;
}

/* ---- box voting: nms.c:110-142 -------------------------------------------- */
void orc_bbox_vote(const float *nms_boxes, long K, const float *sb, long N, float thr, float *res) {
  for (long i = 0; i < K; ++i) {
    float acc[5] = {0, 0, 0, 0, 0};
    for (long j = 0; j < N; ++j) {
      float ov = orc_overlap(sb + 5 * j, nms_boxes + 5 * i);
      if (ov > thr) {
        for (int f = 0; f < 4; ++f) acc[f] += sb[5 * j + f] * sb[5 * j + 4];
        acc[4] += sb[5 * j + 4];
      }
    }
    for (int f = 0; f < 4; ++f) res[5 * i + f] = acc[f] / acc[4];
    res[5 * i + 4] = nms_boxes[5 * i + 4];
  }
}

/* ---- index-returning NMS used by demo.lua: utils.lua:402-462 ---------------
 * torch.sort descending is not stable; ties are ordered by ascending index
 * here (documented decision). IoU uses clamp(w,0), clamp(h,0) and the union
 * order (area_j - inter) + area_c of utils.lua:452. Suppress iff ol > thr. */
typedef struct { float s; int32_t i; } orc_si;
static int orc_cmp_desc(const void *a, const void *b) {
  const orc_si *x = (const orc_si *)a, *y = (const orc_si *)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}
long orc_nms_dense(const float *sb, long N, float thr, int32_t *pick) {
  if (N <= 0) return 0;
  orc_si *o = (orc_si *)malloc(sizeof(orc_si) * N);
  uint8_t *sup = (uint8_t *)calloc(N, 1);
  float *area = (float *)malloc(sizeof(float) * N);
  for (long i = 0; i < N; ++i) { o[i].s = sb[5 * i + 4]; o[i].i = (int32_t)i; }
  qsort(o, N, sizeof(orc_si), orc_cmp_desc);
  for (long c = 0; c < N; ++c) {
    const float *b = sb + 5 * o[c].i;
    area[c] = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  }
  long np = 0;
  for (long c = 0; c < N; ++c) {
    if (sup[c]) continue;
    pick[np++] = o[c].i;
    const float *bc = sb + 5 * o[c].i;
    for (long j = 0; j < N; ++j) {
      const float *bj = sb + 5 * o[j].i;
      float xx1 = bj[0] < bc[0] ? bc[0] : bj[0];   /* clamp(x1[c], inf) */
      float yy1 = bj[1] < bc[1] ? bc[1] : bj[1];
      float xx2 = bj[2] < 0 ? 0 : (bj[2] > bc[2] ? bc[2] : bj[2]); /* clamp(0, x2[c]) */
      float yy2 = bj[3] < 0 ? 0 : (bj[3] > bc[3] ? bc[3] : bj[3]);
      float w = xx2 - xx1 + 1; if (w < 0) w = 0;
      float h = yy2 - yy1 + 1; if (h < 0) h = 0;
      float inter = w * h;
      float uni = (area[j] - inter) + area[c];
      float ol = inter / uni;
      if (ol > thr) sup[j] = 1;
    }
  }
  free(o); free(sup); free(area);
  return np;
}

/* ---- nn.Foveal: modules/Foveal.lua:15-44 ---------------------------------- *
 * Lua numbers are doubles; the FloatTensor constructor rounds once to fp32.  */
void orc_foveal(const float *rois, long R, float *out /* 4R x 5 */) {
  static const double off[4] = {0.0, 0.25, 0.5, 1.5};
  static const double mul[4] = {1.0, 1.5, 2.0, 4.0};
  for (long i = 0; i < R; ++i) {
    const float *b = rois + 5 * i;
    double id = b[0], x = b[1], y = b[2], x2 = b[3], y2 = b[4];
    double w = x2 - x, h = y2 - y;
    float *o = out + 20 * i;
    memcpy(o, b, sizeof(float) * 5);                    /* base[1]:copy(box) */
    for (int k = 1; k < 4; ++k) {
      double rx = x - w * off[k], ry = y - h * off[k];
      double rw = w * mul[k], rh = h * mul[k];
      o[5 * k + 0] = (float)id;
      o[5 * k + 1] = (float)rx;
      o[5 * k + 2] = (float)ry;
      o[5 * k + 3] = (float)(rx + rw);
      o[5 * k + 4] = (float)(ry + rh);
    }
  }
}

/* ---- nn.ContextRegion: modules/ContextRegion.lua:14-32 --------------------- *
 * a,b are computed in double (Lua) and stored in the tensor dtype (fp32 here);
 * the mm row is out[k] = sum_j in[j]*tr[j][k] accumulated in fp32 over j=1..4
 * including the exact-zero terms (which cannot change a finite fp32 sum).    */
void orc_context_region(const float *rois, long R, float scale, float *out) {
  float a = (float)((1.0 + (double)scale) / 2.0);
  float b = (float)((1.0 - (double)scale) / 2.0);
  for (long i = 0; i < R; ++i) {
    const float *r = rois + 5 * i; float *o = out + 5 * i;
    o[0] = r[0];
    o[1] = r[1] * a + r[3] * b;
    o[2] = r[2] * a + r[4] * b;
    o[3] = r[1] * b + r[3] * a;
    o[4] = r[2] * b + r[4] * a;
  }
}

/* ---- nn.BBoxNorm (eval): modules/BBoxNorm.lua:18-32 ------------------------ */
void orc_bbox_norm(float *d, long R, long C4 /* = 4C */, const float *mean4, const float *std4) {
  for (long i = 0; i < R * C4; ++i) d[i] = d[i] * std4[i & 3] + mean4[i & 3];
}

/* ---- utils.convertFrom, tensor branch: utils.lua:226-246 -------------------- *
 * Applied per class block of 4 with the ORIGINAL boxes (ImageDetect.lua:183-185).
 * torch.addcmul(xc, y, w) = xc + y*w evaluated unfused in fp32.              */
void orc_convert_from(const float *deltas, const float *boxes, long R, long C, float *out) {
  for (long i = 0; i < R; ++i) {
    const float *b = boxes + 4 * i;
    float xc = (b[0] + b[2]) * 0.5f, yc = (b[1] + b[3]) * 0.5f;
    float w = b[2] - b[0], h = b[3] - b[1];
    for (long c = 0; c < C; ++c) {
      const float *y = deltas + (i * C + c) * 4; float *o = out + (i * C + c) * 4;
      float t0 = y[0] * w, t1 = y[1] * h;
      float xtc = xc + t0, ytc = yc + t1;
      float wt = expf(y[2]) * w, ht = expf(y[3]) * h;
      float hw = wt * 0.5f, hh = ht * 0.5f;
      o[0] = xtc - hw; o[1] = ytc - hh; o[2] = xtc + hw; o[3] = ytc + hh;
    }
  }
}
/* utils.convertTo (utils.lua:180-199) in double, used only for the round-trip KAT */
void orc_convert_to_f64(const double *bbox, const double *tbox, double *out) {
  double xc = (bbox[0] + bbox[2]) * 0.5, yc = (bbox[1] + bbox[3]) * 0.5;
  double w = bbox[2] - bbox[0], h = bbox[3] - bbox[1];
  double xtc = (tbox[0] + tbox[2]) * 0.5, ytc = (tbox[1] + tbox[3]) * 0.5;
  double wt = tbox[2] - tbox[0], ht = tbox[3] - tbox[1];
  out[0] = (xtc - xc) / w; out[1] = (ytc - yc) / h; out[2] = log(wt / w); out[3] = log(ht / h);
}
void orc_convert_from_f64(const double *bbox, const double *y, double *out) {
  double xc = (bbox[0] + bbox[2]) * 0.5, yc = (bbox[1] + bbox[3]) * 0.5;
  double w = bbox[2] - bbox[0], h = bbox[3] - bbox[1];
  double xtc = xc + y[0] * w, ytc = yc + y[1] * h;
  double wt = w * exp(y[2]), ht = h * exp(y[3]);
  out[0] = xtc - wt / 2; out[1] = ytc - ht / 2; out[2] = xtc + wt / 2; out[3] = ytc + ht / 2;
}

/* ---- Tester_FRCNN.lua:75-78: clamp x to [1,W0], y to [1,H0] in place -------- */
void orc_clamp_boxes(float *boxes, long n_boxes /* R*C */, float W0, float H0) {
  for (long i = 0; i < n_boxes * 2; ++i) {
    float *p = boxes + 2 * i;
    p[0] = p[0] < 1 ? 1 : (p[0] > W0 ? W0 : p[0]);
    p[1] = p[1] < 1 ? 1 : (p[1] > H0 ? H0 : p[1]);
  }
}

/* ---- nn.SoftMax over dim 2 (ImageDetect.lua:189-191): max-shifted ---------- */
void orc_softmax(const float *x, long R, long C, float *out) {
  for (long i = 0; i < R; ++i) {
    const float *r = x + i * C; float *o = out + i * C;
    float m = r[0]; for (long c = 1; c < C; ++c) if (r[c] > m) m = r[c];
    float s = 0; for (long c = 0; c < C; ++c) { o[c] = expf(r[c] - m); s += o[c]; }
    for (long c = 0; c < C; ++c) o[c] = o[c] / s;
  }
}

/* ---- ImageDetect.lua:66-70 project_im_rois (single scale) ------------------ */
void orc_project_rois(const float *boxes, long R, float im_scale, float *rois) {
  for (long i = 0; i < R; ++i) {
    rois[5 * i] = 1.0f;
    for (int k = 0; k < 4; ++k) rois[5 * i + 1 + k] = (boxes[4 * i + k] - 1.0f) * im_scale + 1.0f;
  }
}

/* ---- nn.Normalize(2) per row (model_utils.lua:218): x / sqrt(sum x^2 + 1e-10) */
void orc_l2_normalize(float *x, long rows, long n) {
  for (long r = 0; r < rows; ++r) {
    float *p = x + r * n; double s = 0; /* TH accumulates sums in accreal=double for float */
    for (long i = 0; i < n; ++i) s += (double)(p[i] * p[i]);
    float nrm = sqrtf((float)s + 1e-10f);
    for (long i = 0; i < n; ++i) p[i] = p[i] / nrm;
  }
}

/* ---- inn.ROIPooling forward (imagine-nn, see header) ------------------------
 * fmap N x C x H x W (NCHW), rois R x 5 [batch_idx(1-based), x1, y1, x2, y2]
 * in scaled-image 1-based pixels, out R x C x PH x PW, argmax (may be NULL).
 * variant 1: end = round((x2-1)*s);  variant 2 (default): end = round((x2-1)*s) - 1.
 * len = max(end - start + 1, 1); bins floor/ceil, clipped to the map; an empty
 * bin yields 0 / argmax -1; strict '>' from -FLT_MAX so the first max wins.  */
void orc_roi_pool(const float *fmap, long N, long C, long H, long W,
                  const float *rois, long R, int PW, int PH, float scale, int variant,
                  float *out, int32_t *argmax) {
  (void)N;
  for (long r = 0; r < R; ++r) {
    const float *roi = rois + 5 * r;
    long n = (long)roi[0] - 1;
    int sw = (int)roundf((roi[1] - 1) * scale);
    int sh = (int)roundf((roi[2] - 1) * scale);
    int ew = (int)roundf((roi[3] - 1) * scale);
    int eh = (int)roundf((roi[4] - 1) * scale);
    if (variant == 2) { ew -= 1; eh -= 1; }
    int rw = ew - sw + 1; if (rw < 1) rw = 1;
    int rh = eh - sh + 1; if (rh < 1) rh = 1;
    float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    for (long c = 0; c < C; ++c) {
      const float *plane = fmap + (n * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph) for (int pw = 0; pw < PW; ++pw) {
        int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
        int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
        hs += sh; he += sh; ws += sw; we += sw;
        hs = hs < 0 ? 0 : (hs > H ? (int)H : hs); he = he < 0 ? 0 : (he > H ? (int)H : he);
        ws = ws < 0 ? 0 : (ws > W ? (int)W : ws); we = we < 0 ? 0 : (we > W ? (int)W : we);
        int empty = (he <= hs) || (we <= ws);
        float m = empty ? 0.0f : -FLT_MAX; int32_t mi = -1;
        for (int hh = hs; hh < he; ++hh) for (int ww = ws; ww < we; ++ww) {
          float v = plane[hh * W + ww];
          if (v > m) { m = v; mi = (int32_t)(hh * W + ww); }
        }
        long o = ((r * C + c) * PH + ph) * PW + pw;
        out[o] = m; if (argmax) argmax[o] = mi;
      }
    }
  }
}

/* ---- max-pool k x k stride s, ceil_mode flag (Caffe-converted VGG: ceil) ---- *
 * torch nn.SpatialMaxPooling: out = ceil|floor((in + 2p - k)/s) + 1, and with
 * ceil mode the last window must start inside the (left-padded) input.       */
long orc_pool_out(long in, int k, int s, int p, int ceil_mode) {
  long o = ceil_mode ? (long)ceilf((float)(in + 2 * p - k) / s) + 1 : (long)floorf((float)(in + 2 * p - k) / s) + 1;
  if (ceil_mode && (o - 1) * s >= in + p) --o;
  return o;
}


/* ---- getImages ------------------------------------------------
 * ImageTransformer:updateOutput, modules/ImageTransformer.lua:19-33: index by swap (1-based), mul(scale) when
 * scale ~= 1, add(-mean[i]), div(std[i]) when std — each a separate fp32 tensor op. */
void orc_image_transform(const float *im, long H, long W, const int *swap /* 1-based, or NULL */, float scale,
                         const float *mean, const float *std /* or NULL */, float *out) {
  long n = H * W;
  for (int c = 0; c < 3; ++c) {
    const float *src = im + (long)(swap ? swap[c] - 1 : c) * n;
    float *dst = out + (long)c * n;
    for (long i = 0; i < n; ++i) dst[i] = src[i];
  }
  if (scale != 1.0f)
    for (long i = 0; i < 3 * n; ++i) out[i] = out[i] * scale;
  for (int c = 0; c < 3; ++c) {
    float *dst = out + (long)c * n;
    float nm = (float)(-(double)mean[c]);
    for (long i = 0; i < n; ++i) dst[i] = dst[i] + nm;
    if (std)
      for (long i = 0; i < n; ++i) dst[i] = dst[i] / std[c];
  }
}

/* ImageDetect.lua:31-39 (single scale): im_scale and the size image.scale allocates (Lua numbers truncated to long) */
void orc_get_images_size(long H0, long W0, double scale, double max_size, long *h, long *w, double *im_scale) {
  double smin = H0 < W0 ? (double)H0 : (double)W0, smax = H0 < W0 ? (double)W0 : (double)H0;
  double s = scale / smin;
  if (floor(s * smax + 0.5) > max_size) s = max_size / smax;      /* torch.round */
  *h = (long)((double)H0 * s);
  *w = (long)((double)W0 * s);
  *im_scale = s;
}

/* one row or column, torch image generic/image.c scaleLinear_rowcol as recalled: strided in, strided out, the
 * shrinking branch carries its window state from one output sample to the next */
static void orc_scale_rowcol(const float *src, long src_stride, long src_len, float *dst, long dst_stride, long dst_len) {
  if (dst_len > src_len) {
    if (src_len == 1) {
      for (long di = 0; di < dst_len; ++di) dst[di * dst_stride] = src[0];
      return;
    }
    float scale = (float)(src_len - 1) / (float)(dst_len - 1);
    for (long di = 0; di < dst_len - 1; ++di) {
      float si_f = (float)di * scale;
      long si_i = (long)si_f;
      si_f -= (float)si_i;
      dst[di * dst_stride] = (1 - si_f) * src[si_i * src_stride] + si_f * src[(si_i + 1) * src_stride];
    }
    dst[(dst_len - 1) * dst_stride] = src[(src_len - 1) * src_stride];
  } else if (dst_len < src_len) {
    long si0_i = 0;
    float si0_f = 0;
    float scale = (float)src_len / (float)dst_len;
    for (long di = 0; di < dst_len; ++di) {
      float si1_f = (float)(di + 1) * scale;
      long si1_i = (long)si1_f;
      si1_f -= (float)si1_i;
      float acc = (1 - si0_f) * src[si0_i * src_stride];
      float n = 1 - si0_f;
      for (long si = si0_i + 1; si < si1_i; ++si) {
        acc += src[si * src_stride];
        n += 1;
      }
      if (si1_i < src_len) {
        acc += si1_f * src[si1_i * src_stride];
        n += si1_f;
      }
      dst[di * dst_stride] = acc / n;
      si0_i = si1_i;
      si0_f = si1_f;
    }
  } else {
    for (long di = 0; di < dst_len; ++di) dst[di * dst_stride] = src[di * src_stride];
  }
}

/* image.scale(src, w, h) 'bilinear' (scaleBilinear): rows to the new width into a temporary, then its columns */
void orc_image_scale(const float *src, long C, long H, long W, float *dst, long h, long w) {
  float *tmp = (float *)malloc(sizeof(float) * (size_t)(H * w > 0 ? H * w : 1));
  for (long c = 0; c < C; ++c) {
    for (long j = 0; j < H; ++j) orc_scale_rowcol(src + (c * H + j) * W, 1, W, tmp + j * w, 1, w);
    for (long i = 0; i < w; ++i) orc_scale_rowcol(tmp + i, w, H, dst + c * h * w + i, w, h);
  }
  free(tmp);
}
