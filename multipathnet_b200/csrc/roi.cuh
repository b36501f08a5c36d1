// roi.cuh — job descriptors of the fused Foveal + ROI-pooling kernel (roi.cu), shared with model.cu
#pragma once
#include "common.cuh"
constexpr int ROI_MAX_LEVELS = 6;   // level 0 = the feature map, level k = max over 2^k x 2^k blocks at every position
struct RoiJob {
  const __nv_bfloat16 *hi, *lo;   // feature map planes, NHWC (= level 0)
  const __nv_bfloat16 *hi_lv[ROI_MAX_LEVELS], *lo_lv[ROI_MAX_LEVELS];   // max-pyramid levels (same geometry, ld = C)
  int nlev;                        // number of valid levels incl. level 0
  int H, W, C; long long ld;       // ld = pixel stride (elements)
  float scale;                     // spatial scale
  int region;                      // 0: ROI, 1..3: foveal x1.5, x2, x4
  __nv_bfloat16 *out_hi, *out_lo;  // R x bins x out_ld
  long long out_ld; int out_ch_off;
  int normalize;
};
constexpr int MAX_ROI_JOBS = 16;
struct RoiJobs { RoiJob j[MAX_ROI_JOBS]; int n; };
int mpn_roi_pool_fused_launch(mpn_ctx *ctx, const RoiJobs &jobs, const float *rois_dev, int64_t R, int PW, int PH,
                              int variant);
// builds pyramid level k (block 2^k) from level k-1: out[y][x] = max of the four 2^(k-1) blocks at (y,x),(y,x+s),(y+s,x),(y+s,x+s)
int mpn_maxpyr_launch(mpn_ctx *ctx, const __nv_bfloat16 *ph, const __nv_bfloat16 *pl, int N, int H, int W, int C, long long ld_in,
                      int s, __nv_bfloat16 *oh, __nv_bfloat16 *ol);
int mpn_roi_pool_nchw_launch(mpn_ctx *ctx, const float *fmap_dev, int64_t N, int64_t C, int64_t H, int64_t W,
                             const float *rois_dev, int64_t R, int PW, int PH, float scale, int variant,
                             float *out_dev, int32_t *argmax_dev);
