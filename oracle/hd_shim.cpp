// TEST INFRASTRUCTURE: host build of the product's __host__ __device__ getImages arithmetic
// (multipathnet_b200/csrc/image_scale.cuh), so the CPU suite can run exactly what get_images_kernel runs per pixel and
// compare it with the independent two-pass restatement orc_image_scale. Built with -ffp-contract=off (the device side
// uses *_rn intrinsics). Not linked into the product.
#include "../multipathnet_b200/csrc/image_scale.cuh"

extern "C" void hd_get_images(const float *im, int H0, int W0, const int *swap /* 1-based */, float scale, const float *mean,
                              const float *std /* or NULL */, int h, int w, float *out) {
  mpn_img::TransformedImage I;
  I.im = im; I.im_u8 = nullptr; I.lut = nullptr; I.H0 = H0; I.W0 = W0;
  for (int c = 0; c < 3; ++c) {
    I.t.src_chan[c] = swap[c] - 1;
    I.t.neg_mean[c] = (float)(-(double)mean[c]);
    I.t.std[c] = std ? std[c] : 1.0f;
  }
  I.t.has_scale = scale != 1.0f;
  I.t.scale = scale;
  I.t.has_std = std != nullptr;
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) out[((long)c * h + y) * w + x] = mpn_img::scaled_pixel(I, h, w, c, y, x);
}

// the uint8 source (H0 x W0 x 3 interleaved bytes), with or without the byte -> float table the device path uses
extern "C" void hd_get_images_u8(const unsigned char *im, int H0, int W0, const int *swap /* 1-based */, float scale, const float *mean,
                                 const float *std /* or NULL */, int h, int w, int use_lut, float *out) {
  static float lut[256];
  for (int b = 0; b < 256; ++b) lut[b] = (float)b / 255.0f;
  mpn_img::TransformedImage I;
  I.im = nullptr; I.im_u8 = im; I.lut = use_lut ? lut : nullptr; I.H0 = H0; I.W0 = W0;
  for (int c = 0; c < 3; ++c) {
    I.t.src_chan[c] = swap[c] - 1;
    I.t.neg_mean[c] = (float)(-(double)mean[c]);
    I.t.std[c] = std ? std[c] : 1.0f;
  }
  I.t.has_scale = scale != 1.0f;
  I.t.scale = scale;
  I.t.has_std = std != nullptr;
  for (int c = 0; c < 3; ++c)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) out[((long)c * h + y) * w + x] = mpn_img::scaled_pixel(I, h, w, c, y, x);
}
