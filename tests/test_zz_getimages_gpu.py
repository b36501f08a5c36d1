"""GPU suite, last file on purpose: getImages on the device (mpn_get_images / mpn_model_trunk_image).

The kernel was written after this round's GPU budget was spent, so it has NOT run on a B200 yet; its per-pixel arithmetic
is the __host__ __device__ code the CPU suite already checks bit for bit (tests/test_getimages_cpu.py), what is untested
is the launch itself. Until a GPU run has confirmed them these tests are xfail(strict=False): a pass shows as XPASS, a
failure as XFAIL, neither hides or breaks the verified suite before it (this file sorts last so that even a faulting
kernel cannot disturb another test). Drop the marker once a round has seen them pass."""
import os

import numpy as np
import pytest

import multipathnet_b200 as mpn
from multipathnet_b200 import models, workloads as wl
from multipathnet_b200.image_detect import ImageDetect
from multipathnet_b200.modules import ImageTransformer

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="first GPU run of get_images_kernel (written after the round's GPU budget was spent)")]


@pytest.mark.parametrize("H0,W0,scale,max_size", [(60, 80, 60, 100), (48, 64, 75, 1000), (120, 90, 60, 1000), (50, 200, 100, 300), (333, 500, 600, 1000)])
@pytest.mark.parametrize("kind", ["ross", "imagenet"])
def test_get_images_matches_the_oracle_bit_for_bit(ctx, oracle_built, H0, W0, scale, max_size, kind):
    im = wl.raw_image(H0, W0, H0 + W0)
    ref, s_ref = oracle_built.get_images(im, kind, scale, max_size)
    out, s = ctx.get_images(im, kind, scale, max_size)
    assert s == s_ref and out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_get_images_golden_fixture(ctx):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "getimages_golden.npz"))
    for name in ("grow", "shrink", "capped", "same"):
        scale, max_size, s, inet = g[name + "_cfg"]
        out, so = ctx.get_images(g[name + "_im"], "imagenet" if inet else "ross", scale, max_size)
        assert so == s and np.array_equal(out, g[name + "_out"])


def test_detect_from_the_raw_image_equals_the_host_getimages_path(ctx):
    spec = models.vgg16_fast_rcnn(21, seed=3, width_div=4, fc_dim=256)
    m = mpn.Model(ctx, spec, max_rois=64, max_h=192, max_w=256)
    im = wl.raw_image(96, 128, 7)
    boxes = wl.random_boxes(32, 96, 128, 7)
    host = ImageDetect(m, ImageTransformer("ross"), scale=[120], max_size=200)
    dev = ImageDetect(m, ImageTransformer("ross"), scale=[120], max_size=200, on_device=True)
    s0, b0 = host.detect(im, boxes)
    s1, b1 = dev.detect(im, boxes)
    assert np.array_equal(s0, s1) and np.array_equal(b0, b1)           # same image bits in, same kernels after
    s2, b2 = dev.detect(None, boxes, recompute_features=False)         # cached features (ImageDetect.lua:109-111)
    assert np.array_equal(s1, s2) and np.array_equal(b1, b2)
    m.close()
