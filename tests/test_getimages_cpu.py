"""CPU suite, part 7: getImages — ImageTransformer + image.scale + the im_scale rule of ImageDetect.lua:22-52.
Three statements of the same arithmetic must agree BIT FOR BIT: the two-pass C restatement of image.scale (oracle,
carried window state like the original), the numpy host mirror ImageDetect.getImages uses, and the product's per-pixel
__host__ __device__ code (csrc/image_scale.cuh, the body of get_images_kernel) built for the host by oracle/hd_shim.cpp.
PARITY UNPINNED against Torch's `image` package itself (third-party, absent): hand-computed known answers below pin the
recalled algorithm (corner-aligned interpolation when growing, area average when shrinking, rows first)."""
import ctypes as C
import os

import numpy as np
import pytest

import multipathnet_b200 as mpn
from multipathnet_b200 import workloads as wl
from multipathnet_b200.image_detect import ImageDetect, _get_images_size, _image_scale, _scale_axis
from multipathnet_b200.modules import ImageTransformer


def test_scale_axis_known_answers():
    a = np.array([[1.0, 3.0, 7.0]], np.float32)
    assert np.array_equal(_scale_axis(a, 5), np.array([[1, 2, 3, 5, 7]], np.float32))            # corners aligned, midpoints between
    assert np.array_equal(_scale_axis(a, 3), a) and _scale_axis(a, 3) is not a
    b = np.array([[2.0, 4.0, 6.0, 8.0, 10.0]], np.float32)
    assert np.array_equal(_scale_axis(b[:, :4], 2), np.array([[3, 7]], np.float32))               # 4 -> 2: pair averages
    # 5 -> 2: windows [0, 2.5) and [2.5, 5): (2 + 4 + 0.5*6) / 2.5 and (0.5*6 + 8 + 10) / 2.5
    assert np.allclose(_scale_axis(b, 2), np.array([[9 / 2.5, 21 / 2.5]], np.float32), rtol=1e-6)
    assert np.array_equal(_scale_axis(np.array([[5.0]], np.float32), 4), np.full((1, 4), 5, np.float32))


@pytest.mark.parametrize("H0,W0,h,w", [(20, 30, 33, 50), (40, 56, 17, 24), (24, 32, 24, 32), (30, 20, 45, 20), (31, 47, 30, 61),
                                        (7, 9, 70, 90), (64, 48, 5, 4), (1, 1, 3, 2)])
@pytest.mark.parametrize("kind", ["ross", "imagenet"])
def test_three_statements_agree_bit_for_bit(oracle_built, H0, W0, h, w, kind):
    O = oracle_built
    im = wl.raw_image(H0, W0, H0 * 100 + W0)
    t_orc = O.image_transform(im, kind)
    assert np.array_equal(t_orc, wl.transform(im, kind)) and np.array_equal(t_orc, ImageTransformer(kind).forward(im))
    two_pass = O.image_scale(t_orc, h, w)
    assert np.array_equal(two_pass, _image_scale(t_orc, h, w))                                    # numpy host mirror
    assert np.array_equal(two_pass, O.hd_get_images(im, kind, h, w))                              # the kernel's per-pixel code
    assert two_pass.shape == (3, h, w) and np.isfinite(two_pass).all()
    lo, hi = t_orc.min(), t_orc.max()
    assert two_pass.min() >= lo - 1e-3 * abs(lo) - 1e-5 and two_pass.max() <= hi + 1e-3 * abs(hi) + 1e-5    # convex combinations


def test_size_rule_python_c_abi_and_oracle_agree(oracle_built):
    lib = mpn.load_library()                                       # host-only entry: no GPU needed
    rng = np.random.default_rng(0)
    cases = [(600, 800, 600, 1000), (480, 640, 600, 1000), (375, 500, 600, 1000), (333, 500, 600, 1000), (500, 333, 600, 1000),
             (300, 1000, 600, 1000), (427, 640, 800, 1000), (100, 1234, 600, 1000), (1200, 1600, 600, 1000), (601, 1001, 600, 1000)]
    cases += [(int(a), int(b), 600, 1000) for a, b in rng.integers(50, 1500, (300, 2))]
    cases += [(int(a), int(b), 800, 1000) for a, b in rng.integers(50, 1500, (100, 2))]
    for H0, W0, scale, max_size in cases:
        h, w, s = C.c_int32(), C.c_int32(), C.c_double()
        assert lib.mpn_get_images_size(H0, W0, float(scale), float(max_size), C.byref(h), C.byref(w), C.byref(s)) == 0
        assert (h.value, w.value, s.value) == _get_images_size(H0, W0, scale, max_size) == oracle_built.get_images_size(H0, W0, scale, max_size)
        assert max(h.value, w.value) <= max_size and (min(h.value, w.value) in (scale - 1, scale) or max(h.value, w.value) >= max_size - 1)
    assert _get_images_size(600, 800, 600, 1000) == (600, 800, 1.0)
    assert _get_images_size(300, 1000, 600, 1000)[2] == 1.0                                       # capped by max_size
    assert lib.mpn_get_images_size(0, 10, 600.0, 1000.0, None, None, None) != 0


def test_benchmark_configs_are_the_identity_resize(oracle_built):
    """BASELINE configs 2/3 hand detect() 600 x 800 images at scale 600: image.scale is a copy, im_scale == 1"""
    im = wl.raw_image(60, 80, 3)
    out, s = oracle_built.get_images(im, "ross", 60, 100)
    assert s == 1.0 and np.array_equal(out, wl.transform(im, "ross"))
    assert np.array_equal(oracle_built.hd_get_images(im, "ross", 60, 80), out)


class _FakeModel:
    def __init__(self):
        self.calls = []

    def trunk_image(self, im, kind, scale, max_size):
        self.calls.append(("trunk_image", im.shape, kind, scale, max_size))
        return 1.25, 75, 100

    def detect(self, img, boxes, im_scale, recompute):
        self.calls.append(("detect", None if img is None else img.shape, float(im_scale), bool(recompute)))
        return np.zeros((boxes.shape[0], 3), np.float32), np.zeros((boxes.shape[0], 12), np.float32)


def test_image_detect_host_and_device_getimages_paths():
    im = wl.raw_image(60, 80, 1)
    boxes = np.array([[1, 1, 20, 20]], np.float32)
    m = _FakeModel()
    d = ImageDetect(m, ImageTransformer("ross"), scale=[75], max_size=100)
    img, s = d.getImages(im)
    assert img.shape == (3, 75, 100) and s == 1.25
    d.detect(im, boxes)
    assert m.calls == [("detect", (3, 75, 100), 1.25, True)]
    m.calls.clear()
    d2 = ImageDetect(m, ImageTransformer("ross"), scale=[75], max_size=100, on_device=True)
    d2.detect(im, boxes)
    d2.detect(None, boxes, recompute_features=False)
    assert m.calls == [("trunk_image", (3, 60, 80), "ross", 75, 100), ("detect", None, 1.25, False), ("detect", None, 1.25, False)]


def test_getimages_golden_fixture(oracle_built):
    """committed outputs (tests/golden/make_golden.py): any later change to the restatement or the mirrors is caught"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "getimages_golden.npz"))
    for name in ("grow", "shrink", "capped", "same"):
        im, out = g[name + "_im"], g[name + "_out"]
        scale, max_size, s, inet = g[name + "_cfg"]
        kind = "imagenet" if inet else "ross"
        o, so = oracle_built.get_images(im, kind, scale, max_size)
        assert so == s and np.array_equal(o, out)
        d = ImageDetect(_FakeModel(), ImageTransformer(kind), scale=[scale], max_size=max_size)
        img, sm = d.getImages(im)
        assert sm == s and np.array_equal(img, out)
        assert np.array_equal(oracle_built.hd_get_images(im, kind, out.shape[1], out.shape[2]), out)
    assert g["capped_out"].shape[2] == 90 and g["same_cfg"][2] == 1.0


@pytest.mark.parametrize("kind", ["ross", "imagenet"])
@pytest.mark.parametrize("H0,W0,scale,max_size", [(48, 64, 60, 100), (75, 50, 60, 70), (40, 40, 40, 100)])
def test_u8_source_table_equals_the_division(oracle_built, H0, W0, scale, max_size, kind):
    """the uint8 path's byte -> float table (what the kernel reads since round 2) holds the same correctly rounded quotients as the
    per-sample IEEE division it replaces, and both equal the fp32 path fed with byte / 255: bit for bit, growing and shrinking"""
    O = oracle_built
    rng = np.random.default_rng(H0 + W0)
    im_u8 = rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8)
    im_u8[:3, :5] = 0; im_u8[-2:, -4:] = 255
    h, w, _ = O.get_images_size(H0, W0, scale, max_size)
    a = O.hd_get_images_u8(im_u8, kind, h, w, use_lut=True)
    b = O.hd_get_images_u8(im_u8, kind, h, w, use_lut=False)
    assert np.array_equal(a, b)
    im_f = np.ascontiguousarray(im_u8.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
    assert np.array_equal(a, O.hd_get_images(im_f, kind, h, w))
