"""GPU parity: region modules, BBoxNorm, decode, inn.ROIPooling-compatible op through the C ABI vs the oracle."""
import os

import numpy as np
import pytest

from multipathnet_b200 import modules, utils as U, workloads as wl
from oracle import ref as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rois(n, seed):
    return np.concatenate([np.ones((n, 1), np.float32), wl.random_boxes(n, 600, 800, seed)], 1)


def test_foveal_bit_exact(ctx):
    r = _rois(1000, 1)
    assert np.array_equal(ctx.foveal(r), O.foveal(r))
    g = np.load(os.path.join(GOLD, "ops_golden.npz"))
    assert np.array_equal(modules.Foveal(ctx).forward(g["rois"]), g["foveal"])
    with pytest.raises(ValueError):
        modules.Foveal(ctx).forward(np.zeros((3, 4), np.float32))        # Foveal.lua:17 assert


def test_context_region_bit_exact(ctx):
    r = _rois(500, 2)
    for s in (0.5, 1.5, 2.0, 4.0):
        assert np.array_equal(ctx.context_region(r, s), O.context_region(r, s))
    m = modules.ContextRegion(ctx, 2.0)
    assert np.all(m.updateGradInput(r, None) == 0)                        # ContextRegion.lua:34-37


def test_bbox_norm_modes(ctx):
    d = np.random.default_rng(0).standard_normal((300, 84)).astype(np.float32)
    mean, std = [0.0, 0.01, -0.02, 0.03], [0.1, 0.1, 0.2, 0.2]
    m = modules.BBoxNorm(ctx, mean, std)
    assert np.array_equal(m.forward(d), d)                                # training mode: identity (BBoxNorm.lua:20)
    m.evaluate()
    assert np.array_equal(m.forward(d), O.bbox_norm(d, mean, std))
    with pytest.raises(RuntimeError):
        m.updateGradInput(d, d)                                           # BBoxNorm.lua:35
    with pytest.raises(ValueError):
        m.forward(np.zeros((2, 6), np.float32))


def test_bbox_decode(ctx):
    boxes = wl.random_boxes(1000, 600, 800, 3)
    d = (np.random.default_rng(1).standard_normal((1000, 84)) * 0.3).astype(np.float32)
    got, ref = U.convertFrom(ctx, boxes, d), O.convert_from(d, boxes)
    # only expf may differ by an ulp between libm and CUDA; everything else is the same op order
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-3)
    g = np.load(os.path.join(GOLD, "ops_golden.npz"))
    np.testing.assert_allclose(ctx.bbox_decode(g["deltas"], g["boxes"]), g["decoded"], rtol=2e-6, atol=1e-3)


@pytest.mark.parametrize("variant", [1, 2])
def test_roi_pool_reference_test_shape(ctx, variant):
    """modules/test.lua:60-65 shapes: (1,512,38,50), 40 rois randn*50 (negative / inverted => clipped & empty bins)"""
    rng = np.random.default_rng(5)
    fm = rng.standard_normal((1, 512, 38, 50)).astype(np.float32)
    rois = (rng.standard_normal((40, 5)) * 50).astype(np.float32)
    rois[:, 0] = 1
    out, am = ctx.roi_pool(fm, rois, 7, 7, 1 / 16, variant, with_argmax=True)
    ro, ra = O.roi_pool(fm, rois, 7, 7, 1 / 16, variant, with_argmax=True)
    assert np.array_equal(out, ro) and np.array_equal(am, ra)
    # chunk invariance (the property the reference test asserts with == 0)
    parts = np.concatenate([ctx.roi_pool(fm, rois[:25], 7, 7, 1 / 16, variant), ctx.roi_pool(fm, rois[25:], 7, 7, 1 / 16, variant)])
    assert np.array_equal(parts, out)


def test_roi_pool_realistic_and_batched(ctx):
    rng = np.random.default_rng(6)
    fm = rng.standard_normal((2, 64, 38, 50)).astype(np.float32)
    rois = _rois(300, 7)
    rois[::2, 0] = 2
    for (pw, ph, sc) in [(7, 7, 1 / 16), (6, 6, 1 / 16), (14, 14, 1 / 16)]:
        assert np.array_equal(ctx.roi_pool(fm, rois, pw, ph, sc, 2), O.roi_pool(fm, rois, pw, ph, sc, 2))
    m = modules.ROIPooling(ctx, 7, 7, 1 / 16)
    assert np.array_equal(m.forward((fm, rois)), O.roi_pool(fm, rois, 7, 7, 1 / 16, 2))
    g = np.load(os.path.join(GOLD, "ops_golden.npz"))
    assert np.array_equal(ctx.roi_pool(g["fmap"], g["rois_neg"], 7, 7, 1 / 16, 2), g["roi_v2"])
    assert np.array_equal(ctx.roi_pool(g["fmap"], g["rois_neg"], 7, 7, 1 / 16, 1), g["roi_v1"])


def test_roi_pool_rejects_bad_batch_index(ctx):
    import multipathnet_b200 as mpn
    fm = np.zeros((1, 8, 10, 10), np.float32)
    with pytest.raises(mpn.MpnError):
        ctx.roi_pool(fm, np.array([[3, 1, 1, 5, 5]], np.float32), 7, 7, 1.0)


def test_module_ops_on_device_buffers_equal_the_host_entry_points(ctx):
    """mpn_foveal_dev / mpn_context_region_dev / mpn_bbox_norm_dev (what the Lua modules call for CudaTensors: no host round
    trip) == the host-pointer entry points above, bit for bit (same kernels), which are pinned against the oracle"""
    import torch
    rng = np.random.default_rng(12)
    R = 333
    rois = np.concatenate([np.ones((R, 1), np.float32), wl.random_boxes(R, 600, 800, 12)], 1).astype(np.float32)
    r_d = torch.from_numpy(rois).cuda()
    out_f = torch.empty((4 * R, 5), dtype=torch.float32, device="cuda")
    ctx.foveal_dev(r_d, R, out_f)
    out_c = torch.empty((R, 5), dtype=torch.float32, device="cuda")
    ctx.context_region_dev(r_d, R, 1.5, out_c)
    d = rng.standard_normal((R, 84)).astype(np.float32)
    d_d = torch.from_numpy(d).cuda()
    mean, std = np.float32([0.1, -0.2, 0.05, 0.0]), np.float32([0.1, 0.1, 0.2, 0.2])
    ctx.bbox_norm_dev(d_d, R, 84, mean, std)
    ctx.synchronize()
    assert np.array_equal(out_f.cpu().numpy(), ctx.foveal(rois))
    assert np.array_equal(out_c.cpu().numpy(), ctx.context_region(rois, 1.5))
    assert np.array_equal(d_d.cpu().numpy(), ctx.bbox_norm(d, mean, std))
    assert np.array_equal(out_f.cpu().numpy(), O.foveal(rois))
