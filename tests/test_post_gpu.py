"""GPU parity of the post-NMS device tail: pack_detections_kernel == utils.keep_top_k
(utils.lua:75-96) + the fixed-size record, select_boxes_kernel == nn.SelectBoxes (modules/SelectBoxes.lua:26-56), the
model's detection sink, and the library-issued all-gather in a world of one. Bit-exact (index / copy work)."""
import numpy as np
import pytest

import multipathnet_b200 as mpn
from multipathnet_b200 import dist as mdist, models, utils as U, workloads as wl

pytestmark = pytest.mark.gpu


def synth(R, C, seed, quant=None, per_class=None):
    """scores / boxes + per-class keep lists in non-increasing score order (what nms.c emits)"""
    rng = np.random.default_rng(seed)
    scores = rng.random((R, C)).astype(np.float32)
    if quant:
        scores = (np.floor(scores * quant) / quant).astype(np.float32)
    bboxes = (rng.random((R, 4 * C)) * 500).astype(np.float32)
    keep = np.zeros((C - 1, R), np.int32)
    counts = np.zeros(C - 1, np.int32)
    for j in range(1, C):
        n = per_class if per_class is not None else int(rng.integers(0, R // 2))
        rows = rng.choice(R, size=n, replace=False)
        rows = rows[np.argsort(-scores[rows, j], kind="stable")]
        keep[j - 1, :n] = rows; counts[j - 1] = n
    return scores, bboxes, keep, counts


def host_record(scores, bboxes, keep, counts, top_k=100):
    C = scores.shape[1]
    tables = []
    for j in range(1, C):
        k = keep[j - 1, :counts[j - 1]]
        tables.append(np.concatenate([bboxes[k, 4 * j:4 * j + 4], scores[k, j:j + 1]], 1).astype(np.float32))
    kept, _ = U.keep_top_k([t.copy() for t in tables], top_k)
    return kept, mdist.tables_to_dets(kept)


@pytest.mark.parametrize("R,C,quant,per_class", [(1000, 21, None, None), (1000, 81, None, None), (300, 21, 16, None), (64, 3, None, 5),
                                                  (2000, 81, 64, None), (50, 21, None, 0), (400, 21, None, 3)])
def test_pack_detections_equals_keep_top_k(ctx, R, C, quant, per_class):
    scores, bboxes, keep, counts = synth(R, C, R + C, quant, per_class)
    kept, dets = host_record(scores, bboxes, keep, counts)
    rec = ctx.pack_detections(scores, bboxes, keep, counts, 100)
    assert int(rec[0]) == dets.shape[0]
    if dets.shape[0] <= mpn.MPN_MAX_DET:
        assert np.array_equal(mdist.unpack_record(rec), dets)                     # class-major, emission order, bit for bit
        assert np.all(rec[1 + 6 * dets.shape[0]:] == 0)
        back = mdist.record_to_tables(rec, C)
        assert all(np.array_equal(a.reshape(-1, 5), b) for a, b in zip(kept, back))
    else:                                                                          # ties overflow the record: count says so, loudly on the host
        with pytest.raises(OverflowError):
            mdist.unpack_record(rec)
        assert np.array_equal(rec[1:1 + 6 * mpn.MPN_MAX_DET].reshape(-1, 6), dets[:mpn.MPN_MAX_DET])


def test_pack_detections_all_tied_overflow_and_other_top_k(ctx):
    scores, bboxes, keep, counts = synth(500, 21, 3, quant=1, per_class=20)          # every score 0.0: 400 tied rows
    rec = ctx.pack_detections(scores, bboxes, keep, counts, 100)
    assert int(rec[0]) == 400
    scores, bboxes, keep, counts = synth(500, 21, 4)
    for k in (1, 7, 128):
        _, dets = host_record(scores, bboxes, keep, counts, k)
        assert np.array_equal(mdist.unpack_record(ctx.pack_detections(scores, bboxes, keep, counts, k)), dets)
    with pytest.raises(RuntimeError, match="top_k"):
        ctx.pack_detections(scores, bboxes, keep, counts, 129)


@pytest.mark.parametrize("R,C", [(1, 2), (128, 21), (1000, 81), (2500, 21)])
def test_select_boxes(ctx, R, C):
    rng = np.random.default_rng(R)
    classes = rng.random((R, C)).astype(np.float32)
    classes[::7] = np.round(classes[::7], 1)                                          # tied maxima: the first one wins
    ys = rng.standard_normal((R, 4 * C)).astype(np.float32)
    a = classes.argmax(1)
    want = np.stack([ys[np.arange(R), 4 * a + i] for i in range(4)], 1)
    assert np.array_equal(ctx.select_boxes(classes, ys), want)
    mean, std = np.float32([0.1, -0.2, 0.3, 0.05]), np.float32([0.1, 0.1, 0.2, 0.2])
    assert np.array_equal(ctx.select_boxes(classes, ys, mean, std), want * std + mean)    # output:cmul(sigma):add(mu)
    assert np.array_equal(ctx.select_boxes(classes, ys), mpn.modules.SelectBoxes().forward([classes, ys]))


def test_model_detection_sink_and_world_of_one_gather(ctx):
    """every detect+NMS pass appends the image's record; the records equal keep_top_k of the returned outputs; the library's
    all-gather in a world of one hands them back unchanged; a full sink fails loudly"""
    import torch
    spec = models.vgg16_fast_rcnn(21, seed=7, width_div=4, fc_dim=256)
    m = mpn.Model(ctx, spec, max_rois=512, max_h=256, max_w=320)
    H, W, R, n = 150, 203, 250, 3
    rec_d = torch.zeros((n, mpn.MPN_REC_FLOATS), dtype=torch.float32, device="cuda")
    m.set_detection_sink(rec_d, n, 100)
    want = []
    for i in range(n):
        img = wl.transform(wl.raw_image(H, W, 20 + i), spec.transformer)
        boxes = wl.random_boxes(R, H, W, 20 + i)
        scores, bboxes, keeps = m.detect_nms(img, boxes, 1.0, W, H, -1.5, 0.3)
        tables = [np.concatenate([bboxes[k, 4 * j:4 * j + 4], scores[k, j:j + 1]], 1).astype(np.float32) for j, k in enumerate(keeps, start=1)]
        want.append(mdist.pack_record(mdist.tables_to_dets(tables)))
    assert m.detection_sink_count() == n
    got = rec_d.cpu().numpy()
    assert np.array_equal(got, np.stack(want))
    assert ctx.dist_world() == (0, 1)
    g = mdist.gather_records_dev(ctx, rec_d, n)
    assert g.shape == (1, n, mpn.MPN_REC_FLOATS) and np.array_equal(g[0], got)
    with pytest.raises(RuntimeError, match="sink is full"):
        m.detect_nms(img, boxes, 1.0, W, H, -1.5, 0.3)
    m.set_detection_sink(None, 0)
    m.detect_nms(img, boxes, 1.0, W, H, -1.5, 0.3)
    m.close()


# ---- Tester_FRCNN:testOne with its test-time options on the device (mpn_model_test_one) -------------------
class _OracleNmsBackend:
    """the seam of multipathnet_b200/tester.py with the GPU network behind detect() and the CPU ORACLE (nms.c restatement,
    pinned by the literal build) behind NMS / voting: what the device tail must reproduce bit for bit"""

    def __init__(self, model):
        self.model = model

    def detect(self, img, boxes, im_scale, recompute_features):
        return self.model.detect(img, boxes, im_scale, recompute_features)

    def detect_nms(self, img, boxes, im_scale, W0, H0, thresh, nms_thresh):
        from oracle import ref as O
        scores, bboxes = self.model.detect(img, boxes, im_scale, True)
        bboxes = O.clamp_boxes(bboxes, W0, H0)
        keeps = []
        for j in range(1, scores.shape[1]):
            idx = np.nonzero(scores[:, j] > thresh)[0]
            sb = np.concatenate([bboxes[idx, 4 * j:4 * j + 4], scores[idx, j:j + 1]], 1).astype(np.float32)
            keeps.append(idx[O.nms(sb, nms_thresh)] if len(idx) else idx)
        return scores, bboxes, keeps

    def nms_batched(self, sb, offsets, thr):
        from oracle import ref as O
        return [O.nms(sb[offsets[s]:offsets[s + 1]], thr) for s in range(len(offsets) - 1)]

    def bbox_vote(self, nms_boxes, scored_boxes, thr):
        from oracle import ref as O
        return O.bbox_vote(nms_boxes, scored_boxes, thr)


@pytest.mark.parametrize("num_iter,rbox,voting", [(2, False, True), (1, False, True), (3, True, False), (2, True, True), (2, False, False)])
def test_tester_test_one_on_the_device(ctx, num_iter, rbox, voting):
    """Tester(num_iterative_loc, use_rbox_scores, bbox_voting) at R = 1000: ONE mpn_model_test_one call == the host-side glue
    of the reference's control flow around detect() with the oracle's nms.c / bbox_vote — same boxes, same order, same bits"""
    spec = models.vgg16_fast_rcnn(21, seed=7, width_div=4, fc_dim=256)
    m = mpn.Model(ctx, spec, max_rois=4096, max_h=256, max_w=320)
    raw = wl.raw_image(150, 203, 31)
    boxes = wl.random_boxes(1000, 150, 203, 31)
    tf = mpn.modules.ImageTransformer(spec.transformer)
    kw = dict(scale=[150], max_size=400, num_iterative_loc=num_iter, use_rbox_scores=rbox, bbox_voting=voting, score_thresh=0.02)
    dev = mpn.Tester(m, tf, **kw)
    host = mpn.Tester(m, tf, backend=_OracleNmsBackend(m), device_tail=False, **kw)
    assert dev.device_tail and not host.device_tail
    n0 = ctx.launch_count
    got = dev.testOne(raw, boxes)
    assert ctx.launch_count > n0
    want = host.testOne(raw, boxes)
    assert len(got) == len(want) == spec.num_classes - 1
    for j, (a, b) in enumerate(zip(got, want), start=1):
        assert a.shape == b.shape and np.array_equal(a, b), f"class {j}"
    assert np.array_equal(dev.raw[0], host.raw[0]) and np.array_equal(dev.raw[1], host.raw[1])
    assert sum(len(a) for a in got) > 50
    m.close()
