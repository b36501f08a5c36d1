"""CPU suite, part 4: the host-side glue of fbcoco.Tester_FRCNN (iterative localisation, rbox scores, bbox voting,
keep-top-k / transpose / COCO rows) against a literal restatement of the Lua control flow, with the CPU oracle as the
backend behind the same seam the C ABI sits behind in production (multipathnet_b200/tester.py:_AbiBackend)."""
import numpy as np
import pytest

from multipathnet_b200 import utils as U
from multipathnet_b200.modules import SelectBoxes
from multipathnet_b200.tester import Tester as FrcnnTester
from oracle import ref as O

C = 6            # classes incl. background


class FakeModel:
    C = C
    ctx = None


class IdentityTransformer:
    def forward(self, im):
        return im


def _fake_net(boxes):
    """A deterministic stand-in for detect(): class scores and per-class boxes as smooth functions of the proposals."""
    b = np.asarray(boxes, np.float64)
    ctr = np.stack([(b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2], 1)
    logits = np.stack([np.sin(ctr[:, 0] * (0.01 + 0.003 * c)) + np.cos(ctr[:, 1] * (0.02 - 0.002 * c)) for c in range(C)], 1)
    e = np.exp(logits - logits.max(1, keepdims=True))
    scores = (e / e.sum(1, keepdims=True)).astype(np.float32)
    per = [b + np.array([-c, -0.5 * c, 0.7 * c, c], np.float64) * 1.5 for c in range(C)]
    return scores, np.concatenate(per, 1).astype(np.float32)


class OracleBackend:
    def __init__(self):
        self.calls = []

    def detect(self, img, boxes, im_scale, recompute_features):
        self.calls.append((img is not None, bool(recompute_features), len(boxes)))
        return _fake_net(boxes)

    def detect_nms(self, img, boxes, im_scale, W0, H0, thresh, nms_thresh):
        scores, bboxes = _fake_net(boxes)
        bboxes = O.clamp_boxes(bboxes, W0, H0)
        keeps = []
        for j in range(1, C):
            idx = np.nonzero(scores[:, j] > thresh)[0]
            sb = np.concatenate([bboxes[idx, 4 * j:4 * j + 4], scores[idx, j:j + 1]], 1).astype(np.float32)
            keeps.append(idx[O.nms(sb, nms_thresh)] if len(idx) else idx)
        return scores, bboxes, keeps

    def nms_batched(self, sb, offsets, thr):
        return [O.nms(sb[offsets[s]:offsets[s + 1]], thr) for s in range(len(offsets) - 1)]

    def bbox_vote(self, nms_boxes, scored_boxes, thr):
        return O.bbox_vote(nms_boxes, scored_boxes, thr)


def _lua_testOne(im, boxes, num_iter, use_rbox, thresh, nms_thr, voting, vote_thr, vote_pow):
    """Tester_FRCNN.lua:54-139 restated step by step (torch -> numpy)."""
    H0, W0 = im.shape[1], im.shape[2]
    output, bbox_pred = _fake_net(boxes)
    tmp = bbox_pred.reshape(-1, 2)
    tmp[:, 0] = np.clip(tmp[:, 0], 1, W0); tmp[:, 1] = np.clip(tmp[:, 1], 1, H0)           # :75-78, in place
    all_output, all_bbox = [output], [bbox_pred]
    for _ in range(2, num_iter + 1):
        maxids = np.argmax(output, 1)
        new_boxes = np.stack([bbox_pred[np.arange(len(maxids)), maxids * 4 + i] for i in range(4)], 1)
        output, bbox_pred = _fake_net(new_boxes)
        all_output.append(output); all_bbox.append(bbox_pred)
    if use_rbox:
        all_output.pop(0); all_bbox.pop()
    output = np.concatenate(all_output, 0); bbox_pred = np.concatenate(all_bbox, 0)
    res = []
    for j in range(1, C):
        sc = output[:, j]
        idx = np.nonzero(sc > thresh)[0]
        sb = np.zeros((len(idx), 5), np.float32)
        if len(idx):
            sb[:, :4] = bbox_pred[idx, 4 * j:4 * j + 4]; sb[:, 4] = sc[idx]
        kept = sb[O.nms(sb, nms_thr)] if len(idx) else sb
        if voting and len(kept):
            resc = sb.copy(); resc[:, 4] = np.power(resc[:, 4], np.float32(vote_pow))
            kept = O.bbox_vote(kept, resc, vote_thr)
        res.append(kept)
    return res


def _boxes(n, seed, H=120, W=160):
    rng = np.random.default_rng(seed)
    x1 = rng.uniform(1, W - 30, n); y1 = rng.uniform(1, H - 30, n)
    return np.stack([x1, y1, x1 + rng.uniform(8, 60, n), y1 + rng.uniform(8, 50, n)], 1).astype(np.float32)


@pytest.mark.parametrize("num_iter,use_rbox,voting", [(1, False, False), (1, False, True), (2, False, False), (3, False, True),
                                                      (2, True, False), (3, True, True)])
def test_testOne_matches_the_lua_control_flow(oracle_built, num_iter, use_rbox, voting):
    im = np.zeros((3, 120, 160), np.float32)
    boxes = _boxes(90, 5 + num_iter)
    be = OracleBackend()
    t = FrcnnTester(FakeModel(), IdentityTransformer(), scale=[120], max_size=1000, nms_thresh=0.3, bbox_vote_thresh=0.5,
               score_thresh=0.12, bbox_voting=voting, num_iterative_loc=num_iter, use_rbox_scores=use_rbox,
               bbox_voting_score_pow=2.0, backend=be)
    got = t.testOne(im, boxes)
    want = _lua_testOne(im, boxes, num_iter, use_rbox, 0.12, 0.3, voting, 0.5, 2.0)
    assert len(got) == len(want) == C - 1
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    if num_iter > 1:      # trunk recomputed once, later passes re-use the cached features (ImageDetect.lua:107-111)
        assert [c[1] for c in be.calls] == [True] + [False] * (num_iter - 1)
        assert [c[0] for c in be.calls] == [True] + [False] * (num_iter - 1)


def test_rbox_scores_need_two_passes():
    with pytest.raises(ValueError):
        FrcnnTester(FakeModel(), IdentityTransformer(), use_rbox_scores=True, backend=OracleBackend())


def test_select_boxes_first_maximum_and_denormalise():
    cls = np.array([[0.1, 0.7, 0.7], [0.9, 0.05, 0.05]], np.float32)
    ys = np.arange(24, dtype=np.float32).reshape(2, 12)
    out = SelectBoxes().forward([cls, ys])
    assert np.array_equal(out, np.array([[4, 5, 6, 7], [12, 13, 14, 15]], np.float32))       # ties -> first maximum
    out2 = SelectBoxes(mean=[1, 2, 3, 4], std=[2, 2, 2, 2]).forward([cls, ys])
    assert np.array_equal(out2, out * 2 + np.array([1, 2, 3, 4], np.float32))
    with pytest.raises(ValueError):
        SelectBoxes().forward([cls, ys[:, :8]])


def test_keep_top_k_transpose_and_coco_rows():
    rng = np.random.default_rng(3)
    imgs = [[np.concatenate([rng.uniform(1, 50, (n, 4)), rng.random((n, 1))], 1).astype(np.float32) for n in (3, 0, 5)] for _ in range(2)]
    kept = FrcnnTester.keepTopKPerImage(imgs, 4)
    for per_img, src in zip(kept, imgs):
        allsc = np.sort(np.concatenate([b[:, 4] for b in src if b.size]))[::-1]
        thr = allsc[min(len(allsc), 4) - 1]
        assert sum(len(b) for b in per_img) == int((np.concatenate([b[:, 4] for b in src if b.size]) >= thr).sum())     # `>=` tie rule
    tr = U.transposeBoxes(kept, 3)
    assert len(tr) == 3 and len(tr[0]) == 2 and tr[2][1] is kept[1][2]
    rows = U.coco_results(tr, image_ids=[42, 77], category_ids=[1, 5, 9])
    assert rows.shape[1] == 7 and rows.shape[0] == sum(len(b) for per in kept for b in per)
    first = tr[0][0][0]
    assert np.allclose(rows[0], [42, first[0] - 1, first[1] - 1, first[2] - first[0], first[3] - first[1], first[4], 1])
    assert set(np.unique(rows[:, 6])) <= {1.0, 5.0, 9.0}
