"""GPU suite: several model replicas on ONE GPU (mpn_ctx_create_stream / mpn_ctx_wait_ctx, multipathnet_b200.ModelReplicas) —
the reference's one-replica-per-donkey-thread arrangement (test_runner.lua:55-66) with K threads per GPU. Work of different
replicas runs concurrently on their own streams; every result must equal the single-model result bit for bit, and the join
before the end-of-run gather must order replica 0's stream after the others."""
import numpy as np
import pytest

import multipathnet_b200 as mpn
from multipathnet_b200 import dist as mdist, models, workloads as wl

pytestmark = pytest.mark.gpu

H, W, R = 150, 203, 180


def _inputs(spec, i):
    return wl.transform(wl.raw_image(H, W, 40 + i), spec.transformer), wl.random_boxes(R, H, W, 40 + i)


def test_ctx_with_its_own_stream(ctx):
    c2 = mpn.Context(0, own_stream=True)
    c3 = mpn.Context(0, own_stream=True, priority=-5)        # clamped to the device's range
    assert c2.stream_handle != 0 and c3.stream_handle not in (0, c2.stream_handle)
    assert ctx.stream_handle == 0                               # the session ctx sits on the legacy default stream
    c2.wait_ctx(c3); c2.wait_ctx(c2); ctx.wait_ctx(c2)
    boxes = wl.random_boxes(50, H, W, 1)
    sb = np.concatenate([boxes, np.linspace(1, 0, 50, dtype=np.float32)[:, None]], 1).astype(np.float32)
    assert np.array_equal(c2.nms(sb, 0.3), ctx.nms(sb, 0.3))       # an op on a ctx with its own stream
    c2.synchronize(); c3.synchronize()
    c2.close(); c3.close()


@pytest.mark.parametrize("n_rep", [2, 3])
def test_replicas_equal_single_model_and_join_orders_the_gather(ctx, n_rep):
    import torch
    spec = models.vgg16_fast_rcnn(21, seed=7, width_div=4, fc_dim=256)
    single = mpn.Model(ctx, spec, max_rois=512, max_h=256, max_w=320)
    n_img = 7
    inputs = [_inputs(spec, i) for i in range(n_img)]
    want = [single.detect_nms(im, bx, 1.0, W, H, -1.5, 0.3) for im, bx in inputs]
    want_rec = []
    for scores, bboxes, keeps in want:
        tables = [np.concatenate([bboxes[k, 4 * j:4 * j + 4], scores[k, j:j + 1]], 1).astype(np.float32) for j, k in enumerate(keeps, start=1)]
        want_rec.append(mdist.pack_record(mdist.tables_to_dets(tables)))
    single.close()

    reps = mpn.ModelReplicas(0, spec, n_rep, max_rois=512, max_h=256, max_w=320)
    assert len(reps) == n_rep and len({c.stream_handle for c in reps.ctxs}) == n_rep
    per = (n_img + n_rep - 1) // n_rep
    rec_d = torch.zeros((n_rep, per, mpn.MPN_REC_FLOATS), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()            # the zero fill ran on torch's stream; the replicas' streams are non-blocking
    for k, m in enumerate(reps.models):
        m.set_detection_sink(rec_d[k], per, 100)
    # pipelined public API, images dealt round-robin, two in flight per replica
    got = [None] * n_img
    pending = [[] for _ in range(n_rep)]
    for i, (im, bx) in enumerate(inputs):
        k = i % n_rep
        if len(pending[k]) == 2:
            j, t = pending[k].pop(0)
            got[j] = reps.models[k].detect_nms_wait(t)
        pending[k].append((i, reps.models[k].detect_nms_submit(im, bx, 1.0, W, H, -1.5, 0.3)))
    # the join + ONE gather on replica 0's context, BEFORE any host-side wait: stream order alone must make every record visible
    reps.join()
    g = mdist.gather_records_dev(reps.ctxs[0], rec_d.view(-1, mpn.MPN_REC_FLOATS), n_rep * per)[0].reshape(n_rep, per, -1)
    for k in range(n_rep):
        for j, t in pending[k]:
            got[j] = reps.models[k].detect_nms_wait(t)
    for i in range(n_img):
        (s0, b0, k0), (s1, b1, k1) = want[i], got[i]
        assert np.array_equal(s0, s1) and np.array_equal(b0, b1), f"image {i}"
        assert all(np.array_equal(a, b) for a, b in zip(k0, k1)), f"image {i}"
        assert np.array_equal(g[i % n_rep, i // n_rep], want_rec[i]), f"record of image {i}"
    reps.close()
