"""CPU suite, part 6: the general nn-graph importer (multipathnet_b200.t7.model_from_t7) on graphs assembled the way
models/multipathnet.lua:30-121 and models/resnet.lua:28-50 assemble them (tiny widths).  PARITY UNPINNED against real
.t7 files (none in the image).  The check is two independent evaluations of the same graph: tests/_nn_interp.py walks the
nn modules (batch norm unfused, MulConstant / Narrow / ModeSwitch as modules) and the CPU oracle runs the imported
ModelSpec (batch norm folded, factors folded into conv_mix, heads as column ranges)."""
import io

import numpy as np
import pytest

from multipathnet_b200 import t7
from multipathnet_b200._lib import Model
from multipathnet_b200.t7 import T7Object
from oracle import graphs as G
from _nn_interp import evaluate, _t

O = lambda name, **f: T7Object(name, f)


def _roundtrip(o):
    buf = io.BytesIO()
    t7.save(buf, o)
    buf.seek(0)
    return t7.load(buf)


def _seq(*mods):
    return O("nn.Sequential", modules=list(mods))


def _conv(rng, cin, cout, k=3, s=1, p=None, bias=True, gain=1.0):
    p = (k // 2) if p is None else p
    f = {"nInputPlane": cin, "nOutputPlane": cout, "kW": k, "kH": k, "dW": s, "dH": s, "padW": p, "padH": p, "groups": 1,
         "weight": (rng.standard_normal((cout, cin, k, k)) * gain * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)}
    if bias:
        f["bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)
    return T7Object("cudnn.SpatialConvolution", f)


def _linear(rng, cout, cin, std=None, zero_bias=False):
    return O("nn.Linear", weight=(rng.standard_normal((cout, cin)) * (std or np.sqrt(2.0 / cin))).astype(np.float32),
             bias=np.zeros(cout, np.float32) if zero_bias else (rng.standard_normal(cout) * 0.05).astype(np.float32))


def _relu():
    return O("cudnn.ReLU", inplace=True)


def _pool(k=2, s=2, p=0, ceil=True):
    return O("cudnn.SpatialMaxPooling", kW=k, kH=k, dW=s, dH=s, padW=p, padH=p, ceil_mode=ceil)


def _ident():
    return O("nn.Identity")


def _par(*mods):
    return O("nn.ParallelTable", modules=list(mods))


def _cat(*mods):
    return O("nn.ConcatTable", modules=list(mods))


# ------------------------------------------------------------------------------------------------- MultiPathNet
def _conv345(rng, chans, scales, normalized, use3, use4, P=3):
    """model_utils.lua:209-251"""
    def pool1(idx, nfeat, sc, factor):
        s = _seq(_par(O("nn.SelectTable", index=idx), _ident()), O("inn.ROIPooling", W=P, H=P, spatial_scale=sc))
        if normalized:
            s.fields["modules"] += [O("nn.View", size=[-1, nfeat * P * P]), O("nn.Normalize", p=2, eps=1e-10), O("nn.Contiguous"),
                                    O("nn.View", size=[-1, nfeat, P, P])]
        else:
            s.fields["modules"].append(O("nn.MulConstant", constant_scalar=factor))
        return s
    levels, tot = [pool1(1, chans[0], scales[0], 1.0)], chans[0]
    if use4:
        levels.append(pool1(2, chans[1], scales[1], 1.0 / 30)); tot += chans[1]
    if use3:
        levels.append(pool1(3, chans[2], scales[2], 1.0 / 200)); tot += chans[2]
    join = _seq(_cat(*levels), O("nn.JoinTable", dimension=2))
    if normalized:
        join.fields["modules"].append(O("nn.MulConstant", constant_scalar=1000))
    join.fields["modules"] += [_conv(rng, tot, chans[0], k=1, gain=0.7 if normalized else 30.0), O("nn.View", size=[-1], numInputDims=3)]
    return join


def _tiny_multipathnet(rng, normalized=True, het=True, integral_k=0, C=4, fc=24, P=3):
    c3, c4, c5 = 8, 16, 16
    feats = [_conv(rng, 3, 8, gain=1 / 8.0), _relu(), _pool(), _conv(rng, 8, c3), _relu()]           # ... conv3 tap at 1/2
    conv4 = _seq(_pool(), _conv(rng, c3, c4), _relu())                                               # 1/4
    conv5 = _seq(_pool(), _conv(rng, c4, c5), _relu())                                               # 1/8
    skip = _seq(*feats, _cat(conv4, _ident()), _par(_cat(conv5, _ident()), _ident()), O("nn.FlattenTable"))
    classifier = lambda: _seq(_linear(rng, fc, c5 * P * P), O("nn.ReLU"), O("nn.Dropout", p=0.5, v2=True, inplace=True),
                              _linear(rng, fc, fc), O("nn.ReLU"), O("nn.Dropout", p=0.5, v2=True, inplace=True))
    N = 4
    model = _seq(_par(O("nn.NoBackprop", modules=[O("nn.DataParallelTable", modules=[skip])]), _ident()),
                 _par(_ident(), _seq(O("nn.Foveal"), O("nn.View", size=[-1, N, 5]), O("nn.Transpose", permutations=[[1, 2]]))))
    regions = O("nn.ModelParallelTable", dimension=2, modules=[], gpuAssignments=[])
    chans, scales = (c5, c4, c3), (1 / 8.0, 1 / 4.0, 1 / 2.0)
    for i in range(1, N + 1):
        regions.fields["modules"].append(_seq(_par(_ident(), O("nn.Select", dimension=1, index=i)),
                                              _conv345(rng, chans, scales, normalized, i == 1, i <= 3, P), classifier()))
    if het:
        regions.fields["modules"].append(_seq(_par(_ident(), O("nn.Select", dimension=1, index=2)),
                                              _conv345(rng, chans, scales, normalized, True, True, P), classifier()))
    model.fields["modules"].append(regions)
    cls = [_linear(rng, C, N * fc, 0.05) for _ in range(max(integral_k, 1))]
    bbox = _linear(rng, 4 * C, fc if het else N * fc, 0.02)
    cls_m = _cat(*cls) if integral_k else cls[0]
    if het:
        model.fields["modules"].append(_cat(O("nn.Narrow", dimension=2, index=1, length=N * fc),
                                            O("nn.Narrow", dimension=2, index=N * fc + 1, length=fc)))
        model.fields["modules"].append(_par(cls_m, bbox))
    else:
        model.fields["modules"].append(_cat(cls_m, bbox))
    if integral_k:                                                                                   # model_utils.lua:275-317
        sm = _par(*[_seq(O("nn.SoftMax"), O("nn.View", size=[1, -1, C])) for _ in range(integral_k)])
        model.fields["modules"].append(O("nn.ModeSwitch", train=False, modules=[
            _par(O("nn.SelectTable", index=1), _ident()),
            _seq(_par(_seq(sm, O("nn.JoinTable", dimension=1), O("nn.Mean", dimension=1)), _ident()))]))
        model.fields["noSoftMax"] = True
    model.fields["modules"].append(_par(_ident(), O("nn.BBoxNorm", mean=np.zeros((1, 4), np.float32),
                                                      std=np.array([[0.1, 0.1, 0.2, 0.2]], np.float32))))
    return model


def _inputs(rng, H=40, W=56, R=6):
    img = (rng.standard_normal((3, H, W)) * 40).astype(np.float32)
    x1 = rng.uniform(1, W - 12, R); y1 = rng.uniform(1, H - 12, R)
    rois = np.stack([np.ones(R), x1, y1, x1 + rng.uniform(4, 11, R), y1 + rng.uniform(4, 11, R)], 1).astype(np.float32)
    return img, rois


def _both(model, img, rois, **kw):
    spec = t7.model_from_t7(_roundtrip(model), **kw)
    ref_cls, ref_bbox = evaluate(model, [_t(img)[None], _t(rois)])
    ts = G.trunk_forward(spec, img)
    cls, bbox = G.heads_forward(spec, ts, rois)
    return spec, (ref_cls.numpy(), ref_bbox.numpy()), (cls, bbox)


def _close(a, b, rel=2e-5):
    return np.abs(a - b).max() <= rel * max(np.abs(b).max(), 1e-6)


@pytest.mark.parametrize("normalized,het,integral_k", [(True, True, 0), (False, True, 0), (True, False, 0), (True, True, 3)])
def test_multipathnet_graph_import_matches_module_evaluation(oracle_built, normalized, het, integral_k):
    rng = np.random.default_rng(11 + 2 * normalized + het)
    model = _tiny_multipathnet(rng, normalized, het, integral_k)
    img, rois = _inputs(rng)
    spec, (rc, rb), (c, b) = _both(model, img, rois)
    assert len(spec.towers) == (5 if het else 4) and [t.region for t in spec.towers] == [0, 1, 2, 3] + ([1] if het else [])
    assert [len(t.levels) for t in spec.towers] == [3, 2, 2, 1] + ([3] if het else [])
    assert all(t.normalize == int(normalized) for t in spec.towers) and spec.transformer == "ross"
    assert [s for s, _ in spec.towers[0].levels] == [spec.taps["out1"], spec.taps["out2"], spec.taps["out3"]]     # conv5 | conv4 | conv3
    assert [sc for _, sc in spec.towers[0].levels] == [1 / 8.0, 1 / 4.0, 1 / 2.0]
    assert len(spec.cls_heads) == max(integral_k, 1) and spec.no_softmax == int(integral_k > 0) and spec.has_bbox_norm == 1
    if het:
        assert (spec.cls_heads[0].col_begin, spec.cls_heads[0].col_len) == (0, 96) and (spec.bbox_head.col_begin, spec.bbox_head.col_len) == (96, 24)
    assert rc.shape == c.shape == (6, 4) and rb.shape == b.shape == (6, 16)
    assert _close(c, rc) and _close(b, rb)
    if integral_k:
        assert np.allclose(c.sum(1), 1.0, atol=1e-5)                       # probabilities: mean of K softmaxes
    Model.build_desc(spec)                                                 # the C-ABI description accepts it (host-only)


# ------------------------------------------------------------------------------------------------------ ResNet
def _bn(rng, c, fixed):
    g, beta = rng.uniform(0.5, 1.5, c), rng.standard_normal(c) * 0.1
    mean, var = rng.standard_normal(c) * 0.2, rng.uniform(0.5, 2.0, c)
    if fixed:                                                              # inn.utils.BNtoFixed -> inn.ConstAffine
        a = g / np.sqrt(var + 1e-5)
        return O("inn.ConstAffine", a=a.astype(np.float32), b=(beta - mean * a).astype(np.float32), inplace=True)
    return O("nn.SpatialBatchNormalization", weight=g.astype(np.float32), bias=beta.astype(np.float32), eps=1e-5, momentum=0.1,
             affine=True, train=False, running_mean=mean.astype(np.float32), running_var=var.astype(np.float32))


def _bottleneck(rng, cin, n, stride, fixed):
    """fb.resnet.torch models/resnet.lua bottleneck, shortcut type B"""
    s = _seq(_conv(rng, cin, n, 1, bias=False), _bn(rng, n, fixed), _relu(),
             _conv(rng, n, n, 3, stride, bias=False), _bn(rng, n, fixed), _relu(),
             _conv(rng, n, 4 * n, 1, bias=False, gain=0.5), _bn(rng, 4 * n, fixed))
    short = _seq(_conv(rng, cin, 4 * n, 1, stride, 0, bias=False, gain=0.5), _bn(rng, 4 * n, fixed)) if (cin != 4 * n or stride != 1) else _ident()
    return _seq(_cat(s, short), O("nn.CAddTable", inplace=True), _relu())


def _tiny_resnet(rng, C=4, integral_k=0):
    net = [_conv(rng, 3, 8, 7, 2, 3, bias=False, gain=0.1), _bn(rng, 8, True), _relu(), _pool(3, 2, 1, ceil=False),
           _seq(_bottleneck(rng, 8, 4, 1, True), _bottleneck(rng, 16, 4, 1, True)),                 # layer1
           _seq(_bottleneck(rng, 16, 8, 2, False), _bottleneck(rng, 32, 8, 1, False)),              # layer2 (1/8)
           _seq(_bottleneck(rng, 32, 8, 1, False))]                                                 # layer3
    features = _seq(O("nn.NoBackprop", modules=[_seq(*net[:5])]), *net[5:])                          # utils.disableFeatureBackprop(features, 5)
    classifier = _seq(_seq(_bottleneck(rng, 32, 16, 2, False), _bottleneck(rng, 64, 16, 1, True)),   # layer4
                      O("nn.SpatialAveragePooling", kW=2, kH=2, dW=1, dH=1, padW=0, padH=0), O("nn.View", size=[64], numInputDims=3))
    cls = [_linear(rng, C, 64, 0.05) for _ in range(max(integral_k, 1))]
    model = _seq(_par(O("nn.DataParallelTable", modules=[features]), _ident()), O("inn.ROIPooling", W=4, H=4, spatial_scale=1 / 8.0),
                 O("nn.DataParallelTable", modules=[classifier]), _cat(_cat(*cls) if integral_k else cls[0], _linear(rng, 4 * C, 64, 0.02)))
    if integral_k:
        sm = _par(*[_seq(O("nn.SoftMax"), O("nn.View", size=[1, -1, C])) for _ in range(integral_k)])
        model.fields["modules"].append(O("nn.ModeSwitch", train=False, modules=[
            _par(O("nn.SelectTable", index=1), _ident()),
            _seq(_par(_seq(sm, O("nn.JoinTable", dimension=1), O("nn.Mean", dimension=1)), _ident()))]))
        model.fields["noSoftMax"] = True
    return model


@pytest.mark.parametrize("integral_k", [0, 2])
def test_resnet_graph_import_folds_batchnorm_and_residuals(oracle_built, integral_k):
    rng = np.random.default_rng(21 + integral_k)
    model = _tiny_resnet(rng, integral_k=integral_k)
    img, rois = _inputs(rng, 48, 64, 5)
    spec, (rc, rb), (c, b) = _both(model, img, rois)
    assert spec.transformer == "imagenet" and spec.has_bbox_norm == 0 and len(spec.cls_heads) == max(integral_k, 1)
    convs = [L for L in spec.trunk_layers if L.kind == 1]
    assert len(convs) == 1 + 4 + 3 + 4 + 3 + 3 and sum(L.residual_slot >= 0 for L in convs) == 5
    assert all(L.relu == 1 for L in convs if L.residual_slot >= 0)                     # ReLU after the add rides on the last conv
    for L in spec.trunk_layers:                                                        # a residual is produced before it is read
        if L.residual_slot > 0:
            order = [x.out_slot for x in spec.trunk_layers]
            assert order.index(L.residual_slot) < order.index(L.out_slot)
    tw = spec.towers[0]
    assert (tw.pooled_w, tw.pooled_h, tw.levels[0][1]) == (4, 4, 1 / 8.0) and tw.layers[-1].kind == 3  # global average pool last
    assert _close(c, rc, 5e-5) and _close(b, rb, 5e-5)
    Model.build_desc(spec)


def test_model_from_t7_also_reads_the_flat_fast_rcnn_graph(oracle_built):
    from test_t7_cpu import _tiny_fast_rcnn
    rng = np.random.default_rng(5)
    model, _ = _tiny_fast_rcnn(rng)
    a, b = t7.fast_rcnn_from_t7(model), t7.model_from_t7(model)
    img, rois = _inputs(rng, 24, 32, 4)
    (ca, ba), (cb, bb) = G.heads_forward(a, G.trunk_forward(a, img), rois), G.heads_forward(b, G.trunk_forward(b, img), rois)
    assert np.array_equal(ca, cb) and np.array_equal(ba, bb) and b.bbox_mean == a.bbox_mean and b.has_bbox_norm == 1


def test_graph_importer_refuses_what_it_cannot_represent():
    rng = np.random.default_rng(2)
    m = _tiny_resnet(rng)
    feats = t7.flatten_sequential(m.fields["modules"][0].fields["modules"][0])
    feats[1].typename = "nn.SpatialBatchNormalization"; feats[1].fields.pop("a")      # batch norm without statistics
    with pytest.raises((ValueError, AttributeError, KeyError)):
        t7.model_from_t7(m)
    m2 = _tiny_multipathnet(rng)
    m2.fields["modules"][2].fields["dimension"] = 1
    with pytest.raises(NotImplementedError):
        t7.model_from_t7(m2)
    m3 = _seq(_par(_seq(_pool(), _relu()), _ident()), O("inn.ROIPooling", W=2, H=2, spatial_scale=0.5))
    with pytest.raises(NotImplementedError):                                          # ReLU with no convolution to ride on
        t7.model_from_t7(m3)


def _same_spec(a, b):
    key = lambda L: (L.kind, L.cin, L.cout, L.kh, L.kw, L.stride, L.pad, L.relu, L.ceil_mode, L.residual_slot >= 0)
    assert [key(L) for L in a.trunk_layers] == [key(L) for L in b.trunk_layers]
    assert len(a.towers) == len(b.towers)
    for ta, tb in zip(a.towers, b.towers):
        assert (ta.region, ta.pooled_w, ta.pooled_h, ta.normalize, [sc for _, sc in ta.levels]) == (tb.region, tb.pooled_w, tb.pooled_h, tb.normalize, [sc for _, sc in tb.levels])
        assert [key(L) for L in ta.layers] == [key(L) for L in tb.layers]
        for La, Lb in zip(ta.layers, tb.layers):
            if La.weight >= 0:
                assert np.array_equal(a.weights[La.weight].ravel(), b.weights[Lb.weight].ravel()) and np.array_equal(a.weights[La.bias], b.weights[Lb.bias])
    for La, Lb in zip(a.trunk_layers, b.trunk_layers):
        if La.weight >= 0:
            assert np.array_equal(a.weights[La.weight].ravel(), b.weights[Lb.weight].ravel())
    for ha, hb in zip(list(a.cls_heads) + [a.bbox_head], list(b.cls_heads) + [b.bbox_head]):
        assert (ha.col_begin, ha.col_len, ha.cout) == (hb.col_begin, hb.col_len, hb.cout) and np.array_equal(a.weights[ha.weight], b.weights[hb.weight])
    assert (a.num_classes, a.no_softmax, a.has_bbox_norm) == (b.num_classes, b.no_softmax, b.has_bbox_norm)
    assert np.allclose(a.bbox_std, b.bbox_std) and np.allclose(a.bbox_mean, b.bbox_mean)


@pytest.mark.parametrize("which", ["vgg16_fast_rcnn", "vgg16_multipathnet", "vgg16_multipathnet_integral"])
def test_full_depth_specs_survive_export_save_load_import(oracle_built, which):
    """the bench's own graphs (full VGG-16 depth, narrow widths): ModelSpec -> nn graph -> .t7 bytes -> nn graph -> ModelSpec,
    same layers, same taps in {conv5, conv4, conv3} order, same weights; and the two specs give the same forward."""
    from multipathnet_b200 import models
    spec = {"vgg16_fast_rcnn": lambda: models.vgg16_fast_rcnn(21, seed=1, width_div=8, fc_dim=64),
            "vgg16_multipathnet": lambda: models.vgg16_multipathnet(11, seed=2, width_div=8, fc_dim=32),
            "vgg16_multipathnet_integral": lambda: models.vgg16_multipathnet(11, seed=3, width_div=8, fc_dim=32, integral_k=3)}[which]()
    back = t7.model_from_t7(_roundtrip(t7.model_to_t7(spec)))
    _same_spec(back, spec)
    if which != "vgg16_fast_rcnn":
        assert [s for s, _ in back.towers[0].levels] == [back.taps["out1"], back.taps["out2"], back.taps["out3"]]
    rng = np.random.default_rng(4)
    img, rois = _inputs(rng, 64, 80, 4)
    ca, ba = G.heads_forward(spec, G.trunk_forward(spec, img), rois)
    cb, bb = G.heads_forward(back, G.trunk_forward(back, img), rois)
    assert np.array_equal(ca, cb) and np.array_equal(ba, bb)
