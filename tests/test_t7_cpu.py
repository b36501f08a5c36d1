"""CPU suite, part 5: Torch-7 file reader / writer (multipathnet_b200/t7.py) and the nn-graph importer.
PARITY UNPINNED against real files (no .t7 and no Torch in the image): round trips through this module's own writer,
hand-packed bytes for the cases the writer never produces (strides, offsets, shared storages), and a forward pass of an
imported graph through the CPU oracle against plain PyTorch on the same weights."""
import io
import struct

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from multipathnet_b200 import t7
from multipathnet_b200.t7 import T7Object
from oracle import graphs as G


def _roundtrip(o):
    buf = io.BytesIO()
    t7.save(buf, o)
    buf.seek(0)
    return t7.load(buf)


def test_scalars_tables_and_tensors_round_trip():
    rng = np.random.default_rng(0)
    obj = {"name": "vgg", "n": 3.0, "flag": True, "none_inside": [1.0, "two", False],
           "w": rng.standard_normal((4, 3, 2)).astype(np.float32), "idx": np.arange(5, dtype=np.int64),
           "d": rng.standard_normal((2, 2)), "b": np.array([1, 2, 255], np.uint8), "empty": np.zeros((0,), np.float32)}
    back = _roundtrip(obj)
    assert back["name"] == "vgg" and back["n"] == 3.0 and back["flag"] is True
    assert back["none_inside"] == [1.0, "two", False]
    for k in ("w", "idx", "d", "b"):
        assert back[k].dtype == obj[k].dtype and np.array_equal(back[k], obj[k])
    assert back["empty"].size == 0


def test_shared_objects_are_read_once():
    shared = np.arange(6, dtype=np.float32).reshape(2, 3)
    back = _roundtrip({"a": shared, "b": shared, "t": [shared]})
    assert back["a"] is back["b"] and back["t"][0] is back["a"]          # memoised by index, like torch.load


def _pack_tensor(typename, storage_type, idx, sizes, strides, offset, storage_idx, data=None):
    b = struct.pack("<ii", t7.TYPE_TORCH, idx)
    for s in ("V 1", typename):
        b += struct.pack("<i", len(s)) + s.encode()
    b += struct.pack("<i", len(sizes)) + b"".join(struct.pack("<q", s) for s in sizes) + b"".join(struct.pack("<q", s) for s in strides)
    b += struct.pack("<q", offset)
    b += struct.pack("<ii", t7.TYPE_TORCH, storage_idx)
    if data is not None:
        for s in ("V 1", storage_type):
            b += struct.pack("<i", len(s)) + s.encode()
        b += struct.pack("<q", data.size) + data.tobytes()
    return b


def test_strided_views_offsets_and_shared_storage():
    """What torch.save really writes for a transposed / narrowed tensor: the whole storage + sizes / strides / offset."""
    data = np.arange(24, dtype=np.float32)
    # table {t = storage viewed as 3 x 4 transposed (4 x 3, strides 1,4), n = narrow of the same storage from element 6}
    body = struct.pack("<iii", t7.TYPE_TABLE, 1, 2)
    body += struct.pack("<i", t7.TYPE_STRING) + struct.pack("<i", 1) + b"t"
    body += _pack_tensor("torch.FloatTensor", "torch.FloatStorage", 2, [4, 3], [1, 4], 1, 3, data)
    body += struct.pack("<i", t7.TYPE_STRING) + struct.pack("<i", 1) + b"n"
    body += _pack_tensor("torch.FloatTensor", "torch.FloatStorage", 4, [2, 3], [3, 1], 7, 3, None)      # storage 3 again: reference only
    back = t7.load(io.BytesIO(body))
    assert np.array_equal(back["t"], data[:12].reshape(3, 4).T)
    assert np.array_equal(back["n"], data[6:12].reshape(2, 3))


def test_truncated_file_is_an_error():
    buf = io.BytesIO()
    t7.save(buf, {"w": np.ones((10, 10), np.float32)})
    with pytest.raises(EOFError):
        t7.load(io.BytesIO(buf.getvalue()[:-17]))


def _conv(cin, cout, rng, cudnn=True):
    return T7Object("cudnn.SpatialConvolution" if cudnn else "nn.SpatialConvolution",
                    {"nInputPlane": cin, "nOutputPlane": cout, "kW": 3, "kH": 3, "dW": 1, "dH": 1, "padW": 1, "padH": 1, "groups": 1,
                     "weight": (rng.standard_normal((cout, cin, 3, 3)) * np.sqrt(2.0 / (cin * 9))).astype(np.float32),
                     "bias": (rng.standard_normal(cout) * 0.05).astype(np.float32)})


def _linear(cout, cin, rng, std=None):
    return T7Object("nn.Linear", {"weight": (rng.standard_normal((cout, cin)) * (std or np.sqrt(2.0 / cin))).astype(np.float32),
                                  "bias": (rng.standard_normal(cout) * 0.05).astype(np.float32)})


def _seq(*mods):
    return T7Object("nn.Sequential", {"modules": list(mods)})


def _tiny_fast_rcnn(rng, C=5):
    relu = lambda: T7Object("cudnn.ReLU", {"inplace": True})
    pool = T7Object("cudnn.SpatialMaxPooling", {"kW": 2, "kH": 2, "dW": 2, "dH": 2, "padW": 0, "padH": 0, "ceil_mode": True})
    c1, c2, c3 = _conv(3, 64, rng), _conv(64, 64, rng, cudnn=False), _conv(64, 128, rng)
    frozen = T7Object("nn.NoBackprop", {"modules": [_seq(c1, relu(), c2, relu(), pool)]})        # utils.disableFeatureBackprop
    features = T7Object("nn.DataParallelTable", {"modules": [_seq(frozen, c3, relu())]})         # utils.makeDataParallel
    fc6, fc7 = _linear(96, 128 * 3 * 3, rng), _linear(64, 96, rng)
    top = _seq(fc6, T7Object("nn.ReLU", {}), T7Object("nn.Dropout", {"p": 0.5, "v2": True}), fc7, T7Object("nn.ReLU", {}),
               T7Object("nn.Dropout", {"p": 0.5, "v2": True}))
    cls, bbox = _linear(C, 64, rng, 0.01), _linear(4 * C, 64, rng, 0.001)
    model = _seq(T7Object("nn.ParallelTable", {"modules": [features, T7Object("nn.Identity", {})]}),
                 T7Object("inn.ROIPooling", {"W": 3, "H": 3, "spatial_scale": 0.5}),
                 T7Object("nn.View", {"size": [-1], "numInputDims": 3}), top,
                 T7Object("nn.ConcatTable", {"modules": [cls, bbox]}),
                 T7Object("nn.ParallelTable", {"modules": [T7Object("nn.Identity", {}),
                                                           T7Object("nn.BBoxNorm", {"mean": np.array([[0.0, 0.01, 0.02, 0.03]], np.float32),
                                                                                    "std": np.array([[0.1, 0.1, 0.2, 0.2]], np.float32)})]}))
    return model, (c1, c2, c3, fc6, fc7, cls, bbox)


def test_import_fast_rcnn_graph_and_forward_through_the_oracle(oracle_built):
    rng = np.random.default_rng(5)
    model, (c1, c2, c3, fc6, fc7, cls, bbox) = _tiny_fast_rcnn(rng)
    spec = t7.fast_rcnn_from_t7(_roundtrip(model))                 # through the file format, not the in-memory objects
    assert spec.num_classes == 5 and spec.has_bbox_norm == 1 and spec.bbox_mean[1] == pytest.approx(0.01)
    assert [l.kind for l in spec.trunk_layers] == [1, 1, 2, 1] and [l.relu for l in spec.trunk_layers if l.kind == 1] == [1, 1, 1]
    assert spec.trunk_layers[2].ceil_mode == 1 and spec.towers[0].pooled_w == 3 and spec.towers[0].levels[0][1] == 0.5
    # forward: CPU oracle on the imported spec vs plain PyTorch on the original weights (trunk + heads, pre-softmax parity)
    img = rng.standard_normal((3, 24, 32)).astype(np.float32)
    x = torch.from_numpy(img)[None]
    x = F.relu(F.conv2d(x, torch.from_numpy(c1.weight), torch.from_numpy(c1.bias), padding=1))
    x = F.relu(F.conv2d(x, torch.from_numpy(c2.weight), torch.from_numpy(c2.bias), padding=1))
    x = F.max_pool2d(x, 2, 2, ceil_mode=True)
    x = F.relu(F.conv2d(x, torch.from_numpy(c3.weight), torch.from_numpy(c3.bias), padding=1))
    feats = G.trunk_forward(spec, img)
    assert np.allclose(feats[spec.taps["feat"]].numpy(), x.numpy(), atol=1e-5)
    boxes = np.array([[1, 1, 20, 16], [5, 3, 30, 22], [9, 9, 14, 13]], np.float32)
    scores, bb = G.detect(spec, img, boxes, 1.0)
    assert scores.shape == (3, 5) and bb.shape == (3, 20) and np.allclose(scores.sum(1), 1.0, atol=1e-5)


def test_importer_rejects_what_the_accelerated_path_does_not_run():
    rng = np.random.default_rng(1)
    model, (c1, *_r) = _tiny_fast_rcnn(rng)
    c1.fields["groups"] = 2
    with pytest.raises(NotImplementedError):
        t7.fast_rcnn_from_t7(model)
    with pytest.raises(ValueError):
        t7.fast_rcnn_from_t7(T7Object("nn.Linear", {}))


def test_proposal_file_boxes_are_permuted_like_DataSetJSON():
    """DataSetJSON.lua:157,234: stored y1,x1,y2,x2 -> index(2, {2,1,4,3}) = x1,y1,x2,y2"""
    stored = [np.array([[10, 20, 30, 40], [1, 2, 3, 4]], np.float32), np.zeros((0, 4), np.float32)]
    back = t7.proposals_from_t7(_roundtrip({"boxes": stored, "scores": [np.array([0.9, 0.1], np.float32), np.zeros(0, np.float32)],
                                            "images": ["a.jpg", "b.jpg"]}))
    assert np.array_equal(back["boxes"][0], np.array([[20, 10, 40, 30], [2, 1, 4, 3]], np.float32))
    assert back["boxes"][1].shape == (0, 4) and back["images"] == ["a.jpg", "b.jpg"]


def test_functions_are_skipped_and_file_paths_work(tmp_path):
    """multipathnet.lua:123 stores a Lua function in the model (`model.setPhase2 = ...`): torch.save writes its bytecode
    and upvalue table; the reader skips the bytecode, keeps the upvalues and the memo index (a second reference to the same
    function is just the index)."""
    s = lambda x: struct.pack("<i", t7.TYPE_STRING) + struct.pack("<i", len(x)) + x.encode()
    num = lambda v: struct.pack("<id", t7.TYPE_NUMBER, v)
    upvals = struct.pack("<iii", t7.TYPE_TABLE, 3, 1) + num(1.0) + s("captured")
    body = struct.pack("<iii", t7.TYPE_TABLE, 1, 4)
    body += s("setPhase2") + struct.pack("<iii", t7.TYPE_RECUR_FUNCTION, 2, 5) + b"\x1bLJ\x02\x00" + upvals
    body += s("again") + struct.pack("<ii", t7.TYPE_RECUR_FUNCTION, 2)
    body += s("plain") + struct.pack("<ii", t7.TYPE_FUNCTION, 3) + b"abc" + struct.pack("<i", t7.TYPE_NIL)
    body += s("phase") + num(1.0)
    p = tmp_path / "m.t7"
    p.write_bytes(body)
    back = t7.load(str(p))
    assert back["phase"] == 1.0 and back["setPhase2"].typename == "function" and back["again"] is back["setPhase2"]
    assert back["setPhase2"].upvalues == ["captured"] and back["plain"].upvalues is None
    t7.save(str(p), {"x": np.arange(3, dtype=np.float32)})
    assert np.array_equal(t7.load(str(p))["x"], [0, 1, 2])
    with pytest.raises(ValueError):
        t7.load(io.BytesIO(struct.pack("<i", 42)))
    with pytest.raises(AttributeError):
        T7Object("nn.Linear", {}).weight
    assert "nn.Linear" in repr(T7Object("nn.Linear", {"weight": 1}))


def _pack_obj(idx, version, cls, payload):
    b = struct.pack("<ii", t7.TYPE_TORCH, idx)
    for s in (version, cls):
        b += struct.pack("<i", len(s)) + s.encode()
    return b + payload


def _pack_table(idx, pairs):
    b = struct.pack("<iii", t7.TYPE_TABLE, idx, len(pairs))
    for k, v in pairs:
        b += (struct.pack("<id", t7.TYPE_NUMBER, float(k)) if not isinstance(k, str)
              else struct.pack("<ii", t7.TYPE_STRING, len(k)) + k.encode()) + v
    return b


def test_model_parallel_table_custom_serialisation():
    """ModelParallelTable.lua:607-628 (__write, __version 2): gpuAssignments, each branch as its own object, then the table of
    the remaining fields — packed by hand here, independently of this module's writer."""
    num = lambda v: struct.pack("<id", t7.TYPE_NUMBER, float(v))
    ident = lambda idx, tidx: _pack_obj(idx, "V 1", "nn.Identity", _pack_table(tidx, []))
    gpu = _pack_table(2, [(1, num(1)), (2, num(4))])
    rest = _pack_table(7, [("dimension", num(2)), ("noGradInput", struct.pack("<ii", t7.TYPE_BOOLEAN, 0))])
    mpt = _pack_obj(1, "V 2", "nn.ModelParallelTable", gpu + ident(3, 4) + ident(5, 6) + rest)
    o = t7.load(io.BytesIO(mpt))
    assert o.typename == "nn.ModelParallelTable" and o.dimension == 2.0 and o.noGradInput is False
    assert o.gpuAssignments == [1.0, 4.0] and [m.typename for m in o.modules] == ["nn.Identity"] * 2
    back = _roundtrip(o)                                              # this module's writer emits the same layout
    assert back.gpuAssignments == [1.0, 4.0] and len(back.modules) == 2 and back.dimension == 2.0
    buf = io.BytesIO(); t7.save(buf, o)
    assert b"V 2" in buf.getvalue()
    # a version-1 file holds the plain field table
    v1 = _pack_obj(1, "V 1", "nn.ModelParallelTable", _pack_table(2, [("dimension", num(2)), ("modules", _pack_table(3, [(1, ident(4, 5))]))]))
    o1 = t7.load(io.BytesIO(v1))
    assert o1.dimension == 2.0 and len(o1.modules) == 1
    # cunn's DataParallelTable (newer layout): gpuAssignments, then the field table that holds `modules` itself
    dpt = _pack_obj(1, "V 3", "nn.DataParallelTable", _pack_table(2, [(1, num(1))]) + _pack_table(3, [("dimension", num(1)), ("modules", _pack_table(4, [(1, ident(5, 6))]))]))
    od = t7.load(io.BytesIO(dpt))
    assert od.gpuAssignments == [1.0] and len(od.modules) == 1 and od.dimension == 1.0
    # ... and the older one: gpuAssignments, the replicas, the field table
    dpt2 = _pack_obj(1, "V 2", "nn.DataParallelTable", _pack_table(2, [(1, num(1)), (2, num(2))]) + ident(3, 4) + ident(5, 6) + _pack_table(7, [("dimension", num(1))]))
    od2 = t7.load(io.BytesIO(dpt2))
    assert len(od2.modules) == 2 and t7.flatten_sequential(od2)[0].typename == "nn.Identity"
    with pytest.raises(ValueError):                                   # three branches for two gpuAssignments, no field table
        t7.load(io.BytesIO(_pack_obj(1, "V 2", "nn.ModelParallelTable", _pack_table(2, [(1, num(1))]) + ident(3, 4) + ident(5, 6) + ident(7, 8))))


def test_nobackprop_legacy_inner_field():
    """NoBackprop.lua:34-46: files older than version 2 keep the wrapped module in `inner`"""
    inner = _pack_obj(3, "V 1", "nn.Identity", _pack_table(4, []))
    o = t7.load(io.BytesIO(_pack_obj(1, "V 1", "nn.NoBackprop", _pack_table(2, [("inner", inner)]))))
    assert [m.typename for m in o.modules] == ["nn.Identity"] and "inner" not in o.fields


def test_writer_memo_keeps_temporaries_alive(tmp_path):
    """the writer memoised by id(o) without holding o, so a freed temporary's id could be reused inside one save
    and a later, different table came back as an alias of an earlier one"""
    from multipathnet_b200 import t7
    import io
    w = t7._Writer(io.BytesIO())
    w.obj([1, 2]); w.obj([3, 4])                       # two temporaries that may share an id() once the first is freed
    w.f.seek(0)
    r = t7._Reader(w.f)
    a, b = r.obj(), r.obj()
    assert list(a.values() if isinstance(a, dict) else a) != list(b.values() if isinstance(b, dict) else b)
    tables = [t7.T7Object("nn.ModelParallelTable", {"gpuAssignments": [float(i), float(i + 1)], "modules": [], "dimension": 2.0}) for i in range(20)]
    p = tmp_path / "mpt.t7"
    t7.save(str(p), {"tables": tables})
    back = t7.load(str(p))["tables"]
    got = [list(x.gpuAssignments.values()) if isinstance(x.gpuAssignments, dict) else list(x.gpuAssignments) for x in (back.values() if isinstance(back, dict) else back)]
    assert got == [[float(i), float(i + 1)] for i in range(20)]
