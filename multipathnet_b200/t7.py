"""Torch-7 binary serialisation (`torch.save` / `torch.load`, torch7 File.lua) without Torch — SURVEY 8f-4.

The reference keeps its models and proposals in `.t7` files (`test_runner.lua:31`: `torch.load(opt.test_model)`;
proposals `{boxes, scores, images}`, `DataSetJSON.lua:124-239`). This module reads that format into plain Python / numpy
so real weights and proposals can be fed to the B200 path, and writes it (the subset below) so the reader can be tested
without Torch. PARITY UNPINNED: no `.t7` file and no Torch exist in the build image, so the format is restated from
torch7's public File.lua / Tensor / Storage `read`-`write` methods and pinned only by round trips through this module's
own writer (tests/test_t7_cpu.py); the first real file is the real test.

Format (little-endian, "binary" mode): every object starts with an int32 type tag
  0 nil | 1 number (float64) | 2 string (int32 length + bytes) | 3 table | 4 torch object | 5 boolean (int32)
  | 6/7/8 function (skipped: int32 size + bytecode + upvalue table).
Tables and torch objects are memoised: an int32 index follows the tag, and a repeated index means "the same object".
  table : int32 n, then n x (key object, value object)
  torch : version string "V 1" (int32 length + bytes), class name string, then the class payload:
          torch.*Tensor  : int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storage offset (1-based), storage object
          torch.*Storage : int64 n, then n raw elements
          anything else (nn modules, ...): one object (a table) whose pairs become the fields.
"""
from __future__ import annotations

import struct
from typing import Any, BinaryIO, Dict, List

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN, TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = range(9)

_DTYPES = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Short": np.int16,
           "Byte": np.uint8, "Char": np.int8, "CudaTensor": np.float32, "Cuda": np.float32, "Half": np.float16}


class T7Object:
    """A torch class instance that is not a tensor / storage (nn modules, ...): `typename` + its fields."""

    def __init__(self, typename: str, fields: Dict[Any, Any] = None):
        self.typename = typename
        self.fields = {} if fields is None else fields

    def __getattr__(self, k):
        f = self.__dict__.get("fields", {})
        if k in f:
            return f[k]
        raise AttributeError(f"{self.__dict__.get('typename')} has no field {k!r}")

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def __repr__(self):
        return f"T7Object({self.typename}, fields={list(self.fields)})"


def _as_list_or_dict(d: Dict[Any, Any]):
    """A Lua table with keys 1..n (and nothing else) reads as a list, anything else as a dict."""
    n = len(d)
    if n and all(isinstance(k, float) and k == int(k) for k in d) and sorted(int(k) for k in d) == list(range(1, n + 1)):
        return [d[float(i)] for i in range(1, n + 1)]
    return {(int(k) if isinstance(k, float) and k == int(k) else k): v for k, v in d.items()}


class _Reader:
    def __init__(self, f: BinaryIO):
        self.f = f
        self.memo: Dict[int, Any] = {}

    def _read(self, fmt: str):
        size = struct.calcsize(fmt)
        b = self.f.read(size)
        if len(b) != size:
            raise EOFError("truncated .t7 file")
        return struct.unpack(fmt, b)

    def int32(self) -> int:
        return self._read("<i")[0]

    def int64(self) -> int:
        return self._read("<q")[0]

    def string(self) -> str:
        n = self.int32()
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("truncated .t7 string")
        return b.decode("latin-1")

    def obj(self):
        t = self.int32()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            return self._read("<d")[0]
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_BOOLEAN:
            return self.int32() != 0
        if t in (TYPE_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            idx = self.int32() if t != TYPE_FUNCTION else None
            if idx is not None and idx in self.memo:
                return self.memo[idx]
            n = self.int32()
            self.f.read(n)                       # Lua bytecode: of no use here
            fn = T7Object("function")
            if idx is not None:
                self.memo[idx] = fn
            fn.fields["upvalues"] = self.obj()
            return fn
        if t == TYPE_TABLE:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            holder: Dict[Any, Any] = {}
            self.memo[idx] = holder              # cycles resolve to the raw dict
            n = self.int32()
            for _ in range(n):
                k = self.obj()
                holder[k] = self.obj()
            out = _as_list_or_dict(holder)
            self.memo[idx] = out
            return out
        if t == TYPE_TORCH:
            idx = self.int32()
            if idx in self.memo:
                return self.memo[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version
            if cls.startswith("torch.") and cls.endswith("Storage"):
                dt = _DTYPES[cls[len("torch."):-len("Storage")]]
                n = self.int64()
                raw = self.f.read(n * np.dtype(dt).itemsize)
                if len(raw) != n * np.dtype(dt).itemsize:
                    raise EOFError("truncated .t7 storage")
                a = np.frombuffer(raw, dtype=dt).copy()
                self.memo[idx] = a
                return a
            if cls.startswith("torch.") and cls.endswith("Tensor"):
                nd = self.int32()
                size = [self.int64() for _ in range(nd)]
                stride = [self.int64() for _ in range(nd)]
                off = self.int64() - 1
                storage = self.obj()
                key = cls[len("torch."):-len("Tensor")] or "Float"
                dt = _DTYPES.get(key, np.float32)
                if storage is None or nd == 0:
                    a = np.zeros((0,), dt)
                else:
                    item = storage.dtype.itemsize
                    a = np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[s * item for s in stride]).copy()
                self.memo[idx] = a
                return a
            o = T7Object(cls)
            self.memo[idx] = o
            payload = self.obj()
            if isinstance(payload, dict):
                o.fields.update(payload)
            elif isinstance(payload, list):
                o.fields.update({i + 1: v for i, v in enumerate(payload)})
            return o
        raise ValueError(f"unknown .t7 type tag {t}")


def load(path_or_file) -> Any:
    """torch.load(path) -> numbers (float), str, bool, list / dict (Lua tables), numpy arrays (tensors, storages),
    T7Object (other torch classes, e.g. nn modules)."""
    if hasattr(path_or_file, "read"):
        return _Reader(path_or_file).obj()
    with open(path_or_file, "rb") as f:
        return _Reader(f).obj()


# ------------------------------------------------------------------------------------------ writer (tests, fixtures)
_TENSOR_NAMES = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long",
                 np.dtype(np.int32): "Int", np.dtype(np.uint8): "Byte", np.dtype(np.int16): "Short", np.dtype(np.int8): "Char"}


class _Writer:
    def __init__(self, f: BinaryIO):
        self.f = f
        self.next_idx = 1
        self.seen: Dict[int, int] = {}

    def i32(self, v):
        self.f.write(struct.pack("<i", int(v)))

    def i64(self, v):
        self.f.write(struct.pack("<q", int(v)))

    def string(self, s: str):
        b = s.encode("latin-1")
        self.i32(len(b)); self.f.write(b)

    def _index(self, o) -> bool:
        """writes the memo index; True if the object was written before (nothing more to emit)"""
        if id(o) in self.seen:
            self.i32(self.seen[id(o)])
            return True
        self.seen[id(o)] = self.next_idx
        self.i32(self.next_idx)
        self.next_idx += 1
        return False

    def obj(self, o):
        if o is None:
            self.i32(TYPE_NIL)
        elif isinstance(o, bool):
            self.i32(TYPE_BOOLEAN); self.i32(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.i32(TYPE_NUMBER); self.f.write(struct.pack("<d", float(o)))
        elif isinstance(o, str):
            self.i32(TYPE_STRING); self.string(o)
        elif isinstance(o, np.ndarray):
            self.i32(TYPE_TORCH)
            if self._index(o):
                return
            name = _TENSOR_NAMES[o.dtype]
            self.string("V 1"); self.string(f"torch.{name}Tensor")
            a = np.ascontiguousarray(o)
            self.i32(a.ndim)
            for s in a.shape:
                self.i64(s)
            for s in a.strides:
                self.i64(s // a.itemsize)
            self.i64(1)
            if a.ndim == 0 or a.size == 0:
                self.i32(TYPE_NIL)
            else:
                self.i32(TYPE_TORCH); self.i32(self.next_idx); self.next_idx += 1
                self.string("V 1"); self.string(f"torch.{name}Storage")
                self.i64(a.size); self.f.write(a.tobytes())
        elif isinstance(o, T7Object):
            self.i32(TYPE_TORCH)
            if self._index(o):
                return
            self.string("V 1"); self.string(o.typename)
            self._table(o.fields, fresh_index=True)
        elif isinstance(o, (list, tuple)):
            self.i32(TYPE_TABLE)
            if self._index(o):
                return
            self.i32(len(o))
            for i, v in enumerate(o):
                self.obj(i + 1); self.obj(v)
        elif isinstance(o, dict):
            self.i32(TYPE_TABLE)
            if self._index(o):
                return
            self.i32(len(o))
            for k, v in o.items():
                self.obj(k); self.obj(v)
        else:
            raise TypeError(f"cannot serialise {type(o)} to .t7")

    def _table(self, d: dict, fresh_index: bool):
        self.i32(TYPE_TABLE)
        self.i32(self.next_idx); self.next_idx += 1
        self.i32(len(d))
        for k, v in d.items():
            self.obj(k); self.obj(v)


def save(path_or_file, obj) -> None:
    """torch.save(path, obj) for numbers, strings, booleans, lists / dicts, numpy arrays and T7Object."""
    if hasattr(path_or_file, "write"):
        _Writer(path_or_file).obj(obj)
        return
    with open(path_or_file, "wb") as f:
        _Writer(f).obj(obj)


# ------------------------------------------------------------------------------------------ nn graph -> ModelSpec
def _base(typename: str) -> str:
    return typename.split(".", 1)[-1]


def _children(m: T7Object) -> List[Any]:
    mods = m.get("modules")
    if mods is None and m.get("module") is not None:
        mods = [m.get("module")]
    if isinstance(mods, dict):
        mods = [mods[k] for k in sorted(mods)]
    return list(mods or [])


def flatten_sequential(m) -> List[T7Object]:
    """Depth-first list of the leaf modules of nested nn.Sequential / nn.NoBackprop / nn.DataParallelTable containers
    (utils.disableFeatureBackprop and makeDataParallel wrap parts of the trunk: model_utils.lua:95-103, vgg.lua:18-27)."""
    if not isinstance(m, T7Object):
        raise ValueError("not a torch object")
    b = _base(m.typename)
    if b in ("Sequential", "NoBackprop"):
        out: List[T7Object] = []
        for c in _children(m):
            out += flatten_sequential(c)
        return out
    if b in ("DataParallelTable", "DataParallel"):
        kids = _children(m)
        return flatten_sequential(kids[0]) if kids else []
    return [m]


def fast_rcnn_from_t7(model, num_classes: int = None, name: str = "t7"):
    """The graph `models/vgg.lua:23-31` (or alexnet / any trunk of conv / ReLU / max-pool) returns, as saved by train.lua,
    -> ModelSpec:  Sequential{ ParallelTable{trunk, Identity}, inn.ROIPooling(W,H,s), View, top (Linear/ReLU/Dropout...),
    ConcatTable{Linear cls, Linear bbox} [, BBoxNorm / SoftMax added at test time] }.
    Grouped convolutions, LRN and the MultiPathNet / ResNet graphs (ModelParallelTable towers, residual blocks) are not
    covered here: those specs are built by multipathnet_b200.models from the same Lua files."""
    from ._lib import Head, Layer, ModelSpec, Tower, MPN_LAYER_CONV, MPN_LAYER_FLATTEN, MPN_LAYER_MAXPOOL
    if not isinstance(model, T7Object) or _base(model.typename) != "Sequential":
        raise ValueError("expected the nn.Sequential detection model")
    top_mods = _children(model)
    if not top_mods or _base(top_mods[0].typename) != "ParallelTable":
        raise ValueError("expected nn.ParallelTable{trunk, Identity} first (vgg.lua:23-27)")
    trunk_mods = flatten_sequential(_children(top_mods[0])[0])
    arrays: List[np.ndarray] = []

    def add(a):
        arrays.append(np.ascontiguousarray(a, np.float32))
        return len(arrays) - 1

    trunk: List[Layer] = []
    slot, cin = 0, 3
    for m in trunk_mods:
        b = _base(m.typename)
        if b == "SpatialConvolution" or b == "SpatialConvolutionMM":
            if int(m.get("groups", 1) or 1) != 1:
                raise NotImplementedError("grouped convolution (CaffeNet) is not on the accelerated path")
            cout, cin_m = int(m.nOutputPlane), int(m.nInputPlane)
            kh, kw = int(m.kH), int(m.kW)
            if int(m.dW) != int(m.dH) or int(m.get("padW", 0)) != int(m.get("padH", 0)):
                raise NotImplementedError("anisotropic stride / padding")
            if cin_m != cin:
                raise ValueError(f"conv expects {cin_m} input planes, trunk has {cin}")
            w = np.asarray(m.weight, np.float32).reshape(cout, cin_m, kh, kw)
            trunk.append(Layer(MPN_LAYER_CONV, slot, slot + 1, cin=cin, cout=cout, kh=kh, kw=kw, stride=int(m.dW),
                               pad=int(m.get("padW", 0)), relu=0, weight=add(w), bias=add(np.asarray(m.bias, np.float32).reshape(cout))))
            cin = cout
            slot += 1
        elif b == "ReLU":
            if not trunk or trunk[-1].kind != MPN_LAYER_CONV or trunk[-1].out_slot != slot:
                raise NotImplementedError("ReLU that does not follow a convolution")
            trunk[-1].relu = 1
        elif b == "SpatialMaxPooling":
            if int(m.kW) != int(m.kH) or int(m.dW) != int(m.dH):
                raise NotImplementedError("anisotropic pooling")
            trunk.append(Layer(MPN_LAYER_MAXPOOL, slot, slot + 1, kh=int(m.kH), kw=int(m.kW), stride=int(m.dW),
                               pad=int(m.get("padW", 0)), ceil_mode=1 if m.get("ceil_mode", False) else 0))
            slot += 1
        elif b in ("Dropout", "Identity", "Copy"):
            continue
        else:
            raise NotImplementedError(f"trunk module {m.typename}")
    feat_slot, c5 = slot, cin

    rest = top_mods[1:]
    if not rest or _base(rest[0].typename) != "ROIPooling":
        raise ValueError("expected inn.ROIPooling after the trunk (vgg.lua:28)")
    roi = rest[0]
    pw, ph, scale = int(roi.W), int(roi.H), float(roi.spatial_scale)
    tl = [Layer(MPN_LAYER_FLATTEN, 0, 1)]
    tslot, k_in = 1, c5 * pw * ph
    heads = None
    bbox_mean, bbox_std, has_norm = (0.0, 0.0, 0.0, 0.0), (0.1, 0.1, 0.2, 0.2), 0
    mods: List[T7Object] = []
    for m in rest[1:]:
        mods += flatten_sequential(m) if _base(m.typename) == "Sequential" else [m]
    for m in mods:
        b = _base(m.typename)
        if b in ("View", "Reshape", "Identity", "Copy"):
            continue
        if b == "Dropout":
            if m.get("v2", True) is False:
                raise NotImplementedError("nn.Dropout(v2=false) scales at test time")
            continue
        if b == "Linear" and heads is None:
            w = np.asarray(m.weight, np.float32)
            if w.shape[1] != k_in:
                raise ValueError(f"Linear expects {w.shape[1]} inputs, tower has {k_in}")
            tl.append(Layer(MPN_LAYER_CONV, tslot, tslot + 1, cin=k_in, cout=w.shape[0], relu=0, weight=add(w),
                            bias=add(np.asarray(m.bias, np.float32).reshape(-1))))
            k_in = w.shape[0]
            tslot += 1
        elif b == "ReLU" and heads is None:
            tl[-1].relu = 1
        elif b == "ConcatTable":
            kids = _children(m)
            if len(kids) != 2 or any(_base(k.typename) != "Linear" for k in kids):
                raise NotImplementedError("expected ConcatTable{Linear cls, Linear bbox} (model_utils.lua:105-119)")
            heads = kids
        elif b == "ParallelTable" and heads is not None:
            for k in _children(m):        # test-time tail: {SoftMax, BBoxNorm} on {cls, bbox}
                kb = _base(k.typename)
                if kb == "BBoxNorm":
                    has_norm = 1
                    bbox_mean = tuple(float(x) for x in np.asarray(k.mean).reshape(-1)[:4])
                    bbox_std = tuple(float(x) for x in np.asarray(k.std).reshape(-1)[:4])
        elif b == "SoftMax":
            continue                      # detect() applies the softmax itself (ImageDetect.lua:186-190)
        elif b == "BBoxNorm":
            has_norm = 1
            bbox_mean = tuple(float(x) for x in np.asarray(m.mean).reshape(-1)[:4])
            bbox_std = tuple(float(x) for x in np.asarray(m.std).reshape(-1)[:4])
        else:
            raise NotImplementedError(f"head module {m.typename}")
    if heads is None:
        raise ValueError("no ConcatTable{cls, bbox} head found")
    wc, wb = np.asarray(heads[0].weight, np.float32), np.asarray(heads[1].weight, np.float32)
    C = wc.shape[0]
    if num_classes is not None and C != num_classes:
        raise ValueError(f"class head has {C} outputs, expected {num_classes}")
    if wb.shape[0] != 4 * C or wc.shape[1] != k_in or wb.shape[1] != k_in:
        raise ValueError("class / bbox head sizes do not match the tower")
    cls_head = Head(0, k_in, C, add(wc), add(np.asarray(heads[0].bias, np.float32).reshape(-1)))
    bbox_head = Head(0, k_in, 4 * C, add(wb), add(np.asarray(heads[1].bias, np.float32).reshape(-1)))
    tower = Tower(region=0, levels=[(feat_slot, scale)], pooled_w=pw, pooled_h=ph, normalize=0, layers=tl, out_slot=tslot)
    # ImageDetect applies SoftMax itself unless model.noSoftMax (ImageDetect.lua:186-190): a saved training graph has none
    return ModelSpec(name=name, trunk_layers=trunk, towers=[tower], cls_heads=[cls_head], bbox_head=bbox_head, num_classes=C,
                     weights=arrays, roi_variant=2, no_softmax=0, has_bbox_norm=has_norm,
                     bbox_mean=bbox_mean, bbox_std=bbox_std, transformer="ross", taps={"feat": feat_slot})


def proposals_from_t7(obj) -> Dict[str, Any]:
    """A proposal file `{boxes = {[i] = N_i x 4 (y1, x1, y2, x2)}, scores = {...}, images = {...}}`
    (DataSetJSON.lua:124-239, utils.lua:305-372) -> {'boxes': [N_i x 4 float32 in x1,y1,x2,y2], 'scores': [...], 'images': [...]}."""
    if not isinstance(obj, dict) or "boxes" not in obj:
        raise ValueError("expected a table with a `boxes` field")
    boxes = obj["boxes"]
    if isinstance(boxes, dict):
        boxes = [boxes[k] for k in sorted(boxes)]
    out_boxes = []
    for b in boxes:
        b = np.asarray(b, np.float32).reshape(-1, 4)
        out_boxes.append(b[:, [1, 0, 3, 2]].copy())            # y1,x1,y2,x2 -> x1,y1,x2,y2
    res = {"boxes": out_boxes}
    for k in ("scores", "images"):
        if k in obj:
            v = obj[k]
            res[k] = [v[i] for i in sorted(v)] if isinstance(v, dict) else v
    return res
