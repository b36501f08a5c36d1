"""Test infrastructure: a tiny evaluator of torch `nn` graphs held as T7Object trees (what multipathnet_b200.t7.load
returns), in eval mode, on PyTorch CPU fp32 — tables are Python lists.  It is the independent side of the importer
tests: evaluate(graph) must equal the CPU oracle run on model_from_t7(graph).  Module semantics restated from torch/nn,
imagine-nn and the reference's modules/ (Foveal.lua, BBoxNorm.lua, ModeSwitch.lua:16-20, ModelParallelTable.lua:195-242
concat along `dimension`); ROI pooling and Foveal come from the C oracle."""
import numpy as np
import torch
import torch.nn.functional as F

from multipathnet_b200.t7 import T7Object, _base, _children
from oracle import ref as O


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def evaluate(m: T7Object, x):
    b = _base(m.typename)
    kids = _children(m)
    if b in ("Sequential", "NoBackprop"):
        for c in kids:
            x = evaluate(c, x)
        return x
    if b == "DataParallelTable":
        return evaluate(kids[0], x)
    if b == "ConcatTable":
        return [evaluate(c, x) for c in kids]
    if b == "ParallelTable":
        assert len(kids) == len(x)
        return [evaluate(c, xi) for c, xi in zip(kids, x)]
    if b == "ModelParallelTable":
        return torch.cat([evaluate(c, x) for c in kids], dim=int(m.dimension) - 1)
    if b == "ModeSwitch":
        return evaluate(kids[1], x)                              # eval mode
    if b == "FlattenTable":
        flat = lambda v: [y for e in v for y in flat(e)] if isinstance(v, list) else [v]
        return flat(x)
    if b == "SelectTable":
        return x[int(m.index) - 1]
    if b in ("Identity", "Copy", "Contiguous", "Dropout"):
        return x
    if b in ("SpatialConvolution", "SpatialConvolutionMM"):
        bias = None if m.get("bias") is None else _t(m.bias)
        return F.conv2d(x, _t(m.weight), bias, stride=int(m.dW), padding=int(m.get("padW", 0)))
    if b == "SpatialBatchNormalization":
        return F.batch_norm(x, _t(m.running_mean), _t(m.running_var), _t(m.weight), _t(m.bias), False, 0.0, float(m.eps))
    if b == "ConstAffine":
        return x * _t(m.a).view(1, -1, 1, 1) + _t(m.b).view(1, -1, 1, 1)
    if b == "ReLU":
        return F.relu(x)
    if b == "SpatialMaxPooling":
        return F.max_pool2d(x, int(m.kW), int(m.dW), int(m.get("padW", 0)), ceil_mode=bool(m.get("ceil_mode", False)))
    if b == "SpatialAveragePooling":
        return F.avg_pool2d(x, int(m.kW), int(m.dW))
    if b == "CAddTable":
        return x[0] + x[1]
    if b == "ROIPooling":
        data, rois = x
        return _t(O.roi_pool(data.numpy(), rois.numpy(), int(m.W), int(m.H), np.float32(m.spatial_scale), 2))
    if b == "Foveal":
        return _t(O.foveal(x.numpy()))
    if b == "View":
        size = [int(s) for s in m.size]
        nd = m.get("numInputDims")
        if nd is not None and x.dim() == int(nd) + 1:            # setNumInputDims: a leading batch dimension is kept
            return x.reshape([x.shape[0]] + size)
        return x.reshape(size)
    if b == "Transpose":
        for p in m.permutations:
            x = x.transpose(int(p[0]) - 1, int(p[1]) - 1)
        return x
    if b == "Select":
        return x.select(int(m.dimension) - 1, int(m.index) - 1)
    if b == "Narrow":
        return x.narrow(int(m.dimension) - 1, int(m.index) - 1, int(m.length))
    if b == "Normalize":
        return _t(O.l2_normalize(x.contiguous().numpy()))
    if b == "MulConstant":
        return x * np.float32(m.constant_scalar)
    if b == "JoinTable":
        return torch.cat(list(x), dim=int(m.dimension) - 1)
    if b == "Mean":
        return x.mean(dim=int(m.dimension) - 1)
    if b == "Linear":
        return F.linear(x, _t(m.weight), None if m.get("bias") is None else _t(m.bias))
    if b == "SoftMax":
        return _t(O.softmax(x.numpy()))
    if b == "BBoxNorm":
        return _t(O.bbox_norm(x.numpy(), np.asarray(m.mean).reshape(-1), np.asarray(m.std).reshape(-1)))
    raise NotImplementedError(m.typename)
