"""conv1_1 tensor-core kernel: pipeline wait counters of CTA 0 (MPN_C1_TRACE=1 makes the kernel print them) + timing"""
import os, sys, time
os.environ["MPN_C1_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipathnet_b200 as mpn
ctx = mpn.Context(0)
rng = np.random.default_rng(0)
x = (rng.random((1, 3, 600, 800)) * 255 - 110).astype(np.float32)
w = (rng.standard_normal((64, 3, 3, 3)) / 8).astype(np.float32); b = rng.standard_normal(64).astype(np.float32)
y = ctx.conv_check(x, w, b, stride=1, pad=1, relu=True, impl=2)
ctx.synchronize()
print("out", y.shape, float(np.abs(y).mean()))
