"""pipeline-wait counters of the 3x3 A-reuse kernel on the VGG conv layers (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multipathnet_b200 as mpn
ctx = mpn.Context(0)
names = ["p.wait_emptyA", "p.wait_emptyB", "p.total", "m.wait_fullA", "m.wait_fullB", "m.wait_tempty", "m.total", "e.wait_tfull", "e.store", "e.total"]
for (name, Cin, H, W, Cout) in [("conv1_2", 64, 600, 800, 64), ("conv2_1", 64, 300, 400, 128), ("conv2_2", 128, 300, 400, 128),
                                ("conv3_2", 256, 150, 200, 256), ("conv4_2", 512, 75, 100, 512), ("conv5_1", 512, 38, 50, 512)]:
    ms, bn, cg, mode, dbg = ctx.conv_bench(1, Cin, H, W, Cout)
    fl = 2.0 * Cin * Cout * 9 * H * W * 3
    print(name, f"ms={ms:.4f} BN={bn} CG={cg} mode={mode} issued_TF={fl / ms / 1e9:.0f}", {n: round(v / 1e3, 1) for n, v in zip(names, dbg)}, "(kcycles, CTA 0)")
