"""CPU checks of bench.py's bookkeeping (no GPU): both arms print the SAME `config` dict (the driver compares them), the
issued-MMA accounting knows which Linears take two products, the workload table matches BASELINE.json's configs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from multipathnet_b200 import models  # noqa: E402


def test_both_arms_share_one_config_dict():
    a = bench.bench_config(1, 2)
    b = bench.bench_config(1, 2)
    assert a == b and a["workload"] == bench.WORKLOAD
    assert set(a) >= {"workload", "replicas_per_gpu", "parallelism", "l2", "nms_thr", "score_thresh", "roi_variant", "top_k_per_image"}
    assert bench.bench_config(8, 2)["parallelism"] != a["parallelism"]          # N is part of the config, identically in both arms
    json.dumps(a)


def test_w16_flops_are_fc6_fc7_of_single_tower_graphs():
    vgg = models.vgg16_fast_rcnn(21, seed=1, width_div=8, fc_dim=1024)
    k6 = (512 // 8) * 49
    assert models.w16_flops_per_roi(vgg) == 2.0 * k6 * 1024           # fc6 (K = 3136 >= 2048, 1024 outputs); fc7 has K = 1024 < 2048
    assert models.w16_flops_per_roi(vgg) < models.head_flops_per_roi(vgg)
    small = models.vgg16_fast_rcnn(21, seed=1, width_div=8, fc_dim=256)          # fc_dim < 1024: no w16 layer
    assert models.w16_flops_per_roi(small) == 0.0
    mpn = models.vgg16_multipathnet(21, seed=1, width_div=8, fc_dim=1024)        # multi-tower graphs default to three products
    assert models.w16_flops_per_roi(mpn) == 0.0


def test_workloads_follow_baseline_configs():
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert "VGG-16 Fast R-CNN" in cfgs[1] and "1000 ROIs" in cfgs[1]
    w = bench.WORKLOADS
    assert (w["vgg16_frcnn"]["H"], w["vgg16_frcnn"]["W"], w["vgg16_frcnn"]["R"], w["vgg16_frcnn"]["C"]) == (600, 800, 1000, 21)
    assert w["multipathnet"]["R"] == 1000 and w["multipathnet"]["boxes"] == "sharpmask"
    assert w["resnet50"]["R"] == 2000 and w["resnet50"]["kw"] == {"integral_k": 6}
