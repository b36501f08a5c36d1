"""Multi-GPU: the reference's only inference strategy — one replica per GPU, images dealt round-robin
(test_runner.lua:55-66,91-104) — as one process per GPU over torch.distributed, plus ONE all-gather of
fixed-size padded detection records at the end. No data-path collective: the path shards by image."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

MAX_DET = 128          # reference cap is top-100 per image (Tester_FRCNN.lua:163); ties may exceed => pad to 128
REC = 1 + MAX_DET * 6  # [count, MAX_DET x (x1,y1,x2,y2,score,class)]


def shard_images(n_images: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_images, world))


def pack_record(dets: np.ndarray, top_k: int = 100) -> np.ndarray:
    """dets K x 6 [x1,y1,x2,y2,score,class], class-major as Tester:testOne returns them -> the fixed-size record of
    include/mpn_abi.h: utils.keep_top_k's rule (utils.lua:75-96: every row with score >= the top_k-th largest score, so ties
    at the cut survive and the count may exceed top_k), row order preserved. Host mirror of pack_detections_kernel
    (csrc/post.cu); raises when the ties overflow MAX_DET."""
    rec = np.zeros(REC, np.float32)
    if dets.shape[0]:
        s = np.sort(dets[:, 4])[::-1]
        thresh = s[min(len(s), top_k) - 1]
        d = dets[dets[:, 4] >= thresh]
        if d.shape[0] > MAX_DET:
            raise OverflowError(f"{d.shape[0]} detections tie into the top {top_k}: more than MAX_DET = {MAX_DET}")
        rec[0] = d.shape[0]
        rec[1:1 + d.size] = d.reshape(-1)
    return rec


def unpack_record(rec: np.ndarray) -> np.ndarray:
    k = int(rec[0])
    if k > MAX_DET:
        raise OverflowError(f"detection record overflowed: {k} rows survive keep_top_k, capacity {MAX_DET}")
    return rec[1:1 + 6 * k].reshape(k, 6).copy()


def tables_to_dets(img_boxes: List[np.ndarray]) -> np.ndarray:
    """Tester:testOne's per-class tables (K_j x 5, class j = 1..C-1) -> K x 6 rows with the class in column 6"""
    rows = [np.concatenate([np.asarray(b, np.float32).reshape(-1, 5), np.full((len(b), 1), j, np.float32)], 1)
            for j, b in enumerate(img_boxes, start=1) if len(b)]
    return np.concatenate(rows, 0) if rows else np.zeros((0, 6), np.float32)


def record_to_tables(rec: np.ndarray, num_classes: int) -> List[np.ndarray]:
    """inverse of the packing: the per-class tables after keep_top_k (what Tester:keepTopKPerImage leaves, :163-168)"""
    d = unpack_record(rec)
    return [d[d[:, 5] == j, :5].copy() for j in range(1, num_classes)]


def gather_records_dev(ctx, records_dev, n_records: int) -> np.ndarray:
    """the product collective: this rank's `n_records` packed records (device) -> world x n_records x REC on the host,
    through mpn_dist_all_gather (ncclAllGather issued by libmpn_b200.so; a copy in a world of one)"""
    return ctx.dist_all_gather(records_dev, n_records * REC).reshape(-1, n_records, REC)


def gather_detections(dets: Dict[int, np.ndarray], n_images: int, rank: int, world: int, device="cuda") -> Dict[int, np.ndarray]:
    """every rank contributes the records of its images; one all_gather; every rank returns all images."""
    import torch
    import torch.distributed as dist
    per_rank = (n_images + world - 1) // world
    buf = np.zeros((per_rank, REC), np.float32)
    mine = shard_images(n_images, rank, world)
    for slot, i in enumerate(mine):
        buf[slot] = pack_record(dets[i])
    t = torch.from_numpy(buf).to(device)
    if world > 1:
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)                     # THE collective of the path (NCCL on GPUs, gloo in CPU tests)
    else:
        out = [t]
    res = {}
    for r in range(world):
        a = out[r].cpu().numpy()
        for slot, i in enumerate(shard_images(n_images, r, world)):
            res[i] = unpack_record(a[slot])
    return res
