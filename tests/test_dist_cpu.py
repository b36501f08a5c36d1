"""CPU suite, part 3: the N>1 path (image sharding + one all-gather of padded detections,
SURVEY 8e) exercised with world_size-2 gloo on CPU."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from multipathnet_b200 import dist as mdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_images = 7
    mine = mdist.shard_images(n_images, rank, world)
    dets = {}
    for i in mine:                                    # fake per-image detections: image i has (i % 3) + 1 of them
        k = (i % 3) + 1
        d = np.zeros((k, 6), np.float32); d[:, 4] = np.linspace(0.9, 0.5, k); d[:, 5] = i; d[:, 0] = i
        dets[i] = d
    allr = mdist.gather_detections(dets, n_images, rank, world, device="cpu")
    q.put((rank, mine, {i: allr[i].tolist() for i in allr}))
    dist.destroy_process_group()


def test_shard_and_allgather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    res.sort()
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]          # round-robin, test_runner.lua:91-104
    for _, _, allr in res:                                               # every rank ends with every image
        assert sorted(allr) == list(range(7))
        for i in range(7):
            a = np.array(allr[i], np.float32)
            assert a.shape == ((i % 3) + 1, 6) and np.all(a[:, 5] == i)


def test_pack_unpack_roundtrip():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from multipathnet_b200 import dist as mdist
    rng = np.random.default_rng(0)
    d = rng.random((150, 6)).astype(np.float32)
    rec = mdist.pack_record(d)                                           # top-100 by score (Tester_FRCNN.lua:163)
    out = mdist.unpack_record(rec)
    assert out.shape[0] == 100
    assert np.all(np.diff(out[:, 4]) <= 0)
    assert mdist.unpack_record(mdist.pack_record(np.zeros((0, 6), np.float32))).shape == (0, 6)
