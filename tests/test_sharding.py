"""N>1 path on CPU: block partition of frame pairs and the all-gather of pose records over gloo, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam3d_gx_b200 import _abi, sharding


def test_partition_covers_everything_once():
    for n in (0, 1, 5, 64, 512, 513):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen.extend(sharding.partition(n, world, r))
            assert seen == list(range(n))
            sizes = [len(sharding.partition(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert len(sharding.partition(512, 8, 3)) == 64          # BASELINE config 4: 64 pairs per GPU


def _fake_result(i):
    r = _abi.Result()
    for k in range(16):
        r.T[k] = i * 100.0 + k
    r.norm, r.fitness, r.inliers, r.iterations, r.status = i + 0.5, i * 1e-3, 1000 + i, 10, i % 3
    return r


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [_fake_result(i) for i in sharding.partition(n_total, world, rank)]
    allr = sharding.gather_results(mine, n_total, dist)
    ok = len(allr) == n_total
    for i, r in enumerate(allr):
        ok &= r["inliers"] == 1000 + i and r["status"] == i % 3 and r["T"][3, 3] == i * 100.0 + 15 and abs(r["norm"] - (i + 0.5)) < 1e-12
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 8), (2, 7), (3, 10)])
def test_gather_pose_records_gloo(world, n_total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_record_round_trip():
    recs = [_fake_result(i) for i in range(5)]
    back = sharding.bytes_to_records(sharding.records_to_bytes(recs))
    assert [b["inliers"] for b in back] == [1000 + i for i in range(5)]
    assert _abi.RESULT_BYTES == 160
