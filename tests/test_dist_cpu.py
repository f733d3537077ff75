"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, max-over-ranks
timing, bucketed gradient mean."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tf_ssd_b200 import dist_utils


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, w = dist_utils.init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    dev = torch.device("cpu")
    # sharding: every rank sees its slice of the same global batch
    g = np.arange(37)
    b, e = dist_utils.shard_range(len(g), rank, world)
    mine = torch.tensor(g[b:e].sum(), dtype=torch.float64)
    dist.all_reduce(mine)
    assert float(mine) == float(g.sum())
    # timing: the slowest rank wins
    assert dist_utils.max_over_ranks(1.0 + rank, dev) == float(world)
    # gradient buckets: mean over ranks, small bucket size to force several buckets
    shapes = [("a/kernel", (3, 3, 8, 16)), ("a/bias", (150,)), ("b/kernel", (1, 1, 16, 5)), ("c/kernel", (7,)), ("d/bias", (3,))]
    gb = dist_utils.GradBuckets(shapes, dev, bucket_bytes=2048)
    assert len(gb.buckets) >= 2
    rng = np.random.default_rng(5)          # same stream on every rank
    ref = {}
    for name, shape in shapes:
        per_rank = [rng.standard_normal(shape).astype(np.float32) for _ in range(world)]
        gb.views[name].copy_(torch.from_numpy(per_rank[rank]))
        ref[name] = np.mean(per_rank, axis=0)
    gb.allreduce_mean_()
    for name, _ in shapes:
        assert np.allclose(gb.views[name].numpy(), ref[name], rtol=1e-6, atol=1e-7), name
        assert gb.views[name].data_ptr() % 16 == 0
    dist_utils.barrier()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_range_properties():
    for n in (0, 1, 7, 32, 255, 256):
        for world in (1, 2, 3, 8):
            spans = [dist_utils.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_paths_are_noops():
    gb = dist_utils.GradBuckets([("w", (4, 4))], torch.device("cpu"))
    gb.views["w"].fill_(2.0)
    gb.allreduce_mean_()
    assert float(gb.views["w"].sum()) == 32.0
    assert dist_utils.max_over_ranks(3.5, torch.device("cpu")) == 3.5
