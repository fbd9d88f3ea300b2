"""CPU, world_size 2, gloo: the N>1 path -- interleaved sharding of worlds over ranks, independent stepping of
each shard (here by the CPU oracle standing in for the kernels: no GPU in this tier) and the single end-of-run
gather -- reproduces the single-process result in global world order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _results_for(indices, n_total, steps):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fizz2d_b200 as fz
    from fizz2d_b200.batch import GroupSpec
    from helpers import oracle_from_spec
    sc = fz.load_in(fz.scenes.scene_path("2CHAMBERS_3BALLS"))
    deltas = fz.perturbation_deltas(20260004, n_total, sc.n_free)[indices]
    spec = GroupSpec.from_scene(sc, len(indices), np.ascontiguousarray(deltas), max_calls=128)
    rows = np.zeros((6, len(indices)))
    for k in range(len(indices)):
        ow = oracle_from_spec(spec, k)
        ow.run(steps)
        rows[:4, k] = ow.energy()
        rows[4, k] = ow.heat()
        rows[5, k] = ow.status()
    return rows


def _worker(rank, world_size, port, n_total, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    sys.path.insert(0, ROOT)
    from fizz2d_b200 import distributed as fd
    idx = fd.shard_indices(n_total, rank, world_size)
    assert len(idx) == fd.shard_count(n_total, rank, world_size)
    local = torch.from_numpy(_results_for(idx, n_total, steps))
    full = fd.gather_results(local, n_total)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_two_rank_gather_matches_single_process(tmp_path, n_total):
    steps = 60
    mp.spawn(_worker, args=(2, _free_port(), n_total, steps, str(tmp_path)), nprocs=2, join=True)
    want = _results_for(np.arange(n_total), n_total, steps)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))
        assert got.shape == (6, n_total)
        assert np.array_equal(got.view(np.int64), want.view(np.int64))


def test_shards_partition_the_batch():
    sys.path.insert(0, ROOT)
    from fizz2d_b200 import distributed as fd
    for n, g in ((10, 4), (65536, 8), (5, 8)):
        allidx = np.concatenate([fd.shard_indices(n, r, g) for r in range(g)])
        assert sorted(allidx.tolist()) == list(range(n))
        assert sum(fd.shard_count(n, r, g) for r in range(g)) == n
