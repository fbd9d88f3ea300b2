"""Worker of tests/test_gpu_multi.py (one process per GPU under torchrun, NCCL): every rank steps ITS shard of a perturbed
RAMPWORLD population (world w of the global batch lives on rank w % G) with the CUDA engine, checks a sample of its worlds
against the CPU oracle bit for bit, then all ranks exchange [E,K,P,H,fitness,status] with the path's one collective and
every rank checks the gathered table against what it and (through a second, independent all_gather of the raw state) the
others computed."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fizz2d_b200 as fz
    from fizz2d_b200 import distributed as fd
    from fizz2d_b200.batch import GroupSpec
    from helpers import bits, oracle_from_spec, orc
    n_per, steps = 2048, 400
    total = n_per * world
    sc = fz.load_in(fz.scenes.scene_path("RAMPWORLD_2TRIANGLES_5SQUARES"))
    idx = fd.shard_indices(total, rank, world)
    deltas = np.ascontiguousarray(fz.perturbation_deltas(20260003, total, sc.n_free)[idx])
    spec = GroupSpec.from_scene(sc, n_per, deltas, max_calls=512)
    bw = fz.BatchedWorld([spec], device=local)
    bw.set_tuning("hybrid", 64)
    bw.step(steps)
    com, heat, st, sd, calls = bw.com(), bw.heat(), bw.status(), bw.steps_done(), bw.calls()
    # (1) this rank's shard against the oracle: a spread-out sample, every field
    sample = list(range(rank, n_per, 16))
    worlds = [oracle_from_spec(spec, w) for w in sample]
    orc.run_batch(worlds, steps, max(1, (os.cpu_count() or 2) // world))
    for w, ow in zip(sample, worlds):
        assert ow.steps_done() == sd[w] and ow.status() == st[w] and ow.total_calls() == calls[w], (rank, w)
        assert np.array_equal(bits(ow.com()), bits(com[w])) and bits(ow.heat()) == bits(heat[w]), (rank, w)
    # (2) the collective: global table in world order on every rank
    g = bw.groups[0]
    fz.lib.check(fz.lib.load().fizz_energy(g.handle, torch.cuda.current_stream().cuda_stream))
    mine = fd.pack_results(g.section("energy", 4), g.section("heat", 1)[0], g.section("status", 1, torch.int64)[0])
    full = fd.gather_results(mine, total)
    assert full.shape == (fd.RESULT_ROWS, total)
    dev = mine.device
    assert torch.equal(full[:, torch.from_numpy(idx).to(dev)].view(torch.int64), mine.view(torch.int64))
    # every other rank's columns hold what THAT rank computed (cross-check through an independent all_gather of heat)
    heats = [torch.empty(n_per, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(heats, g.section("heat", 1)[0].contiguous())
    for r in range(world):
        cols = torch.from_numpy(fd.shard_indices(total, r, world)).to(dev)
        assert torch.equal(full[4, cols].view(torch.int64), heats[r].view(torch.int64)), (rank, r)
    # world 0 of the global batch is the unperturbed scene: rank 0 owns it, everybody sees its kinetic energy
    assert float(full[1, 0]) == 53350000.00000001 or steps != 1000
    dist.barrier()
    if rank == 0:
        print("multi-gpu parity OK: %d ranks x %d worlds x %d steps, %d worlds/rank vs oracle" % (world, n_per, steps, len(sample)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
