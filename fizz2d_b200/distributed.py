"""Multi-GPU: worlds never interact (each reference World owns its objects, physics.py:270), so a batch is sharded
over ranks with no data-path collective; the only exchange is ONE end-of-run gather of per-world results
[E_total, K, P, H, fitness, status] (48 B per world) -- SURVEY 8e.

One process per GPU (torchrun); `torch.distributed` (NCCL on GPUs, gloo on CPU for tests) is plumbing only.
World w of the global batch lives on rank w % G (interleaved, so contact-heavy and quiescent worlds -- and the
scene types of a mixed batch -- spread evenly), at local index w // G.
"""
from __future__ import annotations

import numpy as np

RESULT_ROWS = 6  # E_total, K, P, H, fitness, status


def shard_indices(n_worlds: int, rank: int, world_size: int) -> np.ndarray:
    """Global indices of the worlds owned by `rank` (interleaved partition)."""
    return np.arange(rank, n_worlds, world_size, dtype=np.int64)


def shard_count(n_worlds: int, rank: int, world_size: int) -> int:
    return (n_worlds - rank + world_size - 1) // world_size


def gather_results(local, n_worlds: int, group=None):
    """All-gather per-world result rows.  `local`: tensor [RESULT_ROWS, W_local] (any device) of the worlds
    shard_indices() assigns to this rank.  Returns [RESULT_ROWS, n_worlds] in GLOBAL world order on every rank.
    A single collective; ranks with one world fewer are padded."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[1] == n_worlds
        return local
    ws = dist.get_world_size(group)
    wmax = (n_worlds + ws - 1) // ws
    rows = local.shape[0]
    padded = torch.zeros((rows, wmax), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    out = torch.empty((ws, rows, wmax), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), padded.view(-1), group=group)
    # out[r, :, k] is world k*ws + r  ->  [rows, wmax, ws] flattened is global order
    return out.permute(1, 2, 0).reshape(rows, wmax * ws)[:, :n_worlds].contiguous()


def pack_results(energy, fitness, status):
    """[6, W] float64: World.energy() rows, a fitness row and the status word (as float64)."""
    import torch
    out = torch.empty((RESULT_ROWS, energy.shape[1]), dtype=torch.float64, device=energy.device)
    out[:4] = energy
    out[4] = fitness
    out[5] = status.to(torch.float64)
    return out
