"""Multi-GPU layout of the update path: replicas only.

A single MPDO sweep is a chain of dependent steps and does not shard, so the unit of distribution is the
circuit (parameter sweeps, noise realisations, batched circuits): circuit i runs on rank i mod world, one
process per GPU, no collective inside evolve. The only exchange step is one all-gather of the fixed-width
per-circuit readout rows at the end (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard(circuit_ids, rank, world):
    """Round-robin: the circuits this rank evolves."""
    return list(circuit_ids)[rank::world]


def gather_readout(local_rows, total, rank=None, world=None):
    """local_rows [n_local, w] (rows of this rank's circuits, in shard order) -> [total, w] ordered by circuit id."""
    if not dist.is_available() or not dist.is_initialized():
        return local_rows
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    per = (total + world - 1) // world
    w = local_rows.shape[1]
    padded = torch.zeros((per, w), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    flat = torch.empty((world * per, w), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(flat, padded)          # one collective: [world * per, w], rank-major
    parts = flat.view(world, per, w)
    out = torch.empty((total, w), dtype=local_rows.dtype, device=local_rows.device)
    for r in range(world):
        count = len(range(r, total, world))
        out[r::world] = parts[r, :count]
    return out
