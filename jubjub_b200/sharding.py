"""Batch partitioning for the multi-GPU path: rank g of G owns the contiguous block
[g*n/G, (g+1)*n/G) of unit indices.  Pure host logic, no device code."""


def shard_range(n, rank, world):
    """Half-open index range [start, stop) owned by `rank`; blocks differ by at most one unit."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n * rank // world, n * (rank + 1) // world


def gather_offsets(n, world):
    """Start offset of every rank's block in the gathered output (rank order == index order)."""
    return [shard_range(n, r, world)[0] for r in range(world)]


def equal_shards(n, world):
    """The NCCL all-gather path needs equal blocks; returns units per rank or raises."""
    if n % world:
        raise ValueError(f"{n} units do not split evenly over {world} ranks; pad the batch or use shard_range")
    return n // world
