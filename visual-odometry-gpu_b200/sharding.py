"""Frame sharding for multi-GPU runs: frames are independent (the reference keeps no state
between detectAndCompute calls, src/orb.cpp:58-109), so rank g simply owns a contiguous block of the batch.
No collective is involved; results land in per-frame slots indexed by global frame id."""


def shard_range(n_frames, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`: ceil(n/world) frames per rank, last ranks may be short/empty."""
    if world_size < 1 or not (0 <= rank < world_size) or n_frames < 0:
        raise ValueError("bad shard arguments")
    per = (n_frames + world_size - 1) // world_size
    lo = min(n_frames, rank * per)
    hi = min(n_frames, lo + per)
    return lo, hi
