"""Clip sharding for configs 2-4: clips are independent units, so rank r of `world`
simply takes a contiguous block -- there is no collective on the data path."""


def clip_range(n_clips, rank, world):
    """contiguous, balanced (sizes differ by at most 1) block of clip indices for this rank"""
    base, extra = divmod(n_clips, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
