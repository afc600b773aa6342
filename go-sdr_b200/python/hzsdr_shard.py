"""Multi-GPU partitioning of the hot path: one process per GPU.

* Channelizer / independent streams (BASELINE configs 2, 3, 5): stream s -> rank s mod G.  No
  data-path collective; every rank owns its streams' NCO state.
* Beamform (config 4): channel c -> rank c div ceil(C/G) (contiguous blocks, so each rank's partial
  sum keeps the reference's left-to-right channel order inside the shard); every rank computes
  the partial beam of its channels with hzsdr_beamform and the partial beams are summed onto a
  root with ONE collective (NCCL reduce over NVLink on GPUs; any torch.distributed backend in the
  CPU tests).

Pure host logic: no CUDA, no oracle.  Used by bench.py and by tests/test_multi_rank.py (gloo).
"""
from __future__ import annotations


def stream_shard(n_streams: int, world: int, rank: int) -> list[int]:
    """Streams owned by `rank`: s mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world))


def channel_shard(n_channels: int, world: int, rank: int) -> range:
    """Contiguous block of channels owned by `rank` (the last ranks may own fewer / none)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per = -(-n_channels // world)
    lo = min(rank * per, n_channels)
    return range(lo, min(lo + per, n_channels))


def reduce_partial_beams(partial, dist, root: int = 0):
    """Sum per-rank partial beams (a torch tensor of float32 pairs) onto `root` with one collective.
    On GPUs bench.py uses the library's own NCCL communicator (hzsdr_comm_reduce_c64) instead."""
    dist.reduce(partial, dst=root, op=dist.ReduceOp.SUM)
    return partial
