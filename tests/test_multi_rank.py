"""world_size-2 tests of the N>1 host logic on CPU (gloo): the partitioners cover every unit exactly
once, and the sharded Beamform -- per-rank partial beams + ONE reduce -- equals the single-process
result.  The per-rank arithmetic here is the oracle (no GPU in this container); the GPU version of
the same flow is tests/test_gpu_multi.py and bench.py --workload c4."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import go_sdr_oracle as O
import hzsdr_shard as S


def test_partitioners_cover_exactly_once():
    for world in (1, 2, 3, 4, 8):
        streams = sorted(s for r in range(world) for s in S.stream_shard(512, world, r))
        assert streams == list(range(512))
        chans = [c for r in range(world) for c in S.channel_shard(64, world, r)]
        assert chans == list(range(64))  # contiguous and in order: keeps the reference's sum order inside a shard
    assert list(S.channel_shard(5, 8, 7)) == []
    with pytest.raises(ValueError):
        S.stream_shard(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _beam_worker(rank, world, port, nchan, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
        mine = S.channel_shard(nchan, world, rank)
        chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=100 + c, phase=0.37 * c) for c in mine]
        partial = O.beamform(chans, O.FORMAT_U8, w[mine.start:mine.stop]) if len(mine) else np.zeros(n, np.complex64)
        t = torch.from_numpy(np.ascontiguousarray(partial).view(np.float32).copy())
        S.reduce_partial_beams(t, dist, root=0)
        # independent streams: every rank reports what it owns; no collective on the data path
        owned = torch.zeros(512, dtype=torch.int32)
        owned[S.stream_shard(512, world, rank)] = 1
        dist.all_reduce(owned)
        if rank == 0:
            q.put((t.numpy().view(np.complex64).copy(), owned.numpy().copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_beamform_with_one_reduce_gloo():
    world, nchan, n = 2, 8, 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_beam_worker, args=(r, world, port, nchan, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    beam, owned = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=100 + c, phase=0.37 * c) for c in range(nchan)]
    want = O.beamform(chans, O.FORMAT_U8, w)
    # the sharded sum re-associates the fp32 additions: within the 1e-5 bar, not bit-equal
    assert O.rel_l2(beam, want) <= 1e-6
    assert np.all(owned == 1)
