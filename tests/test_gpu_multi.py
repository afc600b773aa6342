"""GPU, >= 2 devices: the sharded Beamform with the library's own NCCL communicator
(hzsdr_comm_*), one process per GPU, against the oracle."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import go_sdr_oracle as O
import hzsdr as H
import hzsdr_shard as S

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, nchan, n, q):
    ctx = H.Context(rank)
    comm = H.Comm(ctx, world, rank, uid)
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    mine = S.channel_shard(nchan, world, rank)
    raw = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=100 + c, phase=0.37 * c) for c in mine]
    chans = [ctx.to_device(r) for r in raw]
    out = ctx.alloc(n * 8)
    ctx.beamform(H.FORMAT_U8, [c.ptr for c in chans], w[mine.start:mine.stop], n, out.ptr)
    comm.reduce_c64(out.ptr, n, 0)
    res = out.download(np.complex64, n)
    if rank == 0:
        q.put(res)
    ctx.sync()
    comm.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_beamform_nccl_reduce():
    world, nchan, n = 2, 16, 1 << 16
    uid = H.Comm.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, nchan, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    beam = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=100 + c, phase=0.37 * c) for c in range(nchan)]
    assert O.rel_l2(beam, O.beamform(chans, O.FORMAT_U8, w)) <= 1e-5
