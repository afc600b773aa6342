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


def _rs_worker(rank, world, nchan, n, qs, q_out):
    ctx = H.Context(rank)
    grp = H.BeamGroup(ctx, world, rank, n)
    for r in range(world):  # exchange IPC handles through the parent's queues
        if r != rank:
            qs[r].put((rank, grp.handle))
    handles = {rank: grp.handle}
    while len(handles) < world:
        r, h = qs[rank].get(timeout=120)
        handles[r] = h
    grp.connect([handles[r] for r in range(world)])
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    mine = S.channel_shard(nchan, world, rank)
    sl = n // world
    out = ctx.alloc(sl * 8)
    results = []
    for step in range(3):  # several steps: exercises the double-buffered staging and the flags
        raw = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=1000 * step + c, phase=0.37 * c) for c in mine]
        chans = [ctx.to_device(r) for r in raw]
        grp.exec(H.FORMAT_U8, [c.ptr for c in chans], w[mine.start:mine.stop], out.ptr)
        grp.join()
        results.append(out.download(np.complex64, sl))
    q_out.put((rank, results))
    ctx.sync()
    grp.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_beamform_reduce_scatter_over_peer_memory():
    world, nchan, n = 2, 16, 1 << 16
    ctx = mp.get_context("spawn")
    qs = [ctx.Queue() for _ in range(world)]
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_rs_worker, args=(r, world, nchan, n, qs, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    sl = n // world
    for step in range(3):
        chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=1000 * step + c, phase=0.37 * c) for c in range(nchan)]
        want = O.beamform(chans, O.FORMAT_U8, w)
        beam = np.concatenate([got[r][step] for r in range(world)])
        assert beam.shape == want.shape and sl * world == n
        assert O.rel_l2(beam, want) <= 1e-5


def _rs_batch_worker(rank, world, nchan, n, nbuf, steps, qs, q_out):
    import time
    ctx = H.Context(rank)
    grp = H.BeamGroup(ctx, world, rank, n, max_batch=nbuf)
    for r in range(world):
        if r != rank:
            qs[r].put((rank, grp.handle))
    handles = {rank: grp.handle}
    while len(handles) < world:
        r, h = qs[rank].get(timeout=120)
        handles[r] = h
    grp.connect([handles[r] for r in range(world)])
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    mine = S.channel_shard(nchan, world, rank)
    sl = n // world
    # 3 distinct buffer sets, used round-robin by (step, k); every (step, k) gets its own output slice
    sets = [[ctx.to_device(O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=1000 * v + c, phase=0.37 * c)) for c in mine] for v in range(3)]
    outs = [[ctx.alloc(sl * 8) for _ in range(nbuf)] for _ in range(steps)]
    for step in range(steps):  # back to back, NO host synchronisation; the ranks' launch times are skewed
        if (step + rank) % 3 == 0:
            time.sleep(0.03)
        ptrs = [[c.ptr for c in sets[(step + k) % 3]] for k in range(nbuf)]
        grp.exec_batch(H.FORMAT_U8, ptrs, w[mine.start:mine.stop], [o.ptr for o in outs[step]])
    grp.join()
    ctx.sync()
    q_out.put((rank, [[o.download(np.complex64, sl) for o in row] for row in outs]))
    ctx.sync()
    grp.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_beamform_batched_back_to_back(world):
    """hzsdr_beam_group_exec_batch: 10 exchanges of 4 buffers enqueued back to back with no host
    synchronisation and skewed rank timing -- the double-buffered staging is only reused after every
    owner has acknowledged reading it -- every slice of every step against the oracle."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    nchan, n, nbuf, steps = 16, 1 << 17, 4, 10
    ctx = mp.get_context("spawn")
    qs = [ctx.Queue() for _ in range(world)]
    q_out = ctx.Queue()
    procs = [ctx.Process(target=_rs_batch_worker, args=(r, world, nchan, n, nbuf, steps, qs, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q_out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = O.beamform_angles(433e6, 30.0, [0.15 * c for c in range(nchan)])
    want = []
    for v in range(3):
        chans = [O.synth_raw(O.FORMAT_U8, n, 2_400_000, 1e5, seed=1000 * v + c, phase=0.37 * c) for c in range(nchan)]
        want.append(O.beamform(chans, O.FORMAT_U8, w))
    for step in range(steps):
        for k in range(nbuf):
            beam = np.concatenate([got[r][step][k] for r in range(world)])
            assert O.rel_l2(beam, want[(step + k) % 3]) <= 1e-5, (step, k)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_contexts_in_one_process():
    """One process, one context per GPU (the layout of a multi-GPU ReadBeamform in Go): every chain
    kernel -- each needs more dynamic shared memory than the default, a per-device function
    attribute, and some read a per-device constant table -- runs on the second device as on the
    first, interleaved, with equal results."""
    c0, c1 = H.Context(0), H.Context(1)
    for fmt, fs, n, f0, taps, nfft, D in [(H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 255, 1024, 10),
                                          (H.FORMAT_I8, 20_000_000, 1 << 17, 2.5e6, 255, 1024, 5),
                                          (H.FORMAT_I16, 8_000_000, 1 << 17, 1e6, 2047, 2048, 4),
                                          (H.FORMAT_U8, 2_400_000, 1 << 17, 3e5, 1023, 8192, 16),
                                          (H.FORMAT_I16, 61_440_000, 1 << 17, 7.68e6, 4095, 16384, 16),
                                          (H.FORMAT_I8, 20_000_000, 1 << 16, 2.5e6, 127, 512, 3)]:
        raw = O.synth_raw(fmt, n, fs, f0, seed=9)
        Hf = O.filter_freq(O.lowpass_taps(taps, 1 / (2 * D + 2)), nfft)
        want, ts = O.chain(raw, fmt, fs, -f0, Hf, D)
        got = []
        for ctx in (c0, c1):
            ch = H.Chain(ctx, fmt, fs, -f0, Hf, D)
            total = ch.out_len(n)
            src, dst = ctx.to_device(raw), ctx.alloc(max(total, 1) * 8)
            assert ch.exec(src.ptr, n, dst.ptr, total) == total and ch.ts == ts
            got.append(dst.download(np.complex64, total))
            ch.close()
        assert np.array_equal(got[0].view(np.uint32), got[1].view(np.uint32))
        assert O.rel_l2(got[1], want) <= 1e-5
    # a long transform on the second device (two tile kernels with their own attributes)
    x = (np.random.default_rng(1).standard_normal(1 << 16) + 0j).astype(np.complex64)
    outs = []
    for ctx in (c0, c1):
        plan = H.FftPlan(ctx, x.size, x.size, H.FFT_FORWARD)
        s, d = ctx.to_device(x), ctx.alloc(x.nbytes)
        plan.transform(s.ptr, d.ptr, 1)
        outs.append(d.download(np.complex64, x.size))
        plan.close()
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
