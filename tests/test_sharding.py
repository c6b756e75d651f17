"""CPU, world_size 2 over gloo: the frame sharding and the single peak gather (gr_doa_b200/sharding.py).  The per-rank
compute is stubbed with the oracle -- what is under test is the host logic that the NCCL path shares."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_doa_b200 import sharding


def test_shard_ranges_partition_the_frames():
    for n in (0, 1, 7, 65536, 1048576 + 3):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_stream_slabs_carry_the_overlap_halo():
    N, ov, nframes, world = 2048, 512, 11, 3
    hop = N - ov
    for r in range(world):
        lo, hi = sharding.shard_range(nframes, r, world)
        s_lo, s_hi = sharding.stream_slab(nframes, r, world, N, ov)
        assert s_lo == lo * hop and s_hi == (hi - 1) * hop + N
        assert s_hi - s_lo == (hi - lo - 1) * hop + N      # hop*(n-1) + snapshot = forecast + history-1


def test_pack_roundtrip_is_bit_exact():
    val = torch.randn(5, 3); loc = torch.randn(5, 3); bins = torch.randint(0, 4096, (5, 3), dtype=torch.int32)
    v, l, b = sharding.unpack_peaks(sharding.pack_peaks(val, loc, bins), 3)
    assert torch.equal(v, val) and torch.equal(l, loc) and torch.equal(b, bins)


def _worker_packed(rank, world, port, q, mode="gather"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, K = 6, 3
    pb = sharding.PeakBuffers(n, K, "cpu", world=world, is_dst=(rank == 0), mode=mode)
    g = torch.Generator().manual_seed(100 + rank)
    pb.val.copy_(torch.randn((n, K), generator=g)); pb.loc.copy_(torch.randn((n, K), generator=g))
    pb.bins.copy_(torch.randint(0, 4096, (n, K), generator=g, dtype=torch.int32))
    got = pb.gather(dst=0)
    if rank == 0:
        val, loc, bins = sharding.PeakBuffers.split(got)
        ok = True
        for r in range(world):
            g = torch.Generator().manual_seed(100 + r)
            ok &= torch.equal(val[r * n:(r + 1) * n], torch.randn((n, K), generator=g))
            ok &= torch.equal(loc[r * n:(r + 1) * n], torch.randn((n, K), generator=g))
            ok &= torch.equal(bins[r * n:(r + 1) * n], torch.randint(0, 4096, (n, K), generator=g, dtype=torch.int32))
        q.put(bool(ok))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["gather", "allgather"])
def test_packed_peak_buffers_gather(mode):
    """Both forms of the one collective: gather to rank 0, or all_gather_into_tensor with the other ranks dropping the result."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_packed, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _worker_exchange(rank, world, port, q):
    """Three pipelined steps through PeakExchange (two slots): every step's gather delivers that step's data."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, K = 5, 2
    ex = sharding.PeakExchange(n, K, "cpu", world=world, is_dst=(rank == 0))
    ok = True
    for step in range(3):
        val, loc, bins = ex.begin()
        g = torch.Generator().manual_seed(1000 * step + rank)
        val.copy_(torch.randn((n, K), generator=g)); loc.copy_(torch.randn((n, K), generator=g))
        bins.copy_(torch.randint(0, 4096, (n, K), generator=g, dtype=torch.int32))
        ex.submit()
        ex.drain()
        res = ex.result()
        if rank == 0:
            v, l, b = res
            for r in range(world):
                g = torch.Generator().manual_seed(1000 * step + r)
                ok &= torch.equal(v[r * n:(r + 1) * n], torch.randn((n, K), generator=g))
                ok &= torch.equal(l[r * n:(r + 1) * n], torch.randn((n, K), generator=g))
                ok &= torch.equal(b[r * n:(r + 1) * n], torch.randint(0, 4096, (n, K), generator=g, dtype=torch.int32))
        else:
            ok &= res is None
    if rank == 0:
        q.put(bool(ok))
    else:
        assert ok
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_peak_exchange():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_exchange, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _worker(rank, world, port, nframes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from gr_doa_b200 import synth
    M, N, T, P, K = 4, 256, 1, 512, 2
    frames, _ = synth.frames_numpy(nframes, M, N, [70.0], seed=99)       # every rank can regenerate any frame
    lo, hi = sharding.shard_range(nframes, rank, world)

    def chain_fn(x):
        val, loc, bins = O.chain_frames(x, 0, 0.5, T, P, K)
        return torch.from_numpy(val), torch.from_numpy(loc), torch.from_numpy(bins)

    res = sharding.run_sharded(chain_fn, frames[lo:hi], nframes, dst=0)
    if rank == 0:
        val, loc, bins = O.chain_frames(frames, 0, 0.5, T, P, K)
        ok = (np.array_equal(res[0].numpy(), val) and np.array_equal(res[1].numpy(), loc) and np.array_equal(res[2].numpy(), bins))
        q.put(bool(ok))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nframes", [10, 7])
def test_two_rank_gather_matches_single_process(nframes):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nframes, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_counter_based_frames_are_the_same_on_host_and_device_side_generators():
    """synth.frames_philox_*: frame f is a pure function of (seed, f) -- numpy and torch produce the same bits, and a shard
    regenerated from its first frame index equals the corresponding slice of the whole batch."""
    import numpy as np
    from gr_doa_b200 import synth
    a, tha = synth.frames_philox_numpy(70000, 9, 8, 128, [40.0, 90.0, 140.0], jitter_deg=5.0, seed=synth.SEED_BASE + 3)
    b, thb = synth.frames_philox_torch(70000, 9, 8, 128, [40.0, 90.0, 140.0], jitter_deg=5.0, seed=synth.SEED_BASE + 3, device="cpu", chunk=4)
    assert np.array_equal(a.view(np.uint32), b.numpy().view(np.uint32)) and np.array_equal(tha, thb.numpy())
    c, _ = synth.frames_philox_numpy(70004, 3, 8, 128, [40.0, 90.0, 140.0], jitter_deg=5.0, seed=synth.SEED_BASE + 3)
    assert np.array_equal(c, a[4:7])
    other, _ = synth.frames_philox_numpy(70000, 2, 8, 128, [40.0, 90.0, 140.0], jitter_deg=5.0, seed=synth.SEED_BASE + 4)
    assert not np.array_equal(other, a[:2])
    assert abs(float((np.abs(a) ** 2).mean()) - 3.1) < 0.15          # three unit tones + noise at -10 dB
