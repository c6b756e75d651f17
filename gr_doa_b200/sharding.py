"""Multi-GPU plumbing for the DoA chain: frames are independent (the reference's work() loops never carry state
across frames: lib/autocorrelate_impl.cc:92, lib/MUSIC_lin_array_impl.cc:121, lib/find_local_max_impl.cc:179), so the
path shards by frame index with no data-path collective.  One process per GPU (torchrun); rank r owns the contiguous
block of frames shard_range(B, r, world).  The only exchange is ONE gather of the per-frame peaks (value, location, bin
= 12*K bytes per frame) to rank 0, over torch.distributed (NCCL on GPUs, gloo in the CPU tests).

Streaming inputs (hop < snapshot) shard the same way: rank r additionally needs the `overlap` samples that precede
its first frame's end -- the host hands every rank its slab plus that halo (stream_slab), exactly GNU Radio's
history()-1 (lib/autocorrelate_impl.cc:57); no inter-GPU sample exchange exists.
"""
import torch
import torch.distributed as dist


def shard_range(nframes: int, rank: int, world: int):
    """Contiguous, balanced block of frame indices [lo, hi) for `rank`; the first nframes % world ranks get one more."""
    base, rem = divmod(nframes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stream_slab(nframes: int, rank: int, world: int, snapshot_size: int, overlap_size: int):
    """Sample range [s_lo, s_hi) of each channel stream that rank needs for its frames (halo included)."""
    lo, hi = shard_range(nframes, rank, world)
    hop = snapshot_size - overlap_size
    if hi == lo:
        return lo * hop, lo * hop
    return lo * hop, (hi - 1) * hop + snapshot_size


def pack_peaks(val: torch.Tensor, loc: torch.Tensor, bins: torch.Tensor) -> torch.Tensor:
    """[n][K] f32, [n][K] f32, [n][K] i32 -> one [n][3K] int32 payload (bit-preserving)."""
    return torch.cat([val.contiguous().view(torch.int32), loc.contiguous().view(torch.int32), bins.to(torch.int32)], dim=1)


def unpack_peaks(payload: torch.Tensor, K: int):
    return (payload[:, :K].contiguous().view(torch.float32), payload[:, K:2 * K].contiguous().view(torch.float32),
            payload[:, 2 * K:].contiguous())


def gather_peaks(val, loc, bins, nframes_total: int, dst: int = 0, group=None):
    """One collective: every rank contributes its shard's peaks; rank `dst` returns (val, loc, bins) for all
    nframes_total frames in frame order, the other ranks return None.  Shards may differ in length by one frame, so
    the payload is padded to the longest shard (the pad rows are dropped on dst)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    K = val.shape[1]
    if world == 1:
        return val, loc, bins
    rank = dist.get_rank(group)
    longest = -(-nframes_total // world)
    payload = pack_peaks(val, loc, bins)
    if payload.shape[0] < longest:
        payload = torch.cat([payload, payload.new_zeros((longest - payload.shape[0], 3 * K))], dim=0)
    if rank == dst:
        bufs = [torch.empty_like(payload) for _ in range(world)]
        dist.gather(payload, bufs, dst=dst, group=group)
        parts = []
        for r in range(world):
            lo, hi = shard_range(nframes_total, r, world)
            parts.append(bufs[r][: hi - lo])
        return unpack_peaks(torch.cat(parts, dim=0), K)
    dist.gather(payload, None, dst=dst, group=group)
    return None


class PeakBuffers:
    """Peaks of one shard in ONE packed int32 buffer [3][n][K] (values | locations | bins; float32 words bit-cast), so that
    the per-step exchange is a single collective on a preallocated buffer with no packing kernels around it:
    `val`, `loc`, `bins` are views the chain writes into; on the destination rank `gathered` is [world][3][n][K]."""

    def __init__(self, n: int, K: int, device, world: int = 1, is_dst: bool = False, mode: str = "gather"):
        """mode "gather": dist.gather to dst (NCCL: grouped send/recv, only dst receives); "allgather":
        dist.all_gather_into_tensor (one ring / NVLS collective, every rank receives world * 36 n bytes and all but dst drop them)."""
        self.n, self.K, self.world, self.mode = n, K, world, mode
        self.buf = torch.empty((3, n, K), dtype=torch.int32, device=device)
        self.val = self.buf[0].view(torch.float32)
        self.loc = self.buf[1].view(torch.float32)
        self.bins = self.buf[2]
        need = world > 1 and (is_dst or mode == "allgather")
        self.gathered = torch.empty((world, 3, n, K), dtype=torch.int32, device=device) if need else None
        self._views = [self.gathered[r] for r in range(world)] if (self.gathered is not None and is_dst) else None
        self.is_dst = is_dst

    def outputs(self):
        return self.val, self.loc, self.bins

    def gather(self, dst: int = 0, group=None):
        """Equal-size shards only (bench / streaming slabs); returns the [world][3][n][K] buffer on dst, None elsewhere."""
        if self.world == 1:
            return self.buf[None]
        if self.mode == "allgather":
            dist.all_gather_into_tensor(self.gathered.view(-1), self.buf.view(-1), group=group)
            return self.gathered if self.is_dst else None
        dist.gather(self.buf, self._views, dst=dst, group=group)
        return self.gathered

    @staticmethod
    def split(gathered):
        """[world][3][n][K] int32 -> (val, loc, bins) as [world*n][K] in frame order (rank-major)."""
        w, _, n, K = gathered.shape
        return (gathered[:, 0].reshape(w * n, K).view(torch.float32), gathered[:, 1].reshape(w * n, K).view(torch.float32),
                gathered[:, 2].reshape(w * n, K))


class PeakExchange:
    """The per-step exchange, pipelined: `depth` PeakBuffers used round-robin and a side stream, so that the gather of step i
    runs under the chain kernel of step i+1 (the chain kernel must leave the collective's kernel a few SMs:
    dev knob chain_sms_reserve).  Per step:  out = ex.begin();  chain.run_device(x, out=out);  ex.submit().
    `drain()` makes the current stream wait for every gather submitted so far.  On CPU tensors (gloo tests) the same calls run
    synchronously."""

    def __init__(self, n: int, K: int, device, world: int = 1, is_dst: bool = False, depth: int = 2, dst: int = 0, group=None,
                 pipelined: bool = True, mode: str = "gather"):
        self.bufs = [PeakBuffers(n, K, device, world=world, is_dst=is_dst, mode=mode) for _ in range(depth)]
        self.world, self.dst, self.group, self.i = world, dst, group, 0
        self.cuda = torch.device(device).type == "cuda" and pipelined   # otherwise: the gather runs in stream order
        if self.cuda and world > 1:
            self.comm = torch.cuda.Stream(device=device)
            self.ready = [torch.cuda.Event() for _ in range(depth)]      # chain kernel of this slot finished
            self.done = [None] * depth                                   # gather of this slot finished
        self.last = None

    def begin(self):
        """Outputs for the next step; the current stream first waits until this slot's previous gather has read them."""
        k = self.i % len(self.bufs)
        if self.cuda and self.world > 1 and self.done[k] is not None:
            torch.cuda.current_stream().wait_event(self.done[k])
        return self.bufs[k].outputs()

    def submit(self):
        """Queue the gather of the slot filled since begin(); returns that slot's PeakBuffers."""
        k = self.i % len(self.bufs)
        pb = self.bufs[k]
        self.i += 1
        self.last = pb
        if self.world == 1:
            return pb
        if not self.cuda:
            pb.gather(dst=self.dst, group=self.group)
            return pb
        self.ready[k].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ready[k])
            pb.gather(dst=self.dst, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.comm)
            self.done[k] = ev
        return pb

    def drain(self):
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm)

    def result(self):
        """(val, loc, bins) of the last submitted step in frame order on dst (after drain()), None elsewhere."""
        pb = self.last
        if self.world == 1:
            return pb.outputs()
        return PeakBuffers.split(pb.gathered) if (pb.gathered is not None and pb.is_dst) else None


def run_sharded(chain_fn, frames_local, nframes_total: int, dst: int = 0, group=None):
    """chain_fn(frames_local) -> (val, loc, bins) for this rank's frames; returns the gathered result on dst."""
    val, loc, bins = chain_fn(frames_local)
    return gather_peaks(val, loc, bins, nframes_total, dst=dst, group=group)
