#!/usr/bin/env python
"""bench.py -- the DoA hot path on B200: autocorrelate -> MUSIC_lin_array -> find_local_max, peaks only.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Headline workload (BASELINE.json configs[2], the configuration the metric is quoted on): 65,536 independent 8-element ULA
frames x 2048 snapshots (8 GiB of complex64, resident in HBM before the timed region), 3 sources at 10 dB SNR, 4096-point
angle scan, K = 3 peaks.  A step = one pass of the whole chain over that batch.  For N > 1 every rank owns a batch of the same
size (weak scaling; frames are independent, no data-path collective) and ONE gather of the per-frame peaks to rank 0 closes
each step.  Inputs (8 GiB) are far larger than L2 (126 MB), so no flush is needed between iterations.

JSON keys beyond the base contract:
  roofline      the chain kernel against measured HBM copy bandwidth (algorithmic bytes / CUDA-event launch time)
  per_step      each of >= 20 further steps between its own pair of events: median / best / worst
  sustained     the same step back to back for >= 2 s (the headline's K steps last tens of milliseconds): ms/step, roofline
                fraction, the time course in slices, SM / memory clocks, power and temperature sampled through NVML
  cpu_baseline  the reference's CPU path on this box's host cores, bounded sample of the same frames: the reference's own
                block sources (oracle/_ref, kind "reference") when that build is present, and the port (oracle/doa_oracle.cpp)
  parity        exemption rates measured on that sample: frames whose peak bins differ from the CPU arm, how many of them
                are near-ties, and how many are unexplained (must be 0); high_snr_informational: frames with different bins at
                the reference app's noise amplitude 5e-3 (46 dB) -- GPU vs port, GPU vs reference build, and the two CPU builds
                against each other (float32 Q at a null is below its own rounding there: they disagree alike; not a gate)
  e2e           the same chain through the host-pointer C-ABI call, H2D/D2H inside the timed region, per-rank H2D GB/s
  other_configs the other BASELINE.json configs (N = 1: cfg1 streaming with overlap 512 and both averaging methods, cfg2
                Root-MUSIC chain, cfg4 large array; every N: cfg5 = 1,048,576 16-element frames x 1024 snapshots STRONG-scaled
                over the ranks, one peak gather per step), each with ms, frames/s, roofline fraction and (N = 1) its CPU arm
  clocks, gpu_launches

--impl reference times the reference arm on the host cores: oracle/_ref (the reference's unmodified lib/*_impl.cc compiled
against the Armadillo / GNU Radio stand-ins, oracle/build_ref.py) when present, else the port; bounded sample per step.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(frames=65536, M=8, N=2048, T=3, P=4096, K=3, d=0.5, thetas=[40.0, 90.0, 140.0], jitter=5.0, snr_db=10.0)
METRIC, UNIT = "doa_frames_per_s", "frames/s"
ALG_BYTES_CHAIN = lambda w: 8 * w["M"] * (w["hop"] if "hop" in w else w["N"]) + 8 * w["K"]     # unique samples in once + K (value, location) out
ALG_BYTES_COV = lambda w: 8 * w["M"] * w["N"] + 8 * w["M"] * w["M"]            # autocorrelate stage: samples in + M*M complex out

# the other BASELINE.json configs (SURVEY 8(d) shapes); cfg5's T, P, K are not given by BASELINE.json: T3, P4096, K3 assumed
CFG1 = dict(name="cfg1", M=4, N=2048, overlap=512, T=1, P=2048, K=1, thetas=[60.0], frames=262144)
CFG2 = dict(name="cfg2", M=4, N=2048, overlap=512, T=2, thetas=[50.0, 110.0], frames=262144, avg=1)
CFG4 = dict(name="cfg4", M=64, N=16384, T=8, P=16384, K=8, thetas=[30.0 + 120.0 * i / 7 for i in range(8)], frames=512)
CFG5 = dict(name="cfg5", M=16, N=1024, T=3, P=4096, K=3, thetas=[40.0, 90.0, 140.0], frames=1048576)


def workload_text(w):
    return (f"cfg3: {w['frames']} independent {w['M']}-element ULA frames x {w['N']} snapshots, {w['T']} sources @ {w['snr_db']} dB, "
            f"{w['P']}-point scan, K={w['K']} peaks (autocorrelate -> MUSIC_lin_array -> find_local_max)")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM / memory clocks, power, temperature and throttle reasons of one GPU through NVML while a timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index, period=0.004, detailed=False):
        super().__init__(daemon=True)
        self.index, self.period, self.detailed = index, period, detailed
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.mem, self.power, self.temp, self.stamps = [], [], [], []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.stamps.append(time.perf_counter())
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                if self.detailed:
                    self.mem.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                    self.temp.append(nv.nvmlDeviceGetTemperature(self.h, nv.NVML_TEMPERATURE_GPU))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        out = {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
               "samples": len(self.samples)}
        if self.detailed and self.power:
            n = len(self.samples)
            q = max(1, n // 4)
            out.update({"sm_mhz_min": min(self.samples), "mem_mhz": statistics.median(self.mem), "mem_mhz_min": min(self.mem),
                        "power_w_first_quarter": statistics.mean(self.power[:q]), "power_w_last_quarter": statistics.mean(self.power[-q:]),
                        "power_w_max": max(self.power), "temp_c_first": self.temp[0], "temp_c_last": self.temp[-1],
                        "sample_period_ms": 1e3 * (self.stamps[-1] - self.stamps[0]) / max(1, n - 1)})
        return out


# ---------------------------------------------------------------------------------------------------------------- CPU arm
def cpu_arm():
    """(module, kind): the reference's own sources when oracle/_ref is built (here or prebuilt), else the port."""
    try:
        from oracle import reference as REF
        if REF.available():
            REF.lib()
            return REF, "reference"
    except Exception:
        pass
    from oracle import oracle as O
    return O, "port"


def cpu_chain(mod, kind, frames, avg, d, T, P, K, cores, root=False):
    """One timed pass of the CPU arm over `frames`: (values, locations, bins or None, seconds).  The reference build runs as
    `cores` single-threaded worker processes (its thousands of tiny BLAS calls per frame serialise the threads of one process on
    OpenBLAS's buffer lock: oracle/ref_worker.py), the port as OpenMP threads over frames."""
    if kind == "reference":
        outs, sec = mod.chain_frames_procs(frames, avg, d, T, P, K, cores, root=root)
        return (outs[0], None, None, sec) if root else (outs[0], outs[1], None, sec)
    t0 = time.perf_counter()
    if root:
        aoa = mod.rootmusic(mod.autocorrelate_frames(frames, avg, nthreads=cores), d, T, frames.shape[1], nthreads=cores)
        return aoa, None, None, time.perf_counter() - t0
    val, loc, bins = mod.chain_frames(frames, avg, d, T, P, K, nthreads=cores)
    return val, loc, bins, time.perf_counter() - t0


def run_reference(args, w):
    """Reference arm: the reference's CPU implementation on all host cores; each step = a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from gr_doa_b200 import synth
    mod, kind = cpu_arm()
    cores = mod.max_threads()
    per_step = (64 if kind == "reference" else 128) * cores     # ~2 ms (reference build) / ~0.5 ms (port) per frame and core
    fr, _ = synth.frames_philox_numpy(0, per_step, w["M"], w["N"], w["thetas"], d=w["d"], snr_db=w["snr_db"], jitter_deg=w["jitter"],
                                      seed=synth.SEED_BASE + 3)      # the first frames of the very batch the GPU arm times
    for _ in range(min(args.warmup, 1)):
        cpu_chain(mod, kind, fr, 0, w["d"], w["T"], w["P"], w["K"], cores)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_chain(mod, kind, fr, 0, w["d"], w["T"], w["P"], w["K"], cores)[3]
    value = per_step * args.steps / dt
    what = ("gr-doa's unmodified lib/{autocorrelate,MUSIC_lin_array,find_local_max}_impl.cc (oracle/_ref: Armadillo / GNU Radio stand-ins, OpenBLAS), "
            "one single-threaded worker process per core (oracle/ref_worker.py), timed pass of the slowest worker" if kind == "reference" else "oracle port (oracle/doa_oracle.cpp), OpenMP over frames")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(w), "sample_frames_per_step": per_step},
        "msamples_per_s_per_stream": value * w["N"] / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{per_step} frames of the workload per step x {args.steps} steps; {what}; BLAS single-threaded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------- helpers
def classify_against_cpu(np, O, sub, bins_gpu, bins_cpu, w, cores):
    """Exemption rates on a CPU-checked sample: frames whose sorted peak bins differ, how many are near-ties (one bin apart and
    closer, in the float64 null spectrum, than 8x the float32 CPU path's own distance from it: tests/parity.py), how many not."""
    bg, bc = np.sort(bins_gpu, axis=1), np.sort(bins_cpu, axis=1)
    diff = np.where((bg != bc).any(axis=1))[0]
    near, unexplained = 0, 0
    if len(diff):
        R = O.autocorrelate_frames(sub[diff], 0, nthreads=cores)
        q64 = O.music_f64(R, w["d"], w["T"], w["M"], w["P"], nthreads=cores)
        q32 = O.music_q(R, w["d"], w["T"], w["M"], w["P"], nthreads=cores)
        for i, f in enumerate(diff):
            noise = float(np.abs(q32[i].astype(np.float64) - q64[i]).max())
            ok = all(a == b or (abs(int(a) - int(b)) <= 1 and abs(q64[i, a] - q64[i, b]) <= 8.0 * noise) for a, b in zip(bg[f], bc[f]))
            near += ok
            unexplained += (not ok)
    n = len(bins_gpu)
    return {"frames_checked": n, "frames_with_different_bins": int(len(diff)), "near_tie_frac": near / n, "unexplained_bins": int(unexplained)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=WORKLOAD["frames"], help="frames per GPU (default: the BASELINE config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sc16", action="store_true", help="skip the informational sc16-input arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (cfg1, cfg2, cfg4, cfg5)")
    ap.add_argument("--cfg5-frames", type=int, default=CFG5["frames"], help="total frames of the strong-scaled cfg5 run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = dict(WORKLOAD, frames=args.frames)
    if args.impl == "reference":
        return run_reference(args, w)

    import numpy as np
    import torch
    import torch.distributed as dist
    import gr_doa_b200 as doa
    from gr_doa_b200 import _lib, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = measured_peaks()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    B, M, N, T, P, K = w["frames"], w["M"], w["N"], w["T"], w["P"], w["K"]
    # counter-based batch: frame f is a pure function of (seed, f); rank r holds frames [r B, (r + 1) B), and any of
    # them can be regenerated on the host (the CPU arm below does, for a slice, and compares)
    x, _ = synth.frames_philox_torch(rank * B, B, M, N, w["thetas"], d=w["d"], snr_db=w["snr_db"], jitter_deg=w["jitter"],
                                     seed=synth.SEED_BASE + 3, device=dev)
    chain = doa.DoaChain(M, N, 0, 0, w["d"], T, P, K, device=local, max_frames=B)
    # the one collective of the path: packed peaks of every shard to rank 0, double-buffered on a side stream so that the
    # gather of step i runs under the chain kernel of step i+1 (every gather completes inside the timed region: drain())
    # measured (B200 x8 boxes): the serial gather costs 0.03 ms per step at 2 GPUs, 0.24 ms at 8; pipelined with SMs left to NCCL:
    # round 1: 8 GPUs 1.87-1.95 -> 1.74 ms per step (2 SMs), 4 GPUs no change, 2 GPUs 2-3 % slower (the reserved SMs);
    # round 2 (profiles/r02_bench_n8*.json, another box): serial gather 2.30 ms, pipelined 2.00 (2 SMs) / 1.83 (1 SM);
    # all_gather_into_tensor instead of gather: 2.07 serial, 2.39 / 2.05 pipelined (2 / 4 SMs) -- every rank then receives
    # 8 x 2.4 MB it drops, and its ring kernel needs more SMs: not better, gather stays
    pipelined = os.environ.get("DOA_PIPELINE", "1" if world > 4 else "0") != "0"
    gather_mode = os.environ.get("DOA_GATHER", "gather")        # "gather" (send/recv to rank 0) or "allgather" (all_gather_into_tensor)
    peaks = sharding.PeakExchange(B, K, dev, world=world, is_dst=(rank == 0), pipelined=pipelined, mode=gather_mode)
    reserve = int(os.environ.get("DOA_SMS_RESERVE", "1")) if (world > 1 and pipelined) else 0
    chain.set_sms_reserve(reserve)
    out = peaks.bufs[0].outputs()
    total = B * world

    def step():
        nonlocal out
        out = peaks.begin()
        chain.run_device(x, out=out)
        peaks.submit()

    for _ in range(args.warmup):
        step()
    peaks.drain()
    chain.set_profiling(True)
    fence()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = 0
    for _ in range(args.steps):
        step()
        launches += chain.launches()
    peaks.drain()
    e1.record()
    fence()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    cov_ms, eig_ms, scan_ms = chain.stage_ms()
    chain.set_profiling(False)
    nlaunch = chain.launches()
    fused = nlaunch == 1          # one persistent kernel for the whole chain: the stage timers read (0, 0, total)
    ms, cov_ms, eig_ms, scan_ms = max_over_ranks([ms, cov_ms, eig_ms, scan_ms])
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3)

    # ---- per-step spread (median and best): each step between its own pair of events, after the contract's region ----
    per_step = None
    if not args.no_sustained:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(args.steps, 20))]
        for a, b in evs:
            a.record()
            step()
            b.record()
        peaks.drain()
        fence()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        per_step = {"steps": len(ts), "median_ms": ts[len(ts) // 2], "best_ms": ts[0], "worst_ms": ts[-1],
                    "note": "rank-local, one event pair per step (kernel + launch gap; the exchange of a multi-rank run is not inside the pair)"}

    # ---- the sustained regime: the same step back to back for >= 2 s, time course in slices, NVML samples --------------------
    sustained = None
    if not args.no_sustained:
        n_slices, slice_steps = 10, max(10, int(0.25 / (ms_per_step * 1e-3)))     # ~0.25 s per slice
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_slices + 1)]
        fence()
        smp = ClockSampler(local, period=0.001, detailed=True)
        smp.start()
        evs[0].record()
        for s in range(n_slices):
            for _ in range(slice_steps):
                step()
            peaks.drain()
            evs[s + 1].record()
        fence()
        sclk = smp.stop()
        course = [evs[s].elapsed_time(evs[s + 1]) / slice_steps for s in range(n_slices)]
        course = [max_over_ranks([c])[0] for c in course]
        sus_ms = sum(course) / n_slices
        sustained = {"seconds": sum(course) * slice_steps * 1e-3, "steps": n_slices * slice_steps, "ms_per_step": sus_ms,
                     "value": total / (sus_ms * 1e-3), "unit": UNIT,
                     "roofline_frac_chain": ALG_BYTES_CHAIN(w) * B / (sus_ms * 1e-3) / 1e9 / peak,
                     "ms_per_step_by_slice": [round(c, 4) for c in course], "vs_burst": sus_ms / ms_per_step, "nvml": sclk}

    # ---- end to end through the host-pointer C-ABI call (pinned host buffers, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        import psutil
        eb = B
        while eb > 4096 and eb * M * N * 8 * world * 2 > psutil.virtual_memory().available:
            eb //= 2
        hx = torch.empty((eb, M, N), dtype=torch.complex64, pin_memory=True)
        hx.copy_(x[:eb])
        hout = (torch.empty((eb, K), dtype=torch.float32, pin_memory=True), torch.empty((eb, K), dtype=torch.float32, pin_memory=True),
                torch.empty((eb, K), dtype=torch.int32, pin_memory=True))
        esteps = max(3, min(args.steps, 10))
        for _ in range(2):
            chain.run_host(hx, out=hout)
        # bare copy rate of this rank while every rank copies (what the platform gives; the chain call cannot beat it)
        dbuf = torch.empty((min(eb, 8192), M, N), dtype=torch.complex64, device=dev)
        fence()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(0, eb - dbuf.shape[0] + 1, dbuf.shape[0]):
            dbuf.copy_(hx[i:i + dbuf.shape[0]], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        bare_gbps = (eb // dbuf.shape[0]) * dbuf.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del dbuf
        fence()
        t0 = time.perf_counter()
        for _ in range(esteps):
            chain.run_host(hx, out=hout)      # synchronous: returns when the peaks are in host memory
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        dt = max_over_ranks([dt_local])[0]
        h2d = eb * M * N * 8
        rates = [h2d * esteps / dt_local / 1e9, bare_gbps]
        if world > 1:
            gl = [torch.zeros(2, device=dev, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(gl, torch.tensor(rates, device=dev, dtype=torch.float64))
            per_rank = [[round(float(v), 2) for v in g.tolist()] for g in gl]
        else:
            per_rank = [[round(v, 2) for v in rates]]
        e2e = {"value": eb * world * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": eb * K * 12,
               "frames_per_step_per_gpu": eb, "steps": esteps, "ms_per_step": dt / esteps * 1e3,
               "h2d_GBps_per_rank": [p[0] for p in per_rank], "bare_pinned_memcpy_GBps_per_rank": [p[1] for p in per_rank],
               "h2d_GBps_aggregate": sum(p[0] for p in per_rank),
               "api": "doa_cuda_chain_run (host pointers, chunked H2D overlapped with kernels, synchronous)",
               "timer": "host wall clock around the synchronous calls, max over ranks",
               "note": "PCIe-bound: the chain call moves its input at the rate a bare cudaMemcpyAsync loop from the same pinned buffer reaches on this rank while all ranks copy"}
        assert torch.equal(hout[2].cuda(), out[2][:eb]), "host path and device path disagree"
        del hx

    # ---- informational: the same frames as sc16 (UHD int16 I/Q) -- NOT the BASELINE wire format ----
    sc16 = None
    if not args.no_e2e and not args.no_sc16:
        try:
            s15 = 1.0 / 32768
            sb = min(B, 32768)
            q = torch.view_as_real(x[:sb]).mul(8192.0).round_().clamp_(-32768, 32767).to(torch.int16)
            fcq = torch.view_as_complex(q.to(torch.float32).mul_(s15))
            ref_q = [t.clone() for t in chain.run_device(fcq)]
            del fcq
            chain.set_input_format("sc16", s15)
            for _ in range(3):
                got_q = chain.run_device(q)
            torch.cuda.synchronize()      # rank-local: no collective inside this optional arm
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(10):
                chain.run_device(q)
            q1.record()
            torch.cuda.synchronize()      # rank-local: no collective inside this optional arm
            same_q = all(torch.equal(a, b) for a, b in zip(got_q, ref_q))
            hq = torch.empty(q.shape, dtype=torch.int16, pin_memory=True)
            hq.copy_(q)
            hout_q = (torch.empty((sb, K), dtype=torch.float32, pin_memory=True), torch.empty((sb, K), dtype=torch.float32, pin_memory=True),
                      torch.empty((sb, K), dtype=torch.int32, pin_memory=True))
            chain.run_host(hq, out=hout_q)
            t0 = time.perf_counter()
            for _ in range(5):
                chain.run_host(hq, out=hout_q)
            torch.cuda.synchronize()
            dtq = (time.perf_counter() - t0) / 5
            sc16 = {"note": "informational, per GPU: int16 I/Q samples read directly by the covariance (half the bytes); not the BASELINE wire format",
                    "frames": sb, "device_ms": q0.elapsed_time(q1) / 10, "device_frames_per_s": sb / (q0.elapsed_time(q1) / 10 * 1e-3),
                    "e2e_frames_per_s": sb / dtq, "h2d_bytes_per_step": sb * M * N * 4,
                    "bit_identical_to_fc32_on_converted_samples": bool(same_q)}
            del q, hq
        except Exception as ex:      # never lose the headline line to the informational arm
            sc16 = {"error": repr(ex)}
        finally:
            chain.set_input_format("fc32")

    # ---- CPU baseline + exemption rates: the reference's CPU path on this box's host cores, bounded sample of the same frames ----
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        cores = O.max_threads()
        ns = min(B, max(2048, 1024 * cores))        # ~0.35 ms per frame and core: about 10 s of CPU work in under a second of wall clock
        sub = x[:ns].cpu().numpy()
        O.chain_frames(sub[:256], 0, w["d"], T, P, K, nthreads=cores)
        t0 = time.perf_counter()
        v_o, l_o, b_o = O.chain_frames(sub, 0, w["d"], T, P, K, nthreads=cores)
        dt_port = time.perf_counter() - t0
        bins_gpu = out[2][:ns].cpu().numpy()
        parity = {"against_port": classify_against_cpu(np, O, sub, bins_gpu, b_o, w, cores)}
        r0 = min(B - 64, 12345)             # any slice of the device batch can be re-derived on the host
        host, _ = synth.frames_philox_numpy(rank * B + r0, 64, M, N, w["thetas"], d=w["d"], snr_db=w["snr_db"], jitter_deg=w["jitter"],
                                            seed=synth.SEED_BASE + 3)
        devs = x[r0:r0 + 64].cpu().numpy()
        parity["host_regeneration"] = {"frames": [r0, r0 + 64], "samples_bit_identical_frac": float((host.view(np.uint32) == devs.view(np.uint32)).mean()),
                                       "max_abs_diff": float(np.abs(host - devs).max())}
        cpu = {"value": ns / dt_port, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ns} frames of the timed batch, oracle port, OpenMP over frames ({cores} threads), BLAS single-threaded",
               "peak_bins_identical_frac": 1.0 - parity["against_port"]["frames_with_different_bins"] / ns}
        # one high-SNR run (noise amplitude 5e-3, the reference app's value: 46 dB) -- informational.  There Q at a
        # null is smaller than the float32 rounding of its own terms (it comes out <= 0 on some frames, 1/Q then flips sign and the
        # dB spectrum's maximum moves): the reference's OWN two CPU builds pick different peaks on ~5 % of the frames, many bins
        # apart, and the GPU differs from either by the same fraction.  The near-tie rule does not apply to that regime.
        def high_snr():
            nh = min(ns, 2048)
            hs, _ = synth.frames_philox_numpy(0, nh, M, N, w["thetas"], d=w["d"], snr_db=46.0206, jitter_deg=w["jitter"], seed=synth.SEED_BASE + 33)
            got_h = chain.run_device(torch.from_numpy(hs).to(dev))
            torch.cuda.synchronize()
            _, _, b_h = O.chain_frames(hs, 0, w["d"], T, P, K, nthreads=cores)

            def differ(a, b):
                a, b = np.sort(a, axis=1), np.sort(b, axis=1)
                dist = np.abs(a.astype(np.int64) - b.astype(np.int64)).max(axis=1)
                return {"frames_checked": int(len(a)), "frames_with_different_bins": int((dist > 0).sum()), "of_which_one_bin_apart": int((dist == 1).sum())}
            res = {"snr_db": 46.02, "gpu_vs_port": differ(got_h[2].cpu().numpy(), b_h),
                   "note": "noise amplitude 5e-3 (apps/run_MUSIC_lin_array_simulation.py:209); not a gate: the null depth (median 1e-6 of the float64 Q) is below "
                           "float32's own error on Q (6e-6; Q <= 0 on ~5 % of the frames), so the reference's two CPU builds disagree with each other as often as the GPU does with either"}
            mod_h, kind_h = cpu_arm()
            if kind_h == "reference":
                nq = min(nh, 32 * cores)
                _, l_q, _, _ = cpu_chain(mod_h, kind_h, hs[:nq], 0, w["d"], T, P, K, cores)
                b_q = np.rint(l_q.astype(np.float64) * P / 180.0).astype(np.int64)
                res["gpu_vs_reference_build"] = differ(got_h[2].cpu().numpy()[:nq], b_q)
                res["port_vs_reference_build"] = differ(b_h[:nq], b_q)
            return res
        try:
            parity["high_snr_informational"] = high_snr()
        except Exception as ex:
            parity["high_snr_informational"] = {"error": repr(ex)}
        mod, kind = cpu_arm()
        if kind == "reference":
            nr = min(ns, 256 * cores)
            v_r, l_r, _, dt_ref = cpu_chain(mod, kind, sub[:nr], 0, w["d"], T, P, K, cores)
            bins_ref = np.rint(l_r.astype(np.float64) * P / 180.0).astype(np.int64)
            parity["against_reference_build"] = classify_against_cpu(np, O, sub[:nr], bins_gpu[:nr], bins_ref, w, cores)
            parity["port_vs_reference_build"] = classify_against_cpu(np, O, sub[:nr], b_o[:nr], bins_ref, w, cores)
            cpu = {"value": nr / dt_ref, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {nr} frames of the timed batch through gr-doa's unmodified block sources (oracle/_ref), one single-threaded worker process per core ({cores}), BLAS single-threaded; timed pass of the slowest worker",
                   "peak_bins_identical_frac": 1.0 - parity["against_reference_build"]["frames_with_different_bins"] / nr,
                   "port": {"value": ns / dt_port, "sample_frames": ns, "threads": cores}}

    del x, peaks
    chain.close()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs -------------------------------------------------------------------------------------
    others = None
    if not args.no_others:
        others = {}
        try:
            if world == 1:
                others.update(bench_small_configs(doa, synth, torch, np, dev, local, peak, args))
            others["cfg5"] = bench_cfg5(doa, synth, sharding, torch, dist, np, dev, local, rank, world, peak, args, fence, max_over_ranks)
        except Exception as ex:     # never lose the headline line to the additional configs
            import traceback
            others["error"] = repr(ex) + " | " + traceback.format_exc()[-600:]

    if rank == 0:
        kern_ms = (cov_ms + eig_ms + scan_ms) if fused else cov_ms
        kern_bytes = (ALG_BYTES_CHAIN(w) if fused else ALG_BYTES_COV(w)) * B
        cov_gbs = kern_bytes / (kern_ms * 1e-3) / 1e9
        chain_gbs = ALG_BYTES_CHAIN(w) * B / ((ms_per_step) * 1e-3) / 1e9      # per GPU
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "cov_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj["fused_dram_bytes_per_frame" if fused else "dram_bytes_per_frame"] * B
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_text(w), "frames_per_gpu": B, "peak_exchange": gather_mode if world > 1 else None, "parallelism": f"frames sharded x{world}, one peak gather per step" + (" (double-buffered on a side stream, under the next step's kernel; " + str(reserve) + " SM left to NCCL)" if (world > 1 and pipelined) else ""),
                       "l2": "inputs (8 GiB/GPU) larger than L2 (126 MB): no flush between iterations",
                       "timer": "CUDA events on the launching stream, max over ranks",
                       "regime": f"burst: {args.steps} steps = {ms:.0f} ms; see `sustained` for >= 2 s of back-to-back steps"},
            "msamples_per_s_per_stream": value * N / 1e6,
            "msamples_per_s_aggregate": value * N * M / 1e6,
            "roofline": {"bound": "hbm", "achieved": cov_gbs, "peak": peak, "unit": "GB/s", "frac": cov_gbs / peak,
                         "traffic": traffic,
                         "kernel": ("chain_ws_kernel<8> (covariance + eigensolver + scan/peaks in one persistent warp-specialised kernel)" if fused
                                    else "cov_small_kernel<8> (covariance, dominant)"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes, "launch_ms": kern_ms,
                         "stage_ms": ({"fused_chain": kern_ms} if fused else {"cov": cov_ms, "eig": eig_ms, "scan_peaks": scan_ms}),
                         "chain": {"algorithmic_bytes_per_frame": ALG_BYTES_CHAIN(w), "achieved": chain_gbs,
                                   "frac": chain_gbs / peak, "note": "whole step (chain kernel" + ("" if fused else "s") + (" + peak gather" if world > 1 else "") + ") per GPU against the same HBM peak"}},
            "per_step": per_step,
            "sustained": sustained,
            "cpu_baseline": cpu,
            "parity": parity,
            "e2e": e2e,
            "sc16_input": sc16,
            "other_configs": others,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def time_calls(torch, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_small_configs(doa, synth, torch, np, dev, local, peak, args):
    """cfg1 (the reference's own app shape: streaming, overlap 512, both averaging methods), cfg2 (Root-MUSIC chain), cfg4 (large
    array) on one GPU, inputs resident; each next to the CPU arm on a bounded sample of the same shape."""
    res = {}
    mod, kind = (None, None) if args.no_cpu else cpu_arm()
    cores = mod.max_threads() if mod else 0

    def cpu_rate(c, avg, n_per_core, root=False):
        if mod is None:
            return None
        n = max(cores, n_per_core * cores)
        fr, _ = synth.frames_numpy(n, c["M"], c["N"], c["thetas"], snr_db=10.0, jitter_deg=2.0, seed=synth.SEED_BASE + 50)
        dt = cpu_chain(mod, kind, fr, avg, 0.5, c["T"], c.get("P", 0), c.get("K", 0), cores, root=root)[3]
        return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{n} independent frames of this shape"}

    # cfg1 / cfg2: M channel streams, frames overlap by 512 samples (hop 1536): the covariance reads them in place
    for c in (CFG1, CFG2):
        hop = c["N"] - c["overlap"]
        Bc = c["frames"]
        L = (Bc - 1) * hop + c["N"]
        xs = synth.stream_torch(c["M"], L, c["thetas"], snr_db=10.0, seed=synth.SEED_BASE + 11, device=dev)     # [M][L]
        wb = dict(M=c["M"], hop=hop, K=c.get("K", 0))
        if c is CFG1:
            for avg in (0, 1):
                ch = doa.DoaChain(c["M"], c["N"], c["overlap"], avg, 0.5, c["T"], c["P"], c["K"], device=local, max_frames=Bc)
                ms = time_calls(torch, lambda: ch.run_device(xs, frame_stride=hop, chan_stride=L, nframes=Bc), 10)
                gbs = ALG_BYTES_CHAIN(wb) * Bc / (ms * 1e-3) / 1e9
                res["cfg1_" + ("fb" if avg else "fwd")] = {
                    "workload": f"{Bc} frames of 4 channel streams, snapshot 2048, overlap 512 (hop 1536), avg_method {avg}, 1 source, P 2048, K 1 (index_max)",
                    "ms": ms, "frames_per_s": Bc / (ms * 1e-3), "msamples_per_s_per_stream": Bc * hop / (ms * 1e-3) / 1e6,
                    "launches": ch.launches(), "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": ALG_BYTES_CHAIN(wb), "achieved": gbs, "frac": gbs / peak},
                    "cpu_baseline": cpu_rate(c, avg, 64)}
                ch.close()
        else:
            rc = doa.RootMusicChain(c["M"], c["N"], c["overlap"], 1, 0.5, c["T"], device=local, max_frames=Bc)
            ms = time_calls(torch, lambda: rc.run_device(xs, frame_stride=hop, chan_stride=L, nframes=Bc), 10)
            bytes_pf = 8 * c["M"] * hop + 4 * c["T"]
            gbs = bytes_pf * Bc / (ms * 1e-3) / 1e9
            aoa = rc.run_device(xs, frame_stride=hop, chan_stride=L, nframes=Bc)
            res["cfg2_rootmusic"] = {
                "workload": f"{Bc} frames of 4 channel streams, snapshot 2048, overlap 512, forward-backward averaging, 2 sources (autocorrelate -> rootMUSIC_linear_array)",
                "ms": ms, "frames_per_s": Bc / (ms * 1e-3), "launches": rc.launches(), "nan_frac": float(torch.isnan(aoa).any(1).float().mean()),
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": bytes_pf, "achieved": gbs, "frac": gbs / peak},
                "cpu_baseline": cpu_rate(c, 1, 128, root=True)}
            if mod is not None:   # Root-MUSIC exemption rate: frames whose selected root is within 4e-4 of the unit circle (tests/parity.py)
                from oracle import oracle as O
                nchk = 4096
                Rr = O.autocorrelate(xs[:, :(nchk - 1) * hop + c["N"]].cpu().numpy(), c["N"], c["overlap"], 1)
                a64, d64 = O.rootmusic_f64(Rr, 0.5, c["T"], c["M"], nthreads=cores, return_dist=True)
                good = np.nanmin(d64, axis=1) >= 4e-4
                err = np.abs(aoa[:nchk].cpu().numpy() - a64)
                res["cfg2_rootmusic"]["parity"] = {"frames_checked": nchk, "near_circle_frac": float((~good).mean()),
                                                   "max_abs_err_deg_vs_float64_on_good_frames": float(err[good].max()) if good.any() else None}
            rc.close()
        del xs
        torch.cuda.empty_cache()

    # cfg4: 64-element array (tensor-core HERK covariance, one-sided block eigensolver, wide scan)
    c = CFG4
    Bc = c["frames"]
    x4, _ = synth.frames_torch(Bc, c["M"], c["N"], c["thetas"], jitter_deg=2.0, device=dev, chunk=32, seed=synth.SEED_BASE + 4)
    ch = doa.DoaChain(c["M"], c["N"], 0, 0, 0.5, c["T"], c["P"], c["K"], device=local, max_frames=Bc)
    ms = time_calls(torch, lambda: ch.run_device(x4), 10)
    ch.set_profiling(True)
    for _ in range(5):
        ch.run_device(x4)
    torch.cuda.synchronize()
    st = ch.stage_ms()
    ch.set_profiling(False)
    wb = dict(M=c["M"], N=c["N"], K=c["K"])
    gbs = ALG_BYTES_CHAIN(wb) * Bc / (ms * 1e-3) / 1e9
    herk_tflops = 3 * 2.0 * 128 * 64 * 2 * c["N"] * Bc / (st[0] * 1e-3) / 1e12        # 3xTF32 real MMA flops of [Z; W] Z^T
    res["cfg4"] = {"workload": f"{Bc} independent 64-element frames x 16384 snapshots, 8 sources, 16384-point scan, K 8",
                   "ms": ms, "frames_per_s": Bc / (ms * 1e-3), "stage_ms": {"cov_herk_tc": st[0], "jacobi": st[1], "scan_peaks": st[2]},
                   "launches": ch.launches(),
                   "roofline": {"bound": "tensor pipe (covariance) / shared-memory bandwidth (eigensolver)", "algorithmic_bytes_per_frame": ALG_BYTES_CHAIN(wb),
                                "achieved_GBps": gbs, "frac_hbm": gbs / peak, "herk_tf32_mma_TFLOPs": herk_tflops,
                                "herk_input_GBps": 8.0 * c["M"] * c["N"] * Bc / (st[0] * 1e-3) / 1e9},
                   "cpu_baseline": cpu_rate(c, 0, 1)}
    ch.close()
    del x4
    torch.cuda.empty_cache()
    return res


def bench_cfg5(doa, synth, sharding, torch, dist, np, dev, local, rank, world, peak, args, fence, max_over_ranks):
    """BASELINE configs[4]: 1,048,576 16-element frames x 1024 snapshots, STRONG-scaled: rank r owns frames
    shard_range(total, r, world) (generated on its device, resident before the timed region), one gather of the peak indices per
    step.  value = total frames / max-over-ranks step time."""
    c = CFG5
    total = args.cfg5_frames
    lo, hi = sharding.shard_range(total, rank, world)
    Bl = hi - lo
    M, N, T, P, K = c["M"], c["N"], c["T"], c["P"], c["K"]
    x5, _ = synth.frames_torch(Bl, M, N, c["thetas"], jitter_deg=5.0, device=dev, chunk=8192, seed=synth.SEED_BASE + 5 + 1000 * rank)
    call = min(Bl, 262144)                                     # frames per chain call: bounds the per-call intermediates (R, G, u)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, device=local, max_frames=call)
    longest = -(-total // world)
    pk = sharding.PeakExchange(longest, K, dev, world=world, is_dst=(rank == 0), pipelined=False, mode=os.environ.get("DOA_GATHER", "gather"))

    def step():
        val, loc, bins = pk.begin()
        for f0 in range(0, Bl, call):
            n = min(call, Bl - f0)
            ch.run_device(x5[f0:f0 + n], out=(val[f0:f0 + n], loc[f0:f0 + n], bins[f0:f0 + n]))
        pk.submit()

    steps = 5
    for _ in range(3):
        step()
    pk.drain()
    fence()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nl = 0
    for _ in range(steps):
        step()
        nl += ch.launches() * (-(-Bl // call))
    pk.drain()
    e1.record()
    fence()
    ms = max_over_ranks([e0.elapsed_time(e1) / steps])[0]
    # kernel-only time of this rank's shard (no gather), and the stage split
    kms = time_calls(torch, lambda: [ch.run_device(x5[f0:f0 + min(call, Bl - f0)]) for f0 in range(0, Bl, call)], 3, warmup=1)
    ch.set_profiling(True)
    for _ in range(2):
        ch.run_device(x5[:call])
    torch.cuda.synchronize()
    st = ch.stage_ms()
    ch.set_profiling(False)
    kms = max_over_ranks([kms])[0]
    bytes_pf = 8 * M * N + 8 * K
    res = {"workload": f"{total} independent 16-element frames x 1024 snapshots, 3 sources, 4096-point scan, K 3 (T, P, K assumed: BASELINE.json gives none), strong-scaled over {world} rank(s)",
           "scaling": "strong", "n_ranks": world, "frames_total": total, "frames_per_gpu": Bl, "ms_per_step": ms,
           "frames_per_s": total / (ms * 1e-3), "msamples_per_s_per_stream": total * N / (ms * 1e-3) / 1e6,
           "kernel_ms_per_step_per_gpu": kms, "exposed_gather_ms": ms - kms, "launches_per_step": ch.launches() * (-(-Bl // call)),
           "stage_ms_per_call": {"frames": call, "cov": st[0], "jacobi": st[1], "scan_peaks": st[2]},
           "roofline": {"bound": "fp32 pipe (8.5 flop/B covariance + 16x16 eigensolver), reported against HBM", "algorithmic_bytes_per_frame": bytes_pf,
                        "achieved_GBps_per_gpu": bytes_pf * Bl / (ms * 1e-3) / 1e9, "frac": bytes_pf * Bl / (ms * 1e-3) / 1e9 / peak,
                        "frac_kernels_only": bytes_pf * Bl / (kms * 1e-3) / 1e9 / peak},
           "timer": "CUDA events, max over ranks; inputs resident (generated per shard on the device); gather of the peaks inside the timed region"}
    if rank == 0 and world == 1 and not args.no_cpu:
        mod, kind = cpu_arm()
        cores = mod.max_threads()
        n = 16 * cores
        fr = x5[:n].cpu().numpy()
        dt = cpu_chain(mod, kind, fr, 0, 0.5, T, P, K, cores)[3]
        res["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind, "sample": f"first {n} frames of the shard"}
    ch.close()
    del x5, pk
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    sys.exit(main())
