#!/usr/bin/env python
"""bench.py -- the DoA hot path on B200: autocorrelate -> MUSIC_lin_array -> find_local_max, peaks only.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 65,536 independent 8-element ULA frames
x 2048 snapshots (8 GiB of complex64, resident in HBM before the timed region), 3 sources at 10 dB SNR, 4096-point angle
scan, K = 3 peaks.  A step = one pass of the whole chain over that batch.  For N > 1 every rank owns a batch of the same
size (weak scaling; frames are independent, no data-path collective) and ONE gather of the per-frame peaks to rank 0
closes each step.  Inputs (8 GiB) are far larger than L2 (126 MB), so no flush is needed between iterations.

JSON keys beyond the base contract: roofline (the covariance kernel -- the dominant one -- against measured HBM copy
bandwidth, plus the whole chain's algorithmic bytes / step time), cpu_baseline (the CPU oracle = LAPACK restatement of
the reference, timed on this box's host cores on a bounded sample of the same frames), e2e (the same chain through the
host-pointer C-ABI call, H2D/D2H inside the timed region), clocks, gpu_launches.

--impl reference times the reference arm: the reference's CPU algorithm (oracle port; the reference itself needs GNU
Radio + Armadillo and cannot be built here) on all host cores, bounded sample per step.
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
ALG_BYTES_CHAIN = lambda w: 8 * w["M"] * w["N"] + 8 * w["K"]              # samples in once + K (value, location) out
ALG_BYTES_COV = lambda w: 8 * w["M"] * w["N"] + 8 * w["M"] * w["M"]       # autocorrelate stage: samples in + M*M complex out


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
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}

    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def run_reference(args, w):
    """Reference arm: the reference's CPU algorithm (oracle port) on all host cores; each step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O
    from gr_doa_b200 import synth
    cores = O.max_threads()
    per_step = 128 * cores                      # ~0.7 ms/frame/core -> ~0.1 s per step
    fr, _ = synth.frames_numpy(per_step, w["M"], w["N"], w["thetas"], d=w["d"], snr_db=w["snr_db"], jitter_deg=w["jitter"],
                               seed=synth.SEED_BASE + 3)
    for _ in range(args.warmup):
        O.chain_frames(fr, 0, w["d"], w["T"], w["P"], w["K"], nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.chain_frames(fr, 0, w["d"], w["T"], w["P"], w["K"], nthreads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(w), "sample_frames_per_step": per_step},
        "msamples_per_s_per_stream": value * w["N"] / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames of the workload per step x {args.steps} steps, OpenMP over frames, BLAS single-threaded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


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

    B, M, N, T, P, K = w["frames"], w["M"], w["N"], w["T"], w["P"], w["K"]
    x, _ = synth.frames_torch(B, M, N, w["thetas"], d=w["d"], snr_db=w["snr_db"], jitter_deg=w["jitter"],
                              seed=synth.SEED_BASE + 3 + 1000 * rank, device=dev)
    chain = doa.DoaChain(M, N, 0, 0, w["d"], T, P, K, device=local, max_frames=B)
    # the one collective of the path: packed peaks of every shard to rank 0, double-buffered on a side stream so that the
    # gather of step i runs under the chain kernel of step i+1 (every gather completes inside the timed region: drain())
    # measured (B200 x8 box): the serial gather costs 0.03 ms per step at 2 GPUs, 0.24 ms at 8; pipelined with 2 SMs left to
    # NCCL: 8 GPUs 1.87-1.95 -> 1.74 ms per step, 4 GPUs no change, 2 GPUs 2-3 % slower (the reserved SMs)
    pipelined = os.environ.get("DOA_PIPELINE", "1" if world > 4 else "0") != "0"
    peaks = sharding.PeakExchange(B, K, dev, world=world, is_dst=(rank == 0), pipelined=pipelined)
    if world > 1 and pipelined:
        _lib.lib().doa_cuda_dev_set(b"chain_sms_reserve", int(os.environ.get("DOA_SMS_RESERVE", "2")))
    out = peaks.bufs[0].outputs()
    total = B * world

    def step():
        nonlocal out
        out = peaks.begin()
        chain.run_device(x, out=out)
        peaks.submit()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    fused = chain.launches() == 1          # one persistent kernel for the whole chain: the stage timers read (0, 0, total)
    if world > 1:
        t = torch.tensor([ms, cov_ms, eig_ms, scan_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, cov_ms, eig_ms, scan_ms = [float(v) for v in t.tolist()]
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3)

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
        fence()
        t0 = time.perf_counter()
        for _ in range(esteps):
            chain.run_host(hx, out=hout)      # synchronous: returns when the peaks are in host memory
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": eb * world * esteps / dt, "unit": UNIT, "h2d_bytes_per_step": eb * M * N * 8, "d2h_bytes_per_step": eb * K * 12,
               "frames_per_step_per_gpu": eb, "steps": esteps, "ms_per_step": dt / esteps * 1e3,
               "api": "doa_cuda_chain_run (host pointers, chunked H2D overlapped with kernels, synchronous)",
               "timer": "host wall clock around the synchronous calls, max over ranks"}
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

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample of the same frames (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        cores = O.max_threads()
        ns = min(B, max(2048, 1024 * cores))        # ~0.35 ms per frame and core: about 10 s of CPU work in under a second of wall clock
        sub = x[:ns].cpu().numpy()
        O.chain_frames(sub[:256], 0, w["d"], T, P, K, nthreads=cores)
        t0 = time.perf_counter()
        v_o, l_o, b_o = O.chain_frames(sub, 0, w["d"], T, P, K, nthreads=cores)
        dt = time.perf_counter() - t0
        same = float((np.sort(out[2][:ns].cpu().numpy(), 1) == np.sort(b_o, 1)).all(1).mean())
        cpu = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ns} frames of the timed batch, OpenMP over frames ({cores} threads), BLAS single-threaded",
               "peak_bins_identical_frac": same}

    if rank == 0:
        peak, peak_src = measured_peaks()
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
            "config": {"workload": workload_text(w), "frames_per_gpu": B, "parallelism": f"frames sharded x{world}, one peak gather per step" + (" (double-buffered on a side stream, under the next step's kernel; 2 SMs left to NCCL)" if (world > 1 and pipelined) else ""),
                       "l2": "inputs (8 GiB/GPU) larger than L2 (126 MB): no flush between iterations",
                       "timer": "CUDA events on the launching stream, max over ranks"},
            "msamples_per_s_per_stream": value * N / 1e6,
            "msamples_per_s_aggregate": value * N * M / 1e6,
            "roofline": {"bound": "hbm", "achieved": cov_gbs, "peak": peak, "unit": "GB/s", "frac": cov_gbs / peak,
                         "traffic": traffic,
                         "kernel": ("chain_ws_kernel<8> (covariance + Jacobi + scan/peaks in one persistent warp-specialised kernel)" if fused
                                    else "cov_small_kernel<8> (covariance, dominant)"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes, "launch_ms": kern_ms,
                         "stage_ms": ({"fused_chain": kern_ms} if fused else {"cov": cov_ms, "eig": eig_ms, "scan_peaks": scan_ms}),
                         "chain": {"algorithmic_bytes_per_frame": ALG_BYTES_CHAIN(w), "achieved": chain_gbs,
                                   "frac": chain_gbs / peak, "note": "whole step (chain kernel" + ("" if fused else "s") + (" + peak gather" if world > 1 else "") + ") per GPU against the same HBM peak"}},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "sc16_input": sc16,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
