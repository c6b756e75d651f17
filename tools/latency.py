"""Per-call latency of the host-pointer entry points (what a GNU Radio work() pays) at scheduler-sized batches, next to the CPU
restatement of the reference on one core.  cfg1 shape (4 elements, 2048 snapshots, overlap 512, P = 2048, K = 1) and cfg3
shape (8 elements, P = 4096, K = 3).  Host wall clock, median of 200 calls after 20 warm-up calls."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gr_doa_b200 import synth
import gr_doa_b200 as doa
from oracle import oracle as O

def med(fn, n=200, warm=20):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6

for name, (M, N, T, P, K, th) in {"cfg1 shape": (4, 2048, 1, 2048, 1, [60.0]), "cfg3 shape": (8, 2048, 3, 4096, 3, [40.0, 90.0, 140.0])}.items():
    for n in (1, 8, 64):
        fr, _ = synth.frames_numpy(n, M, N, th, snr_db=10.0, seed=n)
        ac = doa.autocorrelate(M, N, 0, 0, max_frames=64)
        mu = doa.MUSIC_lin_array(0.5, T, M, P, max_frames=64)
        fl = doa.find_local_max(K, P, 0.0, 180.0, max_frames=64)
        ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=64)
        streams = np.ascontiguousarray(fr.transpose(1, 0, 2).reshape(M, n * N))     # the same frames as M channel streams
        R = ac.work(streams)
        S = mu.work(R)
        t_ac = med(lambda: ac.work(streams))
        t_mu = med(lambda: mu.work(R))
        t_fl = med(lambda: fl.work(S))
        t_ch = med(lambda: ch.run_host(fr))
        # the same call with the caller's buffer page-locked once (doa_cuda_pin_host_buffer), as a flowgraph could do for its ring buffers
        from gr_doa_b200 import _lib
        pinned = _lib.lib().doa_cuda_pin_host_buffer(fr.ctypes.data, fr.nbytes) == 0
        t_chp = med(lambda: ch.run_host(fr)) if pinned else float("nan")
        if pinned: _lib.lib().doa_cuda_unpin_host_buffer(fr.ctypes.data)
        t_cpu = med(lambda: O.chain_frames(fr, 0, 0.5, T, P, K, nthreads=1), n=20, warm=2)
        print(f"{name}, {n:2d} frames per call: autocorrelate {t_ac:7.1f} us  MUSIC {t_mu:7.1f} us  find_local_max {t_fl:7.1f} us  |  fused chain call {t_ch:7.1f} us (buffer page-locked: {t_chp:7.1f} us)"
              f"  |  CPU restatement, 1 core {t_cpu:8.1f} us", flush=True)
