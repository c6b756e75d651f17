"""SURVEY section 8(d), CPU baseline: the oracle port (the reference's Armadillo/LAPACK arithmetic: cgemm-order covariance,
cheevd, the P three-factor products, cgeev, the find_local_max passes) timed on THIS box's host cores at every BASELINE.json
shape -- one thread and all cores (OpenMP over frames, BLAS single-threaded) -- on a bounded sample of frames.
Needs no GPU; run it on the GPU box so the numbers sit next to the GPU ones (gpurun -- python tools/cpu_baseline.py).
Measurement only: the product never calls the oracle."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from gr_doa_b200 import synth
from oracle import oracle as O

cores = O.max_threads()
SHAPES = [
    # name, M, N, T, P, K, avg, thetas, root?
    ("cfg1 MUSIC chain (M4 T1 N2048 ov512 P2048 K1, fwd)", 4, 2048, 1, 2048, 1, 0, [60.0], False),
    ("cfg1 MUSIC chain, forward-backward", 4, 2048, 1, 2048, 1, 1, [60.0], False),
    ("cfg2 Root-MUSIC chain (M4 T2 N2048 FB)", 4, 2048, 2, 0, 0, 1, [50.0, 110.0], True),
    ("cfg3 MUSIC chain (M8 T3 N2048 P4096 K3)", 8, 2048, 3, 4096, 3, 0, [40.0, 90.0, 140.0], False),
    ("cfg5 MUSIC chain (M16 T3 N1024 P4096 K3)", 16, 1024, 3, 4096, 3, 0, [40.0, 90.0, 140.0], False),
    ("cfg4 MUSIC chain (M64 T8 N16384 P16384 K8)", 64, 16384, 8, 16384, 8, 0, [30.0 + 120.0 * i / 7 for i in range(8)], False),
]
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0      # seconds of CPU work per measurement, roughly


def run(fr, M, T, P, K, avg, root, nt):
    if root:
        R = O.autocorrelate_frames(fr, avg, nthreads=nt)
        return O.rootmusic(R, 0.5, T, M, nthreads=nt)
    return O.chain_frames(fr, avg, 0.5, T, P, K, nthreads=nt)


print(json.dumps({"host_cores_used": cores, "nproc": os.cpu_count()}))
for name, M, N, T, P, K, avg, th, root in SHAPES:
    probe = 8 if M < 64 else 2
    fr, _ = synth.frames_numpy(probe, M, N, th, snr_db=10.0, seed=1)
    run(fr, M, T, P, K, avg, root, 1)
    t0 = time.perf_counter(); run(fr, M, T, P, K, avg, root, 1); per = (time.perf_counter() - t0) / probe
    res = {"shape": name, "ms_per_frame_one_thread_probe": round(per * 1e3, 3)}
    for tag, nt in (("one_thread", 1), ("all_cores", cores)):
        n = int(max(nt * 2, min(8192, budget * nt / per)))
        fr, _ = synth.frames_numpy(n, M, N, th, snr_db=10.0, seed=2)
        t0 = time.perf_counter(); run(fr, M, T, P, K, avg, root, nt); dt = time.perf_counter() - t0
        res[tag] = {"frames": n, "frames_per_s": round(n / dt, 1), "threads": nt}
    print(json.dumps(res), flush=True)
