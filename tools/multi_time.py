"""doa_cuda_multi_run (one process, every visible GPU) from pinned host memory at the cfg3 shape: frames/s end to end against
the same call on one device.  usage: multi_time.py [frames_per_device]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gr_doa_b200 as doa
from gr_doa_b200 import synth

per = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M, N, T, P, K = 8, 2048, 3, 4096, 3
G = torch.cuda.device_count()
x, _ = synth.frames_torch(per, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, device="cuda")
host = torch.empty((per * G, M, N), dtype=torch.complex64).pin_memory()
for g in range(G):
    host[g * per:(g + 1) * per].copy_(x)
del x
res = {"devices": G, "frames_per_device": per}
ref = None
for tag, devs in (("one", [0]), ("all", list(range(G)))):
    B = per * len(devs)
    mc = doa.DoaChainMulti(M, N, 0, 0, 0.5, T, P, K, devices=devs, max_frames_per_device=per)
    out = (np.empty((B, K), np.float32), np.empty((B, K), np.float32), np.empty((B, K), np.int32))
    mc.run_host(host[:B], out=out)
    t0 = time.perf_counter()
    for _ in range(5):
        mc.run_host(host[:B], out=out)
    dt = (time.perf_counter() - t0) / 5
    res[tag] = {"frames_per_s": round(B / dt), "ms_per_call": round(dt * 1e3, 2), "h2d_GBps": round(B * M * N * 8 / dt / 1e9, 1)}
    if ref is None:
        ref = [o.copy() for o in out]
    else:   # every device saw the same frames: its block must equal device 0's answer
        res["identical_to_one_device"] = all(np.array_equal(o[g * per:(g + 1) * per], r) for g in range(len(devs)) for o, r in zip(out, ref))
    mc.close()
print(json.dumps(res))
