import os, sys
sys.path.insert(0, "/root/repo")
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
for (B, N, avg) in ((7, 66, 1), (65536, 1024, 0), (1000, 2048, 1)):
    x, _ = synth.frames_torch(B, 16, N, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
    ac = doa.autocorrelate(16, N, 0, avg, max_frames=B)
    res = {}
    for ring in (0, 1):
        doa.set_default_option("cov16_ring", ring)
        for _ in range(2): R = ac.work_device(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): R = ac.work_device(x)
        e1.record(); torch.cuda.synchronize()
        res[ring] = (R.clone(), e0.elapsed_time(e1) / 5)
    same = torch.equal(res[0][0].view(torch.float32), res[1][0].view(torch.float32))
    gb = B * 16 * N * 8 / 1e9
    print(f"B={B} N={N} avg={avg}: LDG {res[0][1]:.3f} ms, ring {res[1][1]:.3f} ms ({gb/res[1][1]*1e3:.0f} GB/s), bit-identical {same}", flush=True)
