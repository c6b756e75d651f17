"""sc16 against fc32 input at the cfg3 and cfg1 shapes: device-resident chain time (CUDA events) and the host-pointer
chain (pinned host frames, H2D inside the timed region).  Same frames, quantised to int16 for the sc16 arm.

    python tools/sc16_time.py [--no-e2e]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gr_doa_b200 as doa
from gr_doa_b200 import _lib, synth

L = _lib.lib()
S15 = 1.0 / 32768
SHAPES = {
    "cfg3 (M8 T3 N2048 P4096 K3)": dict(B=65536, M=8, N=2048, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0]),
    "cfg1 (M4 T1 N2048 P2048 K1)": dict(B=131072, M=4, N=2048, T=1, P=2048, K=1, th=[60.0]),
    "cfg5 shard (M16 T3 N1024 P4096 K3)": dict(B=65536, M=16, N=1024, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0]),
}
only = [a for a in sys.argv[1:] if not a.startswith("--")]
if only:
    SHAPES = {k: v for k, v in SHAPES.items() if any(o in k for o in only)}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, c in SHAPES.items():
    B, M, N, T, P, K = c["B"], c["M"], c["N"], c["T"], c["P"], c["K"]
    x, _ = synth.frames_torch(B, M, N, c["th"], jitter_deg=2.0, device="cuda")
    q = torch.view_as_real(x).mul(8192.0).round_().clamp_(-32768, 32767).to(torch.int16)
    del x
    fc = torch.view_as_complex(q.to(torch.float32).mul_(S15))
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    out = None
    res = {}
    ref = [t.clone() for t in ch.run_device(fc)]
    res["fc32_ms"] = round(timed(lambda: ch.run_device(fc, out=out), 20), 4)
    ch.set_input_format("sc16", S15)
    got = ch.run_device(q)
    res["identical"] = all(torch.equal(a, b) for a, b in zip(got, ref))
    res["sc16_ms"] = round(timed(lambda: ch.run_device(q), 20), 4)
    res["launches"] = ch.launches()
    for k, bytes_per in (("fc32", 8), ("sc16", 4)):
        gb = B * M * N * bytes_per / 1e9
        res[k + "_GBps"] = round(gb / res[k + "_ms"] * 1e3)
        res[k + "_frames_per_s"] = round(B / res[k + "_ms"] * 1e3)
    if "--no-e2e" not in sys.argv and M == 8:
        Bh = 16384
        for k, src in (("sc16", q[:Bh]), ("fc32", fc[:Bh])):
            ch.set_input_format(k, S15)
            host = torch.empty(src.shape, dtype=src.dtype).pin_memory()
            host.copy_(src)
            ch.run_host(host)
            t0 = time.perf_counter()
            for _ in range(5):
                ch.run_host(host)
            dt = (time.perf_counter() - t0) / 5
            res[k + "_e2e_frames_per_s"] = round(Bh / dt)
            res[k + "_e2e_GBps_h2d"] = round(host.numel() * host.element_size() / dt / 1e9, 1)
    print(name, json.dumps(res), flush=True)
    del q, fc, ch
    torch.cuda.empty_cache()
