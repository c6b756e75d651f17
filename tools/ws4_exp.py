"""Fused chain at M = 4 (cfg2 shape through MUSIC: 262,144 frames x 4 x 2048, FB averaging, P = 1024, K = 2): dev knob ws4."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
doa.dev_library().__enter__()   # the -DDOA_DEV_KNOBS build (python -m gr_doa_b200.build --dev): experimental kernel variants
L = _lib.lib()
B, M, N, T, P, K = 262144, 4, 2048, 2, 1024, 2
x, _ = synth.frames_torch(B, M, N, [50.0, 110.0], jitter_deg=2.0, device="cuda", chunk=4096)
ch = doa.DoaChain(M, N, 0, 1, 0.5, T, P, K, max_frames=B)
cfgs = [int(a) for a in sys.argv[1:]] or [0, 1]
ref, same, times, launches = None, {}, {c: [] for c in cfgs}, {}
for c in cfgs:
    doa.set_default_option("ws4", c)
    for _ in range(2): out = ch.run_device(x)
    torch.cuda.synchronize()
    out = [t.clone() for t in out]
    if ref is None: ref = out
    same[c] = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out, ref))
    launches[c] = ch.launches()
for rnd in range(4):
    for c in cfgs:
        doa.set_default_option("ws4", c)
        ch.run_device(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ch.run_device(x)
        e1.record(); torch.cuda.synchronize()
        times[c].append(e0.elapsed_time(e1) / 5)
gb = B * (M * N * 8 + 8 * K) / 1e9
for c in cfgs:
    med = statistics.median(times[c])
    print(f"ws4={c}: min {min(times[c]):.4f} ms  median {med:.4f} ms  frac of HBM {gb/med*1e3/6542.7:.3f}  launches {launches[c]}  bit-identical to first: {same[c]}", flush=True)
doa.set_default_option("ws4", 0)
