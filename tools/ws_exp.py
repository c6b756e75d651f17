"""Fused-kernel configuration sweep at cfg3 (dev knobs ws_split = producers*100 + consumers, ws_stages, ws_nbuf; a configuration is written split|stages|nbuf, e.g. 80824), interleaved
round-robin so that clock/thermal drift hits every configuration alike.  usage: ws_exp.py 8083 4123 ..."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
doa.dev_library().__enter__()   # the -DDOA_DEV_KNOBS build (python -m gr_doa_b200.build --dev): experimental kernel variants
L = _lib.lib()
B, M, N, T, P, K = 65536, 8, 2048, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
cfgs = sys.argv[1:] or ["80824"]            # a trailing 't' = bulk (UBLKCP) ring fills instead of per-lane cp.async (80824 only); 'm' = tensor-map TMA (UTMALDG) fills; 'c' = channel-major fills
def select(cs):
    doa.set_default_option("tma", 0)   # the sweep selects the fill explicitly
    doa.set_default_option("ws_tma", 1 if cs.endswith("t") else 2 if cs.endswith("m") else 0)     # 'm': one tensor-map TMA box per stage
    doa.set_default_option("ws_fill", 2 if cs.endswith("c") else 0)
    c = int(cs.rstrip("tcm"))
    doa.set_default_option("ws_split", c // 100); doa.set_default_option("ws_stages", (c // 10) % 10); doa.set_default_option("ws_nbuf", c % 10)
ref, same, times, launches = None, {}, {c: [] for c in cfgs}, {}
for c in cfgs:
    select(c)
    for _ in range(3): out = ch.run_device(x)
    torch.cuda.synchronize()
    out = [t.clone() for t in out]
    if ref is None: ref = out
    same[c] = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out, ref))
    launches[c] = ch.launches()
for rnd in range(6):
    for c in cfgs:
        select(c)
        ch.run_device(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): ch.run_device(x)
        e1.record(); torch.cuda.synchronize()
        times[c].append(e0.elapsed_time(e1) / 8)
for c in cfgs:
    ms, med = min(times[c]), statistics.median(times[c])
    print(f"split|stages|nbuf={c}: min {ms:.4f} ms  median {med:.4f} ms  {B/med/1e3:.2f} M frames/s  frac {B*131096/med/1e6/6542.7:.3f}  launches {launches[c]}  bit-identical to first: {same[c]}", flush=True)
