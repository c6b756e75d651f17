"""fused16.cu (M = 16: covariance + Jacobi in one persistent kernel) against the two stage kernels: time and bits.  Not a test."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
doa.dev_library().__enter__()
B, M, N, T, P, K = 65536, 16, 1024, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, device="cuda", chunk=8192)
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
ref = None
for name, opts in (("stage kernels", {"fused16": 0}), ("fused16 2 pairs + 4 consumers", {"fused16": 1, "ws_split": 0}),
                   ("fused16 1 pair + 6 consumers", {"fused16": 1, "ws_split": 106}), ("fused16 1 pair + 8 consumers", {"fused16": 1, "ws_split": 108})):
    for k, v in opts.items(): ch.set_option(k, v)
    for _ in range(3): out = ch.run_device(x)
    torch.cuda.synchronize()
    ch.set_profiling(True)
    for _ in range(5): out = ch.run_device(x)
    torch.cuda.synchronize()
    st = ch.stage_ms(); ch.set_profiling(False)
    if ref is None: ref = [t.clone() for t in out]
    same = all(torch.equal(a, b) for a, b in zip(ref, out))
    print(json.dumps({"variant": name, "stage_ms": [round(v, 3) for v in st], "chain_ms": round(sum(st), 3), "launches": ch.launches(), "bit_identical": same,
                      "frac_hbm": round(131096 * B / (sum(st) * 1e-3) / 1e9 / 6542.7, 3)}), flush=True)
