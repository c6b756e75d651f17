"""cfg3: the whole chain in one persistent kernel (shipped) against the split form -- covariance + Jacobi in the persistent kernel,
scan on the tensor cores as a second kernel (scan_tc.cu) -- for several producer/consumer splits.  Dev build.  Not a test."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
doa.dev_library().__enter__()
B, M, N, T, P, K = 65536, 8, 2048, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, device="cuda")
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
ref = None
for name, opts in (("fused (shipped)", {"fused": 1, "ws_split": 808}), ("split 8+8", {"fused": 2, "ws_split": 808}), ("split 10+6", {"fused": 2, "ws_split": 1006}),
                   ("split 12+4", {"fused": 2, "ws_split": 1204}), ("split 12+4, 3 stages", {"fused": 2, "ws_split": 1205}), ("stage kernels", {"fused": 0})):
    for k, v in opts.items(): ch.set_option(k, v)
    for _ in range(3): out = ch.run_device(x)
    torch.cuda.synchronize()
    ch.set_profiling(True)
    for _ in range(10): out = ch.run_device(x)
    torch.cuda.synchronize()
    st = ch.stage_ms(); ch.set_profiling(False)
    if ref is None: ref = [t.clone() for t in out]
    same = float((ref[2] == out[2]).all(1).float().mean())
    print(json.dumps({"variant": name, "stage_ms": [round(v, 3) for v in st], "chain_ms": round(sum(st), 3), "launches": ch.launches(), "bins_identical_frac": same,
                      "frac_hbm": round(131096 * B / (sum(st) * 1e-3) / 1e9 / 6542.7, 3)}), flush=True)
