"""Small-shape exercise of the tensor-core scan (scan_tc.cu) in the unfused chain, for compute-sanitizer --tool memcheck:
even / odd tile counts, the shifted last bin tile, K = 1, partial frame tiles, more frame tiles than one CTA round."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
def chain(B, M, N, T, P, K):
    x, _ = synth.frames_torch(B, M, N, list(np.linspace(40.0, 140.0, T)) if T > 1 else [70.0], jitter_deg=2.0, device="cuda", chunk=64)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    ch.set_option("fused", 0)
    out = ch.run_device(x); torch.cuda.synchronize()
    assert ch.launches() == 3
    return out
chain(70, 8, 128, 3, 1024, 3)
chain(130, 16, 128, 3, 126, 3)        # two bin tiles, the second one shifted onto the first
chain(5, 4, 128, 1, 2048, 1)          # K = 1 (arg-max)
chain(300, 6, 64, 2, 1000, 4)         # runtime-M instantiation, K = 4, P not a multiple of anything
chain(3, 2, 64, 1, 251, 2)
print("sanitize_scan_tc: done")
