"""Small-shape exercise of every kernel family, meant to be run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
doa.dev_library().__enter__()   # the -DDOA_DEV_KNOBS build (python -m gr_doa_b200.build --dev): experimental kernel variants
L = _lib.lib()
def chain(B, M, N, T, P, K, avg=0):
    x, _ = synth.frames_torch(B, M, N, list(np.linspace(40.0, 140.0, T)) if T > 1 else [70.0], jitter_deg=2.0, device="cuda", chunk=64)
    ch = doa.DoaChain(M, N, 0, avg, 0.5, T, P, K, max_frames=B)
    out = ch.run_device(x); torch.cuda.synchronize()
    R = doa.autocorrelate(M, N, 0, avg, max_frames=B).work_device(x)
    mu = doa.MUSIC_lin_array(0.5, T, M, P, max_frames=B); S = mu.work_device(R)
    doa.find_local_max(K, P, 0.0, 180.0, max_frames=B).work_device(S)
    doa.rootMUSIC_linear_array(0.5, T, M, max_frames=B).work_device(R)
    torch.cuda.synchronize()
    return out
chain(70, 8, 256, 3, 1024, 3)              # fused 8+8
chain(70, 4, 192, 2, 512, 2, avg=1)        # fused M = 4
chain(70, 4, 130, 1, 512, 1)               # fused, K = 1
chain(37, 16, 192, 3, 512, 3)              # cov16 ring + group Jacobi 16
chain(5, 64, 1056, 4, 1024, 5)             # tcgen05 HERK + block Jacobi + wide scan
chain(9, 12, 100, 2, 300, 2)               # generic tiled covariance
for cfg in ((412, 3, 2), (416, 4, 2), (812, 3, 2), (808, 3, 3)):
    doa.set_default_option("ws_split", cfg[0]); doa.set_default_option("ws_stages", cfg[1]); doa.set_default_option("ws_nbuf", cfg[2])
    chain(70, 8, 256, 3, 1024, 3)
doa.set_default_option("ws_split", 808); doa.set_default_option("ws_stages", 2); doa.set_default_option("ws_nbuf", 4)
doa.set_default_option("ws_tma", 1); chain(70, 8, 256, 3, 1024, 3); doa.set_default_option("ws_tma", 0)
doa.set_default_option("root_aberth", 0); chain(20, 8, 128, 3, 256, 3); doa.set_default_option("root_aberth", 1)
print("sanitize_small: done")
