"""Small fixed workload for ncu: the chain at the cfg4 shape (M 64, N 16384, T 8, P 16384, K 8) on B frames (default 512): tensor-core
HERK, jacobi_os_block_kernel<64>, scan_peaks_wide_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M, N, T, P, K = 64, 16384, 8, 16384, 8
x, _ = synth.frames_torch(B, M, N, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
for _ in range(3):
    ch.run_device(x)
torch.cuda.synchronize()
ch.set_profiling(True)
for _ in range(5):
    ch.run_device(x)
torch.cuda.synchronize()
print("stage ms (cov, eig, scan):", ch.stage_ms(), "B", B, "launches/call", ch.launches())
