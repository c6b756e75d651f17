"""Small fixed workload for ncu: the chain at the cfg5 shape (M 16, N 1024, P 4096, K 3) on B frames (default 65536): the three
kernels of the shipped path (cov16_ring_kernel, jacobi_group_kernel<16>, scan_tc_kernel<16>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
M, N, T, P, K = 16, 1024, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, device="cuda", chunk=8192)
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
for _ in range(3):
    ch.run_device(x)
torch.cuda.synchronize()
ch.set_profiling(True)
for _ in range(5):
    ch.run_device(x)
torch.cuda.synchronize()
print("stage ms (cov, eig, scan):", ch.stage_ms(), "B", B, "launches/call", ch.launches())
