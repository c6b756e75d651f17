"""Small fixed workload for ncu: the cfg3 chain on B frames (default 8192), 3 warm-up runs + 2 measured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
M, N, T, P, K = 8, 2048, 3, 4096, 3
if len(sys.argv) > 2:
    M, N, T, P, K = [int(v) for v in sys.argv[2:7]]
x, _ = synth.frames_torch(B, M, N, [40, 90, 140][:T] if T <= 3 else list(range(30, 151, 120 // (T - 1)))[:T], jitter_deg=5, device="cuda")
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
ch.set_profiling(True)
for it in range(5):
    out = ch.run_device(x)
torch.cuda.synchronize()
print("stage ms (cov, eig, scan):", ch.stage_ms(), "B", B)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip())
for rep in range(3):
    ch.set_profiling(True)
    for it in range(20):
        out = ch.run_device(x)
    torch.cuda.synchronize()
    print("stage ms (cov, eig, scan) mean of 20:", ch.stage_ms())
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip())
