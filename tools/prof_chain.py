"""Small fixed workload for ncu: the chain on B frames of cfg3 (default 65536), warm-up + a few measured calls.
usage: prof_chain.py [B [fused [sc16]]]   (third argument "sc16": the same frames quantised to int16 I/Q)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
M, N, T, P, K = 8, 2048, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40, 90, 140], jitter_deg=5, device="cuda")
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
if len(sys.argv) > 3 and sys.argv[3] == "sc16":
    x = torch.view_as_real(x).mul(8192.0).round_().clamp_(-32768, 32767).to(torch.int16)
    ch.set_input_format("sc16", 1.0 / 32768)
doa.set_default_option("fused", fused)
for it in range(4):
    out = ch.run_device(x)
torch.cuda.synchronize()
ch.set_profiling(True)
for it in range(10):
    out = ch.run_device(x)
torch.cuda.synchronize()
print("stage ms (cov, eig, scan) mean of 10 [fused: (0, 0, total)]:", ch.stage_ms(), "B", B, "launches/call", ch.launches())
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip())
