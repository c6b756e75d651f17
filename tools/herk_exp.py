import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from gr_doa_b200 import synth, _lib
from oracle import oracle as O
import gr_doa_b200 as doa
L = _lib.lib()
def relfro(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for (B, N, avg) in ((2, 64, 0), (3, 4096, 1), (5, 1000, 0), (300, 2048, 0), (3, 16384, 0), (149, 34, 1)):
    fr, _ = synth.frames_numpy(B, 64, N, [40.0, 75.0, 120.0], snr_db=10.0, seed=N + B)
    R_o = O.autocorrelate_frames(fr, avg, nthreads=8)
    x = torch.from_numpy(fr).cuda()
    ac = doa.autocorrelate(64, N, 0, avg, max_frames=B)
    doa.set_default_option("herk_tc", 0); R0 = ac.work_device(x).cpu().numpy()
    doa.set_default_option("herk_tc", 1); R1 = ac.work_device(x).cpu().numpy()
    print(f"B={B} N={N} avg={avg}: CUDA-core relfro {relfro(R0, R_o):.2e}  tensor-core relfro {relfro(R1, R_o):.2e}  max|diff| {np.abs(R1-R_o).max():.2e}", flush=True)
B, N = 592, 16384
x, _ = synth.frames_torch(B, 64, N, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
ac = doa.autocorrelate(64, N, 0, 0, max_frames=B)
for tc in (0, 1):
    doa.set_default_option("herk_tc", tc)
    for _ in range(2): R = ac.work_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): R = ac.work_device(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = B * 64 * N * 8 / 1e9
    flops = 3 * 2 * 128 * 64 * 2 * N * B
    print(f"herk_tc={tc}: {ms:.3f} ms for {B} frames ({gb:.2f} GB): {gb/ms*1e3:.0f} GB/s = {gb/ms*1e3/6542.7:.3f} of HBM; 3xTF32 real-MMA rate {flops/ms/1e9:.0f} TFLOP/s")
