import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
B, N = int(os.environ.get("HB", 592)), int(os.environ.get("HN", 16384))
x, _ = synth.frames_torch(B, 64, N, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
ac = doa.autocorrelate(64, N, 0, 0, max_frames=B)
doa.set_default_option("herk_tc", 1)
for _ in range(2): R = ac.work_device(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): R = ac.work_device(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gb = B * 64 * N * 8 / 1e9
print(f"mode={os.environ.get('HERK_MODE', '0')}: {ms:.3f} ms for {B} frames x {N} ({gb:.2f} GB): {gb/ms*1e3:.0f} GB/s = {gb/ms*1e3/6542.7:.3f} of HBM")
