"""Root-MUSIC: Aberth-Ehrlich fast path against the Hessenberg-QR kernel (dev knob root_aberth), time and agreement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
for (B, M, N, T, th, snr) in ((65536, 8, 2048, 3, [40.0, 90.0, 140.0], 10.0), (262144, 4, 2048, 2, [50.0, 110.0], 10.0), (20000, 8, 512, 7, list(np.linspace(20, 160, 7)), 20.0),
                              (20000, 16, 256, 5, list(np.linspace(30, 150, 5)), 10.0), (20000, 8, 2048, 3, [40.0, 90.0, 140.0], 40.0)):
    x, _ = synth.frames_torch(B, M, N, th, snr_db=snr, jitter_deg=2.0, device="cuda", chunk=4096)
    R = doa.autocorrelate(M, N, 0, 0, max_frames=B).work_device(x)
    rm = doa.rootMUSIC_linear_array(0.5, T, M, max_frames=B)
    res = {}
    for ab in (0, 1):
        doa.set_default_option("root_aberth", ab)
        for _ in range(2): out = rm.work_device(R)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): out = rm.work_device(R)
        e1.record(); torch.cuda.synchronize()
        res[ab] = (out.clone(), e0.elapsed_time(e1) / 3)
    a, b = res[0][0], res[1][0]
    both = torch.isfinite(a) & torch.isfinite(b)
    d = (a - b).abs()[both]
    print(f"M={M} T={T} snr={snr} B={B}: QR {res[0][1]:.3f} ms, Aberth(+QR fallback) {res[1][1]:.3f} ms; max |angle diff| {d.max().item():.2e} deg, "
          f"frames > 1e-4 deg: {int((d.view(-1) > 1e-4).sum())}, NaN pattern equal: {bool((torch.isfinite(a) == torch.isfinite(b)).all())}", flush=True)
doa.set_default_option("root_aberth", 1)
