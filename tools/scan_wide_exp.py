"""Generic-M scan: CTA-per-frame (scan_peaks_wide_kernel) against warp-per-frame, bits and time (cfg4 shape, 256 frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
for (B, M, N, T, P, K) in ((256, 64, 16384, 8, 16384, 8), (300, 64, 128, 4, 1024, 3), (77, 64, 256, 2, 2000, 2), (33, 64, 130, 5, 333, 12)):
    x, _ = synth.frames_torch(B, M, N, [30.0 + 120.0 * i / max(1, T - 1) for i in range(T)], jitter_deg=2.0, device="cuda", chunk=16)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    res = {}
    for wide in (0, 1):
        doa.set_default_option("scan_wide", wide)
        for _ in range(2): out = ch.run_device(x)
        ch.set_profiling(True)
        for _ in range(5): out = ch.run_device(x)
        torch.cuda.synchronize()
        res[wide] = ([t.clone() for t in out], ch.stage_ms())
        ch.set_profiling(False)
    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(res[0][0], res[1][0]))
    print(f"B={B} M={M} P={P} K={K}: stages (cov, eig, scan) warp-per-frame {[round(v,3) for v in res[0][1]]} ms, CTA-per-frame {[round(v,3) for v in res[1][1]]} ms, bit-identical {same}", flush=True)
doa.set_default_option("scan_wide", 1)
