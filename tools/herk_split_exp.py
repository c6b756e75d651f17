"""A/B of the tensor-core HERK's split tail (option "herk_split"; herk_tc.cu): batches of 64 x N frames that do not fill a
round of the 148 persistent CTAs.  Interleaved timing, bit-identity checked.  gpurun -- python tools/herk_split_exp.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa

N = int(os.environ.get("HN", 16384))
for B in [int(b) for b in os.environ.get("HB", "1,4,32,74,148,222,296,512,592").split(",")]:
    x, _ = synth.frames_torch(B, 64, N, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
    ac = doa.autocorrelate(64, N, 0, 0, max_frames=B)
    res, ms = {}, {0: [], 1: []}
    for rep in range(4):
        for split in (0, 1):
            ac.set_option("herk_split", split)
            res[split] = ac.work_device(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): res[split] = ac.work_device(x)
            e1.record(); torch.cuda.synchronize()
            ms[split].append(e0.elapsed_time(e1) / 5)
    same = bool(torch.equal(torch.view_as_real(res[0]), torch.view_as_real(res[1])))
    gb = B * 64 * N * 8 / 1e9
    a, b = min(ms[0]), min(ms[1])
    print(f"B={B} N={N}: split off {a:.3f} ms, on {b:.3f} ms ({a / b:.2f}x; {gb / b * 1e3:.0f} GB/s = {gb / b * 1e3 / 6542.7:.3f} of HBM); same bits: {same}", flush=True)
    del x, ac
