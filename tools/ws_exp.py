import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
B, M, N, T, P, K = 65536, 8, 2048, 3, 4096, 3
x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
ref = None
for wc in [int(a) for a in sys.argv[1:]] or [4123]:
    L.doa_cuda_dev_set(b"ws_split", wc // 10); L.doa_cuda_dev_set(b"ws_stages", wc % 10)
    for _ in range(3): out = ch.run_device(x)
    torch.cuda.synchronize()
    out = [t.clone() for t in out]
    if ref is None: ref = out
    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out, ref))
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ch.run_device(x)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    ms = min(ts)
    print(f"split/stages={wc}: {ms:.4f} ms  {B/ms/1e3:.2f} M frames/s  frac {B*131096/ms/1e6/6542.7:.3f}  launches {ch.launches()}  bit-identical to first: {same}", flush=True)
