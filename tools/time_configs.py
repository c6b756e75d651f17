"""Times the chain on the other BASELINE.json configs (not bench lines: parity-test shapes), for DESIGN.md."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
L = _lib.lib()
CFG = {
    "cfg1 (M4 T1 N2048 P2048 K1, independent frames)": dict(B=262144, M=4, N=2048, T=1, P=2048, K=1, th=[60.0]),
    "cfg2-shape MUSIC (M4 T2 N2048 P1024 K2, FB)": dict(B=262144, M=4, N=2048, T=2, P=1024, K=2, th=[50.0, 110.0], avg=1),
    "cfg3 (M8 T3 N2048 P4096 K3)": dict(B=65536, M=8, N=2048, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0]),
    "cfg5 shard (M16 T3 N1024 P4096 K3)": dict(B=65536, M=16, N=1024, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0]),
    "cfg4 (M64 T8 N16384 P16384 K8)": dict(B=512, M=64, N=16384, T=8, P=16384, K=8, th=[30.0 + 120.0 * i / 7 for i in range(8)]),
}
only = sys.argv[1:] 
for name, c in CFG.items():
    if only and not any(o in name for o in only): continue
    B, M, N, T, P, K = c["B"], c["M"], c["N"], c["T"], c["P"], c["K"]
    x, _ = synth.frames_torch(B, M, N, c["th"], jitter_deg=2.0, device="cuda", chunk=max(1, 2**28 // (M * N * 8)))
    ch = doa.DoaChain(M, N, 0, c.get("avg", 0), 0.5, T, P, K, max_frames=B)
    res = {}
    for fused in (0, 1):
        doa.set_default_option("fused", fused)
        for _ in range(3): ch.run_device(x)
        torch.cuda.synchronize()
        if fused == 0:
            ch.set_profiling(True)
            for _ in range(5): ch.run_device(x)
            torch.cuda.synchronize()
            res["stages_ms"] = [round(v, 4) for v in ch.stage_ms()]
            ch.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ch.run_device(x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res["fused_ms" if fused else "unfused_ms"] = round(ms, 4)
        res["launches_fused" if fused else "launches_unfused"] = ch.launches()
    best = min(res["fused_ms"], res["unfused_ms"])
    gb = B * M * N * 8 / 1e9
    res.update(frames_per_s=round(B / best * 1e3), input_GB=round(gb, 2), GBps=round(gb / best * 1e3), frac_hbm=round(gb / best * 1e3 / 6542.7, 3))
    print(name, json.dumps(res))
    if M in (4,8,16):
        rm = doa.rootMUSIC_linear_array(0.5, T, M, max_frames=B)
        ac = doa.autocorrelate(M, N, 0, c.get("avg", 0), max_frames=B)
        R = ac.work_device(x)
        for _ in range(2): rm.work_device(R)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3): rm.work_device(R)
        e1.record(); torch.cuda.synchronize()
        print("   rootMUSIC (EVD + companion QR) ms per batch:", round(e0.elapsed_time(e1) / 3, 3), " frames/s", round(B / (e0.elapsed_time(e1) / 3) * 1e3))
    del x, ch
    torch.cuda.empty_cache()
