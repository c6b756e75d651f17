"""Tensor-core scan (scan_tc.cu) against the Horner scan and the oracle: peak bins, values, and time.  Not a test.

    gpurun -- python tools/scan_tc_exp.py
"""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from gr_doa_b200 import synth, _lib
import gr_doa_b200 as doa
doa.dev_library().__enter__()   # the -DDOA_DEV_KNOBS build (python -m gr_doa_b200.build --dev): experimental kernel variants

L = _lib.lib()
SHAPES = {
    "cfg3": dict(M=8, N=2048, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0], avg=0),
    "cfg5": dict(M=16, N=1024, T=3, P=4096, K=3, th=[40.0, 90.0, 140.0], avg=0),
    "cfg1": dict(M=4, N=2048, T=1, P=2048, K=1, th=[60.0], avg=1),
    "cfg2m": dict(M=4, N=2048, T=2, P=1024, K=2, th=[50.0, 110.0], avg=1),
    "m6": dict(M=6, N=512, T=2, P=1024, K=4, th=[50.0, 110.0], avg=0),
}


def parity(name, c, B=2048):
    M, N, T, P, K = c["M"], c["N"], c["T"], c["P"], c["K"]
    fr, _ = synth.frames_numpy(B, M, N, c["th"], jitter_deg=5.0, snr_db=10.0, seed=synth.SEED_BASE + M)
    x = torch.from_numpy(fr).cuda()
    ch = doa.DoaChain(M, N, 0, c["avg"], 0.5, T, P, K, max_frames=B)
    doa.set_default_option("fused", 0)
    out = {}
    for tc in (0, 1):
        doa.set_default_option("scan_tc", tc)
        v, l, b = [t.cpu().numpy() for t in ch.run_device(x)]
        out[tc] = (v, l, b, ch.launches())
    doa.set_default_option("fused", 1)
    vo, lo, bo = O.chain_frames(fr, c["avg"], 0.5, T, P, K, nthreads=O.max_threads())
    res = {"shape": name, "frames": B, "launches": [out[0][3], out[1][3]]}
    for tc in (0, 1):
        v, l, b, _ = out[tc]
        same = (np.sort(b, 1) == np.sort(bo, 1)).all(1)
        res[f"tc{tc}_bins_equal_oracle"] = float(same.mean())
        res[f"tc{tc}_max_bin_dist"] = int(np.abs(np.sort(b, 1) - np.sort(bo, 1)).max())
        res[f"tc{tc}_max_val_diff_on_equal"] = float(np.abs(v[same] - vo[same]).max()) if same.any() else None
    res["tc_vs_horner_bins_equal"] = float((out[0][2] == out[1][2]).all(1).mean())
    res["tc_vs_horner_vals_equal"] = float((out[0][0] == out[1][0]).all(1).mean())
    print(json.dumps(res), flush=True)


def timing(name, c, B=65536):
    M, N, T, P, K = c["M"], c["N"], c["T"], c["P"], c["K"]
    x, _ = synth.frames_torch(B, M, N, c["th"], jitter_deg=5.0, device="cuda", chunk=max(1, 2**28 // (M * N * 8)))
    ch = doa.DoaChain(M, N, 0, c["avg"], 0.5, T, P, K, max_frames=B)
    doa.set_default_option("fused", 0)
    res = {"shape": name, "frames": B}
    for tc in (0, 1, 2, 3):   # 0: Horner scan, 1: tensor-core scan, 2: its scan phase alone, 3: long flats not redone (wrong on those frames)
        doa.set_default_option("scan_tc", min(tc, 1))
        doa.set_default_option("scan_tc_dbg", {0: 0, 1: 0, 2: 1, 3: 2}[tc])
        for _ in range(3): ch.run_device(x)
        torch.cuda.synchronize()
        ch.set_profiling(True)
        for _ in range(5): ch.run_device(x)
        torch.cuda.synchronize()
        res[f"tc{tc}_stages_ms"] = [round(v, 4) for v in ch.stage_ms()]
        ch.set_profiling(False)
    doa.set_default_option("fused", 1)
    doa.set_default_option("scan_tc_dbg", 0)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, c in SHAPES.items():
        if only and name not in only: continue
        parity(name, c)
    for name in ("cfg3", "cfg5", "cfg1"):
        if only and name not in only: continue
        timing(name, SHAPES[name])
