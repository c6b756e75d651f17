"""Standalone block kernels at cfg3 / cfg1 size (device-resident): time and fraction of the HBM roofline of each block's own bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa

def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, (B, M, N, T, P, K, th) in {"cfg3": (65536, 8, 2048, 3, 4096, 3, [40.0, 90.0, 140.0]), "cfg1": (262144, 4, 2048, 1, 2048, 1, [60.0])}.items():
    x, _ = synth.frames_torch(B, M, N, th, jitter_deg=2.0, device="cuda", chunk=4096)
    ac = doa.autocorrelate(M, N, 0, 0, max_frames=B)
    mu = doa.MUSIC_lin_array(0.5, T, M, P, max_frames=B)
    fl = doa.find_local_max(K, P, 0.0, 180.0, max_frames=B)
    rm = doa.rootMUSIC_linear_array(0.5, T, M, max_frames=B)
    R = ac.work_device(x)
    S = mu.work_device(R)
    peak = 6542.7
    t = ev(lambda: ac.work_device(x)); by = B * (M * N * 8 + M * M * 8)
    print(f"{name} autocorrelate          {t:7.3f} ms  {by/t/1e6:7.0f} GB/s  {by/t/1e6/peak:.2f} of HBM")
    t = ev(lambda: mu.work_device(R)); by = B * (M * M * 8 + P * 4)
    print(f"{name} MUSIC_lin_array (spectrum out) {t:7.3f} ms  {by/t/1e6:7.0f} GB/s  {by/t/1e6/peak:.2f} of HBM")
    t = ev(lambda: fl.work_device(S)); by = B * (P * 4 + K * 12)
    print(f"{name} find_local_max         {t:7.3f} ms  {by/t/1e6:7.0f} GB/s  {by/t/1e6/peak:.2f} of HBM")
    t = ev(lambda: rm.work_device(R)); by = B * (M * M * 8 + T * 4)
    print(f"{name} rootMUSIC_linear_array {t:7.3f} ms  {by/t/1e6:7.0f} GB/s  {by/t/1e6/peak:.2f} of HBM", flush=True)
