import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa
for (B, M, N, T) in ((256, 64, 16384, 8), (64, 64, 1024, 4), (1000, 32, 512, 4), (2000, 24, 256, 3)):
    x, _ = synth.frames_torch(B, M, N, [30.0 + 120.0 * i / max(1, T - 1) for i in range(T)], jitter_deg=2.0, device="cuda", chunk=16)
    R = doa.autocorrelate(M, N, 0, 0, max_frames=B).work_device(x)
    mus = doa.MUSIC_lin_array(0.5, T, M, 1024, max_frames=B)
    for _ in range(2): out = mus.noise_subspace_device(R)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = mus.noise_subspace_device(R)
    e1.record(); torch.cuda.synchronize()
    G = out[0].view(B, M, M).transpose(1, 2)
    w, V = torch.linalg.eigh(R.view(B, M, M).transpose(1, 2).to(torch.complex128))
    En = V[:, :, : M - T]
    Gref = (En @ En.conj().transpose(1, 2)).to(torch.complex64)
    print(f"B={B} M={M}: block Jacobi {e0.elapsed_time(e1)/5:.3f} ms; max |G - G_f64| = {(G - Gref).abs().amax().item():.2e}", flush=True)
