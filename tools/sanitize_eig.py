"""Small-shape exercise of the round-2 eigensolvers and of the wide scan's CTA-wide refinement, for compute-sanitizer
(--tool memcheck / racecheck / synccheck): lane-group solver at 8 and 16 elements (partial warps, inputs the factorisation rejects
next to good ones), CTA solver at 24 and 64 elements (covariances, indefinite and zero matrices), generic-M chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa

def herm(B, M, kind, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    A = torch.randn((B, M, M), generator=g, device="cuda") + 1j * torch.randn((B, M, M), generator=g, device="cuda")
    if kind == "spd": H = A @ A.conj().transpose(1, 2) / M + 0.5 * torch.eye(M, device="cuda")
    elif kind == "indef": H = A + A.conj().transpose(1, 2)
    else: H = torch.zeros_like(A)
    return H.to(torch.complex64).transpose(1, 2).contiguous().view(B, M * M)

for M, B in ((8, 13), (16, 7), (24, 5), (64, 3)):
    mus = doa.MUSIC_lin_array(0.5, 2, M, 256, max_frames=4 * B)
    mixed = torch.cat([herm(B, M, "spd", 1), herm(B, M, "indef", 2), herm(2, M, "zero", 3), herm(B, M, "spd", 4)])[: 4 * B]
    for os_ in (1, 0):
        mus.set_option("eig_onesided", os_)
        G, u, w = mus.noise_subspace_device(mixed)
        S = mus.work_device(mixed)
    torch.cuda.synchronize()

def chain(B, M, N, T, P, K):
    x, _ = synth.frames_torch(B, M, N, list(np.linspace(40.0, 140.0, T)), jitter_deg=2.0, device="cuda", chunk=64)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    out = ch.run_device(x); torch.cuda.synchronize()
    return out
chain(5, 64, 1056, 4, 1024, 5)     # HERK + CTA eigensolver + wide scan, K > 4 (two refinement rounds)
chain(6, 24, 200, 3, 600, 3)       # tiled covariance + CTA eigensolver (MP = 32) + wide scan
chain(37, 16, 192, 3, 512, 3)      # lane-group solver inside the three-kernel chain
chain(70, 8, 256, 3, 1024, 3)      # ... and inside the fused kernel
print("sanitize_eig: done")
