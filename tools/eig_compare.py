"""Comparison timing asked for by the north star: the library eigensolver (cuSOLVER, reached through torch.linalg.eigh, which
dispatches batched small Hermitian problems to syevjBatched / XsyevBatched) against jacobi_group_kernel on the covariance
batch of cfg3 (65,536 8x8) and of the cfg5 shard (65,536 16x16).  Timed as a comparison only; the product never calls it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gr_doa_b200 import synth
import gr_doa_b200 as doa

def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for (M, N, T, B) in ((8, 2048, 3, 65536), (16, 1024, 3, 65536)):
    x, _ = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=2.0, device="cuda", chunk=2048)
    R = doa.autocorrelate(M, N, 0, 0, max_frames=B).work_device(x)            # [B][M*M] column-major
    Rm = R.view(B, M, M).transpose(1, 2).contiguous()                          # row-major Hermitian for torch
    mus = doa.MUSIC_lin_array(0.5, T, M, 1024, max_frames=B)
    ours = ev(lambda: mus.noise_subspace_device(R))
    CH = int(os.environ.get("EIG_CHUNK", 8192))       # XsyevBatched rejects the whole batch in one call (INVALID_VALUE at 65,536)
    def lib_all():
        return [torch.linalg.eigh(Rm[i:i + CH]) for i in range(0, B, CH)]
    print(f"M={M}: jacobi_group_kernel {ours:.3f} ms", flush=True)
    lib = ev(lib_all, reps=2)
    V = torch.cat([p[1] for p in lib_all()], 0)
    En = V[:, :, : M - T]
    G_lib = En @ En.conj().transpose(1, 2)
    G_ours = mus.noise_subspace_device(R)[0].view(B, M, M).transpose(1, 2)
    err = (G_lib - G_ours).abs().amax().item()
    print(f"M={M}: jacobi_group_kernel (EVD + noise projector + diagonal sums) {ours:.3f} ms; torch.linalg.eigh -> cuSOLVER in chunks of {CH} (EVD only) {lib:.2f} ms "
          f"for {B} matrices; max |G_ours - G_lib| = {err:.2e}", flush=True)
