"""A/B of the two eigensolvers behind launch_noise_subspace at 8 and 16 elements (option "eig_onesided"): the one-sided Jacobi
on the Cholesky factor (eig_os_device.cuh, default) against the two-sided Jacobi (eig_device.cuh).  Per shape: time of the
kernel (EVD + projector + diagonal sums), error of the projector against a float64 eigendecomposition (torch.linalg.eigh in
float64, comparison only), eigenvalue error, and the chain's peak bins against the two-sided solver's.  Then the inputs the
factorisation rejects (zero, indefinite, rank-deficient, NaN): the fallback must give the two-sided solver's bits."""
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


def g64(R, M, T, chunk=8192):
    out, ws = [], []
    for i in range(0, R.shape[0], chunk):
        Rm = R[i:i + chunk].view(-1, M, M).transpose(1, 2).to(torch.complex128)
        Rm = torch.triu(Rm) + torch.triu(Rm, 1).conj().transpose(1, 2)
        w, V = torch.linalg.eigh(Rm)
        En = V[:, :, : M - T]
        out.append((En @ En.conj().transpose(1, 2)))
        ws.append(w)
    return torch.cat(out), torch.cat(ws)


shapes = [(16, 1024, 3, 65536, 10.0), (8, 2048, 3, 65536, 10.0), (16, 1024, 3, 8192, 40.0), (8, 2048, 7, 8192, 0.0), (16, 64, 15, 8192, -5.0)]
for (M, N, T, B, snr) in shapes:
    th = [20.0 + 140.0 * i / max(T - 1, 1) for i in range(T)]
    x, _ = synth.frames_torch(B, M, N, th, snr_db=snr, jitter_deg=2.0, device="cuda", chunk=2048)
    R = doa.autocorrelate(M, N, 0, 0, max_frames=B).work_device(x)
    Gref, wref = g64(R, M, T)
    mus = doa.MUSIC_lin_array(0.5, T, M, 1024, max_frames=B)
    res = {}
    for os_ in (0, 1):
        mus.set_option("eig_onesided", os_)
        ms = ev(lambda: mus.noise_subspace_device(R))
        G, u, w = mus.noise_subspace_device(R)
        Gm = G.view(B, M, M).transpose(1, 2).to(torch.complex128)
        e = (Gm - Gref).abs().amax(dim=(1, 2))
        ew = (w.double() - wref).abs().amax().item()
        res[os_] = (G, u, w)
        print(f"M={M} N={N} T={T} B={B} snr={snr}: eig_onesided={os_}: {ms:.3f} ms; projector err vs f64 mean {e.mean().item():.2e} "
              f"p99 {e.quantile(0.99).item():.2e} max {e.max().item():.2e}; eigenvalue err {ew:.2e} (largest {wref.abs().max().item():.1f})", flush=True)
    dG = (res[0][0] - res[1][0]).abs().amax().item()
    print(f"   max |G_twosided - G_onesided| = {dG:.2e}")
    if T <= 4:
        P, K = 4096, T
        ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
        ch.set_option("fused", 0)
        bins = {}
        for os_ in (0, 1):
            ch.set_option("eig_onesided", os_)
            bins[os_] = ch.run_device(x)[2].clone()
        nd = (bins[0] != bins[1]).any(1).sum().item()
        print(f"   chain (unfused): frames whose peak bins differ between the two solvers: {nd} of {B} ({nd / B:.4%})")
    del x, R
    torch.cuda.empty_cache()

# inputs the Cholesky factorisation rejects or barely accepts
for M in (8, 16):
    T = 1
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    B = 64
    A = torch.randn((B, M, M), generator=g, device="cuda", dtype=torch.float32) + 1j * torch.randn((B, M, M), generator=g, device="cuda", dtype=torch.float32)
    Hm = (A + A.conj().transpose(1, 2)).to(torch.complex64)                  # indefinite Hermitian
    v = torch.randn((B, M, 2), generator=g, device="cuda", dtype=torch.float32)
    v = torch.view_as_complex(v.contiguous())
    lowrank = (v[:, :, None] * v[:, None, :].conj()).to(torch.complex64)      # rank one, positive semidefinite
    zero = torch.zeros((B, M, M), dtype=torch.complex64, device="cuda")
    nanm = Hm.clone(); nanm[::2, 0, 0] = float("nan")
    mixed = torch.cat([Hm[:8], lowrank[:8] + 0.5 * torch.eye(M, device="cuda")[None], Hm[8:16]])
    mus = doa.MUSIC_lin_array(0.5, T, M, 1024, max_frames=4 * B)
    for name, mat in (("indefinite", Hm), ("rank one", lowrank), ("zero", zero), ("NaN in every other", nanm), ("mixed", mixed)):
        Rin = mat.transpose(1, 2).contiguous().view(mat.shape[0], M * M)
        out = {}
        for os_ in (0, 1):
            mus.set_option("eig_onesided", os_)
            out[os_] = [t.clone() for t in mus.noise_subspace_device(Rin)]
        same = all(torch.equal(torch.nan_to_num(torch.view_as_real(a) if a.is_complex() else a, nan=7.0),
                               torch.nan_to_num(torch.view_as_real(b) if b.is_complex() else b, nan=7.0)) for a, b in zip(out[0], out[1]))
        if "NaN" in name:
            print(f"M={M} {name}: identical to the two-sided solver: {same}")
            continue
        Gref, _ = g64(Rin, M, T)
        e1 = (out[1][0].view(-1, M, M).transpose(1, 2).to(torch.complex128) - Gref).abs().amax().item()
        e0 = (out[0][0].view(-1, M, M).transpose(1, 2).to(torch.complex128) - Gref).abs().amax().item()
        print(f"M={M} {name}: identical to the two-sided solver: {same}; projector err two-sided {e0:.2e}, one-sided {e1:.2e}")
