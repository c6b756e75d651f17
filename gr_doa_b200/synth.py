"""Synthetic uniform-linear-array snapshots, the signal model the reference's apps and QA use.

    x[m, t] = sum_k a_m(theta_k) * s_k[t] + sigma * w[m, t]
    a_m(theta) = exp(-j*2*pi*cos(theta) * d * (M-1-2m)/2)      apps/run_MUSIC_lin_array_simulation.py:71-73,
                                                               examples/@wpi_twinrx_doa_testbench/wpi_twinrx_doa_testbench.m:60-64
    s_k[t] = exp(j*(w_k*t + phase_k)),  w_k = pi/(k+2)         tone model of music_test_input_gen.m:36-37,97
    w ~ CN(0, 1)

Two generators: numpy (host, seeded Philox; used by the parity tests so the CPU checker and the GPU see the same bytes)
and torch (any device; used by bench.py to fill HBM without a host round trip).
"""
import math

import numpy as np

SEED_BASE = 0x0D0A


def steering(thetas_deg, M, d):
    th = np.deg2rad(np.asarray(thetas_deg, dtype=np.float64))
    loc = d * 0.5 * (M - 1 - 2 * np.arange(M))
    return np.exp(-1j * 2 * np.pi * np.cos(th)[..., None] * loc)          # [..., M]


def frames_numpy(B, M, N, thetas_deg, d=0.5, snr_db=10.0, jitter_deg=0.0, seed=SEED_BASE):
    """Independent frames [B][M][N] complex64; returns (frames, true_thetas [B][T])."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    T = len(thetas_deg)
    th = np.asarray(thetas_deg, np.float64)[None, :] + (rng.uniform(-jitter_deg, jitter_deg, (B, T)) if jitter_deg else 0.0)
    th = np.broadcast_to(th, (B, T))
    A = steering(th, M, d)                                                # [B][T][M]
    w = np.pi / (np.arange(T) + 2.0)
    ph = rng.uniform(0, 2 * np.pi, (B, T))
    t = np.arange(N)
    s = np.exp(1j * (w[None, :, None] * t[None, None, :] + ph[:, :, None]))   # [B][T][N]
    x = np.einsum("btm,btn->bmn", A, s)
    sigma = math.sqrt(10.0 ** (-snr_db / 10.0))
    noise = (rng.standard_normal((B, M, N)) + 1j * rng.standard_normal((B, M, N))) * (sigma / math.sqrt(2.0))
    return (x + noise).astype(np.complex64), np.array(th)


def stream_numpy(nframes, M, N, overlap, thetas_deg, d=0.5, snr_db=10.0, seed=SEED_BASE):
    """M continuous channel streams [M][(nframes-1)*hop + N] complex64 for the streaming (hop/overlap) form."""
    hop = N - overlap
    Lx = (nframes - 1) * hop + N
    fr, _ = frames_numpy(1, M, Lx, thetas_deg, d=d, snr_db=snr_db, seed=seed)
    return fr[0]


def frames_torch(B, M, N, thetas_deg, d=0.5, snr_db=10.0, jitter_deg=0.0, seed=SEED_BASE, device="cuda", chunk=4096):
    """Same model generated on `device` in chunks (values differ from frames_numpy: different RNG)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    T = len(thetas_deg)
    out = torch.empty((B, M, N), dtype=torch.complex64, device=device)
    loc = d * 0.5 * (M - 1 - 2 * torch.arange(M, device=device, dtype=torch.float64))
    base = torch.tensor(list(thetas_deg), device=device, dtype=torch.float64)
    w = math.pi / (torch.arange(T, device=device, dtype=torch.float64) + 2.0)
    t = torch.arange(N, device=device, dtype=torch.float64)
    sigma = math.sqrt(10.0 ** (-snr_db / 10.0))
    truth = torch.empty((B, T), dtype=torch.float64, device=device)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        nb = b1 - b0
        th = base[None, :] + (torch.rand((nb, T), generator=g, device=device, dtype=torch.float64) * 2 - 1) * jitter_deg
        truth[b0:b1] = th
        A = torch.exp(-1j * 2 * math.pi * torch.cos(torch.deg2rad(th))[..., None] * loc)       # [nb][T][M]
        ph = torch.rand((nb, T), generator=g, device=device, dtype=torch.float64) * 2 * math.pi
        s = torch.exp(1j * (w[None, :, None] * t[None, None, :] + ph[:, :, None]))             # [nb][T][N]
        x = torch.einsum("btm,btn->bmn", A.to(torch.complex64), s.to(torch.complex64))
        nz = torch.view_as_complex(torch.randn((nb, M, N, 2), generator=g, device=device, dtype=torch.float32))
        out[b0:b1] = x + nz * (sigma / math.sqrt(2.0))
    return out, truth


def stream_torch(M, L, thetas_deg, d=0.5, snr_db=10.0, seed=SEED_BASE, device="cuda", chunk=1 << 22):
    """M continuous channel streams [M][L] complex64 generated on `device` in chunks along time (the streaming form:
    frames of snapshot_size samples every hop samples are read in place from these streams)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    T = len(thetas_deg)
    out = torch.empty((M, L), dtype=torch.complex64, device=device)
    loc = d * 0.5 * (M - 1 - 2 * torch.arange(M, device=device, dtype=torch.float64))
    th = torch.tensor(list(thetas_deg), device=device, dtype=torch.float64)
    A = torch.exp(-1j * 2 * math.pi * torch.cos(torch.deg2rad(th))[:, None] * loc[None, :]).to(torch.complex64)     # [T][M]
    w = math.pi / (torch.arange(T, device=device, dtype=torch.float64) + 2.0)
    ph = torch.rand((T,), generator=g, device=device, dtype=torch.float64) * 2 * math.pi
    sigma = math.sqrt(10.0 ** (-snr_db / 10.0))
    for t0 in range(0, L, chunk):
        t1 = min(L, t0 + chunk)
        t = torch.arange(t0, t1, device=device, dtype=torch.float64)
        s = torch.exp(1j * (torch.remainder(w[:, None] * t[None, :], 2 * math.pi) + ph[:, None])).to(torch.complex64)   # [T][n]
        nz = torch.view_as_complex(torch.randn((M, t1 - t0, 2), generator=g, device=device, dtype=torch.float32))
        out[:, t0:t1] = A.t() @ s + nz * (sigma / math.sqrt(2.0))
    return out
