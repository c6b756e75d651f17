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


# ---------------------------------------------------------------------------------------------------------------------------
# Counter-based generator: frame f of a batch is a pure function of (seed, f), built on Philox4x32-10 with
# the frame index in the counter, and written in integer / float64 arithmetic that numpy and torch (CPU or CUDA) evaluate the same
# way.  Any shard of a batch generated on a device can therefore be re-derived on the host -- bit for bit in the integer stream,
# and to the last float32 bit in the samples except where a float64 transcendental differs by one ulp between libraries right at a
# float32 rounding boundary (about one sample in 10^8).
#   counter = (j, f mod 2^32, f div 2^32, stream)      key = (seed mod 2^32, seed div 2^32)
#   stream 0, call j: the noise of samples 2j and 2j + 1 of the frame's [M][N] block (Box-Muller on two uniforms each)
#   stream 1, call j: uniforms 4j .. 4j + 3 of the frame's parameters: T jitters, then T tone phases
_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK32 = 0xFFFFFFFF


def _philox4x32(xp, c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on arrays of 32-bit values held in 64-bit integers (numpy uint64 or torch int64: the low 64 bits of the
    products are the same in both, the high word is masked after the shift)."""
    for _ in range(10):
        p0 = c0 * _PHILOX_M0
        p1 = c2 * _PHILOX_M1
        hi0, lo0 = (p0 >> 32) & _MASK32, p0 & _MASK32
        hi1, lo1 = (p1 >> 32) & _MASK32, p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + _PHILOX_W0) & _MASK32, (k1 + _PHILOX_W1) & _MASK32
    return c0, c1, c2, c3


def _frames_philox(xp, f0, nframes, M, N, thetas_deg, d, snr_db, jitter_deg, seed, mk):
    """Shared body: `xp` is numpy or torch, `mk` builds integer / float arrays on the right device."""
    T = len(thetas_deg)
    k0, k1 = seed & _MASK32, (seed >> 32) & _MASK32
    f = mk["arange"](f0, f0 + nframes)                                   # [B] frame indices (64-bit)
    flo, fhi = f & _MASK32, (f >> 32) & _MASK32
    two32 = 1.0 / 4294967296.0

    def uniform(x):                                                       # 32-bit integer -> (0, 1), float64
        return (mk["f64"](x) + 0.5) * two32

    # frame parameters (stream 1)
    npar = (2 * T + 3) // 4
    j = mk["arange"](0, npar)
    zero = flo[:, None] * 0
    r = _philox4x32(xp, j[None, :] + zero, flo[:, None] + zero, fhi[:, None] + zero, zero + 1, k0, k1)
    u = xp.stack([uniform(v) for v in r], -1).reshape(nframes, 4 * npar)   # [B][4 npar]
    th = mk["f64c"](thetas_deg)[None, :] + (2.0 * u[:, :T] - 1.0) * jitter_deg
    ph = 2.0 * math.pi * u[:, T:2 * T]
    loc = d * 0.5 * (M - 1 - 2 * mk["f64"](mk["arange"](0, M)))
    A = xp.exp(-2j * math.pi * (xp.cos(th * (math.pi / 180.0))[..., None] * loc))          # [B][T][M] complex128
    w = math.pi / (mk["f64"](mk["arange"](0, T)) + 2.0)
    t = mk["f64"](mk["arange"](0, N))
    s = xp.exp(1j * (w[None, :, None] * t[None, None, :] + ph[:, :, None]))               # [B][T][N]
    sig = (A[:, :, :, None] * s[:, :, None, :]).sum(1)                                     # [B][M][N], fixed summation order over T
    # noise (stream 0): call j covers samples 2j, 2j + 1 of the flattened [M][N] block
    half = (M * N + 1) // 2
    jj = mk["arange"](0, half)
    zero = flo[:, None] * 0
    r0, r1, r2, r3 = _philox4x32(xp, jj[None, :] + zero, flo[:, None] + zero, fhi[:, None] + zero, zero, k0, k1)
    rad_a, ang_a = xp.sqrt(-2.0 * xp.log(uniform(r0))), 2.0 * math.pi * uniform(r1)
    rad_b, ang_b = xp.sqrt(-2.0 * xp.log(uniform(r2))), 2.0 * math.pi * uniform(r3)
    re = xp.stack([rad_a * xp.cos(ang_a), rad_b * xp.cos(ang_b)], -1).reshape(nframes, 2 * half)[:, :M * N]
    im = xp.stack([rad_a * xp.sin(ang_a), rad_b * xp.sin(ang_b)], -1).reshape(nframes, 2 * half)[:, :M * N]
    sigma = math.sqrt(10.0 ** (-snr_db / 10.0)) / math.sqrt(2.0)
    x_re = sig.real + sigma * re.reshape(nframes, M, N)
    x_im = sig.imag + sigma * im.reshape(nframes, M, N)
    return x_re, x_im, th


def frames_philox_numpy(f0, nframes, M, N, thetas_deg, d=0.5, snr_db=10.0, jitter_deg=0.0, seed=SEED_BASE):
    """Frames f0 .. f0 + nframes - 1 of the counter-based batch `seed` on the host: ([nframes][M][N] complex64, thetas [nframes][T])."""
    mk = {"arange": lambda a, b: np.arange(a, b, dtype=np.uint64), "f64": lambda v: v.astype(np.float64),
          "f64c": lambda v: np.asarray(v, np.float64)}
    x_re, x_im, th = _frames_philox(np, int(f0), int(nframes), M, N, list(thetas_deg), d, snr_db, jitter_deg, int(seed), mk)
    out = np.empty((nframes, M, N), np.complex64)
    out.real = x_re.astype(np.float32)
    out.imag = x_im.astype(np.float32)
    return out, th


def frames_philox_torch(f0, nframes, M, N, thetas_deg, d=0.5, snr_db=10.0, jitter_deg=0.0, seed=SEED_BASE, device="cuda", chunk=None):
    """The same frames generated on `device` (in chunks): what bench.py fills HBM with; a rank's shard starts at its own f0."""
    import torch
    mk = {"arange": lambda a, b: torch.arange(a, b, dtype=torch.int64, device=device), "f64": lambda v: v.to(torch.float64),
          "f64c": lambda v: torch.tensor(list(v), dtype=torch.float64, device=device)}
    if chunk is None:
        chunk = max(1, (1 << 24) // (M * N))            # ~16 M samples per chunk: a few hundred MB of float64 temporaries
    out = torch.empty((nframes, M, N), dtype=torch.complex64, device=device)
    truth = torch.empty((nframes, len(thetas_deg)), dtype=torch.float64, device=device)
    ov = torch.view_as_real(out)
    for b0 in range(0, nframes, chunk):
        nb = min(chunk, nframes - b0)
        x_re, x_im, th = _frames_philox(torch, int(f0) + b0, nb, M, N, list(thetas_deg), d, snr_db, jitter_deg, int(seed), mk)
        ov[b0:b0 + nb, :, :, 0] = x_re.to(torch.float32)
        ov[b0:b0 + nb, :, :, 1] = x_im.to(torch.float32)
        truth[b0:b0 + nb] = th
    return out, truth
