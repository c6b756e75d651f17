"""Generates tests/golden/*.npz: small seeded inputs for BASELINE.json's configs together with the CPU oracle's outputs
(and its float64 twins) on them.  Run in the build container:  python tests/golden/make_golden.py
The fixtures let the GPU parity tests run against committed vectors; test_golden.py re-checks the oracle against them.
(The reference itself is C++/GNU Radio and cannot be imported or built here -- the oracle is its restatement.)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O          # noqa: E402
from gr_doa_b200 import synth           # noqa: E402


def chain_case(name, M, T, N, overlap, P, K, thetas, avg, nframes, seed, d=0.5, snr_db=10.0, stream=True):
    if stream:
        x = synth.stream_numpy(nframes, M, N, overlap, thetas, d=d, snr_db=snr_db, seed=seed)   # [M][L]
        R = O.autocorrelate(x, N, overlap, avg)
    else:
        x, _ = synth.frames_numpy(nframes, M, N, thetas, d=d, snr_db=snr_db, jitter_deg=3.0, seed=seed)   # [B][M][N]
        R = O.autocorrelate_frames(x, avg)
    spec = O.music(R, d, T, M, P)
    val, loc, bins = O.find_local_max(spec, K, 0.0, 180.0)
    out = dict(x=x, R=R, spec=spec, val=val, loc=loc, bins=bins, q32=O.music_q(R, d, T, M, P), q64=O.music_f64(R, d, T, M, P),
               aoa=O.rootmusic(R, d, T, M), aoa64=O.rootmusic_f64(R, d, T, M),
               params=np.array([M, T, N, overlap, P, K, avg, nframes, int(stream)], np.int64), d=np.float32(d),
               thetas=np.array(thetas, np.float64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def sc16_case(name, M, T, N, overlap, P, K, thetas, avg, nframes, seed, scale, d=0.5, stream=True):
    """Samples as a radio delivers them (UHD "sc16": int16 I, Q) + what the reference computes after UHD's host-side
    conversion to gr_complex, float(int16) * scale."""
    if stream:
        x = synth.stream_numpy(nframes, M, N, overlap, thetas, d=d, snr_db=10.0, seed=seed)
    else:
        x, _ = synth.frames_numpy(nframes, M, N, thetas, d=d, snr_db=10.0, jitter_deg=3.0, seed=seed)
    q = np.clip(np.rint(np.stack([x.real, x.imag], axis=-1) * 8192.0), -32768, 32767).astype(np.int16)
    f = q.astype(np.float32) * np.float32(scale)
    xc = (f[..., 0] + 1j * f[..., 1]).astype(np.complex64)
    R = O.autocorrelate(xc, N, overlap, avg) if stream else O.autocorrelate_frames(xc, avg)
    spec = O.music(R, d, T, M, P)
    val, loc, bins = O.find_local_max(spec, K, 0.0, 180.0)
    out = dict(q=q, scale=np.float32(scale), R=R, spec=spec, val=val, loc=loc, bins=bins, q32=O.music_q(R, d, T, M, P),
               q64=O.music_f64(R, d, T, M, P), aoa64=O.rootmusic_f64(R, d, T, M),
               params=np.array([M, T, N, overlap, P, K, avg, nframes, int(stream)], np.int64), d=np.float32(d),
               thetas=np.array(thetas, np.float64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def flm_case():
    rng = np.random.Generator(np.random.Philox(key=0x0D0A))
    vecs = []
    for ln in (2 ** 11,):
        t = 2 * np.pi * np.linspace(0, 1, ln)
        vecs.append(np.abs(np.sin(3.14 * t) + 0.5 * np.cos(6.09 * t) + 0.1 * np.sin(10.11 * t + 1 / 6) + 0.1 * np.sin(15.3 * t + 1 / 3)))
        vecs.append(rng.standard_normal(ln))                                   # many noise peaks
        vecs.append(np.round(rng.standard_normal(ln) * 2) / 2)                  # quantised: plateaus and exact ties
        vecs.append(np.linspace(0, 1, ln))                                      # monotone: no peak at all
        v = np.zeros(ln); v[700] = 3.0; vecs.append(v)                          # a single peak (fill-in rule)
        v = np.zeros(ln); v[100:110] = 1.0; v[900] = 2.0; vecs.append(v)        # plateau peak + spike
    vecs = np.array(vecs, np.float32)
    out = dict(vecs=vecs)
    for K in (1, 2, 3, 5, 8):
        val, loc, bins = O.find_local_max(vecs, K, 0.0, 2 * np.pi)
        out[f"val{K}"], out[f"loc{K}"], out[f"bins{K}"] = val, loc, bins
    np.savez_compressed(os.path.join(HERE, "find_local_max.npz"), **out)
    print("find_local_max", vecs.shape)


if __name__ == "__main__":
    S = synth.SEED_BASE
    # BASELINE.json configs[0]: 4-element ULA, 1 source, snapshot 2048, overlap 512, P 2048 (streaming), both averaging modes
    chain_case("cfg1_fwd", 4, 1, 2048, 512, 2048, 1, [60.0], 0, 4, S + 1)
    chain_case("cfg1_fb", 4, 1, 2048, 512, 2048, 1, [60.0], 1, 4, S + 1)
    # configs[1]: Root-MUSIC 4-element, 2 sources, fwd-bwd
    chain_case("cfg2_root", 4, 2, 2048, 512, 1024, 2, [50.0, 110.0], 1, 4, S + 2)
    # configs[2]: independent 8-element frames x 2048 snapshots, 3 sources, 4096-point scan
    chain_case("cfg3_batch", 8, 3, 2048, 0, 4096, 3, [40.0, 90.0, 140.0], 0, 3, S + 3, stream=False)
    # configs[4]: 16-element frames x 1024 snapshots
    chain_case("cfg5_m16", 16, 3, 1024, 0, 4096, 3, [40.0, 90.0, 140.0], 0, 3, S + 5, stream=False)
    flm_case()
    # sc16 wire format in front of the same path (the int16 samples are the fixture; the reference sees them converted)
    sc16_case("sc16_cfg1_fb", 4, 1, 2048, 512, 2048, 1, [60.0], 1, 4, S + 11, 1.0 / 32768)
    sc16_case("sc16_cfg3_batch", 8, 3, 2048, 0, 4096, 3, [40.0, 90.0, 140.0], 0, 3, S + 13, 1.0 / 32767, stream=False)
