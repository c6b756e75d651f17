"""Generates tests/golden/ref_*.npz: seeded inputs for BASELINE.json's configs together with what gr-doa's OWN, unmodified
block sources compute on them (oracle/_ref/libdoa_ref.so: /root/reference/lib/*_impl.cc compiled against the Armadillo /
GNU Radio stand-ins, oracle/build_ref.py).  Run in the build container, where /root/reference exists:

    python tests/golden/make_ref_golden.py

The GPU box has no /root/reference: there the CUDA path and the port (oracle/doa_oracle.cpp) are checked against these
committed vectors (tests/test_reference_build.py).  The float64 twins (q64, aoa64, dist64) come from the port and only serve
the near-tie / near-circle classification of tests/parity.py.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O          # noqa: E402
from oracle import reference as REF     # noqa: E402
from gr_doa_b200 import synth           # noqa: E402


def chain_case(name, M, T, N, overlap, P, K, thetas, avg, nframes, seed, d=0.5, snr_db=10.0, stream=True, keep_input=True):
    if stream:
        x = synth.stream_numpy(nframes, M, N, overlap, thetas, d=d, snr_db=snr_db, seed=seed)   # [M][L]
        R, info = REF.autocorrelate(x, N, overlap, avg)
    else:
        x, _ = synth.frames_numpy(nframes, M, N, thetas, d=d, snr_db=snr_db, jitter_deg=3.0, seed=seed)   # [B][M][N]
        R = REF.autocorrelate_frames(x, avg)
        info = {"forecast": N, "history": 1, "consumed": N}
    spec = REF.music(R, d, T, M, P)
    val, loc = REF.find_local_max(spec, K, 0.0, 180.0)
    aoa, max_streams = REF.rootmusic(R, d, T, M, return_max_streams=True)
    aoa64, dist64 = O.rootmusic_f64(R, d, T, M, return_dist=True)
    out = dict(R=R, spec=spec, val=val, loc=loc, aoa=aoa, aoa64=aoa64, dist64=dist64, q64=O.music_f64(R, d, T, M, P),
               q32=O.music_q(R, d, T, M, P), params=np.array([M, T, N, overlap, P, K, avg, nframes, int(stream), seed], np.int64),
               d=np.float32(d), snr_db=np.float64(snr_db), thetas=np.array(thetas, np.float64),
               sched=np.array([info["forecast"], info["history"], info["consumed"], max_streams], np.int64))
    if keep_input:
        out["x"] = x
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def flm_case():
    """find_local_max on its own: the reference QA's test vector shape (python/qa_find_local_max.py:52-60), noise, plateaus,
    monotone and constant vectors, fewer peaks than requested (the fill-in rule and its index bug)."""
    rng = np.random.Generator(np.random.Philox(key=0x0D0A + 77))
    ln = 2 ** 10
    t = 2 * np.pi * np.linspace(0, 1, ln)
    vecs = [np.abs(np.sin(3.14 * t) + 0.5 * np.cos(6.09 * t) + 0.1 * np.sin(10.11 * t + 1 / 6) + 0.1 * np.sin(15.3 * t + 1 / 3)),
            rng.standard_normal(ln), np.round(rng.standard_normal(ln) * 2) / 2, np.linspace(0, 1, ln), np.linspace(1, 0, ln),
            np.ones(ln), np.concatenate([np.linspace(0, 1, ln // 2), np.ones(ln // 2)]), -np.abs(t - 3.0),
            np.repeat(rng.standard_normal(ln // 4), 4), np.where((np.arange(ln) // 7) % 2 == 0, 1.0, 0.0)]
    for _ in range(22):
        v = rng.standard_normal(ln).cumsum()
        vecs.append(np.round(v * 4) / 4 if _ % 2 else v)
    vecs = np.asarray(vecs, np.float32)
    out = dict(vecs=vecs)
    for K in (1, 2, 3, 4, 8):
        val, loc = REF.find_local_max(vecs, K, 0.0, float(2 * np.pi))
        out[f"val{K}"] = val
        out[f"loc{K}"] = loc
    np.savez_compressed(os.path.join(HERE, "ref_find_local_max.npz"), **out)
    print("ref_find_local_max", vecs.shape)


def calibrate_case():
    M, d, pilot = 4, 0.5, 45.0
    x, _ = synth.frames_numpy(24, M, 2048, [pilot], d=d, snr_db=20.0, seed=0x0D0A + 99)
    gains = np.array([1.0, 0.8 * np.exp(0.3j), 1.2 * np.exp(-0.5j), 0.9 * np.exp(1.1j)], np.complex64)
    R = REF.autocorrelate_frames((x * gains[None, :, None]).astype(np.complex64), 0)
    out = dict(R=R, est=REF.calibrate_lin_array(R, d, M, pilot), gains=gains, params=np.array([M], np.int64), d=np.float32(d), pilot=np.float32(pilot))
    np.savez_compressed(os.path.join(HERE, "ref_calibrate.npz"), **out)
    print("ref_calibrate", out["est"].shape)


if __name__ == "__main__":
    s = synth.SEED_BASE + 100
    # configs[0]: run_MUSIC_lin_array_simulation shape, both averaging methods (streaming, overlap 512)
    chain_case("ref_cfg1_fwd", 4, 1, 2048, 512, 2048, 1, [60.0], 0, 24, s + 1)
    chain_case("ref_cfg1_fb", 4, 1, 2048, 512, 2048, 1, [60.0], 1, 24, s + 1)
    # configs[1]: Root-MUSIC, 2 sources, forward-backward
    chain_case("ref_cfg2_root", 4, 2, 2048, 512, 1024, 2, [50.0, 110.0], 1, 24, s + 2)
    # configs[2]: independent 8-element frames
    chain_case("ref_cfg3_batch", 8, 3, 2048, 0, 4096, 3, [40.0, 90.0, 140.0], 0, 12, s + 3, stream=False)
    # configs[4]: 16-element frames x 1024 snapshots
    chain_case("ref_cfg5_m16", 16, 3, 1024, 0, 4096, 3, [40.0, 90.0, 140.0], 0, 12, s + 5, stream=False)
    # configs[3]: the large array (inputs are regenerated from the seed: 8 MB per frame)
    chain_case("ref_cfg4_m64", 64, 8, 16384, 0, 16384, 8, [30.0 + 120.0 * i / 7 for i in range(8)], 0, 2, s + 4, stream=False, keep_input=False)
    # an odd shape: nothing a power of two, 4 peaks asked of 2 sources (fill-in rule)
    chain_case("ref_odd", 6, 2, 500, 100, 1000, 4, [50.0, 110.0], 1, 16, s + 6)
    flm_case()
    calibrate_case()
