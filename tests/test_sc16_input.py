"""sc16 (UHD int16 I/Q) samples read directly by the covariance kernels.

The reference only ever sees fc32: UHD converts the radio's int16 pairs on the host (python/twinrx_usrp_source.py:57,
cpu_format "fc32") before autocorrelate reads them.  The criterion is therefore: the sc16 path gives what the fc32 path
gives on the converted samples float(int16) * scale -- bit for bit when the scale is a power of two (same accumulation
order, exact conversion, the scale commutes with every rounding), within 2 ulp per entry otherwise -- and that, like the
fc32 path, stays inside the north_star's covariance tolerance against the oracle's autocorrelate on those samples."""
import numpy as np
import pytest

from tests import parity

S15 = 1.0 / 32768


def quantise(x, full_scale=0.25):
    """complex64 [...] -> int16 [..., 2] the way an ADC would deliver it (|x| ~ 1 at a quarter of full scale)."""
    q = np.stack([x.real, x.imag], axis=-1) * (32768 * full_scale)
    return np.clip(np.rint(q), -32768, 32767).astype(np.int16)


def to_fc32(q, scale):
    """What UHD's sc16 -> fc32 converter hands to the reference: float(int16) * scale, in float."""
    f = q.astype(np.float32) * np.float32(scale)
    return (f[..., 0] + 1j * f[..., 1]).astype(np.complex64)


def test_conversion_trick_is_exact_for_every_int16():
    """Host restatement of sc16_to_c64 (csrc/cov_device.cuh): bias to unsigned, splice under the exponent of 2^23, subtract
    2^23 + 32768.  Exact for all 65,536 values, extremes included."""
    i16 = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    w = i16.view(np.uint16).astype(np.uint32)
    spliced = ((w ^ 0x8000) | 0x4B000000).astype(np.uint32)
    f = spliced.view(np.float32) - np.float32(8421376.0)
    assert np.array_equal(f, i16.astype(np.float32))


def test_mirror_rejects_the_wrong_dtype_without_touching_the_device():
    import gr_doa_b200 as doa
    blk = doa.blocks._Block.__new__(doa.blocks._Block)
    blk._sc16 = True
    with pytest.raises(ValueError):
        blk._samples(np.zeros((4, 64), np.complex64))
    with pytest.raises(ValueError):
        blk._samples(np.zeros((4, 64, 3), np.int16))
    assert blk._nsamp(blk._samples(np.zeros((64, 2), np.int16))) == 64


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,overlap,avg", [(4, 2048, 512, 1), (8, 512, 128, 0), (8, 255, 0, 1), (4, 100, 33, 0), (2, 64, 0, 0),
                                             (16, 192, 0, 1), (16, 255, 7, 0), (12, 100, 20, 1), (64, 129, 0, 0)])
def test_sc16_covariance_equals_fc32_on_converted_samples(M, N, overlap, avg):
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    from oracle import oracle
    n = 13
    x = synth.stream_numpy(n, M, N, overlap, [70.0, 110.0][: max(1, min(2, M - 1))], seed=3 * M + N)
    q = quantise(x)
    q[0, 0] = (-32768, 32767)          # the extremes go through the converter too
    q[-1, -1] = (32767, -32768)
    ac = doa.autocorrelate(M, N, overlap, avg, max_frames=16)
    if M == 64:
        ac.set_option("herk_tc", 0)   # the tensor-core HERK is an fc32-only path with its own rounding
    try:
        ref = ac.work(to_fc32(q, S15))
        ac.set_input_format("sc16", S15)
        got = ac.work(q)
        assert got.shape == ref.shape == (n, M * M)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))                       # bit for bit
        exp = oracle.autocorrelate(to_fc32(q, S15), N, overlap, avg)
        assert parity.rel_fro(got, exp) <= parity.COV_REL_FRO
        # UHD's own factor (1/32767, not a power of two): the scale enters once, squared
        s = 1.0 / 32767
        ac.set_input_format("sc16", s)
        got2 = ac.work(q)
        ac.set_input_format("fc32")
        ref2 = ac.work(to_fc32(q, s))
        assert parity.rel_fro(got2, ref2) <= 1e-6
        assert parity.rel_fro(got2, oracle.autocorrelate(to_fc32(q, s), N, overlap, avg)) <= parity.COV_REL_FRO
        assert np.array_equal(ac.work(to_fc32(q, S15)), ref)                                    # and fc32 is back
    finally:
        ac.close()


@pytest.mark.gpu
@pytest.mark.parametrize("M,T,P,K,N", [(8, 3, 4096, 3, 2048), (8, 3, 1024, 3, 200), (4, 1, 2048, 1, 2048), (4, 2, 1024, 2, 512),
                                       (16, 3, 1024, 3, 256)])
def test_chain_on_sc16_equals_chain_on_converted_samples(M, T, P, K, N):
    """Fused kernel at M = 8 / 4 (8-byte cp.async ring slots), three kernels at M = 16; device, host-frames and
    host-streams entry points; with channel gains on top."""
    import torch
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    B = 700
    thetas = list(np.linspace(50.0, 130.0, T))
    fr, _ = synth.frames_numpy(B, M, N, thetas, snr_db=10.0, seed=11 * M + T)
    q = quantise(fr)
    fc = to_fc32(q, S15)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    ref = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(fc).cuda())]
    ref_host = ch.run_host(fc)
    g = (np.linspace(0.8, 1.3, M) * np.exp(1j * np.linspace(-0.5, 0.7, M))).astype(np.complex64)
    ch.set_channel_gains(g)
    ref_g = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(fc).cuda())]
    ch.set_input_format("sc16", S15)
    got_g = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(q).cuda())]
    ch.set_channel_gains(None)
    got = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(q).cuda())]
    got_host = ch.run_host(q)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)
    for a, b in zip(got_g, ref_g):
        assert np.array_equal(a, b)
    for a, b in zip(got_host, ref_host):
        assert np.array_equal(a, b)
    assert np.abs(np.sort(got[1], 1) - np.sort(np.tile(thetas, (B, 1)), 1)).max() < 3.0        # and the sources are found
    # streaming form (hop / overlap framing on the device) through the GNU Radio-facing entry point
    n, ov = 9, N // 4
    xs = quantise(synth.stream_numpy(n, M, N, ov, thetas, seed=5 * M))
    chs = doa.DoaChain(M, N, ov, 1, 0.5, T, P, K, max_frames=16)
    ref_s = chs.run_streams(list(to_fc32(xs, S15)), n)
    chs.set_input_format("sc16", S15)
    got_s = chs.run_streams(list(xs), n)
    for a, b in zip(got_s, ref_s):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_full_size_batch_on_sc16_equals_fc32(doa):
    """BASELINE configs[2] size (65,536 x 8 x 2048): every frame's peaks from the int16 samples are those from the converted
    samples, bit for bit; generated and compared on the device."""
    import torch
    from gr_doa_b200 import synth
    B, M, N, T, P, K = 65536, 8, 2048, 3, 4096, 3
    x, truth = synth.frames_torch(B, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, snr_db=10.0, device="cuda")
    q = torch.view_as_real(x).mul(8192.0).round_().clamp_(-32768, 32767).to(torch.int16)
    del x
    fc = torch.view_as_complex(q.to(torch.float32).mul_(S15))
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    ref = ch.run_device(fc)
    ch.set_input_format("sc16", S15)
    got = ch.run_device(q)
    torch.cuda.synchronize()
    assert ch.launches() == 1
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    err = (got[1].sort(1).values.double() - truth.sort(1).values).abs().max(1).values
    assert (err <= 1.0).double().mean() >= 0.999


@pytest.mark.gpu
def test_error_behaviour_of_set_input_format():
    import gr_doa_b200 as doa
    from gr_doa_b200 import _lib
    from gr_doa_b200._lib import DoaCudaError
    L = _lib.lib()
    ac = doa.autocorrelate(4, 64, 0, 0, max_frames=8)
    with pytest.raises(ValueError):
        ac.set_input_format("sc8")
    for bad in (0.0, -1.0, float("nan"), float("inf")):
        with pytest.raises(DoaCudaError):
            ac.set_input_format("sc16", bad)
    assert L.doa_cuda_set_input_format(ac._h, 7, 1.0) != 0
    mu = doa.MUSIC_lin_array(0.5, 1, 4, 64, max_frames=8)
    assert L.doa_cuda_set_input_format(mu._h, 1, 1.0) != 0                  # not a covariance-producing handle
    ac.set_input_format("sc16")
    with pytest.raises(ValueError):
        ac.work(np.zeros((4, 64), np.complex64))                            # complex samples on an sc16 handle
    assert ac.work(np.zeros((4, 10, 2), np.int16)).shape == (0, 16)         # shorter than one snapshot: no frames
    out = ac.work(np.full((4, 64, 2), 16384, np.int16))                     # x = 0.5 + 0.5j everywhere: R = 0.5
    assert np.array_equal(out, np.full((1, 16), 0.5 + 0j, np.complex64))


@pytest.mark.gpu
@pytest.mark.parametrize("M,T,P,K,N", [(8, 3, 1024, 3, 200), (8, 2, 512, 2, 2048), (4, 1, 512, 1, 136), (4, 2, 1024, 2, 1024)])
def test_channel_major_ring_fills_change_nothing(M, T, P, K, N):
    """Dev knob ws_fill = 2 (a producer lane fills ONE channel with immediates on a single address instead of one slot of every
    channel): same ring contents, hence the same bits, for both sample formats, ragged last chunk (N % 64 != 0) included."""
    import torch
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    B = 900
    fr, _ = synth.frames_numpy(B, M, N, list(np.linspace(50.0, 130.0, T)), snr_db=10.0, seed=M + N)
    q = quantise(fr)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)          # the shipped configuration
    with doa.dev_library():                                          # the variant lives in the -DDOA_DEV_KNOBS build
        dch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    dch.set_option("ws_fill", 2)
    for fmt, x in (("fc32", torch.from_numpy(to_fc32(q, S15)).cuda()), ("sc16", torch.from_numpy(q).cuda())):
        ch.set_input_format(fmt, S15)
        dch.set_input_format(fmt, S15)
        ref = [t.clone() for t in ch.run_device(x)]
        got = dch.run_device(x)
        assert dch.launches() == 1
        assert all(torch.equal(a, b) for a, b in zip(got, ref)), fmt


# ---- committed golden fixtures (tests/golden/sc16_*.npz, written by tests/golden/make_golden.py) -----------------------
SC16_CASES = ["sc16_cfg1_fb", "sc16_cfg3_batch"]


def _load_sc16(name):
    import os
    from tests.test_golden import GOLDEN
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    M, T, N, overlap, P, K, avg, nframes, stream = [int(v) for v in z["params"]]
    return z, dict(M=M, T=T, N=N, overlap=overlap, P=P, K=K, avg=avg, nframes=nframes, stream=bool(stream), d=float(z["d"]),
                   scale=float(z["scale"]))


@pytest.mark.parametrize("name", SC16_CASES)
def test_oracle_reproduces_sc16_golden(oracle, name):
    """CPU: the oracle on float(int16) * scale (what UHD's converter hands the reference) against the committed fixture."""
    z, p = _load_sc16(name)
    xc = to_fc32(z["q"], p["scale"])
    R = oracle.autocorrelate(xc, p["N"], p["overlap"], p["avg"]) if p["stream"] else oracle.autocorrelate_frames(xc, p["avg"])
    assert parity.rel_fro(R, z["R"]) < 1e-6
    val, loc, bins = oracle.find_local_max(oracle.music(z["R"], p["d"], p["T"], p["M"], p["P"]), p["K"], 0.0, 180.0)
    assert np.array_equal(bins, z["bins"])
    assert np.abs(np.sort(z["loc"], axis=1) - np.sort(z["thetas"])[None, :]).max() < 4.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", SC16_CASES)
def test_gpu_sc16_against_golden(name):
    """GPU: int16 samples in, the fixture's covariance (1e-5), peak bins (near-ties classified) and Root-MUSIC angles (1e-4
    degree against the float64 twin) out -- through autocorrelate, the fused chain and the Root-MUSIC chain."""
    import gr_doa_b200 as doa
    z, p = _load_sc16(name)
    M, T, N, ov, P, K, avg, n = p["M"], p["T"], p["N"], p["overlap"], p["P"], p["K"], p["avg"], p["nframes"]
    ac = doa.autocorrelate(M, N, ov, avg, max_frames=n)
    ac.set_input_format("sc16", p["scale"])
    ch = doa.DoaChain(M, N, ov, avg, p["d"], T, P, K, max_frames=n)
    ch.set_input_format("sc16", p["scale"])
    rc = doa.RootMusicChain(M, N, ov, avg, p["d"], T, max_frames=n)
    rc.set_input_format("sc16", p["scale"])
    if p["stream"]:
        R = ac.work(z["q"])
        val, loc, bins = ch.run_streams(list(z["q"]), n)
        aoa = rc.run_streams(list(z["q"]), n)
    else:
        R = ac.work(np.ascontiguousarray(z["q"].transpose(1, 0, 2, 3)).reshape(M, n * N, 2))
        val, loc, bins = ch.run_host(z["q"])
        aoa = rc.run_host(z["q"])
    assert parity.rel_fro(R, z["R"]) <= parity.COV_REL_FRO
    ndiff, unexplained = parity.classify_bins(bins, z["bins"], z["q64"], z["q32"])
    assert unexplained == []
    assert np.abs(aoa - z["aoa64"]).max() <= parity.ROOT_DEG
