"""the antenna_correction / phase_correct_hier step in front of the path, folded into the
covariance as R' = D R D^H.  Oracle: the reference block's own arithmetic (out_k[i] = g_k * in_k[i] in complex64,
lib/antenna_correction_impl.cc:90-96) followed by the autocorrelate restatement."""
import os

import numpy as np
import pytest

from tests import parity


def reference_gains(gain, phase):
    """lib/antenna_correction_impl.cc:65-70: gr_complex(1.0/GainEst, 0) * exp(gr_complex(0, -PhaseEst)), float arithmetic."""
    gain = np.asarray(gain, np.float32); phase = np.asarray(phase, np.float32)
    a = (1.0 / gain.astype(np.float64)).astype(np.float32)
    return (a * np.cos(-phase).astype(np.float32) + 1j * (a * np.sin(-phase).astype(np.float32))).astype(np.complex64)


def test_config_file_reader_follows_the_reference_constructor(tmp_path):
    import gr_doa_b200 as doa
    from gr_doa_b200._lib import DoaCudaError
    gain = [1.0, 0.5, 2.0, 1.25]; phase = [0.0, 0.3, -1.2, 3.0]
    cfg = tmp_path / "antenna.cfg"
    cfg.write_text("".join(f"{g} {p}\n" for g, p in zip(gain, phase)))
    ac = doa.antenna_correction(4, str(cfg))
    exp = reference_gains(gain, phase)
    assert np.abs(ac.gains - exp).max() <= 2e-7 * np.abs(exp).max()
    assert not hasattr(ac, "work")        # the product never multiplies samples on the host
    # the reference's three failure modes (:59-60, :68-69, :73-74)
    with pytest.raises(DoaCudaError, match="Cannot find configuration file"):
        doa.antenna_correction(4, str(tmp_path / "missing.cfg"))
    with pytest.raises(DoaCudaError, match="too many inputs"):
        doa.antenna_correction(3, str(cfg))
    with pytest.raises(DoaCudaError, match="does not have enough inputs"):
        doa.antenna_correction(5, str(cfg))


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,overlap,avg", [(4, 256, 0, 0), (8, 512, 128, 1), (16, 192, 0, 1), (16, 255, 0, 0), (64, 256, 0, 1), (64, 129, 0, 0),
                                             (12, 100, 20, 1), (2, 64, 0, 0)])
def test_folded_gains_match_the_reference_multiply_then_autocorrelate(M, N, overlap, avg):
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    from oracle import oracle
    rng = np.random.default_rng(M * 100 + N)
    g = reference_gains(rng.uniform(0.5, 2.0, M), rng.uniform(-3.0, 3.0, M))
    n = 11
    x = synth.stream_numpy(n, M, N, overlap, [70.0, 110.0][: max(1, min(2, M - 1))], seed=M + N)
    exp = oracle.autocorrelate((g[:, None] * x).astype(np.complex64), N, overlap, avg)
    ac = doa.autocorrelate(M, N, overlap, avg, max_frames=16)
    plain = ac.work(x)
    ac.set_channel_gains(g)
    got = ac.work(x)
    assert parity.rel_fro(got, exp) <= parity.COV_REL_FRO
    Rm = got.reshape(n, M, M)
    assert np.abs(Rm - np.conj(np.transpose(Rm, (0, 2, 1)))).max() <= 4e-6 * np.abs(Rm).max()      # still Hermitian
    if avg == 0:
        assert np.abs(np.imag(np.einsum("bii->bi", Rm))).max() == 0.0                                # exactly real diagonal
    ac.set_channel_gains(None)
    assert np.array_equal(ac.work(x), plain) or parity.rel_fro(ac.work(x), plain) <= 1e-6          # gains removed again


@pytest.mark.gpu
@pytest.mark.parametrize("M,T,P,K", [(8, 3, 4096, 3), (4, 2, 1024, 2), (16, 3, 1024, 3)])
def test_chain_with_gains_equals_chain_on_corrected_samples(M, T, P, K):
    """A miscalibrated array (per-channel gain/phase errors) decoded with the correcting gains folded into the chain gives the
    peaks of the chain run on explicitly corrected samples (fused kernel at M = 8 / 4, three kernels at M = 16)."""
    import torch
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    B, N = 600, 512
    thetas = list(np.linspace(50.0, 130.0, T))
    fr, _ = synth.frames_numpy(B, M, N, thetas, snr_db=10.0, seed=7 * M)
    rng = np.random.default_rng(M)
    err = (rng.uniform(0.7, 1.4, M) * np.exp(1j * rng.uniform(-1.0, 1.0, M))).astype(np.complex64)   # what the hardware did
    g = (1.0 / err).astype(np.complex64)                                                               # what calibration found
    bad = (fr * err[None, :, None]).astype(np.complex64)
    fixed = (bad * g[None, :, None]).astype(np.complex64)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    ref = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(fixed).cuda())]
    ch.set_channel_gains(g)
    got = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(bad).cuda())]
    ch.set_channel_gains(None)
    wrong = [t.cpu().numpy() for t in ch.run_device(torch.from_numpy(bad).cuda())]
    same = (np.sort(got[2], 1) == np.sort(ref[2], 1)).all(1)
    assert same.mean() >= 0.97                                     # R differs by fp32 rounding only: near-ties may flip a bin
    assert np.abs(np.sort(got[2], 1) - np.sort(ref[2], 1)).max() <= 1
    assert np.abs(np.sort(got[1], 1) - np.sort(np.tile(thetas, (B, 1)), 1)).max() < 3.0   # and the sources are found
    assert (np.sort(wrong[2], 1) == np.sort(ref[2], 1)).all(1).mean() < 0.5                # without the gains they are not


@pytest.mark.gpu
def test_error_behaviour_of_the_new_entry_points():
    import ctypes as C
    import gr_doa_b200 as doa
    from gr_doa_b200 import _lib
    from gr_doa_b200._lib import DoaCudaError
    L = _lib.lib()
    ac = doa.autocorrelate(4, 64, 0, 0, max_frames=8)
    with pytest.raises(ValueError):
        ac.set_channel_gains(np.ones(3, np.complex64))                       # wrong length
    with pytest.raises(DoaCudaError):
        ac.set_channel_gains(np.array([1, 1, np.nan, 1], np.complex64))      # non-finite
    mu = doa.MUSIC_lin_array(0.5, 1, 4, 64, max_frames=8)
    g = np.ones(4, np.complex64)
    assert L.doa_cuda_set_channel_gains(mu._h, g.ctypes.data) != 0           # not a covariance-producing handle
    for bad in ((0.5, 1, 45.0), (0.7, 4, 45.0), (0.5, 65, 45.0), (0.5, 4, float("nan"))):
        with pytest.raises(DoaCudaError):
            doa.calibrate_lin_array(*bad)
    cal = doa.calibrate_lin_array(0.5, 4, 45.0, max_frames=4)
    assert cal.work(np.zeros((0, 16), np.complex64)).shape == (0, 4)         # empty batch
    with pytest.raises(DoaCudaError):
        cal.work(np.tile(np.eye(4, dtype=np.complex64).reshape(1, 16), (5, 1)))   # beyond max_frames
