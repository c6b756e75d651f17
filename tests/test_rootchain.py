"""doa_cuda_rootchain_*: autocorrelate -> rootMUSIC_linear_array in one call (BASELINE configs[1]: 4-element ULA, 2 sources,
snapshot 2048, forward-backward averaging).  Criterion: the bits of the two separate stages (same kernels, the covariance
just stays on the device), and through them the north_star's 1e-4 degree against the float64 twin of the reference."""
import ctypes as C

import numpy as np
import pytest

from tests import parity


def test_rootchain_create_validates_like_the_grc_checks():
    from gr_doa_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    E = _lib.EINVAL
    assert L.doa_cuda_rootchain_create(C.byref(h), 1, 2048, 512, 1, C.c_float(0.5), 1, 0, 16) == E     # inputs
    assert L.doa_cuda_rootchain_create(C.byref(h), 4, 2048, 2048, 1, C.c_float(0.5), 2, 0, 16) == E   # overlap < snapshot
    assert L.doa_cuda_rootchain_create(C.byref(h), 4, 2048, 512, 3, C.c_float(0.5), 2, 0, 16) == E    # avg_method
    assert L.doa_cuda_rootchain_create(C.byref(h), 4, 2048, 512, 1, C.c_float(0.5), 4, 0, 16) == E    # inputs > num_targets
    assert L.doa_cuda_rootchain_create(C.byref(h), 4, 2048, 512, 1, C.c_float(0.6), 2, 0, 16) == E    # norm_spacing <= 0.5
    assert not h.value


@pytest.mark.gpu
@pytest.mark.parametrize("M,T,N,overlap,avg", [(4, 2, 2048, 512, 1), (8, 3, 512, 0, 0), (16, 3, 256, 64, 1), (6, 2, 200, 50, 0)])
def test_rootchain_equals_the_two_blocks(oracle, M, T, N, overlap, avg):
    import torch
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    n = 300
    thetas = list(np.linspace(50.0, 110.0, T))
    x = synth.stream_numpy(n, M, N, overlap, thetas, seed=M + T)
    ac = doa.autocorrelate(M, N, overlap, avg, max_frames=n)
    rm = doa.rootMUSIC_linear_array(0.5, T, M, max_frames=n)
    R = ac.work(x)
    ref = rm.work(R)
    rc = doa.RootMusicChain(M, N, overlap, avg, 0.5, T, max_frames=n)
    got = rc.run_streams(list(x), n)
    assert rc.launches() == (2 if M in (4, 8) else 3)      # 4 / 8 elements: covariance + eigensolver in one persistent kernel, then the roots
    rc.set_option("fused", 0)
    assert np.array_equal(rc.run_streams(list(x), n), got) and rc.launches() == 3      # the stage kernels give the same bits
    rc.set_option("fused", 1)
    assert np.array_equal(got, ref)
    a64, d64 = oracle.rootmusic_f64(oracle.autocorrelate(x, N, overlap, avg), 0.5, T, M, return_dist=True)
    worst, near = parity.root_angles_ok(got, a64, d64)
    assert worst <= parity.ROOT_DEG and near <= 6
    assert np.abs(got - np.array(thetas)[None, :]).max() < 2.0           # the reference QA's own bound
    # independent frames: host and device entry points
    hop = N - overlap
    fr = np.stack([x[:, i * hop:i * hop + N] for i in range(64)])
    assert np.array_equal(rc.run_host(fr), ref[:64])
    assert np.array_equal(rc.run_device(torch.from_numpy(fr).cuda()).cpu().numpy(), ref[:64])
    # gains and sc16 reach this handle too
    from tests.test_sc16_input import S15, quantise, to_fc32
    q = quantise(fr)
    g = (np.linspace(0.8, 1.3, M) * np.exp(1j * np.linspace(-0.5, 0.7, M))).astype(np.complex64)
    rc.set_channel_gains(g)
    ref2 = rc.run_host(to_fc32(q, S15))
    rc.set_input_format("sc16", S15)
    assert np.array_equal(rc.run_host(q), ref2)
    from gr_doa_b200._lib import DoaCudaError
    longer = synth.stream_numpy(n + 1, M, N, overlap, thetas, seed=M + T)
    with pytest.raises(DoaCudaError):
        rc.run_streams(list(quantise(longer)), n + 1)                     # beyond max_frames
    with pytest.raises(ValueError):
        rc.run_streams(list(quantise(x)), n + 1)                          # arrays too short for n + 1 frames: refused before the library reads them
