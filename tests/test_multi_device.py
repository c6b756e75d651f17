"""doa_cuda_multi_*: the fused chain over several GPUs from one process.  Frames are independent
(lib/autocorrelate_impl.cc:92, lib/MUSIC_lin_array_impl.cc:121, lib/find_local_max_impl.cc:179), so the criterion is exact:
every frame's peaks equal what one device returns for it, whatever the device list and batch size."""
import ctypes as C

import numpy as np
import pytest


def test_multi_create_validates_arguments_without_a_device():
    from gr_doa_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    dev = (C.c_int * 2)(0, 0)
    args = (4, 2048, 512, 0, C.c_float(0.5), 1, 1024, 1, C.c_float(0), C.c_float(180))
    assert L.doa_cuda_multi_create(C.byref(h), *args, None, 1, 16) == _lib.EINVAL
    assert L.doa_cuda_multi_create(C.byref(h), *args, dev, 0, 16) == _lib.EINVAL
    assert L.doa_cuda_multi_create(C.byref(h), *args, dev, 65, 16) == _lib.EINVAL
    assert not h.value
    assert L.doa_cuda_multi_device_count(None) == _lib.EINVAL
    if L.doa_cuda_device_count() == 0:           # no GPU here: fails loudly, no fallback
        assert L.doa_cuda_multi_create(C.byref(h), *args, dev, 2, 16) == _lib.ECUDA
        assert b"no CUDA device" in L.doa_cuda_last_error(None)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0], "all"])
def test_multi_equals_single_device_chain(devices):
    import torch
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    if devices == "all":
        devices = list(range(torch.cuda.device_count()))
    M, N, T, P, K, B = 8, 512, 3, 2048, 3, 1001
    fr, _ = synth.frames_numpy(B, M, N, [40.0, 90.0, 140.0], jitter_deg=5.0, snr_db=10.0, seed=len(devices))
    ref = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B).run_host(fr)
    mc = doa.DoaChainMulti(M, N, 0, 0, 0.5, T, P, K, devices=devices, max_frames_per_device=512)
    G = len(devices)
    for nb in (1, G - 1, G, G + 1, 700, B):
        if nb < 1 or nb > 512 * G:
            continue
        blocks = mc.blocks(nb)
        assert sum(c for _, c in blocks) == nb and blocks[0][0] == 0
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))          # contiguous, in order
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1                            # balanced
        got = mc.run_host(fr[:nb])
        for a, b in zip(got, ref):
            assert np.array_equal(a, b[:nb]), (nb, devices)
    if 512 * G < B:
        from gr_doa_b200._lib import DoaCudaError
        with pytest.raises(DoaCudaError):
            mc.run_host(fr)                                                                          # beyond the capacity
    # gains and the sample format reach every device
    from tests.test_sc16_input import S15, quantise, to_fc32
    nb = min(B, 512 * G)
    q = quantise(fr[:nb])
    g = (np.linspace(0.8, 1.3, M) * np.exp(1j * np.linspace(-0.5, 0.7, M))).astype(np.complex64)
    one = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=nb)
    one.set_channel_gains(g)
    ref2 = one.run_host(to_fc32(q, S15))
    mc.set_channel_gains(g)
    mc.set_input_format("sc16", S15)
    got2 = mc.run_host(q)
    for a, b in zip(got2, ref2):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_multi_error_paths():
    import gr_doa_b200 as doa
    from gr_doa_b200._lib import DoaCudaError
    with pytest.raises(DoaCudaError):
        doa.DoaChainMulti(8, 512, 0, 0, 0.5, 3, 2048, 3, devices=[0, 999])       # a device that does not exist
    with pytest.raises(DoaCudaError):
        doa.DoaChainMulti(8, 512, 0, 0, 0.5, 8, 2048, 3, devices=[0])            # targets must be < elements
    mc = doa.DoaChainMulti(4, 64, 0, 0, 0.5, 1, 256, 1, devices=[0, 0], max_frames_per_device=4)
    v, l, b = mc.run_host(np.zeros((0, 4, 64), np.complex64))
    assert v.shape == (0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,overlap,avg,sc16", [(4, 2048, 512, 1, False), (8, 256, 32, 0, True)])
def test_multi_streams_equal_single_device(M, N, overlap, avg, sc16):
    """The streaming (hop / overlap) form: every device reads its block of frames, halo included, from the same host streams."""
    import gr_doa_b200 as doa
    from gr_doa_b200 import synth
    from tests.test_sc16_input import S15, quantise, to_fc32
    T, P, K, n = (1, 2048, 1, 97) if M == 4 else (2, 1024, 2, 97)
    x = synth.stream_numpy(n, M, N, overlap, [60.0] if T == 1 else [50.0, 110.0], seed=M)
    one = doa.DoaChain(M, N, overlap, avg, 0.5, T, P, K, max_frames=128)
    mc = doa.DoaChainMulti(M, N, overlap, avg, 0.5, T, P, K, devices=[0, 0, 0], max_frames_per_device=64)
    if sc16:
        x = quantise(x)
        one.set_input_format("sc16", S15)
        mc.set_input_format("sc16", S15)
    for nb in (1, 2, 3, 50, n):
        ref = one.run_streams(list(x), nb)
        got = mc.run_streams(list(x), nb)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b), nb
