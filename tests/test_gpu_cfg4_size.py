"""GPU (-m gpu): BASELINE.json configs[3] at the batch SURVEY section 8(d) names -- 512 frames of 64 elements x 16,384 snapshots,
8 sources, 16,384-point scan, K = 8 (4.3 GB of samples) -- through size-independent properties and an oracle spot check on frames
of the very same device buffer.  512 = 3 rounds of the 148 persistent HERK CTAs + 68 frames that are shared between CTAs
(herk_tc.cu: split tail), so frame independence here also says that a shared frame has the bits of a whole one."""
import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu

B, M, N, T, P, K = 512, 64, 16384, 8, 16384, 8
THETAS = [30.0 + 120.0 * i / 7 for i in range(8)]


@pytest.fixture(scope="module")
def full(doa):
    import torch
    from gr_doa_b200 import synth
    x, truth = synth.frames_torch(B, M, N, THETAS, jitter_deg=2.0, snr_db=10.0, device="cuda", chunk=32)
    ch = doa.DoaChain(M, N, 0, 0, 0.5, T, P, K, max_frames=B)
    val, loc, bins = ch.run_device(x)
    torch.cuda.synchronize()
    yield dict(x=x, truth=truth, ch=ch, val=val, loc=loc, bins=bins, torch=torch)
    del x
    torch.cuda.empty_cache()


def test_every_frame_finds_its_eight_sources(full):
    loc = np.sort(full["loc"].cpu().numpy(), 1)
    truth = np.sort(full["truth"].cpu().numpy(), 1)
    assert np.abs(loc - truth).max() < 0.1                                      # 64 elements, 16 384 snapshots: sharp nulls
    val = full["val"].cpu().numpy()
    assert np.all(val[:, 0] == 0.0) and np.all(np.diff(val, axis=1) <= 0.0)
    assert np.all(np.diff(full["loc"].cpu().numpy(), axis=1) <= 0.0)


def test_frames_are_independent_of_batch_and_position(full):
    """Sub-batches whose frames fall into other rounds of the persistent kernels, are shared between CTAs or not, or arrive
    permuted give every frame the same bits."""
    torch = full["torch"]
    for lo, hi in ((0, 37), (100, 248), (444, 512), (300, 301)):
        v, l, b = full["ch"].run_device(full["x"][lo:hi])
        assert torch.equal(b, full["bins"][lo:hi]) and torch.equal(v, full["val"][lo:hi]) and torch.equal(l, full["loc"][lo:hi])
    perm = torch.randperm(200, device="cuda")
    v, l, b = full["ch"].run_device(full["x"][:200][perm].contiguous())
    assert torch.equal(b, full["bins"][:200][perm]) and torch.equal(v, full["val"][:200][perm])


def test_covariance_at_size_is_split_invariant(full, doa):
    torch = full["torch"]
    ac = doa.autocorrelate(M, N, 0, 0, max_frames=B)
    R1 = ac.work_device(full["x"])
    ac.set_option("herk_split", 0)
    R0 = ac.work_device(full["x"])
    assert torch.equal(torch.view_as_real(R0), torch.view_as_real(R1))


def test_oracle_spot_check_on_the_resident_buffer(full, oracle, doa):
    pick = np.array([0, 147, 443, 444, 479, 511])                               # whole-frame rounds and shared tail frames
    fr = full["x"][full["torch"].from_numpy(pick).cuda()].cpu().numpy()
    nt = oracle.max_threads()
    R = oracle.autocorrelate_frames(fr, 0, nthreads=nt)
    spec = oracle.music(R, 0.5, T, M, P, nthreads=nt)
    q32, q64 = oracle.music_q(R, 0.5, T, M, P, nthreads=nt), oracle.music_f64(R, 0.5, T, M, P, nthreads=nt)
    val_o, loc_o, bins_o = oracle.find_local_max(spec, K, 0.0, 180.0, nthreads=nt)
    ndiff, unexplained = parity.classify_bins(full["bins"].cpu().numpy()[pick], bins_o, q64, q32)
    assert unexplained == []
    # the chain's covariance is the block's: check it through the standalone block on the same frames
    got = doa.autocorrelate(M, N, 0, 0, max_frames=len(pick)).work_device(full["torch"].from_numpy(fr).cuda()).cpu().numpy()
    assert parity.rel_fro(got, R) <= parity.COV_REL_FRO
