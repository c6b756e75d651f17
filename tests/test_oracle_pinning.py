"""Pins the CPU oracle (oracle/doa_oracle.cpp) against what the REFERENCE's own tests hold for this path.

The reference has no golden vectors: python/qa_*.py ask a live Octave (examples/doa_testbench_create.m) for expected
values on unseeded randn data.  Here the same cases run against tests/refmodel.py, a float64 restatement of that
Octave model, with the reference's own tolerances (and much tighter ones where the arithmetic allows).
Everything beyond these bounds is UNPINNED by the reference (oracle header).
"""
import numpy as np
import pytest

from tests import refmodel


# ---- python/qa_autocorrelate.py:40-177 ------------------------------------------------------------------------------
@pytest.mark.parametrize("len_ss,overlap,M,FB", [(2048, 512, 4, False), (1024, 256, 8, True), (256, 32, 4, True)])
def test_autocorrelate_matches_octave_model(oracle, len_ss, overlap, M, FB):
    rng = np.random.default_rng(1234 + M + len_ss)
    num_ss = 200                                     # the QA uses 1500; 200 keeps the CPU suite short
    L = (len_ss - overlap) * num_ss + overlap        # autocorrelate.m:25
    xx = (rng.standard_normal((L, M)) + 1j * rng.standard_normal((L, M))).astype(np.complex64)   # :28
    expected = refmodel.octave_autocorrelate(xx, len_ss, overlap, FB)
    observed = oracle.autocorrelate(xx.T, len_ss, overlap, int(FB))
    assert observed.shape == expected.shape
    assert np.abs(observed - expected).max() <= 1.0                       # the reference's own bound (qa_autocorrelate.py:82)
    assert np.linalg.norm(observed - expected) / np.linalg.norm(expected) < 1e-6
    if FB:   # the quirk: backward term carries an extra 1/N, so FB output is ~half the forward one, not a true average
        fwd = oracle.autocorrelate(xx.T, len_ss, overlap, 0)
        assert np.linalg.norm(observed - 0.5 * fwd) / np.linalg.norm(fwd) < 2.0 / len_ss


# ---- python/qa_MUSIC_lin_array.py:46-155 -------------------------------------------------------------------------------
@pytest.mark.parametrize("aoa,M,d", [(23.0, 8, 0.4), (121.0, 16, 0.5)])
def test_music_known_angle(oracle, aoa, M, d):
    rng = np.random.default_rng(7 + M)
    len_ss, overlap, P, num_ss = 256, 32, 1024, 100
    xx = refmodel.octave_music_input(num_ss, len_ss, overlap, M, d, [aoa], True, rng).astype(np.complex64)
    S = oracle.autocorrelate(xx.T, len_ss, overlap, 1)
    spec = oracle.music(S, d, 1, M, P)
    val, loc, idx = oracle.find_local_max(spec, 1, 0.0, 180.0)
    assert np.all(np.abs(loc[:, 0] - aoa) <= 2.0)                          # qa_MUSIC_lin_array.py:96,152
    assert np.all(np.abs(loc[:, 0] - aoa) <= 180.0 / P)                    # in fact within one bin
    assert np.all(val == 0.0)                                              # peak normalised to 0 dB (:142)
    ref = refmodel.octave_music_doa(refmodel.octave_autocorrelate(xx, len_ss, overlap, True), M, 1, d, P)
    assert np.all(np.abs(loc[:, 0] - ref) <= 180.0 / P + 1e-6)


# ---- python/qa_rootMUSIC_linear_array.py:41-145 -------------------------------------------------------------------------
@pytest.mark.parametrize("aoa,M,d,len_ss,overlap", [(23.0, 8, 0.5, 256, 32), (52.0, 4, 0.5, 1024, 64)])
def test_rootmusic_known_angle(oracle, aoa, M, d, len_ss, overlap):
    rng = np.random.default_rng(11 + M)
    num_ss = 100
    # the QA runs noiseless (snr 1000 dB): the signal root pair then sits ON the unit circle and which twin counts as
    # "strictly inside" (:125) is float32 cgeev rounding luck (at 40 dB about 1 frame in 100 has both twins at
    # |z| = 1.0000001 and the reference picks a spurious root); 20 dB keeps the case well-posed
    xx = refmodel.octave_music_input(num_ss, len_ss, overlap, M, d, [aoa], True, rng, snr_db=20.0).astype(np.complex64)
    S = oracle.autocorrelate(xx.T, len_ss, overlap, 1)
    got = oracle.rootmusic(S, d, 1, M)
    assert np.all(np.abs(got[:, 0] - aoa) <= 2.0)                          # qa_rootMUSIC_linear_array.py:87,141
    ref = refmodel.octave_rmusic(refmodel.octave_autocorrelate(xx, len_ss, overlap, True), M, 1, d)
    assert np.abs(got - ref).max() < 0.05
    twin = oracle.rootmusic_f64(S, d, 1, M)
    assert np.abs(twin - ref).max() < 1e-3


# ---- python/qa_find_local_max.py:40-110 -----------------------------------------------------------------------------------
@pytest.mark.parametrize("which,vector_len,K", [(1, 2 ** 11, 3), (2, 2 ** 12, 5)])
def test_find_local_max_octave_findpeaks(oracle, which, vector_len, K):
    data, t = refmodel.findpeaks_test_vector(which, vector_len)
    data32 = data.astype(np.float32)
    exp_pks, exp_idx = refmodel.octave_findpeaks_topk(data32, K)
    val, loc, idx = oracle.find_local_max(data32, K, 0.0, 2 * np.pi)
    np.testing.assert_allclose(val[0], exp_pks, atol=0.5e-5)              # assertAlmostEqual(..., 5)  (:74,110)
    assert np.array_equal(idx[0], exp_idx)
    # locations: the block's x-axis step is range/len (find_local_max_impl.cc:66), not Octave's linspace range/(len-1),
    # and port 1 is sorted descending by x (:188) -- the reference QA never evaluates this half of its assertion.
    xa = oracle.x_axis(vector_len, 0.0, 2 * np.pi)
    assert np.array_equal(loc[0], np.sort(xa[exp_idx])[::-1])


def test_find_local_max_rules(oracle):
    f = np.float32
    # plateau entered by a rise and left by a fall: peak = FIRST bin of the plateau (:92-107,114)
    v = np.array([0, 1, 3, 3, 3, 1, 0, 2, 0, 0], f)
    val, loc, idx = oracle.find_local_max(v, 2, 0.0, 10.0)
    assert idx[0].tolist() == [2, 7]
    # plateau followed by a further rise is not a peak; end points are never peaks; a trailing flat counts as rising
    v = np.array([5, 1, 2, 2, 3, 4, 4, 4], f)
    val, loc, idx = oracle.find_local_max(v, 2, 0.0, 8.0)
    assert idx[0].tolist() == [0, 0]          # no peak at all -> both entries = global arg-max (first occurrence) (:149-150)
    # fewer peaks than requested: fill-in uses all_pks_sorted_indx(0), an index into the PEAK LIST (reference bug, :152)
    v = np.array([0, 1, 0, 5, 0, 9, 0, 0], f)  # peaks at bins 1,3,5; best is the 3rd peak -> list position 2
    val, loc, idx = oracle.find_local_max(v, 5, 0.0, 8.0)
    assert idx[0].tolist() == [5, 3, 1, 2, 2]
    assert val[0].tolist() == [9.0, 5.0, 1.0, 0.0, 0.0]
    # K == 1 is plain index_max (find_local_max_impl.h:53-56), first occurrence
    v = np.array([1, 7, 3, 7, 0], f)
    val, loc, idx = oracle.find_local_max(v, 1, 0.0, 5.0)
    assert idx[0, 0] == 1 and val[0, 0] == 7.0
    # port 1 is sorted by x descending, decoupled from port 0's height order (:187-188)
    v = np.array([0, 3, 0, 9, 0, 5, 0], f)
    val, loc, idx = oracle.find_local_max(v, 3, 0.0, 7.0)
    assert val[0].tolist() == [9.0, 5.0, 3.0] and idx[0].tolist() == [3, 5, 1]
    assert loc[0].tolist() == [5.0, 3.0, 1.0]


def test_music_tables_follow_constructor(oracle):
    """lib/MUSIC_lin_array_impl.cc:56-86: array positions, float-accumulated theta grid, steering = exp(i*s*loc)."""
    for d, M, P in ((0.5, 8, 4096), (0.4, 4, 1000)):
        loc, th, V = oracle.music_tables(d, M, P)
        assert np.allclose(loc, d * 0.5 * (M - 1 - 2 * np.arange(M)), atol=1e-7)
        # theta is accumulated in float (:69): exact for power-of-two P, drifts by ~1e-5..1e-4 rad otherwise
        assert th[0] == 0.0 and abs(th[-1] - np.pi * (P - 1) / P) < (1e-6 if P == 4096 else 1e-4)
        psi = -2 * np.pi * np.cos(th.astype(np.float64))
        assert np.abs(V - np.exp(1j * psi[:, None] * loc[None, :].astype(np.float64))).max() < 5e-6
        assert np.abs(np.abs(V) - 1).max() < 1e-6
    # power-of-two P keeps the accumulated grid exact
    _, th, _ = oracle.music_tables(0.5, 4, 2048)
    assert np.array_equal(th, (np.pi * (np.arange(2048) * (180.0 / 2048)).astype(np.float32).astype(np.float64) / 180.0).astype(np.float32))


def test_float64_twins_bound_the_oracle_noise(oracle):
    """The reference's fp32 LAPACK path against its own float64 twin on the same covariance."""
    from gr_doa_b200 import synth
    B, M, N, T, P = 128, 8, 2048, 3, 4096
    fr, _ = synth.frames_numpy(B, M, N, [40, 90, 140], jitter_deg=5, snr_db=10, seed=21)
    R = oracle.autocorrelate_frames(fr, 0)
    q32, q64 = oracle.music_q(R, 0.5, T, M, P), oracle.music_f64(R, 0.5, T, M, P)
    assert np.abs(q32 - q64).max() < 5e-5
    a32, a64 = oracle.rootmusic(R, 0.5, T, M), oracle.rootmusic_f64(R, 0.5, T, M)
    assert np.abs(a32 - a64).max() < 0.05
