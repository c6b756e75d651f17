"""calibrate_lin_array (lib/calibrate_lin_array_impl.cc:46-134).  The reference's QA
(python/qa_calibrate_lin_array.py:78-80) perturbs every antenna by a complex gain, feeds the covariance of a pilot at a known
angle and asserts that true_perturbation / estimate is the same number for every element (to one decimal place): the
estimate is an eigenvector, i.e. defined up to a complex factor.  The oracle (the same two cheevd calls) is pinned to that
criterion; the GPU block is compared with the oracle as a direction (|<e_gpu, e_oracle>| = 1) and to the same QA criterion."""
import numpy as np
import pytest

from tests import parity

CASES = [(0.5, 45.0, 4, 1024, 128, 0), (0.5, 60.0, 8, 1024, 128, 1), (0.2, 25.0, 4, 256, 64, 0)]     # the QA's three configurations


def pilot_covariances(oracle, d, pilot, M, N, overlap, avg, n=60, snr=30.0, seed=0):
    from gr_doa_b200 import synth
    rng = np.random.default_rng(seed + M)
    pert = (rng.uniform(0.5, 1.5, M) * np.exp(1j * rng.uniform(-np.pi, np.pi, M))).astype(np.complex64)
    x = synth.stream_numpy(n, M, N, overlap, [pilot], d=d, snr_db=snr, seed=seed + 3)
    R = oracle.autocorrelate((pert[:, None] * x).astype(np.complex64), N, overlap, avg)
    return R, pert


def qa_spread(pert, est):
    """max over frames of |diff(pert / est)| relative to |pert / est| (0 when est is proportional to pert)."""
    ratio = pert[None, :] / est
    return float((np.abs(np.diff(ratio, axis=1)).max(1) / np.abs(ratio).mean(1)).max())


@pytest.mark.parametrize("d,pilot,M,N,overlap,avg", CASES)
def test_oracle_meets_the_reference_qa_criterion(oracle, d, pilot, M, N, overlap, avg):
    R, pert = pilot_covariances(oracle, d, pilot, M, N, overlap, avg)
    est = oracle.calibrate_lin_array(R, d, M, pilot)
    assert np.abs(np.linalg.norm(est, axis=1) - 1.0).max() <= 1e-5
    assert qa_spread(pert, est) <= 0.05           # assertComplexTuplesAlmostEqual(..., places=1)


@pytest.mark.gpu
@pytest.mark.parametrize("d,pilot,M,N,overlap,avg", CASES + [(0.5, 100.0, 16, 512, 0, 1), (0.4, 75.0, 6, 300, 30, 0), (0.5, 30.0, 64, 512, 0, 0)])
def test_gpu_block_matches_the_oracle_direction(oracle, d, pilot, M, N, overlap, avg):
    import gr_doa_b200 as doa
    R, pert = pilot_covariances(oracle, d, pilot, M, N, overlap, avg, n=40)
    est_o = oracle.calibrate_lin_array(R, d, M, pilot, nthreads=4)
    cal = doa.calibrate_lin_array(d, M, pilot, max_frames=64)
    est = cal.work(R)
    assert est.shape == est_o.shape == (R.shape[0], M)
    assert np.abs(np.linalg.norm(est, axis=1) - 1.0).max() <= 1e-5
    cosang = np.abs(np.einsum("bi,bi->b", np.conj(est_o), est))
    assert cosang.min() >= 1.0 - 1e-5                       # same vector up to the unit-modulus factor LAPACK leaves open
    assert qa_spread(pert, est) <= 0.05
    assert np.abs(est - cal.work(R)).max() == 0.0           # deterministic
    v = oracle.calibrate_pilot_vector(d, M, pilot)
    assert np.abs(np.abs(v) - 1.0).max() <= 1e-6
