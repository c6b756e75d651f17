"""Parity pinned by the reference's own code.

oracle/_ref/libdoa_ref.so is gr-doa's unmodified lib/{autocorrelate,MUSIC_lin_array,rootMUSIC_linear_array,find_local_max,
calibrate_lin_array}_impl.cc compiled against stand-ins for Armadillo and GNU Radio (oracle/build_ref.py).  tests/golden/ref_*.npz
are its outputs on seeded inputs (tests/golden/make_ref_golden.py); they travel to the GPU box, the reference tree does not.

  CPU (-m "not gpu"):  the port (oracle/doa_oracle.cpp) against the fixtures, and -- where the reference build exists -- against
                       the build itself on fresh inputs; the fixtures against the known angles.
  GPU (-m gpu):        every stage of libdoa_cuda and the fused chains against the fixtures, with tests/parity.py's criteria.
"""
import glob
import os

import numpy as np
import pytest

from tests import parity
from tests.conftest import has_cuda

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHAIN_CASES = ["ref_cfg1_fb", "ref_cfg1_fwd", "ref_cfg2_root", "ref_cfg3_batch", "ref_cfg4_m64", "ref_cfg5_m16", "ref_odd"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    M, T, N, overlap, P, K, avg, nframes, stream, seed = [int(v) for v in z["params"]]
    p = dict(M=M, T=T, N=N, overlap=overlap, P=P, K=K, avg=avg, nframes=nframes, stream=bool(stream), seed=seed, d=float(z["d"]),
             thetas=[float(t) for t in z["thetas"]], snr_db=float(z["snr_db"]))
    return z, p


def inputs_of(z, p):
    """The fixture's input samples: stored, or (cfg4: 8 MB per frame) regenerated from the recorded seed."""
    if "x" in z.files:
        return z["x"]
    from gr_doa_b200 import synth
    x, _ = synth.frames_numpy(p["nframes"], p["M"], p["N"], p["thetas"], d=p["d"], snr_db=p["snr_db"], jitter_deg=3.0, seed=p["seed"])
    return x


def peak_bins(loc, P):
    """x-axis locations (k * 180 / P accumulated in float, lib/find_local_max_impl.cc:60-69) back to bins."""
    return np.rint(np.asarray(loc, np.float64) * P / 180.0).astype(np.int64)


def test_fixture_set_is_complete():
    have = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz")))
    assert have == sorted(CHAIN_CASES + ["ref_find_local_max", "ref_calibrate"])


@pytest.mark.parametrize("name", CHAIN_CASES)
def test_reference_fixtures_find_the_sources(name):
    z, p = load(name)
    T = p["T"]
    # MUSIC + find_local_max: the T highest peaks sit on the true angles (grid 180/P, 3 degrees of jitter on independent frames)
    tol = 4.0 if not p["stream"] else 1.0
    # (port 1 is sorted descending by x, decoupled from the heights, lib/find_local_max_impl.cc:188: compare as sets)
    if p["K"] == T:
        assert np.abs(np.sort(z["loc"], axis=1) - np.sort(np.asarray(p["thetas"]))[None, :]).max() < tol
    # Root-MUSIC: angles ascending, on the truth
    assert np.abs(z["aoa"] - np.sort(np.asarray(p["thetas"]))[None, :]).max() < tol
    # scheduler contract: forecast = hop * noutput, history = overlap + 1, consume_each(hop * noutput), 1..T output ports
    hop = p["N"] - p["overlap"]
    fc, hist, cons, max_streams = [int(v) for v in z["sched"]]
    assert hist == p["overlap"] + 1 and max_streams == T
    if p["stream"]:
        assert fc == hop * p["nframes"] and cons == hop * p["nframes"]


@pytest.mark.parametrize("name", CHAIN_CASES)
def test_port_matches_reference_fixtures(oracle, name):
    """The restatement against what the reference's own sources computed.  Covariance, Root-MUSIC and find_local_max are the
    same arithmetic statement by statement (bit-equal on the machine that wrote the fixtures; rounding-level tolerances here
    because BLAS/LAPACK kernels are CPU-dispatched); the MUSIC scan differs in the last bits of Q (the reference's per-angle
    products run through cgemm/cgemv inside Armadillo, the port's through plain loops in the same order)."""
    z, p = load(name)
    M, T, P, K, d = p["M"], p["T"], p["P"], p["K"], p["d"]
    x = inputs_of(z, p)
    R = oracle.autocorrelate(x, p["N"], p["overlap"], p["avg"]) if p["stream"] else oracle.autocorrelate_frames(x, p["avg"], nthreads=oracle.max_threads())
    assert parity.rel_fro(R, z["R"]) < 1e-6
    spec = oracle.music(z["R"], d, T, M, P, nthreads=oracle.max_threads())
    assert parity.spectrum_db_error(spec, z["spec"], z["q64"]) < parity.SPECTRUM_DB
    # find_local_max is comparisons and copies: bit-exact on the reference's own spectra
    val, loc, _ = oracle.find_local_max(z["spec"], K, 0.0, 180.0)
    assert np.array_equal(val, z["val"]) and np.array_equal(loc, z["loc"])
    # peaks of the port's own spectra: identical bins except near-ties
    _, loc_p, bins_p = oracle.find_local_max(spec, K, 0.0, 180.0)
    ndiff, unexplained = parity.classify_bins(bins_p, peak_bins(z["loc"], P), z["q64"], z["q32"])
    assert not unexplained and ndiff <= max(1, p["nframes"] // 8)
    # Root-MUSIC: both are float32 cgeev; each within its own noise of the float64 twin, and of each other on good frames
    aoa = oracle.rootmusic(z["R"], d, T, M)
    good = np.nanmin(z["dist64"], axis=1) >= parity.ROOT_NEAR_CIRCLE
    if good.any():
        assert np.abs(aoa[good] - z["aoa"][good]).max() < 5e-2
        assert np.abs(z["aoa"][good] - z["aoa64"][good]).max() < 5e-2
    else:   # the large array: every selected root within 4e-4 of the circle; the angles still sit on the truth
        assert np.abs(aoa - z["aoa"]).max() < 0.5


def test_port_matches_reference_find_local_max_vectors(oracle):
    z = np.load(os.path.join(GOLDEN, "ref_find_local_max.npz"))
    for K in (1, 2, 3, 4, 8):
        val, loc, _ = oracle.find_local_max(z["vecs"], K, 0.0, float(2 * np.pi))
        # heights always; locations whenever the K + 1 highest peaks have distinct heights (equal heights are ordered by an
        # unstable std::sort in the reference, lib/find_local_max_impl.cc:137)
        assert np.array_equal(val, z[f"val{K}"])
        distinct = np.array([len(set(np.round(v, 12))) == len(v) for v in z[f"val{K}"]])
        assert np.array_equal(loc[distinct], z[f"loc{K}"][distinct])


def test_port_matches_reference_calibrate(oracle):
    z = np.load(os.path.join(GOLDEN, "ref_calibrate.npz"))
    M = int(z["params"][0])
    est = oracle.calibrate_lin_array(z["R"], float(z["d"]), M, float(z["pilot"]))
    # an eigenvector: defined up to a unit-modulus factor
    ph = np.sum(est * np.conj(z["est"]), axis=1, keepdims=True)
    assert np.abs(est - z["est"] * ph / np.abs(ph)).max() < 1e-4
    # and proportional to the injected gains (the reference QA's criterion, python/qa_calibrate_lin_array.py)
    g = z["gains"][None, :]
    ratio = z["est"] / g
    assert np.abs(ratio / ratio[:, :1] - 1.0).max() < 0.1


def _ref_or_skip():
    from oracle import reference as REF
    if not REF.available():
        pytest.skip("reference tree and prebuilt oracle/_ref are both absent")
    REF.lib()
    return REF


def test_reference_build_compiles_the_unmodified_sources():
    """The build recipe names the reference's files where they lie and no copy of them exists in this repository."""
    import importlib.util
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("build_ref", os.path.join(here, "oracle", "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    srcs = mod.reference_sources()
    assert len(srcs) == 5 and all(s.startswith(mod.REF + os.sep + "lib" + os.sep) and s.endswith("_impl.cc") for s in srcs)
    for s in srcs:
        assert not os.path.exists(os.path.join(here, "oracle", os.path.basename(s)))


def test_port_against_live_reference_build(oracle):
    """Fresh seeded inputs through both: covariance (incl. the forward-backward 1/N quirk), Root-MUSIC and find_local_max
    bit-equal; spectra within 1e-3 dB outside deep nulls; peaks identical except near-ties."""
    REF = _ref_or_skip()
    from gr_doa_b200 import synth
    nthr = oracle.max_threads()
    stats = []
    for (M, N, ov, avg, T, P, K, th) in [(4, 2048, 512, 1, 1, 2048, 1, [60.0]), (4, 2048, 512, 1, 2, 1024, 2, [50.0, 110.0]),
                                         (8, 2048, 0, 0, 3, 4096, 3, [40.0, 90.0, 140.0]), (16, 1024, 0, 0, 3, 4096, 3, [40.0, 90.0, 140.0]),
                                         (5, 300, 37, 1, 2, 777, 4, [70.0, 120.0])]:
        nfr = 48
        x = synth.stream_numpy(nfr, M, N, ov, th, seed=4242 + M)
        Rr, info = REF.autocorrelate(x, N, ov, avg)
        Ro = oracle.autocorrelate(x, N, ov, avg)
        assert info == {"forecast": (N - ov) * nfr, "history": ov + 1, "consumed": (N - ov) * nfr}
        assert np.array_equal(Rr.view(np.uint32), Ro.view(np.uint32))
        assert np.array_equal(REF.rootmusic(Ro, 0.5, T, M).view(np.uint32), oracle.rootmusic(Ro, 0.5, T, M).view(np.uint32))
        Sr, So = REF.music(Ro, 0.5, T, M, P, nthreads=nthr), oracle.music(Ro, 0.5, T, M, P, nthreads=nthr)
        q64, q32 = oracle.music_f64(Ro, 0.5, T, M, P, nthreads=nthr), oracle.music_q(Ro, 0.5, T, M, P, nthreads=nthr)
        assert parity.spectrum_db_error(So, Sr, q64) < parity.SPECTRUM_DB
        vr, lr = REF.find_local_max(Sr, K, 0.0, 180.0)
        vo, lo, bo = oracle.find_local_max(Sr, K, 0.0, 180.0)
        assert np.array_equal(vr, vo) and np.array_equal(lr, lo)
        _, _, bins_port = oracle.find_local_max(So, K, 0.0, 180.0)
        ndiff, unexplained = parity.classify_bins(bins_port, bo, q64, q32)
        assert not unexplained
        stats.append((M, ndiff, nfr))
    assert sum(s[1] for s in stats) <= 6, stats


def test_reference_rootmusic_exhausted_slots_are_90_degrees(oracle):
    """Fewer than T roots strictly inside the unit circle: the consumed root is overwritten with (inf, 0), index_min of an all-inf
    vector is 0, arg(inf + 0j) = 0 and acos(0) = 90 degrees (lib/rootMUSIC_linear_array_impl.cc:131-141).  Only an EMPTY inside set
    is undefined (Armadillo's index_min throws)."""
    REF = _ref_or_skip()
    M, T = 4, 2
    # a rank-one covariance: one source at high SNR leaves a single root pair near the circle and conjugate-reciprocal pairs elsewhere
    a = np.exp(-2j * np.pi * 0.5 * np.cos(np.deg2rad(70.0)) * (np.arange(M) - (M - 1) / 2))
    R = (np.outer(a, a.conj()) + 1e-3 * np.eye(M)).astype(np.complex64).T.reshape(1, -1)
    ar, ao = REF.rootmusic(R, 0.5, T, M), oracle.rootmusic(R, 0.5, T, M)
    assert np.array_equal(ar.view(np.uint32), ao.view(np.uint32))


# ------------------------------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("name", CHAIN_CASES)
def test_gpu_stages_match_reference_fixtures(doa, oracle, name):
    import torch
    z, p = load(name)
    M, T, N, P, K, d = p["M"], p["T"], p["N"], p["P"], p["K"], p["d"]
    x = inputs_of(z, p)
    n = p["nframes"]
    # stage 1
    ac = doa.autocorrelate(M, N, p["overlap"], p["avg"], max_frames=n)
    R = ac.general_work(n, [x[k] for k in range(M)])[0] if p["stream"] else ac.work_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert parity.rel_fro(R, z["R"]) < parity.COV_REL_FRO
    # stage 2 on the reference's covariances
    mus = doa.MUSIC_lin_array(d, T, M, P, max_frames=n)
    spec = mus.work(z["R"])
    assert parity.spectrum_db_error(spec, z["spec"], z["q64"]) < parity.SPECTRUM_DB
    # stage 4 on the reference's spectra: bit-exact heights, locations where heights are distinct
    flm = doa.find_local_max(K, P, 0.0, 180.0, max_frames=n)
    val, loc = flm.work(z["spec"])
    assert np.array_equal(val, z["val"])
    distinct = np.array([len(set(v)) == len(v) for v in z["val"]])
    assert np.array_equal(loc[distinct], z["loc"][distinct])
    # stage 3 on the reference's covariances: 1e-4 degree against the float64 twin on well-conditioned frames, and no
    # further from the reference's float32 answer than that answer is from the twin
    rm = doa.rootMUSIC_linear_array(d, T, M, max_frames=n)
    aoa = rm.work(z["R"])
    worst, near = parity.root_angles_ok(aoa, z["aoa64"], z["dist64"])
    assert worst <= parity.ROOT_DEG and (near <= max(1, n // 4) or M >= 32)
    good = np.nanmin(z["dist64"], axis=1) >= parity.ROOT_NEAR_CIRCLE
    if good.any():
        assert np.abs(aoa[good] - z["aoa"][good]).max() <= np.abs(z["aoa"][good] - z["aoa64"][good]).max() + parity.ROOT_DEG
    else:
        assert np.abs(aoa - z["aoa"]).max() < 0.5


@gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("name", CHAIN_CASES)
def test_gpu_chain_matches_reference_fixtures(doa, name):
    """The fused autocorrelate -> MUSIC -> find_local_max call against the reference's peaks: bins identical except near-ties,
    heights within the bound the reference's own float32 noise allows."""
    z, p = load(name)
    M, T, N, P, K, d = p["M"], p["T"], p["N"], p["P"], p["K"], p["d"]
    x = inputs_of(z, p)
    n = p["nframes"]
    ch = doa.DoaChain(M, N, p["overlap"], p["avg"], d, T, P, K, max_frames=n)
    val, loc, bins = ch.run_streams([x[k] for k in range(M)], n) if p["stream"] else ch.run_host(x)
    ref_bins_by_loc = peak_bins(z["loc"], P)                       # port 1: descending by x
    ndiff, unexplained = parity.classify_bins(bins, ref_bins_by_loc, z["q64"], z["q32"])
    assert not unexplained, (name, unexplained)
    assert ndiff <= max(1, n // 6)
    same = (np.sort(bins, axis=1) == np.sort(ref_bins_by_loc, axis=1)).all(axis=1)
    assert np.array_equal(loc[same], z["loc"][same])
    # heights (port 0, descending): compare entry by entry on frames whose bins agree
    if K <= T:
        bound = parity.peak_value_bound_db(z["q64"], np.sort(ref_bins_by_loc, axis=1), z["q32"]).max(axis=1)
        assert (np.abs(val[same] - z["val"][same]).max(axis=1) <= bound[same]).all()


@gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_gpu_rootchain_and_calibrate_match_reference_fixtures(doa):
    z, p = load("ref_cfg2_root")
    x = z["x"]
    rc = doa.RootMusicChain(p["M"], p["N"], p["overlap"], p["avg"], p["d"], p["T"], max_frames=p["nframes"])
    aoa = rc.run_streams([x[k] for k in range(p["M"])], p["nframes"])
    worst, near = parity.root_angles_ok(aoa, z["aoa64"], z["dist64"])
    assert worst <= parity.ROOT_DEG and near <= 4
    zc = np.load(os.path.join(GOLDEN, "ref_calibrate.npz"))
    M = int(zc["params"][0])
    cal = doa.calibrate_lin_array(float(zc["d"]), M, float(zc["pilot"]), max_frames=zc["R"].shape[0])
    est = cal.work(zc["R"])
    ph = np.sum(est * np.conj(zc["est"]), axis=1, keepdims=True)
    assert np.abs(est - zc["est"] * ph / np.abs(ph)).max() < 1e-4
