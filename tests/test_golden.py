"""CPU: the oracle against the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).
LAPACK kernels are CPU-dispatched, so floats are compared to rounding-level tolerances and bins exactly."""
import glob
import os

import numpy as np
import pytest

from tests import parity

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cfg*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    M, T, N, overlap, P, K, avg, nframes, stream = [int(v) for v in z["params"]]
    return z, dict(M=M, T=T, N=N, overlap=overlap, P=P, K=K, avg=avg, nframes=nframes, stream=bool(stream), d=float(z["d"]))


def test_fixture_set_is_complete():
    assert CASES == ["cfg1_fb", "cfg1_fwd", "cfg2_root", "cfg3_batch", "cfg5_m16"]
    assert os.path.exists(os.path.join(GOLDEN, "find_local_max.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    z, p = load(name)
    R = oracle.autocorrelate(z["x"], p["N"], p["overlap"], p["avg"]) if p["stream"] else oracle.autocorrelate_frames(z["x"], p["avg"])
    assert parity.rel_fro(R, z["R"]) < 1e-6
    spec = oracle.music(z["R"], p["d"], p["T"], p["M"], p["P"])
    assert parity.spectrum_db_error(spec, z["spec"], z["q64"]) < 1e-4
    val, loc, bins = oracle.find_local_max(z["spec"], p["K"], 0.0, 180.0)
    assert np.array_equal(bins, z["bins"]) and np.array_equal(val, z["val"]) and np.array_equal(loc, z["loc"])
    assert np.abs(oracle.rootmusic_f64(z["R"], p["d"], p["T"], p["M"]) - z["aoa64"]).max() < 1e-6
    # the truth is recovered (grid 180/P): the fixtures are sane, not just self-consistent
    assert np.abs(np.sort(z["loc"], axis=1) - np.sort(z["thetas"])[None, :]).max() < 4.0


def test_find_local_max_golden(oracle):
    z = np.load(os.path.join(GOLDEN, "find_local_max.npz"))
    for K in (1, 2, 3, 5, 8):
        val, loc, bins = oracle.find_local_max(z["vecs"], K, 0.0, 2 * np.pi)
        assert np.array_equal(val, z[f"val{K}"]) and np.array_equal(loc, z[f"loc{K}"]) and np.array_equal(bins, z[f"bins{K}"])
