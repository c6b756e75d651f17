"""ctypes front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see doa_oracle.cpp header): importable
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from the
product package gr_doa_b200.

All arrays are numpy; complex64 matrices are column-major M x M per frame, laid out exactly as the
reference blocks put them on their ports (lib/autocorrelate_impl.cc:103, lib/MUSIC_lin_array_impl.cc:124).
"""
import ctypes as C
import glob
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _find_lapack():
    """OpenBLAS (with LAPACK) bundled in this image: scipy's (scipy_ prefix) first, opencv's as a fallback."""
    import scipy
    sp = os.path.dirname(os.path.dirname(scipy.__file__))
    for pat, prefix in ((os.path.join(sp, "scipy.libs", "libscipy_openblas-*.so"), "scipy_"),
                        (os.path.join(sp, "opencv_python_headless.libs", "libopenblasp-*.so"), "")):
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0], prefix
    raise RuntimeError("no LAPACK-bearing OpenBLAS found in site-packages")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "_build", "libdoa_oracle.so")
    if not os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("oracle_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    L = C.CDLL(path)
    lp, prefix = _find_lapack()
    rc = L.oracle_init(lp.encode(), prefix.encode())
    if rc != 0:
        raise RuntimeError(f"oracle_init failed ({rc}) for {lp}")
    L.lapack_path = lp
    _LIB = L
    return L


def max_threads() -> int:
    """Host threads this process may use: the CPU affinity mask (torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would otherwise turn the 'all host cores' baseline into a single-thread one); every entry point takes the count
    explicitly (`num_threads` clause), so the environment variable does not cap it."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(int(lib().oracle_max_threads()), n)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def autocorrelate(streams, snapshot_size, overlap_size, avg_method, nframes=None, nthreads=1):
    """streams: [M][L] complex64 (M channel streams).  Returns [n][M*M] complex64 (col-major per frame).
    Mirrors autocorrelate_impl::general_work (lib/autocorrelate_impl.cc:82-118)."""
    x = _c64(streams)
    M, Lx = x.shape
    hop = snapshot_size - overlap_size
    nmax = (Lx - snapshot_size) // hop + 1 if Lx >= snapshot_size else 0
    n = nmax if nframes is None else nframes
    assert n <= nmax
    out = np.empty((n, M * M), np.complex64)
    ptrs = (C.c_void_p * M)(*[x[k].ctypes.data for k in range(M)])
    lib().oracle_autocorrelate(ptrs, M, snapshot_size, overlap_size, int(avg_method), n, _p(out), nthreads)
    return out


def autocorrelate_frames(frames, avg_method, nthreads=1):
    """frames: [B][M][N] complex64 independent frames -> [B][M*M]."""
    x = _c64(frames)
    B, M, N = x.shape
    out = np.empty((B, M * M), np.complex64)
    lib().oracle_autocorrelate_frames(_p(x), M, N, int(avg_method), B, _p(out), nthreads)
    return out


def music_tables(norm_spacing, M, P):
    loc = np.empty(M, np.float32)
    th = np.empty(P, np.float32)
    V = np.empty((P, M), np.complex64)   # V[ii, nn] == d_vii_matrix(nn, ii)
    lib().oracle_music_tables(C.c_float(norm_spacing), M, P, _p(loc), _p(th), _p(V))
    return loc, th, V


def music(R, norm_spacing, num_targets, M, P, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], P), np.float32)
    lib().oracle_music(_p(R), R.shape[0], C.c_float(norm_spacing), num_targets, M, P, _p(out), nthreads)
    return out


def music_q(R, norm_spacing, num_targets, M, P, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], P), np.float32)
    lib().oracle_music_q(_p(R), R.shape[0], C.c_float(norm_spacing), num_targets, M, P, _p(out), nthreads)
    return out


def music_f64(R, norm_spacing, num_targets, M, P, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], P), np.float64)
    lib().oracle_music_f64(_p(R), R.shape[0], C.c_float(norm_spacing), num_targets, M, P, _p(out, C.c_double), nthreads)
    return out


def noise_projector(R, num_targets, M, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    G = np.empty_like(R)
    w = np.empty((R.shape[0], M), np.float32)
    lib().oracle_noise_projector(_p(R), R.shape[0], num_targets, M, _p(G), _p(w), nthreads)
    return G, w


def noise_projector_f64(R, num_targets, M, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    G = np.empty(R.shape, np.complex128)
    w = np.empty((R.shape[0], M), np.float64)
    lib().oracle_noise_projector_f64(_p(R), R.shape[0], num_targets, M, _p(G, C.c_double), _p(w, C.c_double), nthreads)
    return G, w


def calibrate_lin_array(R, norm_spacing, M, pilot_angle, nthreads=1):
    """lib/calibrate_lin_array_impl.cc:112-126: [n][M*M] covariances -> [n][M] complex64 gain/phase estimates (the
    eigenvector LAPACK returns, i.e. defined up to a unit-modulus factor)."""
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], M), np.complex64)
    lib().oracle_calibrate_lin_array(_p(R), R.shape[0], C.c_float(norm_spacing), M, C.c_float(pilot_angle), _p(out), nthreads)
    return out


def calibrate_pilot_vector(norm_spacing, M, pilot_angle):
    v = np.empty(M, np.complex64)
    lib().oracle_calibrate_pilot_vector(C.c_float(norm_spacing), M, C.c_float(pilot_angle), _p(v))
    return v


def rootmusic(R, norm_spacing, num_targets, M, nthreads=1, return_roots=False):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], num_targets), np.float32)
    roots = np.empty((R.shape[0], 2 * M - 2), np.complex64)
    lib().oracle_rootmusic(_p(R), R.shape[0], C.c_float(norm_spacing), num_targets, M, _p(out), _p(roots), nthreads)
    return (out, roots) if return_roots else out


def rootmusic_f64(R, norm_spacing, num_targets, M, nthreads=1, return_dist=False):
    """float64 twin of Root-MUSIC.  return_dist: also 1-|z| of the selected roots (the frame's conditioning)."""
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], num_targets), np.float64)
    dist = np.full((R.shape[0], num_targets), np.nan, np.float64)
    lib().oracle_rootmusic_f64(_p(R), R.shape[0], C.c_float(norm_spacing), num_targets, M, _p(out, C.c_double),
                               _p(dist, C.c_double), nthreads)
    return (out, dist) if return_dist else out


def x_axis(length, x_min, x_max):
    x = np.empty(length, np.float32)
    lib().oracle_x_axis(length, C.c_float(x_min), C.c_float(x_max), _p(x))
    return x


def find_local_max(vecs, num_max_vals, x_min, x_max, nthreads=1):
    """vecs [n][len] float32 -> (values [n][K] desc by height, locations [n][K] desc by x, bins [n][K] in value order)."""
    v = np.ascontiguousarray(vecs, np.float32)
    if v.ndim == 1:
        v = v[None, :]
    n, ln = v.shape
    K = num_max_vals
    val = np.empty((n, K), np.float32)
    loc = np.empty((n, K), np.float32)
    idx = np.empty((n, K), np.int32)
    lib().oracle_find_local_max(_p(v), n, K, ln, C.c_float(x_min), C.c_float(x_max), _p(val), _p(loc), _p(idx, C.c_int), nthreads)
    return val, loc, idx


def chain_frames(frames, avg_method, norm_spacing, num_targets, P, K, nthreads=1):
    x = _c64(frames)
    B, M, N = x.shape
    val = np.empty((B, K), np.float32)
    loc = np.empty((B, K), np.float32)
    idx = np.empty((B, K), np.int32)
    lib().oracle_chain_frames(_p(x), B, M, N, int(avg_method), C.c_float(norm_spacing), num_targets, P, K,
                              _p(val), _p(loc), _p(idx, C.c_int), nthreads)
    return val, loc, idx
