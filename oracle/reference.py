"""ctypes front end of oracle/_ref/libdoa_ref.so: gr-doa's own, unmodified block sources compiled against the Armadillo /
GNU Radio stand-ins (oracle/build_ref.py, oracle/arma_shim/armadillo).  TEST INFRASTRUCTURE ONLY: importable from tests/,
tests/golden/make_ref_golden.py and bench.py's CPU arm, never from gr_doa_b200.

Same array conventions as oracle/oracle.py (complex64 covariances column-major M x M per frame).  Every call builds the
block through its factory and runs its work(); MUSIC_lin_array's destructor prints a line on stdout (lib/MUSIC_lin_array_impl.cc:92-95),
which `_quiet()` keeps away from callers that print JSON.
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _load_build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle_build_ref", os.path.join(_HERE, "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available() -> bool:
    """True when the library exists (prebuilt) or can be built here (the reference tree is present)."""
    mod = _load_build_module()
    return os.path.exists(mod.OUT) or mod.available()


def lib(use_herk: int = 0):
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _load_build_module().build()
    if path is None:
        raise RuntimeError("oracle/_ref/libdoa_ref.so is not built and /root/reference is not present")
    from oracle.oracle import _find_lapack
    L = C.CDLL(path)
    lp, prefix = _find_lapack()
    rc = L.ref_init(lp.encode(), prefix.encode(), int(use_herk))
    if rc != 0:
        raise RuntimeError(f"ref_init failed ({rc}) for {lp}")
    _LIB = L
    return L


def set_herk(flag: bool):
    """Sensitivity switch: route X*trans(X) through herk('U') + mirror instead of gemm('N','C') (see the shim's header)."""
    lib().ref_set_herk(int(bool(flag)))


def max_threads() -> int:
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(int(lib().ref_max_threads()), n)


@contextlib.contextmanager
def _quiet():
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def autocorrelate(streams, snapshot_size, overlap_size, avg_method, nframes=None):
    """streams [M][L] complex64 -> ([n][M*M], info) through autocorrelate::make + general_work."""
    x = _c64(streams)
    M, Lx = x.shape
    hop = snapshot_size - overlap_size
    nmax = (Lx - snapshot_size) // hop + 1 if Lx >= snapshot_size else 0
    n = nmax if nframes is None else nframes
    assert 0 < n <= nmax
    out = np.empty((n, M * M), np.complex64)
    ptrs = (C.c_void_p * M)(*[x[k].ctypes.data for k in range(M)])
    fc, hist, cons = C.c_int(0), C.c_int(0), C.c_int(0)
    r = lib().ref_autocorrelate(M, snapshot_size, overlap_size, int(avg_method), ptrs, n, _p(out), C.byref(fc), C.byref(hist), C.byref(cons))
    if r != n:
        raise RuntimeError(f"ref_autocorrelate returned {r}")
    return out, {"forecast": fc.value, "history": hist.value, "consumed": cons.value}


def autocorrelate_frames(frames, avg_method):
    x = _c64(frames)
    return np.concatenate([autocorrelate(f, x.shape[2], 0, avg_method, 1)[0] for f in x], 0)


def music(R, norm_spacing, num_targets, M, P, nthreads=1):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], P), np.float32)
    with _quiet():
        rc = lib().ref_music(C.c_float(norm_spacing), num_targets, M, P, _p(R), R.shape[0], _p(out), nthreads)
    if rc:
        raise RuntimeError(f"ref_music returned {rc}")
    return out


def rootmusic(R, norm_spacing, num_targets, M, nthreads=1, return_max_streams=False):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], num_targets), np.float32)
    ms = C.c_int(0)
    lib().ref_rootmusic(C.c_float(norm_spacing), num_targets, M, _p(R), R.shape[0], _p(out), nthreads, C.byref(ms))
    return (out, ms.value) if return_max_streams else out


def find_local_max(vecs, num_max_vals, x_min, x_max, nthreads=1):
    v = np.ascontiguousarray(vecs, np.float32)
    if v.ndim == 1:
        v = v[None, :]
    n, ln = v.shape
    val = np.empty((n, num_max_vals), np.float32)
    loc = np.empty((n, num_max_vals), np.float32)
    rc = lib().ref_find_local_max(num_max_vals, ln, C.c_float(x_min), C.c_float(x_max), _p(v), n, _p(val), _p(loc), nthreads)
    if rc:
        raise RuntimeError(f"ref_find_local_max returned {rc}")
    return val, loc


def calibrate_lin_array(R, norm_spacing, M, pilot_angle):
    R = _c64(R).reshape(-1, M * M)
    out = np.empty((R.shape[0], M), np.complex64)
    r = lib().ref_calibrate(C.c_float(norm_spacing), M, C.c_float(pilot_angle), _p(R), R.shape[0], _p(out))
    if r != R.shape[0]:
        raise RuntimeError(f"ref_calibrate returned {r}")
    return out


def chain_frames(frames, avg_method, norm_spacing, num_targets, P, K, x_min=0.0, x_max=180.0, nthreads=1, return_spectra=False):
    """autocorrelate -> MUSIC_lin_array -> find_local_max on independent frames [B][M][N]: (values, locations[, spectra])."""
    x = _c64(frames)
    B, M, N = x.shape
    val = np.empty((B, K), np.float32)
    loc = np.empty((B, K), np.float32)
    spec = np.empty((B, P), np.float32) if return_spectra else None
    with _quiet():
        rc = lib().ref_chain_frames(M, N, int(avg_method), C.c_float(norm_spacing), num_targets, P, K, C.c_float(x_min), C.c_float(x_max),
                                    _p(x), B, _p(val), _p(loc), _p(spec) if return_spectra else None, nthreads)
    if rc:
        raise RuntimeError(f"ref_chain_frames returned {rc}")
    return (val, loc, spec) if return_spectra else (val, loc)


def rootchain_frames(frames, avg_method, norm_spacing, num_targets, nthreads=1):
    x = _c64(frames)
    B, M, N = x.shape
    out = np.empty((B, num_targets), np.float32)
    lib().ref_rootchain_frames(M, N, int(avg_method), C.c_float(norm_spacing), num_targets, _p(x), B, _p(out), nthreads)
    return out


def chain_frames_procs(frames, avg_method, norm_spacing, num_targets, P, K, nprocs, root=False):
    """The same flowgraph over `nprocs` single-threaded worker PROCESSES (oracle/ref_worker.py), the way independent GNU Radio
    flowgraphs would use the cores; threads of one process serialise on OpenBLAS's buffer lock.  Returns (outputs, seconds)
    with outputs = (values, locations) -- or (angles,) for root=True -- and seconds = the slowest worker's timed pass
    (start-up, file I/O and warm-up excluded)."""
    import json
    import subprocess
    import tempfile
    x = _c64(frames)
    B = x.shape[0]
    nprocs = max(1, min(int(nprocs), B))
    root_dir = os.path.dirname(_HERE)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm) as td:
        fpath, prefix = os.path.join(td, "frames.npy"), os.path.join(td, "out")
        np.save(fpath, x)
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", PYTHONPATH=root_dir + os.pathsep + os.environ.get("PYTHONPATH", ""),
                   CUDA_VISIBLE_DEVICES="")
        procs = [subprocess.Popen([sys.executable, "-m", "oracle.ref_worker", fpath, prefix, str(w), str(nprocs), str(int(avg_method)),
                                   repr(float(norm_spacing)), str(num_targets), str(P), str(K)] + (["root"] if root else []),
                                  cwd=root_dir, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for w in range(nprocs)]
        stats = []
        for p in procs:
            so, se = p.communicate()
            if p.returncode != 0:
                raise RuntimeError(f"ref_worker failed: {se[-400:]}")
            stats.append(json.loads([l for l in so.splitlines() if l.startswith("{")][-1]))
        parts = [np.load(f"{prefix}.{w}.npy") for w in range(nprocs) if stats[w]["frames"] > 0]
    out = np.concatenate(parts, axis=0)
    seconds = max(s["seconds"] for s in stats)
    if root:
        return (out,), seconds
    return (out[:, :K].copy(), out[:, K:].copy()), seconds
