"""One worker PROCESS of the reference CPU arm (TEST INFRASTRUCTURE, see oracle/reference.py).

The reference's blocks make thousands of tiny BLAS calls per frame (one cgemm + one cgemv per scan angle inside Armadillo);
OpenBLAS serialises concurrent callers of one process on its buffer lock, so threads do not scale.  A GNU Radio user who wants
all cores runs several flowgraph processes: so does this arm.  Each worker loads oracle/_ref/libdoa_ref.so, takes its slice
of the frames from a .npy file, runs warm-up + the timed pass single-threaded and prints one JSON line.

    python -m oracle.ref_worker <frames.npy> <out_prefix> <worker> <nworkers> <avg> <d> <T> <P> <K> [root]
"""
import json
import sys
import time

import numpy as np


def main():
    path, out_prefix, w, nw, avg, d, T, P, K = sys.argv[1:10]
    root = len(sys.argv) > 10 and sys.argv[10] == "root"
    w, nw, avg, T, P, K, d = int(w), int(nw), int(avg), int(T), int(P), int(K), float(d)
    from oracle import reference as REF
    REF.lib()
    fr = np.load(path, mmap_mode="r")
    lo, hi = len(fr) * w // nw, len(fr) * (w + 1) // nw
    mine = np.ascontiguousarray(fr[lo:hi])
    if len(mine) == 0:
        print(json.dumps({"worker": w, "frames": 0, "seconds": 0.0}))
        return
    run = (lambda x: (REF.rootchain_frames(x, avg, d, T, nthreads=1),)) if root else (lambda x: REF.chain_frames(x, avg, d, T, P, K, nthreads=1))
    run(mine[:1])
    t0 = time.perf_counter()
    res = run(mine)
    dt = time.perf_counter() - t0
    np.save(f"{out_prefix}.{w}.npy", np.concatenate([np.asarray(r, np.float32).reshape(len(mine), -1) for r in res], axis=1))
    print(json.dumps({"worker": w, "frames": int(len(mine)), "seconds": dt}))


if __name__ == "__main__":
    main()
