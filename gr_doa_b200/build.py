"""Build libdoa_cuda.so in-tree with nvcc for sm_100a (no torch, no cmake: a handful of .cu files, one shared library).

    python -m gr_doa_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  cudart is linked statically so the
library has no run-time dependency beyond the driver.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libdoa_cuda.so")
SOURCES = ["doa_cuda.cu", "cov.cu", "eig.cu", "eig_block.cu", "scan.cu", "root.cu", "fused.cu", "herk_tc.cu", "scan_tc.cu", "fused16.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "doa_cuda.h")]
    return max(os.path.getmtime(p) for p in paths)


DEV_LIB = os.path.join(HERE, "libdoa_cuda_dev.so")


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    """dev=False: libdoa_cuda.so, the product (shipped kernel configurations only).  dev=True: libdoa_cuda_dev.so, the same
    sources with -DDOA_DEV_KNOBS: every experimental kernel variant and the options that select them (tools/, bit-identity tests)."""
    lib_path = DEV_LIB if dev else LIB
    flags = NVCC_FLAGS + (["-DDOA_DEV_KNOBS"] if dev else [])
    os.makedirs(BUILD, exist_ok=True)
    if not force and os.path.exists(lib_path) and os.path.getmtime(lib_path) >= _deps_mtime():
        return lib_path
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".dev.o" if dev else ".o"))
        srcp = os.path.join(CSRC, src)
        hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
        hdr_m = max([os.path.getmtime(h) for h in hdrs] + [os.path.getmtime(os.path.join(HERE, "..", "include", "doa_cuda.h"))])
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(srcp), hdr_m):
            return obj, ""
        r = subprocess.run([nvcc] + flags + ["-c", srcp, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    r = subprocess.run([nvcc, "-shared", "-o", lib_path] + objs + ["-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev="--dev" in sys.argv))
    if "--all" in sys.argv:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev=True))
