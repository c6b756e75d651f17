"""Recipe for oracle/_ref/libdoa_ref.so: the reference's OWN block sources, unmodified, compiled where they lie.

    python oracle/build_ref.py [--force]

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (present in the build container, absent on the GPU box, which uses the
prebuilt .so: oracle/_ref/ is git-ignored but travels with the gpurun snapshot).  The reference's own build system is not
run (it needs cmake-found GNU Radio, Boost, Armadillo >= 7.300, CppUnit, SWIG and the un-vendored OPINCAA library); instead
g++ compiles the five block sources directly against
  * /root/reference/include            the reference's public headers (doa/*.h), as they are
  * oracle/arma_shim/armadillo         a stand-in for Armadillo (see its header for what it restates)
  * gr_doa_b200/gnuradio/shim          the compile-only stand-in for gnuradio/{block,sync_block,io_signature}.h and boost::shared_ptr
and oracle/ref_harness.cpp (C entry points that call make() + work()).
Flags follow the reference's CMake (Release: -O3; -std=c++11, CMakeLists.txt:28-44; no -march, so no FMA contraction).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DOA_REFERENCE_DIR", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libdoa_ref.so")
BLOCKS = ["autocorrelate", "MUSIC_lin_array", "rootMUSIC_linear_array", "find_local_max", "calibrate_lin_array"]


def reference_sources():
    return [os.path.join(REF, "lib", b + "_impl.cc") for b in BLOCKS]


def available() -> bool:
    return all(os.path.exists(p) for p in reference_sources())


def build(force: bool = False):
    """Returns the path of the library, or None when the reference tree is not there and no prebuilt library exists."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = reference_sources() + [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "arma_shim", "armadillo"), __file__]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++11", "-O3", "-DNDEBUG", "-fopenmp", "-fPIC", "-shared", "-Dgnuradio_doa_EXPORTS",
           "-I", os.path.join(HERE, "arma_shim"), "-I", os.path.join(ROOT, "gr_doa_b200", "gnuradio", "shim"),
           "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "lib"),
           "-o", OUT, os.path.join(HERE, "ref_harness.cpp")] + reference_sources() + ["-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
