"""Compile the four GNU Radio block sources against the compile-only shim (shim/) and link the fake scheduler with
libdoa_cuda.so.  GNU Radio, Boost and SWIG are not in this image; with a real GNU Radio 3.7 the same lib/*.cc build inside
gr-doa's own CMake."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build", "fake_scheduler")


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(HERE, "lib", f) for f in ("autocorrelate_impl.cc", "MUSIC_lin_array_impl.cc",
                                                   "rootMUSIC_linear_array_impl.cc", "find_local_max_impl.cc", "music_chain_impl.cc", "rootmusic_chain_impl.cc", "calibrate_lin_array_impl.cc")]
    srcs.append(os.path.join(HERE, "harness", "fake_scheduler.cc"))
    deps = srcs + [os.path.join(PKG, "libdoa_cuda.so")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return OUT
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Dgnuradio_doa_EXPORTS", "-I", os.path.join(HERE, "shim"),
           "-I", os.path.join(HERE, "include"), "-I", os.path.join(PKG, "..", "include"), "-o", OUT] + srcs + \
          ["-L", PKG, "-ldoa_cuda", "-Wl,-rpath," + PKG]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
