"""Where does a stage of the tensor-core HERK's operand ring spend its time?  A -DDOA_HERK_TRACE=1 build records clock64 in CTA 0 at
the ring's hand-offs (stage full seen by the MMA warp / MMAs issued / converter before the empty wait / empty seen / full arrived).
    python tools/herk_trace.py build          (CPU box: variant library under gr_doa_b200/_ab/)
    gpurun -- python tools/herk_trace.py      (prints medians over the steady state of the first frame)"""
import ctypes, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
AB = os.path.join(ROOT, "gr_doa_b200", "_ab")
LIB = os.path.join(AB, "libherk_trace.so")

if len(sys.argv) > 1 and sys.argv[1] == "build":
    os.makedirs(AB, exist_ok=True)
    extra = sys.argv[2:]
    obj = os.path.join(AB, "herk_trace.o")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-DDOA_HERK_TRACE=1"] + extra +
                          ["-c", os.path.join(ROOT, "gr_doa_b200", "csrc", "herk_tc.cu"), "-o", obj])
    objs = [o for o in glob.glob(os.path.join(ROOT, "gr_doa_b200", "_build", "*.o")) if not o.endswith(".dev.o") and not o.endswith("herk_tc.o")]
    subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs + [obj, "-cudart", "static"])
    print(LIB)
    sys.exit(0)

import numpy as np
import torch
from gr_doa_b200 import _lib, synth
_lib.LIB_PATH = LIB
import gr_doa_b200 as doa
B, N = 148, 16384
x, _ = synth.frames_torch(B, 64, N, [30.0 + 120.0 * i / 7 for i in range(8)], jitter_deg=2.0, device="cuda", chunk=32)
ac = doa.autocorrelate(64, N, 0, 0, max_frames=B)
for _ in range(3):
    ac.work_device(x)
torch.cuda.synchronize()
tr = np.zeros((5, 1024), np.int64)
L = ctypes.CDLL(LIB)
rc = L.doa_herk_trace_dump(tr.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
full, issued, pre, empty, arrive = tr
q = np.arange(64, 1000)
med = lambda v: (int(np.median(v)), int(np.percentile(v, 10)), int(np.percentile(v, 90)))
print("clk (median, p10, p90) over stages 64..999 of CTA 0's first frame:")
print(" stage period (full arrived, q -> q+1)          ", med(np.diff(arrive[64:1000])))
print(" MMA warp sees full after the arrival            ", med(full[q] - arrive[q]))
print(" MMA warp: full seen -> 8 MMAs + commit issued   ", med(issued[q] - full[q]))
print(" MMA warp: issued(q) -> full seen(q+1)           ", med(full[q + 1] - issued[q]))
print(" converter: wait at empty                        ", med(empty[q] - pre[q]))
waited = (empty[q] - pre[q]) > 100
print(" ... fraction of stages where it waited > 100 clk", round(float(waited.mean()), 3))
qq = q[waited & (q >= 68)]
print(" issued(q-4) -> empty seen(q), when it waited    ", med(empty[qq] - issued[qq - 4]))
print(" full arrived(q-4) -> empty seen(q), when waited ", med(empty[qq] - arrive[qq - 4]))
print(" converter: empty seen -> full arrived (convert) ", med(arrive[q] - empty[q]))
print(" converter group cycle (pre(q) -> pre(q+3))      ", med(pre[q + 3] - pre[q]))
print(" converter: full arrived(q) -> before empty(q+3) ", med(pre[q + 3] - arrive[q]))
