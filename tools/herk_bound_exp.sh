#!/bin/bash
# Which side bounds the tensor-core HERK?  Builds timing-only variants of herk_tc.cu (-DDOA_HERK_EXP=mask: results are WRONG by
# construction) next to the product objects and times each in its own process:  bash tools/herk_bound_exp.sh build   (CPU box)
#                                                                               gpurun -- bash tools/herk_bound_exp.sh run
cd "$(dirname "$0")/.."
mode=${1:-run}
masks="0 1 3 4 7"
if [ "$mode" = build ]; then
  mkdir -p gr_doa_b200/_ab
  for m in $masks; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDOA_HERK_EXP=$m -c gr_doa_b200/csrc/herk_tc.cu -o gr_doa_b200/_ab/herk_exp$m.o || exit 1
    objs=$(ls gr_doa_b200/_build/*.o | grep -v "\.dev\.o" | grep -v herk_tc.o)
    nvcc -shared -o gr_doa_b200/_ab/libherk_exp$m.so $objs gr_doa_b200/_ab/herk_exp$m.o -cudart static || exit 1
  done
  ls -la gr_doa_b200/_ab/
else
  for rep in 1 2; do for m in $masks; do
    DOA_AB_LIB=gr_doa_b200/_ab/libherk_exp$m.so HB=${HB:-592} python - <<PY
import os, sys
sys.path.insert(0, ".")
from gr_doa_b200 import _lib
_lib.LIB_PATH = os.path.abspath(os.environ["DOA_AB_LIB"])
exec(open("tools/herk_time.py").read().replace("mode=", "exp mask $m mode="))
PY
  done; done
fi
