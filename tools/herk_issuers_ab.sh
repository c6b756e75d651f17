#!/bin/bash
# One against two MMA-issuing warps in the tensor-core HERK (-DDOA_HERK_ISSUERS; herk_tc.cu), each library in its own process.
#   bash tools/herk_issuers_ab.sh build     (CPU box: gr_doa_b200/_ab/libherk_iss{1,2}.so next to the product objects)
#   gpurun -- bash tools/herk_issuers_ab.sh
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  mkdir -p gr_doa_b200/_ab
  objs=$(ls gr_doa_b200/_build/*.o | grep -v "\.dev\.o" | grep -v herk_tc.o)
  for n in 1 2; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DDOA_HERK_ISSUERS=$n -c gr_doa_b200/csrc/herk_tc.cu -o gr_doa_b200/_ab/herk_iss$n.o || exit 1
    nvcc -shared -o gr_doa_b200/_ab/libherk_iss$n.so $objs gr_doa_b200/_ab/herk_iss$n.o -cudart static || exit 1
  done
  exit 0
fi
for rep in 1 2; do for lib in gr_doa_b200/_ab/libherk_iss2.so gr_doa_b200/_ab/libherk_iss1.so; do for hb in 592 512; do
DOA_AB_LIB=$lib HB=$hb python - <<PY
import os, sys
sys.path.insert(0, ".")
from gr_doa_b200 import _lib
_lib.LIB_PATH = os.path.abspath(os.environ["DOA_AB_LIB"])
exec(open("tools/herk_time.py").read().replace("mode=", os.path.basename(os.environ["DOA_AB_LIB"]) + " mode="))
PY
done; done; done
