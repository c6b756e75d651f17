for rep in 1 2; do for lib in gr_doa_b200/libdoa_cuda.so gr_doa_b200/_ab/libherk_iss1.so; do for hb in 592 512; do
DOA_AB_LIB=$lib HB=$hb python - <<PY
import os, sys
sys.path.insert(0, ".")
from gr_doa_b200 import _lib
_lib.LIB_PATH = os.path.abspath(os.environ["DOA_AB_LIB"])
exec(open("tools/herk_time.py").read().replace("mode=", os.path.basename(os.environ["DOA_AB_LIB"]) + " mode="))
PY
done; done; done
