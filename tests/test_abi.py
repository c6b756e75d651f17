"""CPU: libdoa_cuda.so loads, exports every symbol include/doa_cuda.h declares, validates arguments like the GRC
<check>s do, and FAILS LOUDLY (no fallback) when no CUDA device is present.  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT, has_cuda


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "doa_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(doa_cuda_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gr_doa_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/doa_cuda.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms, "ctypes table and header disagree"
    assert L.doa_cuda_abi_version() == 1


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under gr_doa_b200/ or include/ may import, link or mention it."""
    bad = []
    for base in ("gr_doa_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".h", ".cuh", ".cpp", ".cc")):
                    if re.search(r"\boracle\b", open(os.path.join(dp, fn), errors="ignore").read()):
                        bad.append(os.path.join(dp, fn))
    assert bad == []


def test_argument_validation_matches_grc_checks():
    """grc/doa_autocorrelate.xml:41-43 (overlap < snapshot, inputs > 0, snapshot > 0),
    grc/doa_MUSIC_lin_array.xml:33-35 (inputs > num_targets, norm_spacing <= 0.5)."""
    from gr_doa_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    E = _lib.EINVAL
    assert L.doa_cuda_autocorrelate_create(C.byref(h), 0, 2048, 512, 0, 0, 16) == E
    assert L.doa_cuda_autocorrelate_create(C.byref(h), 4, 0, 0, 0, 0, 16) == E
    assert L.doa_cuda_autocorrelate_create(C.byref(h), 4, 2048, 2048, 0, 0, 16) == E
    assert L.doa_cuda_autocorrelate_create(C.byref(h), 4, 2048, 512, 2, 0, 16) == E
    assert L.doa_cuda_music_create(C.byref(h), C.c_float(0.5), 4, 4, 1024, 0, 16) == E      # targets must be < elements
    assert L.doa_cuda_music_create(C.byref(h), C.c_float(0.6), 1, 4, 1024, 0, 16) == E      # spacing aliases
    assert L.doa_cuda_music_create(C.byref(h), C.c_float(0.5), 1, 4, 1, 0, 16) == E
    assert L.doa_cuda_rootmusic_create(C.byref(h), C.c_float(0.0), 1, 4, 0, 16) == E
    assert L.doa_cuda_find_local_max_create(C.byref(h), 0, 1024, C.c_float(0), C.c_float(180), 0, 16) == E
    assert L.doa_cuda_chain_create(C.byref(h), 8, 2048, 0, 0, C.c_float(0.5), 8, 4096, 3, C.c_float(0), C.c_float(180), 0, 16) == E
    assert not h.value
    assert b"num_targets" in L.doa_cuda_last_error(None)


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_create_fails_loudly_without_a_gpu():
    from gr_doa_b200 import _lib
    import gr_doa_b200 as doa
    L = _lib.lib()
    assert L.doa_cuda_device_count() == 0
    with pytest.raises(_lib.DoaCudaError) as ei:
        doa.autocorrelate(4, 2048, 512, 0)
    assert ei.value.code == _lib.ECUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(_lib.DoaCudaError):
        doa.DoaChain(8, 2048, 0, 0, 0.5, 3, 4096, 3)


def test_pin_host_buffer_rejects_bad_arguments():
    from gr_doa_b200 import _lib
    L = _lib.lib()
    assert L.doa_cuda_pin_host_buffer(None, 4096) == _lib.EINVAL
    import numpy as np
    a = np.zeros(1024, np.float32)
    assert L.doa_cuda_pin_host_buffer(a.ctypes.data, 0) == _lib.EINVAL
    assert L.doa_cuda_unpin_host_buffer(None) == _lib.EINVAL
    if not has_cuda():       # no device: the CUDA error comes back as a code and a text, nothing is pinned
        assert L.doa_cuda_pin_host_buffer(a.ctypes.data, a.nbytes) == _lib.ECUDA
        assert b"cudaHostRegister" in L.doa_cuda_last_error(None)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/doa_cuda.h compiles as strict C99 (no C++ types in any signature) and a C program that
    only uses the header links against libdoa_cuda.so and runs (without a device it gets the documented error code back)."""
    import subprocess
    from gr_doa_b200 import _lib
    src = tmp_path / "c_client.c"
    src.write_text(r'''
#include <stdio.h>
#include "doa_cuda.h"
int main(void) {
  doa_cuda_handle* h = 0;
  int rc;
  if (doa_cuda_abi_version() != 1) return 10;
  rc = doa_cuda_autocorrelate_create(&h, 4, 2048, 2048, 0, 0, 16);          /* overlap must be < snapshot */
  if (rc != DOA_CUDA_EINVAL || h != 0) return 11;
  rc = doa_cuda_autocorrelate_create(&h, 4, 2048, 512, 0, 0, 16);
  if (doa_cuda_device_count() == 0) { if (rc != DOA_CUDA_ECUDA || h != 0) return 12; }
  else { if (rc != DOA_CUDA_OK || doa_cuda_autocorrelate_forecast(h, 3) != 3 * 1536) return 13; doa_cuda_destroy(h); }
  printf("c client ok\n");
  return 0;
}
''')
    exe = tmp_path / "c_client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), "-L", libdir, "-ldoa_cuda", "-Wl,-rpath," + libdir])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "c client ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
