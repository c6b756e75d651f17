"""ctypes binding of libdoa_cuda.so (include/doa_cuda.h).  Fails loudly when the library is missing: there is no
fallback path of any kind in this package."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdoa_cuda.so")
DEV_LIB_PATH = os.path.join(_HERE, "libdoa_cuda_dev.so")      # -DDOA_DEV_KNOBS build: experimental kernel variants (tools/, tests)
_lib = None
_dev_lib = None
_use_dev = False

OK, EINVAL, ECUDA, ENOMEM, ECAPACITY = 0, -1, -2, -3, -4

# every symbol include/doa_cuda.h declares: name -> (restype, argtypes)
_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_hp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "doa_cuda_abi_version": (_i, []),
    "doa_cuda_last_error": (C.c_char_p, [_vp]),
    "doa_cuda_device_count": (_i, []),
    "doa_cuda_autocorrelate_create": (_i, [_hp, _i, _i, _i, _i, _i, _i]),
    "doa_cuda_autocorrelate_run": (_i, [_vp, _hp, _i, _vp]),
    "doa_cuda_autocorrelate_run_device": (_i, [_vp, _vp, _ll, _ll, _i, _vp, _vp]),
    "doa_cuda_autocorrelate_forecast": (_i, [_vp, _i]),
    "doa_cuda_music_create": (_i, [_hp, _f, _i, _i, _i, _i, _i]),
    "doa_cuda_music_run": (_i, [_vp, _vp, _i, _vp]),
    "doa_cuda_music_run_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "doa_cuda_music_get_tables": (_i, [_vp, _vp, _vp, _vp]),
    "doa_cuda_music_noise_subspace_device": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "doa_cuda_rootmusic_create": (_i, [_hp, _f, _i, _i, _i, _i]),
    "doa_cuda_rootmusic_run": (_i, [_vp, _vp, _i, _vp]),
    "doa_cuda_rootmusic_run_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "doa_cuda_find_local_max_create": (_i, [_hp, _i, _i, _f, _f, _i, _i]),
    "doa_cuda_find_local_max_run": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "doa_cuda_find_local_max_run_device": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "doa_cuda_chain_create": (_i, [_hp, _i, _i, _i, _i, _f, _i, _i, _i, _f, _f, _i, _i]),
    "doa_cuda_chain_run_device": (_i, [_vp, _vp, _ll, _ll, _i, _vp, _vp, _vp, _vp]),
    "doa_cuda_chain_run": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "doa_cuda_chain_run_streams": (_i, [_vp, _hp, _i, _vp, _vp, _vp]),
    "doa_cuda_last_launch_count": (_i, [_vp]),
    "doa_cuda_set_profiling": (_i, [_vp, _i]),
    "doa_cuda_chain_stage_ms": (_i, [_vp, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f)]),
    "doa_cuda_calibrate_create": (_i, [_hp, _f, _i, _f, _i, _i]),
    "doa_cuda_calibrate_run": (_i, [_vp, _vp, _i, _vp]),
    "doa_cuda_calibrate_run_device": (_i, [_vp, _vp, _i, _vp, _vp]),
    "doa_cuda_set_channel_gains": (_i, [_vp, _vp]),
    "doa_cuda_antenna_gains_from_file": (_i, [C.c_char_p, _i, _vp]),
    "doa_cuda_set_input_format": (_i, [_vp, _i, _f]),
    "doa_cuda_rootchain_create": (_i, [_hp, _i, _i, _i, _i, _f, _i, _i, _i]),
    "doa_cuda_rootchain_run_device": (_i, [_vp, _vp, _ll, _ll, _i, _vp, _vp]),
    "doa_cuda_rootchain_run": (_i, [_vp, _vp, _i, _vp]),
    "doa_cuda_rootchain_run_streams": (_i, [_vp, _hp, _i, _vp]),
    "doa_cuda_multi_create": (_i, [_hp, _i, _i, _i, _i, _f, _i, _i, _i, _f, _f, C.POINTER(_i), _i, _i]),
    "doa_cuda_multi_run": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "doa_cuda_multi_run_streams": (_i, [_vp, _hp, _i, _vp, _vp, _vp]),
    "doa_cuda_multi_device_count": (_i, [_vp]),
    "doa_cuda_multi_block": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "doa_cuda_pin_host_buffer": (_i, [_vp, C.c_ulonglong]),
    "doa_cuda_unpin_host_buffer": (_i, [_vp]),
    "doa_cuda_set_option": (_i, [_vp, C.c_char_p, _i]),
    "doa_cuda_has_dev_knobs": (_i, []),
    "doa_cuda_destroy": (None, [_vp]),
}


class DoaCudaError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libdoa_cuda error {code}: {text}")
        self.code = code


def _load(path):
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m gr_doa_b200.build" + (" --dev" if path == DEV_LIB_PATH else "") + "` "
                          "(libdoa_cuda is the only compute path; there is no CPU fallback)")
    L = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)   # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    return L


def lib():
    """libdoa_cuda.so (built in-tree by gr_doa_b200.build) -- or, inside a `dev_library()` block, libdoa_cuda_dev.so.
    Raises if absent: no fallback exists."""
    global _lib, _dev_lib
    if _use_dev:
        if _dev_lib is None:
            _dev_lib = _load(DEV_LIB_PATH)
        return _dev_lib
    if _lib is None:
        _lib = _load(LIB_PATH)
    return _lib


class dev_library:
    """Context manager: blocks CREATED inside it are bound to libdoa_cuda_dev.so, the -DDOA_DEV_KNOBS build that also contains
    the experimental kernel configurations (a block keeps the library it was created with)."""

    def __enter__(self):
        global _use_dev
        self._prev = _use_dev
        _use_dev = True
        return lib()

    def __exit__(self, *exc):
        global _use_dev
        _use_dev = self._prev
        return False


def check(rc, handle=None):
    if rc != OK:
        txt = lib().doa_cuda_last_error(handle)
        raise DoaCudaError(rc, txt.decode() if txt else "")
    return rc
