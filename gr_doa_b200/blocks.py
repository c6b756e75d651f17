"""Host-side mirror of the four gr-doa blocks on the DoA hot path, over libdoa_cuda's C ABI.

Same names, constructor parameters and port semantics as the reference's python namespace `doa`
(swig/doa_swig.i:22-36; grc/doa_*.xml <make> lines), so the parity tests read like python/qa_*.py:

    doa.autocorrelate(inputs, snapshot_size, overlap_size, avg_method)            include/doa/autocorrelate.h:56
    doa.MUSIC_lin_array(norm_spacing, num_targets, num_ant_ele, pspectrum_len)    include/doa/MUSIC_lin_array.h:56
    doa.rootMUSIC_linear_array(norm_spacing, num_targets, num_ant_ele)            include/doa/rootMUSIC_linear_array.h:54
    doa.find_local_max(num_max_vals, vector_len, x_min, x_max)                    include/doa/find_local_max.h:56

`work(...)` takes/returns numpy arrays in host memory (what a GNU Radio scheduler hands to work()) and goes through
the *_run entry points; `work_device(...)` takes/returns torch CUDA tensors and goes through *_run_device on torch's
current stream.  torch is used for device memory and streams only.  DoaChain is the fused
autocorrelate -> MUSIC_lin_array -> find_local_max path (peaks only).
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import check


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _stream_ptr(dev=None):
    """torch's current stream ON THE TENSOR'S DEVICE (a handle's stream argument must belong to the handle's device)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_live = weakref.WeakSet()
_default_options = {}


def set_default_option(key, value):
    """Measurement scripts (tools/): set a per-handle option (doa_cuda_set_option) on every live block of this process and on
    every block created from now on.  This convenience is Python-side state; the library itself has none."""
    _default_options[key] = int(value)
    for b in list(_live):
        try:
            b.set_option(key, value)
        except _lib.DoaCudaError:
            pass          # a product-library handle refusing a dev-only variant


class _Block:
    def __init__(self):
        self._h = C.c_void_p()
        self._L = _lib.lib()

    def _created(self, rc):
        if rc != 0:
            txt = self._L.doa_cuda_last_error(None)
            raise _lib.DoaCudaError(rc, txt.decode() if txt else "")
        _live.add(self)
        for k, v in _default_options.items():
            if self._L.doa_cuda_set_option(self._h, k.encode(), v) != 0:
                pass      # not applicable to this library build

    def launches(self):
        return int(self._L.doa_cuda_last_launch_count(self._h))

    def set_option(self, key, value):
        """Per-handle option (include/doa_cuda.h: doa_cuda_set_option): path selection for A/B measurements, SM reserve."""
        rc = self._L.doa_cuda_set_option(self._h, key.encode() if isinstance(key, str) else key, int(value))
        if rc != 0:
            txt = self._L.doa_cuda_last_error(self._h)
            raise _lib.DoaCudaError(rc, txt.decode() if txt else "")

    def _stream_of(self, t):
        """torch's current stream on the tensor's device, which must be the handle's device (include/doa_cuda.h, threading and
        streams: a handle's scratch lives on its device; one call at a time per handle)."""
        dev = getattr(self, "device", None)
        if dev is not None and t.device.index != dev:
            raise ValueError(f"tensor on cuda:{t.device.index}, handle created for cuda:{dev}")
        return _stream_ptr(t.device)

    def set_sms_reserve(self, n):
        """SMs the persistent chain kernel leaves to a collective's kernel (multi-GPU runs)."""
        self.set_option("sms_reserve", n)

    def set_channel_gains(self, gains):
        """Per-channel complex gains applied in front of the covariance as R' = D R D^H (autocorrelate and chain handles):
        what an antenna_correction / phase_correct_hier block upstream would have multiplied into the samples
        (lib/antenna_correction_impl.cc:90-96).  None removes them."""
        import numpy as np
        if gains is None:
            check(self._L.doa_cuda_set_channel_gains(self._h, None), self._h)
            return
        g = np.ascontiguousarray(np.asarray(gains, dtype=np.complex64).reshape(-1))
        if g.size != self.inputs:
            raise ValueError(f"expected {self.inputs} gains, got {g.size}")
        check(self._L.doa_cuda_set_channel_gains(self._h, g.ctypes.data), self._h)

    _sc16 = False

    def set_input_format(self, fmt="fc32", scale=1.0 / 32768):
        """Sample format of the covariance input (autocorrelate and chain handles).  "fc32": complex64 (gr_complex, the
        reference's only format).  "sc16": UHD's cpu_format "sc16" -- int16 arrays / tensors with a trailing axis of 2
        (I, Q); a sample's value is int16 * scale.  Converted exactly inside the covariance kernel (include/doa_cuda.h)."""
        if fmt not in ("fc32", "sc16"):
            raise ValueError("fmt must be 'fc32' or 'sc16'")
        check(self._L.doa_cuda_set_input_format(self._h, 1 if fmt == "sc16" else 0, C.c_float(scale)), self._h)
        self._sc16 = fmt == "sc16"

    def _samples(self, x):
        """Host samples in the handle's input format: contiguous complex64 [...], or int16 [..., 2] for sc16."""
        if not self._sc16:
            return _np(x, np.complex64)
        x = np.asarray(x)
        if x.dtype != np.int16 or x.shape[-1] != 2:
            raise ValueError("sc16 input: int16 array with a trailing (I, Q) axis of 2 expected")
        return np.ascontiguousarray(x)

    def _nsamp(self, x):
        return x.size // 2 if self._sc16 else x.size

    def _checked_streams(self, streams, nframes):
        """M channel streams for nframes frames: each must hold hop * (nframes - 1) + snapshot_size samples (history included),
        as general_work()'s input buffers do (lib/autocorrelate_impl.cc:74-80,95-100); a short array would be read past its end."""
        xs = [self._samples(x) for x in streams]
        if len(xs) != self.inputs:
            raise ValueError(f"expected {self.inputs} channel streams, got {len(xs)}")
        need = (nframes - 1) * self.hop + self.snapshot_size if nframes > 0 else 0
        for x in xs:
            if self._nsamp(x) < need:
                raise ValueError("not enough input items for nframes")
        return xs

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.doa_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class antenna_correction:
    """gr::doa::antenna_correction (lib/antenna_correction_impl.cc:47-99) as a gain source: reads the reference's config file
    ("gain phase" per channel, g_k = (1/gain_k) e^{-j phase_k}, same validation) into `.gains`; hand them to
    autocorrelate.set_channel_gains / DoaChain.set_channel_gains instead of running a multiply pass over the samples.
    There is deliberately no work(): the multiply happens inside the covariance kernels (R' = D R D^H), never on the host."""

    def __init__(self, num_inputs, config_filename):
        import numpy as np
        self.num_inputs = num_inputs
        L = _lib.lib()
        g = np.empty(num_inputs, np.complex64)
        rc = L.doa_cuda_antenna_gains_from_file(str(config_filename).encode(), num_inputs, g.ctypes.data)
        if rc != 0:
            txt = L.doa_cuda_last_error(None)
            raise _lib.DoaCudaError(rc, txt.decode() if txt else "")
        self.gains = g


class autocorrelate(_Block):
    """gr::doa::autocorrelate (lib/autocorrelate_impl.cc:47-118)."""

    def __init__(self, inputs, snapshot_size, overlap_size, avg_method, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.inputs, self.snapshot_size, self.overlap_size, self.avg_method = inputs, snapshot_size, overlap_size, int(avg_method)
        self.hop = snapshot_size - overlap_size
        self.max_frames = max_frames
        self._created(self._L.doa_cuda_autocorrelate_create(C.byref(self._h), inputs, snapshot_size, overlap_size,
                                                            int(avg_method), device, max_frames))

    def history(self):
        return self.overlap_size + 1          # set_history(), lib/autocorrelate_impl.cc:57

    def forecast(self, noutput_items):
        return int(self._L.doa_cuda_autocorrelate_forecast(self._h, noutput_items))

    def general_work(self, noutput_items, input_items):
        """input_items: M arrays of complex64, each with >= hop*(n-1)+snapshot_size samples (history included).
        Returns ([n][M*M] complex64, consumed_per_port = hop*n) like general_work() + consume_each()."""
        xs = [self._samples(x) for x in input_items]
        need = (noutput_items - 1) * self.hop + self.snapshot_size if noutput_items > 0 else 0
        for x in xs:
            if self._nsamp(x) < need:
                raise ValueError("not enough input items for noutput_items")
        out = np.empty((noutput_items, self.inputs * self.inputs), np.complex64)
        ptrs = (C.c_void_p * self.inputs)(*[x.ctypes.data for x in xs])
        check(self._L.doa_cuda_autocorrelate_run(self._h, ptrs, noutput_items, out.ctypes.data), self._h)
        return out, self.hop * noutput_items

    def work(self, streams):
        """Whole-stream convenience: streams [M][L] -> all complete frames."""
        x = self._samples(streams)
        L = x.shape[1]
        n = (L - self.snapshot_size) // self.hop + 1 if L >= self.snapshot_size else 0
        outs = []
        for f0 in range(0, n, self.max_frames):
            nf = min(self.max_frames, n - f0)
            o, _ = self.general_work(nf, [x[k, f0 * self.hop:] for k in range(self.inputs)])
            outs.append(o)
        return np.concatenate(outs) if outs else np.empty((0, self.inputs ** 2), np.complex64)

    def work_device(self, x, frame_stride=None, chan_stride=None, nframes=None):
        """x: torch complex64 CUDA tensor, [B][M][N] independent frames (default strides) or any strided layout
        (sc16 input format: int16 [B][M][N][2])."""
        import torch
        if frame_stride is None:
            B, M, N = x.shape[:3]
            frame_stride, chan_stride, nframes = M * N, N, B
        out = torch.empty((nframes, self.inputs * self.inputs), dtype=torch.complex64, device=x.device)
        check(self._L.doa_cuda_autocorrelate_run_device(self._h, x.data_ptr(), frame_stride, chan_stride, nframes,
                                                        out.data_ptr(), self._stream_of(x)), self._h)
        return out


class MUSIC_lin_array(_Block):
    """gr::doa::MUSIC_lin_array (lib/MUSIC_lin_array_impl.cc:47-150)."""

    def __init__(self, norm_spacing, num_targets, num_ant_ele, pspectrum_len, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.norm_spacing, self.num_targets, self.num_ant_ele, self.pspectrum_len = norm_spacing, num_targets, num_ant_ele, pspectrum_len
        self.max_frames = max_frames
        self.nout_items_total = 0             # public counter of the reference block (lib/MUSIC_lin_array_impl.h:47)
        self._created(self._L.doa_cuda_music_create(C.byref(self._h), C.c_float(norm_spacing), num_targets, num_ant_ele,
                                                    pspectrum_len, device, max_frames))

    def tables(self):
        M, P = self.num_ant_ele, self.pspectrum_len
        loc, th, V = np.empty(M, np.float32), np.empty(P, np.float32), np.empty((P, M), np.complex64)
        check(self._L.doa_cuda_music_get_tables(self._h, loc.ctypes.data, th.ctypes.data, V.ctypes.data), self._h)
        return loc, th, V

    def noise_subspace_device(self, R):
        """eig_sym + U_N*trans(U_N) only: returns (G [n][M*M] c64, u [n][M] c64, eigenvalues [n][M] f32) CUDA tensors."""
        import torch
        n, M = R.shape[0], self.num_ant_ele
        G = torch.empty((n, M * M), dtype=torch.complex64, device=R.device)
        u = torch.empty((n, M), dtype=torch.complex64, device=R.device)
        w = torch.empty((n, M), dtype=torch.float32, device=R.device)
        check(self._L.doa_cuda_music_noise_subspace_device(self._h, R.data_ptr(), n, G.data_ptr(), u.data_ptr(), w.data_ptr(),
                                                           self._stream_of(R)), self._h)
        return G, u, w

    def work(self, R):
        M = self.num_ant_ele
        R = _np(R, np.complex64).reshape(-1, M * M)
        out = np.empty((R.shape[0], self.pspectrum_len), np.float32)
        for f0 in range(0, R.shape[0], self.max_frames):
            nf = min(self.max_frames, R.shape[0] - f0)
            check(self._L.doa_cuda_music_run(self._h, R[f0:].ctypes.data, nf, out[f0:].ctypes.data), self._h)
        self.nout_items_total += R.shape[0]
        return out

    def work_device(self, R):
        import torch
        n = R.shape[0]
        out = torch.empty((n, self.pspectrum_len), dtype=torch.float32, device=R.device)
        check(self._L.doa_cuda_music_run_device(self._h, R.data_ptr(), n, out.data_ptr(), self._stream_of(R)), self._h)
        self.nout_items_total += n
        return out


class rootMUSIC_linear_array(_Block):
    """gr::doa::rootMUSIC_linear_array (lib/rootMUSIC_linear_array_impl.cc:46-152)."""

    def __init__(self, norm_spacing, num_targets, num_ant_ele, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.norm_spacing, self.num_targets, self.num_ant_ele = norm_spacing, num_targets, num_ant_ele
        self.max_frames = max_frames
        self._created(self._L.doa_cuda_rootmusic_create(C.byref(self._h), C.c_float(norm_spacing), num_targets, num_ant_ele,
                                                        device, max_frames))

    def work(self, R):
        M = self.num_ant_ele
        R = _np(R, np.complex64).reshape(-1, M * M)
        out = np.empty((R.shape[0], self.num_targets), np.float32)
        for f0 in range(0, R.shape[0], self.max_frames):
            nf = min(self.max_frames, R.shape[0] - f0)
            check(self._L.doa_cuda_rootmusic_run(self._h, R[f0:].ctypes.data, nf, out[f0:].ctypes.data), self._h)
        return out

    def work_device(self, R):
        import torch
        n = R.shape[0]
        out = torch.empty((n, self.num_targets), dtype=torch.float32, device=R.device)
        check(self._L.doa_cuda_rootmusic_run_device(self._h, R.data_ptr(), n, out.data_ptr(), self._stream_of(R)), self._h)
        return out


class find_local_max(_Block):
    """gr::doa::find_local_max (lib/find_local_max_impl.cc:47-194).  Returns (out0 peak values, out1 locations[, bins])."""

    def __init__(self, num_max_vals, vector_len, x_min, x_max, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.num_max_vals, self.vector_len, self.x_min, self.x_max = num_max_vals, vector_len, x_min, x_max
        self.max_frames = max_frames
        self._created(self._L.doa_cuda_find_local_max_create(C.byref(self._h), num_max_vals, vector_len, C.c_float(x_min),
                                                             C.c_float(x_max), device, max_frames))

    def work(self, vecs, return_bins=False):
        v = _np(vecs, np.float32).reshape(-1, self.vector_len)
        n, K = v.shape[0], self.num_max_vals
        val, loc, bins = np.empty((n, K), np.float32), np.empty((n, K), np.float32), np.empty((n, K), np.int32)
        for f0 in range(0, n, self.max_frames):
            nf = min(self.max_frames, n - f0)
            check(self._L.doa_cuda_find_local_max_run(self._h, v[f0:].ctypes.data, nf, val[f0:].ctypes.data,
                                                      loc[f0:].ctypes.data, bins[f0:].ctypes.data), self._h)
        return (val, loc, bins) if return_bins else (val, loc)

    def work_device(self, vecs):
        import torch
        n, K = vecs.shape[0], self.num_max_vals
        val = torch.empty((n, K), dtype=torch.float32, device=vecs.device)
        loc = torch.empty_like(val)
        bins = torch.empty((n, K), dtype=torch.int32, device=vecs.device)
        check(self._L.doa_cuda_find_local_max_run_device(self._h, vecs.data_ptr(), n, val.data_ptr(), loc.data_ptr(),
                                                         bins.data_ptr(), self._stream_of(vecs)), self._h)
        return val, loc, bins


class calibrate_lin_array(_Block):
    """gr::doa::calibrate_lin_array (lib/calibrate_lin_array_impl.cc:46-134): antenna gain/phase estimates from the covariance
    of a pilot at a known angle.  Output per covariance: num_ant_ele complex values, unit norm, defined up to a unit-modulus
    factor exactly like the reference's (LAPACK-phase) eigenvector."""

    def __init__(self, norm_spacing, num_ant_ele, pilot_angle, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.num_ant_ele, self.inputs = num_ant_ele, num_ant_ele
        self._created(self._L.doa_cuda_calibrate_create(C.byref(self._h), C.c_float(norm_spacing), num_ant_ele, C.c_float(pilot_angle),
                                                        device, max_frames))

    def work(self, R):
        M = self.num_ant_ele
        R = np.ascontiguousarray(np.asarray(R, dtype=np.complex64).reshape(-1, M * M))
        out = np.empty((R.shape[0], M), np.complex64)
        check(self._L.doa_cuda_calibrate_run(self._h, R.ctypes.data, R.shape[0], out.ctypes.data), self._h)
        return out

    def work_device(self, R):
        import torch
        n, M = R.shape[0], self.num_ant_ele
        out = torch.empty((n, M), dtype=torch.complex64, device=R.device)
        check(self._L.doa_cuda_calibrate_run_device(self._h, R.data_ptr(), n, out.data_ptr(), self._stream_of(R)), self._h)
        return out


class DoaChain(_Block):
    """autocorrelate -> MUSIC_lin_array -> find_local_max(K, P, x_min, x_max) in one call, peaks only."""

    def __init__(self, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets, pspectrum_len,
                 num_max_vals, x_min=0.0, x_max=180.0, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.inputs, self.snapshot_size, self.overlap_size = inputs, snapshot_size, overlap_size
        self.hop = snapshot_size - overlap_size
        self.K, self.max_frames = num_max_vals, max_frames
        self._created(self._L.doa_cuda_chain_create(C.byref(self._h), inputs, snapshot_size, overlap_size, int(avg_method),
                                                    C.c_float(norm_spacing), num_targets, pspectrum_len, num_max_vals,
                                                    C.c_float(x_min), C.c_float(x_max), device, max_frames))

    def set_profiling(self, on=True):
        check(self._L.doa_cuda_set_profiling(self._h, int(on)), self._h)

    def stage_ms(self):
        """Mean (cov, eig, scan) CUDA-event ms over the run_device calls since set_profiling(True)."""
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        n = self._L.doa_cuda_chain_stage_ms(self._h, C.byref(a), C.byref(b), C.byref(c))
        if n < 0:
            check(n, self._h)
        return a.value, b.value, c.value

    def run_device(self, x, out=None, frame_stride=None, chan_stride=None, nframes=None):
        """x: torch complex64 CUDA tensor [B][M][N] (or explicit strides; sc16 input format: int16 [B][M][N][2]).
        Returns (val, loc, bins) CUDA tensors."""
        import torch
        if frame_stride is None:
            B, M, N = x.shape[:3]
            frame_stride, chan_stride, nframes = M * N, N, B
        if out is None:
            val = torch.empty((nframes, self.K), dtype=torch.float32, device=x.device)
            out = (val, torch.empty_like(val), torch.empty((nframes, self.K), dtype=torch.int32, device=x.device))
        val, loc, bins = out
        check(self._L.doa_cuda_chain_run_device(self._h, x.data_ptr(), frame_stride, chan_stride, nframes, val.data_ptr(),
                                                loc.data_ptr(), bins.data_ptr(), self._stream_of(x)), self._h)
        return val, loc, bins

    def run_host(self, frames, out=None):
        """frames: host complex64 [B][M][N] (numpy, or a pinned torch tensor; sc16 input format: int16 [B][M][N][2]).
        Returns numpy (val, loc, bins)."""
        if hasattr(frames, "data_ptr"):
            ptr, B = frames.data_ptr(), frames.shape[0]
        else:
            frames = self._samples(frames)
            ptr, B = frames.ctypes.data, frames.shape[0]
        if out is None:
            out = (np.empty((B, self.K), np.float32), np.empty((B, self.K), np.float32), np.empty((B, self.K), np.int32))
        val, loc, bins = out
        p = [o.data_ptr() if hasattr(o, "data_ptr") else o.ctypes.data for o in (val, loc, bins)]
        check(self._L.doa_cuda_chain_run(self._h, ptr, B, p[0], p[1], p[2]), self._h)
        return val, loc, bins

    def run_streams(self, streams, nframes):
        xs = self._checked_streams(streams, nframes)
        val, loc, bins = (np.empty((nframes, self.K), np.float32), np.empty((nframes, self.K), np.float32),
                          np.empty((nframes, self.K), np.int32))
        ptrs = (C.c_void_p * self.inputs)(*[x.ctypes.data for x in xs])
        check(self._L.doa_cuda_chain_run_streams(self._h, ptrs, nframes, val.ctypes.data, loc.ctypes.data, bins.ctypes.data),
              self._h)
        return val, loc, bins


class DoaChainMulti(_Block):
    """DoaChain over several GPUs from ONE process (doa_cuda_multi_*): a batch of independent host frames is cut into contiguous
    blocks, one per entry of `devices`; every device runs the fused chain on its block concurrently and writes its peaks into
    the shared output arrays -- per frame the same bits as DoaChain on one device."""

    def __init__(self, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets, pspectrum_len,
                 num_max_vals, x_min=0.0, x_max=180.0, devices=None, max_frames_per_device=4096):
        super().__init__()
        if devices is None:
            devices = list(range(int(self._L.doa_cuda_device_count())))
        self.devices = [int(d) for d in devices]
        self.inputs, self.snapshot_size, self.overlap_size = inputs, snapshot_size, overlap_size
        self.hop = snapshot_size - overlap_size
        self.K, self.max_frames = num_max_vals, max_frames_per_device * max(1, len(self.devices))
        dv = (C.c_int * max(1, len(self.devices)))(*self.devices)
        self._created(self._L.doa_cuda_multi_create(C.byref(self._h), inputs, snapshot_size, overlap_size, int(avg_method),
                                                    C.c_float(norm_spacing), num_targets, pspectrum_len, num_max_vals,
                                                    C.c_float(x_min), C.c_float(x_max), dv, len(self.devices),
                                                    max_frames_per_device))

    def blocks(self, nframes):
        """[(first, count)] per device for a batch of nframes."""
        out = []
        for g in range(len(self.devices)):
            a, b = C.c_int(), C.c_int()
            check(self._L.doa_cuda_multi_block(self._h, nframes, g, C.byref(a), C.byref(b)), self._h)
            out.append((a.value, b.value))
        return out

    def run_host(self, frames, out=None):
        """frames: host [B][M][N] complex64 (numpy or pinned torch tensor; sc16 format: int16 [B][M][N][2]).
        Returns numpy (val, loc, bins)."""
        if hasattr(frames, "data_ptr"):
            ptr, B = frames.data_ptr(), frames.shape[0]
        else:
            frames = self._samples(frames)
            ptr, B = frames.ctypes.data, frames.shape[0]
        if out is None:
            out = (np.empty((B, self.K), np.float32), np.empty((B, self.K), np.float32), np.empty((B, self.K), np.int32))
        p = [o.data_ptr() if hasattr(o, "data_ptr") else o.ctypes.data for o in out]
        check(self._L.doa_cuda_multi_run(self._h, ptr, B, p[0], p[1], p[2]), self._h)
        return out

    def run_streams(self, streams, nframes):
        """streams: M channel arrays holding hop*(nframes-1)+snapshot_size samples each (a general_work() view)."""
        xs = self._checked_streams(streams, nframes)
        val, loc, bins = (np.empty((nframes, self.K), np.float32), np.empty((nframes, self.K), np.float32),
                          np.empty((nframes, self.K), np.int32))
        ptrs = (C.c_void_p * self.inputs)(*[x.ctypes.data for x in xs])
        check(self._L.doa_cuda_multi_run_streams(self._h, ptrs, nframes, val.ctypes.data, loc.ctypes.data, bins.ctypes.data),
              self._h)
        return val, loc, bins


class RootMusicChain(_Block):
    """autocorrelate -> rootMUSIC_linear_array in one call (doa_cuda_rootchain_*): samples in, num_targets ascending angles
    per frame out; the covariance never leaves the device."""

    def __init__(self, inputs, snapshot_size, overlap_size, avg_method, norm_spacing, num_targets, device=0, max_frames=4096):
        super().__init__()
        self.device = device
        self.inputs, self.snapshot_size, self.overlap_size = inputs, snapshot_size, overlap_size
        self.hop, self.T, self.max_frames = snapshot_size - overlap_size, num_targets, max_frames
        self._created(self._L.doa_cuda_rootchain_create(C.byref(self._h), inputs, snapshot_size, overlap_size, int(avg_method),
                                                        C.c_float(norm_spacing), num_targets, device, max_frames))

    def run_device(self, x, frame_stride=None, chan_stride=None, nframes=None):
        import torch
        if frame_stride is None:
            B, M, N = x.shape[:3]
            frame_stride, chan_stride, nframes = M * N, N, B
        out = torch.empty((nframes, self.T), dtype=torch.float32, device=x.device)
        check(self._L.doa_cuda_rootchain_run_device(self._h, x.data_ptr(), frame_stride, chan_stride, nframes, out.data_ptr(),
                                                    self._stream_of(x)), self._h)
        return out

    def run_host(self, frames):
        frames = self._samples(frames)
        out = np.empty((frames.shape[0], self.T), np.float32)
        check(self._L.doa_cuda_rootchain_run(self._h, frames.ctypes.data, frames.shape[0], out.ctypes.data), self._h)
        return out

    def run_streams(self, streams, nframes):
        xs = self._checked_streams(streams, nframes)
        out = np.empty((nframes, self.T), np.float32)
        ptrs = (C.c_void_p * self.inputs)(*[x.ctypes.data for x in xs])
        check(self._L.doa_cuda_rootchain_run_streams(self._h, ptrs, nframes, out.ctypes.data), self._h)
        return out
