"""ctypes binding of libspleeterrt_b200.so — mirrors the reference's operator interface.

Names follow the reference (james34602/SpleeterRT): ``Separator.process_spleeter`` is
``processSpleeter`` (Executable/spleeter.c:177), ``stft``/``istft`` are stftFix.c:363/496,
``separate`` is the CLI's main.c:762-806 flow for a batch of streams and n_stems nets.
"""
import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
FFTSIZE, HOPSIZE, BINS = 4096, 1024, 2049
COEFF_FLOATS = 9822725

# every symbol include/*.h declares (checked by the CPU test-suite against the built library)
HEADER_SYMBOLS = {
    "srt_b200.h": ["srt_create", "srt_create_cli", "srt_output_pairs", "srt_destroy", "srt_last_error", "srt_half_to_float",
                   "srt_load_coeff_dat", "srt_save_coeff_dat", "srt_model_fp16_nets", "srt_load_model_fp16", "srt_pack_layer", "srt_unet_host",
                   "srt_unet_device", "srt_separate_batch", "srt_separate_batch_async", "srt_batch_wait",
                   "srt_separate_device", "srt_separate_batch_interleaved", "srt_separate_batch_interleaved_async",
                   "srt_separate_device_interleaved", "srt_resample_frames", "srt_resample_host", "srt_resample_device",
                   "srt_resample_plan", "srt_stft_rows",
                   "srt_stft_host", "srt_istft_host", "srt_launch_count", "srt_last_timing", "srt_set_timing",
                   "srt_debug_tensor", "srt_probe_tensor_peak", "srt_probe_copy_bandwidth", "srt_cuda_stream", "srt_host_alloc", "srt_host_free", "srt_synchronize",
                   "srt_stream_create", "srt_stream_process", "srt_stream_destroy", "srt_stream_launch_count"],
    "spleeter.h": ["getCoeffSize", "allocateSpleeterStr", "initSpleeter", "getMaskPtr", "processSpleeter",
                   "freeSpleeter"],
    "stftFix.h": ["InitSTFT", "FreeSTFT", "stft", "istft"],
    "Spleeter4Stems.h": ["Spleeter4StemsInit", "Spleeter4StemsFree", "Spleeter4StemsProcessSamples"],
}
# include/srt_dispatch.h lives in its own library (libspleeterrt_dispatch.so: links libnccl)
DISPATCH_SYMBOLS = ["srt_dispatch_get_id", "srt_dispatch_create", "srt_dispatch_destroy", "srt_dispatch_last_error",
                    "srt_dispatch_broadcast_weights", "srt_dispatch_broadcast_sizes", "srt_dispatch_separate_device", "srt_dispatch_wait",
                    "srt_dispatch_comm_stream", "srt_dispatch_schedule", "srt_dispatch_local_streams",
                    "srt_dispatch_peer_layout", "srt_dispatch_peer_buffers", "srt_dispatch_separate_peer"]


class SrtError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("device", C.c_int), ("n_stems", C.c_int), ("time_step", C.c_int), ("bin_limit", C.c_int),
                ("max_images", C.c_int), ("max_batch_images", C.c_int), ("flavour", C.c_int),
                ("conv_impl", C.c_int), ("cuda_stream", C.c_void_p), ("precision", C.c_int), ("share_weights", C.c_int)]


PRECISION_COMPENSATED, PRECISION_TF32, PRECISION_COMPENSATED_BF16 = 0, 1, 2      # srt_config.precision (include/srt_b200.h)


def _precision(p):
    if p is None:
        return PRECISION_COMPENSATED
    if p in ("compensated", PRECISION_COMPENSATED):
        return PRECISION_COMPENSATED
    if p in ("tf32", PRECISION_TF32) and p is not True:
        return PRECISION_TF32
    if p in ("compensated_bf16", "tf32+bf16", PRECISION_COMPENSATED_BF16):
        return PRECISION_COMPENSATED_BF16
    raise SrtError(f"unknown precision {p!r} (compensated | compensated_bf16 | tf32)")


def lib_path():
    return os.path.join(PKG, "libspleeterrt_b200.so")


_lib = None


def load_library():
    """Load the CUDA library.  Fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise SrtError(f"{path} is missing: build it with `python -m spleeterrt_b200.build` "
                       "(the product has no CPU / PyTorch fallback)")
    lib = C.CDLL(path)
    lib.srt_last_error.restype = C.c_char_p
    lib.srt_create.argtypes = [C.POINTER(_Config), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.srt_destroy.argtypes = [C.c_void_p]
    lib.srt_create_cli.argtypes = [C.POINTER(_Config), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.srt_output_pairs.argtypes = [C.c_void_p]
    lib.srt_unet_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.srt_unet_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.srt_separate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.srt_separate_device.argtypes = lib.srt_separate_batch.argtypes
    lib.srt_separate_batch_async.argtypes = lib.srt_separate_batch.argtypes + [C.POINTER(C.c_int)]
    lib.srt_batch_wait.argtypes = [C.c_void_p, C.c_int]
    lib.srt_separate_batch_interleaved.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.srt_separate_device_interleaved.argtypes = lib.srt_separate_batch_interleaved.argtypes
    lib.srt_separate_batch_interleaved_async.argtypes = lib.srt_separate_batch_interleaved.argtypes + [C.POINTER(C.c_int)]
    lib.srt_stft_rows.restype = C.c_size_t
    lib.srt_stft_rows.argtypes = [C.c_size_t]
    lib.srt_stft_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t] + [C.c_void_p] * 4
    lib.srt_istft_host.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p, C.c_void_p]
    lib.srt_launch_count.restype = C.c_longlong
    lib.srt_launch_count.argtypes = [C.c_void_p]
    lib.srt_last_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    lib.srt_set_timing.argtypes = [C.c_void_p, C.c_int]
    lib.srt_debug_tensor.restype = C.c_longlong
    lib.srt_debug_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
    lib.srt_host_alloc.restype = C.c_void_p
    lib.srt_host_alloc.argtypes = [C.c_size_t]
    lib.srt_host_free.argtypes = [C.c_void_p]
    lib.srt_synchronize.argtypes = [C.c_void_p]
    lib.srt_half_to_float.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.srt_resample_frames.restype = C.c_size_t
    lib.srt_resample_frames.argtypes = [C.c_size_t, C.c_double]
    lib.srt_resample_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.srt_resample_device.argtypes = lib.srt_resample_host.argtypes
    lib.srt_resample_plan.restype = C.c_longlong
    lib.srt_resample_plan.argtypes = [C.c_size_t, C.c_int, C.c_double, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.srt_load_coeff_dat.argtypes = [C.c_char_p, C.c_void_p]
    lib.srt_save_coeff_dat.argtypes = [C.c_char_p, C.c_void_p]
    lib.srt_model_fp16_nets.argtypes = [C.c_char_p]
    lib.srt_load_model_fp16.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
    lib.srt_pack_layer.restype = C.c_longlong
    lib.srt_pack_layer.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.srt_stream_create.argtypes = [C.POINTER(_Config), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.srt_stream_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.srt_stream_destroy.argtypes = [C.c_void_p]
    lib.srt_stream_launch_count.restype = C.c_longlong
    lib.srt_stream_launch_count.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def exported_symbols(path=None):
    """Dynamic symbols of the built library (used by the CPU tests; no GPU needed)."""
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", path or lib_path()], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def dispatch_lib_path():
    return os.path.join(PKG, "libspleeterrt_dispatch.so")


_dlib = None


def load_dispatch_library():
    """libspleeterrt_dispatch.so (include/srt_dispatch.h).  Load torch (its NCCL) first when the process uses torch.distributed."""
    global _dlib
    if _dlib is not None:
        return _dlib
    load_library()
    if not os.path.exists(dispatch_lib_path()):
        raise SrtError("libspleeterrt_dispatch.so has not been built: python -m spleeterrt_b200.build")
    lib = C.CDLL(dispatch_lib_path())
    lib.srt_dispatch_last_error.restype = C.c_char_p
    lib.srt_dispatch_get_id.argtypes = [C.c_void_p]
    lib.srt_dispatch_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.srt_dispatch_destroy.argtypes = [C.c_void_p]
    lib.srt_dispatch_broadcast_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.srt_dispatch_broadcast_sizes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.srt_dispatch_separate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_int]
    lib.srt_dispatch_wait.argtypes = [C.c_void_p]
    lib.srt_dispatch_comm_stream.restype = C.c_void_p
    lib.srt_dispatch_comm_stream.argtypes = [C.c_void_p]
    lib.srt_dispatch_schedule.restype = C.c_longlong
    lib.srt_dispatch_schedule.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong]
    lib.srt_dispatch_local_streams.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.srt_dispatch_peer_layout.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.srt_dispatch_peer_buffers.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.srt_dispatch_separate_peer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    _dlib = lib
    return lib


def dispatch_schedule(world, rank, root, n_samples, pairs, chunks):
    """The point-to-point schedule of one rank (host only): int32[n][6] rows {group, kind, peer, stream, slot, count}."""
    lib = load_dispatch_library()
    ns = (C.c_size_t * len(n_samples))(*[int(x) for x in n_samples])
    n = lib.srt_dispatch_schedule(world, rank, root, ns, len(n_samples), pairs, chunks, None, 0)
    if n < 0:
        raise SrtError(lib.srt_dispatch_last_error().decode())
    rows = np.zeros((max(n, 1), 6), np.int32)
    lib.srt_dispatch_schedule(world, rank, root, ns, len(n_samples), pairs, chunks, rows.ctypes.data, n)
    return rows[:n]


class NcclDispatcher:
    """include/srt_dispatch.h: one per rank.  `exchange_id(id_bytes_or_None) -> id_bytes` is the launcher's job (rank 0 passes its
    fresh id, everybody gets it back): torch.distributed.broadcast_object_list under torchrun, a file or MPI elsewhere."""

    def __init__(self, world, rank, device, exchange_id):
        self.lib = load_dispatch_library()
        self.world, self.rank = world, rank
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self.lib.srt_dispatch_get_id(buf))
        ident = exchange_id(bytes(buf) if rank == 0 else None)
        buf = (C.c_ubyte * 128)(*ident)
        h = C.c_void_p()
        self._check(self.lib.srt_dispatch_create(buf, world, rank, device, C.byref(h)))
        self.h = h

    def _check(self, rc):
        if rc != 0:
            raise SrtError(f"dispatch error {rc}: {self.lib.srt_dispatch_last_error().decode()}")

    def broadcast_nets(self, nets, n_nets, root=0):
        """nets: [(coeff, mode)] on root (None elsewhere) -> the same list on every rank (one ncclBroadcast of the blobs)."""
        flat = np.zeros((n_nets, COEFF_FLOATS), np.float32)
        modes = np.zeros(n_nets, np.uint64)
        if self.rank == root:
            for k, (c, m) in enumerate(nets):
                flat[k] = np.asarray(c, np.float32)
                modes[k] = int(m)
        self._check(self.lib.srt_dispatch_broadcast_weights(self.h, flat.ctypes.data, n_nets, root))
        self._check(self.lib.srt_dispatch_broadcast_sizes(self.h, modes.ctypes.data, n_nets, root))
        return [(np.ascontiguousarray(flat[k]), int(modes[k])) for k in range(n_nets)]

    def separate_device(self, sep, root, pl, pr, n_arr, n_streams, uw, po, chunks=4):
        """pl / pr / po: ctypes arrays of device pointers on root (None elsewhere); n_arr: c_size_t array, same on all ranks."""
        self._check(self.lib.srt_dispatch_separate_device(self.h, sep.h, root, pl, pr, n_arr, n_streams, uw, po, chunks))

    def peer_buffers(self, root, n_samples, pairs):
        """Peer-memory mode: root's PCM / stem buffers for this batch shape as this rank sees them.
        Returns (d_in, d_out, in_off, out_off, in_floats, out_floats): device addresses and per-stream float offsets."""
        ns = len(n_samples)
        n_arr = (C.c_size_t * ns)(*[int(x) for x in n_samples])
        in_off, out_off = (C.c_size_t * ns)(), (C.c_size_t * ns)()
        fi, fo = C.c_size_t(), C.c_size_t()
        self._check(self.lib.srt_dispatch_peer_layout(n_arr, ns, pairs, in_off, out_off, C.byref(fi), C.byref(fo)))
        pi, po = C.c_void_p(), C.c_void_p()
        self._check(self.lib.srt_dispatch_peer_buffers(self.h, root, fi.value, fo.value, C.byref(pi), C.byref(po)))
        return pi.value, po.value, list(in_off), list(out_off), fi.value, fo.value

    def separate_peer(self, sep, root, n_arr, n_streams, uw=None, chunks=4):
        self._check(self.lib.srt_dispatch_separate_peer(self.h, sep.h, root, n_arr, n_streams, uw, chunks))

    def wait(self):
        self._check(self.lib.srt_dispatch_wait(self.h))

    def comm_stream(self):
        return self.lib.srt_dispatch_comm_stream(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_dispatch_destroy(self.h)
            self.h = None


def probe_tensor_peak(kind="tf32", seconds=0.1, device=0):
    """TFLOP/s of back-to-back N = 256 tcgen05 MMAs on every SM (srt_probe_tensor_peak): kind "tf32" or "bf16"."""
    lib = load_library()
    out = C.c_double(0.0)
    lib.srt_probe_tensor_peak.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]
    rc = lib.srt_probe_tensor_peak(device, 0 if kind == "tf32" else 1, float(seconds), C.byref(out))
    if rc != 0:
        raise SrtError(f"srt error {rc}: {lib.srt_last_error().decode()}")
    return out.value


def probe_copy_bandwidth(nbytes=1 << 30, device=0):
    lib = load_library()
    out = C.c_double(0.0)
    lib.srt_probe_copy_bandwidth.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_double)]
    rc = lib.srt_probe_copy_bandwidth(device, int(nbytes), C.byref(out))
    if rc != 0:
        raise SrtError(f"srt error {rc}: {lib.srt_last_error().decode()}")
    return out.value


def pinned_empty(shape, dtype=np.float32):
    """A numpy array in page-locked host memory (srt_host_alloc).  Hand such arrays to separate_async() / separate_raw_async():
    cudaMemcpyAsync from pageable memory is staged synchronously by the driver, which silently turns the three-batch
    pipeline into blocking copies.  The memory is released when the array (and every view of it) is garbage-collected."""
    lib = load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    lib.srt_host_alloc.restype = C.c_void_p
    lib.srt_host_alloc.argtypes = [C.c_size_t]
    lib.srt_host_free.argtypes = [C.c_void_p]
    p = lib.srt_host_alloc(max(n, 1))
    if not p:
        raise SrtError("srt_host_alloc failed")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.srt_host_free(self.ptr)
            except Exception:
                pass
    buf = (C.c_char * max(n, 1)).from_address(p)
    buf._owner = _Owner(p)                      # the ctypes buffer keeps the allocation alive; numpy keeps the buffer
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def half_to_float(halves):
    """fp16 model blob -> fp32, denormals as zero (f32Decompress, main.c:423-434)."""
    h = np.ascontiguousarray(halves, dtype=np.uint16)
    out = np.empty(h.shape, np.float32)
    load_library().srt_half_to_float(h.ctypes.data, out.ctypes.data, h.size)
    return out


TIMING_CATEGORIES = {"down2": 0, "down3": 1, "down4": 2, "down5": 3, "down6": 4, "up1": 5, "up2": 6, "up3": 7,
                     "up4": 8, "up5": 9, "down1": 10, "up6": 11, "up7": 12, "stft": 13, "istft": 14, "ola": 15,
                     "h2d": 16, "d2h": 17}


def _host_check(lib, rc):
    if rc < 0:
        raise SrtError(f"srt error {rc}: {lib.srt_last_error().decode()}")
    return rc


def load_coeff_dat(path):
    """One fp32 spleeterCoeff dump (the VST's *.dat, PluginProcessor.cpp:48-80) -> float32[9822725]."""
    lib = load_library()
    out = np.empty(COEFF_FLOATS, np.float32)
    _host_check(lib, lib.srt_load_coeff_dat(os.fsencode(path), out.ctypes.data))
    return out


def save_coeff_dat(path, coeff):
    lib = load_library()
    coeff = np.ascontiguousarray(coeff, np.float32)
    if coeff.size != COEFF_FLOATS:
        raise SrtError("a net is one spleeterCoeff blob of 9822725 floats")
    _host_check(lib, lib.srt_save_coeff_dat(os.fsencode(path), coeff.ctypes.data))


def load_model_fp16(path):
    """The fp16 model blob (spleeterQuantized) -> list of float32[9822725], one per net (main.c:435-443)."""
    lib = load_library()
    nets = _host_check(lib, lib.srt_model_fp16_nets(os.fsencode(path)))
    out = []
    for k in range(nets):
        w = np.empty(COEFF_FLOATS, np.float32)
        _host_check(lib, lib.srt_load_model_fp16(os.fsencode(path), k, w.ctypes.data))
        out.append(w)
    return out


def pack_layer(coeff, layer, time_step, bin_limit, form=0):
    """Packed tcgen05 B-operand blob of one tensor-core layer (srt_pack_layer)."""
    lib = load_library()
    coeff = np.ascontiguousarray(coeff, np.float32)
    n = _host_check(lib, lib.srt_pack_layer(layer, form, time_step, bin_limit, coeff.ctypes.data, None, 0))
    out = np.empty(n, np.float32)
    _host_check(lib, lib.srt_pack_layer(layer, form, time_step, bin_limit, coeff.ctypes.data, out.ctypes.data, out.size))
    return out


RESAMPLER_INDEX_INC = 491      # Executable/libsamplerate/src_sinc.c:143


def resample_plan(n_in, channels, ratio, coeff_count=22438, index_inc=RESAMPLER_INDEX_INC):
    """(input frame, table offset) per output frame of the reference's one-shot converter; host only."""
    lib = load_library()
    n_out = lib.srt_resample_frames(n_in, float(ratio))
    fr, st = np.zeros(max(n_out, 1), np.int32), np.zeros(max(n_out, 1), np.int32)
    gen = _host_check(lib, lib.srt_resample_plan(n_in, channels, float(ratio), coeff_count, index_inc, n_out, fr.ctypes.data, st.ctypes.data))
    return fr[:gen], st[:gen], n_out


class Separator:
    """One context on one B200 for (n_stems nets, T, F).

    nets: list of (coeff float32[9822725], stemMode) — stemMode 0 = LeakyReLU/ReLU, else ELU
    (initSpleeter, Executable/spleeter.c:130-139).
    """

    def __init__(self, nets, time_step, bin_limit, max_images=1, max_batch_images=0, device=0, flavour=0,
                 conv_impl=None, cuda_stream=None, precision=None, share_weights=False):
        self.lib = load_library()
        if conv_impl is None:
            conv_impl = 1 if os.environ.get("SRT_CONV_IMPL", "") == "simt" else 0
        cfg = _Config(device, len(nets), time_step, bin_limit, max_images, max_batch_images, flavour, conv_impl,
                      cuda_stream, _precision(precision), 1 if share_weights else 0)
        self.S, self.T, self.F = len(nets), time_step, bin_limit
        self.max_images = max_images
        coeffs = [np.ascontiguousarray(c, np.float32) for c, _ in nets]
        for c in coeffs:
            if c.size != COEFF_FLOATS:
                raise SrtError("each net must be one spleeterCoeff blob of 9822725 floats")
        cp = (C.c_void_p * max(self.S, 1))(*[c.ctypes.data for c in coeffs])
        modes = (C.c_int * max(self.S, 1))(*[int(m) for _, m in nets])
        h = C.c_void_p()
        self._check(self.lib.srt_create(C.byref(cfg), cp if self.S else None, modes if self.S else None, C.byref(h)))
        self.h = h
        self._coeffs = coeffs if share_weights else None      # shared sets are keyed by these host pointers: keep them alive

    def _check(self, rc):
        if rc != 0:
            raise SrtError(f"srt error {rc}: {self.lib.srt_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- processSpleeter ------------------------------------------------------------------
    def process_spleeter(self, x):
        """x: float32[n_img][2][T][F] (or [2][T][F]) magnitudes -> masks float32[S][n_img][2][T][F]."""
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 3:
            x = x[None]
        n = x.shape[0]
        assert x.shape[1:] == (2, self.T, self.F)
        y = np.empty((self.S, n, 2, self.T, self.F), np.float32)
        self._check(self.lib.srt_unet_host(self.h, x.ctypes.data, n, y.ctypes.data))
        return y

    def unet_device(self, d_mag_ptr, n_img, d_mask_ptr):
        self._check(self.lib.srt_unet_device(self.h, d_mag_ptr, n_img, d_mask_ptr))

    def debug_tensor(self, name, n_img):
        level = {"skip": lambda i: i, "up": lambda i: 6 - i}
        kind = "skip" if name.startswith("skip") else "up"
        i = int(name[-1])
        lv = 0 if name == "up6" else level[kind](i)
        ch = {"skip": [0, 16, 32, 64, 128, 256, 512], "up": [0, 256, 128, 64, 32, 16, 1]}[kind][i]
        H, W = self.T >> lv, self.F >> lv
        out = np.empty((self.S, n_img, ch, H, W), np.float32)
        rc = self.lib.srt_debug_tensor(self.h, name.encode(), out.ctypes.data, out.size)
        if rc < 0:
            self._check(int(rc))
        return out

    # ---- full path ------------------------------------------------------------------------
    def separate(self, streams, unaffected=None):
        """streams: list of (L, R) float32 arrays.  Returns list of float32[S][2][n]."""
        ns = len(streams)
        Ls, Rs = self._channels(streams)
        n = (C.c_size_t * ns)(*[l.size for l in Ls])
        pl = (C.c_void_p * ns)(*[l.ctypes.data for l in Ls])
        pr = (C.c_void_p * ns)(*[r.ctypes.data for r in Rs])
        outs = [np.empty((self.S, 2, l.size), np.float32) for l in Ls]
        po = (C.c_void_p * (ns * self.S * 2))(*[o[s, c].ctypes.data for o in outs for s in range(self.S) for c in range(2)])
        uw = None
        if unaffected is not None:
            uw = (C.c_float * self.S)(*[float(u) for u in unaffected])
        self._check(self.lib.srt_separate_batch(self.h, pl, pr, n, ns, uw, po))
        return outs

    @staticmethod
    def _channels(streams):
        """contiguous float32 (L, R) per stream; the C side reads n = L.size samples of BOTH, so a shorter R is an error"""
        Ls = [np.ascontiguousarray(l, np.float32).ravel() for l, _ in streams]
        Rs = [np.ascontiguousarray(r, np.float32).ravel() for _, r in streams]
        for i, (l, r) in enumerate(zip(Ls, Rs)):
            if l.size == 0 or l.size != r.size:
                raise SrtError(f"stream {i}: L has {l.size} samples, R has {r.size}; both channels need the same non-zero length")
        return Ls, Rs

    def separate_interleaved(self, frames, unaffected=None):
        """frames: list of float32[n][channels] (channels 1 or 2) or float32[n] (mono): interleaved frames as a WAV
        decoder yields them (main.c:767-769).  Returns a list of float32[pairs][n][2]: interleaved stereo frames per
        output pair, ready for a float32 WAV writer (main.c:806)."""
        ns = len(frames)
        X = [np.ascontiguousarray(x, np.float32).reshape(len(x), -1) for x in frames]
        ch = (C.c_int * ns)(*[x.shape[1] for x in X])
        n = (C.c_size_t * ns)(*[x.shape[0] for x in X])
        pp = (C.c_void_p * ns)(*[x.ctypes.data for x in X])
        outs = [np.empty((self.S, x.shape[0], 2), np.float32) for x in X]
        po = (C.c_void_p * (ns * self.S))(*[o[q].ctypes.data for o in outs for q in range(self.S)])
        uw = (C.c_float * self.S)(*[float(u) for u in unaffected]) if unaffected is not None else None
        self._check(self.lib.srt_separate_batch_interleaved(self.h, pp, ch, n, ns, uw, po))
        return outs

    def separate_async(self, streams, unaffected=None):
        """Like separate(), but only enqueues the batch (srt_separate_batch_async).  Returns a pending handle;
        result(handle) waits and returns the stems.  Up to three batches are in flight per Separator; a fourth submit first drains the oldest.
        The copies overlap the kernels only for page-locked buffers: pass streams allocated with pinned_empty() (outputs are
        allocated that way here); pageable numpy arrays still work but their H2D copies are staged synchronously."""
        ns = len(streams)
        Ls, Rs = self._channels(streams)
        n = (C.c_size_t * ns)(*[l.size for l in Ls])
        pl = (C.c_void_p * ns)(*[l.ctypes.data for l in Ls])
        pr = (C.c_void_p * ns)(*[r.ctypes.data for r in Rs])
        outs = [pinned_empty((self.S, 2, l.size), np.float32) for l in Ls]
        po = (C.c_void_p * (ns * self.S * 2))(*[o[s, c].ctypes.data for o in outs for s in range(self.S) for c in range(2)])
        uw = (C.c_float * self.S)(*[float(u) for u in unaffected]) if unaffected is not None else None
        ticket = self.separate_raw_async(pl, pr, n, ns, uw, po)
        return {"ticket": ticket, "outs": outs, "keepalive": (Ls, Rs, n, pl, pr, po, uw)}

    def result(self, pending):
        self.wait(pending["ticket"])
        return pending["outs"]

    def separate_raw(self, pl, pr, n, ns, uw, po, device=False):
        fn = self.lib.srt_separate_device if device else self.lib.srt_separate_batch
        self._check(fn(self.h, pl, pr, n, ns, uw, po))

    def separate_raw_async(self, pl, pr, n, ns, uw, po):
        """Enqueue one host-pointer batch (srt_separate_batch_async); returns the ticket for wait()."""
        ticket = C.c_int(-1)
        self._check(self.lib.srt_separate_batch_async(self.h, pl, pr, n, ns, uw, po, C.byref(ticket)))
        return ticket.value

    def wait(self, ticket):
        self._check(self.lib.srt_batch_wait(self.h, ticket))

    # ---- sample-rate conversion --------------------------------------------------------------
    def resample(self, x, ratio, table, index_inc=RESAMPLER_INDEX_INC):
        """x: float32[n] or [n][ch] interleaved frames -> (float32[ceil(n*ratio)][ch], frames generated): the
        reference's JamesDSPOfflineResampling (main.c:209-224) with the host's sinc table."""
        x = np.ascontiguousarray(x, np.float32)
        ch = 1 if x.ndim == 1 else x.shape[1]
        table = np.ascontiguousarray(table, np.float32)
        n_out = self.lib.srt_resample_frames(x.shape[0], float(ratio))
        out = np.empty((n_out, ch), np.float32)
        gen = C.c_size_t(0)
        self._check(self.lib.srt_resample_host(self.h, x.ctypes.data, x.shape[0], ch, float(ratio), table.ctypes.data, table.size,
                                               index_inc, out.ctypes.data, n_out, C.byref(gen)))
        return out, gen.value

    # ---- transforms -----------------------------------------------------------------------
    def stft(self, L, R):
        L = np.ascontiguousarray(L, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        rows = self.lib.srt_stft_rows(L.size)
        planes = [np.zeros((rows, FFTSIZE), np.float32) for _ in range(4)]
        self._check(self.lib.srt_stft_host(self.h, L.ctypes.data, R.ctypes.data, L.size, *[p.ctypes.data for p in planes]))
        return planes

    def istft(self, reL, imL, reR, imR):
        planes = [np.ascontiguousarray(a, np.float32) for a in (reL, imL, reR, imR)]
        frames = planes[0].shape[0]
        n = frames * HOPSIZE + FFTSIZE - HOPSIZE
        oL, oR = np.zeros(n, np.float32), np.zeros(n, np.float32)
        self._check(self.lib.srt_istft_host(self.h, *[p.ctypes.data for p in planes], frames, oL.ctypes.data, oR.ctypes.data))
        return oL, oR

    # ---- introspection ----------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.srt_launch_count(self.h))

    def set_timing(self, on=True):
        self.lib.srt_set_timing(self.h, int(on))

    def timing(self, name):
        ms = C.c_float()
        self.lib.srt_last_timing(self.h, TIMING_CATEGORIES[name], C.byref(ms))
        return ms.value

    def synchronize(self):
        self._check(self.lib.srt_synchronize(self.h))


class CliSeparator(Separator):
    """The CLI's output modes (Executable/main.c:776-970) as one device call (srt_create_cli).

    n_outputs = 2: nets = [vocal coeff]            -> separate() returns [vocal, accompaniment = input - vocal]
    n_outputs = 3: nets = [drum coeff, vocal coeff] -> [drum, vocal, accompaniment] (drum net, then the vocal net on
    the residual spectrum).  The activations are the ones main.c picks (drum: ELU, vocal: LeakyReLU/ReLU).
    separate(streams, unaffected=[w]) takes the single unaffectedWeight of main.c:773.
    """

    def __init__(self, coeffs, n_outputs, time_step, bin_limit, max_images=1, max_batch_images=0, device=0,
                 conv_impl=None, cuda_stream=None, precision=None):
        self.lib = load_library()
        if conv_impl is None:
            conv_impl = 1 if os.environ.get("SRT_CONV_IMPL", "") == "simt" else 0
        if len(coeffs) != n_outputs - 1:
            raise SrtError("n_outputs = 2 takes one net (vocal), n_outputs = 3 two (drum, vocal)")
        cfg = _Config(device, 1, time_step, bin_limit, max_images, max_batch_images, 0, conv_impl, cuda_stream, _precision(precision))
        self.T, self.F = time_step, bin_limit
        self.max_images = max_images
        blobs = [np.ascontiguousarray(c, np.float32) for c in coeffs]
        for c in blobs:
            if c.size != COEFF_FLOATS:
                raise SrtError("each net must be one spleeterCoeff blob of 9822725 floats")
        cp = (C.c_void_p * len(blobs))(*[c.ctypes.data for c in blobs])
        h = C.c_void_p()
        self._check(self.lib.srt_create_cli(C.byref(cfg), int(n_outputs), cp, C.byref(h)))
        self.h = h
        self.S = self.lib.srt_output_pairs(self.h)   # output pairs per stream

    def process_spleeter(self, x):
        raise SrtError("a CLI-mode context only offers the full path")


class Streamer:
    """Real-time streaming flavour: mirrors Spleeter4StemsInit / ProcessSamples / Free
    (VST/Source/Spleeter4Stems.h:67-69).  nets: list of coeff arrays (all ELU, as the VST does)."""

    def __init__(self, coeffs, time_step, bin_limit, device=0, unaffected=None, precision=None):
        self.lib = load_library()
        self.S = len(coeffs)
        self._coeffs = [np.ascontiguousarray(c, np.float32) for c in coeffs]
        cfg = _Config(device, self.S, time_step, bin_limit, 1, 1, 1, 0, None, _precision(precision))
        cp = (C.c_void_p * self.S)(*[c.ctypes.data for c in self._coeffs])
        uw = None
        if unaffected is not None:
            uw = (C.c_float * self.S)(*[float(u) for u in unaffected])
        h = C.c_void_p()
        rc = self.lib.srt_stream_create(C.byref(cfg), cp, uw, C.byref(h))
        if rc != 0:
            raise SrtError(f"srt error {rc}: {self.lib.srt_last_error().decode()}")
        self.h = h

    def process(self, L, R):
        """feed one block; returns float32[2S][n] (NaN where the call wrote nothing, as the reference leaves it untouched)"""
        L = np.ascontiguousarray(L, np.float32)
        R = np.ascontiguousarray(R, np.float32)
        out = np.full((2 * self.S, L.size), np.nan, np.float32)
        ptrs = (C.c_void_p * (2 * self.S))(*[out[j].ctypes.data for j in range(2 * self.S)])
        rc = self.lib.srt_stream_process(self.h, L.ctypes.data, R.ctypes.data, L.size, ptrs)
        if rc != 0:
            raise SrtError(f"srt error {rc}: {self.lib.srt_last_error().decode()}")
        return out

    def launch_count(self):
        return int(self.lib.srt_stream_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.srt_stream_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
