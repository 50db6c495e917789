"""spleeterrt_b200 — B200-native implementation of SpleeterRT's spectrogram-to-mask hot path.

This package is a thin ctypes mirror of the C ABI in ``include/srt_b200.h`` (tier B) and of the
reference's own operator interface (``processSpleeter`` / ``stft`` / ``istft``, tier A).  All
compute happens in ``libspleeterrt_b200.so`` (hand-written sm_100a CUDA); there is no CPU or
PyTorch fallback — if the library or a B200 is missing, calls raise.
"""
from .api import (COEFF_FLOATS, FFTSIZE, HOPSIZE, BINS, Separator, CliSeparator, Streamer, SrtError, half_to_float,  # noqa: F401
                  load_coeff_dat, save_coeff_dat, load_model_fp16, pack_layer, resample_plan,
                  lib_path, load_library, exported_symbols, HEADER_SYMBOLS, DISPATCH_SYMBOLS, dispatch_lib_path, load_dispatch_library,
                  dispatch_schedule, NcclDispatcher, probe_tensor_peak, probe_copy_bandwidth, pinned_empty)
