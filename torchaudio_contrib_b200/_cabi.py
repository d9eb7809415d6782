"""ctypes binding of libtac_b200.so (include/tac_b200.h).  No torch extension: tensors cross the
boundary as raw device pointers and the current CUDA stream handle.

The library is REQUIRED: importing the package without it, or calling an operator on a machine
without a CUDA device, raises -- there is no CPU or eager-torch fallback anywhere in this package.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# TAC_B200_LIB: an alternative build of the same library (e.g. the -DTAC_K1_TRACE_BUILD timing build of scripts/)
LIB_PATH = os.environ.get("TAC_B200_LIB") or os.path.join(_PKG, "lib", "libtac_b200.so")

TAC_OK = 0
TAC_ERR_INVALID = -1
TAC_ERR_UNSUPPORTED = -2
TAC_ERR_CUDA = -3
TAC_ERR_WORKSPACE = -4

PAD_MODES = {"reflect": 0, "constant": 1, "replicate": 2, "circular": 3}

_c = ctypes
_i64, _int, _f32, _ptr = _c.c_int64, _c.c_int, _c.c_float, _c.c_void_p

# name -> (restype, argtypes); mirrors include/tac_b200.h one to one
_STFT_ARGS = [_ptr, _i64, _i64, _i64, _ptr, _int, _int, _int, _int, _int]
SIGNATURES = {
    "tac_version": (_int, []),
    "tac_last_error": (_c.c_char_p, []),
    "tac_device_info": (_int, [_c.POINTER(_int)] * 3),
    "tac_stft_num_frames": (_i64, [_i64, _int, _int, _int]),
    "tac_stft_f32": (_int, _STFT_ARGS + [_int, _ptr, _ptr]),
    "tac_spectrogram_f32": (_int, _STFT_ARGS + [_int, _f32, _ptr, _ptr]),
    "tac_complex_norm_f32": (_int, [_ptr, _i64, _f32, _ptr, _ptr]),
    "tac_amplitude_to_db_f32": (_int, [_ptr, _i64, _f32, _f32, _ptr, _ptr]),
    "tac_db_to_amplitude_f32": (_int, [_ptr, _i64, _f32, _ptr, _ptr]),
    "tac_magphase_f32": (_int, [_ptr, _i64, _f32, _ptr, _ptr, _ptr]),
    "tac_phase_vocoder_f32": (_int, [_ptr, _i64, _int, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "tac_phase_vocoder_f64": (_int, [_ptr, _i64, _int, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "tac_phase_vocoder_backward_f32": (_int, [_ptr, _ptr, _i64, _int, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "tac_phase_vocoder_backward_f64": (_int, [_ptr, _ptr, _i64, _int, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr]),
    "tac_fbplan_bytes": (_i64, [_int, _int]),
    "tac_fbplan_build_host": (_int, [_ptr, _int, _int, _ptr, _i64, _c.POINTER(_i64)]),
    "tac_fbplan_band_handle": (_i64, [_ptr]),
    "tac_fbplan_fused_handle": (_i64, [_ptr, _int]),
    "tac_power_mel_f32": (_int, [_ptr, _int, _f32, _i64, _i64, _int, _ptr, _int, _int, _f32, _f32, _ptr, _ptr]),
    "tac_melspec_workspace_bytes": (_i64, [_i64, _i64, _int, _int, _int]),
    "tac_melspec_f32": (_int, _STFT_ARGS + [_f32, _ptr, _int, _int, _f32, _f32, _ptr, _i64, _ptr, _ptr]),
    "tac_melspec_banded_f32": (_int, _STFT_ARGS + [_f32, _ptr, _i64, _int, _int, _f32, _f32, _ptr, _int, _ptr]),
    "tac_melspec_banded_peers_f32": (_int, _STFT_ARGS + [_f32, _ptr, _i64, _int, _int, _f32, _f32, _ptr, _int, _i64, _int, _ptr]),
    "tac_peer_alloc": (_int, [_i64, _c.POINTER(_ptr), _ptr]),
    "tac_peer_open": (_int, [_ptr, _c.POINTER(_ptr)]),
    "tac_peer_close": (_int, [_ptr]),
    "tac_peer_free": (_int, [_ptr]),
    "tac_peer_barrier": (_int, [_ptr, _int, _int, _c.c_uint32, _c.c_double, _ptr]),
    "tac_peer_timed_out": (_int, [_ptr, _c.POINTER(_int)]),
    "tac_mc_supported": (_int, [_c.POINTER(_int)]),
    "tac_mc_create": (_int, [_i64, _int, _c.POINTER(_ptr), _c.POINTER(_int)]),
    "tac_mc_import": (_int, [_int, _i64, _int, _c.POINTER(_ptr)]),
    "tac_mc_add_device": (_int, [_ptr]),
    "tac_mc_bind": (_int, [_ptr, _c.POINTER(_ptr), _c.POINTER(_ptr)]),
    "tac_mc_barrier": (_int, [_ptr, _int, _c.c_uint32, _c.c_double, _ptr]),
    "tac_mc_timed_out": (_int, [_ptr, _c.POINTER(_int)]),
    "tac_mc_free": (_int, [_ptr]),
    "tac_melspec_banded_mc_f32": (_int, _STFT_ARGS + [_f32, _ptr, _i64, _int, _int, _f32, _f32, _ptr, _i64, _int, _ptr]),
    "tac_stft_backward_workspace_bytes": (_i64, [_i64, _i64, _int, _int, _int]),
    "tac_stft_backward_f32": (_int, [_ptr, _i64, _i64, _ptr, _int, _int, _int, _int, _int, _int, _ptr, _ptr, _i64, _ptr]),
    "tac_window_grad_f32": (_int, [_ptr, _i64, _i64, _i64, _ptr, _int, _int, _int, _int, _ptr, _ptr, _i64, _ptr]),
    "tac_spectrogram_backward_f32": (_int, _STFT_ARGS + [_int, _f32, _ptr, _ptr, _ptr, _i64, _ptr]),
    "tac_filterbank_backward_f32": (_int, [_ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _int, _int, _ptr, _ptr]),
    "tac_melspec_backward_workspace_bytes": (_i64, [_i64, _i64, _int, _int, _int]),
    "tac_melspec_backward_f32": (_int, _STFT_ARGS + [_f32, _ptr, _int, _ptr, _i64, _i64, _i64, _ptr, _ptr, _i64, _ptr]),
    "tac_amplitude_to_db_backward_f32": (_int, [_ptr, _ptr, _i64, _f32, _ptr, _ptr]),
    "tac_complex_norm_backward_f32": (_int, [_ptr, _ptr, _i64, _f32, _ptr, _ptr]),
    "tac_mulaw_encode_f32_i64": (_int, [_ptr, _i64, _int, _ptr, _int, _int, _f32, _ptr, _ptr]),
    "tac_mulaw_tables_host": (_int, [_int, _ptr, _int, _c.POINTER(_int), _c.POINTER(_int), _c.POINTER(_f32), _ptr, _c.POINTER(_int)]),
    "tac_mulaw_decode_i64_f32": (_int, [_ptr, _i64, _int, _ptr, _ptr, _ptr]),
    "tac_mulaw_decode_f32_f32": (_int, [_ptr, _i64, _int, _ptr, _ptr, _ptr]),
    "tac_stft_f64": (_int, [_ptr, _i64, _i64, _i64, _ptr, _int, _int, _int, _int, _int, _int, _ptr, _ptr]),
    "tac_complex_norm_f64": (_int, [_ptr, _i64, _c.c_double, _ptr, _ptr]),
    "tac_magphase_f64": (_int, [_ptr, _i64, _c.c_double, _ptr, _ptr, _ptr]),
    "tac_amplitude_to_db_f64": (_int, [_ptr, _i64, _c.c_double, _c.c_double, _ptr, _ptr]),
    "tac_db_to_amplitude_f64": (_int, [_ptr, _i64, _c.c_double, _ptr, _ptr]),
    "tac_apply_filterbank_f64": (_int, [_ptr, _ptr, _i64, _i64, _int, _int, _ptr, _ptr]),
    "tac_mulaw_decode_i64_f64": (_int, [_ptr, _i64, _int, _ptr, _ptr, _ptr]),
    "tac_mulaw_encode_f64_i64": (_int, [_ptr, _i64, _int, _ptr, _ptr]),
    "tac_hpss_f32": (_int, [_ptr, _i64, _int, _int, _int, _f32, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "tac_pipeline_create": (_int, [_ptr, _ptr, _ptr, _c.POINTER(_ptr)]),
    "tac_pipeline_run_host": (_int, [_ptr, _ptr, _i64, _i64, _ptr]),
    "tac_pipeline_destroy": (_int, [_ptr]),
    "tac_launch_count": (_i64, []),
    "tac_profile_enable": (_int, [_int]),
    "tac_profile_read": (_int, [_c.POINTER(_c.c_double), _c.POINTER(_i64)]),
    "tac_mel_kernel_variant": (_int, [_int]),
    "tac_pointwise_backward_f32": (_int, [_int, _ptr, _ptr, _ptr, _i64, _f32, _ptr, _ptr]),
}


class PipelineConfig(ctypes.Structure):
    """struct tac_pipeline_config"""
    _fields_ = [("n_fft", _int), ("hop", _int), ("center", _int), ("pad_mode", _int), ("normalized", _int),
                ("power", _f32), ("n_bins", _int), ("n_bands", _int), ("to_db", _int), ("ref", _f32), ("amin", _f32)]


class TacError(RuntimeError):
    """Non-zero status from the C ABI (message from tac_last_error())."""

    def __init__(self, code, message):
        super().__init__("libtac_b200: %s (status %d)" % (message, code))
        self.code = code


_lib = None


def load(path=None):
    """dlopen the library and declare every prototype of include/tac_b200.h.
    TAC_B200_LIB names another build of the same library (kernel experiments, scripts/build_variant.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if path is None:
        path = os.environ.get("TAC_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            "torchaudio_contrib_b200 needs its CUDA library %s (build it with "
            "`python build_native.py`); there is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def lib():
    return _lib if _lib is not None else load()


def check(status):
    """Map a C-ABI status to the exception the reference's torch call would have raised."""
    if status == TAC_OK:
        return
    msg = lib().tac_last_error().decode("utf-8", "replace")
    if status == TAC_ERR_UNSUPPORTED:
        raise NotImplementedError("libtac_b200: " + msg)
    raise TacError(status, msg)


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (what, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError(
            "%s is on %s: the B200 kernels need a CUDA tensor (no CPU fallback exists; use "
            "HostPipeline for host buffers)" % (what, t.device))


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())
