"""ctypes binding of libarflow_b200.so — the only door between the Python shim and the kernels.

There is no CPU fallback and no alternative backend: if the library is missing, or a tensor is not
a contiguous fp32 CUDA tensor, the call raises.  Non-zero return codes of the C-ABI become
RuntimeError (the reference raises AT_ERROR("CUDA call failed"), correlation_cuda.cc:81-83).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libarflow_b200.so")

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_P = c_void_p  # device pointers travel as void*

# name -> argument types; every function returns int unless listed in _RESTYPES.
PROTOTYPES = {
    "arf_version": [],
    "arf_error_string": [c_int],
    "arf_debug_set": [c_int, c_int],
    "arf_launch_count": [],
    "arf_corr_out_dims": [c_int] * 7 + [ctypes.POINTER(c_int)] * 3,
    "arf_corr_fwd": [_P, _P, _P] + [c_int] * 9 + [_P],
    "arf_corr_bwd": [_P, _P, _P, _P, _P] + [c_int] * 9 + [_P],
    "arf_warp_fwd": [_P, _P, _P] + [c_int] * 6 + [c_float, c_float] + [c_int] * 4 + [_P],
    "arf_warp_bwd": [_P, _P, _P, _P, _P] + [c_int] * 6 + [c_float, c_float] + [c_int] * 4 + [_P],
    "arf_inside_mask": [_P, _P] + [c_int] * 5 + [_P],
    "arf_range_map": [_P, _P] + [c_int] * 4 + [_P],
    "arf_range_map_bwd": [_P, _P, _P] + [c_int] * 4 + [_P],
    "arf_count_to_mask": [_P, _P, ctypes.c_longlong, c_int, c_float, _P],
    "arf_occ_bidir": [_P, _P, _P] + [c_int] * 3 + [c_float, c_float, _P],
    "arf_resize_bilinear_fwd": [_P, _P, ctypes.c_longlong] + [c_int] * 4 + [c_float] * 3 + [c_int, _P],
    "arf_resize_bilinear_bwd": [_P, _P, ctypes.c_longlong] + [c_int] * 4 + [c_float] * 3 + [c_int, _P],
    "arf_census_num_partials": [c_int] * 3,
    "arf_census_fwd": [_P] * 6 + [c_int] * 4 + [c_float] * 3 + [_P],
    "arf_census_bwd": [_P] * 9 + [c_int] * 4 + [c_float] * 3 + [_P],
    "arf_census_fwd_groups": [_P] * 6 + [c_int] * 5 + [c_float] * 3 + [_P],
    "arf_census_bwd_groups": [_P] * 9 + [c_int] * 5 + [c_float] * 3 + [_P],
    "arf_smooth_num_partials": [c_int] * 3,
    "arf_smooth_fwd": [_P] * 4 + [c_int] * 8 + [c_float] * 3 + [_P],
    "arf_smooth_bwd": [_P] * 4 + [c_int] * 8 + [c_float] * 3 + [_P],
    "arf_featnorm_workspace": [ctypes.c_longlong, ctypes.c_longlong],
    "arf_featnorm_fwd": [_P] * 6 + [ctypes.c_longlong, ctypes.c_longlong, _P],
    "arf_featnorm_bwd": [_P] * 9 + [ctypes.c_longlong, ctypes.c_longlong, _P],
    "arf_bias_leaky_num_partials": [ctypes.c_longlong, c_int, ctypes.c_longlong],
    "arf_bias_leaky_fwd": [_P, _P, ctypes.c_longlong, c_int, ctypes.c_longlong, c_float, _P],
    "arf_bias_leaky_bwd": [_P, _P, _P, _P, _P, ctypes.c_longlong, c_int, ctypes.c_longlong, c_float, _P],
    "arf_bias_leaky_nhwc_num_partials": [ctypes.c_longlong, c_int],
    "arf_bias_leaky_nhwc_fwd": [_P, _P, ctypes.c_longlong, c_int, c_float, _P],
    "arf_bias_leaky_nhwc_bwd": [_P, _P, _P, _P, _P, ctypes.c_longlong, c_int, c_float, _P],
    "arf_conv3x3s2_first_wgrad_workspace": [c_int, c_int, c_int],
    "arf_conv3x3s2_first_wgrad": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "arf_image_pair_pack": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_float, c_float, _P],
    "arf_conv3x3_small_fwd": [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "arf_conv3x3_small_bwd_workspace": [c_int, c_int, c_int, c_int, c_int],
    "arf_conv3x3_small_bwd": [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P],
    "arf_bias_leaky_nhwc_bwd_ld": [_P, ctypes.c_longlong, _P, ctypes.c_longlong, _P, _P, _P, ctypes.c_longlong, c_int, c_float, _P],
    "arf_bias_leaky_nhwc_fwd_ld": [_P, _P, ctypes.c_longlong, _P, ctypes.c_longlong, c_int, c_float, _P],
    "arf_nhwc_unpack_add": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, _P],
    "arf_pad_weight": [_P, _P] + [c_int] * 6 + [ctypes.c_longlong] * 4 + [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, _P],
    "arf_nhwc_transpose": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, _P],
    "arf_nhwc_pack_act": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_float, _P],
    "arf_nhwc_unpack_act": [_P, _P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_float, _P],
    "arf_nhwc_pack": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_int, _P],
    "arf_nhwc_unpack": [_P, _P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_int, _P],
    "arf_stencil_mv_fwd": [_P, _P, _P] + [c_int] * 5 + [_P],
    "arf_stencil_mv_bwd": [_P] * 5 + [c_int] * 5 + [_P],
    "arf_trisolve": [_P] * 6 + [ctypes.c_longlong, c_int, c_int, c_int, _P],
    "arf_inv_diag": [_P] * 4 + [ctypes.c_longlong, c_int, c_int, _P],
    "arf_ssim_fwd": [_P] * 4 + [ctypes.c_longlong] + [c_int] * 5 + [_P],
    "arf_ssim_bwd": [_P] * 7 + [ctypes.c_longlong] + [c_int] * 5 + [_P],
    "arf_resampler_fwd": [_P, _P, _P, ctypes.c_longlong, _P] + [c_int] * 4 + [ctypes.c_longlong, _P],
    "arf_comm_flag_bytes": [],
    "arf_comm_alloc": [ctypes.POINTER(c_void_p), c_size_t],
    "arf_comm_free": [_P],
    "arf_comm_ipc_get": [_P, ctypes.c_char_p],
    "arf_comm_ipc_open": [ctypes.c_char_p, ctypes.POINTER(c_void_p)],
    "arf_comm_ipc_close": [_P],
    "arf_allreduce_f32": [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), c_int, c_int, c_size_t, c_size_t, c_float, c_int, _P],
    "arf_comm_error": [_P, _P],
    "arf_resampler_bwd": [_P, _P, _P, ctypes.c_longlong, _P, _P, _P, _P, ctypes.c_longlong] + [c_int] * 4 + [ctypes.c_longlong, _P],
}
_RESTYPES = {"arf_error_string": ctypes.c_char_p, "arf_launch_count": ctypes.c_longlong,
             "arf_featnorm_workspace": ctypes.c_longlong, "arf_bias_leaky_num_partials": ctypes.c_longlong, "arf_bias_leaky_nhwc_num_partials": ctypes.c_longlong,
             "arf_conv3x3_small_bwd_workspace": ctypes.c_longlong,
             "arf_conv3x3s2_first_wgrad_workspace": ctypes.c_longlong}

_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "arflow_b200: %s is missing — run `python -m arflow_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def error_string(code):
    return load().arf_error_string(int(code)).decode()


def check(code, what):
    if code != 0:
        if code == -2:
            raise NotImplementedError("arflow_b200.%s: %s" % (what, error_string(code)))
        raise RuntimeError("arflow_b200.%s failed: %s (code %d)" % (what, error_string(code), code))


def dev_ptr(t, name="tensor", allow_none=False):
    """Raw device pointer of a contiguous fp32 CUDA tensor (no silent copies, no CPU path)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("arflow_b200: %s is None" % name)
    if not t.is_cuda:
        raise RuntimeError("arflow_b200: %s must be a CUDA tensor (no CPU fallback exists)" % name)
    if t.dtype != torch.float32:
        raise TypeError("arflow_b200: %s must be float32, got %s" % (name, t.dtype))
    if not t.is_contiguous():
        raise ValueError("arflow_b200: %s must be contiguous" % name)
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_profile = None   # when a list: (name, int args, start event, end event) per call, see profile_start


def profile_start():
    """Bracket every C-ABI call with CUDA events on the launching stream (bench.py's in-situ kernel
    timings; not for use under CUDA-graph capture)."""
    global _profile
    _profile = []


def profile_stop():
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    return [(n, a, s.elapsed_time(e)) for (n, a, s, e) in rec]


def launch_count():
    return int(load().arf_launch_count())


def call(name, *args):
    """Invoke a C-ABI entry point on the current stream and raise on failure."""
    fn = getattr(load(), name)
    if _profile is None:
        check(fn(*args), name)
        return
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    check(rc, name)
    _profile.append((name, args, s, e))
