"""ctypes binding of libdiffspectra_b200.so (the C-ABI declared in include/diffspectra_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
``DiffSpectraError`` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdiffspectra_b200.so')

DT_F32, DT_BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_TANH, ACT_GELU, ACT_SILU_HALF, ACT_TANH_MIX = 0, 1, 2, 3, 4, 5
MODE_FP32, MODE_BF16 = 0, 1
SPECTRA_VERSIONS = {'uv': 0, 'ir': 1, 'raman': 2, 'allspectra': 3}
MODEL_KINDS = {'DMT': 0, 'DMT_WO_EQ': 1}


class DiffSpectraError(RuntimeError):
    pass


_lib = None

_P, _I, _Z, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
_U64, _I64 = ctypes.c_ulonglong, ctypes.c_longlong
_PLAN = [_P, _I, _I, _I, _I]            # plan blob, B, N, Mn, Mp
# name -> (restype, argtypes): one entry per symbol of include/diffspectra_b200.h (tests/test_host_logic.py checks that
# every declared symbol is exported and listed here)
SIGNATURES = {
    'ds_last_error': (ctypes.c_char_p, []),
    'ds_version': (_I, []),
    'ds_create': (_I, [ctypes.POINTER(_P), _I, _I, _I]),
    'ds_create_model': (_I, [ctypes.POINTER(_P), _I, _I, _I, _I]),
    'ds_destroy': (_I, [_P]),
    'ds_launch_count': (_I64, [_P]),
    'ds_profile_begin': (_I, []),
    'ds_profile_end': (_I, [ctypes.c_char_p, _Z]),
    'ds_packed_weights_bytes': (_Z, [_P]),
    'ds_pack_weights': (_I, [_P, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(_P), _I, _P, _Z, _P]),
    'ds_plan_bytes': (_Z, [_I, _I]),
    'ds_plan_build': (_I, [_P, _P, _I, _I, _P, ctypes.POINTER(_I), ctypes.POINTER(_I), _P]),
    'ds_plan_build_host': (_I, [_P, _I, _I, _P, _Z, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    'ds_plan_layout': (_I, [_I, _I, ctypes.POINTER(_Z), _I]),
    'ds_workspace_bytes': (_Z, [_P, _I, _I, _I]),
    'ds_specformer_workspace_bytes': (_Z, [_P, _I]),
    'ds_specformer_ctx': (_I, [_P, _P, _P, _P, _I, _P, _P, _Z, _P]),
    'ds_denoise': (_I, [_P] + _PLAN + [_P] * 8 + [_P, _Z, _P]),
    'ds_sample_loop': (_I, [_P] + _PLAN + [_P, _P, _P, _P, _I, _I, _P, _P, _P, _U64, _I64, _F, _I, _P, _P, _P, _Z, _P]),
    'ds_sampler_step': (_I, [_P] + _PLAN + [_P, _P, _P, _P, _P, _P, _P, _P, _U64, _I64, _I, _F, _P, _P, _P, _Z, _P]),
    'ds_post_process': (_I, [_P] + _PLAN + [_P] * 6 + [_P, _Z, _P]),
    'ds_record_bytes': (_Z, [_I]),
    'ds_molecule_records': (_I, [_P] + _PLAN + [_P, _P, _I, _P, _Z, _P]),
    'ds_umma2_probe': (_I, [_P, _P, _P, _P, _I, _P]),
    'ds_coord_head': (_I, [_P] + _PLAN + [_P] * 10 + [_P]),
    'ds_gemm': (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    'ds_gemm_fused': (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P, _I, _P, _I, _I, _P, _I, _P, _I, _P, _I, _P, _P, _P, _P]),
}


def _declare(handle):
    for name, (res, args) in SIGNATURES.items():
        if hasattr(handle, name):
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DiffSpectraError(
                'libdiffspectra_b200.so not built (%s); run `python -c "import __graft_entry__ as g; g.build()"`. '
                'There is no CPU / PyTorch fallback for the sampling hot path.' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc, what):
    if rc != 0:
        raise DiffSpectraError('%s failed (%d): %s' % (what, rc, lib().ds_last_error().decode()))


def ptr(t):
    """Device/host pointer of a tensor (or None) as a ctypes void*."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
