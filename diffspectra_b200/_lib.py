"""ctypes binding of libdiffspectra_b200.so (the C-ABI declared in include/diffspectra_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
``DiffSpectraError`` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdiffspectra_b200.so')

DT_F32, DT_BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_TANH, ACT_GELU = 0, 1, 2, 3
MODE_FP32, MODE_BF16 = 0, 1
SPECTRA_VERSIONS = {'uv': 0, 'ir': 1, 'raman': 2, 'allspectra': 3}
MODEL_KINDS = {'DMT': 0, 'DMT_WO_EQ': 1}


class DiffSpectraError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DiffSpectraError(
                'libdiffspectra_b200.so not built (%s); run `python -c "import __graft_entry__ as g; g.build()"`. '
                'There is no CPU / PyTorch fallback for the sampling hot path.' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ds_last_error.restype = ctypes.c_char_p
        _lib.ds_launch_count.restype = ctypes.c_longlong
        for name in ('ds_packed_weights_bytes', 'ds_plan_bytes', 'ds_workspace_bytes'):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = ctypes.c_size_t
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
