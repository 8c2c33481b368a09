"""ctypes binding of libsg3_b200.so -- the only way the Python host reaches the kernels.

The library is built in-tree by build.py; if it is missing or fails to load, importing an op
fails loudly (there is no CPU or PyTorch fallback behind these ops).
Prototypes mirror include/sg3_b200.h one to one.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SG3_B200_LIB') or os.path.join(_HERE, 'libsg3_b200.so')      # the override is for tuning builds (build.py)

SG3_F32, SG3_F16, SG3_F64 = 0, 1, 2
SG3_E_INVALID, SG3_E_NOKERNEL, SG3_E_TOOLARGE = -1, -2, -3
SIGNS_NONE, SIGNS_WRITE, SIGNS_READ = 0, 1, 2
FLRELU_ROUND_TF32 = 1

c_i64x4 = ctypes.c_int64 * 4


class FlreluDesc(ctypes.Structure):
    """struct sg3_flrelu_desc (include/sg3_b200.h)."""
    _fields_ = [
        ('x', ctypes.c_void_p), ('y', ctypes.c_void_p), ('b', ctypes.c_void_p), ('signs', ctypes.c_void_p),
        ('fu', ctypes.c_void_p), ('fd', ctypes.c_void_p),
        ('N', ctypes.c_int32), ('C', ctypes.c_int32), ('inH', ctypes.c_int32), ('inW', ctypes.c_int32),
        ('outH', ctypes.c_int32), ('outW', ctypes.c_int32),
        ('xStride', c_i64x4), ('yStride', c_i64x4), ('bStride', ctypes.c_int64),
        ('up', ctypes.c_int32), ('down', ctypes.c_int32),
        ('fuW', ctypes.c_int32), ('fuH', ctypes.c_int32), ('fdW', ctypes.c_int32), ('fdH', ctypes.c_int32),
        ('px0', ctypes.c_int32), ('py0', ctypes.c_int32),
        ('gain', ctypes.c_float), ('slope', ctypes.c_float), ('clamp', ctypes.c_float),
        ('flip', ctypes.c_int32), ('signMode', ctypes.c_int32),
        ('sH', ctypes.c_int32), ('sWb', ctypes.c_int32), ('sx', ctypes.c_int32), ('sy', ctypes.c_int32),
        ('dtype', ctypes.c_int32), ('flags', ctypes.c_int32),
        ('ysum', ctypes.c_void_p),
    ]


_lib = None

_I, _L, _F, _P = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
_IP = ctypes.POINTER(ctypes.c_int)

_PROTOS = {
    'sg3_abi_version': (ctypes.c_int, []),
    'sg3_error_string': (ctypes.c_char_p, [_I]),
    'sg3_build_info': (ctypes.c_char_p, []),
    'sg3_launch_count': (ctypes.c_ulonglong, []),
    'sg3_filtered_lrelu_shape': (_I, [_I] * 12 + [_IP] * 4),
    'sg3_filtered_lrelu_supported': (_I, [_I] * 6),
    'sg3_filtered_lrelu': (_I, [ctypes.POINTER(FlreluDesc), _P]),
    'sg3_sizeof_flrelu_desc': (ctypes.c_int, []),
    'sg3_filtered_lrelu_act': (_I, [_P, _P, _I, _I, _I, _I, ctypes.POINTER(c_i64x4), _I, _I, _I, _I, _F, _F, _F, _I, _I, _P]),
    'sg3_bias_act': (_I, [_P, _P, _P, _P, _P, _P, _L, ctypes.c_int32, _L, _I, _I, _F, _F, _F, _I, _P]),
    'sg3_upfirdn2d': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, ctypes.POINTER(c_i64x4), ctypes.POINTER(c_i64x4),
                           _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    'sg3_upfirdn2d_sep': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, ctypes.POINTER(c_i64x4), ctypes.POINTER(c_i64x4),
                                _I, _I, _I, _I, _I, _I, _I, _F, _I, _P]),
    'sg3_modconv_weights': (_I, [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_weights_bwd': (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_weights_bwd_taps': (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_tc_supported': (_I, [_I, _I, _I, _I, _I, _I]),
    'sg3_modconv_wgrad': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_wgrad3': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_fwd': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    'sg3_modconv_set_smem_budget': (_I, [_I]),
    'sg3_modconv_fwd_pitched': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
}

EXPORTS = tuple(_PROTOS)


def lib():
    """Load libsg3_b200.so once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build the sm_100a kernels first '
                f'(python stylegan3-editing_b200/build.py, or __graft_entry__.build()). '
                f'There is no CPU / PyTorch fallback for these ops.')
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.sg3_abi_version() != 1 or L.sg3_sizeof_flrelu_desc() != ctypes.sizeof(FlreluDesc):
            raise RuntimeError('libsg3_b200.so ABI mismatch (version or sg3_flrelu_desc layout)')
        _lib = L
    return _lib


class Sg3Error(RuntimeError):
    pass


def check(code, what):
    if code != 0:
        msg = lib().sg3_error_string(code).decode()
        raise Sg3Error(f'{what} failed: {msg} (code {code})')


def dtype_code(dtype):
    import torch
    try:
        return {torch.float32: SG3_F32, torch.float16: SG3_F16, torch.float64: SG3_F64}[dtype]
    except KeyError:
        raise TypeError(f'sg3_b200 ops support float16/float32/float64 tensors, got {dtype}')


def stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, opname):
    if t.device.type != 'cuda':
        raise RuntimeError(f'{opname}: sg3_b200 kernels run on CUDA tensors only (got {t.device}); '
                           f'there is no CPU fallback by design')
