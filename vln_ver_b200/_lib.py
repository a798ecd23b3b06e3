"""ctypes binding of libver_b200.so (the C ABI declared in include/ver_b200.h).

The library is the product: there is NO fallback.  If it cannot be loaded the
import raises, and every op raises if it is handed a non-CUDA tensor.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libver_b200.so')

VER_F32, VER_F16 = 0, 1
ABI_VERSION = 7


class VerError(RuntimeError):
    pass


def _load():
    # built in-tree by `python -m vln_ver_b200.build` / __graft_entry__.build(); build() is a no-op
    # when the source digest matches the stamp written next to the .so
    from . import build as _build
    try:
        _build.build()
    except Exception as e:  # noqa: BLE001
        if not os.path.exists(LIB_PATH):
            raise VerError(
                f'libver_b200.so is missing and could not be built ({e}); '
                'run `python -m vln_ver_b200.build`') from e
        # a library exists but could not be rebuilt (no nvcc on this machine, or a compile error): it may be older than
        # the sources -- say so instead of loading it silently; a changed interface still fails below (missing symbol /
        # ABI version), a changed kernel body would not
        import warnings
        warnings.warn(f'libver_b200.so could not be rebuilt ({type(e).__name__}: {str(e)[:200]}); loading the existing '
                      f'binary, which may not match the sources', RuntimeWarning, stacklevel=2)
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    sig = {
        'ver_abi_version': (c_int, []),
        'ver_last_error': (c_char_p, []),
        'ver_launch_count': (c_int64, []),
        'ver_point_sampling_f32': (c_int, [P, P, ctypes.POINTER(c_double), c_int, c_int, c_int, c_int,
                                           c_int, c_float, c_float, P, P, P, P, P]),
        'ver_visible_index': (c_int, [P, c_int, c_int, c_int, P, P, P]),
        'ver_msda_forward': (c_int, [c_int, P, ctypes.POINTER(c_int32), c_int, P, P, P, c_int, c_int,
                                     c_int, c_int, c_int, c_int, P]),
        'ver_msda_backward': (c_int, [c_int, P, ctypes.POINTER(c_int32), c_int, P, P, P, P, P, P, c_int,
                                      c_int, c_int, c_int, c_int, c_int, P]),
        'ver_msda3d_forward': (c_int, [c_int, P, ctypes.POINTER(c_int32), c_int, P, P, P, c_int, c_int,
                                       c_int, c_int, c_int, c_int, P]),
        'ver_msda3d_backward': (c_int, [c_int, P, ctypes.POINTER(c_int32), c_int, P, P, P, P, P, P, c_int,
                                        c_int, c_int, c_int, c_int, c_int, P]),
        'ver_convt_col2im': (c_int, [c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'ver_convt_im2col': (c_int, [c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
        'ver_sca_forward': (c_int, [c_int, P, c_int, P, c_int, P, P, P] + [c_int] * 10 + [P]),
        'ver_sca_backward': (c_int, [c_int, P, c_int, P, c_int, P, P, P, P, P, P, P] + [c_int] * 10 + [P]),
        'ver_visibility_order_workspace': (c_int, [c_int, c_int, ctypes.POINTER(ctypes.c_size_t)]),
        'ver_visibility_order': (c_int, [P, c_int, c_int, P, P, P, P, ctypes.c_size_t, P]),
        'ver_sca_forward_sorted': (c_int, [P, P, c_int, P, P, P, P, P] + [c_int] * 9 + [P]),
        'ver_value_image_f16': (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
        'ver_value_image16_f16': (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
        'ver_tc6_supported': (c_int, [c_int] * 5),
        'ver_sca_forward_sorted16': (c_int, [P, P, c_int, P, P, P, P, P] + [c_int] * 9 + [P]),
        'ver_feat_embed': (c_int, [c_int, P, P, P, P, c_int, c_int, c_int, c_int, P]),
        'ver_add_layernorm': (c_int, [c_int, P, P, P, P, P, c_int64, c_int, c_float, P]),
        'ver_dropout_add_layernorm_fwd': (c_int, [c_int, P, P, P, P, P, P, P, c_int64, c_int, c_float, c_float,
                                                  ctypes.c_uint64, P, P]),
        'ver_dropout_add_layernorm_bwd_blocks': (c_int, [c_int64]),
        'ver_dropout_add_layernorm_fwd_bits': (c_int, [c_int, P, P, P, P, P, P, P, P, c_int64, c_int, c_float, c_float,
                                                       ctypes.c_uint64, P, P]),
        'ver_dropout_add_layernorm_bwd_bits': (c_int, [c_int, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_float,
                                                       ctypes.c_uint64, P, P]),
        'ver_dropout_add_layernorm_bwd': (c_int, [c_int, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_float,
                                                  ctypes.c_uint64, P, P]),
        'ver_relu_dropout_fwd': (c_int, [c_int, P, P, c_int64, c_float, ctypes.c_uint64, P, P]),
        'ver_linear_supported': (c_int, [c_int, c_int, c_int]),
        'ver_linear_f16': (c_int, [c_int, P, c_int, P, c_int, P, P, c_int, c_int, c_int, c_int, c_float,
                                   ctypes.c_uint64, P, P]),
        'ver_linear_bwd_colsum_rows': (c_int, [c_int]),
        'ver_linear_relu_dropout_bwd_f16': (c_int, [P, c_int, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_float, P]),
        'ver_colsum_partial_rows': (c_int, []),
        'ver_relu_dropout_bwd': (c_int, [c_int, P, P, P, c_int64, c_float, c_int, P, P]),
        'ver_cast_colsum': (c_int, [c_int, P, P, c_int64, c_int, P, P]),
        'ver_colsum_f16': (c_int, [P, c_int64, c_int, P, P]),
        'ver_colsum_fold_scratch_floats': (c_int, [c_int]),
        'ver_colsum_fold': (c_int, [P, c_int64, c_int, P, P, P]),
        'ver_colsum_fold_batched': (c_int, [P, c_int, c_int64, c_int, P, P]),
        'ver_focal_loss': (c_int, [P, P, c_int, P, P, P, P, c_int64, c_int, c_float, c_float, P]),
        'ver_occupancy_decode': (c_int, [P, c_int64, c_int, c_float, P, P, P, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.ver_abi_version() != ABI_VERSION:
        raise VerError(f'libver_b200.so ABI {lib.ver_abi_version()} != expected {ABI_VERSION}')
    return lib, tuple(sig)


lib, EXPORTED = _load()


def check(rc):
    if rc != 0:
        raise VerError(f'libver_b200 error {rc}: {lib.ver_last_error().decode()}')


def launch_count():
    return int(lib.ver_launch_count())
