"""ctypes binding of libadil_b200.so -- the C ABI declared in include/adil_b200.h.

There is no CPU fallback: if the library is missing or a tensor is not on a CUDA device the call raises.
"""
import ctypes
import os

from .build import LIB_PATH

c_float_p = ctypes.POINTER(ctypes.c_float)


class AdamwParams(ctypes.Structure):
    """struct adil_adamw (include/adil_b200.h)."""
    _fields_ = [("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
                ("eps", ctypes.c_double), ("weight_decay", ctypes.c_double), ("step", ctypes.c_longlong)]


# name -> (restype, argtypes); every symbol include/adil_b200.h declares
_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_LL = ctypes.c_longlong
_SZ = ctypes.c_size_t
_HP = ctypes.POINTER(AdamwParams)
_IP = ctypes.POINTER(ctypes.c_int)
_PP = ctypes.POINTER(ctypes.c_void_p)
SIGNATURES = {
    "adil_version": (_I, []),
    "adil_last_error": (ctypes.c_char_p, []),
    "adil_device_info": (_I, [ctypes.POINTER(_I)] * 3),
    "adil_l2_persist": (_I, [_P, _SZ, _P]),
    "adil_set_impl": (_I, [_I]),
    "adil_get_impl": (_I, []),
    "adil_tc_supported": (_I, [_I, _I, _I]),
    "adil_synth": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, c_float_p, c_float_p, _F, _I, _P]),
    "adil_grad_scratch_bytes": (_SZ, [_I, _I]),
    "adil_grad": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, c_float_p, _P, _F, _I, _IP, _P, _SZ, _P]),
    "adil_grad_max_batch": (_I, [_I, _I, _I, _I]),
    "adil_grad_dict_step": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, c_float_p, _HP, _I, _I, _IP, _P, _SZ,
                                 _P]),
    "adil_dict_step": (_I, [_P, _P, _P, _P, _LL, _HP, _I, _P]),
    "adil_dict_step_atoms": (_I, [_P, _P, _P, _P, _I, _I, _HP, _F, _I, _P, _P]),
    "adil_dict_step_peer": (_I, [_PP, _PP, _P, _P, _LL, _LL, _I, _I, _HP, _I, _P, _P, _P]),
    "adil_code_prox_step": (_I, [_P, _P, _P, _I, _I, _I, _F, _I, _F, _P]),
    "adil_code_step": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _HP, _I, _F, _P, _I, _P]),
    "adil_project_rows": (_I, [_P, _I, _I, _I, _F, _P]),
    "adil_project_atoms_scratch_bytes": (_SZ, [_I]),
    "adil_project_atoms": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "adil_adamw_clamp": (_I, [_P, _P, _P, _P, _LL, _HP, _F, _P]),
    "adil_image_errors_scratch_bytes": (_SZ, [_I]),
    "adil_image_errors": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libadil_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` or `python -m dl_attack_on_imagenet_b200.build`; there is no CPU "
                               "fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().adil_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))
