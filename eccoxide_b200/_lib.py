"""ctypes binding of libeccbatch.so (the C ABI in include/eccbatch.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present, importing
works but creating a Context raises.  The library is built in-tree by __graft_entry__.build()
(make -C eccoxide_b200/csrc) and loaded from eccoxide_b200/libeccbatch.so.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeccbatch.so")

ECB_OK = 0
ECB_ERR_CUDA = -1
ECB_ERR_INVALID_ARG = -2
ECB_ERR_NONCANONICAL_SCALAR = -3
ECB_ERR_POINT_NOT_ON_CURVE = -4
ECB_ERR_OOM = -5
ECB_ERR_NOT_READY = -6

CURVE_P256R1 = 0
CURVE_P384R1 = 1
CURVE_BLS12_381_G1 = 2
CURVE_P256K1 = 3

_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_szp = ctypes.POINTER(ctypes.c_size_t)
_int = ctypes.c_int

# name -> (restype, argtypes); every symbol include/eccbatch.h declares
SIGNATURES = {
    "ecb_init": (_int, [ctypes.POINTER(_int), _int, ctypes.POINTER(_vp)]),
    "ecb_destroy": (None, [_vp]),
    "ecb_last_error": (ctypes.c_char_p, [_vp]),
    "ecb_device_count": (_int, [_vp]),
    "ecb_set_option": (_int, [_vp, ctypes.c_char_p, ctypes.c_long]),
    "ecb_launch_count": (ctypes.c_ulonglong, [_vp]),
    "ecb_alloc_pinned": (_vp, [_sz]),
    "ecb_free_pinned": (None, [_vp]),
    "ecb_ed25519_mul_base": (_int, [_vp, _vp, _sz, _vp, _szp]),
    "ecb_ed25519_mul_base_compressed": (_int, [_vp, _vp, _sz, _vp, _szp]),
    "ecb_ed25519_mul": (_int, [_vp, _vp, _vp, _sz, _vp, _szp]),
    "ecb_ed25519_verify_prehashed": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ecb_ed25519_verify": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ecb_x25519": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ecb_x25519_base": (_int, [_vp, _vp, _sz, _vp]),
    "ecb_x25519_base_dev": (_int, [_vp, _int, _vp, _sz, _vp, _vp]),
    "ecb_x448": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ecb_wei_mul": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _vp, _vp, _szp]),
    "ecb_wei_mul_base": (_int, [_vp, _int, _vp, _sz, _vp, _vp, _szp]),
    "ecb_ecdsa_verify_hashed": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _vp, _szp]),
    "ecb_ecdsa_verify": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp, _szp]),
    "ecb_ed25519_mul_base_dev": (_int, [_vp, _int, _vp, _sz, _vp, _vp]),
    "ecb_ed25519_mul_dev": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp]),
    "ecb_x25519_dev": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp]),
    "ecb_wei_mul_dev": (_int, [_vp, _int, _int, _vp, _vp, _sz, _vp, _vp, _vp]),
    "ecb_wei_mul_base_dev": (_int, [_vp, _int, _int, _vp, _sz, _vp, _vp, _vp]),
    "ecb_x448_dev": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ed25519_verify_prehashed_dev": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ecdsa_verify_hashed_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_profile_collect": (_int, [_vp, _int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_int)]),
    "ecb_dev_status": (_int, [_vp, _int, _szp]),
    "ecb_imad_probe": (_int, [_vp, _int, _int, _int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "ecb_latency_probe": (_int, [_vp, _int, _int, _int, _int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "ecb_fieldmul_probe": (_int, [_vp, _int, _int, _int, _int, _int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "ecb_wei_decompress": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp]),
    "ecb_bls12_381_g1_from_compressed": (_int, [_vp, _vp, _sz, _int, _vp, _vp]),
    "ecb_bls12_381_g1_to_compressed": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ecb_bls12_381_g1_from_uncompressed": (_int, [_vp, _vp, _sz, _int, _vp, _vp, _vp]),
    "ecb_bls12_381_g1_to_uncompressed": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "ecb_ed25519_decompress": (_int, [_vp, _vp, _sz, _vp, _vp]),
    "ecb_wei_msm": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp, _szp]),
    "ecb_ristretto255_decompress": (_int, [_vp, _vp, _sz, _vp, _vp]),
    "ecb_ristretto255_compress": (_int, [_vp, _vp, _sz, _vp]),
    "ecb_ristretto255_mul": (_int, [_vp, _vp, _vp, _sz, _vp, _szp]),
    "ecb_ristretto255_mul_base": (_int, [_vp, _vp, _sz, _vp, _szp]),
    "ecb_wei_decompress_dev": (_int, [_vp, _int, _int, _vp, _vp, _sz, _vp, _vp, _vp]),
    "ecb_bls12_381_g1_from_compressed_dev": (_int, [_vp, _int, _vp, _sz, _int, _vp, _vp, _vp]),
    "ecb_ed25519_public_from_seed": (_int, [_vp, _vp, _sz, _vp]),
    "ecb_ed25519_sign": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ecb_ed25519_public_from_seed_dev": (_int, [_vp, _int, _vp, _sz, _vp, _vp]),
    "ecb_ed25519_public_from_seed_vartime": (_int, [_vp, _vp, _sz, _vp]),
    "ecb_ed25519_sign_vartime": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ecb_ed25519_public_from_seed_vartime_dev": (_int, [_vp, _int, _vp, _sz, _vp, _vp]),
    "ecb_ed25519_sign_vartime_dev": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ed25519_mul_base_ct": (_int, [_vp, _vp, _sz, _vp, _szp]),
    "ecb_ed25519_sign_dev": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ecdsa_sign_hashed": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ecdsa_sign_hashed_vartime": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ecdsa_sign_hashed_vartime_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "ecb_ecdsa_sign_vartime": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_ecdsa_sign_hashed_dev": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "ecb_ecdsa_sign": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "ecb_warm": (_int, [_vp, ctypes.c_char_p, _int, _sz]),
    "ecb_get_info": (_int, [_vp, _int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long)]),
    "ecb_debug_fused_trace": (ctypes.c_long, [_vp, _int, _vp, _sz]),
    "ecb_debug_chunk_plan": (ctypes.c_long, [_sz, _sz, _sz, ctypes.c_long, _vp, _sz]),
    "ecb_debug_ed25519_table": (ctypes.c_long, [_vp, _int, _vp, _sz, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
}

_lib = None


class EccBatchError(RuntimeError):
    def __init__(self, code, msg, bad_index=None):
        super().__init__("libeccbatch error %d: %s%s" % (code, msg, "" if bad_index is None else " (element %d)" % bad_index))
        self.code = code
        self.bad_index = bad_index


def load():
    """Load libeccbatch.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EccBatchError(ECB_ERR_CUDA, "%s not built (run __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
