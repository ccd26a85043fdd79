"""coracle.py — ctypes access to oracle/ecc_oracle.c (the C restatement of the reference's CPU path).

TEST INFRASTRUCTURE / CPU BASELINE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs, never by eccoxide_b200/.  build() compiles it in place with
gcc (oracle/Makefile) into oracle/_build/ (git-ignored, travels to the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libecc_oracle.so")
_lib = None

CURVE_IDS = {"p256r1": 0, "p384r1": 1, "bls12_381_g1": 2, "p256k1": 3}
FIELD_BYTES = {0: 32, 1: 48, 2: 48, 3: 32}
SCALAR_BYTES = {0: 32, 1: 48, 2: 32, 3: 32}
MODE_WINDOW, MODE_COMB, MODE_WNAF = 0, 1, 2


class OracleInvalidInput(ValueError):
    def __init__(self, index, code):
        super().__init__("element %d: %s" % (index, {1: "non-canonical scalar", 2: "bad point"}.get(code, code)))
        self.index = index
        self.code = code


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(LIB_PATH) for f in ("ecc_oracle.c", "mont_tmpl.h")
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = ctypes.CDLL(LIB_PATH)
        lib.orc_init.restype = None
        for name in ("orc_ed25519_mul_base", "orc_ed25519_mul", "orc_ed25519_verify_prehashed", "orc_x25519", "orc_x448",
                     "orc_wei_mul", "orc_ecdsa_verify_hashed", "orc_wei_decompress", "orc_bls12_381_g1_from_compressed",
                     "orc_bls12_381_g1_to_compressed"):
            getattr(lib, name).restype = ctypes.c_long
        lib.orc_init()
        _lib = lib
    return _lib


def _rows(a, w):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(-1, w)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _check(bad, code):
    if bad >= 0:
        raise OracleInvalidInput(int(bad), code.value)
    if bad < -1:
        raise ValueError("bad arguments")


def default_threads():
    return os.cpu_count() or 1


def ed25519_mul_base(k_le, nthreads=1):
    k = _rows(k_le, 32)
    out = np.zeros((k.shape[0], 64), dtype=np.uint8)
    code = ctypes.c_int(0)
    _check(load().orc_ed25519_mul_base(_p(k), ctypes.c_size_t(k.shape[0]), _p(out), nthreads, ctypes.byref(code)), code)
    return out


def ed25519_mul(k_le, xy_le, nthreads=1):
    k, p = _rows(k_le, 32), _rows(xy_le, 64)
    out = np.zeros((k.shape[0], 64), dtype=np.uint8)
    code = ctypes.c_int(0)
    _check(load().orc_ed25519_mul(_p(k), _p(p), ctypes.c_size_t(k.shape[0]), _p(out), nthreads, ctypes.byref(code)), code)
    return out


def ed25519_verify_prehashed(a_enc, r_enc, s_le, k_le, nthreads=1):
    a, r, s, k = (_rows(x, 32) for x in (a_enc, r_enc, s_le, k_le))
    ok = np.zeros(a.shape[0], dtype=np.uint8)
    load().orc_ed25519_verify_prehashed(_p(a), _p(r), _p(s), _p(k), ctypes.c_size_t(a.shape[0]), _p(ok), nthreads)
    return ok.astype(bool)


def x25519(k, u, nthreads=1):
    k, u = _rows(k, 32), _rows(u, 32)
    out = np.zeros((k.shape[0], 32), dtype=np.uint8)
    load().orc_x25519(_p(k), _p(u), ctypes.c_size_t(k.shape[0]), _p(out), nthreads)
    return out


def x448(k, u, nthreads=1):
    k, u = _rows(k, 56), _rows(u, 56)
    out = np.zeros((k.shape[0], 56), dtype=np.uint8)
    load().orc_x448(_p(k), _p(u), ctypes.c_size_t(k.shape[0]), _p(out), nthreads)
    return out


def wei_mul(curve, k_be, xy_be=None, inf_in=None, mode=MODE_WINDOW, nthreads=1):
    cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
    fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
    k = _rows(k_be, sb)
    n = k.shape[0]
    p = None if xy_be is None else _rows(xy_be, 2 * fb)
    if p is None and mode != MODE_COMB:
        raise ValueError("points required")
    if inf_in is not None:
        inf_in = np.ascontiguousarray(inf_in, dtype=np.uint8).reshape(n)
    out = np.zeros((n, 2 * fb), dtype=np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    code = ctypes.c_int(0)
    _check(load().orc_wei_mul(cid, mode, _p(k), _p(p), _p(inf_in), ctypes.c_size_t(n), _p(out), _p(inf), nthreads,
                              ctypes.byref(code)), code)
    return out, inf.astype(bool)


def wei_mul_base(curve, k_be, nthreads=1):
    return wei_mul(curve, k_be, None, None, MODE_COMB, nthreads)


def wei_decompress(curve, x_be, sign, nthreads=1):
    """PointAffine::decompress over a batch -> (x || y rows, present)."""
    cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
    fb = FIELD_BYTES[cid]
    x = _rows(x_be, fb)
    n = x.shape[0]
    sg = np.ascontiguousarray(sign, dtype=np.uint8).reshape(n)
    out = np.zeros((n, 2 * fb), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    load().orc_wei_decompress(cid, _p(x), _p(sg), ctypes.c_size_t(n), _p(out), _p(ok), nthreads)
    return out, ok.astype(bool)


def bls12_381_g1_from_compressed(enc, check_subgroup=True, nthreads=1):
    e = _rows(enc, 48)
    n = e.shape[0]
    out = np.zeros((n, 96), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    load().orc_bls12_381_g1_from_compressed(_p(e), ctypes.c_size_t(n), 1 if check_subgroup else 0, _p(out), _p(ok), nthreads)
    return out, ok.astype(bool)


def bls12_381_g1_to_compressed(xy_be, inf=None, nthreads=1):
    p = _rows(xy_be, 96)
    n = p.shape[0]
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8).reshape(n)
    out = np.zeros((n, 48), dtype=np.uint8)
    load().orc_bls12_381_g1_to_compressed(_p(p), _p(inf), ctypes.c_size_t(n), _p(out), nthreads)
    return out


def bls12_381_beta():
    out = np.zeros(48, dtype=np.uint8)
    load().orc_bls12_381_beta(_p(out))
    return out.tobytes()


def ecdsa_verify_hashed(curve, q_xy_be, z_be, rs_be, nthreads=1):
    cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
    fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
    q, z, rs = _rows(q_xy_be, 2 * fb), _rows(z_be, sb), _rows(rs_be, 2 * sb)
    ok = np.zeros(q.shape[0], dtype=np.uint8)
    code = ctypes.c_int(0)
    _check(load().orc_ecdsa_verify_hashed(cid, _p(q), _p(z), _p(rs), ctypes.c_size_t(q.shape[0]), _p(ok), nthreads,
                                          ctypes.byref(code)), code)
    return ok.astype(bool)


def ed25519_comb_entry(i, j):
    out = np.zeros(64, dtype=np.uint8)
    load().orc_ed25519_comb_entry(i, j, _p(out))
    return out.tobytes()


def wei_comb_entry(curve, i, j):
    cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
    out = np.zeros(2 * FIELD_BYTES[cid], dtype=np.uint8)
    load().orc_wei_comb_entry(cid, i, j, _p(out))
    return out.tobytes()
