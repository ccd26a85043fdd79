"""Context: numpy-level access to the C ABI (host buffers in, host buffers out).

All arrays are contiguous uint8 with one row per element in the reference's wire encodings
(see include/eccbatch.h).  Every method is a single C-ABI call; nothing is computed in Python.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import EccBatchError

FIELD_BYTES = {_lib.CURVE_P256R1: 32, _lib.CURVE_P384R1: 48, _lib.CURVE_BLS12_381_G1: 48, _lib.CURVE_P256K1: 32}
SCALAR_BYTES = {_lib.CURVE_P256R1: 32, _lib.CURVE_P384R1: 48, _lib.CURVE_BLS12_381_G1: 32, _lib.CURVE_P256K1: 32}
CURVE_IDS = {"p256r1": _lib.CURVE_P256R1, "p384r1": _lib.CURVE_P384R1, "bls12_381_g1": _lib.CURVE_BLS12_381_G1, "p256k1": _lib.CURVE_P256K1}


def _rows(a, width, name):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim == 1:
        if a.size % width:
            raise ValueError("%s: length %d is not a multiple of %d" % (name, a.size, width))
        a = a.reshape(-1, width)
    if a.ndim != 2 or a.shape[1] != width:
        raise ValueError("%s: expected shape (n, %d), got %s" % (name, width, a.shape))
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _out(out, shape):
    """Caller-provided output (e.g. pinned memory) or a fresh array."""
    if out is None:
        return np.empty(shape, dtype=np.uint8)
    if out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"] or out.size != int(np.prod(shape)):
        raise ValueError("out: expected contiguous uint8 array of %d bytes" % int(np.prod(shape)))
    return out.reshape(shape)


class Context:
    """One libeccbatch context over `devices` (default: CUDA device 0)."""

    def __init__(self, devices=None, ed25519_comb_w=None):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        if devices is None:
            rc = self._lib.ecb_init(None, 0, ctypes.byref(self._ctx))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self._lib.ecb_init(arr, len(devices), ctypes.byref(self._ctx))
        if rc != _lib.ECB_OK:
            self._ctx = None
            raise EccBatchError(rc, "ecb_init failed (no CUDA device? there is no CPU fallback)")
        if ed25519_comb_w is not None:
            self.set_option("ed25519_comb_w", ed25519_comb_w)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.ecb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc, bad=None):
        if rc != _lib.ECB_OK:
            msg = self._lib.ecb_last_error(self._ctx).decode()
            raise EccBatchError(rc, msg, None if bad is None or bad.value == ctypes.c_size_t(-1).value else bad.value)

    def set_option(self, key, value):
        self._check(self._lib.ecb_set_option(self._ctx, key.encode(), int(value)))

    @property
    def handle(self):
        return self._ctx

    @property
    def lib(self):
        return self._lib

    def launch_count(self):
        return int(self._lib.ecb_launch_count(self._ctx))

    def device_count(self):
        return int(self._lib.ecb_device_count(self._ctx))

    # -- edwards25519 -------------------------------------------------------------------------
    def ed25519_mul_base(self, k_le, compressed=False, out=None):
        k = _rows(k_le, 32, "k_le")
        n = k.shape[0]
        out = _out(out, (n, 32 if compressed else 64))
        bad = ctypes.c_size_t()
        fn = self._lib.ecb_ed25519_mul_base_compressed if compressed else self._lib.ecb_ed25519_mul_base
        self._check(fn(self._ctx, _p(k), n, _p(out), ctypes.byref(bad)), bad)
        return out

    def ed25519_mul(self, k_le, xy_le, out=None):
        k = _rows(k_le, 32, "k_le")
        p = _rows(xy_le, 64, "xy_le")
        if k.shape[0] != p.shape[0]:
            raise ValueError("scalar/point count mismatch")
        n = k.shape[0]
        out = _out(out, (n, 64))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ed25519_mul(self._ctx, _p(k), _p(p), n, _p(out), ctypes.byref(bad)), bad)
        return out

    def ed25519_verify_prehashed(self, a_enc, r_enc, s_le, k_le, out=None):
        a, r, s, k = (_rows(x, 32, nm) for x, nm in ((a_enc, "a_enc"), (r_enc, "r_enc"), (s_le, "s_le"), (k_le, "k_le")))
        n = a.shape[0]
        if not (r.shape[0] == s.shape[0] == k.shape[0] == n):
            raise ValueError("count mismatch")
        ok = _out(out, (n,))
        self._check(self._lib.ecb_ed25519_verify_prehashed(self._ctx, _p(a), _p(r), _p(s), _p(k), n, _p(ok)))
        return ok.astype(bool)

    def ed25519_verify(self, a_enc, msgs, sigs, out=None):
        """ed25519 PublicKey::verify on raw messages: `msgs` is a sequence of bytes objects (ragged);
        the challenge hash SHA-512(R || A || M) mod l runs on the device."""
        a = _rows(a_enc, 32, "a_enc")
        s = _rows(sigs, 64, "sigs")
        n = a.shape[0]
        if s.shape[0] != n or len(msgs) != n:
            raise ValueError("count mismatch")
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(m) for m in msgs) + b"\0", dtype=np.uint8)
        ok = _out(out, (n,))
        self._check(self._lib.ecb_ed25519_verify(self._ctx, _p(a), _p(blob), _p(off), _p(s), n, _p(ok)))
        return ok.astype(bool)

    def ed25519_mul_base_ct(self, k_le, out=None):
        """Point::mul_base for SECRET scalars: the constant-time kernel (csrc/ct.cuh), same bytes as ed25519_mul_base."""
        k = _rows(k_le, 32, "k_le")
        n = k.shape[0]
        out = _out(out, (n, 64))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ed25519_mul_base_ct(self._ctx, _p(k), n, _p(out), ctypes.byref(bad)), bad)
        return out

    def ed25519_public_from_seed(self, seeds, out=None, vartime=False):
        """ed25519 SecretKey::public_key over a batch of 32-byte seeds.  Constant-time in the seeds by default;
        vartime=True is the fast form (ecb_ed25519_public_from_seed_vartime)."""
        sd = _rows(seeds, 32, "seeds")
        n = sd.shape[0]
        out = _out(out, (n, 32))
        fn = self._lib.ecb_ed25519_public_from_seed_vartime if vartime else self._lib.ecb_ed25519_public_from_seed
        self._check(fn(self._ctx, _p(sd), n, _p(out)))
        return out

    def ed25519_sign(self, seeds, msgs, pub=None, out=None, vartime=False):
        """ed25519 Keypair::sign (pub given) / SecretKey::sign (pub None) on raw, ragged messages: n x 64 bytes
        R || S.  Constant-time in the seeds by default; vartime=True is the fast form (ecb_ed25519_sign_vartime)."""
        sd = _rows(seeds, 32, "seeds")
        n = sd.shape[0]
        if len(msgs) != n:
            raise ValueError("count mismatch")
        pb = None
        if pub is not None:
            pb = _rows(pub, 32, "pub")
            if pb.shape[0] != n:
                raise ValueError("count mismatch")
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(m) for m in msgs) + b"\0", dtype=np.uint8)
        out = _out(out, (n, 64))
        fn = self._lib.ecb_ed25519_sign_vartime if vartime else self._lib.ecb_ed25519_sign
        self._check(fn(self._ctx, _p(sd), _p(pb), _p(blob), _p(off), n, _p(out)))
        return out

    def ed25519_sign_fixed(self, seeds, msgs, pub=None, out=None, vartime=False):
        """ed25519_sign for n messages of one length given as an (n, w) byte array (no per-message objects)."""
        sd = _rows(seeds, 32, "seeds")
        n = sd.shape[0]
        m = np.ascontiguousarray(msgs, dtype=np.uint8)
        if m.ndim != 2 or m.shape[0] != n:
            raise ValueError("msgs must be (n, w) bytes")
        pb = None if pub is None else _rows(pub, 32, "pub")
        off = np.arange(n + 1, dtype=np.uint64) * np.uint64(m.shape[1])
        out = _out(out, (n, 64))
        fn = self._lib.ecb_ed25519_sign_vartime if vartime else self._lib.ecb_ed25519_sign
        self._check(fn(self._ctx, _p(sd), _p(pb), _p(m), _p(off), n, _p(out)))
        return out

    # -- X25519 / X448 ------------------------------------------------------------------------
    def x25519(self, k, u, out=None):
        k = _rows(k, 32, "k")
        u = _rows(u, 32, "u")
        if k.shape[0] != u.shape[0]:
            raise ValueError("count mismatch")
        n = k.shape[0]
        out = _out(out, (n, 32))
        self._check(self._lib.ecb_x25519(self._ctx, _p(k), _p(u), n, _p(out)))
        return out

    def x25519_base(self, k, out=None):
        """x25519::x25519_base: public keys of `k` (n x 32 raw secret bytes, clamped inside)."""
        k = _rows(k, 32, "k")
        n = k.shape[0]
        out = _out(out, (n, 32))
        self._check(self._lib.ecb_x25519_base(self._ctx, _p(k), n, _p(out)))
        return out

    def x448(self, k, u, out=None):
        k = _rows(k, 56, "k")
        u = _rows(u, 56, "u")
        if k.shape[0] != u.shape[0]:
            raise ValueError("count mismatch")
        n = k.shape[0]
        out = _out(out, (n, 56))
        self._check(self._lib.ecb_x448(self._ctx, _p(k), _p(u), n, _p(out)))
        return out

    # -- Weierstrass --------------------------------------------------------------------------
    def wei_mul(self, curve, k_be, xy_be, inf_in=None, out=None, out_inf=None, in_subgroup=False):
        """&Point * &Scalar.  in_subgroup (bls12_381_g1 only): the points are known to lie in G1, so the scalar may be split
        over the endomorphism (option bls12_381_g1_glv, scoped to this call) — the same result for such points, 1.35x faster."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        if in_subgroup:
            if cid != _lib.CURVE_BLS12_381_G1:
                raise ValueError("in_subgroup applies to bls12_381_g1")
            self.set_option("bls12_381_g1_glv", 1)
            try:
                return self.wei_mul(cid, k_be, xy_be, inf_in, out, out_inf)
            finally:
                self.set_option("bls12_381_g1_glv", 0)
        fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
        k = _rows(k_be, sb, "k_be")
        p = _rows(xy_be, 2 * fb, "xy_be")
        n = k.shape[0]
        if p.shape[0] != n:
            raise ValueError("count mismatch")
        if inf_in is not None:
            inf_in = np.ascontiguousarray(inf_in, dtype=np.uint8).reshape(n)
        out = _out(out, (n, 2 * fb))
        inf = _out(out_inf, (n,))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_wei_mul(self._ctx, cid, _p(k), _p(p), _p(inf_in), n, _p(out), _p(inf), ctypes.byref(bad)), bad)
        return out, inf.astype(bool)

    def wei_mul_base(self, curve, k_be, out=None, out_inf=None):
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
        k = _rows(k_be, sb, "k_be")
        n = k.shape[0]
        out = _out(out, (n, 2 * fb))
        inf = _out(out_inf, (n,))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_wei_mul_base(self._ctx, cid, _p(k), n, _p(out), _p(inf), ctypes.byref(bad)), bad)
        return out, inf.astype(bool)

    def ecdsa_sign_hashed(self, curve, d_be, k_be, z_be, out=None, out_ok=None, vartime=False):
        """ecdsa::sign_hashed over a batch (not constant-time): (r || s rows, present)."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        sb = SCALAR_BYTES[cid]
        d = _rows(d_be, sb, "d_be")
        k = _rows(k_be, sb, "k_be")
        z = _rows(z_be, sb, "z_be")
        n = d.shape[0]
        if not (k.shape[0] == z.shape[0] == n):
            raise ValueError("count mismatch")
        out = _out(out, (n, 2 * sb))
        ok = _out(out_ok, (n,))
        fn = self._lib.ecb_ecdsa_sign_hashed_vartime if vartime else self._lib.ecb_ecdsa_sign_hashed
        self._check(fn(self._ctx, cid, _p(d), _p(k), _p(z), n, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def ecdsa_sign(self, curve, d_be, k_be, msgs, hash_bits=None, out=None, out_ok=None, vartime=False):
        """ecdsa::sign on raw, ragged messages (hash on the device): (r || s rows, present).  Constant-time in the secret
        and the nonce by default; vartime=True is the fast form."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        sb = SCALAR_BYTES[cid]
        d = _rows(d_be, sb, "d_be")
        k = _rows(k_be, sb, "k_be")
        n = d.shape[0]
        if k.shape[0] != n or len(msgs) != n:
            raise ValueError("count mismatch")
        if hash_bits is None:
            hash_bits = 256 if sb == 32 else 384
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(m) for m in msgs) + b"\0", dtype=np.uint8)
        out = _out(out, (n, 2 * sb))
        ok = _out(out_ok, (n,))
        fn = self._lib.ecb_ecdsa_sign_vartime if vartime else self._lib.ecb_ecdsa_sign
        self._check(fn(self._ctx, cid, int(hash_bits), _p(d), _p(k), _p(blob), _p(off), n, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def wei_decompress(self, curve, x_be, sign, out=None, out_ok=None):
        """PointAffine::decompress over a batch (affine.rs:48): (x || y rows, present)."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        fb = FIELD_BYTES[cid]
        x = _rows(x_be, fb, "x_be")
        n = x.shape[0]
        sg = np.ascontiguousarray(sign, dtype=np.uint8).reshape(-1)
        if sg.shape[0] != n:
            raise ValueError("count mismatch")
        out = _out(out, (n, 2 * fb))
        ok = _out(out_ok, (n,))
        self._check(self._lib.ecb_wei_decompress(self._ctx, cid, _p(x), _p(sg), n, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def bls12_381_g1_from_compressed(self, enc, check_subgroup=True, out=None, out_ok=None):
        """PointAffine::from_compressed / from_compressed_oncurve_only (bls12_381/serialize.rs:286, :310)."""
        e = _rows(enc, 48, "enc")
        n = e.shape[0]
        out = _out(out, (n, 96))
        ok = _out(out_ok, (n,))
        self._check(self._lib.ecb_bls12_381_g1_from_compressed(self._ctx, _p(e), n, 1 if check_subgroup else 0, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def bls12_381_g1_from_uncompressed(self, enc, check_subgroup=True, out=None, out_ok=None):
        """PointAffine::from_uncompressed / from_uncompressed_oncurve_only (bls12_381/serialize.rs:330-380):
        (x || y rows, present, was-the-identity-encoding)."""
        e = _rows(enc, 96, "enc")
        n = e.shape[0]
        out = _out(out, (n, 96))
        ok = _out(out_ok, (n,))
        inf = np.zeros(n, dtype=np.uint8)
        self._check(self._lib.ecb_bls12_381_g1_from_uncompressed(self._ctx, _p(e), n, 1 if check_subgroup else 0, _p(out), _p(inf), _p(ok)))
        return out, ok.astype(bool), inf.astype(bool)

    def bls12_381_g1_to_uncompressed(self, xy_be, inf=None, out=None):
        """Point::to_uncompressed (bls12_381/serialize.rs:412)."""
        p = _rows(xy_be, 96, "xy_be")
        n = p.shape[0]
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8).reshape(-1)
        out = _out(out, (n, 96))
        self._check(self._lib.ecb_bls12_381_g1_to_uncompressed(self._ctx, _p(p), _p(inf), n, _p(out)))
        return out

    def ed25519_decompress(self, enc, out=None, out_ok=None):
        """ed25519 decode_point / Point::decompress over a batch: (x || y rows, present)."""
        e = _rows(enc, 32, "enc")
        n = e.shape[0]
        out = _out(out, (n, 64))
        ok = _out(out_ok, (n,))
        self._check(self._lib.ecb_ed25519_decompress(self._ctx, _p(e), n, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def wei_msm(self, curve, k_be, xy_be):
        """sum_i k_i * P_i (bucket method, csrc/msm.cuh) on bls12_381_g1 / p256k1: (x || y bytes, is_identity)."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
        k = _rows(k_be, sb, "k_be")
        p = _rows(xy_be, 2 * fb, "xy_be")
        n = k.shape[0]
        if p.shape[0] != n:
            raise ValueError("scalar/point count mismatch")
        out = np.zeros(2 * fb, dtype=np.uint8)
        inf = np.zeros(1, dtype=np.uint8)
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_wei_msm(self._ctx, cid, _p(k), _p(p), n, _p(out), _p(inf), ctypes.byref(bad)), bad)
        return out, bool(inf[0])

    # -- ristretto255 ---------------------------------------------------------------------------
    def ristretto255_decompress(self, enc, out=None, out_ok=None):
        """RistrettoPoint::decompress over a batch: (Edwards representative x || y rows, present)."""
        e = _rows(enc, 32, "enc")
        n = e.shape[0]
        out = _out(out, (n, 64))
        ok = _out(out_ok, (n,))
        self._check(self._lib.ecb_ristretto255_decompress(self._ctx, _p(e), n, _p(out), _p(ok)))
        return out, ok.astype(bool)

    def ristretto255_compress(self, xy_le, out=None):
        """RistrettoPoint::compress of affine edwards25519 points."""
        p = _rows(xy_le, 64, "xy_le")
        n = p.shape[0]
        out = _out(out, (n, 32))
        self._check(self._lib.ecb_ristretto255_compress(self._ctx, _p(p), n, _p(out)))
        return out

    def ristretto255_mul(self, k_le, enc, out=None):
        """RistrettoPoint::scale on encodings."""
        k, e = _rows(k_le, 32, "k_le"), _rows(enc, 32, "enc")
        n = k.shape[0]
        if e.shape[0] != n:
            raise ValueError("count mismatch")
        out = _out(out, (n, 32))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ristretto255_mul(self._ctx, _p(k), _p(e), n, _p(out), ctypes.byref(bad)), bad)
        return out

    def ristretto255_mul_base(self, k_le, out=None):
        """RistrettoPoint::mul_base, encoded."""
        k = _rows(k_le, 32, "k_le")
        n = k.shape[0]
        out = _out(out, (n, 32))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ristretto255_mul_base(self._ctx, _p(k), n, _p(out), ctypes.byref(bad)), bad)
        return out

    def bls12_381_g1_to_compressed(self, xy_be, inf=None, out=None):
        """Point::to_compressed (bls12_381/serialize.rs:400)."""
        p = _rows(xy_be, 96, "xy_be")
        n = p.shape[0]
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8).reshape(-1)
            if inf.shape[0] != n:
                raise ValueError("count mismatch")
        out = _out(out, (n, 48))
        self._check(self._lib.ecb_bls12_381_g1_to_compressed(self._ctx, _p(p), _p(inf), n, _p(out)))
        return out

    def ecdsa_verify_hashed(self, curve, q_xy_be, z_be, rs_be, out=None):
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
        q = _rows(q_xy_be, 2 * fb, "q_xy_be")
        z = _rows(z_be, sb, "z_be")
        rs = _rows(rs_be, 2 * sb, "rs_be")
        n = q.shape[0]
        if not (z.shape[0] == rs.shape[0] == n):
            raise ValueError("count mismatch")
        ok = _out(out, (n,))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ecdsa_verify_hashed(self._ctx, cid, _p(q), _p(z), _p(rs), n, _p(ok), ctypes.byref(bad)), bad)
        return ok.astype(bool)

    def ecdsa_verify(self, curve, hash_bits, q_xy_be, msgs, rs_be, out=None):
        """ecdsa::verify on raw messages (ragged `msgs`); hash_bits in (256, 384, 512) selects SHA-2."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else curve
        fb, sb = FIELD_BYTES[cid], SCALAR_BYTES[cid]
        q = _rows(q_xy_be, 2 * fb, "q_xy_be")
        rs = _rows(rs_be, 2 * sb, "rs_be")
        n = q.shape[0]
        if rs.shape[0] != n or len(msgs) != n:
            raise ValueError("count mismatch")
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in msgs], dtype=np.uint64)
        blob = np.frombuffer(b"".join(bytes(m) for m in msgs) + b"\0", dtype=np.uint8)
        ok = _out(out, (n,))
        bad = ctypes.c_size_t()
        self._check(self._lib.ecb_ecdsa_verify(self._ctx, cid, int(hash_bits), _p(q), _p(blob), _p(off), _p(rs), n, _p(ok), ctypes.byref(bad)), bad)
        return ok.astype(bool)

    # -- device-resident variants (raw device pointers, enqueue on `stream`, no sync, no allocation) ---
    def warm(self, op, max_n, curve=None):
        """ecb_warm: build the generator comb `op` needs and size its work buffers for batches up to max_n.
        Required before dev_call (the *_dev entry points never allocate)."""
        cid = CURVE_IDS[curve] if isinstance(curve, str) else (-1 if curve is None else int(curve))
        self._check(self._lib.ecb_warm(self._ctx, op.encode(), cid, int(max_n)))

    def get_info(self, key, dev_index=0):
        v = ctypes.c_long()
        self._check(self._lib.ecb_get_info(self._ctx, dev_index, key.encode(), ctypes.byref(v)))
        return int(v.value)

    def dev_call(self, name, *args):
        """Call a *_dev entry point: args are ints (device pointers, sizes, ids) in ABI order after ctx."""
        self._check(getattr(self._lib, name)(self._ctx, *args))

    def dev_status(self, dev_index=0):
        bad = ctypes.c_size_t()
        rc = self._lib.ecb_dev_status(self._ctx, dev_index, ctypes.byref(bad))
        return rc, (None if bad.value == ctypes.c_size_t(-1).value else bad.value)

    def profile_collect(self, dev_index=0):
        m, f, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        self._check(self._lib.ecb_profile_collect(self._ctx, dev_index, ctypes.byref(m), ctypes.byref(f), ctypes.byref(c)))
        return m.value, f.value, c.value

    # -- measurement --------------------------------------------------------------------------
    def imad_probe(self, variant, iters=4096, dev_index=0):
        macs = ctypes.c_double()
        ms = ctypes.c_double()
        self._check(self._lib.ecb_imad_probe(self._ctx, dev_index, variant, iters, ctypes.byref(macs), ctypes.byref(ms)))
        return macs.value, ms.value

    def latency_probe(self, variant, threads=32, reps=16, dev_index=0):
        cyc, mhz = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.ecb_latency_probe(self._ctx, dev_index, variant, threads, reps, ctypes.byref(cyc), ctypes.byref(mhz)))
        return cyc.value, mhz.value

    def fieldmul_probe(self, fp64_num, fp64_den, blocks_per_sm=4, reps=512, dev_index=0):
        v, chk = ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.ecb_fieldmul_probe(self._ctx, dev_index, fp64_num, fp64_den, blocks_per_sm, reps, ctypes.byref(v), ctypes.byref(chk)))
        return v.value, int(chk.value)

    def debug_fused_trace(self, dev_index=0):
        nb = int(self._lib.ecb_debug_fused_trace(self._ctx, dev_index, None, 0))
        if nb < 0:
            self._check(nb)
        out = np.zeros((max(nb, 1), 4), dtype=np.uint64)
        self._lib.ecb_debug_fused_trace(self._ctx, dev_index, _p(out), nb)
        return out[:nb]

    def debug_ed25519_table(self, dev_index=0):
        w = ctypes.c_int()
        nwin = ctypes.c_int()
        ntab = self._lib.ecb_debug_ed25519_table(self._ctx, dev_index, None, 0, ctypes.byref(w), ctypes.byref(nwin))
        if ntab < 0:
            self._check(int(ntab))
        out = np.empty((ntab, 96), dtype=np.uint8)
        self._lib.ecb_debug_ed25519_table(self._ctx, dev_index, _p(out), out.nbytes, ctypes.byref(w), ctypes.byref(nwin))
        return out, w.value, nwin.value
