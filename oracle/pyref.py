"""pyref.py — exact big-integer restatement of the reference's hot-path *semantics*.

TEST INFRASTRUCTURE ONLY.  Nothing under eccoxide_b200/ imports this; it is used by tests/,
__graft_entry__.smoke() and (never on the measured path) bench.py's checker.

Each function states the reference item whose observable result it reproduces.  Because the
reference's outputs on this path are canonical affine / wire bytes (SURVEY.md §8a "parity-relevant
subtleties": Point fields are private, == is projective), a plain affine group law over Python
ints is an exact oracle for them; the C restatement in oracle/ecc_oracle.c follows the reference's
*algorithms* (limbs, comb, windows) and is cross-checked against this file.

Parity pinning: tests/test_oracle_golden.py checks this module against every known-answer vector
extracted from the reference's own tests (tests/golden/reference_vectors.json).
"""
import hashlib

# --------------------------------------------------------------------------------------
# curve25519 / edwards25519   (src/curve/curve25519.rs:39-51, :364-423)
# --------------------------------------------------------------------------------------
P25519 = 2**255 - 19
L25519 = 2**252 + 27742317777372353535851937790883648493
ED_D = (-121665 * pow(121666, -1, P25519)) % P25519
ED_BY = 4 * pow(5, -1, P25519) % P25519
SQRT_M1 = pow(2, (P25519 - 1) // 4, P25519)


def _ed_recover_x(y, sign):
    """Point::decompress (curve25519.rs:772) via sqrt_div (:258): None if no such point."""
    p = P25519
    u = (y * y - 1) % p
    v = (ED_D * y * y + 1) % p
    # sqrt_div candidate r = u v^3 (u v^7)^((p-5)/8)
    r = u * pow(v, 3, p) * pow(u * pow(v, 7, p), (p - 5) // 8, p) % p
    check = v * r * r % p
    if check == u:
        pass
    elif check == (-u) % p:
        r = r * SQRT_M1 % p
    else:
        return None
    if (r & 1) != sign:
        r = (-r) % p
    return r


ED_BX = _ed_recover_x(ED_BY, 0)
ED_B = (ED_BX, ED_BY)
ED_ID = (0, 1)


def ed_add(P, Q):
    """Complete twisted-Edwards addition (a = -1); same group law as Point::add (curve25519.rs:695)."""
    p = P25519
    x1, y1 = P
    x2, y2 = Q
    k = ED_D * x1 * x2 * y1 * y2 % p
    x3 = (x1 * y2 + x2 * y1) * pow(1 + k, -1, p) % p
    y3 = (y1 * y2 + x1 * x2) * pow(1 - k, -1, p) % p
    return (x3, y3)


def ed_mul(k, P):
    """k*P for any integer k >= 0 (Point::scale curve25519.rs:760 / mul_base :840 give the same point)."""
    R = ED_ID
    Q = P
    while k:
        if k & 1:
            R = ed_add(R, Q)
        Q = ed_add(Q, Q)
        k >>= 1
    return R


def ed_on_curve(x, y):
    """Point::from_coordinate check (curve25519.rs:649)."""
    p = P25519
    return (y * y - x * x - 1 - ED_D * x * x * y * y) % p == 0


def ed_encode(P):
    """encode_point (protocol/ed25519.rs:27): y little-endian, bit 255 = low bit of canonical x."""
    x, y = P
    b = bytearray(y.to_bytes(32, "little"))
    b[31] |= (x & 1) << 7
    return bytes(b)


def ed_decode(b):
    """decode_point (protocol/ed25519.rs:38): rejects non-canonical y, (x=0, sign=1), off-curve."""
    sign = b[31] >> 7
    y = int.from_bytes(b, "little") & ((1 << 255) - 1)
    if y >= P25519:
        return None
    if sign == 1 and (y == 1 or y == P25519 - 1):
        return None
    x = _ed_recover_x(y, sign)
    if x is None:
        return None
    return (x, y)


# ---- ristretto255 (RFC 9496; src/curve/curve25519/ristretto255.rs:73-134) ---------------------------------
def _fe_neg(x):
    return x & 1   # is_negative: the low bit of the canonical representative


def _sqrt_ratio_m1(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1: (was_square, nonnegative r) with r = sqrt(u / v) or sqrt(i u / v)."""
    p = P25519
    r = u * pow(v, 3, p) * pow(u * pow(v, 7, p) % p, (p - 5) // 8, p) % p
    check = v * r * r % p
    correct = check == u % p
    flipped = check == (-u) % p
    flipped_i = check == (-u) * SQRT_M1 % p
    if flipped or flipped_i:
        r = r * SQRT_M1 % p
    if _fe_neg(r):
        r = p - r
    return correct or flipped, r


RISTRETTO_INVSQRT_A_MINUS_D = _sqrt_ratio_m1(1, (-1 - ED_D) % P25519)[1]   # RFC 9496 Appendix A; sign pinned by the KATs


def ristretto255_decompress(b):
    """RistrettoPoint::decompress (ristretto255.rs:105): an edwards25519 representative (x, y), or None."""
    p = P25519
    s = int.from_bytes(b, "little")
    if s >= p or _fe_neg(s):
        return None
    ss = s * s % p
    u1, u2 = (1 - ss) % p, (1 + ss) % p
    u2_sq = u2 * u2 % p
    v = (-(ED_D * u1 * u1) - u2_sq) % p
    was_square, invsqrt = _sqrt_ratio_m1(1, v * u2_sq % p)
    den_x = invsqrt * u2 % p
    den_y = invsqrt * den_x * v % p
    x = 2 * s * den_x % p
    if _fe_neg(x):
        x = p - x
    y = u1 * den_y % p
    t = x * y % p
    if not was_square or _fe_neg(t) or y == 0:
        return None
    return (x, y)


def ristretto255_compress(P):
    """RistrettoPoint::compress (ristretto255.rs:73) of the affine edwards25519 point P = (x, y)."""
    p = P25519
    x0, y0 = P
    z0, t0 = 1, x0 * y0 % p
    u1 = (z0 + y0) * (z0 - y0) % p
    u2 = x0 * y0 % p
    _, invsqrt = _sqrt_ratio_m1(1, u1 * u2 * u2 % p)
    den1, den2 = invsqrt * u1 % p, invsqrt * u2 % p
    z_inv = den1 * den2 * t0 % p
    ix, iy = x0 * SQRT_M1 % p, y0 * SQRT_M1 % p
    ench = den1 * RISTRETTO_INVSQRT_A_MINUS_D % p
    rotate = _fe_neg(t0 * z_inv % p)
    x, y, den_inv = (iy, ix, ench) if rotate else (x0, y0, den2)
    if _fe_neg(x * z_inv % p):
        y = (-y) % p
    s = den_inv * (z0 - y) % p
    if _fe_neg(s):
        s = p - s
    return s.to_bytes(32, "little")


def ristretto255_mul(k_le32, enc):
    """RistrettoPoint::scale (ristretto255.rs:152): encoding of k * P, or None for an invalid encoding / scalar."""
    k = int.from_bytes(k_le32, "little")
    P = ristretto255_decompress(enc)
    if P is None or k >= L25519:
        return None
    return ristretto255_compress(ed_mul(k, P))


def ristretto255_mul_base(k_le32):
    k = int.from_bytes(k_le32, "little")
    return None if k >= L25519 else ristretto255_compress(ed_mul(k, ED_B))


def ed25519_mul_base_xy(k_le32):
    """Point::mul_base(&Scalar) then to_affine, as 64 bytes x_le || y_le.  Scalar must be canonical."""
    k = int.from_bytes(k_le32, "little")
    assert k < L25519
    x, y = ed_mul(k, ED_B)
    return x.to_bytes(32, "little") + y.to_bytes(32, "little")


def ed25519_mul_xy(k_le32, xy64):
    """&Point * &Scalar (curve25519.rs:1274) then to_affine."""
    k = int.from_bytes(k_le32, "little")
    P = (int.from_bytes(xy64[:32], "little"), int.from_bytes(xy64[32:], "little"))
    x, y = ed_mul(k, P)
    return x.to_bytes(32, "little") + y.to_bytes(32, "little")


def ed25519_verify_prehashed(a_enc, r_enc, s_le, k_le):
    """protocol/ed25519.rs:119-147 with k = SHA-512(R||A||M) mod l supplied by the caller."""
    A = ed_decode(a_enc)
    if A is None:
        return False
    R = ed_decode(r_enc)
    if R is None:
        return False
    s = int.from_bytes(s_le, "little")
    if s >= L25519:
        return False
    k = int.from_bytes(k_le, "little")
    negA = ((-A[0]) % P25519, A[1])
    lhs = ed_add(ed_mul(s, ED_B), ed_mul(k, negA))
    return lhs == R


def ed25519_hash_k(r_enc, a_enc, msg):
    """reduce_wide_le(SHA-512(R || A || M)) (protocol/ed25519.rs:21, :139)."""
    h = hashlib.sha512(r_enc + a_enc + msg).digest()
    return (int.from_bytes(h, "little") % L25519).to_bytes(32, "little")


def ed25519_verify(pub, msg, sig):
    return ed25519_verify_prehashed(pub, sig[:32], sig[32:], ed25519_hash_k(sig[:32], pub, msg))


def ed25519_public_from_seed(seed):
    """public_from_seed (protocol/ed25519.rs:84): clamp, reduce mod l, mul_base, encode."""
    h = hashlib.sha512(seed).digest()
    a = bytearray(h[:32])
    a[0] &= 248
    a[31] &= 127
    a[31] |= 64
    k = int.from_bytes(a, "little") % L25519
    return ed_encode(ed_mul(k, ED_B))


def ed25519_sign(seed, msg):
    """sign (protocol/ed25519.rs:112)."""
    h = hashlib.sha512(seed).digest()
    a = bytearray(h[:32])
    a[0] &= 248
    a[31] &= 127
    a[31] |= 64
    a = int.from_bytes(a, "little") % L25519
    prefix = h[32:]
    pub = ed_encode(ed_mul(a, ED_B))
    r = int.from_bytes(hashlib.sha512(prefix + msg).digest(), "little") % L25519
    R = ed_encode(ed_mul(r, ED_B))
    k = int.from_bytes(hashlib.sha512(R + pub + msg).digest(), "little") % L25519
    s = (r + k * a) % L25519
    return R + s.to_bytes(32, "little")


# ---- X25519 (protocol/x25519.rs:36, ladder curve25519.rs:474-513) ----
def _ladder(p, a24, x1, k, nbits):
    """The reference's ladder verbatim (all nbits bits, MSB first, final x2 * z2^(p-2), 0 for z2 = 0)."""
    x2, z2, x3, z3 = 1, 0, x1, 1
    swap = 0
    for t in range(nbits - 1, -1, -1):
        bit = (k >> t) & 1
        swap ^= bit
        if swap:
            x2, x3 = x3, x2
            z2, z3 = z3, z2
        swap = bit
        a = (x2 + z2) % p
        aa = a * a % p
        b = (x2 - z2) % p
        bb = b * b % p
        e = (aa - bb) % p
        c = (x3 + z3) % p
        d = (x3 - z3) % p
        da = d * a % p
        cb = c * b % p
        x3 = (da + cb) ** 2 % p
        z3 = x1 * (da - cb) ** 2 % p
        x2 = aa * bb % p
        z2 = e * (bb + a24 * e) % p
    if swap:
        x2, x3 = x3, x2
        z2, z3 = z3, z2
    return x2 * pow(z2, p - 2, p) % p


def x25519(scalar32, u32):
    k = bytearray(scalar32)
    k[0] &= 248
    k[31] &= 127
    k[31] |= 64
    u = int.from_bytes(u32, "little") & ((1 << 255) - 1)  # decode_u: mask bit 255, no canonical check
    return _ladder(P25519, 121666, u % P25519, int.from_bytes(k, "little"), 256).to_bytes(32, "little")


# ---- X448 (protocol/x448.rs:34, ladder curve448.rs:263-302) ----
P448 = 2**448 - 2**224 - 1


def x448(scalar56, u56):
    k = bytearray(scalar56)
    k[0] &= 252
    k[55] |= 128
    u = int.from_bytes(u56, "little")
    return _ladder(P448, 39082, u % P448, int.from_bytes(k, "little"), 448).to_bytes(56, "little")


# --------------------------------------------------------------------------------------
# short Weierstrass curves (src/curve/projective.rs, params/sec2.rs, params/bls12_381.rs)
# --------------------------------------------------------------------------------------
class WCurve:
    def __init__(self, name, p, n, a, b, gx, gy, fbytes, sbytes):
        self.name, self.p, self.n, self.a, self.b = name, p, n, a % p, b
        self.G = (gx, gy)
        self.fbytes, self.sbytes = fbytes, sbytes

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - (x * x * x + self.a * x + self.b)) % self.p == 0

    def add(self, P, Q):
        p = self.p
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1 + self.a) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    def mul(self, k, P):
        R = None
        Q = P
        while k:
            if k & 1:
                R = self.add(R, Q)
            Q = self.add(Q, Q)
            k >>= 1
        return R

    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.p)

    def enc(self, P):
        """affine x || y big-endian (PointAffine::to_coordinate + FieldElement::to_bytes)."""
        return P[0].to_bytes(self.fbytes, "big") + P[1].to_bytes(self.fbytes, "big")

    def dec(self, b):
        return (int.from_bytes(b[: self.fbytes], "big"), int.from_bytes(b[self.fbytes:], "big"))


P256 = WCurve(
    "p256r1",
    0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff,
    0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551,
    -3,
    0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b,
    0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296,
    0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5,
    32, 32)
P384 = WCurve(
    "p384r1",
    0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffff0000000000000000ffffffff,
    0xffffffffffffffffffffffffffffffffffffffffffffffffc7634d81f4372ddf581a0db248b0a77aecec196accc52973,
    -3,
    0xb3312fa7e23ee7e4988e056be3f82d19181d9c6efe8141120314088f5013875ac656398d8a2ed19d2a85c8edd3ec2aef,
    0xaa87ca22be8b05378eb1c71ef320ad746e1d3b628ba79b9859f741e082542a385502f25dbf55296c3a545e3872760ab7,
    0x3617de4a96262c6f5d9e98bf9292dc29f8f41dbd289a147ce9da3113b5f0b8c00a60b1ce1d7e819d7a431d7c90ea0e5f,
    48, 48)
BLSG1 = WCurve(
    "bls12_381_g1",
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    0, 4,
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
    48, 32)
K256 = WCurve(   # secp256k1: the reference's p256k1 (src/curve/sec2/p256k1.rs, parameters src/params/sec2.rs:908-)
    "p256k1",
    2**256 - 2**32 - 977,
    0xfffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141,
    0, 7,
    0x79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798,
    0x483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8,
    32, 32)
WCURVES = {"p256r1": P256, "p384r1": P384, "bls12_381_g1": BLSG1, "p256k1": K256}


def wei_mul(curve, k_be, xy_be):
    """&Point * &Scalar (fiat/curve_macros.rs:321 -> projective.rs:871/842) then to_affine.
    Returns (xy bytes or zeros, is_infinity)."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    k = int.from_bytes(k_be, "big")
    assert k < c.n
    R = c.mul(k, c.dec(xy_be))
    if R is None:
        return bytes(2 * c.fbytes), 1
    return c.enc(R), 0


def wei_mul_base(curve, k_be):
    """Point::mul_base (fiat/curve_macros.rs:55 -> projective.rs:965/945) then to_affine."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    return wei_mul(c, k_be, c.enc(c.G))


def wei_decompress(curve, x_be, sign):
    """PointAffine::decompress(&x, sign) (src/curve/affine.rs:48, fiat/curve_macros.rs:221):
    y = sqrt(x^3 + a x + b) = (..)^((p+1)/4) (p256r1.rs:68, p384r1.rs:71, bls12_381/fp.rs:64; all three
    primes are 3 mod 4), present iff y^2 is the right-hand side; the root with the requested parity
    (Sign::Positive = even, Negative = odd of the canonical value, field_macros.rs:557).
    x_be must be canonical (FieldElement::from_bytes -> None otherwise, field_macros.rs:15).
    Returns x || y bytes, or None."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    x = int.from_bytes(x_be, "big")
    if x >= c.p:
        return None
    yy = (x * x * x + c.a * x + c.b) % c.p
    y = pow(yy, (c.p + 1) // 4, c.p)
    if y * y % c.p != yy:
        return None
    if (y & 1) != (1 if sign else 0):
        y = (-y) % c.p
    return c.enc((x, y))


BLS_X_ABS = 0xd201000000010000  # |x|, the BLS12-381 seed (params/bls12_381.rs)


def bls_g1_in_subgroup(P):
    """PointAffine::is_in_subgroup (bls12_381/g1.rs:105), by the DEFINITION [r]P = infinity rather
    than the reference's endomorphism test: same predicate, independent computation."""
    return BLSG1.mul(BLSG1.n, P) is None


def bls_g1_from_compressed(enc, check_subgroup=True):
    """PointAffine::from_compressed / from_compressed_oncurve_only (bls12_381/serialize.rs:286-321 with
    read_compressed_flags :117 and read_compressed_affine :172): 48 bytes -> x || y (96 bytes) or None.
    The identity encoding is None as well (the affine type cannot hold it)."""
    c = BLSG1
    flags = enc[0] & 0xE0
    if not flags & 0x80:
        return None
    if flags & 0x40:
        return None  # infinity (canonical or not): no affine point either way
    sort = 1 if flags & 0x20 else 0
    x = int.from_bytes(bytes([enc[0] & 0x1F]) + bytes(enc[1:]), "big")
    if x >= c.p:
        return None
    yy = (x * x * x + c.b) % c.p
    y = pow(yy, (c.p + 1) // 4, c.p)
    if y * y % c.p != yy:
        return None
    if (1 if y > (c.p - 1) // 2 else 0) != sort:
        y = (-y) % c.p
    if check_subgroup and not bls_g1_in_subgroup((x, y)):
        return None
    return c.enc((x, y))


def bls_g1_to_compressed(xy_be, inf=0):
    """Point::to_compressed (serialize.rs:400-420): x with the compression flag and the sort flag
    (y > (p-1)/2); the identity is 0xc0 followed by zeros."""
    c = BLSG1
    if inf:
        return bytes([0xC0]) + bytes(47)
    x, y = c.dec(xy_be)
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80 | (0x20 if y > (c.p - 1) // 2 else 0)
    return bytes(out)


def bls_g1_from_uncompressed(enc, check_subgroup=True):
    """PointAffine::from_uncompressed / _oncurve_only (bls12_381/serialize.rs:330-380, flags :129-140): x || y or None."""
    c = BLSG1
    flags = enc[0] >> 5
    if flags & 0b101:
        return None
    if flags & 0b010:
        return None          # the identity (valid only with a zero payload) has no affine form either way
    x, y = int.from_bytes(enc[:48], "big"), int.from_bytes(enc[48:], "big")
    if x >= c.p or y >= c.p or not c.on_curve((x, y)):
        return None
    if check_subgroup and not bls_g1_in_subgroup((x, y)):
        return None
    return bytes(enc)


def bls_g1_to_uncompressed(xy_be, inf=0):
    """Point::to_uncompressed (bls12_381/serialize.rs:412)."""
    return bytes([0x40]) + bytes(95) if inf else bytes(xy_be)


def ecdsa_verify_hashed(curve, q_xy_be, z_be, rs_be):
    """verify_hashed (protocol/ecdsa.rs:205-222); Signature::from_bytes (:399) rejects zero or
    non-canonical r, s; z is reduced mod n as digest_to_scalar (:340) would."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    sb = c.sbytes
    r = int.from_bytes(rs_be[:sb], "big")
    s = int.from_bytes(rs_be[sb:], "big")
    if r == 0 or s == 0 or r >= c.n or s >= c.n:
        return False
    z = int.from_bytes(z_be, "big") % c.n
    Q = c.dec(q_xy_be)
    if not c.on_curve(Q):
        return False
    si = pow(s, -1, c.n)
    u1, u2 = z * si % c.n, r * si % c.n
    R = c.add(c.mul(u1, c.G), c.mul(u2, Q))
    if R is None:
        return False
    return R[0] % c.n == r


def ecdsa_digest_to_scalar(curve, digest):
    """digest_to_scalar / bits2int (protocol/ecdsa.rs:340-352), big-endian scalar bytes."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    nbits = c.n.bit_length()
    e = int.from_bytes(digest, "big")
    if len(digest) * 8 > nbits:
        e >>= len(digest) * 8 - nbits
    return (e % c.n).to_bytes(c.sbytes, "big")


def ecdsa_sign_hashed(curve, d, k, z):
    """sign_hashed (protocol/ecdsa.rs:165-184) with integers; returns r||s big-endian."""
    c = WCURVES[curve] if isinstance(curve, str) else curve
    R = c.mul(k, c.G)
    r = R[0] % c.n
    s = pow(k, -1, c.n) * (z + r * d) % c.n
    return r.to_bytes(c.sbytes, "big") + s.to_bytes(c.sbytes, "big")
