#!/usr/bin/env python3
"""Generate eccoxide_b200/csrc/params_gen.cuh: Montgomery constants of the in-scope fields/curves.

Inputs are the standard domain parameters (SEC2 / FIPS 186 / BLS12-381), the same values the
reference keeps in src/params/sec2.rs:1041- (p256r1, p384r1) and src/params/bls12_381.rs:25-85;
they are written here from the standards, and tests/test_params.py re-checks them against the
golden file extracted from the reference.  Everything derived (R, R^2, -p^-1 mod 2^32, Montgomery
forms of b, 3b, Gx, Gy) is computed, not typed.
"""
import os

FIELDS = {}
CURVES = {}

P256_P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
P256_N = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
P256_B = 0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b
P256_GX = 0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296
P256_GY = 0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5

P384_P = 0xfffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffff0000000000000000ffffffff
P384_N = 0xffffffffffffffffffffffffffffffffffffffffffffffffc7634d81f4372ddf581a0db248b0a77aecec196accc52973
P384_B = 0xb3312fa7e23ee7e4988e056be3f82d19181d9c6efe8141120314088f5013875ac656398d8a2ed19d2a85c8edd3ec2aef
P384_GX = 0xaa87ca22be8b05378eb1c71ef320ad746e1d3b628ba79b9859f741e082542a385502f25dbf55296c3a545e3872760ab7
P384_GY = 0x3617de4a96262c6f5d9e98bf9292dc29f8f41dbd289a147ce9da3113b5f0b8c00a60b1ce1d7e819d7a431d7c90ea0e5f

# secp256k1 (the reference's p256k1, src/params/sec2.rs:908-; SEC 2 v2 section 2.4.1)
K256_P = 2**256 - 2**32 - 977
K256_N = 0xfffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141
K256_GX = 0x79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798
K256_GY = 0x483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8

BLS_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
BLS_R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
BLS_GX = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
BLS_GY = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1

# edwards25519 group order l (src/curve/curve25519.rs:46 ORDER_LIMBS): scalar field of Ed25519
ED_L = 2**252 + 27742317777372353535851937790883648493


def iroot(n, k):
    """floor(n ** (1/k)) for big integers."""
    lo, hi = 0, 1 << (n.bit_length() // k + 1)
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if mid**k <= n:
            lo = mid
        else:
            hi = mid - 1
    return lo


def sha512_constants():
    """FIPS 180-4: K[t] = first 64 bits of the fractional part of the cube root of the t-th prime,
    H0[i] = the same for the square roots of the first 8 primes.  Computed, not typed."""
    primes, c = [], 2
    while len(primes) < 80:
        if all(c % q for q in primes if q * q <= c):
            primes.append(c)
        c += 1
    K = [iroot(pr << (3 * 64), 3) & ((1 << 64) - 1) for pr in primes]
    H0 = [iroot(pr << (2 * 64), 2) & ((1 << 64) - 1) for pr in primes[:8]]
    H384 = [iroot(pr << (2 * 64), 2) & ((1 << 64) - 1) for pr in primes[8:16]]
    K256 = [iroot(pr << (3 * 32), 3) & 0xffffffff for pr in primes[:64]]
    H256 = [iroot(pr << (2 * 32), 2) & 0xffffffff for pr in primes[:8]]
    return K, H0, H384, K256, H256


def limbs(x, n):
    return [(x >> (32 * i)) & 0xffffffff for i in range(n)]


def arr(name, x, n):
    return "ECB_CONST u32 %s[%d] = {%s};\n" % (name, n, ", ".join("0x%08xu" % w for w in limbs(x, n)))


PLAIN_FIELDS = {"K256_FP"}   # no Montgomery domain: pseudo-Mersenne folding, R = 1 (mont.cuh kind 3)


def field(name, p, n):
    R = 1 if name in PLAIN_FIELDS else 1 << (32 * n)
    inv = (-pow(p, -1, 1 << 32)) % (1 << 32)
    s = "// ---- field %s: p = 0x%x\n" % (name, p)
    s += arr(name + "_MOD", p, n)
    s += arr(name + "_R1", R % p, n)
    s += arr(name + "_R2", (R * R) % p, n)
    s += ("struct %s {\n    static constexpr int N = %d;\n    static constexpr u32 INV = 0x%08xu;\n"
          "    ECB_DEV static u32 mod(int i) { return %s_MOD[i]; }\n"
          "    ECB_DEV static u32 r1(int i) { return %s_R1[i]; }\n"
          "    ECB_DEV static u32 r2(int i) { return %s_R2[i]; }\n};\n\n") % (name, n, inv, name, name, name)
    return s


def mont(x, p, n):
    if p == K256_P:   # plain field (PLAIN_FIELDS)
        return x % p
    return (x << (32 * n)) % p


def bls_beta():
    """The cube root of unity beta in Fp with (beta x, y) = [-x^2](x, y) on G1 (x = the curve seed):
    the reference pins its BETA the same way (bls12_381/g1.rs:440-450); tests compare the bytes."""
    p, r = BLS_P, BLS_R

    def add(P, Q):
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % p == 0:
                return None
            lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
        else:
            lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
        x = (lam * lam - P[0] - Q[0]) % p
        return (x, (lam * (P[0] - x) - P[1]) % p)

    def mul(k, P):
        R = None
        while k:
            if k & 1:
                R = add(R, P)
            P = add(P, P)
            k >>= 1
        return R

    G = (BLS_GX, BLS_GY)
    target = mul((-(0xd201000000010000 ** 2)) % r, G)
    g = 2
    while pow(g, (p - 1) // 3, p) == 1:
        g += 1
    beta = pow(g, (p - 1) // 3, p)
    for cand in (beta, beta * beta % p):
        if (cand * G[0] % p, G[1]) == target:
            return cand
    raise AssertionError("no cube root of unity acts as [-x^2] on G1")


def safegcd_jump_table(K=4):
    """Jump table of the safegcd divsteps (csrc/modinv.cuh sg_divsteps30_jump): K = 4 steps at a time.
    K steps depend only on f mod 2^K (odd), g mod 2^K and on where zeta sits relative to zero, so they are tabulated:
    index = ((clamp(zeta, -6, 4) + 6) * 8 + (f >> 1 & 7)) * 16 + (g & 15); entry = the K-step transition matrix
    (u, v, q, r), [f', g'] = M [f, g] / 2^K, as four 6-bit signed fields, and the new zeta as sgn * zeta + off
    (off: 6 bits signed at bit 24, sgn at bit 30).  The two clamped classes are affine in zeta: zeta >= 4 never
    swaps within four steps, zeta <= -6 swaps at most once (at the first odd g) and stays positive afterwards.
    The device function is checked against 30 single steps in tests/test_hostsim.py."""
    M = 0xffffffff

    def step(zeta, f, g, u, v, q, r):   # one divstep, as sg_divsteps30
        c1, c2 = zeta < 0, g & 1
        x, y, z = ((-f) & M, (-u) & M, (-v) & M) if c1 else (f, u, v)
        if c2:
            g, q, r = (g + x) & M, (q + y) & M, (r + z) & M
        c = c1 and c2
        zeta = (~zeta if c else zeta) - 1
        if c:
            f, u, v = (f + g) & M, (u + q) & M, (v + r) & M
        return zeta, f, g >> 1, (u << 1) & M, (v << 1) & M, q, r

    def run(zeta, f, g):
        u, v, q, r = 1, 0, 0, 1
        for _ in range(K):
            zeta, f, g, u, v, q, r = step(zeta, f, g, u, v, q, r)
        sg = lambda t: t - (1 << 32) if t >> 31 else t
        return zeta, (sg(u), sg(v), sg(q), sg(r))

    tbl = []
    for zc in range(11):
        zs = [zc - 6] if 0 < zc < 10 else ([-6, -7] if zc == 0 else [4, 5])
        for fi in range(8):
            for gi in range(16):
                outs = [run(z, 2 * fi + 1, gi) for z in zs]
                assert len({o[1] for o in outs}) == 1
                sgn = (outs[0][0] - outs[1][0]) * (zs[0] - zs[1]) if len(zs) == 2 else 1
                off = outs[0][0] - sgn * zs[0]
                assert sgn in (1, -1) and all(o[0] == sgn * z + off for z, o in zip(zs, outs))
                u, v, q, r = outs[0][1]
                assert all(-32 <= t <= 31 for t in (u, v, q, r, off))
                tbl.append((u & 63) | ((v & 63) << 6) | ((q & 63) << 12) | ((r & 63) << 18) | ((off & 63) << 24) | ((1 if sgn < 0 else 0) << 30))
    return tbl


def ristretto_invsqrt_a_minus_d():
    """RFC 9496 Appendix A: 1 / sqrt(a - d) on edwards25519 (a = -1), the non-negative (even) root; the parity tests
    compare it with the bytes the reference holds (src/curve/curve25519/ristretto255.rs:31)."""
    p = 2**255 - 19
    d = (-121665 * pow(121666, -1, p)) % p
    v = (-1 - d) % p
    sm1 = pow(2, (p - 1) // 4, p)
    r = pow(v, 3, p) * pow(pow(v, 7, p), (p - 5) // 8, p) % p      # SQRT_RATIO_M1(1, v)
    chk = v * r * r % p
    assert chk in (1, p - 1)
    if chk == p - 1:
        r = r * sm1 % p
    return p - r if r & 1 else r


def main():
    out = ("// GENERATED by tools/gen_params.py — do not edit.\n"
           "// Montgomery-domain constants for the Weierstrass curves of the batch path\n"
           "// (reference: src/params/sec2.rs:1041- , src/params/bls12_381.rs:25-85).\n"
           "#pragma once\n#include \"limb.cuh\"\nnamespace ecb {\n\n")
    out += field("P256_FP", P256_P, 8)
    out += field("P256_FN", P256_N, 8)
    out += field("P384_FP", P384_P, 12)
    out += field("P384_FN", P384_N, 12)
    out += field("BLS_FP", BLS_P, 12)
    out += field("BLS_FR", BLS_R, 8)
    out += field("ED_FN", ED_L, 8)
    out += field("K256_FP", K256_P, 8)
    out += field("K256_FN", K256_N, 8)
    K, H0, H384, K256, H256 = sha512_constants()
    out += "// ---- SHA-512 (FIPS 180-4) round constants and initial hash value\n"
    out += "ECB_CONST unsigned long long SHA512_K[80] = {%s};\n" % ", ".join("0x%016xull" % k for k in K)
    out += "ECB_CONST unsigned long long SHA512_H0[8] = {%s};\n" % ", ".join("0x%016xull" % h for h in H0)
    out += "ECB_CONST unsigned long long SHA384_H0[8] = {%s};\n" % ", ".join("0x%016xull" % h for h in H384)
    out += "ECB_CONST u32 SHA256_K[64] = {%s};\n" % ", ".join("0x%08xu" % k for k in K256)
    out += "ECB_CONST u32 SHA256_H0[8] = {%s};\n\n" % ", ".join("0x%08xu" % h for h in H256)
    for cname, p, n, b, gx, gy in (("P256", P256_P, 8, P256_B, P256_GX, P256_GY),
                                   ("P384", P384_P, 12, P384_B, P384_GX, P384_GY),
                                   ("BLSG1", BLS_P, 12, 4, BLS_GX, BLS_GY),
                                   ("K256", K256_P, 8, 7, K256_GX, K256_GY)):
        out += "// ---- curve %s (Montgomery-domain constants)\n" % cname
        out += arr(cname + "_B", mont(b, p, n), n)
        out += arr(cname + "_B3", mont(3 * b % p, p, n), n)
        out += arr(cname + "_GX", mont(gx, p, n), n)
        out += arr(cname + "_GY", mont(gy, p, n), n)
        assert p % 4 == 3
        out += "// (p + 1) / 4: the square-root exponent (p = 3 mod 4)\n"
        out += arr(cname + "_SQRT_E", (p + 1) // 4, n)
        if cname == "BLSG1":
            out += "// beta: (beta x, y) = [-x^2](x, y) on G1 (subgroup test, bls12_381/g1.rs:55, :105), Montgomery domain\n"
            out += arr("BLSG1_BETA", mont(bls_beta(), p, n), n)
            xsq = 0xd201000000010000 ** 2
            assert (xsq ** 2 - xsq + 1) == BLS_R and xsq.bit_length() == 128
            out += "// x^2 (x = the curve seed, r = x^4 - x^2 + 1): the GLV split k = q x^2 + rem, k P = rem P - q (beta x, y) on G1\n"
            out += arr("BLSG1_XSQ", xsq, 4)
        out += "\n"
    out += "// ---- ristretto255 (RFC 9496 Appendix A): 1 / sqrt(a - d), little-endian limbs\n"
    out += arr("RISTRETTO_INVSQRT_A_MINUS_D", ristretto_invsqrt_a_minus_d(), 8)
    jt = safegcd_jump_table()
    out += "\n// ---- safegcd jump table (csrc/modinv.cuh sg_divsteps30_jump): 4 divsteps per entry, see tools/gen_params.py\n"
    out += "#define SG_JUMP_WORDS %d\n" % len(jt)
    out += "ECB_GTABLE u32 SG_JUMP4[SG_JUMP_WORDS] = {%s};\n" % ", ".join("0x%08xu" % e for e in jt)
    out += "\n}  // namespace ecb\n"
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "eccoxide_b200", "csrc", "params_gen.cuh")
    with open(dst, "w") as f:
        f.write(out)
    print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    main()
