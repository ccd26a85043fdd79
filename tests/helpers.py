"""Shared input generators for the parity tests (seeded, synthetic: SURVEY.md §8d)."""
import hashlib

import numpy as np

from oracle import pyref as R

SEEDS = {"ed25519": 0xECC00001, "x25519": 0xECC00002, "p256": 0xECC00003, "bls": 0xECC00004, "sweep": 0xECC00005}


def rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def rows(lst, width=None):
    a = np.frombuffer(b"".join(lst), dtype=np.uint8).copy()  # writable
    return a.reshape(len(lst), -1) if width is None else a.reshape(-1, width)


def rand_bytes(g, n, width):
    return g.integers(0, 256, size=(n, width), dtype=np.uint8)


def scalars_mod(g, n, order, nbytes, endian):
    """n uniform scalars: wide random bytes reduced mod `order` (as init_from_wide_bytes does)."""
    wide = rand_bytes(g, n, nbytes + 16)
    out = [(int.from_bytes(wide[i].tobytes(), "little") % order).to_bytes(nbytes, endian) for i in range(n)]
    return rows(out)


def fullwidth_scalars(order, seed_u64, count=5):
    """Scalar::from_u64(seed) squared `count` times in the scalar field (reference edge lists)."""
    out, v = [], seed_u64 % order
    for _ in range(count):
        v = v * v % order
        out.append(v)
    return out


def ed_edge_scalars(golden):
    L = R.L25519
    vals = list(golden["ed25519_edge_scalars_u64"]) + fullwidth_scalars(L, golden["fullwidth_seed_u64"]) + [L - 1, L - 2]
    return vals


def wei_edge_scalars(golden, n):
    vals = list(golden["weierstrass_edge_scalars_u64"]) + fullwidth_scalars(n, golden["fullwidth_seed_u64"]) + [n - 1, n - 2, n - 20]
    return vals


def ed_points(g, n):
    """n random points of the prime-order subgroup plus a few small-order / special points, affine LE."""
    base = []
    for _ in range(min(n, 24)):
        k = int.from_bytes(g.bytes(40), "little") % R.L25519
        base.append(R.ed_mul(k or 1, R.ED_B))
    pts = [base[i % len(base)] for i in range(n)]
    return rows([x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts])


def bls_cofactor_points():
    """Points of E(Fp): y^2 = x^3 + 4 outside the prime-order subgroup G1: the two points of order 3, (0, +-2), and
    sums of them with subgroup points (order 3 r).  Point::mul accepts any curve point (g1.rs:375-377), and these
    are the inputs an attacker would choose: j * P is the identity for small j."""
    c = R.WCURVES["bls12_381_g1"]
    t3 = [(0, 2), (0, c.p - 2)]
    assert all(c.on_curve(P) and c.mul(3, P) is None for P in t3)
    mixed = [c.add(c.mul(k, c.G), T) for k, T in ((5, t3[0]), (0xDEADBEEF, t3[1]))]
    return t3 + mixed


def wei_points(curve, g, n):
    c = R.WCURVES[curve]
    base = []
    for _ in range(min(n, 16)):
        k = int.from_bytes(g.bytes(60), "little") % c.n
        base.append(c.mul(k or 1, c.G))
    if curve == "bls12_381_g1" and n >= 12:   # small-order and mixed-order inputs among the first rows
        base[1:1] = bls_cofactor_points()
    return rows([c.enc(base[i % len(base)]) for i in range(n)])


def ecdsa_batch(curve, g, n, corrupt_every=4):
    """Synthetic signatures: valid ones plus corrupted r / z / Q and zero / >= n cases."""
    c = R.WCURVES[curve]
    Q, Z, RS, exp = [], [], [], []
    keys = []
    for _ in range(min(n, 8)):
        d = int.from_bytes(g.bytes(60), "little") % (c.n - 1) + 1
        keys.append((d, c.mul(d, c.G)))
    for i in range(n):
        d, q = keys[i % len(keys)]
        k = int.from_bytes(g.bytes(60), "little") % (c.n - 1) + 1
        zb = bytearray(g.bytes(c.sbytes))
        z = int.from_bytes(zb, "big")
        rs = bytearray(R.ecdsa_sign_hashed(c, d, k, z % c.n))
        qb = bytearray(c.enc(q))
        kind = i % (3 * corrupt_every)
        if kind == 1:
            rs[5] ^= 1
        elif kind == 1 + corrupt_every:
            zb[3] ^= 1
        elif kind == 1 + 2 * corrupt_every:
            rs[c.sbytes + 7] ^= 0x10
        if i == 7:
            rs[: c.sbytes] = bytes(c.sbytes)  # r = 0
        if i == 11:
            rs[c.sbytes:] = bytes(c.sbytes)  # s = 0
        if i == 13:
            rs[: c.sbytes] = c.n.to_bytes(c.sbytes, "big")  # r = n (non canonical)
        if i == 17:
            rs[c.sbytes:] = (c.n + 1).to_bytes(c.sbytes, "big") if c.n + 1 < 1 << (8 * c.sbytes) else rs[c.sbytes:]
        Q.append(bytes(qb))
        Z.append(bytes(zb))
        RS.append(bytes(rs))
    return rows(Q), rows(Z), rows(RS)


def ed25519_sig_batch(g, n):
    A, Rr, S, K = [], [], [], []
    for i in range(n):
        seed = g.bytes(32)
        msg = g.bytes(i % 50)
        pub = R.ed25519_public_from_seed(seed)
        sig = bytearray(R.ed25519_sign(seed, msg))
        if i % 5 == 1:
            sig[3] ^= 1
        if i % 5 == 2:
            sig[40] ^= 4
        if i % 5 == 3:
            msg = msg + b"x"
        if i == 9:
            sig[32:] = R.L25519.to_bytes(32, "little")  # S = l: non-canonical
        k = R.ed25519_hash_k(bytes(sig[:32]), pub, msg)
        A.append(pub)
        Rr.append(bytes(sig[:32]))
        S.append(bytes(sig[32:]))
        K.append(k)
    return rows(A), rows(Rr), rows(S), rows(K)


def sha(alg, msg):
    return getattr(hashlib, alg)(msg).digest()


def checksum(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ed25519_identity_r_cases(g):
    """Prehashed triples whose R is the identity: S = k a mod l makes [S]B - [k]A = (0, 1).
    Returns (A, R, S, K) rows: the canonical encoding of the identity (accepted), the same with the
    sign bit set (x = 0, sign = 1: decode_point rejects), and y = 1 + p (non-canonical: rejected)."""
    L = R.L25519
    A, Rr, S, K = [], [], [], []
    ident = (1).to_bytes(32, "little")
    encs = [ident, (1 | 1 << 255).to_bytes(32, "little"), (1 + R.P25519).to_bytes(32, "little")]
    for enc in encs:
        a = int.from_bytes(g.bytes(40), "little") % (L - 1) + 1
        k = int.from_bytes(g.bytes(40), "little") % L
        A.append(R.ed_encode(R.ed_mul(a, R.ED_B)))
        Rr.append(enc)
        S.append((k * a % L).to_bytes(32, "little"))
        K.append(k.to_bytes(32, "little"))
    return rows(A), rows(Rr), rows(S), rows(K)
