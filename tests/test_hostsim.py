"""Host simulation of the device code (CPU only): the very same .cuh headers that nvcc compiles for
sm_100a are compiled with g++ and an emulated carry flag (ECB_HOSTSIM) and run thread by thread,
then compared with the oracle.  This checks the primitive sequences (carry chains, reductions,
formulas, recoding, batch inversion) without a GPU; the `-m gpu` tests check the real thing."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import ed_edge_scalars, rand_bytes, rng, rows, scalars_mod, wei_edge_scalars, wei_points, ecdsa_batch, ed25519_sig_batch
from oracle import pyref as R

HS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
VP = ctypes.c_void_p


@pytest.fixture(scope="module")
def hs():
    subprocess.check_call(["make", "-s", "-j4", "-C", HS])
    f = ctypes.CDLL(os.path.join(HS, "_build", "hostsim_fields.so"))
    k = ctypes.CDLL(os.path.join(HS, "_build", "hostsim_kernels.so"))
    for fn in ("hs_ed25519_mul_base", "hs_ed25519_mul_base_ct", "hs_ed25519_mul_base_lanes", "hs_ed25519_mul", "hs_wei_mul", "hs_wei_mul_base", "hs_wei_mul_base_ct", "hs_wei_msm", "hs_wei_msm2", "hs_ecdsa_verify", "hs_bls_g1_mul_glv"):
        getattr(k, fn).restype = ctypes.c_ulonglong
    k.hs_wei_dbl.restype = ctypes.c_uint
    return f, k


def p(a):
    return None if a is None else a.ctypes.data_as(VP)


def words(x, n):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def val(a):
    return sum(int(w) << (32 * i) for i, w in enumerate(a))


def test_fe25519_ops(hs):
    f, _ = hs
    g = rng(1)
    P = R.P25519
    cases = [(0, 0), (1, P - 1), (P, P), (2**256 - 1, 2**256 - 1), (2**255 + 18, 19), (P + 5, 2**256 - 39)]
    cases += [(int.from_bytes(g.bytes(32), "little"), int.from_bytes(g.bytes(32), "little")) for _ in range(200)]
    r = np.zeros(8, dtype=np.uint32)
    for a, b in cases:
        aw, bw = words(a, 8), words(b, 8)
        for op, exp in ((0, a * b), (1, a * a), (2, a + b), (3, a - b), (4, -a), (9, a * b), (10, b * b)):
            f.hs_fe25519(op, p(aw), p(bw), p(r))
            assert val(r) % P == exp % P, (op, a, b)
        f.hs_fe25519(5, p(aw), p(bw), p(r))
        assert val(r) == a % P
        assert f.hs_fe25519_is_canonical(p(aw)) == (1 if a < P else 0)
    for a, _ in cases[:20]:
        aw = words(a, 8)
        f.hs_fe25519(6, p(aw), p(aw), p(r))
        assert val(r) % P == pow(a, P - 2, P)
        f.hs_fe25519(8, p(aw), p(aw), p(r))
        assert val(r) % P == pow(a, (P - 5) // 8, P)


def test_fe43_fp64_field(hs):
    """GF(2^255-19) on the FP64 pipe (csrc/fe43.cuh: 6 signed 43-bit limbs in doubles, products split by FMA):
    exact for every operand shape incl. all-ones limbs, loosened (unreduced sums) inputs and negative limbs,
    and the outputs stay inside the tight bound the exactness argument needs."""
    f, _ = hs
    f.hs_fe43_maxlimb.restype = ctypes.c_double
    g = rng(43)
    P = R.P25519
    top = 2**256 - 1
    cases = [(0, 0), (1, 1), (P - 1, P - 1), (top, top), (2**255 - 1, top), (P, 2), (2**43 - 1, 2**215 * (2**41 - 1))]
    cases += [(sum((2**43 - 1) << (43 * i) for i in range(0, 6, 2)) & top, sum((2**43 - 1) << (43 * i) for i in range(1, 6, 2)) & top)]
    cases += [(_structured(g, 8), _structured(g, 8)) for _ in range(150)]
    cases += [(int.from_bytes(g.bytes(32), "little"), int.from_bytes(g.bytes(32), "little")) for _ in range(300)]
    r = np.zeros(8, dtype=np.uint32)
    tight = 2.0**42 + 2.0**16
    for a, b in cases:
        aw, bw = words(a, 8), words(b, 8)
        for loosen in (0, 1, 4):
            k = loosen + 1
            for op, exp in ((0, k * a * k * b), (1, k * a * k * a), (2, k * a + k * b), (3, k * a - k * b), (4, (-k * a - b) * k * b), (5, k * a)):
                f.hs_fe43(op, p(aw), p(bw), loosen, p(r))
                assert val(r) % P == exp % P, (op, loosen, hex(a), hex(b))
            assert f.hs_fe43_maxlimb(0, p(aw), p(bw), loosen) <= tight and f.hs_fe43_maxlimb(1, p(aw), p(bw), loosen) <= tight


K256_P, K256_N = 2**256 - 2**32 - 977, 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
MONT_FIELDS = [(0, R.P256.p, 8), (1, R.P256.n, 8), (2, R.P384.p, 12), (3, R.P384.n, 12), (4, R.BLSG1.p, 12), (5, R.BLSG1.n, 8), (6, K256_P, 8), (7, K256_N, 8)]
LOOSE_FIELDS = (0, 2, 6)   # p256r1 / p384r1 / p256k1 field elements are kept "loose" (< 2^(32 n), congruent): see mont.cuh
PLAIN_FIELDS = (6,)        # secp256k1's field has no Montgomery domain (R = 1): pseudo-Mersenne folding


@pytest.mark.parametrize("field,mod,n", MONT_FIELDS)
def test_montgomery_fields(hs, field, mod, n):
    f, _ = hs
    g = rng(field)
    Rm = 1 << (32 * n)
    Rd = 1 if field in PLAIN_FIELDS else Rm      # the domain's radix
    Ri = pow(Rd, -1, mod)
    r = np.zeros(n, dtype=np.uint32)
    cases = [(0, 0), (1, mod - 1), (mod - 1, mod - 1), (Rm % mod, 1)]
    cases += [(int.from_bytes(g.bytes(48), "little") % mod, int.from_bytes(g.bytes(48), "little") % mod) for _ in range(100)]
    loose = field in LOOSE_FIELDS
    if loose:  # any n-limb value is a valid operand, including the ones that need a second fold
        K = Rm - mod
        cases += [(Rm - 1, Rm - 1), (Rm - 1, mod), (mod, mod), (Rm - 2**200, Rm - 5), (0, Rm - 1), (3, Rm - 2), (mod + 1, 2),
                  (Rm - 1, Rm - K), (Rm - K, Rm - K), (K - 1, K), (0, K), (K, Rm - 1), (Rm - K - 1, K + 1)]
        cases += [(int.from_bytes(g.bytes(4 * n), "little"), int.from_bytes(g.bytes(4 * n), "little")) for _ in range(150)]
        # structured limbs (all-ones / all-zero words) exercise the carry paths of the multiplication-free reduction
        for _ in range(150):
            pick = lambda: sum([0, 0xFFFFFFFF, 1, 0xFFFFFFFE, int(g.integers(0, 2**32))][int(g.integers(0, 5))] << (32 * i) for i in range(n))
            cases.append((pick(), pick()))
    for a, b in cases:
        aw, bw = words(a, n), words(b, n)
        for op, exp in ((0, a * b * Ri), (1, a * a * Ri), (2, a + b), (3, a - b), (4, -a), (5, a * Rd), (7, a * Ri), (13, a + b), (14, a - b), (15, -a)):
            if op == 5 and a >= mod:
                continue  # to_mont takes canonical wire values
            f.hs_mont(field, op, p(aw), p(bw), p(r))
            if loose and op != 7:
                assert val(r) % mod == exp % mod and val(r) < Rm, (field, op, a, b)
            else:
                assert val(r) == exp % mod, (field, op)
    a = cases[5][0]
    f.hs_mont(field, 6, p(words(a * Rd % mod, n)), p(words(0, n)), p(r))
    assert val(r) % mod == pow(a, -1, mod) * Rd % mod


def test_ed25519_mul_base_and_table(hs, golden, coracle):
    _, k = hs
    g = rng(2)
    kb = np.concatenate([rows([v.to_bytes(32, "little") for v in ed_edge_scalars(golden)]), scalars_mod(g, 40, R.L25519, 32, "little")])
    n = kb.shape[0]
    exp = coracle.ed25519_mul_base(kb)
    for W in (4, 7):
        ntab = k.hs_ed25519_table_entries(W)
        table = np.zeros((ntab, 24), dtype=np.uint32)
        k.hs_ed25519_build_table(W, p(table))
        out = np.zeros((n, 64), dtype=np.uint8)
        st = k.hs_ed25519_mul_base(p(kb), ctypes.c_size_t(n), W, p(table), p(out), 0)
        assert st == 2**64 - 1 and np.array_equal(out, exp)
        # the fused small-batch kernel's lane split: windows l, l + lanes, ... per lane, butterfly of complete additions
        for lanes in (1, 2, 4, 8):
            out2 = np.zeros((n, 64), dtype=np.uint8)
            st = k.hs_ed25519_mul_base_lanes(p(kb), ctypes.c_size_t(n), W, p(table), lanes, p(out2))
            assert st == 2**64 - 1 and np.array_equal(out2, exp), (W, lanes)
    bad = kb.copy()
    bad[3] = np.frombuffer(R.L25519.to_bytes(32, "little"), dtype=np.uint8)
    st = k.hs_ed25519_mul_base(p(bad), ctypes.c_size_t(n), W, p(table), p(out), 0)
    assert st == (3 << 8) | 1


def test_ed25519_mul_base_constant_time_form(hs, golden, coracle):
    """ct.cuh: every window scans all 8 entries of the W = 4 comb with masks (the reference's select_from_table,
    curve25519.rs:862-869) and the inversion is the Fermat chain — same bytes as the variable-time comb, on the
    reference's edge scalars (0, 1, l - 1, window boundaries) and random ones, affine and compressed."""
    _, k = hs
    g = rng(21)
    kb = np.concatenate([rows([v.to_bytes(32, "little") for v in ed_edge_scalars(golden)]), scalars_mod(g, 60, R.L25519, 32, "little")])
    n = kb.shape[0]
    table = np.zeros((k.hs_ed25519_table_entries(4), 24), dtype=np.uint32)
    assert table.shape[0] == 64 * 8
    k.hs_ed25519_build_table(4, p(table))
    out = np.zeros((n, 64), dtype=np.uint8)
    st = k.hs_ed25519_mul_base_ct(p(kb), ctypes.c_size_t(n), p(table), p(out), 0)
    exp = coracle.ed25519_mul_base(kb)
    assert st == 2**64 - 1 and np.array_equal(out, exp)
    enc = np.zeros((n, 32), dtype=np.uint8)
    k.hs_ed25519_mul_base_ct(p(kb), ctypes.c_size_t(n), p(table), p(enc), 1)
    want = exp[:, 32:].copy()
    want[:, 31] |= (exp[:, 0] & 1) << 7
    assert np.array_equal(enc, want)


def test_x25519_base_via_comb(hs, golden, coracle):
    """x25519_base(k) = x25519(k, 9) computed as the Montgomery u of clamp(k) * B on edwards25519."""
    _, k = hs
    g = rng(77)
    n = 40
    ks = rand_bytes(g, n, 32)
    ks[0] = 0; ks[1] = 0xFF
    ks[2] = np.frombuffer(bytes.fromhex(golden["x25519"]["iterated_once"]["k"]), dtype=np.uint8)
    ks[3] = np.frombuffer(bytes.fromhex(golden["x25519"]["dh_6_1"]["a"]), dtype=np.uint8)
    nine = np.tile(np.frombuffer((9).to_bytes(32, "little"), dtype=np.uint8), (n, 1))
    exp = coracle.x25519(ks, nine)
    assert exp[2].tobytes().hex() == golden["x25519"]["iterated_once"]["r"]
    for W in (4, 7):
        table = np.zeros((k.hs_ed25519_table_entries(W), 24), dtype=np.uint32)
        k.hs_ed25519_build_table(W, p(table))
        out = np.zeros((n, 32), dtype=np.uint8)
        k.hs_x25519_base(p(ks), ctypes.c_size_t(n), W, p(table), p(out))
        assert np.array_equal(out, exp)


def test_ed25519_mul_x25519_x448(hs, coracle):
    _, k = hs
    g = rng(3)
    n = 24
    kb = scalars_mod(g, n, R.L25519, 32, "little")
    from helpers import ed_points

    pts = ed_points(g, n)
    out = np.zeros((n, 64), dtype=np.uint8)
    assert k.hs_ed25519_mul(p(kb), p(pts), ctypes.c_size_t(n), p(out)) == 2**64 - 1
    assert np.array_equal(out, coracle.ed25519_mul(kb, pts))
    ks, us = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
    us[0] = 0; us[1] = 0xFF
    o = np.zeros((n, 32), dtype=np.uint8)
    k.hs_x25519(p(ks), p(us), ctypes.c_size_t(n), p(o))
    assert np.array_equal(o, coracle.x25519(ks, us))
    n = 6
    ks, us = rand_bytes(g, n, 56), rand_bytes(g, n, 56)
    us[0] = 0xFF
    o = np.zeros((n, 56), dtype=np.uint8)
    k.hs_x448(p(ks), p(us), ctypes.c_size_t(n), p(o))
    assert np.array_equal(o, coracle.x448(ks, us))


@pytest.mark.parametrize("cid,curve,n,am3", [(0, "p256r1", 8, True), (1, "p384r1", 12, True), (2, "bls12_381_g1", 12, False), (3, "p256k1", 8, False)])
def test_wei_doubling_lazy_folds(hs, cid, curve, n, am3):
    """WeiJ::dbl_z<true> (weier.cuh): the field additions of a doubling only COUNT a carry out of their fold; the window
    loop (kernels.cuh wei_window_loop) is redone with the checked forms if any counter is set at its end.  The formulas are
    polynomial maps, so ANY coordinate words are valid inputs: random ones, and loose representatives at the top of the
    range, where folds do carry.  The checked forms (dbl, dbl_z<false>) and the lazy one are compared with the formulas
    evaluated on big integers: the lazy form must be right whenever its counters are zero, the cases must include some
    where they are not, and the checked forms never count."""
    _, k = hs
    c = R.WCURVES[curve]
    p_ = c.p
    Rm = 1 << (32 * n)
    loose = cid in (0, 1, 3)
    Rd = 1 if cid == 3 else Rm                     # domain radix: products are a b / Rd
    Ri = pow(Rd, -1, p_)
    g = rng(930 + cid)

    def expect(X, Y, Z):
        m = lambda a, b: a * b * Ri % p_           # the field product in the domain
        if am3:
            delta, gamma = m(Z, Z), m(Y, Y)
            beta = m(X, gamma)
            alpha = 3 * m(X - delta, X + delta)
            Z3 = m(Y + Z, Y + Z) - gamma - delta
            X3 = m(alpha, alpha) - 8 * beta
            Y3 = m(alpha, 4 * beta - X3) - 8 * m(gamma, gamma)
        else:
            A, B = m(X, X), m(Y, Y)
            Cc = m(B, B)
            D = 2 * (m(X + B, X + B) - A - Cc)
            E = 3 * A
            X3 = m(E, E) - 2 * D
            Z3 = 2 * m(Y, Z)
            Y3 = m(E, D - X3) - 8 * Cc
        return X3 % p_, Y3 % p_, Z3 % p_

    top = Rm if loose else p_
    K = Rm - p_
    edge = [top - 1, top - 2, 0, 1, p_ - 1, (top - K) % top, (top - K - 1) % top, (top - 2 * K) % top, top >> 1, (top >> 1) + 1, top - (top >> 3), (top >> 2) + 5]
    cases = [(a, b, d) for a in edge[:8] for b in edge[:6] for d in edge[:6]]
    cases += [tuple(edge[int(g.integers(0, len(edge)))] for _ in range(3)) for _ in range(300)]
    cases += [tuple(int.from_bytes(g.bytes(4 * n), "little") % top for _ in range(3)) for _ in range(300)]
    out = np.zeros(3 * n, dtype=np.uint32)
    flagged = 0
    for X, Y, Z in cases:
        inp = np.concatenate([words(X, n), words(Y, n), words(Z, n)])
        want = expect(X, Y, Z)
        for which in (0, 1, 2):
            fl = k.hs_wei_dbl(cid, p(inp), p(out), which)
            got = tuple(val(out[i * n:(i + 1) * n]) for i in range(3))
            if which == 2 and fl:
                flagged += 1
                continue
            assert which == 2 or fl == 0
            assert all(v < top for v in got), (which, hex(X), hex(Y), hex(Z))
            assert tuple(v % p_ for v in got) == want, (which, hex(X), hex(Y), hex(Z))
    assert (flagged > 0) == loose, flagged


@pytest.mark.parametrize("cid,curve", [(0, "p256r1"), (1, "p384r1"), (2, "bls12_381_g1"), (3, "p256k1")])
def test_wei_mul(hs, golden, coracle, cid, curve):
    _, k = hs
    c = R.WCURVES[curve]
    g = rng(10 + cid)
    kb = np.concatenate([rows([v.to_bytes(c.sbytes, "big") for v in wei_edge_scalars(golden, c.n)] + [(c.n - j).to_bytes(c.sbytes, "big") for j in range(1, 21)]), scalars_mod(g, 6, c.n, c.sbytes, "big")])
    n = kb.shape[0]
    pts = wei_points(curve, g, n)
    out = np.zeros((n, 2 * c.fbytes), dtype=np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    assert k.hs_wei_mul(cid, p(kb), p(pts), None, ctypes.c_size_t(n), p(out), p(inf)) == 2**64 - 1
    exp, einf = coracle.wei_mul(curve, kb, pts)
    assert np.array_equal(out, exp) and np.array_equal(inf.astype(bool), einf)
    # fixed base: the signed-digit generator comb, two window widths
    exp, einf = coracle.wei_mul_base(curve, kb)
    for W in (4, 5):
        assert k.hs_wei_mul_base(cid, p(kb), ctypes.c_size_t(n), W, p(out), p(inf)) == 2**64 - 1
        assert np.array_equal(out, exp) and np.array_equal(inf.astype(bool), einf)


@pytest.mark.parametrize("cid,curve", [(0, "p256r1"), (1, "p384r1")])
def test_weierstrass_mul_base_constant_time_form(hs, golden, coracle, cid, curve):
    """ct.cuh on p256r1 / p384r1: masked scans of the W = 4 comb and the complete projective addition (RCB Algorithm 4,
    the reference's own formulas, projective.rs:340-423) — same bytes as the reference's mul_base for the NIST KAT
    scalars, the edge scalars (0 -> infinity, 1, n - 1, n - 2) and the scalars that hit the Jacobian exceptional
    cases in a signed radix-16 comb (2^257 mod n and neighbours), plus random ones."""
    _, k = hs
    c = R.WCURVES[curve]
    g = rng(400 + cid)
    vals = [0, 1, 2, 15, 16, 17, c.n - 1, c.n - 2, (1 << (8 * c.sbytes + 1)) % c.n, ((1 << (8 * c.sbytes + 1)) + 1) % c.n, (1 << (8 * c.sbytes - 4)) % c.n]
    vals += [v % c.n for v in wei_edge_scalars(golden, c.n)[:24]]
    kb = np.concatenate([rows([v.to_bytes(c.sbytes, "big") for v in vals]), scalars_mod(g, 24, c.n, c.sbytes, "big")])
    n = kb.shape[0]
    out = np.zeros((n, 2 * c.fbytes), dtype=np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    assert k.hs_wei_mul_base_ct(cid, p(kb), ctypes.c_size_t(n), p(out), p(inf)) == 2**64 - 1
    exp, einf = coracle.wei_mul_base(curve, kb)
    assert np.array_equal(inf.astype(bool), einf) and np.array_equal(out, exp) and einf[0] and not einf[-24:].any()


@pytest.mark.parametrize("cid,curve", [(0, "p256r1"), (1, "p384r1")])
def test_ecdsa_sign(hs, golden, coracle, cid, curve):
    """Device code of sign_hashed (ecdsa.rs:165-184) against the RFC 6979 vectors (d, k, message -> r, s;
    ecdsa.rs:808-878), the big-int oracle on random (d, k, z), and the refusals: zero secret, zero nonce,
    non-canonical scalars.  Every produced signature verifies (C oracle)."""
    import hashlib

    _, k = hs
    c = R.WCURVES[curve]
    v = golden["ecdsa_rfc6979"][curve]
    d0 = int(v["d"], 16)
    ds, ks, zs, want = [], [], [], []
    for kat in v["kats"]:
        dg = hashlib.new(kat["alg"], kat["message"].encode()).digest()
        z = int.from_bytes(R.ecdsa_digest_to_scalar(c, dg), "big")
        ds.append(d0)
        ks.append(int(kat["k"], 16))
        zs.append(z)
        want.append(kat["r"].rjust(2 * c.sbytes, "0") + kat["s"].rjust(2 * c.sbytes, "0"))
    g = rng(90 + cid)
    for _ in range(12):
        ds.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1)
        ks.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1)
        zs.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % c.n)
    nk = len(v["kats"])
    bad = [(0, 5, 7), (5, 0, 7), (c.n, 5, 7), (5, c.n, 7), (5, 6, c.n)]
    for d, kk, z in bad:
        ds.append(d)
        ks.append(kk)
        zs.append(z)
    n = len(ds)
    tob = lambda xs: rows([x.to_bytes(c.sbytes, "big") for x in xs])
    db, kb, zb = tob(ds), tob(ks), tob(zs)
    rs = np.zeros((n, 2 * c.sbytes), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    k.hs_ecdsa_sign(cid, p(db), p(kb), p(zb), ctypes.c_size_t(n), p(rs), p(ok))
    assert list(ok) == [1] * (n - len(bad)) + [0] * len(bad) and not rs[n - len(bad):].any()
    for i in range(nk):
        assert rs[i].tobytes().hex() == want[i]
    for i in range(n - len(bad)):
        assert rs[i].tobytes() == R.ecdsa_sign_hashed(c, ds[i], ks[i], zs[i])
    m = n - len(bad)
    q = rows([c.enc(c.mul(d, c.G)) for d in ds[:m]])
    assert coracle.ecdsa_verify_hashed(curve, q, zb[:m], rs[:m]).all()


def test_ed25519_keygen_and_sign(hs, golden):
    """Device code of expand_secret / public_from_seed / sign_with_public (ed25519.rs:61-110) against
    RFC 8032 TEST 1-3 (seed, public key, message, signature: ed25519.rs:271-290) and the big-int oracle
    on random seeds with message lengths around the SHA-512 padding boundaries."""
    _, k = hs
    W = 6
    table = np.zeros((k.hs_ed25519_table_entries(W), 24), dtype=np.uint32)
    k.hs_ed25519_build_table(W, p(table))
    g = rng(70)
    seeds = [bytes.fromhex(v["seed"]) for v in golden["ed25519_rfc8032"]]
    msgs = [bytes.fromhex(v["message"]) for v in golden["ed25519_rfc8032"]]
    for ln in (0, 1, 47, 48, 79, 80, 111, 112, 175, 176, 300):
        seeds.append(g.bytes(32))
        msgs.append(g.bytes(ln))
    n = len(seeds)
    sd = rows(seeds)
    pub = np.zeros((n, 32), dtype=np.uint8)
    k.hs_ed25519_public_from_seed(p(sd), ctypes.c_size_t(n), W, p(table), p(pub))
    for i, v in enumerate(golden["ed25519_rfc8032"]):
        assert pub[i].tobytes().hex() == v["public"]
    assert [r.tobytes() for r in pub] == [R.ed25519_public_from_seed(s) for s in seeds]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    blob = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    sig = np.zeros((n, 64), dtype=np.uint8)
    k.hs_ed25519_sign(p(sd), p(pub), p(blob), p(off), ctypes.c_size_t(n), W, p(table), p(sig))
    for i, v in enumerate(golden["ed25519_rfc8032"]):
        assert sig[i].tobytes().hex() == v["signature"]
    for i in range(n):
        assert sig[i].tobytes() == R.ed25519_sign(seeds[i], msgs[i])
        assert R.ed25519_verify(pub[i].tobytes(), msgs[i], sig[i].tobytes())


@pytest.mark.parametrize("cid,curve", [(0, "p256r1"), (1, "p384r1"), (2, "bls12_381_g1"), (3, "p256k1")])
def test_wei_decompress(hs, coracle, cid, curve):
    """Device code of PointAffine::decompress (kernels3.cuh) against the oracle: both parities, an x
    with a non-square right-hand side, x = 0, x >= p."""
    _, k = hs
    c = R.WCURVES[curve]
    g = rng(60 + cid)
    n = 40
    pts = wei_points(curve, g, n)
    x = np.ascontiguousarray(pts[:, : c.fbytes])
    sign = (g.integers(0, 2, n)).astype(np.uint8)
    x[n - 3] = 0
    x[n - 2] = np.frombuffer(c.p.to_bytes(c.fbytes, "big"), dtype=np.uint8)
    x[n - 1] = 0xFF
    for i in range(0, 12):  # random field elements: about half are not x-coordinates
        x[i] = np.frombuffer((int.from_bytes(g.bytes(c.fbytes + 8), "big") % c.p).to_bytes(c.fbytes, "big"), dtype=np.uint8)
    out = np.zeros((n, 2 * c.fbytes), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    k.hs_wei_decompress(cid, p(x), p(sign), ctypes.c_size_t(n), p(out), p(ok))
    exp, eok = coracle.wei_decompress(curve, x, sign)
    assert np.array_equal(ok.astype(bool), eok) and np.array_equal(out, exp)
    assert eok.any() and not eok.all()
    for i in (0, 5, 20):
        want = R.wei_decompress(curve, x[i].tobytes(), int(sign[i]))
        assert (out[i].tobytes() if ok[i] else None) == want


def test_bls_g1_encodings(hs, coracle, golden):
    """Device code of from_compressed / to_compressed / is_in_subgroup against the reference's KATs
    (g1.rs:605-680, OFF_SUBGROUP :313-368) and the oracle on random curve points inside and outside G1."""
    _, k = hs
    c = R.BLSG1
    v = golden["bls12_381_g1"]
    encs = [bytes.fromhex(e["bytes"]) for e in v["compressed"]] + [bytes.fromhex(o["compressed"]) for o in v["off_subgroup"]]
    x = 300
    while len(encs) < 16:  # raw curve points (outside G1) and their cofactor-cleared images (inside)
        x += 1
        rhs = (x**3 + 4) % c.p
        y = pow(rhs, (c.p + 1) // 4, c.p)
        if y * y % c.p == rhs:
            encs.append(R.bls_g1_to_compressed(c.enc((x, y))))
            encs.append(R.bls_g1_to_compressed(c.enc(c.mul(0xd201000000010001, (x, y)))))
    g0 = encs[0]
    encs += [bytes([g0[0] & 0x7F]) + g0[1:], bytes([0xC0]) + bytes(47), bytes([0xE0]) + bytes(47), bytes([g0[0] ^ 0x20]) + g0[1:],
             bytes([0x80 | (c.p >> 376)]) + (c.p & ((1 << 376) - 1)).to_bytes(47, "big")]
    e = rows(encs)
    n = e.shape[0]
    for check in (1, 0):
        out = np.zeros((n, 96), dtype=np.uint8)
        ok = np.zeros(n, dtype=np.uint8)
        k.hs_bls_g1_from_compressed(p(e), ctypes.c_size_t(n), check, p(out), p(ok))
        exp, eok = coracle.bls12_381_g1_from_compressed(e, bool(check))
        assert np.array_equal(ok.astype(bool), eok) and np.array_equal(out, exp), check
        for i in range(n):
            assert (out[i].tobytes() if ok[i] else None) == R.bls_g1_from_compressed(encs[i], bool(check)), (i, check)
    assert list(ok[5:8]) == [1, 1, 1]                    # OFF_SUBGROUP decodes with _oncurve_only ...
    uncompressed = [bytes.fromhex(o["uncompressed"]) for o in v["off_subgroup"]]
    assert [out[5 + i].tobytes() for i in range(3)] == uncompressed
    # to_compressed inverts it, identity included
    good = ok.astype(bool)
    inf = np.zeros(n, dtype=np.uint8)
    inf[-1] = 1
    enc2 = np.zeros((n, 48), dtype=np.uint8)
    k.hs_bls_g1_to_compressed(p(out), p(inf), ctypes.c_size_t(n), p(enc2))
    assert np.array_equal(enc2[good], e[good])
    assert enc2[-1].tobytes() == bytes([0xC0]) + bytes(47)
    assert np.array_equal(enc2, coracle.bls12_381_g1_to_compressed(out, inf))


@pytest.mark.parametrize("cid,curve", [(0, "p256r1"), (1, "p384r1")])
def test_ecdsa(hs, coracle, cid, curve):
    _, k = hs
    g = rng(20 + cid)
    q, z, rs = ecdsa_batch(curve, g, 24)
    ok = np.zeros(24, dtype=np.uint8)
    assert k.hs_ecdsa_verify(cid, p(q), p(z), p(rs), ctypes.c_size_t(24), p(ok)) == 2**64 - 1
    assert np.array_equal(ok.astype(bool), coracle.ecdsa_verify_hashed(curve, q, z, rs))


def test_ed25519_verify(hs, coracle):
    _, k = hs
    g = rng(30)
    a, r, s, kk = ed25519_sig_batch(g, 20)
    W = 5
    table = np.zeros((k.hs_ed25519_table_entries(W), 24), dtype=np.uint32)
    k.hs_ed25519_build_table(W, p(table))
    ok = np.zeros(20, dtype=np.uint8)
    k.hs_ed25519_verify(p(a), p(r), p(s), p(kk), ctypes.c_size_t(20), W, p(table), p(ok))
    assert np.array_equal(ok.astype(bool), coracle.ed25519_verify_prehashed(a, r, s, kk))
    # R = identity: canonical encoding accepted, (x = 0, sign = 1) and non-canonical y rejected — the
    # kernel compares encodings instead of decoding R, which must give the same three answers
    from helpers import ed25519_identity_r_cases

    a, r, s, kk = ed25519_identity_r_cases(g)
    ok = np.zeros(3, dtype=np.uint8)
    k.hs_ed25519_verify(p(a), p(r), p(s), p(kk), ctypes.c_size_t(3), W, p(table), p(ok))
    assert ok.tolist() == [1, 0, 0]
    assert coracle.ed25519_verify_prehashed(a, r, s, kk).tolist() == [True, False, False]
    assert [R.ed25519_verify_prehashed(a[i].tobytes(), r[i].tobytes(), s[i].tobytes(), kk[i].tobytes()) for i in range(3)] == [True, False, False]


def _structured(g, n):
    """Operands made of the limb values that break carry handling: 0, 1, ff..f, ff..e and random."""
    choices = [0, 1, 0xFFFFFFFF, 0xFFFFFFFE]
    return sum(int(choices[int(c)] if c < 4 else int(g.integers(0, 1 << 32))) << (32 * i) for i, c in enumerate(g.integers(0, 5, size=n)))


@pytest.mark.parametrize("field,mod,n", MONT_FIELDS)
def test_montgomery_fields_structured_operands(hs, field, mod, n):
    """A dropped top carry in the p384 reduction only showed with limbs like (1, 1, ff..f): random
    operands never hit it.  Every field gets the structured stress."""
    f, _ = hs
    g = rng(100 + field)
    Rm = 1 << (32 * n)
    Ri = pow(1 if field in PLAIN_FIELDS else Rm, -1, mod)
    loose = field in (0, 6)
    r = np.zeros(n, dtype=np.uint32)
    for _ in range(1500):
        a = _structured(g, n) % (Rm if loose else mod)
        b = _structured(g, n) % (Rm if loose else mod)
        for op, exp, bb in ((0, a * b * Ri, b), (1, a * a * Ri, a), (2, a + b, b), (3, a - b, b)):
            f.hs_mont(field, op, p(words(a, n)), p(words(bb, n)), p(r))
            if loose:
                assert val(r) % mod == exp % mod and val(r) < Rm, (field, op, hex(a), hex(b))
            else:
                assert val(r) == exp % mod, (field, op, hex(a), hex(b))


@pytest.mark.parametrize("field,mod,n", [(0, R.P256.p, 8), (2, R.P384.p, 12), (4, R.BLSG1.p, 12), (5, R.BLSG1.n, 8), (6, K256_P, 8)])
def test_montgomery_merged_forms(hs, field, mod, n):
    """K a (K = 2, 3, 4, 8), a - b - c and a - b - 2c in one pass with a single fold of the accumulated carries /
    borrows (mont.cuh: mul_small, sub2, sub_2x) — on the loose fields every n-limb operand is valid, including the
    ones whose fold carries again; on the canonical fields the plain sequences must give canonical results."""
    f, _ = hs
    g = rng(300 + field)
    Rm = 1 << (32 * n)
    loose = field in LOOSE_FIELDS
    top = Rm if loose else mod
    K = Rm - mod
    r = np.zeros(n, dtype=np.uint32)
    cases = []
    if loose:
        edge = [0, 1, K - 1, K, K + 1, mod - 1, mod, mod + 1, Rm - 1, Rm - 2, Rm - K, Rm - K - 1, Rm - K + 1, Rm - 2 * K, Rm - 3 * K, Rm - 8 * K,
                (Rm >> 1) - 1, Rm >> 1, (Rm >> 2) + 1, (Rm >> 3) - 1, Rm - (Rm >> 3), Rm - (Rm >> 2), 2 * K, 3 * K, 7 * K, 8 * K]
        # values for which K a mod 2^(32 n) lands within a few multiples of 2^256 - p of the top (second fold)
        for k in (2, 3, 4, 8):
            for j in range(1, k):
                for d in (-2 * K, -K - 1, -K, -K + 1, -1, 0, 1):
                    edge.append(((j * Rm + Rm + d) // k) % Rm)
    else:
        edge = [0, 1, 2, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1, mod // 3, mod // 3 + 1, mod // 8, mod - mod // 8]
    for a in edge:
        for b in edge[:12]:
            cases.append((a, b, edge[(a + b) % len(edge)]))
    for _ in range(400):
        cases.append(tuple(_structured(g, n) % top for _ in range(3)))
    for _ in range(400):
        cases.append(tuple(int.from_bytes(g.bytes(4 * n), "little") % top for _ in range(3)))

    def check(exp, what):
        if loose:
            assert val(r) % mod == exp % mod and val(r) < Rm, what
        else:
            assert val(r) == exp % mod, what

    for a, b, c in cases:
        aw, bw, cw = words(a, n), words(b, n), words(c, n)
        for op, k in ((8, 2), (9, 3), (10, 4), (11, 8), (12, 8)):
            f.hs_mont(field, op, p(aw), p(bw), p(r))
            check(k * a, (field, op, hex(a)))
        for op, exp in ((0, a - b - c), (1, a - b - 2 * c), (2, a - b - c), (3, a - b - 2 * c)):
            f.hs_mont3(field, op, p(aw), p(bw), p(cw), p(r))
            check(exp, (field, op, hex(a), hex(b), hex(c)))


def test_fe25519_and_fe448_structured_operands(hs):
    f, k = hs
    g = rng(200)
    r8, r14 = np.zeros(8, dtype=np.uint32), np.zeros(14, dtype=np.uint32)
    for _ in range(1500):
        a, b = _structured(g, 8), _structured(g, 8)
        for op, exp in ((0, a * b), (1, a * a), (2, a + b), (3, a - b)):
            f.hs_fe25519(op, p(words(a, 8)), p(words(b, 8)), p(r8))
            assert val(r8) % R.P25519 == exp % R.P25519, (op, hex(a), hex(b))
        f.hs_fe25519(5, p(words(a, 8)), p(words(b, 8)), p(r8))
        assert val(r8) == a % R.P25519
        a, b = _structured(g, 14), _structured(g, 14)
        for op, exp in ((0, a * b), (1, a * a), (2, a + b), (3, a - b)):
            k.hs_fe448(op, p(words(a, 14)), p(words(b, 14)), p(r14))
            assert val(r14) % R.P448 == exp % R.P448, (op, hex(a), hex(b))
        k.hs_fe448(5, p(words(a, 14)), p(words(b, 14)), p(r14))
        assert val(r14) == a % R.P448


def test_sha512_and_ed25519_challenge(hs):
    """Device SHA-512 and k = SHA-512(R || A || M) mod l against hashlib / the big-int oracle."""
    import hashlib

    _, k = hs
    g = rng(512)
    dg = np.zeros(64, dtype=np.uint8)
    for ln in list(range(0, 20)) + [55, 56, 63, 64, 110, 111, 112, 113, 127, 128, 129, 239, 240, 241, 255, 256, 300, 1000]:
        m = np.frombuffer(g.bytes(ln) if ln else b"", dtype=np.uint8).copy() if ln else np.zeros(0, dtype=np.uint8)
        buf = np.concatenate([m, np.zeros(1, dtype=np.uint8)])  # non-empty buffer for ctypes
        k.hs_sha512(p(buf), ctypes.c_size_t(ln), p(dg))
        assert dg.tobytes() == hashlib.sha512(m.tobytes()).digest(), ln
    n = 40
    lens = [0, 1, 47, 48, 63, 64, 65, 111, 112, 175, 176, 177, 300] + [int(x) for x in g.integers(0, 400, size=n - 13)]
    msgs = [g.bytes(l) if l else b"" for l in lens]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    blob = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    a = rand_bytes(g, n, 32)
    sig = rand_bytes(g, n, 64)
    r = np.zeros((n, 32), dtype=np.uint8); s = np.zeros((n, 32), dtype=np.uint8); kk = np.zeros((n, 32), dtype=np.uint8)
    k.hs_ed25519_hash_k(p(a), p(sig), p(blob), p(off), ctypes.c_size_t(n), p(r), p(s), p(kk))
    for i in range(n):
        assert kk[i].tobytes() == R.ed25519_hash_k(sig[i, :32].tobytes(), a[i].tobytes(), msgs[i]), i
    assert np.array_equal(r, sig[:, :32]) and np.array_equal(s, sig[:, 32:])


def test_ecdsa_digest_to_scalar_on_device_code(hs):
    """SHA-256 / SHA-384 / SHA-512 + bits2int (left-pad or keep the leading bytes) vs hashlib / the oracle."""
    import hashlib

    _, k = hs
    g = rng(256)
    lens = [0, 1, 55, 56, 63, 64, 65, 111, 112, 119, 120, 127, 128, 129, 500] + [int(x) for x in g.integers(0, 300, size=25)]
    n = len(lens)
    msgs = [g.bytes(l) if l else b"" for l in lens]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    blob = np.frombuffer(b"".join(msgs) + b"\0", dtype=np.uint8).copy()
    for curve, sb in (("p256r1", 32), ("p384r1", 48)):
        c = R.WCURVES[curve]
        for hname, hid in (("sha256", 256), ("sha384", 384), ("sha512", 512)):
            z = np.zeros((n, sb), dtype=np.uint8)
            k.hs_ecdsa_hash_z(p(blob), p(off), ctypes.c_size_t(n), hid, sb, p(z))
            for i in range(n):
                d = hashlib.new(hname, msgs[i]).digest()
                want = d[:sb] if len(d) >= sb else bytes(sb - len(d)) + d
                assert z[i].tobytes() == want, (curve, hname, i)
                # after the reduction mod n this is the reference's digest_to_scalar
                assert (int.from_bytes(want, "big") % c.n).to_bytes(sb, "big") == R.ecdsa_digest_to_scalar(c, d)


@pytest.mark.parametrize("field,mod,n", MONT_FIELDS)
def test_safegcd_inversion_montgomery_fields(hs, field, mod, n):
    """Field inverses come from safegcd divsteps (csrc/modinv.cuh), not from a Fermat chain: check
    the whole range of operand shapes, including 0 -> 0."""
    f, _ = hs
    g = rng(300 + field)
    Rm = 1 << (32 * n)
    r = np.zeros(n, dtype=np.uint32)
    vals = [1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, 1 << 200, (1 << (32 * n - 1)) % mod, Rm % mod]
    vals += [_structured(g, n) % mod for _ in range(150)] + [int.from_bytes(g.bytes(64), "little") % mod for _ in range(150)]
    for a in vals:
        if a == 0:
            continue
        Rd = 1 if field in PLAIN_FIELDS else Rm
        f.hs_mont(field, 6, p(words(a * Rd % mod, n)), p(words(0, n)), p(r))
        assert val(r) % mod == pow(a, -1, mod) * Rd % mod, hex(a)
    f.hs_mont(field, 6, p(words(0, n)), p(words(0, n)), p(r))
    assert val(r) == 0


def test_safegcd_inversion_25519_and_448(hs):
    f, k = hs
    g = rng(400)
    r8, r14 = np.zeros(8, dtype=np.uint32), np.zeros(14, dtype=np.uint32)
    P, Q = R.P25519, R.P448
    for a in [1, 2, P - 1, P, P + 1, 2**256 - 1, 19, 38, 2**255] + [_structured(g, 8) for _ in range(150)] + [int.from_bytes(g.bytes(32), "little") for _ in range(150)]:
        f.hs_fe25519(6, p(words(a, 8)), p(words(0, 8)), p(r8))
        assert val(r8) % P == (pow(a % P, -1, P) if a % P else 0), hex(a)
    for a in [1, 2, Q - 1, Q, Q + 1, 2**448 - 1, 2**224, 2**447] + [_structured(g, 14) for _ in range(100)] + [int.from_bytes(g.bytes(56), "little") for _ in range(100)]:
        k.hs_fe448(6, p(words(a, 14)), p(words(0, 14)), p(r14))
        assert val(r14) % Q == (pow(a % Q, -1, Q) if a % Q else 0), hex(a)


def test_divsteps_variants_agree(hs):
    """The three forms of 30 divsteps (modinv.cuh: one step at a time; zero runs shifted out by ctz with up to four
    bits cancelled per trip; four steps per lookup in the generated jump table SG_JUMP4, which the warp-level
    inversion uses) give the same transition matrix and the same zeta — on random words, at every zeta around the
    clamp boundaries of the table, and along the trajectories of real inversions."""
    f, _ = hs
    g = rng(77)
    t = [np.zeros(4, dtype=np.int32) for _ in range(3)]
    cases = []
    for _ in range(3000):
        z = int(g.integers(-40, 40))
        cases.append((z, int(g.integers(0, 1 << 30)) | 1, int(g.integers(0, 1 << 30))))
    for z in (-31, -30, -8, -7, -6, -5, -4, -3, -2, -1, 0, 1, 2, 3, 4, 5, 6, 29, 30, 31):
        for g0 in (0, 1, 2, 4, 8, 1 << 29, (1 << 30) - 1, (1 << 30) - 2):
            cases.append((z, (1 << 30) - 19, g0))
    P = R.P25519
    for _ in range(40):
        a, fv, zeta = int.from_bytes(g.bytes(32), "little") % P, P, -1
        while a:
            f0, g0 = fv & 0x3FFFFFFF, a & 0x3FFFFFFF
            cases.append((zeta, f0, g0))
            zeta = f.hs_divsteps30(0, zeta, f0, g0, p(t[0]))
            u, v, q, r = (int(x) for x in t[0])
            fv, a = (u * fv + v * a) >> 30, (q * fv + r * a) >> 30
    for z, f0, g0 in cases:
        zs = [f.hs_divsteps30(k, z, f0, g0, p(t[k])) for k in range(3)]
        assert zs[0] == zs[1] == zs[2], (z, f0, g0)
        assert np.array_equal(t[0], t[1]) and np.array_equal(t[0], t[2]), (z, f0, g0)


def test_ristretto255_encodings(hs, golden):
    """ristretto.cuh against RFC 9496 as the reference holds it (ristretto255.rs:341-398): the encodings of 0 B .. 15 B,
    the 17 encodings that must be refused, round trips, and compress(P + T) == compress(P) for the 8-torsion T
    (equality independent of the representative, :466) — bit-exact with the big-integer oracle."""
    _, k = hs
    v = golden["ristretto255"]
    mult = rows([bytes.fromhex(x) for x in v["multiples"]])
    bad = rows([bytes.fromhex(x) for x in v["bad"]])
    enc = np.concatenate([mult, bad])
    n = enc.shape[0]
    xy = np.zeros((n, 64), dtype=np.uint8)
    ok = np.zeros(n, dtype=np.uint8)
    k.hs_ristretto255_decompress(p(enc), ctypes.c_size_t(n), p(xy), p(ok))
    assert ok[:16].all() and not ok[16:].any() and not xy[16:].any()
    for i in range(16):
        P = R.ristretto255_decompress(enc[i].tobytes())
        assert xy[i].tobytes() == P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")
    back = np.zeros((16, 32), dtype=np.uint8)
    k.hs_ristretto255_compress(p(xy[:16].copy()), ctypes.c_size_t(16), p(back))
    assert np.array_equal(back, mult)
    # multiples of B computed on edwards25519 (any representative) compress to the RFC's encodings
    pts = [R.ed_mul(i, R.ED_B) if i else R.ED_ID for i in range(16)]
    t4 = (R.SQRT_M1, 0)                                        # a point of order 4
    pts2 = [R.ed_add(P, t4) for P in pts] + [R.ed_add(P, (0, R.P25519 - 1)) for P in pts]
    allp = rows([x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts + pts2])
    out = np.zeros((48, 32), dtype=np.uint8)
    k.hs_ristretto255_compress(p(allp), ctypes.c_size_t(48), p(out))
    assert np.array_equal(out[:16], mult) and np.array_equal(out[16:32], mult) and np.array_equal(out[32:], mult)


def test_bls12_381_g1_mul_through_the_endomorphism(hs, coracle):
    """Option bls12_381_g1_glv (kernels3.cuh): k = q x^2 + rem exactly, both parts below 2^128, for edge and random
    scalars; k P = rem P - q phi(P) equals the plain window kernel, the C oracle and the big-integer group law for
    points of G1 (edge scalars 0, 1, r - 1, x^2, x^2 +- 1, multiples of x^2, identity inputs); a non-canonical scalar
    and an off-curve point are refused with their index as in the plain kernel."""
    _, k = hs
    c = R.WCURVES["bls12_381_g1"]
    g = rng(910)
    xsq = 0xD201000000010000 ** 2
    ks = [0, 1, 2, c.n - 1, c.n - 2, xsq - 1, xsq, xsq + 1, 2 * xsq, xsq * xsq % c.n, (xsq - 1) * xsq, (1 << 128) - 1, 1 << 128, (1 << 254) + 1,
          xsq * ((1 << 127) - 1) % c.n, 0xFFFFFFFF, 1 << 32, (1 << 32) - 1 + xsq]
    ks += [int.from_bytes(g.bytes(40), "big") % c.n for _ in range(60)]
    k1 = np.zeros(5, dtype=np.uint32)
    k2 = np.zeros(5, dtype=np.uint32)
    for kk in ks:
        k.hs_bls_glv_split(p(words(kk, 8)), p(k1), p(k2))
        rem, q = val(k1), val(k2)
        assert q * xsq + rem == kk and rem < xsq and q < (1 << 128) and k1[4] == 0 and k2[4] == 0, hex(kk)
    pts = [c.mul(int.from_bytes(g.bytes(40), "big") % c.n or 1, c.G) for _ in range(len(ks))]
    pts[3] = c.G
    kb = rows([v.to_bytes(32, "big") for v in ks])
    pb = rows([c.enc(P) for P in pts])
    n = len(ks)
    inf_in = np.zeros(n, dtype=np.uint8)
    inf_in[5] = 1
    out = np.zeros((n, 96), dtype=np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    assert k.hs_bls_g1_mul_glv(p(kb), p(pb), p(inf_in), ctypes.c_size_t(n), p(out), p(inf)) == 2**64 - 1
    out2 = np.zeros((n, 96), dtype=np.uint8)
    inf2 = np.zeros(n, dtype=np.uint8)
    assert k.hs_wei_mul(2, p(kb), p(pb), p(inf_in), ctypes.c_size_t(n), p(out2), p(inf2)) == 2**64 - 1
    assert np.array_equal(out, out2) and np.array_equal(inf, inf2)
    exp, einf = coracle.wei_mul("bls12_381_g1", kb, pb, inf_in=inf_in)
    assert np.array_equal(out, exp) and np.array_equal(inf.astype(bool), einf)
    for i in (0, 1, 3, 6, 20):
        want = c.mul(ks[i], pts[i]) if ks[i] else None
        assert (inf[i] == 1) == (want is None) and (want is None or out[i].tobytes() == c.enc(want))
    bad = kb.copy()
    bad[4] = 0xFF
    assert k.hs_bls_g1_mul_glv(p(bad), p(pb), p(inf_in), ctypes.c_size_t(n), p(out), p(inf)) == (4 << 8) | 1
    badp = pb.copy()
    badp[9, -1] ^= 1
    assert k.hs_bls_g1_mul_glv(p(kb), p(badp), p(inf_in), ctypes.c_size_t(n), p(out), p(inf)) == (9 << 8) | 2


@pytest.mark.parametrize("cid,curve", [(2, "bls12_381_g1"), (3, "p256k1")])
def test_wei_msm_bucket_method(hs, cid, curve):
    """msm.cuh (signed windows, counting sort, bucket sums, j * B, tree sum, Horner) against the big-integer group law:
    sum_i k_i P_i for random (k, P), with the cases the Jacobian additions must survive inside a bucket — the same
    point twice with the same digit (P + P), P and -P with the same scalar (cancel to the identity), zero scalars —
    and, on BLS12-381, points outside the prime-order subgroup; window widths 4, 5 and 9."""
    _, k = hs
    c = R.WCURVES[curve]
    g = rng(700 + cid)
    pts = [c.mul(int.from_bytes(g.bytes(40), "big") % c.n or 1, c.G) for _ in range(40)]
    if curve == "bls12_381_g1":
        from helpers import bls_cofactor_points

        pts += bls_cofactor_points()
    ks = [int.from_bytes(g.bytes(40), "big") % c.n for _ in pts]
    pts += [pts[0], pts[1], c.neg(pts[2]), pts[3]]          # duplicates / opposite points ...
    ks += [ks[0], ks[1], ks[2], 0]                           # ... with equal scalars, and a zero scalar
    ks[5] = c.n - 1
    ks[6] = 1
    want = None
    for kk, P in zip(ks, pts):
        want = c.add(want, c.mul(kk, P)) if kk else want
    kb = rows([v.to_bytes(c.sbytes, "big") for v in ks])
    pb = rows([c.enc(P) for P in pts])
    for cw in (4, 5, 9):
        out = np.zeros(2 * c.fbytes, dtype=np.uint8)
        inf = np.zeros(1, dtype=np.uint8)
        assert k.hs_wei_msm(cid, p(kb), p(pb), ctypes.c_size_t(len(ks)), cw, p(out), p(inf)) == 2**64 - 1
        assert not inf[0] and out.tobytes() == c.enc(want), cw
    # everything cancels: k P + k (-P) -> the identity
    kb2 = rows([ks[0].to_bytes(c.sbytes, "big")] * 2)
    pb2 = rows([c.enc(pts[0]), c.enc(c.neg(pts[0]))])
    out = np.zeros(2 * c.fbytes, dtype=np.uint8)
    inf = np.zeros(1, dtype=np.uint8)
    k.hs_wei_msm(cid, p(kb2), p(pb2), ctypes.c_size_t(2), 4, p(out), p(inf))
    assert inf[0] and not out.any()


@pytest.mark.parametrize("cid,curve", [(2, "bls12_381_g1"), (3, "p256k1")])
def test_wei_msm_skewed_buckets(hs, cid, curve):
    """The bucket sums are balanced over segments of the sorted index array (msm.cuh: msm_segment_body / msm_finish_body /
    msm_chunk_body), so the distribution of the digits must not matter: equal scalars (every point in ONE bucket per
    window), scalars with the top bit set (the short top window of a 256-bit scalar: half of the points share a bucket),
    a few distinct scalars, and random ones — for segment lengths that cut buckets into many pieces, with the warp-level
    sum of a heavy bucket's pieces replayed lane by lane (heavy threshold 0, 2, 16), and running-sum chunks of 1 .. 32."""
    _, k = hs
    c = R.WCURVES[curve]
    g = rng(800 + cid)
    base = [c.mul(int.from_bytes(g.bytes(40), "big") % c.n or 1, c.G) for _ in range(12)]
    n = 150
    pts = [base[int(g.integers(0, len(base)))] for _ in range(n)]
    pts = [P if g.integers(0, 4) else c.neg(P) for P in pts]
    pb = rows([c.enc(P) for P in pts])
    top = 1 << (c.n.bit_length() - 1)
    families = {
        "equal": [int.from_bytes(g.bytes(40), "big") % c.n] * n,
        "top_bit": [(top | int.from_bytes(g.bytes(40), "big")) % c.n for _ in range(n)],
        "few": [[1, 2, c.n - 1, 0x10001, top][int(g.integers(0, 5))] for _ in range(n)],
        "random": [int.from_bytes(g.bytes(40), "big") % c.n for _ in range(n)],
    }
    out = np.zeros(2 * c.fbytes, dtype=np.uint8)
    inf = np.zeros(1, dtype=np.uint8)
    for name, ks in families.items():
        want = None
        for kk, P in zip(ks, pts):
            want = c.add(want, c.mul(kk, P)) if kk else want
        kb = rows([v.to_bytes(c.sbytes, "big") for v in ks])
        for cw, S, CH, heavy in ((4, 1, 1, 0), (5, 3, 2, 2), (7, 2, 32, 16), (6, 8, 4, 0), (9, 5, 32, 2), (4, 512, 8, 16)):
            st = k.hs_wei_msm2(cid, p(kb), p(pb), ctypes.c_size_t(n), cw, ctypes.c_size_t(S), CH, heavy, p(out), p(inf))
            assert st == 2**64 - 1
            if want is None:
                assert inf[0] and not out.any(), (name, cw, S, CH, heavy)
            else:
                assert not inf[0] and out.tobytes() == c.enc(want), (name, cw, S, CH, heavy)
