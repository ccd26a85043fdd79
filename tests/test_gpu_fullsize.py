"""BASELINE.json's full batch sizes (2^20) through the C ABI, checked with size-independent properties.

Configs 3a, 3b and 4 at n = 2^20 (config 1 and 2 at full size live in test_gpu_parity.py).  Every
element is checked — not a sample — by a property that runs a *different* kernel family:
  k * (t * G)  ==  (k t mod n) * G      variable-base windows (k_wei_mul) vs the generator comb
and the verification batches carry their expected accept bits by construction (1/16 corrupted).
A sampled comparison with the C oracle pins the bytes themselves.
"""
import numpy as np
import pytest

from helpers import rng
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 1 << 20


def _scalars(g, n, c, clear_bits):
    a = g.integers(0, 256, size=(n, c.sbytes), dtype=np.uint8)
    a[:, 0] &= 0xFF >> clear_bits
    return a


def _mulmod_rows(ka, ta, c, period):
    """(k_i * t_(i mod period)) mod n as big-endian rows."""
    t = [int.from_bytes(r.tobytes(), "big") for r in ta[:period]]
    out = np.empty((ka.shape[0], c.sbytes), dtype=np.uint8)
    buf = bytearray()
    for i, r in enumerate(ka):
        buf += (int.from_bytes(r.tobytes(), "big") * t[i % period] % c.n).to_bytes(c.sbytes, "big")
    out[:] = np.frombuffer(bytes(buf), dtype=np.uint8).reshape(-1, c.sbytes)
    return out


@pytest.mark.parametrize("curve,clear", [("p256r1", 1), ("bls12_381_g1", 2)])
def test_variable_base_full_batch_equals_fixed_base_of_the_product(ctx, coracle, curve, clear):
    """Configs 3a / 4: 2^20 (k, P = t G) pairs; k P from k_wei_mul must equal (k t) G from the comb
    kernel for every element, and a sample must equal the oracle's bytes."""
    c = R.WCURVES[curve]
    g = rng(0xECC00003 if curve == "p256r1" else 0xECC00004)
    period = 1 << 12
    t = _scalars(g, period, c, clear)
    k = _scalars(g, N, c, clear)
    base, inf = ctx.wei_mul_base(curve, t)
    assert not inf.any()
    pts = np.ascontiguousarray(np.tile(base, (N // period, 1)))
    got, ginf = ctx.wei_mul(curve, k, pts)
    want, winf = ctx.wei_mul_base(curve, _mulmod_rows(k, t, c, period))
    assert np.array_equal(ginf, winf) and np.array_equal(got, want)
    idx = g.integers(0, N, size=1024)
    exp, einf = coracle.wei_mul(curve, k[idx], pts[idx], nthreads=coracle.default_threads())
    assert np.array_equal(got[idx], exp) and np.array_equal(ginf[idx], einf)


def test_ecdsa_p256_full_batch_accepts_exactly_the_untampered(ctx, coracle):
    """Config 3b: 2^20 synthetic P-256 signatures, every 16th corrupted (bench.py's generator):
    the accept bits must be exactly the construction's, and a sample must agree with the oracle."""
    import bench

    q, z, rs = bench.make_inputs("p256_ecdsa_verify", N, ctx, 0xECC0003B)
    ok = ctx.ecdsa_verify_hashed("p256r1", q, z, rs)
    assert np.array_equal(ok, np.arange(N) % 16 != 5)
    idx = rng(9).integers(0, N, size=1024)
    assert np.array_equal(ok[idx], coracle.ecdsa_verify_hashed("p256r1", q[idx], z[idx], rs[idx], coracle.default_threads()))


def test_ed25519_verify_full_batch_accepts_exactly_the_untampered(ctx, coracle):
    """North star verify_batch at 2^20: R = r B, A = a B, S = r + k a with every 16th S corrupted."""
    import bench

    a, r, s, k = bench.make_inputs("ed25519_verify", N, ctx, 0xECC00016)
    ok = ctx.ed25519_verify_prehashed(a, r, s, k)
    assert np.array_equal(ok, np.arange(N) % 16 != 5)
    idx = rng(10).integers(0, N, size=1024)
    assert np.array_equal(ok[idx], coracle.ed25519_verify_prehashed(a[idx], r[idx], s[idx], k[idx], coracle.default_threads()))


# ---- config 5 members at 2^18 and everything above 2^20 (100 % up to 2^20, sampled above) -------
def test_p384_variable_base_2p18_equals_fixed_base_of_the_product(ctx, coracle):
    """Config 5, p384r1 Point::mul at n = 2^18: every element against the comb kernel through
    k (t G) == (k t mod n) G, and 1024 sampled elements against the C oracle's bytes."""
    c = R.WCURVES["p384r1"]
    n = 1 << 18
    g = rng(0xECC00005)
    period = 1 << 11
    t, k = _scalars(g, period, c, 1), _scalars(g, n, c, 1)
    base, inf = ctx.wei_mul_base("p384r1", t)
    assert not inf.any()
    pts = np.ascontiguousarray(np.tile(base, (n // period, 1)))
    got, ginf = ctx.wei_mul("p384r1", k, pts)
    want, winf = ctx.wei_mul_base("p384r1", _mulmod_rows(k, t, c, period))
    assert np.array_equal(ginf, winf) and np.array_equal(got, want)
    idx = g.integers(0, n, size=1024)
    exp, einf = coracle.wei_mul("p384r1", k[idx], pts[idx], nthreads=coracle.default_threads())
    assert np.array_equal(got[idx], exp) and np.array_equal(ginf[idx], einf)


def test_x448_2p18_diffie_hellman_commutes_for_every_element(ctx, coracle):
    """Config 5, X448 at n = 2^18: a (b G) == b (a G) for every one of the 2^18 pairs (four ladder batches), and
    2048 sampled (k, u) pairs with arbitrary u against the C oracle."""
    n = 1 << 18
    g = rng(0xECC00055)
    a, b = g.integers(0, 256, size=(n, 56), dtype=np.uint8), g.integers(0, 256, size=(n, 56), dtype=np.uint8)
    five = np.zeros((n, 56), dtype=np.uint8)
    five[:, 0] = 5                                          # the base point u = 5 (protocol/x448.rs)
    ag, bg = ctx.x448(a, five), ctx.x448(b, five)
    assert np.array_equal(ctx.x448(a, bg), ctx.x448(b, ag))
    u = g.integers(0, 256, size=(n, 56), dtype=np.uint8)
    out = ctx.x448(a, u)
    idx = g.integers(0, n, size=2048)
    assert np.array_equal(out[idx], coracle.x448(a[idx], u[idx], coracle.default_threads()))
    assert np.array_equal(ag[idx], coracle.x448(a[idx], five[idx], coracle.default_threads()))


def _periodic_check(full, period):
    """Every row of `full` equals the row of the first period it repeats: all elements are compared."""
    n = full.shape[0]
    assert n % period == 0
    return bool((full.reshape(n // period, period, -1) == full[:period][None]).all())


@pytest.mark.parametrize("logn", [22, 24])
def test_ed25519_mul_base_above_2p20(ctx, coracle, logn):
    """Config 5 sizes above 2^20 through the host entry point (dozens of pipeline chunks): the scalars repeat with
    period 2^16, so EVERY output row is compared with its first occurrence, and the first 2^16 rows — a sample
    of 2^16 as SURVEY §8d asks — with the C oracle."""
    n, period = 1 << logn, 1 << 16
    g = rng(0xECC05000 + logn)
    k = g.integers(0, 256, size=(period, 32), dtype=np.uint8)
    k[:, 31] &= 0x0F
    out = ctx.ed25519_mul_base(np.ascontiguousarray(np.tile(k, (n // period, 1))))
    assert _periodic_check(out, period)
    assert np.array_equal(out[:period], coracle.ed25519_mul_base(k, coracle.default_threads()))


@pytest.mark.parametrize("curve,logn,clear", [("p256r1", 22, 1), ("p256r1", 24, 1), ("p384r1", 22, 1)])
def test_weierstrass_mul_above_2p20(ctx, coracle, curve, logn, clear):
    """Config 5: p256r1 Point::mul at 2^22 and 2^24, p384r1 at 2^22 — every row against its first occurrence
    (inputs repeat with period 2^14), the first period against the comb of the product, 1024 rows against the oracle."""
    c = R.WCURVES[curve]
    n, period = 1 << logn, 1 << 14
    g = rng(0xECC05100 + logn + c.sbytes)
    t, k = _scalars(g, period, c, clear), _scalars(g, period, c, clear)
    base, inf = ctx.wei_mul_base(curve, t)
    assert not inf.any()
    reps = n // period
    got, ginf = ctx.wei_mul(curve, np.ascontiguousarray(np.tile(k, (reps, 1))), np.ascontiguousarray(np.tile(base, (reps, 1))))
    assert _periodic_check(got, period) and _periodic_check(ginf.reshape(n, 1), period)
    want, winf = ctx.wei_mul_base(curve, _mulmod_rows(k, t, c, period))
    assert np.array_equal(got[:period], want) and np.array_equal(ginf[:period], winf)
    idx = g.integers(0, period, size=1024)
    exp, einf = coracle.wei_mul(curve, k[idx], base[idx], nthreads=coracle.default_threads())
    assert np.array_equal(got[idx], exp) and np.array_equal(ginf[idx], einf)


def test_x448_2p22(ctx, coracle):
    """Config 5: X448 at 2^22, inputs repeating with period 2^14: every row against its first occurrence, 1024 rows
    against the C oracle."""
    n, period = 1 << 22, 1 << 14
    g = rng(0xECC05448)
    k, u = g.integers(0, 256, size=(period, 56), dtype=np.uint8), g.integers(0, 256, size=(period, 56), dtype=np.uint8)
    out = ctx.x448(np.ascontiguousarray(np.tile(k, (n // period, 1))), np.ascontiguousarray(np.tile(u, (n // period, 1))))
    assert _periodic_check(out, period)
    idx = g.integers(0, period, size=1024)
    assert np.array_equal(out[idx], coracle.x448(k[idx], u[idx], coracle.default_threads()))
