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
