"""Parity of the CUDA path with the reference (through the oracle), all through the C ABI.

Every test calls libeccbatch.so via eccoxide_b200.Context (ctypes, one C-ABI call per method) on
cuda:0 and compares byte for byte with
  - the reference's golden vectors (tests/golden/reference_vectors.json),
  - oracle/ecc_oracle.c / oracle/pyref.py on the same seeded inputs (small sizes),
  - size-independent properties at BASELINE.json's full batch sizes (sampled oracle check,
    homomorphism, DH commutativity, fixed-base == variable-base on the generator).
Bar: bit-exact (integer / byte work).
"""
import hashlib
import os

import numpy as np
import pytest

from helpers import (SEEDS, ecdsa_batch, ed25519_sig_batch, ed_edge_scalars, ed_points, rand_bytes, rng, rows, scalars_mod,
                     wei_edge_scalars, wei_points)
from oracle import pyref as R

pytestmark = pytest.mark.gpu
H = bytes.fromhex
CURVES = ("p256r1", "p384r1", "bls12_381_g1", "p256k1")
NT = None


def threads(coracle):
    return coracle.default_threads()


# ---- library / device sanity -----------------------------------------------------------------
def test_native_library_is_loaded_and_counts_launches(ctx):
    import eccoxide_b200._lib as L

    assert L._lib is not None and L.LIB_PATH.endswith("libeccbatch.so")
    before = ctx.launch_count()
    ctx.ed25519_mul_base(rows([(5).to_bytes(32, "little")]))
    assert ctx.launch_count() >= before + 1  # small batch: ONE fused kernel (comb + block-level inversion + encoding)
    ctx.set_option("ed25519_fused", 0)
    before = ctx.launch_count()
    ctx.ed25519_mul_base(rows([(5).to_bytes(32, "little")]))
    ctx.set_option("ed25519_fused", 1)
    assert ctx.launch_count() >= before + 2  # large-batch form: scalar-mult kernel + batch inversion


def test_imad_probe_reports_a_plausible_peak(ctx):
    lo, _ = ctx.imad_probe(0, 512)
    wide, _ = ctx.imad_probe(2, 512)
    assert 5e12 < lo < 40e12 and 2e12 < wide < 40e12


# ---- Ed25519 fixed base (config 1) --------------------------------------------------------------
def test_ed25519_mul_base_golden_and_edges(ctx, golden, coracle):
    ks = []
    for v in golden["ed25519_rfc8032"]:
        h = bytearray(hashlib.sha512(H(v["seed"])).digest()[:32])
        h[0] &= 248; h[31] &= 127; h[31] |= 64
        ks.append((int.from_bytes(h, "little") % R.L25519).to_bytes(32, "little"))
    enc = ctx.ed25519_mul_base(rows(ks), compressed=True)
    assert [enc[i].tobytes().hex() for i in range(3)] == [v["public"] for v in golden["ed25519_rfc8032"]]
    vals = ed_edge_scalars(golden)
    kb = rows([v.to_bytes(32, "little") for v in vals])
    got = ctx.ed25519_mul_base(kb)
    assert np.array_equal(got, coracle.ed25519_mul_base(kb))
    assert got[0].tobytes() == (0).to_bytes(32, "little") + (1).to_bytes(32, "little")
    for i in (1, 5, len(vals) - 1):
        assert got[i].tobytes() == R.ed25519_mul_base_xy(kb[i].tobytes())


@pytest.mark.parametrize("w", [4, 5, 8, 11])
def test_ed25519_mul_base_every_comb_width(w, coracle, golden):
    """The comb width is a tunable of the GPU schedule; results must not depend on it."""
    from eccoxide_b200 import Context

    g = rng(SEEDS["ed25519"] + w)
    kb = np.concatenate([scalars_mod(g, 600, R.L25519, 32, "little"), rows([v.to_bytes(32, "little") for v in ed_edge_scalars(golden)])])
    with Context(ed25519_comb_w=w) as c:
        assert np.array_equal(c.ed25519_mul_base(kb), coracle.ed25519_mul_base(kb, threads(coracle)))
        table, tw, nwin = c.debug_ed25519_table()
        assert tw == w and table.shape[0] == nwin << (w - 1)
        # table entry (i, j) = niels(j * 2^(w i) * B): check a few against the big-int oracle
        for (i, j) in ((0, 1), (1, 2), (nwin - 1, 1), (2, 1 << (w - 1))):
            x, y = R.ed_mul(j << (w * i), R.ED_B)
            p = R.P25519
            exp = ((y + x) % p).to_bytes(32, "little") + ((y - x) % p).to_bytes(32, "little") + (2 * R.ED_D * x * y % p).to_bytes(32, "little")
            assert table[i * (1 << (w - 1)) + j - 1].tobytes() == exp


@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8])
def test_ed25519_mul_base_fused_small_batch_every_lane_count(lanes, coracle, golden):
    """The small-batch kernel (fused.cuh: windows split over lanes, block-level inversion in the same launch)
    against the C oracle at 2^10 ... 2^17 and ragged sizes, for every lane count, plus the edge scalars;
    the two-kernel large-batch form (ed25519_fused = 0) must give the same bytes."""
    from eccoxide_b200 import Context

    g = rng(SEEDS["ed25519"] + 100 + lanes)
    edge = rows([v.to_bytes(32, "little") for v in ed_edge_scalars(golden)])
    big = scalars_mod(g, (1 << 17) + 77, R.L25519, 32, "little")
    exp_big = coracle.ed25519_mul_base(big, threads(coracle))
    with Context() as c:
        c.set_option("ed25519_lanes", lanes)
        c.ed25519_mul_base(edge[:2])                  # builds the comb table (its launches are not the call's)
        cap = 148 * 480 // max(lanes, 1)
        for n in (1, 31, 33, 1 << 10, 1 << 12, (1 << 13) + 5, 1 << 14, 1 << 15, 1 << 16, 75776, (1 << 17) + 77):
            kb = big[:n].copy()
            m = min(n, edge.shape[0])
            kb[:m] = edge[:m]
            exp = exp_big[:n].copy()
            exp[:m] = coracle.ed25519_mul_base(edge[:m])
            c.set_option("ed25519_fused", 1)
            l0 = c.launch_count()
            got = c.ed25519_mul_base(kb)
            used_fused = c.launch_count() - l0 == 1
            assert np.array_equal(got, exp), (lanes, n)
            if lanes and n <= cap and n < 16384:       # below one pipeline chunk the host path is a single launch
                assert used_fused, (lanes, n)
            enc = c.ed25519_mul_base(kb, compressed=True)
            assert np.array_equal(enc[:, :31], exp[:, 32:63]) and np.array_equal(enc[:, 31] & 0x7F, exp[:, 63]) and np.array_equal(enc[:, 31] >> 7, exp[:, 0] & 1)
            if lanes == 0:
                c.set_option("ed25519_fused", 0)
                assert np.array_equal(c.ed25519_mul_base(kb), exp), ("split", n)
                c.set_option("ed25519_fused", 2)   # the fused body with the large-batch launch shape
                assert np.array_equal(c.ed25519_mul_base(kb), exp), ("fused2", n)
        # non-canonical scalar is still reported with its index
        bad = big[:5000].copy()
        bad[4321] = np.frombuffer(R.L25519.to_bytes(32, "little"), dtype=np.uint8)
        c.set_option("ed25519_fused", 1)
        with pytest.raises(Exception) as ei:
            c.ed25519_mul_base(bad)
        assert getattr(ei.value, "bad_index", None) == 4321


def test_x25519_base_fused_small_batch(ctx, coracle):
    g = rng(SEEDS["x25519"] + 9)
    for n in (1, 100, 4096, 1 << 15, 1 << 16):
        ks = rand_bytes(g, n, 32)
        nine = np.tile(np.frombuffer((9).to_bytes(32, "little"), dtype=np.uint8), (n, 1))
        assert np.array_equal(ctx.x25519_base(ks), coracle.x25519(ks, nine, threads(coracle))), n


@pytest.mark.parametrize("stride", [24, 32])
def test_ed25519_comb_entry_stride(stride, coracle):
    """Comb entries packed (96 B) or one per 128-byte line: same results, same table contents."""
    from eccoxide_b200 import Context

    g = rng(SEEDS["ed25519"] + 7)
    kb = scalars_mod(g, 3000, R.L25519, 32, "little")
    with Context(ed25519_comb_w=9) as c:
        c.set_option("ed25519_entry_stride", stride)
        exp = coracle.ed25519_mul_base(kb, threads(coracle))
        assert np.array_equal(c.ed25519_mul_base(kb), exp)
        c.set_option("ed25519_fused", 0)
        assert np.array_equal(c.ed25519_mul_base(kb), exp)
        table, tw, nwin = c.debug_ed25519_table()
        x, y = R.ed_mul(3 << 9, R.ED_B)
        p = R.P25519
        assert table[(1 << 8) + 2].tobytes() == ((y + x) % p).to_bytes(32, "little") + ((y - x) % p).to_bytes(32, "little") + (2 * R.ED_D * x * y % p).to_bytes(32, "little")


def test_ed25519_mul_base_config1_full_batch(ctx, coracle):
    """Config 1 at full size (2^16 random scalars): 100 % against the C oracle, plus libsodium sample."""
    g = rng(SEEDS["ed25519"])
    kb = scalars_mod(g, 1 << 16, R.L25519, 32, "little")
    got = ctx.ed25519_mul_base(kb)
    assert np.array_equal(got, coracle.ed25519_mul_base(kb, threads(coracle)))
    enc = ctx.ed25519_mul_base(kb[:2048], compressed=True)
    nb = pytest.importorskip("nacl.bindings")
    for i in range(0, 2048, 64):
        if any(kb[i]):
            assert enc[i].tobytes() == nb.crypto_scalarmult_ed25519_base_noclamp(kb[i].tobytes())
    # compressed == encode(affine)
    for i in range(0, 2048, 97):
        x = int.from_bytes(got[i, :32].tobytes(), "little"); y = int.from_bytes(got[i, 32:].tobytes(), "little")
        assert enc[i].tobytes() == R.ed_encode((x, y))


def test_ed25519_mul_base_large_batch_properties(ctx, coracle):
    """2^20 scalars: sampled oracle check + homomorphism (a+b)B = aB + bB on the outputs."""
    g = rng(SEEDS["ed25519"] + 1)
    n = 1 << 20
    kb = rand_bytes(g, n, 32)
    kb[:, 31] &= 0x0F
    got = ctx.ed25519_mul_base(kb)
    idx = g.integers(0, n, size=4096)
    assert np.array_equal(got[idx], coracle.ed25519_mul_base(kb[idx], threads(coracle)))
    for i in range(0, 64, 2):
        a, b = int.from_bytes(kb[i].tobytes(), "little"), int.from_bytes(kb[i + 1].tobytes(), "little")
        pa = (int.from_bytes(got[i, :32].tobytes(), "little"), int.from_bytes(got[i, 32:].tobytes(), "little"))
        pb = (int.from_bytes(got[i + 1, :32].tobytes(), "little"), int.from_bytes(got[i + 1, 32:].tobytes(), "little"))
        assert R.ed_add(pa, pb) == R.ed_mul((a + b) % R.L25519, R.ED_B)


def test_ed25519_noncanonical_scalar_is_reported(ctx):
    from eccoxide_b200 import EccBatchError

    kb = rows([(7).to_bytes(32, "little")] * 300)
    kb = kb.copy()
    kb[123] = np.frombuffer(R.L25519.to_bytes(32, "little"), dtype=np.uint8)
    kb[200] = 0xFF
    with pytest.raises(EccBatchError) as e:
        ctx.ed25519_mul_base(kb)
    assert e.value.code == -3 and e.value.bad_index == 123


def test_empty_batches(ctx):
    assert ctx.ed25519_mul_base(np.zeros((0, 32), dtype=np.uint8)).shape == (0, 64)
    assert ctx.x25519(np.zeros((0, 32), dtype=np.uint8), np.zeros((0, 32), dtype=np.uint8)).shape == (0, 32)
    out, inf = ctx.wei_mul("p256r1", np.zeros((0, 32), dtype=np.uint8), np.zeros((0, 64), dtype=np.uint8))
    assert out.shape == (0, 64) and inf.shape == (0,)


@pytest.mark.parametrize("n", [1, 2, 31, 33, 127, 129, 1000])
def test_ragged_batch_sizes(ctx, coracle, n):
    g = rng(n)
    kb = scalars_mod(g, n, R.L25519, 32, "little")
    assert np.array_equal(ctx.ed25519_mul_base(kb), coracle.ed25519_mul_base(kb))
    k, u = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
    assert np.array_equal(ctx.x25519(k, u), coracle.x25519(k, u))


# ---- Ed25519 variable base ------------------------------------------------------------------------
def test_ed25519_mul_variable_base(ctx, golden, coracle):
    g = rng(SEEDS["ed25519"] + 2)
    vals = ed_edge_scalars(golden)
    kb = np.concatenate([rows([v.to_bytes(32, "little") for v in vals]), scalars_mod(g, 1000, R.L25519, 32, "little")])
    pts = ed_points(g, kb.shape[0]).copy()
    # a few special points: identity, the order-2 point (0,-1), an order-4 point (sqrt(-1), 0), an order-8 point
    special = [(0, 1), (0, R.P25519 - 1), (R.SQRT_M1, 0)]
    for i, (x, y) in enumerate(special):
        pts[i] = np.frombuffer(x.to_bytes(32, "little") + y.to_bytes(32, "little"), dtype=np.uint8)
    got = ctx.ed25519_mul(kb, pts)
    assert np.array_equal(got, coracle.ed25519_mul(kb, pts, threads(coracle)))
    for i in (0, 1, 2, 30, 500):
        assert got[i].tobytes() == R.ed25519_mul_xy(kb[i].tobytes(), pts[i].tobytes())
    # mul_base_matches_scale (curve25519.rs:1374)
    gpt = np.tile(np.frombuffer(R.ED_BX.to_bytes(32, "little") + R.ED_BY.to_bytes(32, "little"), dtype=np.uint8), (kb.shape[0], 1))
    assert np.array_equal(ctx.ed25519_mul(kb, gpt), ctx.ed25519_mul_base(kb))


def test_ed25519_mul_rejects_bad_points(ctx):
    from eccoxide_b200 import EccBatchError

    good = R.ED_BX.to_bytes(32, "little") + R.ED_BY.to_bytes(32, "little")
    pts = rows([good] * 50).copy()
    pts[17, 0] ^= 1
    with pytest.raises(EccBatchError) as e:
        ctx.ed25519_mul(rows([(3).to_bytes(32, "little")] * 50), pts)
    assert e.value.code == -4 and e.value.bad_index == 17
    # non-canonical coordinate (x + p) is rejected like FieldElement::from_bytes does
    pts = rows([good] * 4).copy()
    pts[2, :32] = np.frombuffer((R.ED_BX + R.P25519).to_bytes(32, "little"), dtype=np.uint8)
    with pytest.raises(EccBatchError) as e:
        ctx.ed25519_mul(rows([(3).to_bytes(32, "little")] * 4), pts)
    assert e.value.bad_index == 2


# ---- X25519 (config 2) ------------------------------------------------------------------------------
def test_x25519_golden_and_edges(ctx, golden, coracle):
    v = golden["x25519"]
    ks = [H(e["k"]) for e in v["rfc7748_5_2"]] + [H(v["iterated_once"]["k"])]
    us = [H(e["u"]) for e in v["rfc7748_5_2"]] + [(9).to_bytes(32, "little")]
    rs = [H(e["r"]) for e in v["rfc7748_5_2"]] + [H(v["iterated_once"]["r"])]
    nine = (9).to_bytes(32, "little")
    a, b = H(v["dh_6_1"]["a"]), H(v["dh_6_1"]["b"])
    pubs = ctx.x25519(rows([a, b]), rows([nine, nine]))
    ks += [a, b]; us += [pubs[1].tobytes(), pubs[0].tobytes()]; rs += [H(v["dh_6_1"]["shared"])] * 2
    assert np.array_equal(ctx.x25519(rows(ks), rows(us)), rows(rs))
    p = R.P25519
    g = rng(1)
    eus = [0, 1, p - 1, p, p + 1, 2**255 - 1, 2**256 - 1, 9 | 1 << 255,
           325606250916557431795983626356110631294008115727848805560023387167927233504,
           39382357235489614581723060781553021112529911719440698176882885853963445705823]
    eus = rows([x.to_bytes(32, "little") for x in eus])
    eks = rand_bytes(g, eus.shape[0], 32)
    got = ctx.x25519(eks, eus)
    assert np.array_equal(got, coracle.x25519(eks, eus))
    assert not got[0].any() and not got[1].any() and not got[3].any()  # low-order inputs -> zero output


def test_x25519_config2_full_batch(ctx, coracle):
    """Config 2 at full size (2^20 pairs): sampled 2^14 oracle check, libsodium sample, DH commutativity."""
    g = rng(SEEDS["x25519"])
    n = 1 << 20
    k, u = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
    got = ctx.x25519(k, u)
    idx = np.concatenate([np.arange(4096), g.integers(0, n, size=12288)])
    assert np.array_equal(got[idx], coracle.x25519(k[idx], u[idx], threads(coracle)))
    nb = pytest.importorskip("nacl.bindings")
    for i in range(0, 2048, 32):
        try:
            assert got[i].tobytes() == nb.crypto_scalarmult(k[i].tobytes(), u[i].tobytes())
        except Exception:
            assert not got[i].any()
    # DH: x25519(a, x25519(b, 9)) == x25519(b, x25519(a, 9))
    m = 1 << 12
    nine = np.tile(np.frombuffer((9).to_bytes(32, "little"), dtype=np.uint8), (m, 1))
    pa, pb = ctx.x25519(k[:m], nine), ctx.x25519(u[:m], nine)
    assert np.array_equal(ctx.x25519(k[:m], pb), ctx.x25519(u[:m], pa))


def test_x25519_base_public_keys(ctx, golden, coracle):
    """x25519_base (x25519.rs:49): fixed-base comb + Montgomery map == the ladder on u = 9, byte for byte."""
    v = golden["x25519"]
    g = rng(SEEDS["x25519"] + 9)
    n = 1 << 16
    ks = rand_bytes(g, n, 32)
    ks[0] = 0; ks[1] = 0xFF
    ks[2] = np.frombuffer(H(v["iterated_once"]["k"]), dtype=np.uint8)
    ks[3] = np.frombuffer(H(v["dh_6_1"]["a"]), dtype=np.uint8)
    ks[4] = np.frombuffer(H(v["dh_6_1"]["b"]), dtype=np.uint8)
    got = ctx.x25519_base(ks)
    assert got[2].tobytes() == H(v["iterated_once"]["r"])
    nine = np.tile(np.frombuffer((9).to_bytes(32, "little"), dtype=np.uint8), (n, 1))
    assert np.array_equal(got, ctx.x25519(ks, nine))                       # the ladder kernel, 100 %
    assert np.array_equal(got[:4096], coracle.x25519(ks[:4096], nine[:4096], threads(coracle)))
    # RFC 7748 6.1: shared secret from the two public keys
    shared = ctx.x25519(ks[3:5], got[[4, 3]])
    assert shared[0].tobytes() == shared[1].tobytes() == H(v["dh_6_1"]["shared"])


def test_x448(ctx, golden, coracle):
    v = golden["x448"]
    ks = [H(e["k"]) for e in v["rfc7748_5_2"]]
    us = [H(e["u"]) for e in v["rfc7748_5_2"]]
    rs = [H(e["r"]) for e in v["rfc7748_5_2"]]
    d = v["dh_6_2"]
    five = (5).to_bytes(56, "little")
    ks += [H(d["a"]), H(d["b"]), H(d["a"])]; us += [five, five, H(d["b_pub"])]; rs += [H(d["a_pub"]), H(d["b_pub"]), H(d["shared"])]
    assert np.array_equal(ctx.x448(rows(ks), rows(us)), rows(rs))
    g = rng(SEEDS["sweep"])
    n = 3000
    k, u = rand_bytes(g, n, 56), rand_bytes(g, n, 56)
    u[0] = 0; u[1] = 0xFF; u[2] = np.frombuffer(R.P448.to_bytes(56, "little"), dtype=np.uint8)
    u[3] = np.frombuffer((1).to_bytes(56, "little"), dtype=np.uint8)
    assert np.array_equal(ctx.x448(k, u), coracle.x448(k, u, threads(coracle)))


# ---- Weierstrass variable base (configs 3a, 4, 5) -----------------------------------------------------
@pytest.mark.parametrize("curve,key", [("p256r1", "nist_p256"), ("p384r1", "nist_p384")])
def test_wei_nist_kats(ctx, golden, curve, key):
    c = R.WCURVES[curve]
    kats = golden[key]
    ks = rows([int(v["k"], 16).to_bytes(c.sbytes, "big") for v in kats])
    exp = rows([int(v["x"], 16).to_bytes(c.fbytes, "big") + int(v["y"], 16).to_bytes(c.fbytes, "big") for v in kats])
    gpt = np.tile(np.frombuffer(c.enc(c.G), dtype=np.uint8), (len(kats), 1))
    out, inf = ctx.wei_mul(curve, ks, gpt)
    assert not inf.any() and np.array_equal(out, exp)
    out, inf = ctx.wei_mul_base(curve, ks)
    assert not inf.any() and np.array_equal(out, exp)


def test_bls_g1_kats(ctx, golden):
    c = R.BLSG1
    for ent in golden["bls12_381_g1"]["uncompressed"]:
        out, inf = ctx.wei_mul_base("bls12_381_g1", rows([ent["k"].to_bytes(32, "big")]))
        assert out[0].tobytes().hex() == ent["bytes"] and not inf[0]
    for ent in golden["bls12_381_g1"]["compressed"]:
        out, _ = ctx.wei_mul("bls12_381_g1", rows([ent["k"].to_bytes(32, "big")]), rows([c.enc(c.G)]))
        x, y = c.dec(out[0].tobytes())
        b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80 | (0x20 if y > (c.p - 1) // 2 else 0)
        assert bytes(b).hex() == ent["bytes"]


@pytest.mark.parametrize("curve", CURVES)
def test_wei_mul_random_and_edges(ctx, golden, coracle, curve):
    c = R.WCURVES[curve]
    g = rng(SEEDS["p256"] + len(curve))
    vals = wei_edge_scalars(golden, c.n)
    kb = np.concatenate([rows([v.to_bytes(c.sbytes, "big") for v in vals]), scalars_mod(g, 1500, c.n, c.sbytes, "big")])
    pts = wei_points(curve, g, kb.shape[0])
    got, inf = ctx.wei_mul(curve, kb, pts)
    exp, einf = coracle.wei_mul(curve, kb, pts, nthreads=threads(coracle))
    assert np.array_equal(got, exp) and np.array_equal(inf, einf)
    assert inf[0] and not got[0].any()  # k = 0 -> identity flag, zero bytes
    for i in (1, 17, 40):
        assert (got[i].tobytes(), int(inf[i])) == R.wei_mul(c, kb[i].tobytes(), pts[i].tobytes())
    # fixed base == variable base on the generator (mul_base_matches_generic, completeness.rs:97)
    gb, gi = ctx.wei_mul_base(curve, kb)
    gv, gvi = ctx.wei_mul(curve, kb, np.tile(np.frombuffer(c.enc(c.G), dtype=np.uint8), (kb.shape[0], 1)))
    assert np.array_equal(gb, gv) and np.array_equal(gi, gvi)
    eb, ebi = coracle.wei_mul_base(curve, kb[:200], threads(coracle))
    assert np.array_equal(gb[:200], eb) and np.array_equal(gi[:200], ebi)
    # identity inputs
    out, oinf = ctx.wei_mul(curve, kb[:64], pts[:64], inf_in=np.ones(64, dtype=np.uint8))
    assert oinf.all() and not out.any()
    # k = n - 1 gives -P (for points of the prime-order subgroup: rows 5.. — rows 1-4 of the BLS batch have order 3 or 3 r)
    km1 = rows([(c.n - 1).to_bytes(c.sbytes, "big")] * 8)
    out, _ = ctx.wei_mul(curve, km1, pts[5:13])
    for i in range(8):
        x, y = c.dec(pts[5 + i].tobytes())
        assert out[i].tobytes() == c.enc((x, (-y) % c.p))


@pytest.mark.parametrize("curve", CURVES)
def test_wei_mul_rejects_bad_inputs(ctx, curve):
    from eccoxide_b200 import EccBatchError

    c = R.WCURVES[curve]
    good = c.enc(c.G)
    one = (1).to_bytes(c.sbytes, "big")
    pts = rows([good] * 40).copy()
    pts[9, -1] ^= 1
    with pytest.raises(EccBatchError) as e:
        ctx.wei_mul(curve, rows([one] * 40), pts)
    assert e.value.code == -4 and e.value.bad_index == 9
    ks = rows([one] * 40).copy()
    ks[33] = np.frombuffer(c.n.to_bytes(c.sbytes, "big"), dtype=np.uint8)
    with pytest.raises(EccBatchError) as e:
        ctx.wei_mul(curve, ks, rows([good] * 40))
    assert e.value.code == -3 and e.value.bad_index == 33
    # x >= p is non-canonical
    pts = rows([good] * 3).copy()
    pts[1, : c.fbytes] = np.frombuffer((c.p + 1).to_bytes(c.fbytes, "big"), dtype=np.uint8) if c.p + 1 < 1 << (8 * c.fbytes) else pts[1, : c.fbytes]
    if c.p + 1 < 1 << (8 * c.fbytes):
        with pytest.raises(EccBatchError):
            ctx.wei_mul(curve, rows([one] * 3), pts)


def test_bls_off_subgroup_points(ctx, coracle):
    c = R.BLSG1
    pts, x = [], 1
    while len(pts) < 6:
        x += 1
        rhs = (x**3 + 4) % c.p
        y = pow(rhs, (c.p + 1) // 4, c.p)
        if y * y % c.p == rhs:
            pts.append((x, y))
    g = rng(5)
    kb = scalars_mod(g, len(pts), c.n, 32, "big")
    pb = rows([c.enc(p) for p in pts])
    got, inf = ctx.wei_mul("bls12_381_g1", kb, pb)
    exp, einf = coracle.wei_mul("bls12_381_g1", kb, pb)
    assert np.array_equal(got, exp) and np.array_equal(inf, einf)


def test_p256_config3a_large_batch(ctx, coracle):
    """Config 3a at 2^18 (CI-sized; bench.py runs 2^20): sampled oracle + OpenSSL ECDH + distributivity."""
    c = R.P256
    g = rng(SEEDS["p256"])
    n = 1 << 18
    kb = rand_bytes(g, n, 32)
    kb[:, 0] &= 0x7F
    uniq, _ = ctx.wei_mul_base("p256r1", kb[:4096])
    pts = np.tile(uniq, (n // 4096, 1))
    got, inf = ctx.wei_mul("p256r1", kb, pts)
    idx = g.integers(0, n, size=2048)
    exp, einf = coracle.wei_mul("p256r1", kb[idx], pts[idx], nthreads=threads(coracle))
    assert np.array_equal(got[idx], exp) and np.array_equal(inf[idx], einf)
    # k*(t*G) == (k*t mod n)*G
    for i in range(8):
        k, t = int.from_bytes(kb[i].tobytes(), "big"), int.from_bytes(kb[i % 4096].tobytes(), "big")
        assert got[i].tobytes() == c.enc(c.mul(k * t % c.n, c.G))
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    for i in range(4):
        priv = ec.derive_private_key(int.from_bytes(kb[i].tobytes(), "big"), ec.SECP256R1())
        x, y = c.dec(pts[i].tobytes())
        peer = ec.EllipticCurvePublicNumbers(x, y, ec.SECP256R1()).public_key()
        assert priv.exchange(ec.ECDH(), peer) == got[i, :32].tobytes()


# ---- ECDSA verify (config 3b) ------------------------------------------------------------------------
@pytest.mark.parametrize("curve", ["p256r1", "p384r1"])
def test_ecdsa_rfc6979_and_synthetic(ctx, golden, coracle, curve):
    c, v = R.WCURVES[curve], golden["ecdsa_rfc6979"][curve]
    Q = c.enc((int(v["qx"], 16), int(v["qy"], 16)))
    Qs, Zs, RSs, exp = [], [], [], []
    for kat in v["kats"]:
        z = R.ecdsa_digest_to_scalar(c, hashlib.new(kat["alg"], kat["message"].encode()).digest())
        rs = int(kat["r"], 16).to_bytes(c.sbytes, "big") + int(kat["s"], 16).to_bytes(c.sbytes, "big")
        for tamper in range(3):
            z2, rs2 = bytearray(z), bytearray(rs)
            if tamper == 1:
                z2[-1] ^= 1
            if tamper == 2:
                rs2[-1] ^= 1
            Qs.append(Q); Zs.append(bytes(z2)); RSs.append(bytes(rs2)); exp.append(tamper == 0)
    assert ctx.ecdsa_verify_hashed(curve, rows(Qs), rows(Zs), rows(RSs)).tolist() == exp
    g = rng(SEEDS["p256"] + 7)
    q, z, rs = ecdsa_batch(curve, g, 600)
    got = ctx.ecdsa_verify_hashed(curve, q, z, rs)
    want = coracle.ecdsa_verify_hashed(curve, q, z, rs, threads(coracle))
    assert np.array_equal(got, want)
    assert got.sum() > 300 and (~got).sum() > 100
    for i in (0, 1, 5, 7, 11, 13):
        assert bool(got[i]) == R.ecdsa_verify_hashed(c, q[i].tobytes(), z[i].tobytes(), rs[i].tobytes())


def test_ecdsa_p256_openssl_signatures(ctx):
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric.utils import decode_dss_signature

    c = R.P256
    Qs, Zs, RSs, exp = [], [], [], []
    for i in range(48):
        key = ec.generate_private_key(ec.SECP256R1())
        msg = b"block %d" % i
        r, s = decode_dss_signature(key.sign(msg, ec.ECDSA(hashes.SHA256())))
        pn = key.public_key().public_numbers()
        z = bytearray(hashlib.sha256(msg).digest())
        if i % 6 == 5:
            z[0] ^= 0x80
        Qs.append(c.enc((pn.x, pn.y))); Zs.append(bytes(z)); RSs.append(r.to_bytes(32, "big") + s.to_bytes(32, "big")); exp.append(i % 6 != 5)
    assert ctx.ecdsa_verify_hashed("p256r1", rows(Qs), rows(Zs), rows(RSs)).tolist() == exp


def test_ecdsa_off_curve_key_fails_the_call(ctx):
    from eccoxide_b200 import EccBatchError

    c = R.P256
    q = rows([c.enc(c.G)] * 10).copy()
    q[4, 5] ^= 2
    one = (1).to_bytes(32, "big")
    with pytest.raises(EccBatchError) as e:
        ctx.ecdsa_verify_hashed("p256r1", q, rows([one] * 10), rows([one + one] * 10))
    assert e.value.code == -4 and e.value.bad_index == 4


# ---- Ed25519 verify -------------------------------------------------------------------------------------
def test_ed25519_verify(ctx, golden, coracle):
    A, Rr, S, K, exp = [], [], [], [], []
    for v in golden["ed25519_rfc8032"]:
        pub, msg, sig = H(v["public"]), H(v["message"]), H(v["signature"])
        for tamper in (False, True):
            s2 = bytearray(sig)
            if tamper:
                s2[33] ^= 1
            A.append(pub); Rr.append(bytes(s2[:32])); S.append(bytes(s2[32:])); K.append(R.ed25519_hash_k(bytes(s2[:32]), pub, msg)); exp.append(not tamper)
    assert ctx.ed25519_verify_prehashed(rows(A), rows(Rr), rows(S), rows(K)).tolist() == exp
    g = rng(99)
    a, r, s, k = ed25519_sig_batch(g, 400)
    # decode_point rejections: non-canonical y, x = 0 with sign bit, not on curve
    a = a.copy(); r = r.copy()
    a[20] = np.frombuffer(R.P25519.to_bytes(32, "little"), dtype=np.uint8)
    a[21] = np.frombuffer((1 | 1 << 255).to_bytes(32, "little"), dtype=np.uint8)
    r[22] = np.frombuffer((2).to_bytes(32, "little"), dtype=np.uint8)
    got = ctx.ed25519_verify_prehashed(a, r, s, k)
    want = coracle.ed25519_verify_prehashed(a, r, s, k, threads(coracle))
    assert np.array_equal(got, want)
    assert got.sum() > 100 and (~got).sum() > 100 and not got[20] and not got[21] and not got[9]
    for i in (0, 1, 2, 3, 9, 20):
        assert bool(got[i]) == R.ed25519_verify_prehashed(a[i].tobytes(), r[i].tobytes(), s[i].tobytes(), k[i].tobytes())
    # R = identity in its canonical encoding / with the sign bit / non-canonical: the kernel compares
    # encode_point(lhs) with R's bytes instead of decoding R — same three answers as decode + compare
    from helpers import ed25519_identity_r_cases

    a, r, s, k = ed25519_identity_r_cases(g)
    assert ctx.ed25519_verify_prehashed(a, r, s, k).tolist() == [True, False, False]
    assert coracle.ed25519_verify_prehashed(a, r, s, k).tolist() == [True, False, False]


def test_ed25519_verify_raw_messages(ctx, golden, coracle):
    """PublicKey::verify(msg, sig) with the challenge hash on the device (SHA-512 + reduction mod l)."""
    pubs, msgs, sigs, exp = [], [], [], []
    for v in golden["ed25519_rfc8032"]:
        pub, msg, sig = H(v["public"]), H(v["message"]), H(v["signature"])
        pubs += [pub, pub, pub]; sigs += [sig, sig, bytes([sig[0] ^ 1]) + sig[1:]]; msgs += [msg, msg + b"!", msg]; exp += [True, False, False]
    assert ctx.ed25519_verify(rows(pubs), msgs, rows(sigs)).tolist() == exp
    g = rng(8032)
    n = 3000
    lens = [0, 1, 47, 48, 63, 64, 65, 111, 112, 175, 176, 177, 239, 240, 1000] + [int(x) for x in g.integers(0, 300, size=n - 15)]
    pubs, msgs, sigs = [], [], []
    seeds = [g.bytes(32) for _ in range(16)]
    keys = [(s, R.ed25519_public_from_seed(s)) for s in seeds]
    for i, ln in enumerate(lens):
        seed, pub = keys[i % 16]
        msg = g.bytes(ln) if ln else b""
        sig = bytearray(R.ed25519_sign(seed, msg)) if i < 400 else None
        if sig is None:  # signing in Python is slow: reuse a signature on a different message (invalid) for the bulk
            sig = bytearray(sigs[i % 400])
        if i % 7 == 3:
            sig[5] ^= 0x10
        pubs.append(pub); msgs.append(msg); sigs.append(bytes(sig))
    got = ctx.ed25519_verify(rows(pubs), msgs, rows(sigs))
    ks = rows([R.ed25519_hash_k(s[:32], p_, m) for p_, m, s in zip(pubs, msgs, sigs)])
    sg = rows(sigs)
    want = coracle.ed25519_verify_prehashed(rows(pubs), sg[:, :32].copy(), sg[:, 32:].copy(), ks, threads(coracle))
    assert np.array_equal(got, want)
    assert got[:400].sum() > 300 and not got[400:].any()
    for i in (0, 1, 2, 3, 14, 399):
        assert bool(got[i]) == R.ed25519_verify(pubs[i], msgs[i], sigs[i])
    # same answers as the prehashed entry point
    assert np.array_equal(got, ctx.ed25519_verify_prehashed(rows(pubs), sg[:, :32].copy(), sg[:, 32:].copy(), ks))


@pytest.mark.parametrize("curve", ["p256r1", "p384r1"])
def test_ecdsa_verify_raw_messages(ctx, golden, coracle, curve):
    """ecdsa::verify(public, message, sig): SHA-2 + bits2int on the device, all three hash sizes."""
    c, v = R.WCURVES[curve], golden["ecdsa_rfc6979"][curve]
    Q = c.enc((int(v["qx"], 16), int(v["qy"], 16)))
    for kat in v["kats"]:
        bits = int(kat["alg"][3:])
        rs = int(kat["r"], 16).to_bytes(c.sbytes, "big") + int(kat["s"], 16).to_bytes(c.sbytes, "big")
        msg = kat["message"].encode()
        got = ctx.ecdsa_verify(curve, bits, rows([Q, Q, Q]), [msg, msg + b".", msg], rows([rs, rs, rs[:-1] + bytes([rs[-1] ^ 1])]))
        assert got.tolist() == [True, False, False], kat
    g = rng(6979)
    n = 600
    d = int.from_bytes(g.bytes(40), "little") % (c.n - 1) + 1
    q = c.enc(c.mul(d, c.G))
    lens = [0, 1, 55, 56, 63, 64, 65, 111, 112, 127, 128, 129, 400] + [int(x) for x in g.integers(0, 200, size=n - 13)]
    msgs = [g.bytes(l) if l else b"" for l in lens]
    for bits, hname in ((256, "sha256"), (384, "sha384"), (512, "sha512")):
        zs, rss = [], []
        for i, m in enumerate(msgs):
            z = R.ecdsa_digest_to_scalar(c, hashlib.new(hname, m).digest())
            k = int.from_bytes(g.bytes(60), "little") % (c.n - 1) + 1
            rs = bytearray(R.ecdsa_sign_hashed(c, d, k, int.from_bytes(z, "big"))) if i < 60 else bytearray(rss[i % 60])
            if i % 5 == 2:
                rs[7] ^= 4
            zs.append(z); rss.append(bytes(rs))
        got = ctx.ecdsa_verify(curve, bits, rows([q] * n), msgs, rows(rss))
        want = coracle.ecdsa_verify_hashed(curve, rows([q] * n), rows(zs), rows(rss), threads(coracle))
        assert np.array_equal(got, want) and got[:60].sum() >= 40 and not got[60:].any()
        assert np.array_equal(got, ctx.ecdsa_verify_hashed(curve, rows([q] * n), rows(zs), rows(rss)))


# ---- device-resident entry points ---------------------------------------------------------------------
def test_device_resident_entry_points_match_host_entry_points(ctx):
    torch = pytest.importorskip("torch")
    g = rng(4242)
    n = 5000
    kb = scalars_mod(g, n, R.L25519, 32, "little")
    d_k = torch.from_numpy(kb).cuda()
    d_out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.warm("ed25519_mul_base", n)
    ctx.warm("x25519", n)
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert ctx.dev_status(0) == (0, None)
    assert np.array_equal(d_out.cpu().numpy(), ctx.ed25519_mul_base(kb))
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), 0, d_out.data_ptr(), st)   # n = 0: a no-op, as the host path
    k, u = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
    d_o = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    d_kk, d_u = torch.from_numpy(k).cuda(), torch.from_numpy(u).cuda()  # keep alive: the allocator reuses freed blocks
    ctx.dev_call("ecb_x25519_dev", 0, d_kk.data_ptr(), d_u.data_ptr(), n, d_o.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), ctx.x25519(k, u))
    # invalid element is reported through ecb_dev_status
    kb2 = kb.copy(); kb2[77] = 0xFF
    d_k2 = torch.from_numpy(kb2).cuda()
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k2.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert ctx.dev_status(0) == (-3, 77)


def test_device_resident_entry_points_never_allocate():
    """*_dev calls only enqueue: without ecb_warm (no comb table, no work buffers) they fail with ECB_ERR_NOT_READY
    instead of building a table inside a call documented as asynchronous; a warm for a smaller batch does not
    cover a larger one; after ecb_warm they run, and ecb_get_info reports what was built."""
    torch = pytest.importorskip("torch")
    from eccoxide_b200 import Context, EccBatchError

    g = rng(99)
    n = 1 << 21   # the two-kernel form: 201 MB of projective planes, more than the W = 16 table build leaves behind
    kb = np.tile(scalars_mod(g, 1 << 12, R.L25519, 32, "little"), (n >> 12, 1))
    d_k = torch.from_numpy(kb).cuda()
    d_out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    with Context() as c:
        c.set_option("ed25519_comb_w", 16)
        assert c.get_info("ed25519_comb_w") == 0
        with pytest.raises(EccBatchError) as e:
            c.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), 1 << 12, d_out.data_ptr(), st)
        assert e.value.code == -6
        c.warm("ed25519_mul_base", 1 << 12)
        assert c.get_info("ed25519_comb_w") == 16 and c.get_info("ed25519_comb_windows") == 16
        c.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), 1 << 12, d_out.data_ptr(), st)
        with pytest.raises(EccBatchError) as e:   # a warm for 2^12 does not cover 2^21
            c.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), n, d_out.data_ptr(), st)
        assert e.value.code == -6
        c.warm("ed25519_mul_base", n)
        c.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), n, d_out.data_ptr(), st)
        torch.cuda.synchronize()
        assert c.dev_status(0) == (0, None)
        assert np.array_equal(d_out[:4096].cpu().numpy(), c.ed25519_mul_base(kb[:4096]))
        with pytest.raises(EccBatchError) as e:
            c.dev_call("ecb_wei_mul_base_dev", 0, 0, d_k.data_ptr(), 1024, d_out.data_ptr(), 0, st)
        assert e.value.code == -6
        with pytest.raises(EccBatchError):
            c.warm("no_such_op", 16)


def test_device_resident_large_batch_is_split_over_streams_and_stays_exact(ctx, coracle):
    """With option dev_split, n >= 3 * 2^16 takes the fork/join path (three sub-batches on the slot streams)."""
    torch = pytest.importorskip("torch")
    g = rng(777)
    n = (1 << 18) + 12345
    kb = rand_bytes(g, n, 32)
    kb[:, 31] &= 0x0F
    d_k = torch.from_numpy(kb).cuda()
    d_out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ctx.set_option("dev_split", 1)
    ctx.warm("ed25519_mul_base", n)
    ctx.warm("wei_mul_base", n, "p256r1")
    ctx.set_option("dev_split", 0)
    ctx.warm("ed25519_mul_base", n)
    ctx.set_option("dev_split", 1)
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert ctx.dev_status(0) == (0, None)
    got = d_out.cpu().numpy()
    idx = np.concatenate([np.arange(64), np.arange(n // 3 - 32, n // 3 + 32), np.arange(2 * n // 3 - 32, 2 * n // 3 + 32), np.arange(n - 64, n),
                          g.integers(0, n, size=2048)])
    assert np.array_equal(got[idx], coracle.ed25519_mul_base(kb[idx], threads(coracle)))
    ctx.set_option("dev_split", 0)
    d_out2 = torch.empty_like(d_out)
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k.data_ptr(), n, d_out2.data_ptr(), st)
    torch.cuda.synchronize()
    ctx.set_option("dev_split", 1)
    assert torch.equal(d_out, d_out2)
    # an invalid scalar in the last sub-batch is reported with its absolute index
    bad = n - 1000
    kb2 = kb.copy(); kb2[bad] = 0xFF
    d_k2 = torch.from_numpy(kb2).cuda()
    ctx.dev_call("ecb_ed25519_mul_base_dev", 0, d_k2.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    assert ctx.dev_status(0) == (-3, bad)
    # p256 fixed base through the same path
    kp = rand_bytes(g, n, 32); kp[:, 0] &= 0x7F
    d_kp = torch.from_numpy(kp).cuda()
    d_xy = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); d_inf = torch.empty((n,), dtype=torch.uint8, device="cuda")
    ctx.dev_call("ecb_wei_mul_base_dev", 0, 0, d_kp.data_ptr(), n, d_xy.data_ptr(), d_inf.data_ptr(), st)
    torch.cuda.synchronize()
    assert ctx.dev_status(0) == (0, None)
    sel = g.integers(0, n, size=1024)
    exp, einf = coracle.wei_mul_base("p256r1", kp[sel], threads(coracle))
    assert np.array_equal(d_xy.cpu().numpy()[sel], exp) and np.array_equal(d_inf.cpu().numpy()[sel].astype(bool), einf)
    ctx.set_option("dev_split", 0)


def test_context_over_all_visible_devices_shards_by_contiguous_slice(coracle):
    """ecb_init over every visible GPU: one host thread + streams per device, slice [g n/G, (g+1) n/G)."""
    torch = pytest.importorskip("torch")
    from eccoxide_b200 import Context, EccBatchError

    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("MULTI-DEVICE SHARDING NOT EXERCISED: this lease shows %d GPU; run with gpurun --gpus 2 (profiles/ records the runs that did)" % ndev)
    g = rng(31337)
    n = 100003
    with Context(devices=list(range(ndev))) as c:
        assert c.device_count() == ndev
        kb = scalars_mod(g, n, R.L25519, 32, "little")
        assert np.array_equal(c.ed25519_mul_base(kb), coracle.ed25519_mul_base(kb, threads(coracle)))
        k, u = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
        assert np.array_equal(c.x25519(k, u), coracle.x25519(k, u, threads(coracle)))
        # multi-scalar multiplication: every device reduces its slice, the partial sums meet on the first device
        cb = R.WCURVES["bls12_381_g1"]
        m, period = 1 << 15, 256
        tb = scalars_mod(g, period, cb.n, 32, "big")
        basepts, _ = c.wei_mul_base("bls12_381_g1", tb)
        km = scalars_mod(g, m, cb.n, 32, "big")
        tv = [int.from_bytes(r.tobytes(), "big") for r in tb]
        total = sum(int.from_bytes(km[i].tobytes(), "big") * tv[i % period] for i in range(m)) % cb.n
        want, _ = c.wei_mul_base("bls12_381_g1", rows([total.to_bytes(32, "big")]))
        got, ginf = c.wei_msm("bls12_381_g1", km, np.ascontiguousarray(np.tile(basepts, (m // period, 1))))
        assert not ginf and got.tobytes() == want[0].tobytes()
        kb[n - 7] = 0xFF  # first offender reported with its global index whichever device owns it
        kb[n // 2 + 1] = 0xFF
        with pytest.raises(EccBatchError) as e:
            c.ed25519_mul_base(kb)
        assert e.value.code == -3 and e.value.bad_index == n // 2 + 1


def test_host_entry_points_across_pipeline_chunks(ctx, coracle):
    """A batch spanning several pipeline chunks (default 189 440 elements, 4 stream slots): exact
    results at the chunk borders and the first offender reported with its global index."""
    from eccoxide_b200 import EccBatchError

    g = rng(2718)
    n = 3 * 189440 + 777
    kb = rand_bytes(g, n, 32)
    kb[:, 31] &= 0x0F
    got = ctx.ed25519_mul_base(kb)
    # every element: the chunk borders move with the schedule (ramped chunk sizes, option "ramp")
    assert np.array_equal(got, coracle.ed25519_mul_base(kb, threads(coracle)))
    ctx.set_option("ramp", 0)
    try:
        assert np.array_equal(ctx.ed25519_mul_base(kb), got)
    finally:
        ctx.set_option("ramp", 2)
    kb[2 * 189440 + 5] = 0xFF
    kb[3 * 189440 + 9] = 0xFF
    with pytest.raises(EccBatchError) as e:
        ctx.ed25519_mul_base(kb)
    assert e.value.code == -3 and e.value.bad_index == 2 * 189440 + 5
    # a small chunk size exercises slot reuse many times
    ctx.set_option("chunk", 1000)
    try:
        k, u = rand_bytes(g, 25000, 32), rand_bytes(g, 25000, 32)
        assert np.array_equal(ctx.x25519(k, u), coracle.x25519(k, u, threads(coracle)))
    finally:
        ctx.set_option("chunk", 189440)


# ---- Ed25519 key generation and signing ----------------------------------------------
def test_ed25519_keygen_and_sign(ctx, golden):
    """SecretKey::public_key / Keypair::sign / SecretKey::sign (ed25519.rs:61-120) through the C ABI:
    RFC 8032 TEST 1-3 (seed -> public key -> signature, ed25519.rs:271-290); 3000 random seeds with
    ragged messages bit-exact against OpenSSL's deterministic Ed25519 (same RFC) and, sampled, the
    big-int oracle; every signature accepted by the library's own verification (which the parity
    tests above pin to the reference) and rejected once the message changes."""
    kats = golden["ed25519_rfc8032"]
    seeds = [bytes.fromhex(v["seed"]) for v in kats]
    msgs = [bytes.fromhex(v["message"]) for v in kats]
    g = rng(8032)
    n = 3000
    for i in range(n - len(kats)):
        seeds.append(g.bytes(32))
        msgs.append(g.bytes(int(g.integers(0, 300)) if i % 50 else 0))
    sd = rows(seeds)
    pub = ctx.ed25519_public_from_seed(sd)
    for i, v in enumerate(kats):
        assert pub[i].tobytes().hex() == v["public"]
    sig = ctx.ed25519_sign(sd, msgs, pub=pub)
    sig2 = ctx.ed25519_sign(sd, msgs)          # SecretKey::sign: A derived on the device
    assert np.array_equal(sig, sig2)
    for i, v in enumerate(kats):
        assert sig[i].tobytes().hex() == v["signature"]
    for i in (3, 4, 5, 53, 1000, n - 1):
        assert pub[i].tobytes() == R.ed25519_public_from_seed(seeds[i])
        assert sig[i].tobytes() == R.ed25519_sign(seeds[i], msgs[i])
    ed = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ed25519")
    from cryptography.hazmat.primitives import serialization as ser

    for i in range(n):
        key = ed.Ed25519PrivateKey.from_private_bytes(seeds[i])
        assert key.public_key().public_bytes(ser.Encoding.Raw, ser.PublicFormat.Raw) == pub[i].tobytes(), i
        assert key.sign(msgs[i]) == sig[i].tobytes(), i
    assert ctx.ed25519_verify(pub, msgs, sig).all()
    tampered = [m + b"!" for m in msgs]
    assert not ctx.ed25519_verify(pub, tampered, sig).any()
    assert ctx.ed25519_public_from_seed(np.zeros((0, 32), dtype=np.uint8)).shape == (0, 32)
    # the calls above ran the constant-time kernels (the entry points under the reference's names); the
    # *_vartime forms — large comb indexed by the digits, safegcd — must give the same bytes
    assert np.array_equal(ctx.ed25519_public_from_seed(sd, vartime=True), pub)
    assert np.array_equal(ctx.ed25519_sign(sd, msgs, pub=pub, vartime=True), sig)
    assert np.array_equal(ctx.ed25519_sign(sd, msgs, vartime=True), sig)


def test_ed25519_mul_base_constant_time_kernel(ctx, coracle, golden):
    """ecb_ed25519_mul_base_ct (csrc/ct.cuh: masked scans of a 48 KB comb staged in shared memory, Fermat
    inversion): Point::mul_base bytes for the reference's edge scalars and a 2^16 batch, 100 % against the C oracle;
    a non-canonical scalar is refused with its index as in the variable-time form."""
    from eccoxide_b200 import EccBatchError

    g = rng(0xC7)
    kb = np.concatenate([rows([v.to_bytes(32, "little") for v in ed_edge_scalars(golden)]), scalars_mod(g, (1 << 16) - 7, R.L25519, 32, "little")])
    got = ctx.ed25519_mul_base_ct(kb)
    assert np.array_equal(got, coracle.ed25519_mul_base(kb, threads(coracle)))
    assert np.array_equal(got[:5000], ctx.ed25519_mul_base(kb[:5000]))
    for n in (1, 127, 129):
        assert np.array_equal(ctx.ed25519_mul_base_ct(kb[:n]), got[:n])
    bad = kb[:300].copy()
    bad[211] = 0xFF
    with pytest.raises(EccBatchError) as e:
        ctx.ed25519_mul_base_ct(bad)
    assert e.value.code == -3 and e.value.bad_index == 211


@pytest.mark.parametrize("curve", ["p256r1", "p384r1"])
def test_ecdsa_sign_hashed(ctx, golden, coracle, curve):
    """ecdsa::sign_hashed (ecdsa.rs:165-184) through the C ABI: RFC 6979 A.2.5 / A.2.6 (d, k, message ->
    r, s), 2000 random (d, k, z) bit-exact with the big-int oracle (sampled) and accepted by the
    library's verification, the C oracle and OpenSSL; zero / non-canonical secrets and nonces refused."""
    c = R.WCURVES[curve]
    v = golden["ecdsa_rfc6979"][curve]
    d0 = int(v["d"], 16)
    ds, ks, zs, want = [], [], [], []
    for kat in v["kats"]:
        dg = hashlib.new(kat["alg"], kat["message"].encode()).digest()
        ds.append(d0)
        ks.append(int(kat["k"], 16))
        zs.append(int.from_bytes(R.ecdsa_digest_to_scalar(c, dg), "big"))
        want.append(kat["r"].rjust(2 * c.sbytes, "0") + kat["s"].rjust(2 * c.sbytes, "0"))
    nk = len(want)
    g = rng(6979 + c.sbytes)
    n_rand = 2000
    for _ in range(n_rand):
        ds.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1)
        ks.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1)
        zs.append(int.from_bytes(g.bytes(c.sbytes + 8), "big") % c.n)
    bad = [(0, 5, 7), (5, 0, 7), (c.n, 5, 7), (5, c.n, 7), (5, 6, c.n), (2**(8 * c.sbytes) - 1, 3, 3)]
    for d, kk, z in bad:
        ds.append(d)
        ks.append(kk)
        zs.append(z)
    tob = lambda xs: rows([x.to_bytes(c.sbytes, "big") for x in xs])
    db, kb, zb = tob(ds), tob(ks), tob(zs)
    rs, ok = ctx.ecdsa_sign_hashed(curve, db, kb, zb)          # the constant-time kernels (the reference's names)
    rs_v, ok_v = ctx.ecdsa_sign_hashed(curve, db, kb, zb, vartime=True)
    assert np.array_equal(rs, rs_v) and np.array_equal(ok, ok_v)
    m = nk + n_rand
    assert ok[:m].all() and not ok[m:].any() and not rs[m:].any()
    for i in range(nk):
        assert rs[i].tobytes().hex() == want[i]
    for i in list(range(nk, nk + 40)) + [m - 1]:
        assert rs[i].tobytes() == R.ecdsa_sign_hashed(c, ds[i], ks[i], zs[i])
    q, inf = ctx.wei_mul_base(curve, db[:m])
    assert not inf.any()
    assert ctx.ecdsa_verify_hashed(curve, q, zb[:m], rs[:m]).all()
    assert coracle.ecdsa_verify_hashed(curve, q, zb[:m], rs[:m], threads(coracle)).all()
    zb2 = zb[:m].copy()
    zb2[:, -1] ^= 1
    assert not ctx.ecdsa_verify_hashed(curve, q, zb2, rs[:m]).any()
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.asymmetric.utils import Prehashed, encode_dss_signature

    crv, alg = (ec.SECP256R1(), hashes.SHA256()) if curve == "p256r1" else (ec.SECP384R1(), hashes.SHA384())
    for i in range(nk, nk + 64):
        x, y = c.dec(q[i].tobytes())
        pub = ec.EllipticCurvePublicNumbers(x, y, crv).public_key()
        r_, s_ = int.from_bytes(rs[i, : c.sbytes].tobytes(), "big"), int.from_bytes(rs[i, c.sbytes:].tobytes(), "big")
        pub.verify(encode_dss_signature(r_, s_), zb[i].tobytes(), ec.ECDSA(Prehashed(alg)))   # raises on a bad signature


@pytest.mark.parametrize("curve", ["p256r1", "p384r1"])
def test_ecdsa_sign_raw_messages(ctx, golden, curve):
    """ecdsa::sign (ecdsa.rs:192) with the hash on the device: the RFC 6979 vectors straight from their
    messages for SHA-256 / 384 / 512 (ecdsa.rs:808-878), ragged random messages against hashlib + the
    big-int oracle, and acceptance by the library's raw-message verification."""
    c = R.WCURVES[curve]
    v = golden["ecdsa_rfc6979"][curve]
    d0 = int(v["d"], 16)
    g = rng(1979 + c.sbytes)
    assert len([k for k in v["kats"] if k["alg"] in ("sha256", "sha384", "sha512")]) == len(v["kats"]) >= 3
    for bits in (256, 384, 512):
        kats = [k for k in v["kats"] if k["alg"] == "sha%d" % bits]   # (a curve's vectors need not use every hash)
        msgs = [k["message"].encode() for k in kats] + [g.bytes(int(g.integers(0, 200))) for _ in range(300)] + [b""]
        ds = [d0] * len(kats) + [int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1 for _ in range(301)]
        ks = [int(k["k"], 16) for k in kats] + [int.from_bytes(g.bytes(c.sbytes + 8), "big") % (c.n - 1) + 1 for _ in range(301)]
        tob = lambda xs: rows([x.to_bytes(c.sbytes, "big") for x in xs])
        rs, ok = ctx.ecdsa_sign(curve, tob(ds), tob(ks), msgs, hash_bits=bits)
        assert ok.all()
        for i, k in enumerate(kats):
            assert rs[i].tobytes().hex() == k["r"].rjust(2 * c.sbytes, "0") + k["s"].rjust(2 * c.sbytes, "0")
        for i in range(len(msgs)):
            z = int.from_bytes(R.ecdsa_digest_to_scalar(c, hashlib.new("sha%d" % bits, msgs[i]).digest()), "big")
            assert rs[i].tobytes() == R.ecdsa_sign_hashed(c, ds[i], ks[i], z), (bits, i)
        q, _ = ctx.wei_mul_base(curve, tob(ds))
        assert ctx.ecdsa_verify(curve, bits, q, msgs, rs).all()


# ---- wire formats either side of the path -------------------------------------------
@pytest.mark.parametrize("curve", CURVES)
def test_wei_decompress(ctx, coracle, golden, curve):
    """PointAffine::decompress (affine.rs:48) through the C ABI: x-coordinates of random points with
    both parities, the NIST KAT points, random field elements (about half are not x-coordinates),
    x = 0, x = p, x = 2^(8 FB) - 1; bit-exact with the oracle, zero bytes where the reference gives None."""
    c = R.WCURVES[curve]
    g = rng(SEEDS["p256"] + 31 + len(curve))
    n = 3000
    pts = wei_points(curve, g, n)
    x = np.ascontiguousarray(pts[:, : c.fbytes])
    sign = g.integers(0, 2, n).astype(np.uint8)
    rnd = rows([(int.from_bytes(g.bytes(c.fbytes + 8), "big") % c.p).to_bytes(c.fbytes, "big") for _ in range(200)])
    x[100:300] = rnd
    x[0] = 0
    x[1] = np.frombuffer(c.p.to_bytes(c.fbytes, "big"), dtype=np.uint8)
    x[2] = 0xFF
    kat = {"p256r1": golden["nist_p256"], "p384r1": golden["nist_p384"], "p256k1": golden["p256k1_sage"]}.get(curve)
    if kat:
        for i, e in enumerate(kat[:20]):
            x[10 + i] = np.frombuffer(bytes.fromhex(e["x"]), dtype=np.uint8)
            sign[10 + i] = int(e["y"], 16) & 1
    got, ok = ctx.wei_decompress(curve, x, sign)
    exp, eok = coracle.wei_decompress(curve, x, sign, threads(coracle))
    assert np.array_equal(ok, eok) and np.array_equal(got, exp)
    assert not ok[1] and not ok[2] and not got[1].any() and 40 < int((~ok[100:300]).sum()) < 160
    if kat:
        for i, e in enumerate(kat[:20]):
            assert ok[10 + i] and got[10 + i].tobytes().hex() == e["x"] + e["y"]
    # the points the library itself produces decompress back to themselves
    assert ok[300:].all() and np.array_equal(got[300:][sign[300:] == (pts[300:, -1] & 1)], pts[300:][sign[300:] == (pts[300:, -1] & 1)])
    assert ctx.wei_decompress(curve, np.zeros((0, c.fbytes), dtype=np.uint8), np.zeros(0, dtype=np.uint8))[0].shape == (0, 2 * c.fbytes)


def test_p256k1_reference_kats(ctx, coracle, golden):
    """secp256k1 — the reference's p256k1 (src/curve/sec2/p256k1.rs; the a = 0 path shared with BLS12-381 G1) — through
    the C ABI: the reference's Sage-generated k G for k = 1..100 (src/tests/sage.rs) by the comb and by the
    variable-base windows on G, 4000 random (k, P) against the C oracle, OpenSSL's secp256k1 public keys."""
    c = R.WCURVES["p256k1"]
    kats = golden["p256k1_sage"]
    ks = rows([e["k"].to_bytes(32, "big") for e in kats])
    want = rows([H(e["x"] + e["y"]) for e in kats])
    got, inf = ctx.wei_mul_base("p256k1", ks)
    assert not inf.any() and np.array_equal(got, want)
    gen = np.tile(np.frombuffer(c.enc(c.G), dtype=np.uint8), (len(kats), 1))
    got, inf = ctx.wei_mul("p256k1", ks, gen)
    assert not inf.any() and np.array_equal(got, want)
    g = rng(0x256C1)
    kb = scalars_mod(g, 4000, c.n, 32, "big")
    pts = wei_points("p256k1", g, 4000)
    got, inf = ctx.wei_mul("p256k1", kb, pts)
    exp, einf = coracle.wei_mul("p256k1", kb, pts, nthreads=threads(coracle))
    assert np.array_equal(got, exp) and np.array_equal(inf, einf)
    ec = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ec")
    pub, _ = ctx.wei_mul_base("p256k1", kb[:64])
    for i in range(64):
        nums = ec.derive_private_key(int.from_bytes(kb[i].tobytes(), "big"), ec.SECP256K1()).public_key().public_numbers()
        assert pub[i].tobytes() == nums.x.to_bytes(32, "big") + nums.y.to_bytes(32, "big")


def test_bls_g1_small_order_and_mixed_order_inputs(ctx, coracle):
    """Point::mul on BLS12-381 accepts any point of E(Fp) (g1.rs:375-377): the points of order 3, (0, +-2), and points of
    order 3 r.  For those the per-thread table of multiples holds the identity (3 P = O), which the Jacobian addition
    has to treat as such: every digit pattern, including scalars = 0, 1, 2 mod 3 and window digits 3 and 6."""
    from helpers import bls_cofactor_points

    c = R.BLSG1
    g = rng(0xB153)
    pts_list = bls_cofactor_points()
    ks = [0, 1, 2, 3, 4, 6, 10, 0x33, 0x63, 0x123456789ABCDEF, c.n - 1, c.n - 2] + [int.from_bytes(g.bytes(40), "big") % c.n for _ in range(60)]
    kb = rows([k.to_bytes(32, "big") for k in ks for _ in pts_list])
    pts = rows([c.enc(P) for _ in ks for P in pts_list])
    got, inf = ctx.wei_mul("bls12_381_g1", kb, pts)
    exp, einf = coracle.wei_mul("bls12_381_g1", kb, pts, nthreads=threads(coracle))
    assert np.array_equal(inf, einf) and np.array_equal(got, exp)
    for i in (4, 5, 8, 9, 20, 41):     # the big-integer group law agrees: (k mod 3) P for the order-3 points
        assert (got[i].tobytes(), int(inf[i])) == R.wei_mul(c, kb[i].tobytes(), pts[i].tobytes())
    assert inf[:4].all() and inf[12] and inf[13]      # k = 0; 3 * (0, +-2) = O


def test_ed25519_decompress(ctx, golden):
    """decode_point / Point::decompress (ed25519.rs:38-59, curve25519.rs:772) as a batch entry point: encodings of
    library-produced points round-trip, RFC 8032 public keys decode to points that satisfy the curve, random
    32-byte strings (about half are not points), non-canonical y, x = 0 with the sign bit; bit-exact with the
    big-integer oracle, zero bytes where the reference returns None; the output feeds ecb_ed25519_mul."""
    g = rng(0xDEC0)
    kb = scalars_mod(g, 600, R.L25519, 32, "little")
    xy = ctx.ed25519_mul_base(kb)
    enc = ctx.ed25519_mul_base(kb, compressed=True)
    rnd = rand_bytes(g, 400, 32)
    bad = [(R.P25519).to_bytes(32, "little"), (R.P25519 + 1).to_bytes(32, "little"), bytes([1] + [0] * 30 + [0x80]),
           (R.P25519 - 1 | (1 << 255)).to_bytes(32, "little"), b"\xff" * 32]
    kat = [bytes.fromhex(v["public"]) for v in golden["ed25519_rfc8032"]]
    allenc = np.concatenate([enc, rnd, rows(bad), rows(kat)])
    out, ok = ctx.ed25519_decompress(allenc)
    assert ok[:600].all() and np.array_equal(out[:600], xy)
    assert not ok[1000:1005].any() and not out[1000:1005].any() and ok[1005:].all()
    for i in range(600, allenc.shape[0]):
        P = R.ed_decode(allenc[i].tobytes())
        if P is None:
            assert not ok[i] and not out[i].any(), i
        else:
            assert ok[i] and out[i].tobytes() == P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little"), i
    assert 100 < int(ok[600:1000].sum()) < 300
    k2 = scalars_mod(g, 600, R.L25519, 32, "little")
    assert np.array_equal(ctx.ed25519_mul(k2, out[:600]), ctx.ed25519_mul(k2, xy))
    assert ctx.ed25519_decompress(np.zeros((0, 32), dtype=np.uint8))[0].shape == (0, 64)


def test_bls12_381_g1_mul_through_the_endomorphism(ctx, coracle):
    """Option bls12_381_g1_glv: for points of G1 ecb_wei_mul gives the same bytes with and without it (2^15 random
    pairs + edge scalars + identity inputs), equal to the C oracle on a sample; refusals keep their index; the option
    is off by default and off again afterwards (off-subgroup inputs are only defined without it)."""
    from eccoxide_b200 import EccBatchError

    c = R.WCURVES["bls12_381_g1"]
    g = rng(0x6171)
    n = 1 << 15
    xsq = 0xD201000000010000 ** 2
    base, binf = ctx.wei_mul_base("bls12_381_g1", scalars_mod(g, 1 << 10, c.n, 32, "big"))
    assert not binf.any()
    pts = np.ascontiguousarray(np.tile(base, (n // base.shape[0], 1)))
    kb = scalars_mod(g, n, c.n, 32, "big")
    edge = [0, 1, c.n - 1, xsq - 1, xsq, xsq + 1, 2 * xsq, (xsq - 1) * xsq, (1 << 128) - 1, 1 << 128, (1 << 254) + 1]
    for i, v in enumerate(edge):
        kb[i] = np.frombuffer(v.to_bytes(32, "big"), dtype=np.uint8)
    inf_in = np.zeros(n, dtype=np.uint8)
    inf_in[40:44] = 1
    plain, pinf = ctx.wei_mul("bls12_381_g1", kb, pts, inf_in=inf_in)
    try:
        ctx.set_option("bls12_381_g1_glv", 1)
        fast, finf = ctx.wei_mul("bls12_381_g1", kb, pts, inf_in=inf_in)
        bad = kb.copy()
        bad[77] = 0xFF
        with pytest.raises(EccBatchError) as e:
            ctx.wei_mul("bls12_381_g1", bad, pts, inf_in=inf_in)
        assert e.value.code == -3 and e.value.bad_index == 77
        badp = pts.copy()
        badp[91, -1] ^= 1
        with pytest.raises(EccBatchError) as e:
            ctx.wei_mul("bls12_381_g1", kb, badp, inf_in=inf_in)
        assert e.value.code == -4 and e.value.bad_index == 91
    finally:
        ctx.set_option("bls12_381_g1_glv", 0)
    assert np.array_equal(fast, plain) and np.array_equal(finf, pinf)
    assert finf[0] and finf[40:44].all() and not finf[1:40].any()
    m = 2048
    exp, einf = coracle.wei_mul("bls12_381_g1", kb[:m], pts[:m], inf_in=inf_in[:m], nthreads=os.cpu_count() or 1)
    assert np.array_equal(fast[:m], exp) and np.array_equal(finf[:m], einf)


@pytest.mark.parametrize("curve", ["bls12_381_g1", "p256k1"])
def test_wei_msm(ctx, coracle, curve):
    """ecb_wei_msm (csrc/msm.cuh, the bucket method): sum_i k_i P_i.  Small batches against the big-integer group law
    (duplicates, P and -P with equal scalars, zero scalars, n = 0 and 1, and on BLS12-381 points outside G1); 2^18
    points P_i = t_i G against (sum k_i t_i mod n) G from the generator comb — an exact check of the whole sum;
    invalid scalars / points are refused with their index."""
    from eccoxide_b200 import EccBatchError
    from helpers import bls_cofactor_points

    c = R.WCURVES[curve]
    g = rng(0x3530 + len(curve))
    pts = [c.mul(int.from_bytes(g.bytes(40), "big") % c.n or 1, c.G) for _ in range(60)]
    if curve == "bls12_381_g1":
        pts += bls_cofactor_points()
    ks = [int.from_bytes(g.bytes(40), "big") % c.n for _ in pts]
    pts += [pts[0], pts[1], c.neg(pts[2]), pts[3]]
    ks += [ks[0], ks[1], ks[2], 0]
    want = None
    for kk, P in zip(ks, pts):
        want = c.add(want, c.mul(kk, P)) if kk else want
    kb = rows([v.to_bytes(c.sbytes, "big") for v in ks])
    pb = rows([c.enc(P) for P in pts])
    out, inf = ctx.wei_msm(curve, kb, pb)
    assert not inf and out.tobytes() == c.enc(want)
    out, inf = ctx.wei_msm(curve, kb[:1], pb[:1])
    assert not inf and out.tobytes() == c.enc(c.mul(ks[0], pts[0]))
    out, inf = ctx.wei_msm(curve, kb[:0], pb[:0])
    assert inf and not out.any()
    out, inf = ctx.wei_msm(curve, rows([kb[0].tobytes()] * 2), rows([c.enc(pts[0]), c.enc(c.neg(pts[0]))]))
    assert inf and not out.any()
    # a large batch: P_i = t_(i mod period) G
    n, period = 1 << 18, 1 << 10
    tb = scalars_mod(g, period, c.n, c.sbytes, "big")
    base, binf = ctx.wei_mul_base(curve, tb)
    assert not binf.any()
    kbig = g.integers(0, 256, size=(n, c.sbytes), dtype=np.uint8)
    kbig[:, 0] &= 0x3F if curve == "bls12_381_g1" else 0x7F
    tv = [int.from_bytes(r.tobytes(), "big") for r in tb]
    total = sum(int.from_bytes(kbig[i].tobytes(), "big") * tv[i % period] for i in range(n)) % c.n
    exp, einf = ctx.wei_mul_base(curve, rows([total.to_bytes(c.sbytes, "big")]))
    out, inf = ctx.wei_msm(curve, kbig, np.ascontiguousarray(np.tile(base, (n // period, 1))))
    assert inf == bool(einf[0]) and out.tobytes() == exp[0].tobytes()
    assert out.tobytes() == coracle.wei_mul_base(curve, rows([total.to_bytes(c.sbytes, "big")]))[0][0].tobytes()
    # refusals
    kbad = kb.copy()
    kbad[7] = 0xFF
    with pytest.raises(EccBatchError) as e:
        ctx.wei_msm(curve, kbad, pb)
    assert e.value.code == -3 and e.value.bad_index == 7
    pbad = pb.copy()
    pbad[11, -1] ^= 1
    with pytest.raises(EccBatchError) as e:
        ctx.wei_msm(curve, kb, pbad)
    assert e.value.code == -4 and e.value.bad_index == 11


@pytest.mark.parametrize("curve", ["bls12_381_g1", "p256k1"])
def test_wei_msm_skewed_digit_distributions(ctx, coracle, curve):
    """The bucket sums are balanced over segments of the sorted index array, so the running time and the result must not
    depend on how the digits are distributed.  2^16 points P_i = t_i G with (a) scalars = 64 uniform bytes mod the order —
    on p256k1 the top window of such a scalar holds one bit, so half of all points share ONE bucket (a thread per bucket
    took 0.8 s on 2^18 points) —, (b) ONE scalar for every point (one bucket per window holds everything: the warp-level
    sum of a heavy bucket's pieces), (c) sixteen distinct scalars; each against (sum k_i t_i) G from the generator comb."""
    import time

    c = R.WCURVES[curve]
    g = rng(0x3541 + len(curve))
    n, period = 1 << 16, 1 << 9
    tb = scalars_mod(g, period, c.n, c.sbytes, "big")
    base, binf = ctx.wei_mul_base(curve, tb)
    assert not binf.any()
    pts = np.ascontiguousarray(np.tile(base, (n // period, 1)))
    tv = [int.from_bytes(r.tobytes(), "big") for r in tb]
    uniform = scalars_mod(g, n, c.n, c.sbytes, "big")
    one = np.ascontiguousarray(np.tile(scalars_mod(g, 1, c.n, c.sbytes, "big"), (n, 1)))
    few = scalars_mod(g, 16, c.n, c.sbytes, "big")[g.integers(0, 16, size=n)]
    times = {}
    for name, kb in (("uniform", uniform), ("one", one), ("few", np.ascontiguousarray(few))):
        total = sum(int.from_bytes(kb[i].tobytes(), "big") * tv[i % period] for i in range(n)) % c.n
        exp, einf = ctx.wei_mul_base(curve, rows([total.to_bytes(c.sbytes, "big")]))
        ctx.wei_msm(curve, kb, pts)
        t0 = time.perf_counter()
        out, inf = ctx.wei_msm(curve, kb, pts)
        times[name] = time.perf_counter() - t0
        assert inf == bool(einf[0]) and out.tobytes() == exp[0].tobytes(), name
    # no distribution may cost an order of magnitude more than the uniform one (a thread per bucket: x100 and more)
    assert max(times.values()) < 10 * times["uniform"] + 0.05, times


def test_ristretto255(ctx, golden):
    """ristretto255 (src/curve/curve25519/ristretto255.rs; RFC 9496) through the C ABI: mul_base gives the RFC's
    encodings of 0 B .. 15 B (:341-358), the 17 bad encodings are refused (:380-398) — by decompress with ok = 0 and
    by mul with the offender's index —, scale on encodings agrees with the big-integer oracle for 300 random (k, P)
    and with mul_base when P = B, and the encoding does not depend on the Edwards representative."""
    from eccoxide_b200 import EccBatchError

    v = golden["ristretto255"]
    mult = rows([bytes.fromhex(x) for x in v["multiples"]])
    bad = rows([bytes.fromhex(x) for x in v["bad"]])
    ks = rows([i.to_bytes(32, "little") for i in range(16)])
    assert np.array_equal(ctx.ristretto255_mul_base(ks), mult)
    xy, ok = ctx.ristretto255_decompress(np.concatenate([mult, bad]))
    assert ok[:16].all() and not ok[16:].any() and not xy[16:].any()
    assert np.array_equal(ctx.ristretto255_compress(xy[:16]), mult)
    g = rng(9496)
    n = 300
    kb = scalars_mod(g, n, R.L25519, 32, "little")
    pk = scalars_mod(g, n, R.L25519, 32, "little")
    enc = ctx.ristretto255_mul_base(pk)
    out = ctx.ristretto255_mul(kb, enc)
    for i in range(0, n, 7):
        assert enc[i].tobytes() == R.ristretto255_mul_base(pk[i].tobytes())
        assert out[i].tobytes() == R.ristretto255_mul(kb[i].tobytes(), enc[i].tobytes()), i
    gen = np.tile(mult[1], (n, 1))
    assert np.array_equal(ctx.ristretto255_mul(kb, gen), ctx.ristretto255_mul_base(kb))
    # the same group element through another Edwards representative (P + a point of order 4) encodes identically
    pts = ctx.ed25519_mul_base(pk[:64])
    t4 = (R.SQRT_M1, 0)
    moved = rows([b"".join(c.to_bytes(32, "little") for c in R.ed_add((int.from_bytes(r[:32].tobytes(), "little"), int.from_bytes(r[32:].tobytes(), "little")), t4)) for r in pts])
    assert np.array_equal(ctx.ristretto255_compress(moved), enc[:64])
    mixed = enc[:40].copy()
    mixed[23] = bad[3]
    with pytest.raises(EccBatchError) as e:
        ctx.ristretto255_mul(kb[:40], mixed)
    assert e.value.code == -4 and e.value.bad_index == 23
    kbad = kb[:40].copy()
    kbad[5] = 0xFF
    with pytest.raises(EccBatchError) as e:
        ctx.ristretto255_mul_base(kbad)
    assert e.value.code == -3 and e.value.bad_index == 5


def test_bls_g1_uncompressed_encodings(ctx, golden):
    """The 96-byte flavour (serialize.rs:330-420): the reference's uncompressed KATs (g1.rs:605-680) and OFF_SUBGROUP
    encodings (accepted by _oncurve_only, refused with the subgroup check), flag misuse, the identity encoding
    (refused as the reference's PointAffine cannot hold it, reported in `inf`), non-canonical and off-curve
    coordinates; to_uncompressed round trip including the identity."""
    c = R.BLSG1
    v = golden["bls12_381_g1"]
    kat = [bytes.fromhex(e["bytes"]) for e in v["uncompressed"]]
    off_u = [bytes.fromhex(o["uncompressed"]) for o in v["off_subgroup"]]
    g0 = kat[0]
    ident = bytes([0x40]) + bytes(95)
    bad = [bytes([g0[0] | 0x80]) + g0[1:], bytes([g0[0] | 0x20]) + g0[1:], bytes([0x40]) + bytes(94) + b"\x01", bytes([0x60]) + bytes(95),
           g0[:95] + bytes([g0[95] ^ 1]), c.p.to_bytes(48, "big") + g0[48:], g0[:48] + (c.p + 2).to_bytes(48, "big")]
    enc = rows(kat + off_u + [ident] + bad)
    nk, no = len(kat), len(off_u)
    out, ok, inf = ctx.bls12_381_g1_from_uncompressed(enc, True)
    assert list(ok) == [True] * nk + [False] * (no + 1 + len(bad)) and not out[nk:].any()
    assert list(inf) == [False] * (nk + no) + [True] + [False] * len(bad)
    out2, ok2, _ = ctx.bls12_381_g1_from_uncompressed(enc, False)
    assert list(ok2) == [True] * (nk + no) + [False] * (1 + len(bad)) and np.array_equal(out2[: nk + no], enc[: nk + no])
    for i in range(enc.shape[0]):
        for chk, o, k in ((True, out, ok), (False, out2, ok2)):
            want = R.bls_g1_from_uncompressed(enc[i].tobytes(), chk)
            assert (want is not None) == bool(k[i]) and (want is None or o[i].tobytes() == want), (i, chk)
    g = rng(0x96)
    pts, pinf = ctx.wei_mul_base("bls12_381_g1", np.concatenate([np.zeros((1, 32), dtype=np.uint8), scalars_mod(g, 300, c.n, 32, "big")]))
    assert pinf[0] and not pinf[1:].any()
    u = ctx.bls12_381_g1_to_uncompressed(pts, pinf)
    assert u[0].tobytes() == ident and np.array_equal(u[1:], pts[1:])
    back, bok, binf = ctx.bls12_381_g1_from_uncompressed(u, True)
    assert not bok[0] and binf[0] and bok[1:].all() and np.array_equal(back[1:], pts[1:])


def test_bls_g1_standard_encodings(ctx, coracle, golden):
    """bls12_381/serialize.rs through the C ABI: the reference's compressed KATs (g1.rs:605-680) and
    OFF_SUBGROUP encodings (:313-368), flag misuse, and a batch of library-produced points mixed with
    raw curve points outside G1: to_compressed -> from_compressed round trip, subgroup check on/off."""
    c = R.BLSG1
    v = golden["bls12_381_g1"]
    kat = [bytes.fromhex(e["bytes"]) for e in v["compressed"]]
    off_c = [bytes.fromhex(o["compressed"]) for o in v["off_subgroup"]]
    off_u = [bytes.fromhex(o["uncompressed"]) for o in v["off_subgroup"]]
    g0 = kat[0]
    bad = [bytes([g0[0] & 0x7F]) + g0[1:], bytes([0xC0]) + bytes(47), bytes([0xE0]) + bytes(47), bytes([0xC0]) + bytes(46) + b"\x01",
           bytes([0x80 | (c.p >> 376)]) + (c.p & ((1 << 376) - 1)).to_bytes(47, "big")]
    enc = rows(kat + off_c + bad)
    out, ok = ctx.bls12_381_g1_from_compressed(enc, True)
    assert list(ok) == [True] * 5 + [False] * (3 + len(bad)) and not out[5:].any()
    for i, e in enumerate(v["compressed"]):
        assert out[i].tobytes() == R.wei_mul_base(c, e["k"].to_bytes(32, "big"))[0]
    out, ok = ctx.bls12_381_g1_from_compressed(enc, False)
    assert list(ok) == [True] * 8 + [False] * len(bad)
    assert [out[5 + i].tobytes() for i in range(3)] == off_u
    assert [r.tobytes() for r in ctx.bls12_381_g1_to_compressed(rows(off_u))] == off_c
    # batch: k*G from the comb (in G1), every 7th replaced by a raw curve point (outside G1 almost surely)
    g = rng(4711)
    n = 2000
    ks = scalars_mod(g, n, c.n, 32, "big")
    ks[3] = 0
    pts, inf = ctx.wei_mul_base("bls12_381_g1", ks)
    raw = []
    x = 1000
    while len(raw) < (n + 6) // 7:
        x += 1
        rhs = (x**3 + 4) % c.p
        y = pow(rhs, (c.p + 1) // 4, c.p)
        if y * y % c.p == rhs:
            raw.append(c.enc((x, y if len(raw) & 1 else c.p - y)))
    pts[::7] = rows(raw)
    inf[::7] = False
    inf[3] = True
    enc = ctx.bls12_381_g1_to_compressed(pts, inf)
    assert np.array_equal(enc, coracle.bls12_381_g1_to_compressed(pts, inf, threads(coracle)))
    assert enc[3].tobytes() == bytes([0xC0]) + bytes(47)
    for check in (True, False):
        got, ok = ctx.bls12_381_g1_from_compressed(enc, check)
        exp, eok = coracle.bls12_381_g1_from_compressed(enc, check, threads(coracle))
        assert np.array_equal(ok, eok) and np.array_equal(got, exp), check
        fin = ~inf
        assert not ok[3]
        if check:
            assert not ok[::7].any() and ok[fin & (np.arange(n) % 7 != 0)].all()
        else:
            assert ok[fin].all() and np.array_equal(got[fin], pts[fin])


@pytest.mark.parametrize("mode", [0, 2])
def test_batch_inversion_forms_agree(ctx, coracle, mode):
    """Option inv_block: one inversion per thread (0) or one per block (2) for EVERY field — the
    default picks per field; both forms must give the reference's bytes, including Z = 0 elements
    (zero scalars -> infinity) and ragged sizes that leave threads of the last block without work."""
    g = rng(4242 + mode)
    ctx.set_option("inv_block", mode)
    try:
        for n in (1, 130, 5000):
            kb = scalars_mod(g, n, R.L25519, 32, "little")
            kb[0] = 0
            assert np.array_equal(ctx.ed25519_mul_base(kb), coracle.ed25519_mul_base(kb))
            k, u = rand_bytes(g, n, 32), rand_bytes(g, n, 32)
            u[n // 2] = 0                                    # z2 = 0 -> output 0
            assert np.array_equal(ctx.x25519(k, u), coracle.x25519(k, u))
        for curve in CURVES:
            c = R.WCURVES[curve]
            ks = scalars_mod(g, 300, c.n, c.sbytes, "big")
            ks[3] = 0
            want, winf = coracle.wei_mul_base(curve, ks)
            got, inf = ctx.wei_mul_base(curve, ks)
            assert np.array_equal(inf, winf) and np.array_equal(got, want), curve
        k, u = rand_bytes(g, 200, 56), rand_bytes(g, 200, 56)
        assert np.array_equal(ctx.x448(k, u), coracle.x448(k, u))
    finally:
        ctx.set_option("inv_block", 1)


def test_options_are_validated(ctx):
    from eccoxide_b200 import EccBatchError

    for key, val in (("no_such_option", 1), ("ed25519_comb_w", 3), ("ed25519_comb_w", 27), ("chunk", 0), ("p256r1_comb_w", 99), ("inv_per_thread", 0), ("inv_block", 3), ("ramp", 5)):
        with pytest.raises(EccBatchError) as e:
            ctx.set_option(key, val)
        assert e.value.code == -2
