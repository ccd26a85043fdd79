#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the hot path into tests/golden/.

Run in the build container (where /root/reference is mounted):
    python tools/extract_reference_vectors.py
It only *reads* literals out of the reference's test and parameter files (no code is copied) and
writes tests/golden/reference_vectors.json, which is committed so the tests never need
/root/reference at run time (the GPU box does not have it).

Sources (paths relative to /root/reference):
  src/tests/kats_data.rs            NIST point-multiplication KATs, P-256 = KATS[104:156], P-384 = [156:208]
                                    (index ranges from src/tests/kats.rs:35-38)
  src/protocol/x25519.rs:118-160    RFC 7748 §5.2 / §6.1
  src/curve/curve25519.rs:1629-1643 ladder u-coordinates of k*(u=9), k = 2, 5, 7
  src/protocol/ed25519.rs:271-290   RFC 8032 §7.1 TEST 1-3
  src/protocol/ecdsa.rs:808-878     RFC 6979 A.2.5 (P-256), A.2.6 (P-384)
  src/curve/bls12_381/g1.rs:605-680 compressed / uncompressed k*G; :313-368 OFF_SUBGROUP encodings
  src/protocol/x448.rs:116-160      RFC 7748 §5.2 / §6.2
  src/params/comb/*.rs              generator comb tables: SHA-256 of the whole table + sample entries
  src/params/sec2.rs, bls12_381.rs  domain parameters
  src/curve/curve25519.rs:1392-1429, src/tests/completeness.rs:68-93   edge-scalar lists
"""
import hashlib
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_vectors.json")


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def bytes_of(block):
    return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{1,2})\b", block))


def const_bytes(src, name, start=0):
    """Bytes of `pub const NAME: [u8; N] = [ ... ];` first occurrence at or after `start`."""
    m = re.compile(r"pub const %s: \[u8; \d+\] = \[(.*?)\];" % name, re.S).search(src, start)
    return bytes_of(m.group(1))


out = {"_generated_by": "tools/extract_reference_vectors.py", "_source": "vincenthz/eccoxide (reference tree)"}

# ---- NIST KATs -----------------------------------------------------------------------------
kd = read("src/tests/kats_data.rs")
kvs = []
for m in re.finditer(r"KV \{(.*?)\n    \},", kd, re.S):
    body = m.group(1)
    ent = {}
    for fld in ("k", "x", "y"):
        fm = re.search(r"\b%s: &\[(.*?)\]" % fld, body, re.S)
        ent[fld] = bytes_of(fm.group(1)).hex()
    kvs.append(ent)
assert len(kvs) == 780, len(kvs)
out["nist_p256"] = kvs[104:156]
out["nist_p384"] = kvs[156:208]

# ---- X25519 --------------------------------------------------------------------------------
xs = read("src/protocol/x25519.rs")
hexes = re.findall(r'h\(\s*"([0-9a-f]+)"', xs)
# order in the file: v1 k,u,r ; v2 k,u,r ; iterated k,r ; DH a, b, shared
out["x25519"] = {
    "rfc7748_5_2": [{"k": hexes[0], "u": hexes[1], "r": hexes[2]}, {"k": hexes[3], "u": hexes[4], "r": hexes[5]}],
    "iterated_once": {"k": hexes[6], "r": hexes[7]},
    "dh_6_1": {"a": hexes[8], "b": hexes[9], "shared": hexes[10]},
}
cs = read("src/curve/curve25519.rs")
lad = {}
for k in (2, 5, 7):
    m = re.search(r"const LADDER_9_%d: \[u8; 32\] = \[(.*?)\];" % k, cs, re.S)
    lad[str(k)] = bytes_of(m.group(1)).hex()
out["x25519"]["ladder_u9"] = lad

# ---- Ed25519 -------------------------------------------------------------------------------
es = read("src/protocol/ed25519.rs")
vecs = []
for m in re.finditer(r'Vector \{\s*seed: "([0-9a-f]*)",\s*public: "([0-9a-f]*)",\s*message: "([0-9a-f]*)",\s*signature: "([0-9a-f]*)",', es):
    vecs.append({"seed": m.group(1), "public": m.group(2), "message": m.group(3), "signature": m.group(4)})
assert len(vecs) == 3
out["ed25519_rfc8032"] = vecs

# edge scalars (u64 values) of wnaf_test_scalars and completeness::ct_matches_vartime
m = re.search(r"fn wnaf_test_scalars\(\).*?\[(.*?)\]\s*\.iter\(\)", cs, re.S)
out["ed25519_edge_scalars_u64"] = [int(x.replace("_", "").replace("u64", ""), 0) for x in re.findall(r"(0x[0-9a-f_]+|\d+)(?:u64)?,", m.group(1))]
comp = read("src/tests/completeness.rs")
m = re.search(r"fn ct_matches_vartime\(\).*?for v in \[(.*?)\] \{", comp, re.S)
out["weierstrass_edge_scalars_u64"] = [int(x.replace("_", "").replace("u64", ""), 0) for x in re.findall(r"(0x[0-9a-f_]+|\d+)(?:u64)?,", m.group(1))]
out["fullwidth_seed_u64"] = 0xfedcba9876543210  # squared 5 times in the scalar field (curve25519.rs:1419, completeness.rs:88)

# ---- ECDSA RFC 6979 ------------------------------------------------------------------------
ec = read("src/protocol/ecdsa.rs")
ecd = {}
for curve in ("p256r1", "p384r1"):
    m = re.search(r"ecdsa_test!\(\s*%s,(.*?)\n    \);" % curve, ec, re.S)
    body = m.group(1)
    strs = re.findall(r'"([0-9A-Fa-f]+)"', body.split("&[")[0])
    d, qx, qy = strs[0], strs[1], strs[2]
    kats = []
    for km in re.finditer(r'Kat \{\s*alg: Alg::(\w+),\s*message: "(\w*)",\s*k: "([0-9A-F]+)",\s*r: "([0-9A-F]+)",\s*s: "([0-9A-F]+)",', body):
        kats.append({"alg": km.group(1).lower(), "message": km.group(2), "k": km.group(3).lower(), "r": km.group(4).lower(), "s": km.group(5).lower()})
    assert kats
    ecd[curve] = {"d": d.lower(), "qx": qx.lower(), "qy": qy.lower(), "kats": kats}
out["ecdsa_rfc6979"] = ecd

# ---- BLS12-381 G1 --------------------------------------------------------------------------
g1 = read("src/curve/bls12_381/g1.rs")
i0 = g1.index("fn serialization_kat()")
seg = g1[i0:g1.index("for (k, expected) in COMPRESSED", i0)]
comp_seg, uncomp_seg = seg.split("const UNCOMPRESSED")


def kv_pairs(s):
    res = []
    for m in re.finditer(r"\(\s*(0x[0-9a-f_]+|\d+),\s*\[(.*?)\],\s*\)", s, re.S):
        res.append({"k": int(m.group(1).replace("_", ""), 0), "bytes": bytes_of(m.group(2)).hex()})
    return res


out["bls12_381_g1"] = {"compressed": kv_pairs(comp_seg), "uncompressed": kv_pairs(uncomp_seg)}
assert len(out["bls12_381_g1"]["compressed"]) == 5 and len(out["bls12_381_g1"]["uncompressed"]) == 2
# points of the curve outside the prime-order subgroup: (compressed, uncompressed) pairs (g1.rs:313-368 OFF_SUBGROUP)
i1 = g1.index("const OFF_SUBGROUP")
i1 = g1.index("= &[", i1) + 4
off_seg = g1[i1:g1.index("\n        ];", i1)]
offs = []
for m in re.finditer(r"\(\s*\[(.*?)\],\s*\[(.*?)\],\s*\)", off_seg, re.S):
    c, u = bytes_of(m.group(1)), bytes_of(m.group(2))
    assert len(c) == 48 and len(u) == 96
    offs.append({"compressed": c.hex(), "uncompressed": u.hex()})
assert len(offs) >= 2
out["bls12_381_g1"]["off_subgroup"] = offs

# ---- X448 ----------------------------------------------------------------------------------
x4 = read("src/protocol/x448.rs")
hx = re.findall(r'h\(\s*"([0-9a-f]+)"', x4)
out["x448"] = {
    "rfc7748_5_2": [{"k": hx[0], "u": hx[1], "r": hx[2]}, {"k": hx[3], "u": hx[4], "r": hx[5]}],
    "dh_6_2": {"a": hx[6], "b": hx[7], "a_pub": hx[8], "b_pub": hx[9], "shared": hx[10]},
}

# ---- comb tables ---------------------------------------------------------------------------
combs = {}
for name, fs in (("curve25519", 32), ("p256r1", 32), ("p384r1", 48), ("bls12_381", 48), ("p256k1", 32)):
    src = read("src/params/comb/%s.rs" % name)
    i0 = src.index("pub static COMB_TABLE")
    i1 = src.find("pub const WNAF_BASE_W", i0)
    tab = src[i0:i1 if i1 > 0 else len(src)]
    tab = tab[tab.index("= [") :]
    raw = bytes_of(tab)
    nwin = int(re.search(r"pub const COMB_WINDOWS: usize = (\d+);", src).group(1))
    assert len(raw) == nwin * 15 * 2 * fs, (name, len(raw))
    ent = lambda i, j: raw[(i * 15 + j) * 2 * fs : (i * 15 + j + 1) * 2 * fs].hex()
    samples = {"%d,%d" % (i, j): ent(i, j) for (i, j) in ((0, 0), (0, 14), (1, 0), (5, 6), (nwin - 1, 0), (nwin - 1, 14))}
    combs[name] = {"windows": nwin, "field_bytes": fs, "sha256": hashlib.sha256(raw).hexdigest(), "samples": samples}
out["comb_tables"] = combs

# ---- p256k1 (secp256k1): k * G for k = 1..100 (src/tests/sage.rs:9-1320, Sage-generated) ----------------
sage = read("src/tests/sage.rs")
seg = sage[sage.index("const KATS: [KAT; 100]"):sage.index("use crate::curve::sec2::p256k1")]
k1 = []
for m in re.finditer(r"KAT \{\s*n: (\d+),\s*x: \[(.*?)\],\s*y: \[(.*?)\],\s*\}", seg, re.S):
    k1.append({"k": int(m.group(1)), "x": bytes_of(m.group(2)).hex(), "y": bytes_of(m.group(3)).hex()})
assert len(k1) == 100 and k1[0]["k"] == 1 and k1[99]["k"] == 100
out["p256k1_sage"] = k1

# ---- ristretto255: RFC 9496 vectors held by src/curve/curve25519/ristretto255.rs:341-398 -----------------
rs = read("src/curve/curve25519/ristretto255.rs")
def _strs(name):
    seg = rs[rs.index("const %s:" % name):]
    seg = seg[:seg.index("];")]
    return re.findall(r'"([0-9a-f]+)"', seg)
uni = _strs("UNIFORM")
out["ristretto255"] = {"multiples": _strs("MULTIPLES"), "bad": _strs("BAD"),
                       "uniform": [{"input": uni[2 * i], "encoding": uni[2 * i + 1]} for i in range(len(uni) // 2)]}
assert len(out["ristretto255"]["multiples"]) == 16 and len(out["ristretto255"]["bad"]) == 17

# ---- domain parameters -----------------------------------------------------------------------
sec2 = read("src/params/sec2.rs")
params = {}
for curve in ("p256r1", "p384r1", "p256k1"):
    at = sec2.index("pub mod %s {" % curve)
    params[curve] = {n.lower().replace("_bytes", ""): const_bytes(sec2, n, at).hex() for n in ("P_BYTES", "ORDER_BYTES", "B_BYTES", "GX_BYTES", "GY_BYTES")}
bl = read("src/params/bls12_381.rs")
params["bls12_381_g1"] = {n.lower().replace("_bytes", ""): const_bytes(bl, n).hex() for n in ("P_BYTES", "ORDER_BYTES", "B_BYTES", "GX_BYTES", "GY_BYTES", "BETA_BYTES")}
out["params"] = params

os.makedirs(os.path.dirname(OUT), exist_ok=True)
with open(OUT, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
