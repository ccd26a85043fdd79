#!/usr/bin/env python3
"""First-contact GPU script: IMAD probe, small-batch parity against oracle/pyref.py, rough timings.
Writes gpurun_out/first_run.json.  (Development aid; the real checks live in tests/ and bench.py.)"""
import json, os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eccoxide_b200 import Context
from oracle import pyref as R

res = {}
os.makedirs("gpurun_out", exist_ok=True)
def dump():
    json.dump(res, open("gpurun_out/first_run.json", "w"), indent=1)

ctx = Context()
for v, name in ((0, "imad_lo"), (1, "imad_wide"), (2, "imad_wide_x_chain"), (3, "imad_hi")):
    best = 0
    for _ in range(3):
        macs, ms = ctx.imad_probe(v, 8192)
        best = max(best, macs)
    res["probe_" + name] = {"macs_per_s": best, "ms": ms}
    print("probe", name, "%.3f T mac/s" % (best / 1e12), "%.2f ms" % ms, flush=True)
dump()

random.seed(11)
rng = np.random.default_rng(5)
L = R.L25519
def rand_scalars_l(n):
    return np.frombuffer(b"".join((random.getrandbits(512) % L).to_bytes(32, "little") for _ in range(n)), dtype=np.uint8).reshape(n, 32).copy()

def check(name, fn):
    t0 = time.time()
    try:
        ok = fn()
        res["parity_" + name] = bool(ok)
        print("parity", name, "OK" if ok else "MISMATCH", "%.1fs" % (time.time() - t0), flush=True)
    except Exception as e:
        res["parity_" + name] = "error: %r" % (e,)
        print("parity", name, "ERROR", repr(e), flush=True)
    dump()

def par_ed_base():
    for w in (4, 8, 10):
        ctx.set_option("ed25519_comb_w", w)
        ks = [0, 1, 2, L - 1, L - 2, 2**252, 2**252 - 1] + [random.randrange(L) for _ in range(121)]
        kb = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in ks), dtype=np.uint8).reshape(-1, 32)
        out = ctx.ed25519_mul_base(kb)
        for i, k in enumerate(ks):
            if out[i].tobytes() != R.ed25519_mul_base_xy(k.to_bytes(32, "little")):
                print("mismatch w", w, "i", i); return False
        outc = ctx.ed25519_mul_base(kb, compressed=True)
        for i, k in enumerate(ks):
            if outc[i].tobytes() != R.ed_encode(R.ed_mul(k, R.ED_B)): return False
    ctx.set_option("ed25519_comb_w", 8)
    return True
check("ed25519_mul_base", par_ed_base)

def par_x25519():
    n = 200
    k = rng.integers(0, 256, (n, 32), dtype=np.uint8); u = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for i, v in enumerate([0, 1, R.P25519 - 1, R.P25519, R.P25519 + 1, 2**255 - 1]):
        u[i] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    out = ctx.x25519(k, u)
    return all(out[i].tobytes() == R.x25519(k[i].tobytes(), u[i].tobytes()) for i in range(n))
check("x25519", par_x25519)

def par_ed_mul():
    n = 64
    ks = [random.randrange(L) for _ in range(n)]; ks[0] = 0; ks[1] = 1; ks[2] = L - 1
    pts = [R.ed_mul(random.randrange(1, L), R.ED_B) for _ in range(n)]
    kb = np.frombuffer(b"".join(k.to_bytes(32, "little") for k in ks), dtype=np.uint8).reshape(-1, 32)
    pb = np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts), dtype=np.uint8).reshape(-1, 64)
    out = ctx.ed25519_mul(kb, pb)
    return all(out[i].tobytes() == R.ed25519_mul_xy(kb[i].tobytes(), pb[i].tobytes()) for i in range(n))
check("ed25519_mul", par_ed_mul)

def par_wei(curve):
    c = R.WCURVES[curve]; n = 48
    ks = [random.randrange(c.n) for _ in range(n)]; ks[0] = 0; ks[1] = 1; ks[2] = c.n - 1; ks[3] = 2
    pts = [c.mul(random.randrange(1, c.n), c.G) for _ in range(n)]
    kb = np.frombuffer(b"".join(k.to_bytes(c.sbytes, "big") for k in ks), dtype=np.uint8).reshape(n, -1)
    pb = np.frombuffer(b"".join(c.enc(p) for p in pts), dtype=np.uint8).reshape(n, -1)
    out, inf = ctx.wei_mul(curve, kb, pb)
    for i in range(n):
        e, ei = R.wei_mul(c, kb[i].tobytes(), pb[i].tobytes())
        if bool(inf[i]) != bool(ei) or out[i].tobytes() != e: print("mismatch", curve, i); return False
    outb, infb = ctx.wei_mul_base(curve, kb)
    for i in range(n):
        e, ei = R.wei_mul_base(c, kb[i].tobytes())
        if bool(infb[i]) != bool(ei) or outb[i].tobytes() != e: print("mismatch base", curve, i); return False
    return True
for cv in ("p256r1", "p384r1", "bls12_381_g1"):
    check("wei_mul_" + cv, lambda cv=cv: par_wei(cv))

def par_x448():
    n = 40
    k = rng.integers(0, 256, (n, 56), dtype=np.uint8); u = rng.integers(0, 256, (n, 56), dtype=np.uint8)
    out = ctx.x448(k, u)
    return all(out[i].tobytes() == R.x448(k[i].tobytes(), u[i].tobytes()) for i in range(n))
check("x448", par_x448)

def par_ecdsa(curve):
    c = R.WCURVES[curve]; n = 40
    Q = []; Z = []; RS = []; exp = []
    for i in range(n):
        d = random.randrange(1, c.n); k = random.randrange(1, c.n); z = random.getrandbits(8 * c.sbytes)
        rs = bytearray(R.ecdsa_sign_hashed(c, d, k, z % c.n)); zb = bytearray(z.to_bytes(c.sbytes, "big"))
        if i % 4 == 1: rs[5] ^= 1
        if i % 4 == 2: zb[3] ^= 1
        if i == 7: rs[:c.sbytes] = bytes(c.sbytes)
        q = c.enc(c.mul(d, c.G))
        Q.append(q); Z.append(bytes(zb)); RS.append(bytes(rs)); exp.append(R.ecdsa_verify_hashed(c, q, bytes(zb), bytes(rs)))
    f = lambda l: np.frombuffer(b"".join(l), dtype=np.uint8).reshape(n, -1)
    ok = ctx.ecdsa_verify_hashed(curve, f(Q), f(Z), f(RS))
    return [bool(x) for x in ok] == exp
for cv in ("p256r1", "p384r1"):
    check("ecdsa_" + cv, lambda cv=cv: par_ecdsa(cv))

def par_edverify():
    n = 40; A = []; Rr = []; S = []; K = []; exp = []
    for i in range(n):
        seed = os.urandom(32); msg = os.urandom(i)
        pub = R.ed25519_public_from_seed(seed); sig = bytearray(R.ed25519_sign(seed, msg))
        if i % 4 == 1: sig[3] ^= 1
        if i % 4 == 2: sig[40] ^= 4
        k = R.ed25519_hash_k(bytes(sig[:32]), pub, msg)
        A.append(pub); Rr.append(bytes(sig[:32])); S.append(bytes(sig[32:])); K.append(k)
        exp.append(R.ed25519_verify_prehashed(pub, bytes(sig[:32]), bytes(sig[32:]), k))
    f = lambda l: np.frombuffer(b"".join(l), dtype=np.uint8).reshape(n, 32)
    ok = ctx.ed25519_verify_prehashed(f(A), f(Rr), f(S), f(K))
    return [bool(x) for x in ok] == exp
check("ed25519_verify", par_edverify)

# ---- rough end-to-end timings (host buffers, pageable) -------------------------------------------
def timeit(name, fn, n, reps=3):
    fn()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    res["time_" + name] = {"n": n, "s": best, "ops_per_s": n / best}
    print("time", name, "n=%d %.2f ms  %.2f M/s" % (n, best * 1e3, n / best / 1e6), flush=True)
    dump()

n = 1 << 20
kb = rand_scalars_l(1 << 14); kb = np.tile(kb, (n >> 14, 1))
kb[:, 0] ^= np.arange(n, dtype=np.uint32).astype(np.uint8)  # cheap diversity, still < l since top byte untouched
for w in (4, 6, 8, 10, 12, 14, 16):
    try:
        ctx.set_option("ed25519_comb_w", w)
        t0 = time.perf_counter(); ctx.ed25519_mul_base(kb[:1024]); tb = time.perf_counter() - t0
        res["ed_table_build_w%d_s" % w] = tb
        timeit("ed25519_mul_base_w%d" % w, lambda: ctx.ed25519_mul_base(kb), n)
    except Exception as e:
        print("w", w, "failed", repr(e)); res["time_ed25519_mul_base_w%d" % w] = repr(e); dump()
ctx.set_option("ed25519_comb_w", 8)
timeit("ed25519_mul_base_2p16", lambda: ctx.ed25519_mul_base(kb[: 1 << 16]), 1 << 16)
k32 = rng.integers(0, 256, (n, 32), dtype=np.uint8); u32 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
timeit("x25519", lambda: ctx.x25519(k32, u32), n)
nn = 1 << 18
for cv in ("p256r1", "p384r1", "bls12_381_g1"):
    c = R.WCURVES[cv]
    ks = np.frombuffer(b"".join((random.getrandbits(640) % c.n).to_bytes(c.sbytes, "big") for _ in range(1024)), dtype=np.uint8).reshape(1024, -1)
    ks = np.tile(ks, (nn // 1024, 1))
    pts = [c.enc(c.mul(random.randrange(1, c.n), c.G)) for _ in range(64)]
    pb = np.tile(np.frombuffer(b"".join(pts), dtype=np.uint8).reshape(64, -1), (nn // 64, 1))
    timeit("wei_mul_" + cv, lambda: ctx.wei_mul(cv, ks, pb), nn, reps=2)
pts = [R.ed_mul(random.randrange(1, L), R.ED_B) for _ in range(64)]
pb = np.tile(np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts), dtype=np.uint8).reshape(64, 64), (nn // 64, 1))
timeit("ed25519_mul", lambda: ctx.ed25519_mul(kb[:nn], pb), nn, reps=2)
k56 = rng.integers(0, 256, (nn, 56), dtype=np.uint8); u56 = rng.integers(0, 256, (nn, 56), dtype=np.uint8)
timeit("x448", lambda: ctx.x448(k56, u56), nn, reps=2)
res["launches"] = ctx.launch_count()
dump()
print("done")
